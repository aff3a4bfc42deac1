// rvs_treeg.cuh -- K2 for wave 1 (MCTS(batch_size=1)) with SEVERAL games per warp.
//
// Profiling the warp-per-game kernel (profiles/ncu_selfplay_r1.txt) showed ~95 % of its warp
// instructions in the rollout, whose direction-sliced board ops use only 8 distinct lanes
// (lane & 7 = direction): the other 24 lanes repeat the same work.  Here a game is owned by a
// GROUP of LPG lanes (8, 4 or 2); every lane evaluates 8/LPG directions of the flip / move
// generation scans (independent dependency chains = ILP inside the thread) and one warp
// instruction advances 32/LPG independent games.  With LPG = 4 a lane owns the direction pair
// (+s, -s): the same shift amount, once on the normal and once on the bit-reversed board.
//
// The warp stays CONVERGED: every loop runs while ANY of its groups still has work and a group
// that is done is predicated off.  That is what lets the group reductions be plain full-mask
// SHFL.BFLY butterflies.  (Per-group member masks do not work: REDUX writes one uniform register
// per warp, so nvcc serialises a masked __reduce_*_sync over the distinct masks with MATCH.ANY
// loops -- measured 1.8x SLOWER than one warp per game.)
//
// Semantics are those of rvs_tree.cuh (same citations: mcts.py:84-114 score, :409-444 traverse,
// :544-623 process, :625-640 backup); only the work distribution differs:
//   * children are scanned LPG per step (first chunk / lowest lane keeps ties = first-max rule),
//   * the path lives in shared memory (64 ints per game) instead of one node per lane,
//   * expansion: lane l creates the children whose squares lie in its 8/LPG board rows -- LAZILY: an expanded node
//     keeps only its legal mask until a traverse first descends through it (materialize_g),
//   * the position of a node is stored when the node is first reached (cx.brd), so a traverse applies one move per
//     simulation instead of replaying make_move along the path (select_one_g),
//   * flips come from a shared table of rays + one carry ripple instead of a flood (flip_ray / flip_carry),
//   * the k-th legal square of a rollout ply is found by the lane whose rows hold it, through a 256 x 8
//     select-in-byte table in shared memory, and reaches the group through one warp-wide REDUX.OR (8-lane groups)
//     or a shared slot (grp_nth_post).
// None of this changes a search: visit counts, root N and root W stay bit-identical to the oracle (tests/test_gpu_mcts.py).
#pragma once
#include "rvs_tree.cuh"

namespace rvs {

template <int LPG>
struct Grp {
    static constexpr int ND = 8 / LPG;    // directions per lane
    static constexpr int RPL = 8 / LPG;   // board rows per lane
    int sh;              // first lane of the group inside the warp
    int lane;            // 0..LPG-1 inside the group
    uint64_t below;      // squares of the rows owned by lower lanes
    uint64_t belowd;     // the same set in the lane's OWN domain (LPG == 8: bit-reversed for odd lanes; else = below)
    uint64_t oned;       // square 0 in the lane's own domain
    int bsh;             // LPG == 8: index (0..7) of this lane's row byte inside an own-domain mask
    int flip63;          // LPG == 8: 63 for lanes that work on bit-reversed boards (square s = bit 63 - s), else 0
    uint32_t lut;        // shared-window address: lut[byte * 8 + j] = position of the j-th set bit of byte
    uint32_t lutd;       // the table for own-domain row bytes: odd lanes of 8-lane groups see their row bit-reversed
                         // (lut + 2048: lutr[byte * 8 + j] = lut[rev8(byte) * 8 + j])
    uint32_t gather;     // shared-window address: 3 x 64-byte exchange buffers of the group (grp_or64_own8)
    uint32_t rays;       // shared-window address of this lane's first column of the flip-ray table (ray_init)
    uint32_t path;       // shared-window address: int[kMaxPath + 1] nodes of the current path; the extra word
                         // is the broadcast slot of the 4- and 2-lane groups (grp_nth_post)
                         // (32-bit shared addresses: a generic pointer costs an S2R + LEA per access)
    DirLane d[ND];       // this lane's directions; d[j].neg == (j & 1) when ND >= 2
};

// lut[0 .. 2048): select-in-byte; lut[2048 .. 4096): the same for a bit-reversed byte (positions still count from
// the least significant bit of the NORMAL byte)
template <int LPG>
struct LutCfg { static constexpr int kBytes = (LPG == 8 ? 2 : 1) * 256 * 8; };  // the reversed-byte table only serves 8-lane groups
template <int LPG>
__device__ __forceinline__ void lut_init(uint8_t* lut, int tid, int nthreads) {
    for (int e = tid; e < LutCfg<LPG>::kBytes; e += nthreads) {
        unsigned b = (unsigned)(e >> 3) & 0xFFu;
        if (e >= 256 * 8) b = __brev(b) >> 24;
        int j = e & 7, pos = 0;
        for (int i = 0; i < 8; ++i)
            if ((b >> i) & 1u) {
                if (j == 0) { pos = i; break; }
                --j;
            }
        lut[e] = (uint8_t)pos;
    }
}

// Flip-ray table (flip_ray / flip_carry in rvs_board.cuh): row sqd (the move square in the working domain of the
// direction) holds the ray of every direction.
//   LPG == 8: 64 rows x 128 bytes, the eight direction lanes 16 bytes apart, each entry twice: groups 0 / 2 of a warp
//             read the copy at +0, groups 1 / 3 the copy at +8, so the 16 lanes the LSU serves together (two groups,
//             LDS.64) touch 32 distinct banks whatever their squares are;
//   LPG < 8:  64 rows x 64 bytes [sqd][direction]; a lane reads its 8 / LPG directions (even ones at row idx, odd
//             ones -- right shifts, bit-reversed boards -- at row 63 - idx).
template <int LPG>
struct RayCfg {
    static constexpr int kWords64 = LPG == 8 ? 64 * 16 : 64 * 8;
    static constexpr unsigned kRowBytes = LPG == 8 ? 128u : 64u;
};
template <int RULES, int LPG>
__device__ __forceinline__ void ray_init(uint64_t* rays, int tid, int nthreads) {
    for (int e = tid; e < 64 * 8; e += nthreads) {
        const int sqd = e >> 3, dir = e & 7;
        const uint64_t R = flip_ray(make_dir<RULES>(dir), sqd);
        if constexpr (LPG == 8) {
            rays[sqd * 16 + dir * 2] = R;
            rays[sqd * 16 + dir * 2 + 1] = R;
        } else {
            rays[e] = R;
        }
    }
}

// word offset of the 3 x 16-word exchange buffers of 8-lane group gi inside the CTA's gather array (kGatherWords words)
constexpr int kGatherWords = 264;
__device__ __forceinline__ int gather_off(int gi) { return gi == 0 ? 0 : (gi == 1 ? 144 : (gi == 2 ? 68 : 212)); }

template <int RULES, int LPG>
__device__ __forceinline__ Grp<LPG> make_grp(int lane32, const uint8_t* lut, int* path, const void* gather = nullptr,
                                             const uint64_t* rays = nullptr) {
    Grp<LPG> g;
    g.sh = lane32 & ~(LPG - 1);
    g.lane = lane32 & (LPG - 1);
    g.below = Grp<LPG>::RPL * g.lane == 0 ? 0ULL : ((1ULL << (8 * Grp<LPG>::RPL * g.lane)) - 1ULL);
    const bool own_rev = LPG == 8 && (g.lane & 1);  // make_dir: odd directions are right shifts = bit-reversed boards
    g.belowd = own_rev ? brev64(g.below) : g.below;
    g.oned = own_rev ? (1ULL << 63) : 1ULL;
    g.bsh = own_rev ? 7 - g.lane : g.lane;
    g.flip63 = own_rev ? 63 : 0;
    g.lut = (uint32_t)__cvta_generic_to_shared(lut);
    g.lutd = g.lut + (own_rev ? 256u * 8u : 0u);
    g.path = (uint32_t)__cvta_generic_to_shared(path);
    g.rays = !rays ? 0u
             : (uint32_t)__cvta_generic_to_shared(rays) +
                   (LPG == 8 ? 16u * (uint32_t)g.lane + 8u * (uint32_t)((lane32 >> 3) & 1) : 8u * (uint32_t)(Grp<LPG>::ND * g.lane));
    // 8-lane groups: the four groups of a warp access their exchange buffers with the same instructions, which the LSU
    // serves half a warp (two groups) at a time.  The buffers of groups 0 / 1 / 2 / 3 start at banks 0 / 16 / 4 / 20
    // (gather_off): the two 64-byte stores of a half warp (STS.64, 16 banks each) and the four distinct 16-byte chunks
    // of a half warp's LDS.128 (2 groups x 2 domains) then never share a bank.  (With starts 0 / 4 / 16 / 20 the loads
    // were conflict-free but every STS.64 took 4 wavefronts instead of 2: ncu, round 2.)
    g.gather = gather ? (uint32_t)__cvta_generic_to_shared(gather) + 4u * (uint32_t)gather_off(lane32 >> 3) : 0u;
#pragma unroll
    for (int j = 0; j < Grp<LPG>::ND; ++j) g.d[j] = make_dir<RULES>(g.lane * Grp<LPG>::ND + j);
    return g;
}

__device__ __forceinline__ unsigned lds_u8(uint32_t a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// the load only for lanes with p (the others get 0): 32 scattered byte loads are a ~5-way bank conflict, the one lane
// per group that holds the square is not
__device__ __forceinline__ unsigned lds_u8_if(uint32_t a, bool p) {
    unsigned v = 0;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.shared.u8 %0, [%1]; }" : "+r"(v) : "r"(a), "r"((unsigned)p));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- butterflies over the LPG lanes of a group; the whole warp must be converged ---------------
template <int LPG>
__device__ __forceinline__ uint64_t grp_or64(uint64_t x) {
    unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
#pragma unroll
    for (int o = 1; o < LPG; o <<= 1) {
        lo |= __shfl_xor_sync(kFull, lo, o);
        hi |= __shfl_xor_sync(kFull, hi, o);
    }
    return ((uint64_t)hi << 32) | lo;
}
// LPG == 8 only (one direction per lane, lanes 2a / 2a+1 own +s / -s): OR over the group of partial
// masks that every lane holds in ITS OWN domain (normal for even lanes, bit-reversed for odd ones);
// the result is again in the lane's own domain.  Partners of the first step are in opposite domains
// (one unconditional brev), partners of the later steps in the same one.
__device__ __forceinline__ uint64_t grp_or64_own8_shfl(uint64_t x) {
    unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
    const unsigned plo = __shfl_xor_sync(kFull, lo, 1), phi = __shfl_xor_sync(kFull, hi, 1);
    lo |= __brev(phi);  // brev64(partner): halves swap
    hi |= __brev(plo);
#pragma unroll
    for (int o = 2; o < 8; o <<= 1) {
        lo |= __shfl_xor_sync(kFull, lo, o);
        hi |= __shfl_xor_sync(kFull, hi, o);
    }
    return ((uint64_t)hi << 32) | lo;
}
// The same reduction as an all-gather through shared memory: three dependent SHFL steps (~3 x 27 cycles on
// the rollout's critical path) become one STS -> LDS round trip plus two levels of 3-input ORs.  Lanes of
// equal parity (= equal domain) write next to each other, so a lane reads its own domain's four partial
// masks and the other domain's four with two 16-byte loads each.  BUF selects one of three buffers so that
// back-to-back reductions never reuse the buffer other lanes may still be reading.
template <int BUF>
__device__ __forceinline__ uint64_t grp_or64_own8(const Grp<8>& g, uint64_t x) {
#if defined(RVS_K1_SHFL)  // A/B switch: measured 2.39e8 (butterflies) vs 2.50e8 (all-gather) sims/s at 4096 games
    return grp_or64_own8_shfl(x);
#else
    const uint32_t base = g.gather + BUF * 64;
    const uint32_t par = (uint32_t)(g.lane & 1);
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(base + par * 32 + (uint32_t)(g.lane >> 1) * 8), "r"((unsigned)x),
                 "r"((unsigned)(x >> 32))
                 : "memory");
    __syncwarp();
    unsigned s[8], t[8];
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]) : "r"(base + par * 32) : "memory");
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]) : "r"(base + par * 32 + 16) : "memory");
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]) : "r"(base + (par ^ 1u) * 32) : "memory");
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "r"(base + (par ^ 1u) * 32 + 16) : "memory");
    const unsigned lo = (s[0] | s[2]) | (s[4] | s[6]) | __brev((t[1] | t[3]) | (t[5] | t[7]));
    const unsigned hi = (s[1] | s[3]) | (s[5] | s[7]) | __brev((t[0] | t[2]) | (t[4] | t[6]));
    return ((uint64_t)hi << 32) | lo;
#endif
}
template <int LPG>
__device__ __forceinline__ unsigned grp_max(unsigned x) {
#pragma unroll
    for (int o = 1; o < LPG; o <<= 1) {
        const unsigned y = __shfl_xor_sync(kFull, x, o);
        x = x > y ? x : y;
    }
    return x;
}
template <int LPG>
__device__ __forceinline__ int4 grp_shfl4(const int4& v, int src) {
    return make_int4(__shfl_sync(kFull, v.x, src, LPG), __shfl_sync(kFull, v.y, src, LPG), __shfl_sync(kFull, v.z, src, LPG),
                     __shfl_sync(kFull, v.w, src, LPG));
}

// the rows of a group-uniform mask that this lane owns, as the low 8*RPL bits of a word (RPL <= 4)
template <int LPG>
__device__ __forceinline__ unsigned grp_slice(const Grp<LPG>& g, uint64_t m) {
    constexpr int RPL = Grp<LPG>::RPL;
    if constexpr (RPL == 4) {
        return g.lane == 0 ? (unsigned)m : (unsigned)(m >> 32);
    } else {
        constexpr int per_word = 4 / RPL;  // slices per 32-bit word
        const unsigned w = g.lane < per_word ? (unsigned)m : (unsigned)(m >> 32);
        return (w >> ((g.lane & (per_word - 1)) * 8 * RPL)) & ((1u << (8 * RPL)) - 1u);
    }
}

// k-th (0-based) set bit (ascending squares) of a group-uniform mask with more than k bits, in two halves so that
// the rollout can run the first half (pure register work + one table lookup) in the shadow of the previous ply's
// vote.  `md` is the mask in the lane's OWN domain (LPG == 8: bit-reversed for odd lanes; otherwise normal).
struct NthPrep {
    bool hit;  // this lane's rows hold the bit
    int pos;   // its square (normal numbering), valid when hit
};
template <int LPG>
__device__ __forceinline__ NthPrep grp_nth_prep(const Grp<LPG>& g, uint64_t md, int k) {
    constexpr int RPL = Grp<LPG>::RPL;
    NthPrep r;
    if constexpr (LPG == 8) {
        const unsigned byte = __byte_perm((unsigned)md, (unsigned)(md >> 32), (unsigned)g.bsh) & 0xFFu;  // one PRMT
        const int j = k - popc64(md & g.belowd);
        r.hit = (unsigned)j < (unsigned)__popc(byte);
        r.pos = g.lane * 8 + (int)lds_u8_if(g.lutd + byte * 8 + (j & 7), r.hit);
    } else {
        unsigned slice = grp_slice(g, md);
        int j = k - popc64(md & g.below);
        r.hit = (unsigned)j < (unsigned)__popc(slice);
        int row = 0;
#pragma unroll
        for (int q = 0; q + 1 < RPL; ++q) {  // walk to the row of the slice that holds bit j
            const int n = __popc(slice & 0xFFu);
            const bool next = j >= n && row == q;
            j = next ? j - n : j;
            slice = next ? slice >> 8 : slice;
            row = next ? q + 1 : row;
        }
        r.pos = (g.lane * RPL + row) * 8 + (int)lds_u8_if(g.lut + (slice & 0xFFu) * 8 + (j & 7), r.hit);
    }
    return r;
}
// exactly one lane of the group holds the bit.  8-lane groups learn the square through ONE warp-wide OR reduction
// (REDUX.OR): the squares of the four groups ride in the four bytes of the reduced word.  Same-box A/B on B200 at
// 4096 games: +1.9 % over posting the square through a shared slot (STS -> LDS round trip, two warp barriers), which
// in turn was ~45 cycles shorter than ballot + find-first-set + SHFL.  Smaller groups keep the shared slot: the eight
// groups of LPG == 4 would need two reductions, measured 2.4 % SLOWER at 16 384 games.
template <int LPG>
__device__ __forceinline__ int grp_nth_post(const Grp<LPG>& g, const NthPrep& p) {
    if constexpr (LPG == 8) {
        const unsigned all = __reduce_or_sync(kFull, p.hit ? (unsigned)p.pos << g.sh : 0u);  // g.sh = 8 x group index
        return (int)((all >> g.sh) & 0x3Fu);
    } else {
        if (p.hit) sts_s32(g.path + 4 * kMaxPath, p.pos);
        __syncwarp();
        const int sq = lds_s32(g.path + 4 * kMaxPath);
        __syncwarp();
        return sq;
    }
}

// A position held by a group: side to move / opponent in the normal [0] and the bit-reversed [1]
// domain (popcounts are domain independent).  With one direction per lane (LPG == 8) only [0] is
// used and holds the boards in the lane's OWN domain.
struct GBoard {
    uint64_t P[2], O[2];
    int side;   // 1 BLACK, 2 WHITE
    int flags;  // F_OVER | winner | F_PASSED like Board::flags
};

template <int LPG>
__device__ __forceinline__ GBoard gboard_load(const Grp<LPG>& g, const Board& b) {
    const bool blk = b.side == 1;
    const uint64_t P = blk ? b.black : b.white, O = blk ? b.white : b.black;
    if constexpr (Grp<LPG>::ND == 1) return GBoard{{to_dom(P, g.d[0].neg), 0ULL}, {to_dom(O, g.d[0].neg), 0ULL}, b.side, b.flags};
    return GBoard{{P, brev64(P)}, {O, brev64(O)}, b.side, b.flags};
}

template <int LPG>
__device__ __forceinline__ bool dir_neg(const Grp<LPG>& g, int j) {
    return Grp<LPG>::ND >= 2 ? (j & 1) != 0 : g.d[0].neg;  // compile-time unless a lane owns a single direction
}

// Board.get_valid_moves for side P against O (both domains given), group-uniform result in the lane's OWN domain
// (LPG == 8: bit-reversed for odd lanes -- popcounts, emptiness and grp_nth_prep do not need the normal form)
template <int LPG, int BUF = 1>
__device__ __forceinline__ uint64_t grp_legal_own(const Grp<LPG>& g, const uint64_t (&P)[2], const uint64_t (&O)[2]) {
    if constexpr (Grp<LPG>::ND == 1)  // boards and partial masks in the lane's own domain
        return grp_or64_own8<BUF>(g, legal_raw(g.d[0], P[0], O[0]));
    uint64_t xn = 0, xr = 0;
#pragma unroll
    for (int j = 0; j < Grp<LPG>::ND; ++j) {
        if (dir_neg(g, j)) xr |= legal_raw(g.d[j], P[1], O[1]);
        else xn |= legal_raw(g.d[j], P[0], O[0]);
    }
    return grp_or64<LPG>(xn | brev64(xr));
}
template <int LPG>
__device__ __forceinline__ uint64_t own_to_normal(const Grp<LPG>& g, uint64_t md) {
    if constexpr (Grp<LPG>::ND == 1) return to_dom(md, g.d[0].neg);
    return md;
}
// the same in the normal domain
template <int LPG, int BUF = 1>
__device__ __forceinline__ uint64_t grp_legal(const Grp<LPG>& g, const uint64_t (&P)[2], const uint64_t (&O)[2]) {
    return own_to_normal(g, grp_legal_own<LPG, BUF>(g, P, O));
}

struct MoveOut {
    uint64_t P[2], O[2];  // mover / opponent after the flips (roles not swapped yet)
    uint64_t lm_opp;      // legal mask of the opponent, in the lane's OWN domain (grp_legal_own)
};

// flips of move idx by the side to move + the opponent's reply mask (board.py:181-240)
template <int LPG>
__device__ __forceinline__ MoveOut grp_flip(const Grp<LPG>& g, const GBoard& c, int idx) {
    MoveOut m;
    if constexpr (Grp<LPG>::ND == 1) {  // everything in the lane's own domain
        const int sqd = idx ^ g.flip63;
        const uint64_t mv = 1ULL << sqd;
        uint64_t R;  // this lane's ray of the move square: one table load instead of the five-step flood
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(R) : "r"(g.rays + RayCfg<LPG>::kRowBytes * (uint32_t)sqd));
        const uint64_t f = grp_or64_own8<0>(g, flip_carry(R, c.P[0], c.O[0]));
        m.P[0] = c.P[0] ^ (mv | f); m.P[1] = 0ULL;
        m.O[0] = c.O[0] ^ f;        m.O[1] = 0ULL;
    } else {
        const uint64_t mvn = 1ULL << idx, mvr = 1ULL << (63 - idx);
        uint64_t fn = 0, fr = 0;
        const uint32_t rown = g.rays + RayCfg<LPG>::kRowBytes * (uint32_t)idx, rowr = g.rays + RayCfg<LPG>::kRowBytes * (uint32_t)(63 - idx);
#pragma unroll
        for (int j = 0; j < Grp<LPG>::ND; ++j) {
            uint64_t R;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(R) : "r"((dir_neg(g, j) ? rowr : rown) + 8u * (uint32_t)j));
            if (dir_neg(g, j)) fr |= flip_carry(R, c.P[1], c.O[1]);
            else fn |= flip_carry(R, c.P[0], c.O[0]);
        }
        const uint64_t f = grp_or64<LPG>(fn | brev64(fr));
        const uint64_t fb = brev64(f);
        m.P[0] = c.P[0] ^ (mvn | f); m.P[1] = c.P[1] ^ (mvr | fb);
        m.O[0] = c.O[0] ^ f;         m.O[1] = c.O[1] ^ fb;
    }
    m.lm_opp = grp_legal_own(g, m.O, m.P);
    return m;
}

__device__ __forceinline__ int over_flags(const GBoard& c, uint64_t P, uint64_t O) {  // board.py:246-249, 363-373
    const int np = popc64(P), no = popc64(O);
    const int nb = c.side == 1 ? np : no, nw = c.side == 1 ? no : np;
    const int w = nb > nw ? 1 : (nw > nb ? 2 : 0);
    return F_PASSED | F_OVER | (w << F_WIN_SHIFT);
}

// apply_move() by a group (board.py:181-251).  Groups with act == false keep their position, get 0.
template <int LPG>
__device__ __forceinline__ uint64_t grp_apply_move(const Grp<LPG>& g, GBoard& c, int idx, bool act) {
    const MoveOut m = grp_flip(g, c, act ? idx : 0);
    uint64_t lm = own_to_normal(g, m.lm_opp);
    const bool pass = act && lm == 0;
    if (__any_sync(kFull, pass)) {  // rare, warp-uniform branch: auto-pass (board.py:242-249)
        const uint64_t lm2 = grp_legal<LPG, 2>(g, m.P, m.O);
        if (pass) {
            c.P[0] = m.P[0]; c.P[1] = m.P[1]; c.O[0] = m.O[0]; c.O[1] = m.O[1];
            c.flags = lm2 == 0 ? over_flags(c, m.P[0], m.O[0]) : (int)F_PASSED;
            lm = lm2;
        }
    }
    if (act && !pass) {
        c.P[0] = m.O[0]; c.P[1] = m.O[1]; c.O[0] = m.P[0]; c.O[1] = m.P[1];
        c.side = 3 - c.side;
        c.flags = 0;
    }
    return act ? lm : 0ULL;
}

// random_playout() of one position per group.  A group whose game has ended keeps executing the
// loop body on a dead position (cheaper than predicating every state update); its result was
// captured when the game ended.  Returns the winner (0 draw, 1 black, 2 white); plies are counted
// into `plies`.  lm (normal domain) == 0 on entry means "no rollout for this group".
//
// The loop is software-pipelined around its dependency chain (move square -> flips -> exchange -> move generation ->
// exchange -> next square): the random draw of the NEXT ply is taken at the top of a ply, and the register half of
// the next square selection (grp_nth_prep, incl. its table lookup) is issued BEFORE the branch on the pass vote, so
// both run in the shadow of the exchanges / the vote instead of extending the chain.  The legal mask stays in the
// lane's own domain from one ply to the next (no bit reversal on the chain).  The draws consumed per ply are the
// same as in the plain loop (one roll_next per ply, in order; an auto-pass consumes none).
template <int LPG>
__device__ __forceinline__ int grp_random_playout(const Grp<LPG>& g, GBoard c, uint64_t lm_normal, uint64_t stream, int& plies) {
    uint32_t rs = roll_init(stream);
    bool done = lm_normal == 0;
    int winner = 0;
    plies = 0;
    if (__all_sync(kFull, done)) return 0;
    // a dead group plays square 0 over and over; nothing of it is read
    uint64_t lm = done ? g.oned : (Grp<LPG>::ND == 1 ? to_dom(lm_normal, g.d[0].neg) : lm_normal);
    uint32_t r = roll_next(rs);  // the draw of the first ply
    NthPrep np = grp_nth_prep(g, lm, roll_pick(r, popc64(lm)));
    while (true) {
        r = roll_next(rs);  // the draw of the next ply
        const int idx = grp_nth_post(g, np);
        const MoveOut m = grp_flip(g, c, idx);
        plies += done ? 0 : 1;
        lm = m.lm_opp;
        const bool pass = !done && lm == 0;
        const bool any_pass = __any_sync(kFull, pass);
        // default: the opponent moves next
        c.P[0] = m.O[0]; c.P[1] = m.O[1]; c.O[0] = m.P[0]; c.O[1] = m.P[1];
        c.side = 3 - c.side;
        np = grp_nth_prep(g, lm, roll_pick(r, popc64(lm)));  // speculative: right unless this group passes
        if (any_pass) {  // rare: auto-pass or game over (board.py:242-249)
            const uint64_t lm2 = grp_legal_own<LPG, 2>(g, m.P, m.O);
            if (pass) {
                c.side = 3 - c.side;  // the mover keeps the turn
                c.P[0] = m.P[0]; c.P[1] = m.P[1]; c.O[0] = m.O[0]; c.O[1] = m.O[1];
                lm = lm2;
                if (lm2 == 0) {
                    winner = (over_flags(c, m.P[0], m.O[0]) & F_WIN_MASK) >> F_WIN_SHIFT;
                    done = true;
                }
            }
            if (done) lm = g.oned;  // keep the dead group's ply well defined
            if (__all_sync(kFull, done)) break;
            np = grp_nth_prep(g, lm, roll_pick(r, popc64(lm)));  // same draw, the mask after the pass
        }
    }
    return winner;
}

// per-group view of one game's tree + running counters (TreeCtx of rvs_tree.cuh, LPG lanes wide)
template <int LPG>
struct TreeCtxG {
    int4* hot;
    int4* cold;
    int4* brd;  // position of every visited node {black, white} (select_one_g); its side to move sits in hot.z
    int cap;
    int n_nodes;
    float c_puct;
    int overflow;
    unsigned steps, sims, evals, bytes, created;
    Grp<LPG> g;
    // shared-memory staging of the HOT NODE ROWS of the search: the root and its children (every simulation scans
    // and backs up exactly these rows).  stage = generic address of this group's block of (1 + kStageRows) 32-byte
    // rows {hot, cold} in shared memory; nc0 = staged children (0: nothing staged, all rows live in HBM/L1)
    int4* stage;
    int fc0;
    unsigned nc0;
};

// rows a group can stage: the root + its children (REF rules: up to ~33 legal squares incl. phantom moves)
template <int LPG>
struct StageCfg { static constexpr int kRows = LPG == 8 ? 32 : (LPG == 4 ? 16 : 0); };  // LPG == 4: 16 one-warp CTAs per SM must fit

// address of a node's hot / cold row: the staged copy when the node is the root or one of its staged children.
// One generic-space access serves both cases, so groups of a warp that sit at different tree levels do not diverge.
template <int LPG>
__device__ __forceinline__ int4* hot_at(const TreeCtxG<LPG>& cx, int node) {
    if constexpr (StageCfg<LPG>::kRows == 0) return &cx.hot[node];
    const unsigned rel = (unsigned)(node - cx.fc0);
    const bool st = cx.nc0 != 0u && (node == 0 || rel < cx.nc0);
    return st ? cx.stage + 2 * (node == 0 ? 0 : 1 + (int)rel) : &cx.hot[node];
}
template <int LPG>
__device__ __forceinline__ int4* cold_at(const TreeCtxG<LPG>& cx, int node) {
    if constexpr (StageCfg<LPG>::kRows == 0) return &cx.cold[node];
    const unsigned rel = (unsigned)(node - cx.fc0);
    const bool st = cx.nc0 != 0u && (node == 0 || rel < cx.nc0);
    return st ? cx.stage + 2 * (node == 0 ? 0 : 1 + (int)rel) + 1 : &cx.cold[node];
}

// after the root expansion (and the root noise): copy the root row and its children into the group's block
template <int LPG>
__device__ __forceinline__ void stage_root(TreeCtxG<LPG>& cx, bool act) {
    if constexpr (StageCfg<LPG>::kRows != 0) {
        unsigned nc = 0;
        int fc = 0;
        if (act) {
            const int4 c = cx.cold[0];
            nc = (unsigned)(c.z & 0xFF);
            fc = c.y;
            if (nc > (unsigned)StageCfg<LPG>::kRows) nc = 0;  // does not fit: this search stays in HBM/L1
        }
        for (unsigned i = cx.g.lane; i < nc + (nc ? 1u : 0u); i += LPG) {
            const int node = i == 0 ? 0 : fc + (int)i - 1;
            cx.stage[2 * i] = cx.hot[node];
            cx.stage[2 * i + 1] = cx.cold[node];
        }
        __syncwarp();
        cx.fc0 = fc;
        cx.nc0 = nc;
    }
}
// end of the search: the hot rows (N, W, VL, cached score) go back to HBM, where the move choice reads them
template <int LPG>
__device__ __forceinline__ void unstage_root(TreeCtxG<LPG>& cx) {
    if constexpr (StageCfg<LPG>::kRows != 0) {
        __syncwarp();
        const unsigned nc = cx.nc0;
        for (unsigned i = cx.g.lane; i < nc + (nc ? 1u : 0u); i += LPG) {
            const int node = i == 0 ? 0 : cx.fc0 + (int)i - 1;
            cx.hot[node] = cx.stage[2 * i];
        }
        cx.nc0 = 0u;
        __syncwarp();
    }
}

// LAZY CHILD ROWS.  node.expand (mcts.py:141-161) creates one child per legal move, but with 100 simulations per move
// ~70 % of the expanded nodes are never selected again, so their children are never looked at: 4096 games x ~900 rows x
// 32 B per ply is the 126 MB L2 once over (ncu: L2 hit rate 52 %, the child-row loads of the scan wait ~280 cycles on
// average).  With the uniform prior of the built-in evaluators an unvisited child carries no information beyond its
// square, so expanding a node (other than the root, whose rows are staged) only records the legal mask in the node's
// cold row {P, mask lo, meta | kLazyKids, mask hi}; the rows are created -- in the same ascending square order, with
// the same contents -- by materialize_g the first time a traverse descends through the node.  Searches are unchanged
// node for node; only row indices differ, and those are not observable.
constexpr int kLazyKids = 1 << 20;  // cold.z: the children of this node exist only as the legal mask in cold.y / cold.w
constexpr float kUniformPrior = 1.0f / 65.0f;

// child rows of `lm` at fc, fc + 1, ... (lane l creates the children whose squares lie in its rows)
template <int LPG>
__device__ __forceinline__ void create_children_g(TreeCtxG<LPG>& cx, int fc, uint64_t lm, float prior, int turn) {
    const Grp<LPG>& g = cx.g;
    unsigned slice = grp_slice(g, lm);
    int i = fc + popc64(lm & g.below);
    while (slice) {
        const int sq = g.lane * 8 * Grp<LPG>::RPL + (__ffs(slice) - 1);
        slice &= slice - 1;
        cx.hot[i] = make_int4(0, 0, 0, 0);
        cx.cold[i] = make_int4(__float_as_int(prior), -1, (sq << 8) | (turn << 16), 0);
        ++i;
    }
}

// the traverse is about to scan the children of `node` (cold row c, group-uniform): create them if they are still lazy
template <int LPG>
__device__ __forceinline__ void materialize_g(TreeCtxG<LPG>& cx, int node, int4& c, bool& going, bool lazy) {
    if (lazy) {
        const uint64_t lm = (uint64_t)(unsigned)c.y | ((uint64_t)(unsigned)c.w << 32);
        const int nc = c.z & 0xFF;
        if (cx.n_nodes + nc > cx.cap) {
            cx.overflow |= 1;
            going = false;  // the node stays a leaf of this simulation
        } else {
            const int fc = cx.n_nodes;
            create_children_g(cx, fc, lm, kUniformPrior, 3 - ((c.z >> 16) & 3));  // mcts.py:618
            c.y = fc;
            c.w = 0;
            c.z &= ~kLazyKids;
            if (cx.g.lane == 0) *cold_at(cx, node) = c;
            cx.n_nodes += nc;
            cx.bytes += 32u * (unsigned)nc;
        }
    }
    __syncwarp();
}

// MCTS._backpropagate_path (mcts.py:625-640): lane l owns path nodes l, l+LPG, ...
template <int LPG>
__device__ __forceinline__ void backup_path_g(TreeCtxG<LPG>& cx, int plen, float v, bool act) {
    const int n_upd = act ? plen : 0;
    for (int d = cx.g.lane; d < n_upd; d += LPG) {
        const int n = lds_s32(cx.g.path + 4 * d);
        int4* hp = hot_at(cx, n);
        int4 h = *hp;
        const float sv = ((plen - 1 - d) & 1) ? -v : v;
        h.x += 1;
        h.y = __float_as_int(__fadd_rn(__int_as_float(h.y), sv));
        int vl = h.z & kVLMask;
        if (vl > 0) --vl;
        h.z = (h.z & ~(kVLMask | kCacheValid)) | vl;
        *hp = h;
    }
    cx.bytes += 32u * (unsigned)n_upd;
    __syncwarp();
}

// hot.z bits of a node whose position is stored in cx.brd (set when the node is first reached as a leaf)
constexpr int kBoardWhite = 1 << 20;  // WHITE is to move in the stored position
constexpr int kHasBoard = 1 << 21;

// MCTS._traverse (mcts.py:409-444); the path is left in cx.g.path[0..plen).
// `b` enters as the root position and leaves as the position of the leaf.  The reference replays game.make_move along
// the whole path in every simulation (mcts.py:439).  Here the position of a node is stored the first time the node
// is reached (cx.brd, 16 bytes), so the descent itself touches no board: the leaf's position is its parent's stored
// position plus ONE move (none for a terminal leaf, whose stored value is backed up) -- the same position, since
// make_move is deterministic, including its auto-passes.
template <int LPG>
__device__ __forceinline__ int select_one_g(TreeCtxG<LPG>& cx, GBoard& b, int& plen, int& leaf_vlf, uint64_t& leaf_lm, bool act) {
    const Grp<LPG>& g = cx.g;
    int node = 0;
    plen = 1;
    int4 h = make_int4(0, 0, 0, 0), c = make_int4(0, 0, 0, 0);
    if (act) {
        if (g.lane == 0) sts_s32(g.path, 0);
        h = *hot_at(cx, 0);
        c = *cold_at(cx, 0);
        cx.bytes += 32;
    }
    const unsigned key_floor = ordered_key(-INFINITY);
    bool going = act;
    int parent = 0, parent_z = 0, mv = 0;  // the node above the current one, its hot.z, and the move that leads here
    leaf_lm = 0;
    while (true) {
        const int nchild = c.z & 0xFF;
        going = going && nchild != 0 && !(h.z & kTerminal);
        if (!__any_sync(kFull, going)) break;
        {
            const bool lazy = going && (c.z & kLazyKids);
            if (__any_sync(kFull, lazy)) materialize_g(cx, node, c, going, lazy);
        }
        if (going) {
            h.z += 1;  // node.virtual_loss += 1 (mcts.py:416)
            if (g.lane == 0) reinterpret_cast<int*>(hot_at(cx, node))[2] = h.z;
            cx.bytes += 32u * (unsigned)nchild;
        }
        // a node with children has been backed up at least once (N >= 1); groups that are not descending any more
        // still execute this line, and sqrt(0) would take the IEEE square root's slow path (a call, 33 % of the
        // level steps of a warp: ncu)
        const float sq = __fsqrt_rn((float)(going ? h.x : 1));
        const int fc = c.y;
        const int nscan = going ? nchild : 0;
        unsigned best_key = key_floor;
        int best_i = -1;
        int4 bh = h, bc = c;
        for (int base = 0; __any_sync(kFull, base < nscan); base += LPG) {
            const int i = base + g.lane;
            int4 ch = make_int4(0, 0, 0, 0), cc = make_int4(0, 0, 0, 0);
            unsigned key = 0;
            // uniform control flow: the score of a visited child without a valid cache is computed by the whole warp
            // whenever any lane needs it (lanes that do not get harmless operands), instead of a divergent branch per
            // lane with its reconvergence points
            const bool in = i < nscan;
            int4* hp = hot_at(cx, fc + i);
            if (in) {
                ch = *hp;
                cc = *cold_at(cx, fc + i);
            }
            const bool need = in && ch.x != 0 && !(ch.z & kCacheValid);
            if (__any_sync(kFull, need)) {
                const float s = score_child(need ? ch.x : 1, need ? __int_as_float(ch.y) : 1.0f, need ? (ch.z & kVLMask) : 0,
                                            need ? __int_as_float(cc.x) : 1.0f, (cc.z >> 16) & 3, cx.c_puct, need ? sq : 1.0f);
                if (need) {
                    ch.w = __float_as_int(s);
                    ch.z |= kCacheValid;
                    reinterpret_cast<int2*>(hp)[1] = make_int2(ch.z, ch.w);
                }
            }
            {
                const float score = ch.x == 0 ? INFINITY : __int_as_float(ch.w);  // mcts.py:96-100
                key = in ? ((score == score) ? ordered_key(score) : 0u) : 0u;
            }
            const unsigned mx = grp_max<LPG>(key);
            const unsigned bal = __ballot_sync(kFull, key == mx) >> g.sh;  // never empty inside the group: mx is one of its keys
            const int src = (__ffs(bal) - 1) & (LPG - 1);
            const int4 nh = grp_shfl4<LPG>(ch, src), nc = grp_shfl4<LPG>(cc, src);
            if (mx > best_key) {  // strict: an earlier chunk keeps ties (mcts.py:425)
                best_key = mx;
                best_i = base + src;
                bh = nh;
                bc = nc;
            }
        }
        if (going && best_i < 0) { cx.overflow |= 2; going = false; }
        if (going) {
            if (plen < kMaxPath) {
                parent = node;
                parent_z = h.z;
                mv = (bc.z >> 8) & 0x3F;  // next_move (mcts.py:439)
                node = fc + best_i;
                h = bh;
                c = bc;
                if (g.lane == 0) sts_s32(g.path + 4 * plen, node);
                ++plen;
            } else {
                cx.overflow |= 4;
                going = false;
            }
        }
    }
    // the leaf's position: game.make_move(*next_move) on the parent's stored position (mcts.py:439)
    const bool step = act && node != 0 && !(h.z & kTerminal);
    if (__any_sync(kFull, step)) {
        if (step && parent != 0) {
            const int4 pb = cx.brd[parent];
            const uint64_t black = (uint64_t)(unsigned)pb.x | ((uint64_t)(unsigned)pb.y << 32);
            const uint64_t white = (uint64_t)(unsigned)pb.z | ((uint64_t)(unsigned)pb.w << 32);
            b = gboard_load(g, Board{black, white, (uint8_t)((parent_z & kBoardWhite) ? 2 : 1), 0});
            cx.bytes += 16;
        }
        const uint64_t lm = grp_apply_move(g, b, mv, step);
        if (step) {
            leaf_lm = lm;
            ++cx.steps;
            h.z = (h.z & ~kBoardWhite) | kHasBoard | (b.side == 2 ? kBoardWhite : 0);
            if (g.lane == 0) {  // lane 0 works in the normal domain
                const uint64_t black = b.side == 1 ? b.P[0] : b.O[0], white = b.side == 1 ? b.O[0] : b.P[0];
                cx.brd[node] = make_int4((int)(unsigned)black, (int)(unsigned)(black >> 32), (int)(unsigned)white, (int)(unsigned)(white >> 32));
                reinterpret_cast<int*>(hot_at(cx, node))[2] = h.z;
            }
            cx.bytes += 16;
        }
    }
    if (__any_sync(kFull, act && node == 0)) {  // the root itself is the leaf (first simulation of a search)
        const uint64_t lm0 = grp_legal<LPG, 2>(g, b.P, b.O);
        if (act && node == 0) leaf_lm = lm0;
    }
    leaf_vlf = h.z;
    __syncwarp();
    return node;
}

// node.expand (mcts.py:141-161, 605-618) with the uniform prior of the built-in evaluators:
// lane l creates the children whose squares lie in its rows; child index = rank of the square
template <int LPG>
__device__ __forceinline__ void expand_node_g(TreeCtxG<LPG>& cx, int node, uint64_t lm, float prior, bool act) {
    const Grp<LPG>& g = cx.g;
    if (act) {
        int4* cp = cold_at(cx, node);
        int4 c = *cp;
        const int nc = popc64(lm);
        if ((c.z & 0xFF) != 0) {
            // 'if action not in self.children' (mcts.py:154): already expanded, nothing to add
        } else if (node != 0 && prior == kUniformPrior) {
            // lazy: the legal mask stands for the children until a traverse needs their rows (materialize_g)
            if (g.lane == 0) {
                c.y = (int)(unsigned)lm;
                c.w = (int)(unsigned)(lm >> 32);
                c.z = (c.z & ~0xFF) | nc | kLazyKids;
                *cp = c;
            }
            cx.created += (unsigned)nc;
            cx.bytes += 32u;
        } else if (cx.n_nodes + nc > cx.cap) {
            cx.overflow |= 1;
        } else {
            const int fc = cx.n_nodes;
            create_children_g(cx, fc, lm, prior, 3 - ((c.z >> 16) & 3));  // mcts.py:618
            if (g.lane == 0) {
                c.y = fc;
                c.z = (c.z & ~0xFF) | nc;
                *cp = c;
            }
            cx.n_nodes += nc;
            cx.created += (unsigned)nc;
            cx.bytes += 32u * (unsigned)nc;
        }
    }
    __syncwarp();
}

// one simulation per group (select -> evaluate -> expand -> backup), same results as simulate_one()
template <int EVAL, int LPG>
__device__ __forceinline__ void simulate_one_g(TreeCtxG<LPG>& cx, const GBoard& root, uint64_t stream_for_sim, bool act) {
    GBoard b = root;
    int plen, vlf;
    uint64_t lm;  // legal mask of the leaf position
    const int node = select_one_g(cx, b, plen, vlf, lm, act);
    if (act) ++cx.sims;
    const bool term = act && (vlf & kTerminal);          // mcts.py:364-366: back its stored value up
    const bool dead = act && !term && lm == 0;           // mcts.py:567-579: flag terminal, ABSOLUTE value
    const bool eval = act && !term && lm != 0;
    float v = 0.0f;
    if (EVAL == RVS_EVAL_E0) {
        v = __fdiv_rn((float)(popc64(b.P[0]) - popc64(b.O[0])), 64.0f);
    } else {
        int plies;
        const int w = grp_random_playout(cx.g, b, eval ? lm : 0ULL, stream_for_sim, plies);
        if (eval) cx.steps += (unsigned)plies;
        v = w == 0 ? 0.0f : (w == b.side ? 1.0f : -1.0f);
    }
    if (eval) ++cx.evals;
    if (term) v = term_value_of(vlf);
    if (dead) {
        const int w = (b.flags & F_WIN_MASK) >> F_WIN_SHIFT;
        const int code = !(b.flags & F_OVER) ? 0 : (w == 1 ? 1 : (w == 2 ? 2 : 0));
        if (cx.g.lane == 0) {
            int* z = &reinterpret_cast<int*>(hot_at(cx, node))[2];
            *z = (*z & ~(3 << kTermShift)) | kTerminal | (code << kTermShift);
        }
        v = code == 1 ? 1.0f : (code == 2 ? -1.0f : 0.0f);
    }
    expand_node_g(cx, node, lm, kUniformPrior, eval);  // ends with __syncwarp(): the flag above is visible
    backup_path_g(cx, plen, v, act);
}

}  // namespace rvs

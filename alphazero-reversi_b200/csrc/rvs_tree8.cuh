// rvs_tree8.cuh -- K2 for wave 1 (MCTS(batch_size=1)) with FOUR games per warp.
//
// Profiling the warp-per-game kernel (profiles/ncu_selfplay_r1.txt) showed ~95 % of its warp
// instructions in the rollout, whose direction-sliced board ops use only 8 distinct lanes
// (lane & 7 = direction): the other 24 lanes repeat the same work.  Here a game is owned by an
// 8-lane GROUP (lane & 7 = direction, group = lane >> 3), so one warp instruction advances four
// independent games and the same 4096 games need a quarter of the issue slots.
//
// The warp stays CONVERGED: every loop runs while ANY of its four groups still has work and a
// group that is done is predicated off.  That is what lets the group reductions be plain
// full-mask SHFL.BFLY butterflies over 8 lanes.  (Per-group member masks do not work: REDUX
// writes one uniform register per warp, so nvcc serialises a masked __reduce_*_sync over the
// distinct masks with MATCH.ANY loops -- measured 1.8x SLOWER than one warp per game.)
//
// Semantics are those of rvs_tree.cuh (same citations: mcts.py:84-114 score, :409-444 traverse,
// :544-623 process, :625-640 backup); only the work distribution differs:
//   * children are scanned 8 per step (first chunk / lowest lane keeps ties = first-max rule),
//   * the path lives in shared memory (64 ints per group) instead of one node per lane,
//   * expansion: lane l creates the children of board row l (byte l of the legal mask),
//   * the k-th legal square of a rollout ply is found by the lane whose byte holds it, through a
//     256 x 8 select-in-byte table in shared memory.
#pragma once
#include "rvs_tree.cuh"

namespace rvs {

struct Grp {
    int sh;              // first lane of the group: 0, 8, 16, 24
    int lane;            // 0..7 inside the group
    uint64_t below;      // bits of rows < lane: (1 << 8*lane) - 1
    const uint8_t* lut;  // shared: lut[byte * 8 + j] = position of the j-th set bit of byte
    int* path;           // shared: [kMaxPath] nodes of the current path
};

__device__ __forceinline__ void lut_init(uint8_t* lut, int tid, int nthreads) {
    for (int e = tid; e < 256 * 8; e += nthreads) {
        unsigned b = (unsigned)(e >> 3);
        int j = e & 7, pos = 0;
        for (int i = 0; i < 8; ++i)
            if ((b >> i) & 1u) {
                if (j == 0) { pos = i; break; }
                --j;
            }
        lut[e] = (uint8_t)pos;
    }
}

__device__ __forceinline__ Grp make_grp(int lane32, const uint8_t* lut, int* path) {
    Grp g;
    g.sh = lane32 & 24;
    g.lane = lane32 & 7;
    g.below = (1ULL << (8 * g.lane)) - 1ULL;
    g.lut = lut;
    g.path = path;
    return g;
}

// ---- 8-lane butterflies; the whole warp must be converged -------------------------------------
__device__ __forceinline__ uint64_t grp_or64(uint64_t x) {
    unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        lo |= __shfl_xor_sync(kFull, lo, o);
        hi |= __shfl_xor_sync(kFull, hi, o);
    }
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ unsigned grp_max(unsigned x) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const unsigned y = __shfl_xor_sync(kFull, x, o);
        x = x > y ? x : y;
    }
    return x;
}
__device__ __forceinline__ int4 grp_shfl4(const int4& v, int src) {
    return make_int4(__shfl_sync(kFull, v.x, src, 8), __shfl_sync(kFull, v.y, src, 8), __shfl_sync(kFull, v.z, src, 8),
                     __shfl_sync(kFull, v.w, src, 8));
}

// k-th (0-based) set bit of a group-uniform mask: lane l looks at row l.  -1 when m has <= k bits.
__device__ __forceinline__ int grp_nth_set_bit(const Grp& g, uint64_t m, int k) {
    const unsigned w = g.lane < 4 ? (unsigned)m : (unsigned)(m >> 32);
    const unsigned byte = (w >> ((g.lane & 3) * 8)) & 0xFFu;
    const int j = k - popc64(m & g.below);
    const bool hit = (unsigned)j < (unsigned)__popc(byte);
    const unsigned pos = hit ? (unsigned)(g.lane * 8 + g.lut[byte * 8 + (j & 7)] + 1) : 0u;
    return (int)grp_max(pos) - 1;
}

__device__ __forceinline__ uint64_t grp_legal(const DirLane& L, const CoopBoard& c) {
    return grp_or64(legal_part(L, c.Pd, c.Od));
}

// coop_apply_move() by an 8-lane group (board.py:181-251).  Groups with act == false keep their
// position and get 0.
__device__ __forceinline__ uint64_t grp_apply_move(const DirLane& L, CoopBoard& c, int idx, bool act) {
    idx = act ? idx : 0;
    const uint64_t mvd = 1ULL << (L.neg ? 63 - idx : idx);
    const uint64_t f = grp_or64(flip_part(L, c.Pd, c.Od, mvd));
    const uint64_t fd = to_dom(f, L.neg);
    const uint64_t P = c.Pd ^ (mvd | fd), O = c.Od ^ fd;
    uint64_t lm = grp_or64(legal_part(L, O, P));  // opponent to move
    const bool pass = act && lm == 0;
    if (__any_sync(kFull, pass)) {                // rare, warp-uniform branch
        const uint64_t lm2 = grp_or64(legal_part(L, P, O));  // auto-pass (board.py:242-249)
        if (pass) {
            c.Pd = P; c.Od = O;
            c.flags = F_PASSED;
            if (lm2 == 0) {
                const int np = popc64(P), no = popc64(O);
                const int nb = c.side == 1 ? np : no, nw = c.side == 1 ? no : np;
                const int w = nb > nw ? 1 : (nw > nb ? 2 : 0);
                c.flags = F_PASSED | F_OVER | (w << F_WIN_SHIFT);
            }
            lm = lm2;
        }
    }
    if (act && !pass) { c.Pd = O; c.Od = P; c.side = 3 - c.side; c.flags = 0; }
    return act ? lm : 0ULL;
}

// random_playout() of up to four positions at once; a group with lm == 0 sits the loop out
__device__ __forceinline__ int grp_random_playout(const Grp& g, const DirLane& L, CoopBoard& c, uint64_t lm, uint64_t stream) {
    int plies = 0;
    uint32_t rs = roll_init(stream);
    while (__any_sync(kFull, lm != 0)) {
        const bool a = lm != 0;
        const int k = roll_pick(roll_next(rs), popc64(lm));
        lm = grp_apply_move(L, c, grp_nth_set_bit(g, lm, k), a);
        plies += a ? 1 : 0;
    }
    return plies;
}

// per-group view of one game's tree + running counters (TreeCtx of rvs_tree.cuh, 8 lanes wide)
struct TreeCtx8 {
    int4* hot;
    int4* cold;
    int cap;
    int n_nodes;
    float c_puct;
    int overflow;
    unsigned steps, sims, evals, bytes, created;
    DirLane dir;
    Grp g;
};

// MCTS._backpropagate_path (mcts.py:625-640): lane l owns path nodes l, l+8, ...
__device__ __forceinline__ void backup_path8(TreeCtx8& cx, int plen, float v, bool act) {
    const int n_upd = act ? plen : 0;
    for (int d = cx.g.lane; d < n_upd; d += 8) {
        const int n = cx.g.path[d];
        int4 h = cx.hot[n];
        const float sv = ((plen - 1 - d) & 1) ? -v : v;
        h.x += 1;
        h.y = __float_as_int(__fadd_rn(__int_as_float(h.y), sv));
        int vl = h.z & kVLMask;
        if (vl > 0) --vl;
        h.z = (h.z & ~(kVLMask | kCacheValid)) | vl;
        cx.hot[n] = h;
    }
    cx.bytes += 32u * (unsigned)n_upd;
    __syncwarp();
}

// MCTS._traverse (mcts.py:409-444); the path is left in cx.g.path[0..plen)
__device__ __forceinline__ int select_one8(TreeCtx8& cx, CoopBoard& b, int& plen, int& leaf_vlf, bool act) {
    const Grp& g = cx.g;
    int node = 0;
    plen = 1;
    int4 h = make_int4(0, 0, 0, 0), c = make_int4(0, 0, 0, 0);
    if (act) {
        if (g.lane == 0) g.path[0] = 0;
        h = cx.hot[0];
        c = cx.cold[0];
        cx.bytes += 32;
    }
    const unsigned key_floor = ordered_key(-INFINITY);
    bool going = act;
    while (true) {
        const int nchild = c.z & 0xFF;
        going = going && nchild != 0 && !(h.z & kTerminal);
        if (!__any_sync(kFull, going)) break;
        if (going) {
            h.z += 1;  // node.virtual_loss += 1 (mcts.py:416)
            if (g.lane == 0) reinterpret_cast<int*>(&cx.hot[node])[2] = h.z;
            cx.bytes += 32u * (unsigned)nchild;
        }
        const float sq = __fsqrt_rn((float)h.x);
        const int fc = c.y;
        const int nscan = going ? nchild : 0;
        unsigned best_key = key_floor;
        int best_i = -1;
        int4 bh = h, bc = c;
        for (int base = 0; __any_sync(kFull, base < nscan); base += 8) {
            const int i = base + g.lane;
            int4 ch = make_int4(0, 0, 0, 0), cc = make_int4(0, 0, 0, 0);
            unsigned key = 0;
            if (i < nscan) {
                ch = cx.hot[fc + i];
                cc = cx.cold[fc + i];
                float score;
                if (ch.x == 0) {
                    score = INFINITY;  // mcts.py:96-97
                } else if (ch.z & kCacheValid) {
                    score = __int_as_float(ch.w);  // mcts.py:99-100
                } else {
                    score = score_child(ch.x, __int_as_float(ch.y), ch.z & kVLMask, __int_as_float(cc.x),
                                        (cc.z >> 16) & 3, cx.c_puct, sq);
                    ch.w = __float_as_int(score);
                    ch.z |= kCacheValid;
                    reinterpret_cast<int2*>(&cx.hot[fc + i])[1] = make_int2(ch.z, ch.w);
                }
                key = (score == score) ? ordered_key(score) : 0u;
            }
            const unsigned mx = grp_max(key);
            const unsigned bal = (__ballot_sync(kFull, key == mx) >> g.sh) & 0xFFu;  // never empty: mx is one of the keys
            const int src = __ffs(bal) - 1;
            const int4 nh = grp_shfl4(ch, src), nc = grp_shfl4(cc, src);
            if (mx > best_key) {  // strict: an earlier chunk keeps ties (mcts.py:425)
                best_key = mx;
                best_i = base + src;
                bh = nh;
                bc = nc;
            }
        }
        if (going && best_i < 0) { cx.overflow |= 2; going = false; }
        grp_apply_move(cx.dir, b, (bc.z >> 8) & 0x3F, going);  // game.make_move(*next_move) (mcts.py:439)
        if (going) {
            ++cx.steps;
            node = fc + best_i;
            h = bh;
            c = bc;
            if (plen < kMaxPath) {
                if (g.lane == 0) g.path[plen] = node;
                ++plen;
            } else {
                cx.overflow |= 4;
                going = false;
            }
        }
    }
    leaf_vlf = h.z;
    __syncwarp();
    return node;
}

// node.expand (mcts.py:141-161, 605-618) with the uniform prior of the built-in evaluators:
// lane l creates the children whose squares lie in board row l; child index = rank of the square
__device__ __forceinline__ void expand_node8(TreeCtx8& cx, int node, uint64_t lm, float prior, bool act) {
    const Grp& g = cx.g;
    if (act) {
        int4 c = cx.cold[node];
        const int nc = popc64(lm);
        if ((c.z & 0xFF) != 0) {
            // 'if action not in self.children' (mcts.py:154): already expanded, nothing to add
        } else if (cx.n_nodes + nc > cx.cap) {
            cx.overflow |= 1;
        } else {
            const int fc = cx.n_nodes;
            const int turn = 3 - ((c.z >> 16) & 3);  // mcts.py:618
            const unsigned w = g.lane < 4 ? (unsigned)lm : (unsigned)(lm >> 32);
            unsigned byte = (w >> ((g.lane & 3) * 8)) & 0xFFu;
            int i = fc + popc64(lm & g.below);
            while (byte) {
                const int sq = g.lane * 8 + (__ffs(byte) - 1);
                byte &= byte - 1;
                cx.hot[i] = make_int4(0, 0, 0, 0);
                cx.cold[i] = make_int4(__float_as_int(prior), -1, (sq << 8) | (turn << 16), 0);
                ++i;
            }
            if (g.lane == 0) {
                c.y = fc;
                c.z = (c.z & ~0xFF) | nc;
                cx.cold[node] = c;
            }
            cx.n_nodes += nc;
            cx.created += (unsigned)nc;
            cx.bytes += 32u * (unsigned)nc;
        }
    }
    __syncwarp();
}

// one simulation per group (select -> evaluate -> expand -> backup), same results as simulate_one()
template <int EVAL>
__device__ __forceinline__ void simulate_one8(TreeCtx8& cx, const CoopBoard& root_c, uint64_t stream_for_sim, bool act) {
    CoopBoard b = root_c;
    int plen, vlf;
    const int node = select_one8(cx, b, plen, vlf, act);
    if (act) ++cx.sims;
    const bool term = act && (vlf & kTerminal);          // mcts.py:364-366: back its stored value up
    const uint64_t lm = grp_legal(cx.dir, b);
    const bool dead = act && !term && lm == 0;           // mcts.py:567-579: flag terminal, ABSOLUTE value
    const bool eval = act && !term && lm != 0;
    float v = 0.0f;
    if (EVAL == RVS_EVAL_E0) {
        v = __fdiv_rn((float)(popc64(b.Pd) - popc64(b.Od)), 64.0f);
    } else {
        const int leaf_side = b.side;
        CoopBoard r = b;
        const int plies = grp_random_playout(cx.g, cx.dir, r, eval ? lm : 0ULL, stream_for_sim);
        if (eval) cx.steps += (unsigned)plies;
        const int w = (r.flags & F_WIN_MASK) >> F_WIN_SHIFT;
        v = (!(r.flags & F_OVER) || w == 0) ? 0.0f : (w == leaf_side ? 1.0f : -1.0f);
    }
    if (eval) ++cx.evals;
    if (term) v = term_value_of(vlf);
    if (dead) {
        const int w = (b.flags & F_WIN_MASK) >> F_WIN_SHIFT;
        const int code = !(b.flags & F_OVER) ? 0 : (w == 1 ? 1 : (w == 2 ? 2 : 0));
        if (cx.g.lane == 0) {
            int* z = &reinterpret_cast<int*>(&cx.hot[node])[2];
            *z = (*z & ~(3 << kTermShift)) | kTerminal | (code << kTermShift);
        }
        v = code == 1 ? 1.0f : (code == 2 ? -1.0f : 0.0f);
    }
    expand_node8(cx, node, lm, 1.0f / 65.0f, eval);  // ends with __syncwarp(): the flag above is visible
    backup_path8(cx, plen, v, act);
}

}  // namespace rvs

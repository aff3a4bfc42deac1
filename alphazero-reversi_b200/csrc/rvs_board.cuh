// rvs_board.cuh -- K1: bitboard move generation / flipping / apply for sm_100a.
//
// Packed uint64 black/white pairs, bit i = row*8+col, LSB first
// (reference: src/game/board.py:31-32,49,170).  Two rule sets:
//   RULES_REF    bug-compatible with the reference board (graded mode): no file masks in
//                move generation (board.py:102-124) and abs(d)-indexed edge masks in the
//                flip scan (board.py:196-208).
//   RULES_STRICT true Othello.
//
// The formulas are branch-free Kogge-Stone style fills.  REF move generation floods
// exactly 1+5 steps like the reference loop (board.py:114), done as 1,1,2,2 (doubling
// with a run-of-two mask) instead of six unit steps; see legal_dir().
//
// Everything here is `__host__ __device__` so that the *same bit formulas* the kernels
// execute can be compiled by g++ for the CPU-side unit tests (tests/test_board_formulas.py
// builds csrc/rvs_hostcheck.cpp).  That host build is a test artefact, not a product path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RVS_HD __host__ __device__ __forceinline__
#else
#define RVS_HD inline
#endif

namespace rvs {

enum : int { RULES_REF = 0, RULES_STRICT = 1 };

constexpr uint64_t kNotA = 0xFEFEFEFEFEFEFEFEULL;  // col != 0
constexpr uint64_t kNotH = 0x7F7F7F7F7F7F7F7FULL;  // col != 7
constexpr uint64_t kAll = 0xFFFFFFFFFFFFFFFFULL;
constexpr uint64_t kStartBlack = 0x0000000810000000ULL;  // board.py:31
constexpr uint64_t kStartWhite = 0x0000001008000000ULL;  // board.py:32

RVS_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
RVS_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

template <int S>
RVS_HD uint64_t shl_dir(uint64_t x) {  // S > 0: <<S ; S < 0: >>-S  (compile-time direction)
    if constexpr (S > 0) return x << S; else return x >> (-S);
}

// ---- move generation -------------------------------------------------------------
// One direction of Board.get_valid_moves (board.py:102-124).  `Om` is the opponent set a
// run may pass through: plain O in REF rules; O restricted to the non-wrapping files in
// STRICT rules (masking the inner columns is the classic equivalent of masking each shift).
// Flood: c1 = sh(P)&Om, then five more unit steps in the reference; here unit, then two
// double steps through Om2 = Om & sh(Om)  ->  reach 1+1+2+2 = 6 cells, identical set.
template <int S>
RVS_HD uint64_t legal_dir(uint64_t P, uint64_t Om, uint64_t E) {
    uint64_t c = shl_dir<S>(P) & Om;           // run length >= 1
    c |= shl_dir<S>(c) & Om;                   // <= 2
    uint64_t Om2 = Om & shl_dir<S>(Om);
    c |= shl_dir<2 * S>(c) & Om2;              // <= 4
    c |= shl_dir<2 * S>(c) & Om2;              // <= 6  (reference stops here: range(5))
    return shl_dir<S>(c) & E;
}

template <int RULES>
RVS_HD uint64_t legal_moves(uint64_t P, uint64_t O) {
    const uint64_t E = ~(P | O);
    if constexpr (RULES == RULES_REF) {
        // shifts wrap around rows exactly like the reference (no file masks)
        return legal_dir<1>(P, O, E) | legal_dir<-1>(P, O, E) | legal_dir<8>(P, O, E) |
               legal_dir<-8>(P, O, E) | legal_dir<9>(P, O, E) | legal_dir<-9>(P, O, E) |
               legal_dir<7>(P, O, E) | legal_dir<-7>(P, O, E);
    } else {
        const uint64_t Oh = O & (kNotA & kNotH);  // inner columns: no horizontal wrap
        return legal_dir<1>(P, Oh, E) | legal_dir<-1>(P, Oh, E) | legal_dir<8>(P, O, E) |
               legal_dir<-8>(P, O, E) | legal_dir<9>(P, Oh, E) | legal_dir<-9>(P, Oh, E) |
               legal_dir<7>(P, Oh, E) | legal_dir<-7>(P, Oh, E);
    }
}

// ---- flips -----------------------------------------------------------------------
// One direction of the flip scan (board.py:205-219).  Om = cells a line may pass through
// (O & m), Pm = cells that may close it (P & m); m is the reference's (mis-indexed) edge
// mask.  The reference walks <= 7 steps; a line of 7 opponents cannot be closed, so
// flooding 6 cells and testing the next one is equivalent (proof in DESIGN.md "K1").
template <int S>
RVS_HD uint64_t flip_dir(uint64_t mv, uint64_t Om, uint64_t Pm) {
    uint64_t x = shl_dir<S>(mv) & Om;          // 1
    x |= shl_dir<S>(x) & Om;                   // 2
    uint64_t Om2 = Om & shl_dir<S>(Om);
    x |= shl_dir<2 * S>(x) & Om2;              // 4
    x |= shl_dir<2 * S>(x) & Om2;              // 6
    // x is the contiguous run starting next to mv; the closing cell is the one after it
    uint64_t end = shl_dir<S>(x) & ~x & Pm;
    return end ? x : 0ULL;
}

template <int RULES>
RVS_HD uint64_t flip_mask(uint64_t P, uint64_t O, uint64_t mv) {
    if constexpr (RULES == RULES_REF) {
        // edge_masks.get(abs(d)) (board.py:196-208): |d|==1 -> notA, 7 -> notA, 9 -> notH, 8 -> all
        const uint64_t O1 = O & kNotA, P1 = P & kNotA;  // d = +-1 and +-7
        const uint64_t O9 = O & kNotH, P9 = P & kNotH;  // d = +-9
        return flip_dir<1>(mv, O1, P1) | flip_dir<-1>(mv, O1, P1) | flip_dir<8>(mv, O, P) |
               flip_dir<-8>(mv, O, P) | flip_dir<7>(mv, O1, P1) | flip_dir<-7>(mv, O1, P1) |
               flip_dir<9>(mv, O9, P9) | flip_dir<-9>(mv, O9, P9);
    } else {
        // true Othello: a step that wraps lands on the far file, so mask the landing file
        const uint64_t Oe = O & kNotA, Pe = P & kNotA;  // moving east-ish (+1,+9,-7): never land on col 0
        const uint64_t Ow = O & kNotH, Pw = P & kNotH;  // moving west-ish (-1,-9,+7): never land on col 7
        return flip_dir<1>(mv, Oe, Pe) | flip_dir<-1>(mv, Ow, Pw) | flip_dir<8>(mv, O, P) |
               flip_dir<-8>(mv, O, P) | flip_dir<7>(mv, Ow, Pw) | flip_dir<-7>(mv, Oe, Pe) |
               flip_dir<9>(mv, Oe, Pe) | flip_dir<-9>(mv, Ow, Pw);
    }
}

// ---- game state --------------------------------------------------------------------
// flags byte: bit0 game_over, bits1-2 winner (0 draw,1 black,2 white), bit3 last move was
// followed by an auto-pass (Board.passed_moves_in_a_row, board.py:244).
enum : uint8_t { F_OVER = 1, F_WIN_SHIFT = 1, F_WIN_MASK = 6, F_PASSED = 8 };

struct Board {
    uint64_t black, white;
    uint8_t side;   // 1 BLACK, 2 WHITE
    uint8_t flags;
};

RVS_HD Board start_board() { return Board{kStartBlack, kStartWhite, 1, 0}; }
RVS_HD bool is_over(const Board& b) { return b.flags & F_OVER; }
RVS_HD int winner_of(const Board& b) { return (b.flags & F_WIN_MASK) >> F_WIN_SHIFT; }

template <int RULES>
RVS_HD uint64_t board_legal(const Board& b) {
    return b.side == 1 ? legal_moves<RULES>(b.black, b.white) : legal_moves<RULES>(b.white, b.black);
}

// Board.make_move after the legality check (board.py:181-251): flips, side switch,
// auto-pass, terminal + winner.  `next_legal` receives the legal mask of the side to move
// afterwards (0 when the game is over) so callers never recompute it.
template <int RULES>
RVS_HD void apply_move(Board& b, int idx, uint64_t& next_legal) {
    const bool blk = b.side == 1;
    uint64_t P = blk ? b.black : b.white;
    uint64_t O = blk ? b.white : b.black;
    const uint64_t mv = 1ULL << idx;
    const uint64_t f = flip_mask<RULES>(P, O, mv);
    P ^= mv | f;
    O ^= f;
    b.black = blk ? P : O;
    b.white = blk ? O : P;
    uint8_t flags = 0;
    uint64_t lm = legal_moves<RULES>(O, P);  // opponent to move
    uint8_t side = (uint8_t)(3 - b.side);
    if (lm == 0) {                           // board.py:242-249
        lm = legal_moves<RULES>(P, O);
        side = b.side;
        flags = F_PASSED;
        if (lm == 0) {
            const int nb = popc64(b.black), nw = popc64(b.white);
            const int w = nb > nw ? 1 : (nw > nb ? 2 : 0);
            flags = (uint8_t)(F_PASSED | F_OVER | (w << F_WIN_SHIFT));
        }
    }
    b.side = side;
    b.flags = flags;
    next_legal = lm;
}

// ReversiGame.make_move contract (game.py:47-48,70): false when over or not in the legal mask
template <int RULES>
RVS_HD bool try_move(Board& b, int idx, uint64_t& next_legal) {
    if (is_over(b) || idx < 0 || idx > 63) return false;
    if (!((board_legal<RULES>(b) >> idx) & 1)) return false;
    apply_move<RULES>(b, idx, next_legal);
    return true;
}

// k-th (0-based) set bit of m, ascending
RVS_HD int nth_set_bit(uint64_t m, int k) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    int nlo = __popc(lo);
    if (k < nlo) return __fns(lo, 0, k + 1);
    return 32 + __fns(hi, 0, k - nlo + 1);
#else
    while (k--) m &= m - 1;
    return __builtin_ctzll(m);
#endif
}

// ---- shared counter RNG (DESIGN.md "RNG"; mirrored by oracle/rvs_oracle.c) ----------
RVS_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
RVS_HD uint64_t stream_seed(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = mix64(seed + 0x9E3779B97F4A7C15ULL * (a + 1));
    return mix64(x ^ (0xD1B54A32D192ED03ULL * (b + 1)));
}
RVS_HD uint64_t rng_next(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ULL;
    return mix64(s);
}
RVS_HD int rng_pick(uint64_t r, int n) { return (int)(((r >> 32) * (uint64_t)n) >> 32); }

// playout move picker: PCG-RXS-M-XS-32 seeded from a 64-bit stream id (~9 SASS instructions per
// draw instead of ~25 for splitmix64); move = k-th legal square with k = (r * n) >> 32
RVS_HD uint32_t roll_init(uint64_t stream) { return (uint32_t)(stream ^ (stream >> 32)); }
RVS_HD uint32_t roll_next(uint32_t& s) {
    s = s * 747796405u + 2891336453u;
    const uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
RVS_HD int roll_pick(uint32_t r, int n) { return (int)(((uint64_t)r * (uint64_t)n) >> 32); }

// ---- direction-sliced formulas (warp-cooperative board ops) ---------------------------------
// In the tree kernels one warp owns one game, so a scalar apply_move would leave 31 lanes idle.
// Instead lane l evaluates direction (l & 7) of the flip / move-generation scan and the warp
// ORs the eight partial masks with two REDUX instructions (coop_* below).  To keep every lane
// on the same instruction stream, "negative" directions (>>) run on bit-reversed boards with
// a left shift: brev(x >> s) == brev(x) << s.  The per-direction functions are plain
// host/device code so the CPU test tier checks OR_d part(d) == oracle.
RVS_HD uint64_t brev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}

// Shifts of the direction scans.  The shift amount is per lane (a register), so x << s costs two ALU-pipe SHF, and the
// ALU pipe (LOP3 + SHF, one warp instruction per two cycles) is what the rollout saturates.  HYBRID form (default on the
// device): the low word is one IMAD on the otherwise idle FMA pipe (lo * 2^s), the high word one funnel shift --
// 14 of the ~94 ALU instructions of a rollout ply move over: +2 % at 4096 games, +8 % at 16 384 (same-box A/B on
// B200, gpurun_out/ab_k2_s4.log).  Writing the WHOLE shift as multiplies (IMAD.WIDE.U32 + IMAD, RVS_SHL_MUL) puts two
// dependent FMA-pipe instructions on the flood chain and was ~15 % SLOWER end to end (round 1); RVS_SHL_PLAIN keeps the
// two-SHF form for A/B runs.
struct DirLane {
    uint32_t m1, m2;  // 2^s and 2^(2s)
    int s, s2;      // |shift| in {1, 7, 8, 9} and twice that
    bool neg;       // direction is a right shift -> work on bit-reversed boards
    uint64_t fm;    // flip scan: mask of cells a line may pass through / close on (working domain)
    uint64_t gm;    // move generation: mask applied to the opponent set (working domain)
};

template <int RULES>
RVS_HD DirLane make_dir(int d) {
    // d: 0 +1(E) 1 -1(W) 2 +8(S) 3 -8(N) 4 +9(SE) 5 -9(NW) 6 +7(SW) 7 -7(NE)
    DirLane L;
    const int a = d >> 1;
    L.s = a == 0 ? 1 : (a == 1 ? 8 : (a == 2 ? 9 : 7));
    L.neg = d & 1;
    L.s2 = 2 * L.s;
    L.m1 = 1u << L.s;
    L.m2 = 1u << (2 * L.s);
    uint64_t fm, gm;
    if (RULES == RULES_REF) {
        fm = a == 0 ? kNotA : (a == 1 ? kAll : (a == 2 ? kNotH : kNotA));  // edge_masks.get(abs(d)) (board.py:196-208)
        gm = kAll;                                                         // no file masks (board.py:102-124)
    } else {
        // landing file: east-ish steps (+1,+9,-7) never land on col 0, west-ish (-1,-9,+7) never on col 7
        const bool east = (d == 0 || d == 4 || d == 7);
        fm = a == 1 ? kAll : (east ? kNotA : kNotH);
        gm = a == 1 ? kAll : (kNotA & kNotH);
    }
    L.fm = L.neg ? brev64(fm) : fm;
    L.gm = L.neg ? brev64(gm) : gm;
    return L;
}

RVS_HD uint64_t shl_mul(uint64_t x, uint32_t m) {
    const uint64_t p = (uint64_t)(uint32_t)x * m;
    const uint32_t hi = (uint32_t)(x >> 32) * m + (uint32_t)(p >> 32);
    return ((uint64_t)hi << 32) | (uint32_t)p;
}
#if defined(RVS_SHL_MUL)
RVS_HD uint64_t sh1(const DirLane& L, uint64_t x) { return shl_mul(x, L.m1); }
RVS_HD uint64_t sh2(const DirLane& L, uint64_t x) { return shl_mul(x, L.m2); }
#elif !defined(RVS_SHL_PLAIN) && defined(__CUDA_ARCH__)
// low word on the FMA pipe (one IMAD), high word one funnel shift on the ALU pipe
__device__ __forceinline__ uint64_t shl_hyb(uint64_t x, uint32_t m, int s) {
    uint32_t lo;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"((uint32_t)x), "r"(m));
    const uint32_t hi = __funnelshift_l((uint32_t)x, (uint32_t)(x >> 32), s);
    return ((uint64_t)hi << 32) | lo;
}
RVS_HD uint64_t sh1(const DirLane& L, uint64_t x) { return shl_hyb(x, L.m1, L.s); }
RVS_HD uint64_t sh2(const DirLane& L, uint64_t x) { return shl_hyb(x, L.m2, L.s2); }
#else
RVS_HD uint64_t sh1(const DirLane& L, uint64_t x) { return x << L.s; }
RVS_HD uint64_t sh2(const DirLane& L, uint64_t x) { return x << L.s2; }
#endif

RVS_HD uint64_t to_dom(uint64_t x, bool neg) { return neg ? brev64(x) : x; }

// contribution of one direction to Board.get_valid_moves; Pd/Od and the result in the lane's
// working domain
RVS_HD uint64_t legal_raw(const DirLane& L, uint64_t Pd, uint64_t Od) {
    const uint64_t E = ~(Pd | Od), Om = Od & L.gm;
    uint64_t c = sh1(L, Pd) & Om;
    c |= sh1(L, c) & Om;
    const uint64_t Om2 = Om & sh1(L, Om);
    c |= sh2(L, c) & Om2;
    c |= sh2(L, c) & Om2;
    return sh1(L, c) & E;
}
// same, result in the normal domain
RVS_HD uint64_t legal_part(const DirLane& L, uint64_t Pd, uint64_t Od) { return to_dom(legal_raw(L, Pd, Od), L.neg); }

// contribution of one direction to the flip scan for the move bit `mvd` (all in the working domain)
RVS_HD uint64_t flip_raw(const DirLane& L, uint64_t Pd, uint64_t Od, uint64_t mvd) {
    const uint64_t Om = Od & L.fm, Pm = Pd & L.fm;
    uint64_t x = sh1(L, mvd) & Om;
    x |= sh1(L, x) & Om;
    const uint64_t Om2 = Om & sh1(L, Om);
    x |= sh2(L, x) & Om2;
    x |= sh2(L, x) & Om2;
    const uint64_t end = sh1(L, x) & ~x & Pm;
    return end ? x : 0ULL;
}
// The same scan without the flood.  flip_ray = the cells the scan of move square `sqd` (working domain) can touch:
// sqd + s, sqd + 2s, ... (at most 7: six cells to pass through and the one that closes the line), cut before the
// first cell that leaves the board or lies outside fm (such a cell can neither be passed nor close a line).
// flip_carry: with every non-ray bit and every opponent cell set, adding 1 ripples a carry from bit 0 through the
// move square and along the run of opponent cells; it stops on the first ray cell that is not an opponent disc
// (or runs off the top when there is none).  That cell closes the line iff it holds a disc of the mover, and the
// opponent ray cells the ripple cleared are the flips.  Equal to flip_raw for every position (tests/test_capi_cpu.py checks the host
// build on random positions; the GPU parity tests cover the kernels).
RVS_HD uint64_t flip_ray(const DirLane& L, int sqd) {
    uint64_t R = 0;
    for (int i = 1; i <= 7; ++i) {
        const int c = sqd + i * L.s;
        if (c > 63 || !((L.fm >> c) & 1ULL)) break;
        R |= 1ULL << c;
    }
    return R;
}
RVS_HD uint64_t flip_carry(uint64_t R, uint64_t Pd, uint64_t Od) {
    const uint64_t X = (Od | ~R) + 1ULL;
    const uint64_t closed = X & R & Pd;  // the cell the carry stopped on, if it is a ray cell with a disc of the mover
    const uint64_t run = Od & R & ~X;    // the opponent cells the carry went through (cleared by the ripple)
    return closed ? run : 0ULL;
}
RVS_HD uint64_t flip_part(const DirLane& L, uint64_t Pd, uint64_t Od, uint64_t mvd) {
    return to_dom(flip_raw(L, Pd, Od, mvd), L.neg);
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t warp_or64(uint64_t x) {
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)x);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(x >> 32));
    return ((uint64_t)hi << 32) | lo;
}

// A position held by the warp: every lane keeps the mover's / opponent's discs in ITS OWN
// working domain (bit-reversed for right-shift directions), so a ply costs one brev of the
// reduced flip mask instead of re-reversing both boards.  popcounts are domain independent.
struct CoopBoard {
    uint64_t Pd, Od;  // side to move / opponent, lane domain
    int side;         // 1 BLACK, 2 WHITE
    int flags;        // F_OVER | winner | F_PASSED like Board::flags
};

__device__ __forceinline__ CoopBoard coop_load(const DirLane& L, const Board& b) {
    const bool blk = b.side == 1;
    return CoopBoard{to_dom(blk ? b.black : b.white, L.neg), to_dom(blk ? b.white : b.black, L.neg), b.side, b.flags};
}
// normal-domain view (every lane gets the same Board)
__device__ __forceinline__ Board coop_store(const DirLane& L, const CoopBoard& c) {
    const uint64_t P = to_dom(c.Pd, L.neg), O = to_dom(c.Od, L.neg);
    return Board{c.side == 1 ? P : O, c.side == 1 ? O : P, (uint8_t)c.side, (uint8_t)c.flags};
}

// Board.get_valid_moves for the side to move, by the whole warp (normal domain, warp-uniform)
__device__ __forceinline__ uint64_t coop_legal(const DirLane& L, const CoopBoard& c) {
    return warp_or64(legal_part(L, c.Pd, c.Od));
}

// k-th set bit of a warp-uniform mask, lanes test bits (lane) and (lane+32) in parallel
__device__ __forceinline__ int coop_nth_set_bit(uint64_t m, int k, int lane) {
    const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
    const unsigned lt = (1u << lane) - 1u;
    const int nlo = __popc(lo);
    const bool hit_lo = ((lo >> lane) & 1u) && __popc(lo & lt) == k;
    const bool hit_hi = ((hi >> lane) & 1u) && (nlo + __popc(hi & lt)) == k;
    const unsigned blo = __ballot_sync(0xffffffffu, hit_lo), bhi = __ballot_sync(0xffffffffu, hit_hi);
    return blo ? (__ffs(blo) - 1) : (31 + __ffs(bhi));
}

// apply_move() by the whole warp (board.py:181-251).  Returns the legal mask of the side to move
// afterwards (0 when the game is over); every lane computes the same side / flags.
__device__ __forceinline__ uint64_t coop_apply_move(const DirLane& L, CoopBoard& c, int idx) {
    const uint64_t mvd = 1ULL << (L.neg ? 63 - idx : idx);
    const uint64_t f = warp_or64(flip_part(L, c.Pd, c.Od, mvd));
    const uint64_t fd = to_dom(f, L.neg);
    const uint64_t P = c.Pd ^ (mvd | fd), O = c.Od ^ fd;
    uint64_t lm = warp_or64(legal_part(L, O, P));  // opponent to move
    if (lm) {                                      // warp-uniform branch
        c.Pd = O; c.Od = P; c.side = 3 - c.side; c.flags = 0;
        return lm;
    }
    c.Pd = P; c.Od = O;                            // auto-pass (board.py:242-249)
    lm = warp_or64(legal_part(L, P, O));
    c.flags = F_PASSED;
    if (lm == 0) {
        const int np = popc64(P), no = popc64(O);
        const int nb = c.side == 1 ? np : no, nw = c.side == 1 ? no : np;
        const int w = nb > nw ? 1 : (nw > nb ? 2 : 0);
        c.flags = F_PASSED | F_OVER | (w << F_WIN_SHIFT);
    }
    return lm;
}

// random_playout() by the whole warp from a position whose legal mask `lm` is known
__device__ __forceinline__ int coop_random_playout(const DirLane& L, CoopBoard& c, uint64_t lm, uint64_t stream, int lane) {
    int plies = 0;
    uint32_t rs = roll_init(stream);
    while (lm) {
        const int k = roll_pick(roll_next(rs), popc64(lm));
        lm = coop_apply_move(L, c, coop_nth_set_bit(lm, k, lane));
        ++plies;
    }
    return plies;
}
#endif

// one uniform random playout to the end (config 1 / rollout evaluator); returns plies
template <int RULES>
RVS_HD int random_playout(Board& b, uint64_t stream) {
    int plies = 0;
    if (is_over(b)) return 0;
    uint64_t lm = board_legal<RULES>(b);
    uint32_t rs = roll_init(stream);
    while (lm) {
        const int k = roll_pick(roll_next(rs), popc64(lm));
        apply_move<RULES>(b, nth_set_bit(lm, k), lm);
        ++plies;
    }
    return plies;
}

}  // namespace rvs

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

// one uniform random playout to the end (config 1 / rollout evaluator); returns plies
template <int RULES>
RVS_HD int random_playout(Board& b, uint64_t stream) {
    int plies = 0;
    if (is_over(b)) return 0;
    uint64_t lm = board_legal<RULES>(b);
    while (lm) {
        const int k = rng_pick(rng_next(stream), popc64(lm));
        apply_move<RULES>(b, nth_set_bit(lm, k), lm);
        ++plies;
    }
    return plies;
}

}  // namespace rvs

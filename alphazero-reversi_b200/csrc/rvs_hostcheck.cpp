// rvs_hostcheck.cpp -- TEST ARTEFACT: compiles the kernels' bit formulas (rvs_board.cuh)
// with g++ so the CPU-only test tier can compare them with the oracle before any GPU time
// is spent.  Never loaded by the product package.
#include <math.h>

#include "rvs_board.cuh"
#include "rvs_noise.cuh"

using namespace rvs;

extern "C" {
uint64_t hc_legal(uint64_t P, uint64_t O, int rules) {
    return rules == RULES_STRICT ? legal_moves<RULES_STRICT>(P, O) : legal_moves<RULES_REF>(P, O);
}
uint64_t hc_flips(uint64_t P, uint64_t O, int idx, int rules) {
    return rules == RULES_STRICT ? flip_mask<RULES_STRICT>(P, O, 1ULL << idx)
                                 : flip_mask<RULES_REF>(P, O, 1ULL << idx);
}
int hc_try_move(uint64_t* black, uint64_t* white, uint8_t* side, uint8_t* flags, int idx, int rules,
                uint64_t* next_legal) {
    Board b{*black, *white, *side, *flags};
    uint64_t nl = 0;
    bool ok = rules == RULES_STRICT ? try_move<RULES_STRICT>(b, idx, nl) : try_move<RULES_REF>(b, idx, nl);
    *black = b.black; *white = b.white; *side = b.side; *flags = b.flags; *next_legal = nl;
    return ok ? 1 : 0;
}
int hc_playout(uint64_t* black, uint64_t* white, uint8_t* side, uint8_t* flags, uint64_t stream, int rules) {
    Board b{*black, *white, *side, *flags};
    int p = rules == RULES_STRICT ? random_playout<RULES_STRICT>(b, stream) : random_playout<RULES_REF>(b, stream);
    *black = b.black; *white = b.white; *side = b.side; *flags = b.flags;
    return p;
}
// direction-sliced variants: OR over the eight DirLane parts (what the warp computes with REDUX)
uint64_t hc_legal_sliced(uint64_t P, uint64_t O, int rules) {
    uint64_t v = 0;
    for (int d = 0; d < 8; ++d) {
        const DirLane L = rules == RULES_STRICT ? make_dir<RULES_STRICT>(d) : make_dir<RULES_REF>(d);
        v |= legal_part(L, to_dom(P, L.neg), to_dom(O, L.neg));
    }
    return v;
}
uint64_t hc_flips_sliced(uint64_t P, uint64_t O, int idx, int rules) {
    uint64_t v = 0;
    for (int d = 0; d < 8; ++d) {
        const DirLane L = rules == RULES_STRICT ? make_dir<RULES_STRICT>(d) : make_dir<RULES_REF>(d);
        const uint64_t mvd = L.neg ? (1ULL << (63 - idx)) : (1ULL << idx);
        v |= flip_part(L, to_dom(P, L.neg), to_dom(O, L.neg), mvd);
    }
    return v;
}
// the table form of the 8-lane groups: ray of the move square + carry ripple (flip_ray / flip_carry)
uint64_t hc_flips_carry(uint64_t P, uint64_t O, int idx, int rules) {
    uint64_t v = 0;
    for (int d = 0; d < 8; ++d) {
        const DirLane L = rules == RULES_STRICT ? make_dir<RULES_STRICT>(d) : make_dir<RULES_REF>(d);
        v |= to_dom(flip_carry(flip_ray(L, L.neg ? 63 - idx : idx), to_dom(P, L.neg), to_dom(O, L.neg)), L.neg);
    }
    return v;
}
int hc_nth_set_bit(uint64_t m, int k) { return nth_set_bit(m, k); }
uint64_t hc_stream_seed(uint64_t s, uint64_t a, uint64_t b) { return stream_seed(s, a, b); }
// Dirichlet root noise (rvs_noise.cuh): the product's formulas against the oracle's restatement
double hc_det_log(double x) { return det_log(x); }
double hc_det_exp(double x) { return det_exp(x); }
void hc_dirichlet(double alpha, int k, uint64_t stream, float* eta) { noise_dirichlet(alpha, k, stream, eta); }
float hc_noise_mix(float p, float eta, float eps) { return noise_mix(p, eta, eps); }
}

// rvs_noise.cuh -- Dirichlet noise on the root priors (BASELINE config 4).
//
// NEW-ENGINE FEATURE: the reference carries dirichlet_alpha / dirichlet_epsilon through its config
// (src/config.py:25-26, src/self_play/self_play.py:18-47) but never applies them, so
// there is no reference behaviour to match.  The specification below is this engine's own; the
// oracle restates it independently (oracle/rvs_oracle.c: root_noise) and the two must agree BIT FOR
// BIT, which is why every floating-point step is a single IEEE f64 operation (+ - * / sqrt, never an
// FMA) and log / exp are the fixed polynomial algorithms written out here, not library calls.
//
//   After the root's children are created by the first expansion of a search:
//     P'_i = f32( f32(1-eps) * P_i  +  f32(eps) * f32(eta_i) )        (two f32 products, one f32 sum)
//     eta  = softmax( lg ),  lg_i = log Gamma(alpha) deviate of child i (children in square order)
//   Gamma(a >= 1): Marsaglia-Tsang (d = a - 1/3, c = 1/sqrt(9d), accept log u < z^2/2 + d - dv + d log v)
//   Gamma(a <  1): log g = log Gamma(a+1) + log(u)/a                   (kept in log space: no underflow)
//   normal z: Marsaglia polar method;  uniforms u = ((r >> 11) + 1) * 2^-53 from rng_next()
//   stream: stream_seed(seed, game_id, 0xD1000000 + search_id)
#pragma once
#include "rvs_board.cuh"

namespace rvs {

#if defined(__CUDA_ARCH__)
#define RVS_DADD(a, b) __dadd_rn((a), (b))
#define RVS_DMUL(a, b) __dmul_rn((a), (b))
#define RVS_DDIV(a, b) __ddiv_rn((a), (b))
#define RVS_DSQRT(a) __dsqrt_rn((a))
#define RVS_FADD(a, b) __fadd_rn((a), (b))
#define RVS_FMUL(a, b) __fmul_rn((a), (b))
#else  // host build (tests/test_board_formulas.py) is compiled with -ffp-contract=off
#define RVS_DADD(a, b) ((a) + (b))
#define RVS_DMUL(a, b) ((a) * (b))
#define RVS_DDIV(a, b) ((a) / (b))
#define RVS_DSQRT(a) sqrt((a))
#define RVS_FADD(a, b) ((a) + (b))
#define RVS_FMUL(a, b) ((a) * (b))
#endif

RVS_HD double bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    __builtin_memcpy(&d, &u, 8);
    return d;
#endif
}
RVS_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    __builtin_memcpy(&u, &d, 8);
    return u;
#endif
}

// natural log of a positive normal double: x = m * 2^e with m in [sqrt(1/2), sqrt(2)),
// log m = 2 * atanh(s), s = (m-1)/(m+1), odd series up to s^23 (|s| <= 0.1716: error < 1e-18)
RVS_HD double det_log(double x) {
    uint64_t u = double_to_bits(x);
    int e = (int)((u >> 52) & 0x7FF) - 1023;
    u = (u & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;  // m in [1, 2)
    double m = bits_to_double(u);
    if (m > 1.4142135623730951) { m = RVS_DMUL(m, 0.5); e += 1; }
    const double s = RVS_DDIV(RVS_DADD(m, -1.0), RVS_DADD(m, 1.0));
    const double s2 = RVS_DMUL(s, s);
    double p = 1.0 / 23.0;
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 21.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 19.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 17.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 15.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 13.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 11.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 9.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 7.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 5.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0 / 3.0);
    p = RVS_DADD(RVS_DMUL(p, s2), 1.0);
    const double lm = RVS_DMUL(RVS_DMUL(2.0, s), p);
    return RVS_DADD(RVS_DMUL((double)e, 0.6931471805599453), lm);
}

// exp(x) for x <= 0 (softmax terms): x = k ln2 + r, |r| <= ln2/2, Taylor to r^13, scaled by 2^k;
// anything below 2^-1000 is returned as 0
RVS_HD double det_exp(double x) {
    if (!(x > -690.0)) return 0.0;
    const double kf = floor(RVS_DADD(RVS_DMUL(x, 1.4426950408889634), 0.5));
    const double r = RVS_DADD(x, -RVS_DMUL(kf, 0.6931471805599453));
    double p = 1.0 / 6227020800.0;
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 479001600.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 39916800.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 3628800.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 362880.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 40320.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 5040.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 720.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 120.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 24.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0 / 6.0);
    p = RVS_DADD(RVS_DMUL(p, r), 0.5);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0);
    p = RVS_DADD(RVS_DMUL(p, r), 1.0);
    const int k = (int)kf;  // in [-996, 1]
    return RVS_DMUL(p, bits_to_double((uint64_t)(k + 1023) << 52));
}

RVS_HD double noise_uniform(uint64_t& s) {  // (0, 1]
    return RVS_DMUL((double)((rng_next(s) >> 11) + 1ULL), 1.0 / 9007199254740992.0);
}

RVS_HD double noise_normal(uint64_t& s) {  // Marsaglia polar, one deviate per accepted pair
    while (true) {
        const double v1 = RVS_DADD(RVS_DMUL(2.0, noise_uniform(s)), -1.0);
        const double v2 = RVS_DADD(RVS_DMUL(2.0, noise_uniform(s)), -1.0);
        const double q = RVS_DADD(RVS_DMUL(v1, v1), RVS_DMUL(v2, v2));
        if (q >= 1.0 || q == 0.0) continue;
        return RVS_DMUL(v1, RVS_DSQRT(RVS_DDIV(RVS_DMUL(-2.0, det_log(q)), q)));
    }
}

// log of a Gamma(a, 1) deviate
RVS_HD double noise_log_gamma(double a, uint64_t& s) {
    const double a1 = a < 1.0 ? RVS_DADD(a, 1.0) : a;
    const double d = RVS_DADD(a1, -1.0 / 3.0);
    const double c = RVS_DDIV(1.0, RVS_DSQRT(RVS_DMUL(9.0, d)));
    double lg;
    while (true) {
        const double z = noise_normal(s);
        const double t = RVS_DADD(1.0, RVS_DMUL(c, z));
        if (t <= 0.0) continue;
        const double v = RVS_DMUL(RVS_DMUL(t, t), t);
        const double lu = det_log(noise_uniform(s));
        const double lv = det_log(v);
        // log u < z^2/2 + d - d v + d log v
        const double rhs = RVS_DADD(RVS_DADD(RVS_DADD(RVS_DMUL(0.5, RVS_DMUL(z, z)), d), -RVS_DMUL(d, v)), RVS_DMUL(d, lv));
        if (lu < rhs) { lg = RVS_DADD(det_log(d), lv); break; }
    }
    if (a < 1.0) lg = RVS_DADD(lg, RVS_DDIV(det_log(noise_uniform(s)), a));
    return lg;
}

// eta[0..k) ~ Dirichlet(alpha) as f32 (k <= 64); deterministic in (stream, alpha, k)
RVS_HD void noise_dirichlet(double alpha, int k, uint64_t stream, float* eta) {
    double lg[64];
    uint64_t s = stream;
    double mx = -1e300;
    for (int i = 0; i < k; ++i) {
        lg[i] = noise_log_gamma(alpha, s);
        if (lg[i] > mx) mx = lg[i];
    }
    double sum = 0.0;
    for (int i = 0; i < k; ++i) {
        lg[i] = det_exp(RVS_DADD(lg[i], -mx));
        sum = RVS_DADD(sum, lg[i]);
    }
    for (int i = 0; i < k; ++i) eta[i] = (float)RVS_DDIV(lg[i], sum);
}

RVS_HD float noise_mix(float prior, float eta, float eps) {
    return RVS_FADD(RVS_FMUL(RVS_FADD(1.0f, -eps), prior), RVS_FMUL(eps, eta));
}

}  // namespace rvs

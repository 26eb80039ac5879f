// Rounding-exact building blocks of the two-stream sweep (shared by fband.cu and fband_cp.cu).
//
// Several terms of the reference's flux update are differences of nearly equal products (the direct-beam
// source is G * (M F_i - P F_{i+1})-like, and G itself reaches 1e8), so the place where a compiler fuses a
// multiply-add changes the result at the 1e-10 level.  To agree with the reference's kernels.cu as built
// by nvcc (and to make the column-serial and the layer-parallel kernel agree with each other), every
// such term is written with explicit round-to-nearest intrinsics in the operation order and fusion that nvcc/ptxas 12.9 emit for
// the reference's source expressions (PTX + SASS of K:1443-1451, 1497-1505, 1640-1683, 1744-1787; verified by
// the bit-for-bit comparisons in tests/test_gpu_parity.py).
#pragma once

// min(0, Fa/(-mu) * (g1*c1 + g2*c2) - Fb/(-mu) * m1 * m2)       (K:1447-1449 and its seven siblings)
__device__ __forceinline__ double beam_source(double Fa, double Fb, double neg_mu, double c1, double g1,
                                              double c2, double g2, double m1, double m2) {
    const double s = __fma_rn(c2, g2, __dmul_rn(c1, g1));
    const double t1 = __dmul_rn(s, __ddiv_rn(Fa, neg_mu));
    const double u = __dmul_rn(m1, __ddiv_rn(Fb, neg_mu));
    return fmin(__fma_rn(-m2, u, t1), 0.0);  // ptxas fuses the last product into the subtraction
}

// 2 pi eps (1 - w0) / (E - w0)                                   (K:1451)
__device__ __forceinline__ double source_factor(double epsi, double w0, double E) {
    const double two_pi_eps = __dmul_rn(epsi, 6.283185307179586);  // 2.0 * PI folded by the compiler
    return __ddiv_rn(__dmul_rn(two_pi_eps, __dsub_rn(1.0, w0)), __dsub_rn(E, w0));
}

// 1/M * (P F - N F_opp + fac * planck_terms + direct_terms)      (K:1443-1451)
__device__ __forceinline__ double sweep_update(double invM, double P, double N, double F, double F_opp,
                                               double fac, double planck_terms, double direct_terms) {
    const double flux_terms = __fma_rn(P, F, -__dmul_rn(N, F_opp));
    return __dmul_rn(invM, __dadd_rn(__fma_rn(fac, planck_terms, flux_terms), direct_terms));
}

// isothermal layer: B (N + M - P)                                (K:1445)
__device__ __forceinline__ double planck_iso(double B, double M, double N, double P) {
    return __dmul_rn(__dsub_rn(__dadd_rn(M, N), P), B);
}

// optically thin half-layer: (Ba + Bb)/2 (N + M - P)             (K:1642)
__device__ __forceinline__ double planck_thin(double Ba, double Bb, double M, double N, double P) {
    return __dmul_rn(__dsub_rn(__dadd_rn(M, N), P), __dmul_rn(__dadd_rn(Ba, Bb), 0.5));
}

// eps / (E (1 - w0 g0))                                          (K:1648)
__device__ __forceinline__ double gradient_factor(double epsi, double w0, double g0, double E) {
    return __ddiv_rn(epsi, __dmul_rn(__dsub_rn(1.0, __dmul_rn(w0, g0)), E));
}

// downward form: B1 (M + N) - B2 P + pre (P - M + N) pgrad       (K:1648, 1675)
__device__ __forceinline__ double planck_grad_down(double B1, double B2, double M, double N, double P,
                                                   double pre, double pgrad) {
    const double base = __fma_rn(__dadd_rn(M, N), B1, -__dmul_rn(P, B2));
    return __fma_rn(__dmul_rn(__dadd_rn(N, __dsub_rn(P, M)), pre), pgrad, base);
}

// upward form: B1 (M + N) - B2 P + pre pgrad (M - P - N)         (K:1752, 1779)
__device__ __forceinline__ double planck_grad_up(double B1, double B2, double M, double N, double P,
                                                 double pre, double pgrad) {
    const double base = __fma_rn(__dadd_rn(M, N), B1, -__dmul_rn(P, B2));
    return __fma_rn(__dsub_rn(__dsub_rn(M, P), N), __dmul_rn(pre, pgrad), base);
}

// surface: A (F_dir + F_down) + (1 - A) pi (1 - w0)/(E - w0) B_surf   (K:1469-1474)
__device__ __forceinline__ double boa_flux(double A_s, double Fdir0, double Fd0, double w0, double E,
                                           double B_surf) {
    const double emis = __dmul_rn(
        __ddiv_rn(__dmul_rn(__dmul_rn(__dsub_rn(1.0, A_s), 3.141592653589793), __dsub_rn(1.0, w0)),
                  __dsub_rn(E, w0)),
        B_surf);
    return __fma_rn(A_s, __dadd_rn(Fdir0, Fd0), emis);
}

// the emission part of boa_flux alone: boa_flux(A, Fdir, Fd, ...) == fma(A, Fdir + Fd, boa_emission(A, ...))
__device__ __forceinline__ double boa_emission(double A_s, double w0, double E, double B_surf) {
    return __dmul_rn(
        __ddiv_rn(__dmul_rn(__dmul_rn(__dsub_rn(1.0, A_s), 3.141592653589793), __dsub_rn(1.0, w0)),
                  __dsub_rn(E, w0)),
        B_surf);
}

__device__ __forceinline__ double tiny_to_abs(double f) { return fabs(f) < 1e-100 ? fabs(f) : f; }

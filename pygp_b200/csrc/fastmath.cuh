// fastmath.cuh -- short FP64 elementary functions for the covariance epilogues.
//
// Why: the Gram / trace tile kernels are bound by FP64 instruction issue (64 FP64
// lanes per SM; a warp-wide DFMA occupies the pipe for two cycles), and ncu's source
// view of round 1's gram_kernel<Matern5> put 79 % of the executed instructions in
// the inlined library sqrt / exp (~130 per entry, 29 of them FP64, plus ~24
// constant-materialising moves per exp).  The versions here cost
//     sqrt_pos   6 FP64 (hardware 20-bit seed + one third-order step)
//     exp_tab   10 FP64 + 1 shared load + ~6 integer instructions
// and agree with the correctly rounded results to <= 2 ulp (tests/test_fastmath_gpu.py
// sweeps them against libm through pgp_dev_fastmath), far inside the 1e-10 parity
// tolerance of pygp's kernels (se.py:53-66, matern.py:44-90).
#pragma once

#include <cstdint>

namespace pgp {
namespace fm {

// 2^(j/64), j = 0..63, correctly rounded
__device__ const double kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

constexpr int kExpTabDoubles = 64;

// copy the table into shared memory (call before a __syncthreads)
__device__ __forceinline__ void load_exp_tab(double* s_tab, int tid, int nthreads) {
    for (int i = tid; i < kExpTabDoubles; i += nthreads) s_tab[i] = kExp2Tab[i];
}

// Constants that do not fit a 32-bit immediate (FP64 operands take an immediate for the
// HIGH word only) live in the constant bank, where DFMA reads them as an operand: as C++
// literals ptxas rebuilt each of them with two moves before every use (~28 moves per entry).
__constant__ double kFmC[8] = {
    0x1.71547652b82fep+6,      // [0] 64 / ln 2
    0x1.fdf473de6af28p-28,     // [1] ln 2 / 64 - 0x1.62e42p-7
    1.0 / 120, 1.0 / 24, 1.0 / 6,   // [2..4]
    1.0 / 3,                   // [5]
    0.0, 0.0};

// sqrt(d) for d >= 0, exactly 0 at d = 0: hardware seed y ~ d^-1/2 (2^-20), then
// sqrt = t (1 + e/2 + 3 e^2/8), t = d y, e = 1 - t y.
__device__ __forceinline__ double sqrt_pos(double d) {
    const double dp = d + 0x1p-996;        // == d for d > 2^-943; keeps the seed finite at d = 0 (t = 0 y = 0)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(dp));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(t * e, p, t);
}

// sqrt(d) and 1/sqrt(d) together (the Matern-1/2 gradient divides by r)
__device__ __forceinline__ double sqrt_rsqrt_pos(double d, double& rinv) {
    const double dp = d + 0x1p-996;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(dp));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    rinv = fma(y * e, p, y);
    return fma(t * e, p, t);
}

// exp(x): x = (64 k + j) ln2/64 + q, |q| <= ln2/128;  exp = 2^k * T[j] * (1 + q s(q)),
// s = 1 + q/2 + q^2/6 + q^3/24 + q^4/120 (truncation 3.5e-17 relative).
// `bad` is OR-ed with 1 when the result is not a normal double (x < -708, x > 709): the
// caller recomputes such entries with the library exp in ONE cold block (inlining the
// library call at every entry bloated the kernel by a third).
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ s_tab, int& bad) {
    constexpr double kMagic = 6755399441055744.0;                // 1.5 * 2^52 (high word only)
    constexpr double kHi = 0x1.62e42p-7;                         // ln 2 / 64, 21 significant bits (high word only)
    const double t = fma(x, kFmC[0], kMagic);
    const int n = __double2loint(t);
    const double nd = t - kMagic;
    double q = fma(nd, -kHi, x);
    q = fma(nd, -kFmC[1], q);
    const double T = s_tab[n & 63];
    double s = fma(q, kFmC[2], kFmC[3]);
    s = fma(s, q, kFmC[4]);
    s = fma(s, q, 0.5);
    s = fma(s, q, 1.0);
    const double r = fma(T * q, s, T);
    bad |= (unsigned)(n + 64 * 1020) > (unsigned)(2 * 64 * 1020);
    return __hiloint2double(__double2hiint(r) + ((n >> 6) << 20), __double2loint(r));
}

// exp_tab for the gradient / trace kernels, where an entry below the normal range only ever
// enters a sum or a derivative matrix: arguments under -708 are evaluated AT -708 (absolute
// error < 4e-308), NaN propagates, arguments above 709 (sf > 1e150) are not supported there.
__device__ __forceinline__ double exp_tab_clamped(double x, const double* __restrict__ s_tab) {
    int bad = 0;
    return exp_tab(x < -708.0 ? -708.0 : x, s_tab, bad);
}

}  // namespace fm
}  // namespace pgp

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

// ---- log(x) for x >= 1, ABSOLUTE accuracy ~2e-16 ------------------------------------------------
// The rational-quadratic kernel needs E^-alpha = exp(-alpha log E) with E = 1 + D / (2 alpha) >= 1
// (pygp/kernels/rq.py:56-63): only the absolute error of log E reaches the result (times alpha), so the
// table method needs no care near x = 1.  x = 2^e m, m in [1, 2); j = top 7 mantissa bits; c_j ~ 1 / (centre
// of interval j); r = m c_j - 1 (one fma, |r| <= 2^-8); log x = e ln2 + T_j + log1p(r), T_j = -log c_j.
// Cost: 13 FP64 + one 16-byte shared load + ~8 integer instructions (library pow: ~150 instructions).
constexpr int kLogTabDoubles = 256;
__device__ const double kLogTab[kLogTabDoubles] = {
    0x1.fe01fe0000000p-1, 0x1.ff00ac2b10bc0p-9, 0x1.fa11ca0000000p-1, 0x1.7dc49e7810addp-7,
    0x1.f6310a0000000p-1, 0x1.3cea5df46a5c8p-6, 0x1.f25f640000000p-1, 0x1.b9fc0afaf91a1p-6,
    0x1.ee9c800000000p-1, 0x1.1b0d90923d990p-5, 0x1.eae8080000000p-1, 0x1.58a5b57c8e4dcp-5,
    0x1.e741aa0000000p-1, 0x1.95c836cc8e3f4p-5, 0x1.e3a9180000000p-1, 0x1.d276b22db0b5dp-5,
    0x1.e01e020000000p-1, 0x1.075982498e472p-4, 0x1.dca01e0000000p-1, 0x1.253f6120a1419p-4,
    0x1.d92f220000000p-1, 0x1.42edcd9a646f2p-4, 0x1.d5cac80000000p-1, 0x1.60658ad3750c4p-4,
    0x1.d272ca0000000p-1, 0x1.7da76907b12cfp-4, 0x1.cf26e60000000p-1, 0x1.9ab42252033afp-4,
    0x1.cbe6da0000000p-1, 0x1.b78c7d2b0edb1p-4, 0x1.c8b2660000000p-1, 0x1.d4313a96cb361p-4,
    0x1.c5894e0000000p-1, 0x1.f0a30391162cap-4, 0x1.c26b540000000p-1, 0x1.06714f3ca5972p-3,
    0x1.bf583e0000000p-1, 0x1.14785c6e742bep-3, 0x1.bc4fd60000000p-1, 0x1.2266f328a5acep-3,
    0x1.b951e20000000p-1, 0x1.303d74c647fddp-3, 0x1.b65e2e0000000p-1, 0x1.3dfc2c26cc62bp-3,
    0x1.b374840000000p-1, 0x1.4ba37269a55f0p-3, 0x1.b094b40000000p-1, 0x1.5933896982097p-3,
    0x1.adbe880000000p-1, 0x1.66acd4072ad51p-3, 0x1.aaf1d20000000p-1, 0x1.740f93fc037bap-3,
    0x1.a82e660000000p-1, 0x1.815c059c357ffp-3, 0x1.a574100000000p-1, 0x1.8e92902886d46p-3,
    0x1.a2c2a80000000p-1, 0x1.9bb36547dfb89p-3, 0x1.a01a020000000p-1, 0x1.a8becdf082f1cp-3,
    0x1.9d79f20000000p-1, 0x1.b5b51740fb5abp-3, 0x1.9ae24e0000000p-1, 0x1.c2968890c18cbp-3,
    0x1.9852f00000000p-1, 0x1.cf6359209c5eep-3, 0x1.95cbb00000000p-1, 0x1.dc1bcdcabec8bp-3,
    0x1.934c680000000p-1, 0x1.e8c0250aa5a60p-3, 0x1.90d4f20000000p-1, 0x1.f550a0ecb7b4bp-3,
    0x1.8e65280000000p-1, 0x1.00e6c38ad501ep-2, 0x1.8bfce80000000p-1, 0x1.071b860cd590dp-2,
    0x1.899c100000000p-1, 0x1.0d46b3d9ab750p-2, 0x1.87427c0000000p-1, 0x1.13686fa13a8b1p-2,
    0x1.84f00c0000000p-1, 0x1.1980d34542370p-2, 0x1.82a4a00000000p-1, 0x1.1f8ffa248a2f3p-2,
    0x1.8060180000000p-1, 0x1.2596011df763ap-2, 0x1.7e22560000000p-1, 0x1.2b93013789d31p-2,
    0x1.7beb3a0000000p-1, 0x1.31871a4144190p-2, 0x1.79baa60000000p-1, 0x1.37726827fd863p-2,
    0x1.7790820000000p-1, 0x1.3d54f7e81f71cp-2, 0x1.756cac0000000p-1, 0x1.432ef2f84e814p-2,
    0x1.734f0c0000000p-1, 0x1.490068ec009d2p-2, 0x1.7137860000000p-1, 0x1.4ec9758200275p-2,
    0x1.6f26020000000p-1, 0x1.548a2aa6dd268p-2, 0x1.6d1a620000000p-1, 0x1.5a42ac334cfe4p-2,
    0x1.6b14900000000p-1, 0x1.5ff308ea793dbp-2, 0x1.6914740000000p-1, 0x1.659b56383e1f4p-2,
    0x1.6719f40000000p-1, 0x1.6b3bb05b59444p-2, 0x1.6524f80000000p-1, 0x1.70d42f1789238p-2,
    0x1.63356c0000000p-1, 0x1.7664dfcb9dbd2p-2, 0x1.614b360000000p-1, 0x1.7bede21f7afc4p-2,
    0x1.5f66440000000p-1, 0x1.816f3fb20d49fp-2, 0x1.5d867c0000000p-1, 0x1.86e91a5b30ba1p-2,
    0x1.5babcc0000000p-1, 0x1.8c5b7dad8b48dp-2, 0x1.59d6200000000p-1, 0x1.91c67bf45a84dp-2,
    0x1.5805600000000p-1, 0x1.972a345135159p-2, 0x1.56397c0000000p-1, 0x1.9c86af25c0865p-2,
    0x1.54725e0000000p-1, 0x1.a1dc07915b999p-2, 0x1.52aff60000000p-1, 0x1.a72a47a2bd9f0p-2,
    0x1.50f22e0000000p-1, 0x1.ac718c598b0e4p-2, 0x1.4f38f60000000p-1, 0x1.b1b1e177dfc5cp-2,
    0x1.4d843c0000000p-1, 0x1.b6eb599bcf35ep-2, 0x1.4bd3ee0000000p-1, 0x1.bc1e083cdad0bp-2,
    0x1.4a27fa0000000p-1, 0x1.c14a01ad5f034p-2, 0x1.4880520000000p-1, 0x1.c66f4ea3f6ff8p-2,
    0x1.46dce40000000p-1, 0x1.cb8e04fcd7ad4p-2, 0x1.453d9e0000000p-1, 0x1.d0a63b7321e65p-2,
    0x1.43a2740000000p-1, 0x1.d5b7f6a62c696p-2, 0x1.420b520000000p-1, 0x1.dac35526c5957p-2,
    0x1.40782e0000000p-1, 0x1.dfc856946d5c7p-2, 0x1.3ee8f40000000p-1, 0x1.e4c71b0e87705p-2,
    0x1.3d5d9a0000000p-1, 0x1.e9bfa37586206p-2, 0x1.3bd60e0000000p-1, 0x1.eeb20b000ddf8p-2,
    0x1.3a52440000000p-1, 0x1.f39e5a4011e60p-2, 0x1.38d22e0000000p-1, 0x1.f884a0dbe9ecfp-2,
    0x1.3755be0000000p-1, 0x1.fd64ef2361583p-2, 0x1.35dce60000000p-1, 0x1.011fab085ff8ap-1,
    0x1.34679a0000000p-1, 0x1.0389f052e6342p-1, 0x1.32f5ce0000000p-1, 0x1.05f14d38645a4p-1,
    0x1.3187760000000p-1, 0x1.0855c7c6b4511p-1, 0x1.301c820000000p-1, 0x1.0ab76d0ee14d7p-1,
    0x1.2eb4ea0000000p-1, 0x1.0d163d019d6b8p-1, 0x1.2d50a00000000p-1, 0x1.0f7241e9b497dp-1,
    0x1.2bef980000000p-1, 0x1.11cb83007cd02p-1, 0x1.2a91ca0000000p-1, 0x1.142200ec43d4dp-1,
    0x1.2937260000000p-1, 0x1.1675ca44ba60fp-1, 0x1.27dfa40000000p-1, 0x1.18c6e0335cf09p-1,
    0x1.268b380000000p-1, 0x1.1b154affda29fp-1, 0x1.2539d80000000p-1, 0x1.1d610fbe77003p-1,
    0x1.23eb7a0000000p-1, 0x1.1faa33be70950p-1, 0x1.22a0120000000p-1, 0x1.21f0c0105beecp-1,
    0x1.2157980000000p-1, 0x1.2434b6fc83934p-1, 0x1.2012020000000p-1, 0x1.26761e85430e9p-1,
    0x1.1ecf440000000p-1, 0x1.28b5007b60783p-1, 0x1.1d8f560000000p-1, 0x1.2af15fd0640b0p-1,
    0x1.1c52300000000p-1, 0x1.2d2b3fa2edc9ep-1, 0x1.1b17c60000000p-1, 0x1.2f62aa7b09549p-1,
    0x1.19e0120000000p-1, 0x1.3197a0487fe6cp-1, 0x1.18ab080000000p-1, 0x1.33ca2c0b28995p-1,
    0x1.1778a20000000p-1, 0x1.35fa4e1336ea2p-1, 0x1.1648d60000000p-1, 0x1.38280e2b8798bp-1,
    0x1.151b9a0000000p-1, 0x1.3a53745debdfap-1, 0x1.13f0e80000000p-1, 0x1.3c7c81877320fp-1,
    0x1.12c8b80000000p-1, 0x1.3ea33a5eb2f61p-1, 0x1.11a3020000000p-1, 0x1.40c7a3ca0dcebp-1,
    0x1.107fbc0000000p-1, 0x1.42e9c6a1f80bfp-1, 0x1.0f5ee00000000p-1, 0x1.4509a4733bb0cp-1,
    0x1.0e40660000000p-1, 0x1.472742b53aab3p-1, 0x1.0d24460000000p-1, 0x1.4942a7102fc0dp-1,
    0x1.0c0a780000000p-1, 0x1.4b5bd75d6e276p-1, 0x1.0af2f80000000p-1, 0x1.4d72d1fb9fd0bp-1,
    0x1.09ddba0000000p-1, 0x1.4f87a4c3026ebp-1, 0x1.08cabc0000000p-1, 0x1.519a4a87a3450p-1,
    0x1.07b9f20000000p-1, 0x1.53aad18999b82p-1, 0x1.06ab5a0000000p-1, 0x1.55b934dd40bcep-1,
    0x1.059eea0000000p-1, 0x1.57c57f416f191p-1, 0x1.04949c0000000p-1, 0x1.59cfb3dbae887p-1,
    0x1.038c6c0000000p-1, 0x1.5bd7d20271c77p-1, 0x1.0286500000000p-1, 0x1.5ddde50149924p-1,
    0x1.0182440000000p-1, 0x1.5fe1ec791891ep-1, 0x1.0080400000000p-1, 0x1.61e3f01a46467p-1};

__device__ __forceinline__ void load_log_tab(double* s_tab, int tid, int nthreads) {
    for (int i = tid; i < kLogTabDoubles; i += nthreads) s_tab[i] = kLogTab[i];
}

// [0..4] log1p: -1/2 1/3 -1/4 1/5 -1/6 1/7 ... ; sin / cos kernels on [-pi/4, pi/4] (fdlibm k_sin.c / k_cos.c);
// Cody-Waite pi/2 in three parts (33 + 33 + 53 bits: k pio2_1 exact for k < 2^20)
__constant__ double kFmL[8] = {-0.5, 1.0 / 3, -0.25, 0.2, -1.0 / 6, 1.0 / 7,
                               0x1.a39ef35793c76p-33 /* ln2 - 0x1.62e42feep-1 */, 0.0};
__constant__ double kFmS[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                               2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
__constant__ double kFmCo[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
__constant__ double kFmP[4] = {6.36619772367581382433e-01, 1.57079632673412561417e+00, 6.07710050630396597660e-11,
                               2.02226624871116645580e-21};

__device__ __forceinline__ double log_ge1_tab(double x, const double* __restrict__ s_logtab, int& bad) {
    const int hi = __double2hiint(x);
    bad |= (unsigned)hi >= 0x7ff00000u;                 // inf / NaN / negative: the caller's cold path
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> 13) & 127;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double2 ct = *reinterpret_cast<const double2*>(s_logtab + 2 * j);
    const double r = fma(m, ct.x, -1.0);
    double q = fma(r, kFmL[5], kFmL[4]);
    q = fma(q, r, kFmL[3]);
    q = fma(q, r, kFmL[2]);
    q = fma(q, r, kFmL[1]);
    q = fma(q, r, kFmL[0]);
    const double l1p = fma(r * r, q, r);
    const double ed = __hiloint2double(0x43300000, e) - 4503599627370496.0;     // (double)e for e >= 0
    return fma(ed, 0x1.62e42feep-1, ct.y) + fma(ed, kFmL[6], l1p);
}

// +-sin(x) for |x| < 2^20 pi / 2 (else `bad`): x = k pi/2 + f, |f| <= pi/4; |sin x| = |sin f| (k even) or
// |cos f| (k odd).  The sign is NOT resolved: the periodic kernel only ever squares the sine
// (pygp/kernels/periodic.py:53-59).
__device__ __forceinline__ double sin_unsigned_cw(double x, int& bad) {
    constexpr double kMagic = 6755399441055744.0;
    const double t = fma(x, kFmP[0], kMagic);
    const int k = __double2loint(t);
    const double kd = t - kMagic;
    double f = fma(kd, -kFmP[1], x);
    f = fma(kd, -kFmP[2], f);
    f = fma(kd, -kFmP[3], f);
    bad |= !(fabs(x) < 1.6e6);
    const double z = f * f;
    double ps = fma(z, kFmS[5], kFmS[4]);
    double pc = fma(z, kFmCo[5], kFmCo[4]);
    ps = fma(ps, z, kFmS[3]);   pc = fma(pc, z, kFmCo[3]);
    ps = fma(ps, z, kFmS[2]);   pc = fma(pc, z, kFmCo[2]);
    ps = fma(ps, z, kFmS[1]);   pc = fma(pc, z, kFmCo[1]);
    ps = fma(ps, z, kFmS[0]);   pc = fma(pc, z, kFmCo[0]);
    const double sn = fma(f * z, ps, f);                         // sin f
    const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));        // cos f
    return (k & 1) ? cs : sn;
}

}  // namespace fm
}  // namespace pgp

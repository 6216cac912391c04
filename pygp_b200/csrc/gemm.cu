// gemm.cu -- FP64 tensor-core GEMM for sm_100a:  C = beta C + alpha op(A) op(B)^T
// (NT form for the factorisation; TN / NN forms for FITC, see gemm.cuh).
//
// Why mma.sync and not tcgen05: tcgen05.mma has no f64 kind; on sm_100a every
// FP64 mma shape lowers to DMMA.8x8x4 (checked with cuobjdump), so the kernel is
// written against m8n8k4 directly.  Bound: FP64 tensor pipe.
//
// Tiling: CTA tile 128 x 128 x 32, 16 warps as 4 (M) x 4 (N), warp tile 32 x 32
// = 4 x 4 DMMA tiles = 32 accumulator doubles per lane (the 8-warp 64 x 32
// layouts stay selectable through PGP_GEMM_VARIANT for tuning; measured within
// 3 % of each other, profiles/r01_gemm_variants.txt).  Operands are staged by a
// 3-deep cp.async ring (16-byte chunks, zero-fill predication at the M/N/K
// edges, so no padding of the matrices is needed).  Shared rows are padded to
// BK + 4 doubles: the DMMA fragment loads (lane -> row lane/4, k lane%4) of a
// half-warp then fall in 16 distinct 8-byte banks (ncu: 0 bank conflicts).
// Fragments are double-buffered in registers across the 4-wide k steps.
//
// Structure flags let the same kernel serve every O(N^3) step of the factor,
// the triangular inverse and V V^T without touching zero blocks (gemm.cuh).

#include "gemm.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace pgp {

namespace {

// CTAs are rasterised in groups of GROUP_M tile rows (column index fastest
// inside a group) so that the ~148 CTAs in flight share a ~12 x 12 block of
// tiles: each operand k-slice is then fetched from HBM once per wave and
// served to the other CTAs from L2.
constexpr int GROUP_M = 12;

// T = CTA tile edge (T x T outputs): 128 for the big updates, 64 for launches with
// too few 128-tiles to fill the GPU (leaf levels of the recursions, batched small
// blocks), where a 128 x 128 tile computes mostly padding on a fraction of the SMs.
template <int T_, int BK_, int STAGES_>
struct Cfg {
    static constexpr int BM = T_, BN = T_;
    static constexpr int BK = BK_, STAGES = STAGES_;
    static constexpr int LDS = BK + 4;  // padded shared row (doubles): conflict-free fragment loads
    static constexpr int LDT = BM + 4;  // row of a transposed-operand tile ([k][m]); 132 = 4 mod 16
    static constexpr int TILE_DOUBLES = BM * LDS > BK * LDT ? BM * LDS : BK * LDT;
    static constexpr int STAGE_DOUBLES = 2 * TILE_DOUBLES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);
};

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace

// WM x WN warps; each warp owns a (BM/WM) x (BN/WN) tile = MI x NJ DMMA tiles.
// TA / TB: the operand is stored with the contraction index as its ROW index
// (A is (K, M) row-major / B is (K, N) row-major); its tile is then staged as
// [k][m] with a row of BM + 4 doubles, which keeps the fragment loads
// (lane -> m = lane/4, k = lane%4) conflict-free as well: 132 = 4 (mod 16).
// STAIR: staircase operands (gemm.cuh) -- a separate instantiation, so that the kernels of the factorisation
// keep their register allocation (the 16-warp variant sits at its 128-register cap).
template <int WM, int WN, class C, bool TA, bool TB, bool STAIR = false>
__global__ void __launch_bounds__(WM * WN * 32, C::BM == 64 ? 2 : 1) gemm_kernel(GemmArgs a, int tm, int tn) {
    constexpr int BM = C::BM, BN = C::BN;
    constexpr int THREADS = WM * WN * 32;
    constexpr int MI = BM / WM / 8, NJ = BN / WN / 8;
    constexpr int BK = C::BK, STAGES = C::STAGES, LDS = C::LDS, LDT = C::LDT;
    constexpr int TILE_DOUBLES = C::TILE_DOUBLES, STAGE_DOUBLES = C::STAGE_DOUBLES;
    constexpr int CPR = BK / 2;                      // 16-byte chunks per operand row (k contiguous)
    constexpr int CPT = BM / 2;                      // ... per row of a transposed tile (m contiguous)
    constexpr int NCH = (BM * CPR) / THREADS;        // chunks per thread per operand tile
    static_assert(BM == BN, "square CTA tile assumed by the copy descriptors");
    extern __shared__ __align__(16) double smem[];

    // grouped rasterisation of the linear CTA index
    int tile_m, tile_n;
    {
        const int pid = blockIdx.x;
        const int per_group = GROUP_M * tn;
        const int group = pid / per_group;
        const int first_m = group * GROUP_M;
        const int gsz = min(tm - first_m, GROUP_M);
        const int local = pid - group * per_group;
        tile_m = first_m + local % gsz;
        tile_n = local / gsz;
    }
    const int64_t m0 = (int64_t)tile_m * BM;
    const int64_t n0 = (int64_t)tile_n * BN;
    if (STAIR && a.ragged_mstep) {          // ragged batch: this member's row count
        a.M -= (int64_t)blockIdx.z * a.ragged_mstep;
        if (m0 >= a.M) return;
    }
    if (a.tri && n0 > m0 + BM - 1 + a.tri_off) return;
    // staircase operands (gemm.cuh): first column at which this tile row's block exists, relative to the
    // operand's column 0 (32-bit arithmetic: the prologue must not cost the main loop registers)
    int stair_rel = 0;
    if (STAIR && a.stair && tile_m * BM >= (int)a.stair_front)
        stair_rel = (int)(a.stair_first - a.stair_off) + (tile_m * BM - (int)a.stair_front) / (int)a.stair_nb * (int)a.stair_step;
    if (STAIR && (a.stair & 2) && tile_n * BN + BN - 1 < stair_rel) return;

    const int b = blockIdx.z;
    const double* __restrict__ A = a.A + (int64_t)b * a.strideA;
    const double* __restrict__ B = a.B + (int64_t)b * a.strideB;
    double* __restrict__ Cm = a.C + (int64_t)b * a.strideC;

    // contraction range of this CTA: [ks, ke)
    int64_t ks = 0, ke = a.K;
    if (a.krow) {
        ks = m0 + a.krow_off;
        if (ks < 0) ks = 0;
        ks = ks / BK * BK;
        if (ks > a.K) ks = a.K;
    }
    if (STAIR && (a.stair & 1) && stair_rel > 0) {
        ks = stair_rel / BK * BK;
        if (ks > a.K) ks = a.K;
    }
    if (a.kcol) {  // B is upper triangular in (k, col): the contraction may stop at the tile's last column
        int64_t kend = n0 + BN + a.kcol_off;
        if (kend < ke) ke = kend;
        if (ke < ks) ke = ks;
    }
    if (a.splitk > 1) {  // blockIdx.y = slice of the contraction; partial result to the workspace
        const int64_t per = (a.K + a.splitk - 1) / a.splitk;
        const int64_t kc = (per + BK - 1) / BK * BK;
        ks = (int64_t)blockIdx.y * kc;
        ke = min(a.K, ks + kc);
        if (ks > ke) ks = ke;
        Cm = a.ws + ((int64_t)b * a.splitk + blockIdx.y) * a.M * a.ldws;
    }
    const int KT = (int)((ke - ks + BK - 1) / BK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;

    // per-thread copy descriptors, fixed for the whole k loop.  Chunk q of a thread
    // is chunk (tid + q THREADS) of the tile, so consecutive q differ by a constant
    // row (normal: THREADS / CPR rows) or k (transposed: THREADS / CPT) offset and
    // one base pointer + one shared-memory offset per operand is all that stays
    // live in registers across the loop.
    //   normal operand:     chunk c -> row c / CPR, k offset 2 (c % CPR)
    //   transposed operand: chunk c -> k row c / CPT, m offset 2 (c % CPT)
    constexpr int RQ = THREADS / CPR;   // rows between consecutive chunks of a thread (normal)
    constexpr int KQ = THREADS / CPT;   // k rows between them (transposed)
    static_assert(THREADS % CPR == 0 && THREADS % CPT == 0, "chunk stride must be whole rows");
    const double* srcA0;
    const double* srcB0;
    int soffA0, soffB0, kA0, kB0, okA, okB;   // ok*: normal -> bit q = row valid; transposed -> bytes valid along m/n
    {
        const int c = threadIdx.x;
        if (!TA) {
            const int row = c / CPR;
            kA0 = (c % CPR) * 2;
            soffA0 = row * LDS + kA0;
            okA = 0;
#pragma unroll
            for (int q = 0; q < NCH; ++q) okA |= (m0 + row + q * RQ < a.M) << q;
            srcA0 = A + (m0 + row) * a.lda + ks + kA0;
        } else {
            const int kr = c / CPT, mo = (c % CPT) * 2;
            kA0 = kr;
            soffA0 = kr * LDT + mo;
            const int64_t rem = a.M - (m0 + mo);
            okA = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
            srcA0 = A + (ks + kr) * a.lda + (okA ? m0 + mo : 0);
        }
        if (!TB) {
            const int row = c / CPR;
            kB0 = (c % CPR) * 2;
            soffB0 = row * LDS + kB0;
            okB = 0;
#pragma unroll
            for (int q = 0; q < NCH; ++q) okB |= (n0 + row + q * RQ < a.N) << q;
            srcB0 = B + (n0 + row) * a.ldb + ks + kB0;
        } else {
            const int kr = c / CPT, no = (c % CPT) * 2;
            kB0 = kr;
            soffB0 = kr * LDT + no;
            const int64_t rem = a.N - (n0 + no);
            okB = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
            srcB0 = B + (ks + kr) * a.ldb + (okB ? n0 + no : 0);
        }
    }
    // one 16-byte copy of k-tile `ktile` into ring slot `slot`: c < NCH -> chunk c
    // of the A tile, else chunk c - NCH of the B tile (c is a compile-time constant
    // at every call site)
    auto load_chunk = [&](int slot, int ktile, int c) {
        double* As = smem + slot * STAGE_DOUBLES;
        double* Bs = As + TILE_DOUBLES;
        const int64_t koff = (int64_t)ktile * BK;
        const int kleft = (int)min((int64_t)BK, ke - ks - koff);  // valid k in this tile
        if (c < NCH) {
            const int q = c;
            if (!TA) {
                const int rem = kleft - kA0;
                const int bytes = ((okA >> q) & 1) ? (rem >= 2 ? 16 : (rem == 1 ? 8 : 0)) : 0;
                cp_async16(As + soffA0 + q * RQ * LDS, bytes ? srcA0 + (int64_t)q * RQ * a.lda + koff : A, bytes);
            } else {
                const int bytes = kA0 + q * KQ < kleft ? okA : 0;
                cp_async16(As + soffA0 + q * KQ * LDT, bytes ? srcA0 + (koff + q * KQ) * a.lda : A, bytes);
            }
        } else {
            const int q = c - NCH;
            if (!TB) {
                const int rem = kleft - kB0;
                const int bytes = ((okB >> q) & 1) ? (rem >= 2 ? 16 : (rem == 1 ? 8 : 0)) : 0;
                cp_async16(Bs + soffB0 + q * RQ * LDS, bytes ? srcB0 + (int64_t)q * RQ * a.ldb + koff : B, bytes);
            } else {
                const int bytes = kB0 + q * KQ < kleft ? okB : 0;
                cp_async16(Bs + soffB0 + q * KQ * LDT, bytes ? srcB0 + (koff + q * KQ) * a.ldb : B, bytes);
            }
        }
    };
    auto load_stage = [&](int slot, int ktile) {
#pragma unroll
        for (int c = 0; c < 2 * NCH; ++c) load_chunk(slot, ktile, c);
    };

    // fast path for interior tiles (every row / column of the CTA tile valid and the
    // k-tile complete): no predicates, one pointer bump per k-tile, immediates per
    // chunk.  The general path above costs ~40 scalar instructions per copy, which
    // the 16 warps execute between their DMMAs.
    const bool full_mn = (m0 + BM <= a.M) && (n0 + BN <= a.N);
    const unsigned sdA0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)soffA0 * 8u;
    const unsigned sdB0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)(TILE_DOUBLES + soffB0) * 8u;
    const int64_t stepA = TA ? (int64_t)BK * a.lda : BK, stepB = TB ? (int64_t)BK * a.ldb : BK;
    const int64_t qA = TA ? (int64_t)KQ * a.lda : (int64_t)RQ * a.lda;   // source stride between chunks
    const int64_t qB = TB ? (int64_t)KQ * a.ldb : (int64_t)RQ * a.ldb;
    constexpr unsigned QSA = (TA ? KQ * LDT : RQ * LDS) * 8u;            // destination stride (bytes)
    constexpr unsigned QSB = (TB ? KQ * LDT : RQ * LDS) * 8u;
    const double* gA = srcA0 + (int64_t)(STAGES - 1) * stepA;             // k-tile kt + STAGES - 1
    const double* gB = srcB0 + (int64_t)(STAGES - 1) * stepB;
    auto load_fast = [&](int slot, int c) {
        const unsigned so = (unsigned)slot * (unsigned)(STAGE_DOUBLES * 8);
        if (c < NCH) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdA0 + so + (unsigned)c * QSA),
                         "l"(gA + (int64_t)c * qA));
        } else {
            const int q = c - NCH;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdB0 + so + (unsigned)q * QSB),
                         "l"(gB + (int64_t)q * qB));
        }
    };

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    // fragment addressing: element (row r of the warp tile, k) lives at
    //   normal: r * LDS + k          transposed: k * LDT + r
    constexpr int A_RS = TA ? 1 : LDS, A_KS = TA ? LDT : 1;
    constexpr int B_RS = TB ? 1 : LDS, B_KS = TB ? LDT : 1;

    // The copies of k-tile kt + STAGES - 1 are issued INTERLEAVED with the DMMAs of
    // k-tile kt (CPS per 4-wide k-step) instead of in one burst after the barrier:
    // a burst of 2 NCH LDGSTS per thread from all 16 warps backs up the LSU and
    // starves the tensor pipe for several hundred cycles per k-tile
    // (tools/dmma_probe.cu: the same loop without the copies runs at 35.9 TFLOP/s).
    constexpr int KSTEPS = BK / 4;
    constexpr int CPS = (2 * NCH + KSTEPS - 1) / KSTEPS;
    auto main_loop = [&](auto fast_tag, int kt_begin, int kt_end) {
        constexpr bool FAST = decltype(fast_tag)::value;
        for (int kt = kt_begin; kt < kt_end; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int nk = kt + STAGES - 1;
            const bool do_load = FAST || nk < KT;
            const int nslot = nk % STAGES;
            const double* As = smem + (kt % STAGES) * STAGE_DOUBLES + (wm * (BM / WM) + g) * A_RS + t * A_KS;
            const double* Bs = smem + (kt % STAGES) * STAGE_DOUBLES + TILE_DOUBLES + (wn * (BN / WN) + g) * B_RS + t * B_KS;
            // fragments double-buffered in registers: the loads of step kk+1 are in
            // flight while the DMMAs of step kk issue
            double af[2][MI], bf[2][NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[0][i] = As[i * 8 * A_RS];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[0][j] = Bs[j * 8 * B_RS];
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < KSTEPS) {
#pragma unroll
                    for (int i = 0; i < MI; ++i) af[nxt][i] = As[i * 8 * A_RS + (kk + 1) * 4 * A_KS];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) bf[nxt][j] = Bs[j * 8 * B_RS + (kk + 1) * 4 * B_KS];
                }
                if (FAST) {
#pragma unroll
                    for (int c = kk * CPS; c < (kk + 1) * CPS && c < 2 * NCH; ++c) load_fast(nslot, c);
                } else if (do_load) {
#pragma unroll
                    for (int c = kk * CPS; c < (kk + 1) * CPS && c < 2 * NCH; ++c) load_chunk(nslot, nk, c);
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
            }
            cp_async_commit();
            if (FAST) {
                gA += stepA;
                gB += stepB;
            }
        }
    };
    // iterations whose prefetched k-tile (kt + STAGES - 1) is complete run the
    // predicate-free loop when the CTA tile is interior; the rest (k tail, edge
    // tiles) run the general one
    int kt_fast = 0;
    if (full_mn) {
        int64_t nfull = (ke - ks) / BK - (STAGES - 1);
        kt_fast = (int)(nfull > 0 ? nfull : 0);
        if (kt_fast > KT) kt_fast = KT;
    }
    main_loop(std::true_type{}, 0, kt_fast);
    main_loop(std::false_type{}, kt_fast, KT);
    cp_async_wait<0>();

    // epilogue: lane owns C[row = 8i + g][col = 8j + 2t, +1] of its warp tile
    const bool part = a.splitk > 1;
    const double alpha = part ? 1.0 : a.alpha, beta = part ? 0.0 : a.beta;
    const int64_t ldc = part ? a.ldws : a.ldc;
    const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cm) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int64_t row = m0 + wm * (BM / WM) + i * 8 + g;
        if (row >= a.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int64_t col = n0 + wn * (BN / WN) + j * 8 + 2 * t;
            if (col >= a.N) continue;
            double* p = Cm + row * ldc + col;
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
            if (vec_ok && col + 1 < a.N) {
                if (beta != 0.0) {
                    double2 old = *reinterpret_cast<const double2*>(p);
                    v0 += beta * old.x;
                    v1 += beta * old.y;
                }
                *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
            } else {
                if (beta != 0.0) v0 += beta * p[0];
                p[0] = v0;
                if (col + 1 < a.N) {
                    if (beta != 0.0) v1 += beta * p[1];
                    p[1] = v1;
                }
            }
        }
    }
}

// C = beta C + alpha sum_s ws[s]  (fixed summation order: deterministic)
__global__ void splitk_reduce_kernel(GemmArgs a) {
    const int b = blockIdx.z;
    const int64_t total = a.M * a.N;
    double* Cm = a.C + (int64_t)b * a.strideC;
    const double* ws = a.ws + (int64_t)b * a.splitk * a.M * a.ldws;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx / a.N, c = idx - r * a.N;
        if (a.tri && c > r + a.tri_off) continue;
        double s = 0.0;
        for (int k = 0; k < a.splitk; ++k) s += ws[((int64_t)k * a.M + r) * a.ldws + c];
        double* p = Cm + r * a.ldc + c;
        *p = a.beta != 0.0 ? a.beta * *p + a.alpha * s : a.alpha * s;
    }
}

namespace {
template <int WM, int WN, class C, bool TA, bool TB>
int launch_variant(pgp_ctx* ctx, const GemmArgs& a, int64_t tm, int64_t tn) {
    dim3 grid((unsigned)(tm * tn), (unsigned)std::max(a.splitk, 1), a.batch);
    if (a.stair || a.ragged_mstep) {
        if (TA) return ctx->fail(PGP_E_ARG, "gemm: staircase operands need A stored (M, K)");
        auto kern = gemm_kernel<WM, WN, C, false, TB, true>;
        PGP_TRY(ensure_dyn_smem(ctx, kern, C::SMEM));
        kern<<<grid, WM * WN * 32, C::SMEM, ctx->stream>>>(a, (int)tm, (int)tn);
        return 0;
    }
    auto kern = gemm_kernel<WM, WN, C, TA, TB>;
    PGP_TRY(ensure_dyn_smem(ctx, kern, C::SMEM));
    kern<<<grid, WM * WN * 32, C::SMEM, ctx->stream>>>(a, (int)tm, (int)tn);
    return 0;
}
}  // namespace

int launch_gemm(pgp_ctx* ctx, const GemmArgs& a_in) {
    GemmArgs a = a_in;
    if (a.M <= 0 || a.N <= 0) return 0;
    if ((a.lda & 1) || (a.ldb & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
        (reinterpret_cast<uintptr_t>(a.B) & 15) || ((a.strideA | a.strideB) & 1))
        return ctx->fail(PGP_E_ARG, "gemm: A and B must be 16-byte aligned with even leading dimensions");
    if (a.transA && a.krow) return ctx->fail(PGP_E_ARG, "gemm: krow needs A stored (M, K)");
    if (a.kcol && !a.transB) return ctx->fail(PGP_E_ARG, "gemm: kcol needs B stored (K, N)");
    static const int variant = [] { const char* e = getenv("PGP_GEMM_VARIANT"); return e ? atoi(e) : 0; }();
    // CTA tile: 64 x 64 when 128 x 128 tiles would leave most SMs idle
    auto count_tiles = [&](int64_t T) {           // CTA tiles that do work (tri skips those above the diagonal)
        const int64_t tmT = ceil_div(a.M, T), tnT = ceil_div(a.N, T);
        if (!a.tri) return tmT * tnT;
        int64_t cnt = 0;
        for (int64_t ti = 0; ti < tmT; ++ti) {
            int64_t last = (ti * T + T - 1 + a.tri_off) / T;          // last tile column on / below the diagonal
            if (ti * T + T - 1 + a.tri_off < 0) continue;
            cnt += std::min(tnT, last + 1);
        }
        return cnt;
    };
    const int64_t tiles128 = count_tiles(128) * a.batch;
    static const int force_tile = [] { const char* e = getenv("PGP_GEMM_TILE"); return e ? atoi(e) : 0; }();
    // ... or when the output is at most 64 wide (K = 64 leaf updates of the recursions, C[M x 64] -= A B^T): a
    // 128-wide tile would compute half padding
    static const int narrow64 = [] { const char* e = getenv("PGP_GEMM_NARROW64"); return e ? atoi(e) : 1; }();
    const bool small = force_tile ? force_tile == 64
                                  : (tiles128 < (int64_t)ctx->sm_count || (narrow64 && (a.N <= 64 || a.M <= 64)));
    const int BM = small ? 64 : 128, BN = BM;
    int64_t tm = ceil_div(a.M, (int64_t)BM), tn = ceil_div(a.N, (int64_t)BN);
    if (tm * tn > 0x7fffffffLL || a.batch > 65535) return ctx->fail(PGP_E_ARG, "gemm: grid too large");
    // split the contraction when the output has too few tiles to fill the GPU
    // (FITC: p x p results contracted over n >> p): partials go to a workspace
    // and are summed in a fixed order
    a.splitk = 1;
    if (a_in.splitk != 1 && !a.krow && !a.kcol && !a.stair && !a.ragged_mstep && a.batch == 1) {
        int64_t tiles = count_tiles(BM);
        int64_t want = a_in.splitk > 1 ? a_in.splitk : (2 * ctx->sm_count) / std::max<int64_t>(tiles, 1);
        int64_t max_by_k = a.K / 2048;  // keep >= 2048 contraction steps per slice
        int64_t S = std::min<int64_t>(std::min<int64_t>(want, max_by_k), 64);
        if (a_in.splitk > 1) S = std::min<int64_t>(a_in.splitk, std::max<int64_t>(a.K / 32, 1));
        if (S > 1) {
            a.splitk = (int)S;
            a.ldws = round_up(a.N, 2);
            size_t need = (size_t)S * a.M * a.ldws;
            if (ctx->gemm_ws_doubles < need) {
                PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if (ctx->gemm_ws) dev_free(ctx, ctx->gemm_ws);
                ctx->gemm_ws = nullptr;
                ctx->gemm_ws_doubles = 0;
                PGP_TRY(dev_alloc(ctx, &ctx->gemm_ws, need));
                ctx->gemm_ws_doubles = need;
            }
            a.ws = ctx->gemm_ws;
        }
    }
    // algorithmic flops of this launch (roofline numerator): per tile row, the
    // columns on/below the diagonal times the contraction length actually needed
    double flops = 0.0;
    for (int64_t ti = 0; ti < tm; ++ti) {
        double rows = (double)std::min<int64_t>(BM, a.M - ti * (int64_t)BM);
        double mid = (double)(ti * BM) + 0.5 * (rows - 1.0);
        double cols = (double)a.N, klen = (double)a.K;
        if (a.tri) cols = std::min(std::max(mid + (double)a.tri_off + 1.0, 0.0), (double)a.N);
        if (a.krow) klen = std::min(std::max((double)a.K - (mid + (double)a.krow_off), 0.0), (double)a.K);
        if (a.stair) {
            const int64_t r0 = ti * (int64_t)BM;
            const double S = r0 >= a.stair_front ? (double)(a.stair_first + (r0 - a.stair_front) / a.stair_nb * a.stair_step) : 0.0;
            if (a.stair & 1) klen = std::min(std::max((double)a.K - (S - (double)a.stair_off), 0.0), klen);
            if (a.stair & 2) cols = std::min(std::max((double)a.N - (S - (double)a.stair_off), 0.0), cols);
        }
        if (a.kcol) {
            // sum over the columns j of (j + 1 + off - kstart) clipped to [0, K - kstart]; exact for the
            // regular triangular cases used here (kstart = K - klen)
            double kstart = (double)a.K - klen, hi = (double)a.N + (double)a.kcol_off;   // k-end of the last column
            double lo_end = 1.0 + (double)a.kcol_off;
            double avg_end = 0.5 * (std::min(std::max(lo_end, kstart), (double)a.K) + std::min(std::max(hi, kstart), (double)a.K));
            klen = std::max(avg_end - kstart, 0.0);
        }
        flops += 2.0 * rows * cols * klen;
    }
    double batch_flops = flops * a.batch;
    if (a.ragged_mstep) {       // member b has M - b mstep rows (rectangular count; tri only trims the top block)
        batch_flops = 0.0;
        for (int b = 0; b < a.batch; ++b)
            batch_flops += 2.0 * (double)std::max<int64_t>(a.M - b * a.ragged_mstep, 0) * (double)a.N * (double)a.K;
    }
    {
        Launch L(ctx, PC_GEMM, batch_flops);
        L.shape(a.M, a.N, a.K, a.tri | (a.krow << 1) | (a.transA << 2) | (a.transB << 3) | (a.kcol << 4) | (a.batch << 8));
        int rc;
        if (small) {
            if (a.transA && a.transB) rc = launch_variant<2, 2, Cfg<64, 32, 3>, true, true>(ctx, a, tm, tn);
            else if (a.transA) rc = launch_variant<2, 2, Cfg<64, 32, 3>, true, false>(ctx, a, tm, tn);
            else if (a.transB) rc = launch_variant<2, 2, Cfg<64, 32, 3>, false, true>(ctx, a, tm, tn);
            else rc = launch_variant<2, 2, Cfg<64, 32, 3>, false, false>(ctx, a, tm, tn);
        } else if (variant == 0 && !a.transA && a.K >= 4096) {
            // long contractions: 8 warps with 64 x 32 warp tiles (0.375 fragment loads per DMMA instead of
            // 0.5) are ~2 % ahead (35.3 vs 34.7 TFLOP/s at 8192^3); for K <= 2048 the 16-warp layout wins
            if (a.transB) rc = launch_variant<2, 4, Cfg<128, 32, 3>, false, true>(ctx, a, tm, tn);
            else rc = launch_variant<2, 4, Cfg<128, 32, 3>, false, false>(ctx, a, tm, tn);
        } else if (a.transA && a.transB) rc = launch_variant<4, 4, Cfg<128, 32, 3>, true, true>(ctx, a, tm, tn);
        else if (a.transA) rc = launch_variant<4, 4, Cfg<128, 32, 3>, true, false>(ctx, a, tm, tn);
        else if (a.transB) rc = launch_variant<4, 4, Cfg<128, 32, 3>, false, true>(ctx, a, tm, tn);
        else switch (variant) {
            case 1: rc = launch_variant<4, 4, Cfg<128, 16, 4>, false, false>(ctx, a, tm, tn); break;
            case 2: rc = launch_variant<2, 4, Cfg<128, 32, 3>, false, false>(ctx, a, tm, tn); break;
            case 4: rc = launch_variant<2, 4, Cfg<128, 16, 4>, false, false>(ctx, a, tm, tn); break;
            default: rc = launch_variant<4, 4, Cfg<128, 32, 3>, false, false>(ctx, a, tm, tn); break;
        }
        PGP_TRY(rc);
        PGP_TRY(check_launch(ctx, "gemm_kernel"));
    }
    if (a.splitk > 1) {
        Launch L(ctx, PC_OTHER, 8.0 * a.splitk * a.M * a.N);
        int blocks = (int)std::min<int64_t>(ceil_div(a.M * a.N, 256), (int64_t)ctx->sm_count * 8);
        splitk_reduce_kernel<<<dim3(blocks, 1, a.batch), 256, 0, ctx->stream>>>(a);
        PGP_TRY(check_launch(ctx, "splitk_reduce_kernel"));
    }
    return 0;
}

int launch_gemm_nt(pgp_ctx* ctx, const GemmArgs& a) {
    GemmArgs g = a;
    g.transA = g.transB = 0;
    g.splitk = 1;
    return launch_gemm(ctx, g);
}

}  // namespace pgp

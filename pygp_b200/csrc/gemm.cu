// gemm.cu -- FP64 tensor-core GEMM for sm_100a:  C = beta C + alpha A B^T.
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

#include <cstdlib>

namespace pgp {

namespace {

constexpr int BM = 128, BN = 128;
// CTAs are rasterised in groups of GROUP_M tile rows (column index fastest
// inside a group) so that the ~148 CTAs in flight share a ~12 x 12 block of
// tiles: each operand k-slice is then fetched from HBM once per wave and
// served to the other CTAs from L2.
constexpr int GROUP_M = 12;

template <int BK_, int STAGES_>
struct Cfg {
    static constexpr int BK = BK_, STAGES = STAGES_;
    static constexpr int LDS = BK + 4;  // padded shared row (doubles): conflict-free fragment loads
    static constexpr int STAGE_DOUBLES = (BM + BN) * LDS;
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
template <int WM, int WN, class C>
__global__ void __launch_bounds__(WM * WN * 32, 1) gemm_nt_kernel(GemmArgs a, int tm, int tn) {
    constexpr int THREADS = WM * WN * 32;
    constexpr int MI = BM / WM / 8, NJ = BN / WN / 8;
    constexpr int BK = C::BK, STAGES = C::STAGES, LDS = C::LDS, STAGE_DOUBLES = C::STAGE_DOUBLES;
    constexpr int CPR = BK / 2;                      // 16-byte chunks per operand row
    constexpr int NCH = (BM * CPR) / THREADS;        // chunks per thread per operand tile
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
    if (a.tri && n0 > m0 + BM - 1 + a.tri_off) return;

    const int b = blockIdx.z;
    const double* __restrict__ A = a.A + (int64_t)b * a.strideA;
    const double* __restrict__ B = a.B + (int64_t)b * a.strideB;
    double* __restrict__ Cm = a.C + (int64_t)b * a.strideC;

    int64_t ks = 0;
    if (a.krow) {
        ks = m0 + a.krow_off;
        if (ks < 0) ks = 0;
        ks = ks / BK * BK;
        if (ks > a.K) ks = a.K;
    }
    const int KT = (int)((a.K - ks + BK - 1) / BK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;

    // per-thread copy descriptors, fixed for the whole k loop: chunk q of this
    // thread covers row (tid + q THREADS) / CPR, k offset 2 ((tid + q THREADS) % CPR)
    const double* srcA[NCH];
    const double* srcB[NCH];
    int soff[NCH], kof[NCH];
    bool okA[NCH], okB[NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
        int c = threadIdx.x + q * THREADS;
        int row = c / CPR;
        kof[q] = (c % CPR) * 2;
        soff[q] = row * LDS + kof[q];
        okA[q] = m0 + row < a.M;
        okB[q] = n0 + row < a.N;
        srcA[q] = A + (okA[q] ? (m0 + row) * a.lda : 0) + ks + kof[q];
        srcB[q] = B + (okB[q] ? (n0 + row) * a.ldb : 0) + ks + kof[q];
    }
    auto load_stage = [&](int slot, int ktile) {
        double* As = smem + slot * STAGE_DOUBLES;
        double* Bs = As + BM * LDS;
        const int64_t koff = (int64_t)ktile * BK;
        const int64_t kleft = a.K - ks - koff;  // valid k from this tile's start
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
            int64_t rem = kleft - kof[q];
            int bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
            cp_async16(As + soff[q], okA[q] && bytes ? srcA[q] + koff : A, okA[q] ? bytes : 0);
            cp_async16(Bs + soff[q], okB[q] && bytes ? srcB[q] + koff : B, okB[q] ? bytes : 0);
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

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * STAGE_DOUBLES + (wm * (BM / WM) + g) * LDS + t;
        const double* Bs = smem + (kt % STAGES) * STAGE_DOUBLES + BM * LDS + (wn * (BN / WN) + g) * LDS + t;
        // fragments double-buffered in registers: the loads of step kk+1 are in
        // flight while the DMMAs of step kk issue
        double af[2][MI], bf[2][NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) af[0][i] = As[i * 8 * LDS];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[0][j] = Bs[j * 8 * LDS];
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            if (kk + 1 < BK / 4) {
#pragma unroll
                for (int i = 0; i < MI; ++i) af[nxt][i] = As[i * 8 * LDS + (kk + 1) * 4];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[nxt][j] = Bs[j * 8 * LDS + (kk + 1) * 4];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: lane owns C[row = 8i + g][col = 8j + 2t, +1] of its warp tile
    const double alpha = a.alpha, beta = a.beta;
    const bool vec_ok = ((a.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cm) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int64_t row = m0 + wm * (BM / WM) + i * 8 + g;
        if (row >= a.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int64_t col = n0 + wn * (BN / WN) + j * 8 + 2 * t;
            if (col >= a.N) continue;
            double* p = Cm + row * a.ldc + col;
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

namespace {
template <int WM, int WN, class C>
int launch_variant(pgp_ctx* ctx, const GemmArgs& a, int64_t tm, int64_t tn) {
    auto kern = gemm_nt_kernel<WM, WN, C>;
    PGP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    dim3 grid((unsigned)(tm * tn), 1, a.batch);
    kern<<<grid, WM * WN * 32, C::SMEM, ctx->stream>>>(a, (int)tm, (int)tn);
    return 0;
}
}  // namespace

int launch_gemm_nt(pgp_ctx* ctx, const GemmArgs& a) {
    if (a.M <= 0 || a.N <= 0) return 0;
    if ((a.lda & 1) || (a.ldb & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
        (reinterpret_cast<uintptr_t>(a.B) & 15) || ((a.strideA | a.strideB) & 1))
        return ctx->fail(PGP_E_ARG, "gemm_nt: A and B must be 16-byte aligned with even leading dimensions");
    static const int variant = [] { const char* e = getenv("PGP_GEMM_VARIANT"); return e ? atoi(e) : 0; }();
    int64_t tm = ceil_div(a.M, BM), tn = ceil_div(a.N, BN);
    if (tm * tn > 0x7fffffffLL || a.batch > 65535) return ctx->fail(PGP_E_ARG, "gemm_nt: grid too large");
    // algorithmic flops of this launch (roofline numerator): per tile row, the
    // columns on/below the diagonal times the contraction length actually needed
    double flops = 0.0;
    for (int64_t ti = 0; ti < tm; ++ti) {
        double rows = (double)std::min<int64_t>(BM, a.M - ti * BM);
        double mid = (double)(ti * BM) + 0.5 * (rows - 1.0);
        double cols = (double)a.N, klen = (double)a.K;
        if (a.tri) cols = std::min(std::max(mid + (double)a.tri_off + 1.0, 0.0), (double)a.N);
        if (a.krow) klen = std::min(std::max((double)a.K - (mid + (double)a.krow_off), 0.0), (double)a.K);
        flops += 2.0 * rows * cols * klen;
    }
    Launch L(ctx, PC_GEMM, flops * a.batch);
    int rc;
    switch (variant) {
        case 1: rc = launch_variant<4, 4, Cfg<16, 4>>(ctx, a, tm, tn); break;
        case 2: rc = launch_variant<2, 4, Cfg<32, 3>>(ctx, a, tm, tn); break;
        case 4: rc = launch_variant<2, 4, Cfg<16, 4>>(ctx, a, tm, tn); break;
        default: rc = launch_variant<4, 4, Cfg<32, 3>>(ctx, a, tm, tn); break;
    }
    PGP_TRY(rc);
    return check_launch(ctx, "gemm_nt_kernel");
}

}  // namespace pgp

// gemm.cu -- FP64 tensor-core GEMM for sm_100a:  C = beta C + alpha A B^T.
//
// Why mma.sync and not tcgen05: tcgen05.mma has no f64 kind; on sm_100a every
// FP64 mma shape lowers to DMMA.8x8x4 (checked with cuobjdump), so the kernel is
// written against m8n8k4 directly.  Bound: FP64 tensor pipe.
//
// Tiling: CTA tile 128 x 128 x 16, 8 warps as 2 (M) x 4 (N), warp tile 64 x 32
// = 8 x 4 DMMA tiles = 64 accumulator doubles per lane.  Operands are staged by
// a 4-deep cp.async ring (16-byte chunks, zero-fill predication at the M/N/K
// edges, so no padding of the matrices is needed).  Shared rows are padded to
// 20 doubles: the DMMA fragment loads (lane -> row lane/4, k lane%4) of a
// half-warp then fall in 16 distinct 8-byte banks.
//
// Structure flags let the same kernel serve every O(N^3) step of the factor,
// the triangular inverse and V V^T without touching zero blocks (gemm.cuh).

#include "gemm.cuh"

#include <cstdlib>

namespace pgp {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int LDS = BK + 4;  // padded shared row (doubles)
constexpr int STAGE_DOUBLES = (BM + BN) * LDS;
constexpr size_t GEMM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);
// CTAs are rasterised in groups of GROUP_M tile rows (column index fastest
// inside a group) so that the ~148 CTAs in flight share a ~12 x 12 block of
// tiles: each operand k-slice is then fetched from HBM once per wave and
// served to the other CTAs from L2.
constexpr int GROUP_M = 12;

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

// one operand tile (128 rows x 16 k) -> shared; rows >= rows_valid and
// k >= k_valid are zero-filled
template <int THREADS>
__device__ __forceinline__ void load_tile(double* sdst, const double* g, int64_t ld, int64_t row0,
                                          int64_t rows_total, int64_t k0, int64_t k_end) {
#pragma unroll
    for (int q = 0; q < (BM * BK / 2) / THREADS; ++q) {
        int c = threadIdx.x + q * THREADS;
        int row = c >> 3;
        int kc = (c & 7) * 2;
        int64_t gr = row0 + row;
        int64_t gk = k0 + kc;
        int64_t rem = k_end - gk;
        int bytes = (gr < rows_total && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
        const double* src = bytes ? g + gr * ld + gk : g;
        cp_async16(sdst + row * LDS + kc, src, bytes);
    }
}

}  // namespace

// WM x WN warps; each warp owns a (BM/WM) x (BN/WN) tile = MI x NJ DMMA tiles.
template <int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32, 1) gemm_nt_kernel(GemmArgs a, int tm, int tn) {
    constexpr int THREADS = WM * WN * 32;
    constexpr int MI = BM / WM / 8, NJ = BN / WN / 8;
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
    double* __restrict__ C = a.C + (int64_t)b * a.strideC;

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

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue: fill STAGES-1 slots
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            double* As = smem + s * STAGE_DOUBLES;
            double* Bs = As + BM * LDS;
            load_tile<THREADS>(As, A, a.lda, m0, a.M, ks + (int64_t)s * BK, a.K);
            load_tile<THREADS>(Bs, B, a.ldb, n0, a.N, ks + (int64_t)s * BK, a.K);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) {
                int s = nk % STAGES;
                double* As = smem + s * STAGE_DOUBLES;
                double* Bs = As + BM * LDS;
                load_tile<THREADS>(As, A, a.lda, m0, a.M, ks + (int64_t)nk * BK, a.K);
                load_tile<THREADS>(Bs, B, a.ldb, n0, a.N, ks + (int64_t)nk * BK, a.K);
            }
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * STAGE_DOUBLES + (wm * (BM / WM) + g) * LDS + t;
        const double* Bs = smem + (kt % STAGES) * STAGE_DOUBLES + BM * LDS + (wn * (BN / WN) + g) * LDS + t;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[MI], bf[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = As[i * 8 * LDS + kk * 4];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = Bs[j * 8 * LDS + kk * 4];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: lane owns C[row = 8i + g][col = 8j + 2t, +1] of its warp tile
    const double alpha = a.alpha, beta = a.beta;
    const bool vec_ok = ((a.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int64_t row = m0 + wm * (BM / WM) + i * 8 + g;
        if (row >= a.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int64_t col = n0 + wn * (BN / WN) + j * 8 + 2 * t;
            if (col >= a.N) continue;
            double* p = C + row * a.ldc + col;
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

int launch_gemm_nt(pgp_ctx* ctx, const GemmArgs& a) {
    if (a.M <= 0 || a.N <= 0) return 0;
    if ((a.lda & 1) || (a.ldb & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
        (reinterpret_cast<uintptr_t>(a.B) & 15) || ((a.strideA | a.strideB) & 1))
        return ctx->fail(PGP_E_ARG, "gemm_nt: A and B must be 16-byte aligned with even leading dimensions");
    static const int variant = [] { const char* e = getenv("PGP_GEMM_VARIANT"); return e ? atoi(e) : 0; }();
    PGP_CUDA(ctx, cudaFuncSetAttribute(gemm_nt_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)GEMM_SMEM));
    PGP_CUDA(ctx, cudaFuncSetAttribute(gemm_nt_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)GEMM_SMEM));
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
    dim3 grid((unsigned)(tm * tn), 1, a.batch);
    if (variant == 1)
        gemm_nt_kernel<4, 4><<<grid, 512, GEMM_SMEM, ctx->stream>>>(a, (int)tm, (int)tn);
    else
        gemm_nt_kernel<2, 4><<<grid, 256, GEMM_SMEM, ctx->stream>>>(a, (int)tm, (int)tn);
    return check_launch(ctx, "gemm_nt_kernel");
}

}  // namespace pgp

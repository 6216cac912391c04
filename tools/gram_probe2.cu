// gram_probe2.cu -- round-2 prototype bench for the Gram tile kernel's epilogue and staging.
// NOT part of the product.  Variants of the symmetric Gram build (lower tiles computed, upper
// tiles mirrored through shared memory; exact direct-difference distances) for SE and Matern-5/2:
//
//   lib       round 1's structure: library sqrt / exp, 4 x 4 register tile, 256 threads
//   fm        fastmath.cuh epilogue (table exp, seed + third-order sqrt), epilogue one micro-tile
//             row (4 entries) at a time so the chains interleave and the row is stored at once
//   fm8       the same with 8 x 4 register tiles, 128 threads per tile
//   fm_tma    fm with the inputs in DIMENSION-MAJOR layout Zt[k][npad], staged by the bulk-copy
//             engine (cp.async.bulk ... mbarrier::complete_tx::bytes, one 512-byte row per issuing
//             thread) instead of per-thread loads + shared stores
//   *_nomirror  lower tiles only (what ExactGP._update needs): 4 N^2 bytes
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o bin/gram_probe2 gram_probe2.cu
//   bin/gram_probe2 [N=32768]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../pygp_b200/csrc/fastmath.cuh"

using namespace pgp;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

constexpr int T = 64;
constexpr double kThird = 1.0 / 3.0;

enum { K_SE = 0, K_M5 = 1 };

template <int KT>
__device__ __forceinline__ double value_lib(double D, double c) {
    if (KT == K_SE) return exp(c - D / 2);
    const double r = sqrt(D);
    return exp(c - r) * (1 + r * (1 + r * kThird));
}

template <int KT>
__device__ __forceinline__ double value_fm(double D, double c, const double* tab, int& bad) {
    if (KT == K_SE) return fm::exp_tab(fma(D, -0.5, c), tab, bad);
    const double r = fm::sqrt_pos(D);
    const double e = fm::exp_tab(c - r, tab, bad);
    return e * fma(r, fma(r, fm::kFmC[5], 1.0), 1.0);
}

// cold path: a thread whose fast epilogue left the normal range (exp_tab's `bad`) redoes its
// micro-tile with libm from the staged inputs and overwrites what it stored
template <int KT, int RA>
__device__ __noinline__ void slow_redo(const double* Zs1, const double* Zs2, double* Tm, int d, double c, double* out,
                                       int64_t ld, int64_t n, int64_t i0, int64_t j0, bool mirror) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int a = 0; a < RA; ++a)
        for (int b = 0; b < 4; ++b) {
            const int row = ty + (T / RA) * a, col = 2 * tx + 32 * (b >> 1) + (b & 1);
            double D = 0.0;
            for (int k = 0; k < d; ++k) {
                const double df = Zs1[k * T + row] - Zs2[k * T + col];
                D = fma(df, df, D);
            }
            const double v = value_lib<KT>(D, c);
            if (i0 + row < n && j0 + col < n) out[(i0 + row) * ld + j0 + col] = v;
            if (mirror) Tm[row * (T + 1) + col] = v;
        }
}

template <int KT>
__global__ void naive_kernel(const double* Z, int64_t n, int d, double c, double* out, int64_t ld) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    double D = 0.0;
    for (int k = 0; k < d; ++k) {
        const double df = Z[i * d + k] - Z[j * d + k];
        D += df * df;
    }
    out[i * ld + j] = value_lib<KT>(D, c);
}

__device__ __forceinline__ void tri_decode(int64_t idx, int* ti, int* tj) {
    int i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((int64_t)(i + 1) * (i + 2) / 2 <= idx) ++i;
    while ((int64_t)i * (i + 1) / 2 > idx) --i;
    *ti = i;
    *tj = (int)(idx - (int64_t)i * (i + 1) / 2);
}

// ---- bulk-copy (TMA engine) helpers -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- the tile kernel ----------------------------------------------------------------------
// RA x 4 micro-tile: rows ty + (64 / RA) a, columns 2 tx + 32 (b >> 1) + (b & 1); 16 * 64 / RA threads
template <int KT, int RA, bool FM, bool TMA, bool MIRROR, int MINB>
__global__ void __launch_bounds__(16 * (T / RA), MINB)
tile_kernel(const double* __restrict__ Z, int64_t n, int64_t npad, int d, double c, double* __restrict__ out, int64_t ld) {
    constexpr int NT = 16 * (T / RA);
    extern __shared__ __align__(128) double sm[];
    double* tab = sm;                       // exp table first: its address is a compile-time offset
    uint64_t* bar = reinterpret_cast<uint64_t*>(tab + fm::kExpTabDoubles);
    double* Tm = tab + fm::kExpTabDoubles + 16;   // [64][65] mirror scratch
    double* Zs1 = Tm + (MIRROR ? T * (T + 1) + 15 : 0);   // [d][64]; 16-byte aligned (bulk copies)
    Zs1 = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(Zs1) + 127) & ~(uintptr_t)127);
    double* Zs2 = Zs1 + d * T;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int ti, tj;
    tri_decode(blockIdx.x, &ti, &tj);
    const int64_t i0 = (int64_t)ti * T, j0 = (int64_t)tj * T;

    if (TMA) {
        // Z is dimension-major [d][npad]: row k of a tile's inputs is 512 contiguous bytes
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)(2 * d * T * sizeof(double)));
        __syncwarp();
        if (tid < 2 * d) {
            const int k = tid >> 1, which = tid & 1;
            bulk_g2s((which ? Zs2 : Zs1) + k * T, Z + (int64_t)k * npad + (which ? j0 : i0), T * sizeof(double), bar);
        }
        if (FM) fm::load_exp_tab(tab, tid, NT);
        mbar_wait(bar, 0);
        __syncthreads();
    } else {
        for (int idx = tid; idx < d * T; idx += NT) {
            const int k = idx / T, r = idx % T;
            const int64_t g1 = i0 + r, g2 = j0 + r;
            Zs1[idx] = g1 < n ? Z[g1 * d + k] : 0.0;
            Zs2[idx] = g2 < n ? Z[g2 * d + k] : 0.0;
        }
        if (FM) fm::load_exp_tab(tab, tid, NT);
        __syncthreads();
    }

    double D[RA][4];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) D[a][b] = 0.0;
#pragma unroll 4
    for (int k = 0; k < d; ++k) {
        double zi[RA], zj[4];
#pragma unroll
        for (int a = 0; a < RA; ++a) zi[a] = Zs1[k * T + ty + (T / RA) * a];
        const double2 q0 = *reinterpret_cast<const double2*>(&Zs2[k * T + 2 * tx]);
        const double2 q1 = *reinterpret_cast<const double2*>(&Zs2[k * T + 2 * tx + 32]);
        zj[0] = q0.x; zj[1] = q0.y; zj[2] = q1.x; zj[3] = q1.y;
#pragma unroll
        for (int a = 0; a < RA; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double df = zi[a] - zj[b];
                D[a][b] = fma(df, df, D[a][b]);
            }
    }
    constexpr int TP = T + 1;
    const bool mirror = MIRROR && ti != tj;
    int bad = 0;
#pragma unroll
    for (int a = 0; a < RA; ++a) {
        double v[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) v[b] = FM ? value_fm<KT>(D[a][b], c, tab, bad) : value_lib<KT>(D[a][b], c);
        const int row = ty + (T / RA) * a;
        const int64_t gi = i0 + row;
        if (gi < n) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t gj = j0 + 2 * tx + 32 * h;
                double* dst = out + gi * ld + gj;
                if (gj + 1 < n) *reinterpret_cast<double2*>(dst) = make_double2(v[2 * h], v[2 * h + 1]);
                else if (gj < n) dst[0] = v[2 * h];
            }
        }
        if (mirror) {
#pragma unroll
            for (int b = 0; b < 4; ++b) Tm[row * TP + 2 * tx + 32 * (b >> 1) + (b & 1)] = v[b];
        }
    }
    if (FM && __builtin_expect(bad, 0)) slow_redo<KT, RA>(Zs1, Zs2, Tm, d, c, out, ld, n, i0, j0, mirror);
    if (mirror) {
        __syncthreads();
#pragma unroll
        for (int a = 0; a < RA; ++a) {
            const int row = ty + (T / RA) * a;
            const int64_t gi = j0 + row;
            if (gi >= n) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cc = 2 * tx + 32 * h;
                const int64_t gj = i0 + cc;
                double* dst = out + gi * ld + gj;
                const double v0 = Tm[cc * TP + row], v1 = Tm[(cc + 1) * TP + row];
                if (gj + 1 < n) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                else if (gj < n) dst[0] = v0;
            }
        }
    }
}

struct Ctx {
    int64_t n, npad, ld;
    int d;
    double c;
    double *dZ, *dZt, *dOut;
    std::vector<double> ref, got;
    int64_t nref;
    cudaEvent_t e0, e1;
};

template <int KT, int RA, bool FM, bool TMA, bool MIRROR, int MINB>
void run(Ctx& C, const char* name) {
    auto kern = tile_kernel<KT, RA, FM, TMA, MIRROR, MINB>;
    const int64_t t1 = (C.n + T - 1) / T, ntiles = t1 * (t1 + 1) / 2;
    const size_t smem = ((size_t)2 * C.d * T + (MIRROR ? T * (T + 1) + 15 : 0) + fm::kExpTabDoubles + 16 + 16) * 8;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 16 * (T / RA), smem));
    auto launch = [&] {
        kern<<<(unsigned)ntiles, 16 * (T / RA), smem>>>(TMA ? C.dZt : C.dZ, C.n, C.npad, C.d, C.c, C.dOut, C.ld);
    };
    CK(cudaMemset(C.dOut, 0, (size_t)C.n * C.ld * 8));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy2D(C.got.data(), C.nref * 8, C.dOut, C.ld * 8, C.nref * 8, C.nref, cudaMemcpyDeviceToHost));
    double err = 0.0;
    for (int64_t i = 0; i < C.nref; ++i)
        for (int64_t j = 0; j < C.nref; ++j) {
            if (!MIRROR && j > i) continue;
            const double r = C.ref[i * C.nref + j], g = C.got[i * C.nref + j];
            err = std::fmax(err, std::fabs(g - r) / std::fmax(std::fabs(r), 1e-300));
        }
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(C.e0));
        launch();
        CK(cudaEventRecord(C.e1));
        CK(cudaEventSynchronize(C.e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, C.e0, C.e1));
        best = ms < best ? ms : best;
    }
    const double bytes = (MIRROR ? 8.0 : 4.0) * C.n * C.n;
    printf("{\"variant\": \"%s\", \"kernel\": \"%s\", \"n\": %lld, \"d\": %d, \"ms\": %.3f, \"GBps\": %.1f, \"hbm_frac_of_6545\": %.3f, "
           "\"regs\": %d, \"ctas_per_sm\": %d, \"max_rel_err_vs_lib\": %.3e}\n",
           name, KT == K_SE ? "se" : "matern5", (long long)C.n, C.d, best, bytes / best / 1e6, bytes / best / 1e6 / 6545.3,
           fa.numRegs, occ, err);
    fflush(stdout);
}

template <int KT>
void suite(int64_t n, int d) {
    Ctx C;
    C.n = n; C.d = d; C.c = 0.3;
    C.npad = (n + T - 1) / T * T;
    C.ld = (n + 15) / 16 * 16;
    std::vector<double> hZ((size_t)n * d), hZt((size_t)C.npad * d, 0.0);
    srand(0);
    const double scale = KT == K_SE ? 1.0 / (0.5 * std::sqrt((double)d)) : std::sqrt(5.0) / (0.5 * std::sqrt((double)d));
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < d; ++k) {
            const double v = (rand() / (double)RAND_MAX) * scale;
            hZ[i * d + k] = v;
            hZt[(size_t)k * C.npad + i] = v;
        }
    CK(cudaMalloc(&C.dZ, hZ.size() * 8));
    CK(cudaMalloc(&C.dZt, hZt.size() * 8));
    CK(cudaMalloc(&C.dOut, (size_t)n * C.ld * 8));
    CK(cudaMemcpy(C.dZ, hZ.data(), hZ.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(C.dZt, hZt.data(), hZt.size() * 8, cudaMemcpyHostToDevice));
    C.nref = n < 2048 ? n : 2048;
    double* dRef;
    CK(cudaMalloc(&dRef, (size_t)C.nref * C.nref * 8));
    naive_kernel<KT><<<dim3((unsigned)((C.nref + 255) / 256), (unsigned)C.nref), 256>>>(C.dZ, C.nref, d, C.c, dRef, C.nref);
    CK(cudaDeviceSynchronize());
    C.ref.resize((size_t)C.nref * C.nref);
    C.got.resize((size_t)C.nref * C.nref);
    CK(cudaMemcpy(C.ref.data(), dRef, C.ref.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaEventCreate(&C.e0));
    CK(cudaEventCreate(&C.e1));

    run<KT, 4, false, false, true, 1>(C, "lib_4x4");
    run<KT, 4, true, false, true, 1>(C, "fm_4x4");
    run<KT, 4, true, false, true, 3>(C, "fm_4x4_min3");
    run<KT, 4, true, false, true, 4>(C, "fm_4x4_min4");
    run<KT, 4, true, false, true, 5>(C, "fm_4x4_min5");
    run<KT, 8, true, false, true, 1>(C, "fm_8x4");
    run<KT, 8, true, false, true, 6>(C, "fm_8x4_min6");
    run<KT, 4, true, true, true, 1>(C, "fm_tma_4x4");
    run<KT, 4, true, true, true, 4>(C, "fm_tma_4x4_min4");
    run<KT, 8, true, true, true, 1>(C, "fm_tma_8x4");
    run<KT, 4, true, false, false, 1>(C, "fm_4x4_nomirror");
    run<KT, 4, true, true, false, 4>(C, "fm_tma_4x4_min4_nomirror");

    CK(cudaFree(C.dZ)); CK(cudaFree(C.dZt)); CK(cudaFree(C.dOut)); CK(cudaFree(dRef));
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 32768;
    suite<K_M5>(n, 16);
    suite<K_SE>(n, 8);
    suite<K_SE>(n, 1);
    return 0;
}

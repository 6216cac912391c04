// gram_probe.cu -- prototype bench for the round-2 Gram tile kernel (plan item 1).
// NOT part of the product: a stand-alone experiment that times structural variants of the
// symmetric Matern-5/2 Gram build (N x N, d dimensions, exact direct differences) against each
// other and checks them against a naive kernel.  Written at the end of round 1 from the ncu
// capture of the production kernel (latency-bound at 25-37 % occupancy: every CTA runs
// stage -> barrier -> distances -> epilogue -> store -> barrier -> mirror with nothing overlapped);
// first run: profiles/r01g_gram_probe.txt (v0 2.63 TB/s, v1 2.38-2.45, v2 2.65-2.72 at d = 16, N = 32768).
//
//   v0  the production structure: one CTA per lower tile, inputs staged [k][row] by plain loads
//   v1  persistent CTAs walking the lower tiles row by row; the inputs come from a
//       DIMENSION-MAJOR copy Zt[k][N] (what scale_kernel would write), so a tile's inputs are d
//       contiguous 512-byte segments staged with cp.async into a double buffer while the
//       previous tile computes; the row operand is reused along a tile row
//   v2  v1 with the epilogue's exp as one polynomial per entry on constants held in constant
//       memory (no library call; see profiles/r01f_gram_batched_epilogue_experiment.txt)
//   v3  v2 with 8 x 4 register tiles (128 threads per tile): 8 + 2 shared loads per 64 FP64
//       operations instead of 4 + 2 per 32 -- NOT YET RUN (added after the GPU budget of round 1 ended)
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o bin/gram_probe gram_probe.cu
//   bin/gram_probe [N=32768] [d=16]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

constexpr int T = 64;            // tile edge
constexpr int THREADS = 256;     // 16 x 16 threads, 4 x 4 entries each
constexpr int MAXD = 32;
constexpr double kThird = 1.0 / 3.0;

__device__ __forceinline__ double matern5(double D, double two_logsf) {
    const double r = sqrt(D);
    return exp(two_logsf - r) * (1 + r * (1 + r * kThird));
}

// ---- naive reference: one thread per entry, Z row-major [N][d] ---------------------------------
__global__ void naive_kernel(const double* Z, int64_t n, int d, double two_logsf, double* out, int64_t ld) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    double D = 0.0;
    for (int k = 0; k < d; ++k) {
        const double df = Z[i * d + k] - Z[j * d + k];
        D += df * df;
    }
    out[i * ld + j] = matern5(D, two_logsf);
}

__device__ __forceinline__ void tri_decode(int64_t idx, int* ti, int* tj) {
    int i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((int64_t)(i + 1) * (i + 2) / 2 <= idx) ++i;
    while ((int64_t)i * (i + 1) / 2 > idx) --i;
    *ti = i;
    *tj = (int)(idx - (int64_t)i * (i + 1) / 2);
}

// the RA x 4 micro-tile of a thread (RA = 4: 256 threads, RA = 8: 128 threads per tile):
// rows ty + (64 / RA) a, columns 2 tx + 32 (b >> 1) + (b & 1)
template <int RA>
struct Map {
    int tx, ty;
    __device__ Map() : tx(threadIdx.x & 15), ty(threadIdx.x >> 4) {}
    __device__ int row(int a) const { return ty + (T / RA) * a; }
    __device__ int col(int b) const { return 2 * tx + 32 * (b >> 1) + (b & 1); }
};

template <bool POLY>
__device__ __forceinline__ double value(double D, double two_logsf);

template <>
__device__ __forceinline__ double value<false>(double D, double two_logsf) { return matern5(D, two_logsf); }

// exp(x) = 2^n P13(r), sqrt(D) = D rsqrt(D) (third-order step from the hardware seed); |x| <= 700 assumed by
// the probe's data (the product patches the rest in a cold block)
__constant__ double kC[18] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                              1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
                              1.4426950408889634074, 6755399441055744.0, 6.93147180369123816490e-01,
                              1.90821492927058770002e-10};
template <>
__device__ __forceinline__ double value<true>(double D, double two_logsf) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(D));
    const double t0 = D * y, e0 = fma(-t0, y, 1.0);
    y = fma(y * e0, fma(0.375, e0, 0.5), y);
    const double r = D > 1e-280 ? D * y : 0.0;
    const double x = two_logsf - r;
    const double t = fma(x, kC[14], kC[15]);
    const int n = __double2loint(t);
    const double nd = t - kC[15];
    const double q = fma(nd, -kC[17], fma(nd, -kC[16], x));
    double p = fma(kC[13], q, kC[12]);
#pragma unroll
    for (int k = 11; k >= 0; --k) p = fma(p, q, kC[k]);
    const double e = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return e * (1 + r * (1 + r * kThird));
}

// distances + epilogue + stores of one tile from staged inputs Zs1 / Zs2 ([k][T] each); Tm = mirror scratch
template <bool POLY, int RA = 4>
__device__ __forceinline__ void tile_compute(const double* Zs1, const double* Zs2, double* Tm, int d, double two_logsf,
                                             double* out, int64_t ld, int64_t n, int ti, int tj) {
    Map<RA> t;
    double D[RA][4];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) D[a][b] = 0.0;
    for (int k = 0; k < d; ++k) {
        double zi[RA], zj[4];
#pragma unroll
        for (int a = 0; a < RA; ++a) zi[a] = Zs1[k * T + t.row(a)];
        const double2 q0 = *reinterpret_cast<const double2*>(&Zs2[k * T + 2 * t.tx]);
        const double2 q1 = *reinterpret_cast<const double2*>(&Zs2[k * T + 2 * t.tx + 32]);
        zj[0] = q0.x; zj[1] = q0.y; zj[2] = q1.x; zj[3] = q1.y;
#pragma unroll
        for (int a = 0; a < RA; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double df = zi[a] - zj[b];
                D[a][b] += df * df;
            }
    }
    double res[RA][4];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) res[a][b] = value<POLY>(D[a][b], two_logsf);
    const int64_t i0 = (int64_t)ti * T, j0 = (int64_t)tj * T;
#pragma unroll
    for (int a = 0; a < RA; ++a) {
        const int64_t gi = i0 + t.row(a);
        if (gi >= n) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t gj = j0 + t.col(2 * h);
            double* dst = out + gi * ld + gj;
            if (gj + 1 < n) *reinterpret_cast<double2*>(dst) = make_double2(res[a][2 * h], res[a][2 * h + 1]);
            else if (gj < n) dst[0] = res[a][2 * h];
        }
    }
    if (ti != tj) {   // mirrored tile through a shared transpose
        constexpr int TP = T + 1;
#pragma unroll
        for (int a = 0; a < RA; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) Tm[t.row(a) * TP + t.col(b)] = res[a][b];
        __syncthreads();
#pragma unroll
        for (int a = 0; a < RA; ++a) {
            const int64_t gi = j0 + t.row(a);
            if (gi >= n) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = t.col(2 * h);
                const int64_t gj = i0 + c;
                double* dst = out + gi * ld + gj;
                const double v0 = Tm[c * TP + t.row(a)], v1 = Tm[(c + 1) * TP + t.row(a)];
                if (gj + 1 < n) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                else if (gj < n) dst[0] = v0;
            }
        }
    }
}

// ---- v0: one CTA per lower tile, row-major Z staged transposed by plain loads (production) -------
__global__ void __launch_bounds__(THREADS) v0_kernel(const double* Z, int64_t n, int d, double two_logsf, double* out,
                                                     int64_t ld) {
    extern __shared__ __align__(16) double sm[];
    double* Zs1 = sm;
    double* Zs2 = Zs1 + d * T;
    double* Tm = Zs2 + d * T;
    int ti, tj;
    tri_decode(blockIdx.x, &ti, &tj);
    for (int idx = threadIdx.x; idx < d * T; idx += THREADS) {
        const int k = idx / T, r = idx % T;
        const int64_t g1 = (int64_t)ti * T + r, g2 = (int64_t)tj * T + r;
        Zs1[idx] = g1 < n ? Z[g1 * d + k] : 0.0;
        Zs2[idx] = g2 < n ? Z[g2 * d + k] : 0.0;
    }
    __syncthreads();
    tile_compute<false>(Zs1, Zs2, Tm, d, two_logsf, out, ld, n, ti, tj);
}

// ---- v1 / v2: persistent CTAs, dimension-major inputs, cp.async double buffer ------------------
__device__ __forceinline__ void cp16(double* sdst, const double* gsrc, int bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(bytes));
}

// stage rows [r0, r0 + T) of every dimension of Zt ([d][npad], npad a multiple of T and of 2) into dst [k][T]
__device__ __forceinline__ void stage_async(double* dst, const double* Zt, int64_t npad, int64_t r0, int d) {
    for (int idx = threadIdx.x; idx < d * (T / 2); idx += blockDim.x) {
        const int k = idx / (T / 2), c = (idx % (T / 2)) * 2;
        cp16(dst + k * T + c, Zt + (int64_t)k * npad + r0 + c, 16);
    }
}

template <bool POLY, int RA = 4>
__global__ void __launch_bounds__(16 * (T / RA)) v1_kernel(const double* Zt, int64_t n, int64_t npad, int d, double two_logsf,
                                                     double* out, int64_t ld, int64_t ntiles) {
    extern __shared__ __align__(16) double sm[];
    // [2] x (Zs1, Zs2) + mirror scratch
    double* buf[2] = {sm, sm + 2 * d * T};
    double* Tm = sm + 4 * d * T;
    int64_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    int ti, tj;
    tri_decode(tile, &ti, &tj);
    stage_async(buf[0], Zt, npad, (int64_t)ti * T, d);
    stage_async(buf[0] + d * T, Zt, npad, (int64_t)tj * T, d);
    asm volatile("cp.async.commit_group;\n" ::);
    int cur = 0;
    while (true) {
        const int64_t next = tile + gridDim.x;
        int ni = 0, nj = 0;
        if (next < ntiles) {
            tri_decode(next, &ni, &nj);
            stage_async(buf[cur ^ 1], Zt, npad, (int64_t)ni * T, d);
            stage_async(buf[cur ^ 1] + d * T, Zt, npad, (int64_t)nj * T, d);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 1;\n" ::);    // the current tile's inputs have landed
        __syncthreads();
        tile_compute<POLY, RA>(buf[cur], buf[cur] + d * T, Tm, d, two_logsf, out, ld, n, ti, tj);
        __syncthreads();                                // everybody is done with buf[cur] and Tm
        if (next >= ntiles) break;
        tile = next; ti = ni; tj = nj;
        cur ^= 1;
    }
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 32768;
    const int d = argc > 2 ? atoi(argv[2]) : 16;
    if (d > MAXD) return 1;
    const int64_t npad = (n + T - 1) / T * T, ld = (n + 15) / 16 * 16;
    const double two_logsf = 0.0;
    std::vector<double> hZ((size_t)n * d), hZt((size_t)npad * d, 0.0);
    srand(0);
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < d; ++k) {
            const double v = (rand() / (double)RAND_MAX) / (0.5 * std::sqrt((double)d)) * std::sqrt(5.0);
            hZ[i * d + k] = v;
            hZt[(size_t)k * npad + i] = v;
        }
    double *dZ, *dZt, *dOut, *dRef;
    CK(cudaMalloc(&dZ, hZ.size() * 8));
    CK(cudaMalloc(&dZt, hZt.size() * 8));
    CK(cudaMalloc(&dOut, (size_t)n * ld * 8));
    CK(cudaMemcpy(dZ, hZ.data(), hZ.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dZt, hZt.data(), hZt.size() * 8, cudaMemcpyHostToDevice));
    const int64_t nref = n < 4096 ? n : 4096;          // reference on the leading block only
    CK(cudaMalloc(&dRef, (size_t)nref * nref * 8));
    naive_kernel<<<dim3((unsigned)((nref + 255) / 256), (unsigned)nref), 256>>>(dZ, nref, d, two_logsf, dRef, nref);
    CK(cudaDeviceSynchronize());
    std::vector<double> ref((size_t)nref * nref), got((size_t)nref * nref);
    CK(cudaMemcpy(ref.data(), dRef, ref.size() * 8, cudaMemcpyDeviceToHost));

    const int64_t t1 = (n + T - 1) / T, ntiles = t1 * (t1 + 1) / 2;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto launch) {
        CK(cudaMemset(dOut, 0, (size_t)n * ld * 8));
        launch();
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy2D(got.data(), nref * 8, dOut, ld * 8, nref * 8, nref, cudaMemcpyDeviceToHost));
        double err = 0.0;
        for (size_t i = 0; i < ref.size(); ++i) err = std::fmax(err, std::fabs(got[i] - ref[i]) / std::fmax(std::fabs(ref[i]), 1e-300));
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        printf("{\"variant\": \"%s\", \"n\": %lld, \"d\": %d, \"ms\": %.3f, \"GBps\": %.1f, \"max_rel_err_vs_naive\": %.3e}\n", name,
               (long long)n, d, best, 8.0 * n * n / best / 1e6, err);
    };
    {
        const size_t smem = ((size_t)2 * d * T + T * (T + 1)) * 8;
        CK(cudaFuncSetAttribute(v0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        run("v0_one_cta_per_tile", [&] { v0_kernel<<<(unsigned)ntiles, THREADS, smem>>>(dZ, n, d, two_logsf, dOut, ld); });
    }
    for (int per_sm : {2, 3, 4}) {
        const size_t smem = ((size_t)4 * d * T + T * (T + 1)) * 8;
        const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sms * per_sm);
        char name[96];
        CK(cudaFuncSetAttribute(v1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        snprintf(name, sizeof name, "v1_persistent_%d_per_sm", per_sm);
        run(name, [&] { v1_kernel<false><<<grid, THREADS, smem>>>(dZt, n, npad, d, two_logsf, dOut, ld, ntiles); });
        CK(cudaFuncSetAttribute(v1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        snprintf(name, sizeof name, "v2_persistent_poly_exp_%d_per_sm", per_sm);
        run(name, [&] { v1_kernel<true><<<grid, THREADS, smem>>>(dZt, n, npad, d, two_logsf, dOut, ld, ntiles); });
    }
    for (int per_sm : {3, 4, 6}) {      // v3: 8 x 4 register tiles, 128 threads
        const size_t smem = ((size_t)4 * d * T + T * (T + 1)) * 8;
        const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sms * per_sm);
        char name[96];
        CK(cudaFuncSetAttribute(v1_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        snprintf(name, sizeof name, "v3_persistent_poly_exp_8x4_%d_per_sm", per_sm);
        run(name, [&] { v1_kernel<true, 8><<<grid, 128, smem>>>(dZt, n, npad, d, two_logsf, dOut, ld, ntiles); });
    }
    return 0;
}

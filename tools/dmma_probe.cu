// dmma_probe.cu -- what limits the DMMA inner loop of gemm_kernel?  Replays the
// loop (fragment loads from shared memory + m8n8k4 DMMAs) WITHOUT global traffic,
// sweeping warps per SM, warp-tile shape, fragment source (registers / LDS) and
// a per-k-tile __syncthreads, and prints TFLOP/s for each combination.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: operands in registers (pure pipe); 1: LDS fragments, no barrier;
// 2: LDS fragments + __syncthreads every BK=32 (8 k-steps); 3: as 2 with a second
// barrier (models wait + sync)
template <int MI, int NJ, int MODE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) loop_kernel(double* out, int ktiles) {
    extern __shared__ double sm[];
    constexpr int LDS = 36;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 256 * LDS; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* As = sm + ((warp & 3) * 8 * MI % 128 + g) * LDS + t;
    const double* Bs = sm + 128 * LDS + ((warp >> 2) * 8 * NJ % 128 + g) * LDS + t;
    double ra = 1.0 + lane * 1e-3, rb = 1.0 - lane * 1e-3;
    for (int kt = 0; kt < ktiles; ++kt) {
        if (MODE >= 2) __syncthreads();
        if (MODE >= 3) __syncthreads();
        double af[2][MI], bf[2][NJ];
        if (MODE >= 1) {
#pragma unroll
            for (int i = 0; i < MI; ++i) af[0][i] = As[i * 8 * LDS];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[0][j] = Bs[j * 8 * LDS];
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            if (MODE >= 1 && kk + 1 < 8) {
#pragma unroll
                for (int i = 0; i < MI; ++i) af[nxt][i] = As[i * 8 * LDS + (kk + 1) * 4];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[nxt][j] = Bs[j * 8 * LDS + (kk + 1) * 4];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (MODE == 0) dmma(acc[i][j][0], acc[i][j][1], ra, rb);
                    else dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
                }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// MODE 4: the full software pipeline of gemm_kernel: 3-stage ring filled by
// cp.async (LDGSTS, 16 B per thread, 8 per thread per k-tile, interleaved with the
// DMMAs), fragments read from the ring.  MODE 5: same copies, all issued in one
// burst after the barrier.  MODE 6: as 4 but fragments read from a FIXED stage
// (copies are pure background traffic).
__device__ __forceinline__ void cp16(double* sdst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) pipe_kernel(double* out, const double* __restrict__ G, int ktiles, int ld) {
    extern __shared__ double sm[];
    constexpr int LDS = 36, STAGE = 256 * LDS, STAGES = 3, MI = 4, NJ = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    for (int i = threadIdx.x; i < STAGES * STAGE; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // chunk c = tid + q * 512: row c / 16 (of 256 rows: A then B), k offset 2 (c % 16)
    const int row0 = threadIdx.x / 16, kof = (threadIdx.x % 16) * 2;
    const double* src0 = G + ((size_t)blockIdx.x * 256 + row0) * ld + kof;
    const int soff0 = row0 * LDS + kof;
    auto load_chunk = [&](int slot, int kt, int q) {
        cp16(sm + slot * STAGE + soff0 + q * 32 * LDS, src0 + (size_t)q * 32 * ld + (size_t)(kt % 64) * 32);
    };
    for (int s = 0; s < STAGES - 1; ++s) {
#pragma unroll
        for (int q = 0; q < 8; ++q) load_chunk(s, s, q);
        asm volatile("cp.async.commit_group;\n" ::);
    }
    for (int kt = 0; kt < ktiles; ++kt) {
        asm volatile("cp.async.wait_group 1;\n" ::);
        __syncthreads();
        const int nk = kt + STAGES - 1, nslot = nk % STAGES;
        if (MODE == 5) {
#pragma unroll
            for (int q = 0; q < 8; ++q) load_chunk(nslot, nk, q);
        }
        const int rs = MODE == 6 ? 0 : kt % STAGES;
        const double* As = sm + rs * STAGE + (wm * 32 + g) * LDS + t;
        const double* Bs = sm + rs * STAGE + 128 * LDS + (wn * 32 + g) * LDS + t;
        double af[2][MI], bf[2][NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) af[0][i] = As[i * 8 * LDS];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[0][j] = Bs[j * 8 * LDS];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            if (kk + 1 < 8) {
#pragma unroll
                for (int i = 0; i < MI; ++i) af[nxt][i] = As[i * 8 * LDS + (kk + 1) * 4];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[nxt][j] = Bs[j * 8 * LDS + (kk + 1) * 4];
            }
            if (MODE != 5) load_chunk(nslot, nk, kk);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run_pipe(const char* name, double* d, const double* G, int sms) {
    auto kern = pipe_kernel<MODE>;
    size_t smem = 3 * 256 * 36 * sizeof(double);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ktiles = 2000, ld = 64 * 32;
    kern<<<sms, 512, smem>>>(d, G, 50, ld);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        kern<<<sms, 512, smem>>>(d, G, ktiles, ld);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double flops = (double)sms * 16 * ktiles * 8.0 * 16 * 512.0;
    printf("%-52s : %7.2f TFLOP/s%s\n", name, flops / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : "  (LAUNCH FAILED)");
}

template <int MI, int NJ, int MODE, int MAXT = (MI * NJ > 16 ? 256 : 512), int MINB = 1>
static void run(const char* name, double* d, int sms, int warps_per_cta, int ctas_per_sm) {
    auto kern = loop_kernel<MI, NJ, MODE, MAXT, MINB>;
    size_t smem = 256 * 36 * sizeof(double);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ktiles = 2000;
    int blocks = sms * ctas_per_sm, threads = warps_per_cta * 32;
    kern<<<blocks, threads, smem>>>(d, 50);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        kern<<<blocks, threads, smem>>>(d, ktiles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double flops = (double)blocks * warps_per_cta * ktiles * 8.0 * MI * NJ * 512.0;
    printf("%-34s warps/CTA %2d CTAs/SM %d : %7.2f TFLOP/s%s\n", name, warps_per_cta, ctas_per_sm,
           flops / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : "  (LAUNCH FAILED)");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, sizeof(double) * sms * 4 * 1024);
    for (int w : {4, 8, 16}) {
        run<4, 4, 0>("reg operands 32x32", d, sms, w, 1);
        run<4, 4, 1>("LDS frags 32x32", d, sms, w, 1);
        run<4, 4, 2>("LDS frags 32x32 + 1 barrier/ktile", d, sms, w, 1);
        run<4, 4, 3>("LDS frags 32x32 + 2 barriers/ktile", d, sms, w, 1);
    }
    for (int w : {4, 8}) {
        run<8, 4, 0>("reg operands 64x32", d, sms, w, 1);
        run<8, 4, 1>("LDS frags 64x32", d, sms, w, 1);
        run<8, 4, 2>("LDS frags 64x32 + 1 barrier/ktile", d, sms, w, 1);
        run<8, 8, 0>("reg operands 64x64", d, sms, w, 1);
        run<8, 8, 1>("LDS frags 64x64", d, sms, w, 1);
        run<8, 8, 2>("LDS frags 64x64 + 1 barrier/ktile", d, sms, w, 1);
    }
    run<4, 4, 2>("LDS frags 32x32 + 1 barrier/ktile", d, sms, 8, 2);
    run<4, 4, 2>("LDS frags 32x32 + 1 barrier/ktile", d, sms, 4, 4);
    run<2, 4, 2, 512, 2>("LDS frags 16x32 + 1 barrier/ktile", d, sms, 16, 2);
    double* G; cudaMalloc(&G, sizeof(double) * (size_t)sms * 256 * 64 * 32);
    cudaMemset(G, 0, sizeof(double) * (size_t)sms * 256 * 64 * 32);
    run_pipe<4>("pipeline: cp.async interleaved, frags from ring", d, G, sms);
    run_pipe<5>("pipeline: cp.async burst after barrier", d, G, sms);
    run_pipe<6>("pipeline: cp.async interleaved, frags from fixed stage", d, G, sms);
    return 0;
}

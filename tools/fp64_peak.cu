// fp64_peak.cu -- measures the two FP64 pipes of the GPU it runs on:
//   DMMA.8x8x4 (mma.sync m8n8k4 f64, the tensor pipe the GEMM uses) and DFMA.
// Prints one JSON line; bench.py / DESIGN.md use it as the FP64 roofline
// denominator next to cuBLAS DGEMM (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_loop(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_loop(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static double time_ms(K kern, double* d, int blocks, int threads, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<blocks, threads>>>(d, iters / 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) return 1e30;   // launch failed: never the best
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    double best_dmma = 0, best_dfma = 0; int wd = 0, wf = 0;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = 128;
        int blocks = sms * (warps * 32 / threads);
        double ms = time_ms(dmma_loop, d, blocks, threads, iters);
        double tf = (double)blocks * (threads / 32) * iters * 16 * 512.0 / (ms * 1e-3) / 1e12;
        fprintf(stderr, "warps/SM %2d: dmma %.2f TF", warps, tf);
        if (tf > best_dmma) { best_dmma = tf; wd = warps; }
        ms = time_ms(dfma_loop, d, blocks, threads, iters);
        double tf2 = (double)blocks * threads * iters * 16 * 2.0 / (ms * 1e-3) / 1e12;
        fprintf(stderr, "  dfma %.2f TF\n", tf2);
        if (tf2 > best_dfma) { best_dfma = tf2; wf = warps; }
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dmma_tflops\": %.2f, \"dmma_warps_per_sm\": %d, "
           "\"dfma_tflops\": %.2f, \"dfma_warps_per_sm\": %d}\n", p.name, sms, best_dmma, wd, best_dfma, wf);
    return 0;
}

// chol.cu -- blocked FP64 Cholesky, triangular solve, triangular inverse and
// V V^T on row-major device buffers.
//
// Replaces scipy.linalg.cholesky / solve_triangular / cho_solve at
// pygp/inference/exact.py:54-55,88,128-129 (LAPACK dpotrf / dtrtrs / dpotrs).
//
// Formulation (oracle/blocked_model.py is the numpy model of exactly this):
//   chol_rec(j0, n): factor columns [j0, j0+n) for ALL rows below them
//       n <= 64 : potrf_base (one CTA, shared memory) + trsm_base (one thread
//                 per row, register-resident forward substitution)
//       else    : chol_rec(left half); trapezoid update of the right half by
//                 one DMMA GEMM (tiles above the diagonal skipped); recurse.
//   Every flop above the 64-wide base case is a DMMA GEMM with K >= 64, and
//   three quarters of them have K >= n/4.  Appending r = y - mean as row n
//   turns the TRSV of exact.py:55 into one more row of the same solves.
//   inv_upper: V = L^-T as a structured TRSM of the identity (n^3/3 flops),
//   syrk_upper_lower: K~^-1 = V V^T on the lower tiles with k >= row (n^3/3).

#include "chol.cuh"
#include "spec.cuh"

namespace pgp {

namespace {

inline int64_t split_point(int64_t n) {
    int64_t h = ceil_div(n / 2, (int64_t)kNB) * kNB;
    if (h <= 0) h = kNB;
    if (h >= n) h = n - kNB;
    return h;
}

// ---------------------------------------------------------------------------
// potrf_base: one CTA factors the n x n (n <= 64) diagonal block at (j0, j0)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) potrf_base_kernel(double* F, int64_t ld, int64_t bstride,
                                                         int64_t j0, int n, int* info) {
    __shared__ double A[kNB][kNB + 1];
    __shared__ double col[kNB];
    __shared__ double diag[kNB];
    double* Fb = F + (int64_t)blockIdx.x * bstride + j0 * ld + j0;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < n * n; idx += 256) {
        int r = idx / n, c = idx - r * n;
        if (c <= r) A[r][c] = Fb[(int64_t)r * ld + c];
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    for (int k = 0; k < n; ++k) {
        double d = A[k][k];
        double s;
        if (d > 0.0) {
            s = sqrt(d);
        } else {
            // not positive definite (or NaN): report the first failing minor
            if (tid == 0) atomicCAS(info + blockIdx.x, 0, (int)(j0 + k + 1));
            s = nan("");
        }
        if (tid == k) diag[k] = s;
        if (tid > k && tid < n) {
            double l = A[tid][k] / s;
            col[tid] = l;
            A[tid][k] = l;
        }
        __syncthreads();
        for (int i = k + 1 + ty; i < n; i += 16) {
            double li = col[i];
            for (int j = k + 1 + tx; j <= i; j += 16) A[i][j] -= li * col[j];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < n * n; idx += 256) {
        int r = idx / n, c = idx - r * n;
        if (c < r) Fb[(int64_t)r * ld + c] = A[r][c];
        else if (c == r) Fb[(int64_t)r * ld + c] = diag[r];
    }
}

// ---------------------------------------------------------------------------
// trsm_base: X = B T^-T for the n <= 64 columns [j0, j0+n) of `rows` rows of B,
// T = L[j0.., j0..] lower.  One thread per row: the row lives in registers and
// is forward-substituted against T broadcast from shared memory.
// IDENT: B is the identity (rows are rows j0.. of I): used by the inverse.
// ---------------------------------------------------------------------------
constexpr int kTrsmRows = 128;

template <bool IDENT>
__global__ void __launch_bounds__(kTrsmRows) trsm_base_kernel(double* B, int64_t ldb, int64_t bstrideB,
                                                              int64_t rows, const double* L, int64_t ldl,
                                                              int64_t bstrideL, int64_t j0, int n) {
    extern __shared__ __align__(16) double sm[];
    double* Lt = sm;                          // [64][64]  Lt[k][j] = T[j][k], j > k
    double* rinv = Lt + kNB * kNB;            // [64]
    double* Bt = rinv + kNB;                  // [128][65]
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const double* Lb = L + (int64_t)b * bstrideL + j0 * ldl + j0;
    double* Bb = B + (int64_t)b * bstrideB + j0;
    const int64_t r0 = (int64_t)blockIdx.x * kTrsmRows;

    for (int idx = tid; idx < kNB * kNB; idx += kTrsmRows) {
        int j = idx >> 6, k = idx & 63;  // read row j of T along k (coalesced)
        double v = 0.0;
        if (j < n && k < j) v = Lb[(int64_t)j * ldl + k];
        Lt[k * kNB + j] = v;
        if (j == k) rinv[k] = (k < n) ? 1.0 / Lb[(int64_t)k * ldl + k] : 1.0;
    }
    for (int idx = tid; idx < kTrsmRows * kNB; idx += kTrsmRows) {
        int r = idx >> 6, c = idx & 63;
        double v = 0.0;
        if (IDENT) {
            v = (r0 + r == c) ? 1.0 : 0.0;
        } else if (r0 + r < rows && c < n) {
            v = Bb[(r0 + r) * ldb + c];
        }
        Bt[r * (kNB + 1) + c] = v;
    }
    __syncthreads();

    double x[kNB];
#pragma unroll
    for (int c = 0; c < kNB; ++c) x[c] = Bt[tid * (kNB + 1) + c];
#pragma unroll
    for (int k = 0; k < kNB; ++k) {
        x[k] *= rinv[k];
        const double xk = x[k];
#pragma unroll
        for (int j = k + 1; j < kNB; ++j) x[j] -= xk * Lt[k * kNB + j];
    }
#pragma unroll
    for (int c = 0; c < kNB; ++c) Bt[tid * (kNB + 1) + c] = x[c];
    __syncthreads();

    for (int idx = tid; idx < kTrsmRows * kNB; idx += kTrsmRows) {
        int r = idx >> 6, c = idx & 63;
        if (r0 + r < rows && c < n) Bb[(r0 + r) * ldb + c] = Bt[r * (kNB + 1) + c];
    }
}

constexpr size_t kTrsmSmem = (kNB * kNB + kNB + kTrsmRows * (kNB + 1)) * sizeof(double);

int launch_potrf_base(pgp_ctx* ctx, const Mat& F, int64_t j0, int n, int* d_info) {
    Launch L(ctx, PC_POTRF, (double)n * n * n / 3.0 * F.batch);
    potrf_base_kernel<<<F.batch, 256, 0, ctx->stream>>>(F.p, F.ld, F.bstride, j0, n, d_info);
    return check_launch(ctx, "potrf_base_kernel");
}

template <bool IDENT>
int launch_trsm_base(pgp_ctx* ctx, const Mat& B, int64_t rows, const Mat& L, int64_t j0, int n) {
    if (rows <= 0) return 0;
    auto kern = trsm_base_kernel<IDENT>;
    PGP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem));
    int64_t blocks = ceil_div(rows, kTrsmRows);
    Launch Lc(ctx, PC_TRSM, (double)rows * n * n * B.batch);
    kern<<<dim3((unsigned)blocks, B.batch), kTrsmRows, kTrsmSmem, ctx->stream>>>(
        B.p, B.ld, B.bstride, rows, L.p, L.ld, L.bstride, j0, n);
    return check_launch(ctx, "trsm_base_kernel");
}

int gemm_update(pgp_ctx* ctx, const double* A, int64_t lda, int64_t sA, const double* Bm, int64_t ldb,
                int64_t sB, double* C, int64_t ldc, int64_t sC, int64_t M, int64_t N, int64_t K, double alpha,
                double beta, int tri, int krow, int batch) {
    GemmArgs g;
    g.A = A; g.lda = lda; g.strideA = sA;
    g.B = Bm; g.ldb = ldb; g.strideB = sB;
    g.C = C; g.ldc = ldc; g.strideC = sC;
    g.M = M; g.N = N; g.K = K;
    g.alpha = alpha; g.beta = beta;
    g.tri = tri; g.krow = krow;
    g.batch = batch;
    return launch_gemm_nt(ctx, g);
}

// factor columns [j0, j0+n) of the `mrows`-row buffer
int chol_rec(pgp_ctx* ctx, const Mat& F, int64_t j0, int64_t n, int64_t mrows, int* d_info) {
    if (n <= kNB) {
        PGP_TRY(launch_potrf_base(ctx, F, j0, (int)n, d_info));
        Mat B = F;
        B.p = F.p + (j0 + n) * F.ld;
        return launch_trsm_base<false>(ctx, B, mrows - (j0 + n), F, j0, (int)n);
    }
    int64_t n1 = split_point(n), n2 = n - n1, c0 = j0 + n1;
    PGP_TRY(chol_rec(ctx, F, j0, n1, mrows, d_info));
    // rows c0.., cols [c0, c0+n2) -= P P2^T, P = rows c0.. of cols [j0, c0)
    const double* P = F.p + c0 * F.ld + j0;
    PGP_TRY(gemm_update(ctx, P, F.ld, F.bstride, P, F.ld, F.bstride, F.p + c0 * F.ld + c0, F.ld, F.bstride,
                        mrows - c0, n2, n1, -1.0, 1.0, /*tri=*/1, /*krow=*/0, F.batch));
    return chol_rec(ctx, F, c0, n2, mrows, d_info);
}

int trsm_rec(pgp_ctx* ctx, const Mat& B, int64_t rows, const Mat& L, int64_t j0, int64_t n) {
    if (n <= kNB) return launch_trsm_base<false>(ctx, B, rows, L, j0, (int)n);
    int64_t n1 = split_point(n), n2 = n - n1, c0 = j0 + n1;
    PGP_TRY(trsm_rec(ctx, B, rows, L, j0, n1));
    PGP_TRY(gemm_update(ctx, B.p + j0, B.ld, B.bstride, L.p + c0 * L.ld + j0, L.ld, L.bstride, B.p + c0, B.ld,
                        B.bstride, rows, n2, n1, -1.0, 1.0, 0, 0, B.batch));
    return trsm_rec(ctx, B, rows, L, c0, n2);
}

int inv_rec(pgp_ctx* ctx, const Mat& G, const Mat& L, int64_t j0, int64_t n) {
    if (n <= kNB) {
        Mat B = G;
        B.p = G.p + j0 * G.ld;
        return launch_trsm_base<true>(ctx, B, n, L, j0, (int)n);
    }
    int64_t n1 = split_point(n), n2 = n - n1, c0 = j0 + n1;
    PGP_TRY(inv_rec(ctx, G, L, j0, n1));
    // G[j0:c0, c0:c0+n2] = -V11 L21^T ; V11 upper triangular -> k >= row
    PGP_TRY(gemm_update(ctx, G.p + j0 * G.ld + j0, G.ld, G.bstride, L.p + c0 * L.ld + j0, L.ld, L.bstride,
                        G.p + j0 * G.ld + c0, G.ld, G.bstride, n1, n2, n1, -1.0, 0.0, 0, /*krow=*/1, G.batch));
    Mat B = G;
    B.p = G.p + j0 * G.ld;
    PGP_TRY(trsm_rec(ctx, B, n1, L, c0, n2));
    return inv_rec(ctx, G, L, c0, n2);
}

}  // namespace

int potrf_lower(pgp_ctx* ctx, const Mat& F, int64_t n, int64_t extra, int* d_info) {
    if (n <= 0) return 0;
    return chol_rec(ctx, F, 0, n, n + extra, d_info);
}

int trsm_right_lt(pgp_ctx* ctx, const Mat& B, int64_t rows, const Mat& L, int64_t n) {
    if (n <= 0 || rows <= 0) return 0;
    return trsm_rec(ctx, B, rows, L, 0, n);
}

int inv_upper(pgp_ctx* ctx, const Mat& G, const Mat& L, int64_t n) {
    if (n <= 0) return 0;
    return inv_rec(ctx, G, L, 0, n);
}

int syrk_upper_lower(pgp_ctx* ctx, const Mat& H, const Mat& G, int64_t n) {
    if (n <= 0) return 0;
    return gemm_update(ctx, G.p, G.ld, G.bstride, G.p, G.ld, G.bstride, H.p, H.ld, H.bstride, n, n, n, 1.0, 0.0,
                       /*tri=*/1, /*krow=*/1, G.batch);
}

// ---------------------------------------------------------------------------
// small reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
    return r;
}

__global__ void loglik_kernel(const double* F, int64_t ld, int64_t bstride, int64_t n, double* out) {
    __shared__ double red[8];
    const double* Fb = F + (int64_t)blockIdx.x * bstride;
    double sa = 0.0, sl = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double a = Fb[n * ld + i];
        sa += a * a;
        sl += log(Fb[i * ld + i]);
    }
    sa = block_sum_256(sa, red);
    sl = block_sum_256(sl, red);
    if (threadIdx.x == 0)
        out[blockIdx.x] = -0.5 * sa - 0.5 * log(2 * kPi) * (double)n - sl;
}

int launch_loglik(pgp_ctx* ctx, const Mat& F, int64_t n, double* d_out) {
    Launch L(ctx, PC_OTHER, 16.0 * n * F.batch);
    loglik_kernel<<<F.batch, 256, 0, ctx->stream>>>(F.p, F.ld, F.bstride, n, d_out);
    return check_launch(ctx, "loglik_kernel");
}

__global__ void gemv_upper_kernel(const double* G, int64_t ld, const double* a, int64_t n, double* alpha) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const double* g = G + row * ld;
    double s = 0.0;
    for (int64_t k = row + lane; k < n; k += 32) s += g[k] * a[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) alpha[row] = s;
}

int launch_gemv_upper(pgp_ctx* ctx, const double* G, int64_t ld, const double* a, int64_t n, double* alpha) {
    Launch L(ctx, PC_OTHER, 4.0 * n * n);
    gemv_upper_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, ctx->stream>>>(G, ld, a, n, alpha);
    return check_launch(ctx, "gemv_upper_kernel");
}

__global__ void set_residual_kernel(double* F, int64_t ld, int64_t bstride, int64_t n, const double* y,
                                    const DevSpec* spec) {
    const double mean = spec[blockIdx.y].h.mean;
    double* row = F + (int64_t)blockIdx.y * bstride + n * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        row[i] = y[i] - mean;
}

int launch_set_residual(pgp_ctx* ctx, const Mat& F, int64_t n, const double* d_y, const DevSpec* d_spec) {
    int blocks = (int)std::min<int64_t>(ceil_div(n, 256), 148);
    Launch L(ctx, PC_OTHER, 16.0 * n * F.batch);
    set_residual_kernel<<<dim3(blocks, F.batch), 256, 0, ctx->stream>>>(F.p, F.ld, F.bstride, n, d_y, d_spec);
    return check_launch(ctx, "set_residual_kernel");
}

// one warp per test point
__global__ void predict_reduce_kernel(const double* B, int64_t ld, int64_t rows, int64_t n, const double* a,
                                      const DevSpec* spec, double* mu, double* s2, int64_t bsB, int64_t bsA,
                                      int64_t bsO) {
    __shared__ double kdiag;
    const int b = blockIdx.y;
    const DevSpecHdr& S = spec[b].h;
    if (threadIdx.x == 0) {
        PartVal pv[kMaxParts];
        double val[kMaxNodes];
        for (int p = 0; p < S.n_parts; ++p) part_eval<false>(S.parts[p], 0.0, pv[p]);
        kdiag = tree_forward(S, pv, val);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double* v = B + (int64_t)b * bsB + row * ld;
    const double* ab = a + (int64_t)b * bsA;
    double sm = 0.0, sv = 0.0;
    for (int64_t k = lane; k < n; k += 32) {
        double x = v[k];
        sm += x * ab[k];
        sv += x * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if (lane == 0) {
        mu[(int64_t)b * bsO + row] = S.mean + sm;
        s2[(int64_t)b * bsO + row] = kdiag - sv;
    }
}

int launch_predict_reduce(pgp_ctx* ctx, const double* B, int64_t ld, int64_t rows, int64_t n, const double* a,
                          const DevSpec* d_spec, double* mu, double* s2, int batch, int64_t bstrideB,
                          int64_t bstrideA, int64_t bstrideOut) {
    if (rows <= 0) return 0;
    Launch L(ctx, PC_OTHER, 8.0 * rows * n * batch);
    predict_reduce_kernel<<<dim3((unsigned)ceil_div(rows, 8), batch), 256, 0, ctx->stream>>>(
        B, ld, rows, n, a, d_spec, mu, s2, bstrideB, bstrideA, bstrideOut);
    return check_launch(ctx, "predict_reduce_kernel");
}

__global__ void extract_upper_kernel(const double* F, int64_t ld, int64_t n, double* R) {
    // R[i][j] = L[j][i] for j >= i else 0 ; tile transpose through shared memory
    __shared__ double tile[32][33];
    int64_t bi = (int64_t)blockIdx.y * 32, bj = (int64_t)blockIdx.x * 32;
    // read L[bj + ty][bi + tx] (row of L contiguous)
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t lr = bj + r, lc = bi + threadIdx.x;
        tile[r][threadIdx.x] = (lr < n && lc < n && lc <= lr) ? F[lr * ld + lc] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = bi + r, j = bj + threadIdx.x;
        if (i < n && j < n) R[i * n + j] = tile[threadIdx.x][r];
    }
}

int launch_extract_upper(pgp_ctx* ctx, const double* F, int64_t ld, int64_t n, double* R) {
    if (n <= 0) return 0;
    unsigned t = (unsigned)ceil_div(n, 32);
    Launch L(ctx, PC_OTHER, 16.0 * n * n);
    extract_upper_kernel<<<dim3(t, t), dim3(32, 8), 0, ctx->stream>>>(F, ld, n, R);
    return check_launch(ctx, "extract_upper_kernel");
}

}  // namespace pgp

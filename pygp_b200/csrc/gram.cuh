// gram.cuh -- host-callable launchers of the covariance tile kernels (gram.cu).
#pragma once

#include "spec.cuh"

namespace pgp {

constexpr int kTile = 64;  // covariance tile edge (entries)

// Scaled inputs are held DIMENSION-MAJOR: Z[b][p][k][i] = X[i][k] / ell[b][p][k] for i < n
// (pygp/kernels/_distances.py:17-23), zero for n <= i < z_stride(n).  Row k of a tile's 64
// inputs is then 512 contiguous bytes, which the tile kernels stage with one bulk-copy
// (cp.async.bulk, the TMA engine) per row instead of 8 strided loads + shared stores per
// thread; the padding lets every tile read 64 rows without a bound check.
inline int64_t z_stride(int64_t n) { return round_up(n, kTile) + kTile; }
inline size_t z_doubles(int n_parts, int ndim, int64_t n) { return (size_t)n_parts * ndim * z_stride(n); }
// batch > 1: spec[b], Z[b][...] (batch stride z_doubles) with X shared.
int launch_scale(pgp_ctx* ctx, const DevSpec* d_spec, const double* d_X, int64_t n, int ndim,
                 int n_parts, double* d_Z, int batch);

struct GramArgs {
    const DevSpec* spec = nullptr;   // device, [batch]
    const double* Z1 = nullptr;      // [batch][parts][d][zd1]
    const double* Z2 = nullptr;      // [batch][parts][d][zd2] (may alias Z1)
    int64_t n1 = 0, n2 = 0;
    int64_t zd1 = 0, zd2 = 0;        // doubles between consecutive dimensions (0: z_stride(n1) / z_stride(n2));
                                     // Z1 + r0 with the parent's stride is a row window of a larger array
    int ndim = 0, n_parts = 0;
    double* out = nullptr;           // [batch] (n1, ldo)
    int64_t ldo = 0;
    int64_t out_bstride = 0;
    int lower_only = 0;              // skip tiles strictly above the diagonal
    int symmetric = 0;               // Z2 == Z1, full square: compute lower tiles, mirror them
    int add_noise = 0;               // out[i][i] += spec.sn2
    int hidx = -1;                   // >= 0: write d/d hyper[hidx] instead of K
    int xdim = -1;                   // >= 0: write d k(x1, x2) / d x1[xdim] instead of K (gradx)
    int ydim = -1;                   // with xdim: write d2 k / d x1[xdim] d x2[ydim] (gradxy; SE leaves only)
    int64_t ostride = 1;             // doubles between consecutive columns of `out`
    int batch = 1;
    int single_type = -1;            // leaf type when n_parts == 1 (fast path)
};
int launch_gram(pgp_ctx* ctx, const GramArgs& a);

struct TraceArgs {
    const DevSpec* spec = nullptr;
    const double* Z = nullptr;       // [parts][d][z_stride(n)]
    int64_t n = 0;
    int64_t zd = 0;                  // 0: z_stride(n)
    int ndim = 0, n_parts = 0, nhyper = 0;
    const double* P = nullptr;       // lower triangle of K~^-1, (n, ldp)
    int64_t ldp = 0;
    const double* alpha = nullptr;   // (n)
    double* partials = nullptr;      // [n_cta][nhyper + 1]
    double* dlZ = nullptr;           // device (nhyper + 2): final gradient
    int single_type = -1;
};
// number of CTAs (= rows of `partials`) the trace launch will use
int64_t trace_cta_count(int64_t n);
// dlZ = [-sn2 tr(Q), -1/2 sum(Q o dK_h)..., sum(alpha)]   (exact.py:131-141)
int launch_trace(pgp_ctx* ctx, const TraceArgs& a);

// out[h] += scale * sum_ij W_ij dK_h(x1_i, x2_j), h < nhyper, over a full
// rectangular block; W dense (mode 0) or FITC's Cxu built on the fly (mode 1).
struct TraceRectArgs {
    const DevSpec* spec = nullptr;
    const double* Z1 = nullptr;      // [parts][d][z_stride(n1)]
    const double* Z2 = nullptr;      // [parts][d][z_stride(n2)]
    int64_t n1 = 0, n2 = 0;
    int64_t zd1 = 0, zd2 = 0;        // 0: z_stride(n1) / z_stride(n2)
    int ndim = 0, n_parts = 0, nhyper = 0;
    int mode = 0;
    int sym = 0;                     // mode 0, n1 == n2, X1 == X2: Wd holds the lower triangle of a symmetric W
    const double* Wd = nullptr;      // mode 0: (n1, ldw)
    const double* Bt = nullptr;      // mode 1: (n1, ldw)
    const double* T2 = nullptr;      // mode 1: (n1, ldw)
    const double* al = nullptr;      // mode 1: (n1) alpha
    const double* q = nullptr;       // mode 1: (n1); nullptr = all ones (DTC)
    const double* wv = nullptr;      // mode 1: (n2) w
    int64_t ldw = 0;
    double scale = 1.0;
    double* partials = nullptr;      // [trace_rect_cta_count][nhyper + 1]
    double* out = nullptr;           // device (nhyper), accumulated into
    int single_type = -1;
};
int64_t trace_rect_cta_count(int64_t n1, int64_t n2);
int launch_trace_rect(pgp_ctx* ctx, const TraceRectArgs& a);

// Block-column partition of the gradient trace (dist.cu): this rank holds the rows of K~^-1 that belong
// to its block columns, B[r][i] = K~^-1[i][c(r)] valid for i >= c(r), where local row r = q nb + o is
// global column c = (rank + q size) nb + o.  partials[cta][0] = sum of Q_cc, [1 + h] = the rank's
// share of sum_ij Q_ij dK_h,ij (entries i > c counted twice, i == c once), Q = K~^-1 - alpha alpha^T.
struct TraceDistArgs {
    const DevSpec* spec = nullptr;
    const double* Z = nullptr;       // [parts][d][zd], all n rows
    int64_t zd = 0;
    int64_t n = 0;
    int ndim = 0, n_parts = 0, nhyper = 0;
    const double* B = nullptr;       // (rows_local, ldb)
    int64_t ldb = 0;
    int64_t rows_local = 0;
    int64_t nb = 0;
    int rank = 0, size = 1;
    const double* alpha = nullptr;   // (n)
    double* partials = nullptr;      // [trace_dist_cta_count][nhyper + 1]
    int single_type = -1;
};
int64_t trace_dist_cta_count(int64_t rows_local, int64_t n);
int launch_trace_dist(pgp_ctx* ctx, const TraceDistArgs& a);

// out[h][i] = d k(x_i, x_i) / d hyper_h for h < nhyper (hmode=1) or k(x_i,x_i) (hmode=0)
int launch_diag(pgp_ctx* ctx, const DevSpec* d_spec, int64_t n, int hmode, int nhyper, double* d_out);

// elementwise fastmath.cuh functions on device arrays (accuracy tests)
int launch_fastmath(pgp_ctx* ctx, int which, const double* d_x, int64_t n, double* d_out);

}  // namespace pgp

// gram.cuh -- host-callable launchers of the covariance tile kernels (gram.cu).
#pragma once

#include "spec.cuh"

namespace pgp {

constexpr int kTile = 64;  // covariance tile edge (entries)

// Z[p][i][k] = X[i][k] / ell[p][k]   (pygp/kernels/_distances.py:17-23)
// batch > 1: spec[b], Z[b][...] with X shared.
int launch_scale(pgp_ctx* ctx, const DevSpec* d_spec, const double* d_X, int64_t n, int ndim,
                 int n_parts, double* d_Z, int batch);

struct GramArgs {
    const DevSpec* spec = nullptr;   // device, [batch]
    const double* Z1 = nullptr;      // [batch][parts][n1][d]
    const double* Z2 = nullptr;      // [batch][parts][n2][d] (may alias Z1)
    int64_t n1 = 0, n2 = 0;
    int64_t zs1 = 0, zs2 = 0;        // doubles between the leaves' copies (0: n1 * ndim / n2 * ndim);
                                     // lets Z1 / Z2 be a row window of a larger scaled array
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
    const double* Z = nullptr;       // [parts][n][d]
    int64_t n = 0;
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
    const double* Z1 = nullptr;      // [parts][n1][d]
    const double* Z2 = nullptr;      // [parts][n2][d]
    int64_t n1 = 0, n2 = 0;
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

// out[h][i] = d k(x_i, x_i) / d hyper_h for h < nhyper (hmode=1) or k(x_i,x_i) (hmode=0)
int launch_diag(pgp_ctx* ctx, const DevSpec* d_spec, int64_t n, int hmode, int nhyper, double* d_out);

}  // namespace pgp

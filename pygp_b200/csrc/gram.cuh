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
    int ndim = 0, n_parts = 0;
    double* out = nullptr;           // [batch] (n1, ldo)
    int64_t ldo = 0;
    int64_t out_bstride = 0;
    int lower_only = 0;              // skip tiles strictly above the diagonal
    int add_noise = 0;               // out[i][i] += spec.sn2
    int hidx = -1;                   // >= 0: write d/d hyper[hidx] instead of K
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

// out[h][i] = d k(x_i, x_i) / d hyper_h for h < nhyper (hmode=1) or k(x_i,x_i) (hmode=0)
int launch_diag(pgp_ctx* ctx, const DevSpec* d_spec, int64_t n, int hmode, int nhyper, double* d_out);

}  // namespace pgp

/* Plain-C client of libpygp_b200.so: the same calls a cgo / JNI / ctypes binding
 * makes (include/pygp_b200.h), with no Python and no torch in the process.
 * ExactGP with an SE-ARD kernel on a small synthetic problem: update,
 * loglikelihood(grad) and posterior at three points; prints them as one line
 *   lZ dlZ[0..nh-1] mu[0..2] s2[0..2]
 * tests/test_abi.py compiles it (CPU) and tests/test_exact_gpu.py runs it and
 * compares with the Python host path (GPU). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pygp_b200.h"

int main(void) {
    enum { N = 200, D = 3, M = 3 };
    static double X[N * D], y[N], Xs[M * D];
    unsigned s = 12345u;
    for (int i = 0; i < N * D; ++i) { s = s * 1664525u + 1013904223u; X[i] = (double)(s >> 8) / 16777216.0; }
    for (int i = 0; i < N; ++i) y[i] = sin(3.0 * (X[i * D] + X[i * D + 1] + X[i * D + 2]));
    for (int i = 0; i < M * D; ++i) { s = s * 1664525u + 1013904223u; Xs[i] = (double)(s >> 8) / 16777216.0; }

    pgp_kernel_spec spec;
    memset(&spec, 0, sizeof spec);
    spec.ndim = D;
    spec.nhyper = 1 + D;                 /* log sf, log ell_1..D   (pygp/kernels/se.py:46-51) */
    spec.n_parts = 1;
    spec.n_ops = 1;
    spec.parts[0].type = PGP_SE;
    spec.parts[0].iso = 0;
    spec.parts[0].hyper_offset = 0;
    spec.parts[0].nhyper = 1 + D;
    spec.ops[0].op = PGP_OP_PUSH;
    spec.ops[0].arg = 0;

    /* GP hyper vector [log sn | kernel | mean]  (pygp/inference/_base.py:91-105) */
    double hyp[1 + 1 + D + 1] = {log(0.1), log(1.2), log(0.5), log(0.6), log(0.7), 0.05};
    const int nh = 1 + spec.nhyper + 1;

    pgp_ctx* ctx = NULL;
    int rc = pgp_ctx_create(0, &ctx);
    if (rc) { fprintf(stderr, "ctx: %s\n", pgp_last_error(NULL)); return 2; }
    pgp_model* m = NULL;
    if ((rc = pgp_exact_create(ctx, &spec, X, y, N, &m))) { fprintf(stderr, "create: %s\n", pgp_last_error(ctx)); return 3; }
    if ((rc = pgp_exact_update(m, hyp))) { fprintf(stderr, "update: %d %s\n", rc, pgp_last_error(ctx)); return 4; }
    double lZ, dlZ[8], mu[M], s2[M];
    if ((rc = pgp_exact_loglike(m, 1, &lZ, dlZ))) { fprintf(stderr, "loglike: %s\n", pgp_last_error(ctx)); return 5; }
    if ((rc = pgp_exact_predict(m, Xs, M, mu, s2))) { fprintf(stderr, "predict: %s\n", pgp_last_error(ctx)); return 6; }
    printf("%.17g", lZ);
    for (int i = 0; i < nh; ++i) printf(" %.17g", dlZ[i]);
    for (int i = 0; i < M; ++i) printf(" %.17g", mu[i]);
    for (int i = 0; i < M; ++i) printf(" %.17g", s2[i]);
    printf("\n");
    /* the batched entry point: two hyper vectors in one call, values and gradients (optimization.py:54-62 over
     * restarts); the first one is `hyp`, so it must reproduce the single-model result */
    {
        double hyps[2 * (1 + 1 + D + 1)], blZ[2], bdlZ[2 * 8];
        int32_t binfo[2];
        for (int i = 0; i < nh; ++i) { hyps[i] = hyp[i]; hyps[nh + i] = hyp[i] + 0.05; }
        if ((rc = pgp_batched_loglike(ctx, &spec, X, y, N, hyps, 2, blZ, bdlZ, binfo))) {
            fprintf(stderr, "batched: %s\n", pgp_last_error(ctx));
            return 8;
        }
        if (binfo[0] || binfo[1] || fabs(blZ[0] - lZ) > 1e-10 * fabs(lZ)) { fprintf(stderr, "batched lZ mismatch\n"); return 9; }
        for (int i = 0; i < nh; ++i)
            if (fabs(bdlZ[i] - dlZ[i]) > 1e-8 * (1.0 + fabs(dlZ[i]))) { fprintf(stderr, "batched dlZ mismatch\n"); return 10; }
    }
    /* the distributed entry points on a communicator of ONE rank (no NCCL needed): block-column factorisation and
     * block-column gradient must reproduce pgp_exact_update / pgp_exact_loglike */
    {
        pgp_dist* comm = NULL;
        double lZd, dlZd[8];
        if ((rc = pgp_dist_init(ctx, 1, 0, NULL, &comm))) { fprintf(stderr, "dist init: %s\n", pgp_last_error(ctx)); return 11; }
        if ((rc = pgp_dist_exact_update(comm, m, hyp, 64))) { fprintf(stderr, "dist update: %d %s\n", rc, pgp_last_error(ctx)); return 12; }
        if ((rc = pgp_dist_exact_loglike(comm, m, 64, 1, &lZd, dlZd))) { fprintf(stderr, "dist loglike: %s\n", pgp_last_error(ctx)); return 13; }
        if (fabs(lZd - lZ) > 1e-10 * fabs(lZ)) { fprintf(stderr, "dist lZ mismatch %.17g %.17g\n", lZd, lZ); return 14; }
        for (int i = 0; i < nh; ++i)
            if (fabs(dlZd[i] - dlZ[i]) > 1e-8 * (1.0 + fabs(dlZ[i]))) { fprintf(stderr, "dist dlZ mismatch\n"); return 15; }
        pgp_dist_destroy(comm);
    }
    /* error convention: a non positive-definite matrix comes back as LAPACK info > 0 */
    double bad[1 + 1 + D + 1] = {log(1e-12), log(1.0), log(50.0), log(50.0), log(50.0), 0.0};
    rc = pgp_exact_update(m, bad);
    fprintf(stderr, "singular update -> rc %d (%s)\n", rc, pgp_last_error(ctx));
    pgp_model_destroy(m);
    pgp_ctx_destroy(ctx);
    return rc > 0 ? 0 : 7;
}

// fitc.cu -- FITC sparse pseudo-input GP (pygp/inference/fitc.py).
// Placeholder while the ExactGP path is brought up: entry points exist so the
// ABI is complete, and fail loudly (no CPU fallback).
#include "common.cuh"

struct pgp_fitc { pgp_ctx* ctx; };

extern "C" int pgp_fitc_create(pgp_ctx* ctx, const pgp_kernel_spec*, const double*, int64_t, const double*,
                               const double*, int64_t, pgp_fitc** out) {
    if (out) *out = nullptr;
    if (!ctx) return PGP_E_ARG;
    return ctx->fail(PGP_E_STATE, "FITC device path not implemented yet");
}
extern "C" void pgp_fitc_destroy(pgp_fitc* f) { delete f; }
extern "C" int pgp_fitc_update(pgp_fitc* f, const double*) { return f ? f->ctx->fail(PGP_E_STATE, "FITC not implemented") : PGP_E_ARG; }
extern "C" int pgp_fitc_loglike(pgp_fitc* f, int, double*, double*) { return f ? f->ctx->fail(PGP_E_STATE, "FITC not implemented") : PGP_E_ARG; }
extern "C" int pgp_fitc_predict(pgp_fitc* f, const double*, int64_t, double*, double*) { return f ? f->ctx->fail(PGP_E_STATE, "FITC not implemented") : PGP_E_ARG; }

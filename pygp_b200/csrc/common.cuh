// common.cuh -- context, error plumbing, launch accounting shared by all
// translation units of libpygp_b200.so.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pygp_b200.h"

namespace pgp {

enum ProfClass { PC_GEMM = 0, PC_GRAM = 1, PC_TRACE = 2, PC_POTRF = 3, PC_TRSM = 4, PC_OTHER = 5 };

struct ProfRec {
    cudaEvent_t e0, e1;
    int cls;
    double work;
    int64_t m = 0, n = 0, k = 0;   // GEMM shape (PGP_PROF_DUMP)
    int flags = 0;
};

struct Status {
    int code = 0;
    bool ok() const { return code == 0; }
};

}  // namespace pgp

struct pgp_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // panel stream of the lookahead Cholesky (created on first use)
    std::vector<cudaStream_t> aux;    // extra streams of the batched gradient (independent problems side by side)
    std::vector<cudaEvent_t> sync_events;   // reusable untimed events for cross-stream ordering
    std::string err;
    int64_t launches = 0;
    bool profile = false;
    std::vector<pgp::ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;
    // small pinned staging area for hypers / results
    double* h_pin = nullptr;
    size_t h_pin_doubles = 0;
    // split-K workspace of the GEMM (grown on demand, stream-ordered reuse)
    double* gemm_ws = nullptr;
    size_t gemm_ws_doubles = 0;
    // cache of large device buffers released by destroyed models: cudaMalloc /
    // cudaFree of multi-GiB buffers cost ~0.1 s each, more than the transfer of
    // the inputs they serve (the e2e path creates a model per call)
    struct PoolEntry { void* p; size_t bytes; };
    std::unordered_map<void*, size_t> live;                   // every device allocation handed out -> its size
    std::vector<PoolEntry> pool;
    size_t pool_bytes = 0;
    static constexpr size_t kPoolMaxEntries = 512;            // distinct cached buffers
    static constexpr size_t kPoolCap = (size_t)96 << 30;      // keep at most 96 GiB cached

    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e),
                 cudaGetErrorString(e), file, line, what);
        err = buf;
        return e == cudaErrorMemoryAllocation ? PGP_E_NOMEM : PGP_E_CUDA;
    }
};

#define PGP_CUDA(ctx, call)                                                      \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define PGP_TRY(expr)                   \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != 0) return rc__;     \
    } while (0)

namespace pgp {

// RAII bracket around one kernel launch: counts it and, when profiling is on,
// records an event pair on the context stream.
struct Launch {
    pgp_ctx* ctx;
    int idx = -1;
    Launch(pgp_ctx* c, int cls, double work) : ctx(c) {
        ctx->launches++;
        if (ctx->profile) {
            ProfRec r;
            r.cls = cls;
            r.work = work;
            cudaEventCreate(&r.e0);
            cudaEventCreate(&r.e1);
            cudaEventRecord(r.e0, ctx->stream);
            ctx->prof.push_back(r);
            idx = (int)ctx->prof.size() - 1;
        }
    }
    void shape(int64_t m, int64_t n, int64_t k, int flags) {
        if (idx >= 0) {
            ProfRec& r = ctx->prof[idx];
            r.m = m; r.n = n; r.k = k; r.flags = flags;
        }
    }
    ~Launch() {
        if (idx >= 0) cudaEventRecord(ctx->prof[idx].e1, ctx->stream);
    }
};

inline int check_launch(pgp_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ctx->cuda_fail(e, what, __FILE__, __LINE__);
    return 0;
}

// ---- device memory ---------------------------------------------------------------
// Every device buffer of the library comes from dev_alloc and goes back through
// dev_free, which parks it in a per-context cache (exact-size reuse) instead of
// calling cudaFree: cudaMalloc / cudaFree of multi-GiB buffers cost ~0.1 s each, and
// even a 4 MB cudaFree was measured at up to 0.6 s while tens of GiB are mapped
// (tools/e2e_probe.py) -- more than the H2D copy of the inputs the buffers serve.
// Reuse is stream-ordered: all work of a context runs on (or is joined back to) its
// stream, so a recycled buffer cannot be touched by an earlier user any more.
inline int dev_alloc_bytes(pgp_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return 0;
    for (size_t i = 0; i < ctx->pool.size(); ++i) {
        if (ctx->pool[i].bytes == bytes) {
            *p = ctx->pool[i].p;
            ctx->pool_bytes -= bytes;
            ctx->pool.erase(ctx->pool.begin() + i);
            ctx->live[*p] = bytes;
            return 0;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && !ctx->pool.empty()) {   // give the cache back and retry
        cudaGetLastError();
        for (auto& en : ctx->pool) cudaFree(en.p);
        ctx->pool.clear();
        ctx->pool_bytes = 0;
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        char buf[256];
        snprintf(buf, sizeof buf, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return ctx->fail(PGP_E_NOMEM, buf);
    }
    ctx->live[*p] = bytes;
    return 0;
}

inline void dev_free(pgp_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->live.find(p);
    if (it == ctx->live.end()) {        // not ours (should not happen): release it the slow way
        cudaFree(p);
        return;
    }
    const size_t bytes = it->second;
    ctx->live.erase(it);
    if (ctx->pool_bytes + bytes <= pgp_ctx::kPoolCap && ctx->pool.size() < pgp_ctx::kPoolMaxEntries) {
        ctx->pool.push_back({p, bytes});
        ctx->pool_bytes += bytes;
    } else {
        cudaFree(p);
    }
}

template <class T>
inline int dev_alloc(pgp_ctx* ctx, T** p, size_t count) {
    return dev_alloc_bytes(ctx, reinterpret_cast<void**>(p), count * sizeof(T));
}

// free and forget: most owners keep the pointer in a struct field that is tested later
template <class T>
inline void dev_free_null(pgp_ctx* ctx, T*& p) {
    dev_free(ctx, p);
    p = nullptr;
}

inline void pool_release(pgp_ctx* ctx) {
    for (auto& e : ctx->pool) cudaFree(e.p);
    ctx->pool.clear();
    ctx->pool_bytes = 0;
}

// opt a kernel in to `bytes` of dynamic shared memory.  The attribute is sticky, so
// it is set only when a launch needs more than any earlier one did (a high-water
// mark per (device, kernel entry point)): cudaFuncSetAttribute costs ~1-2 us of
// host time, which shows when thousands of short kernels are chained.
inline int ensure_dyn_smem_ptr(pgp_ctx* ctx, const void* kernel, size_t bytes) {
    static std::map<std::pair<int, const void*>, size_t> high;
    static std::mutex mu;                       // contexts of different devices may live on different threads
    std::lock_guard<std::mutex> lock(mu);
    size_t& h = high[std::make_pair(ctx->device, kernel)];
    if (bytes <= h) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "cudaFuncSetAttribute", __FILE__, __LINE__);
    h = bytes;
    return 0;
}
template <class K>
inline int ensure_dyn_smem(pgp_ctx* ctx, K kernel, size_t bytes) {
    return ensure_dyn_smem_ptr(ctx, reinterpret_cast<const void*>(kernel), bytes);
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// leading dimension of every square work buffer: rows start 128-byte aligned
inline int64_t lead_dim(int64_t n) { return round_up(n, 16); }

}  // namespace pgp

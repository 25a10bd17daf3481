// tgp_ctx.cuh — handle, device workspace arena and host<->device staging shared by the ABI files.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/tgp_b200.h"

namespace tgp {

struct Arena {
    struct Block { char* p; size_t cap; };
    std::vector<Block> blocks;
    size_t used = 0;       // in the last block
    size_t total_req = 0;  // bytes requested since reset

    // At the start of a call: collapse to one block large enough for what the last call needed.
    cudaError_t reset() {
        cudaError_t rc = cudaSuccess;
        if (blocks.size() > 1) {
            size_t want = total_req + (total_req >> 3) + (1 << 20);
            for (auto& b : blocks) cudaFree(b.p);
            blocks.clear();
            char* p = nullptr;
            rc = cudaMalloc(&p, want);
            if (rc == cudaSuccess) blocks.push_back({p, want});
        }
        used = 0;
        total_req = 0;
        return rc;
    }
    void* alloc(size_t n) {
        n = (n + 255) & ~size_t(255);
        total_req += n;
        if (blocks.empty() || used + n > blocks.back().cap) {
            size_t cap = n > (size_t(8) << 20) ? n : (size_t(8) << 20);
            char* p = nullptr;
            if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
            blocks.push_back({p, cap});
            used = 0;
        }
        void* r = blocks.back().p + used;
        used += n;
        return r;
    }
    void release() {
        for (auto& b : blocks) cudaFree(b.p);
        blocks.clear();
    }
};

}  // namespace tgp

// State kept between the two phases of a time-sharded steady-state run (tgp_shard_phase1 / tgp_shard_phase2).
struct tgp_shard_state {
    bool active = false;
    int D = 0, rank = 0, world = 0;
    int64_t T = 0;
    bool fused_xchg = false;     // phase 1 shipped its record through the peer-memory exchange (phase 2 must wait there)
    const double* dy = nullptr;
    alignas(8) char work[256];   // tgp::SSWork<D>: device pointers of the run's workspace
};

// Persistent state of the one-launch steady-state logpdf (tgp_fir.cuh): the cached plan of the last model and a workspace that
// outlives calls (the tile words are tagged with a per-call epoch instead of being cleared).
struct tgp_fir_state {
    std::vector<unsigned char> key;      // bytes the cached plan was built from
    std::vector<unsigned char> plan;     // tgp::FirPlan<D>
    std::vector<double> upload;          // host copy of the device tables (kept alive for the asynchronous upload)
    int status = 1;
    long long bad_step = -1;
    double* dev = nullptr;   size_t dev_cap = 0;     // transient table + lane powers
    unsigned* counters = nullptr;
    double* partials = nullptr; size_t partials_cap = 0;
    double* result = nullptr;
    unsigned long long epoch = 0;
    void release() {
        if (dev) cudaFree(dev);
        if (counters) cudaFree(counters);
        if (partials) cudaFree(partials);
        if (result) cudaFree(result);
        dev = nullptr; counters = nullptr; partials = nullptr; result = nullptr;
        dev_cap = partials_cap = 0;
        key.clear();
    }
};

struct tgp_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;     // side stream for work that can overlap the main chain (created on first use)
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    tgp::Arena arena;
    std::string err;
    int64_t launches = 0, h2d = 0, d2h = 0;
    int algo = TGP_ALGO_AUTO;
    int chunk = 0;
    int dense_math = TGP_DENSE_F64;   // arithmetic of the large-state path (TGP_OPT_DENSE_MATH)
    double ss_tol = 1e-13;
    int64_t ss_prefix = 0;   // 0 = auto
    int sm_count = 148;
    double* pinned = nullptr;  // small pinned scratch for scalar results
    struct Pending { void* host; const void* dev; size_t bytes; size_t width, hpitch, dpitch, rows; };
    std::vector<Pending> pending;  // host outputs to copy back at the end of the call
    // optional per-kernel timing (TGP_OPT_TIMING): CUDA events around every launch on h->stream
    bool timing = false;
    struct Span { const char* name; cudaEvent_t t0, t1; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> ev_pool;
    const char* cur_name = nullptr;
    cudaEvent_t cur_t0 = nullptr;
    tgp_shard_state shard;
    const unsigned long long* deferred_res = nullptr;   // result block {err, lml, converged} of an un-synchronised call
    int64_t deferred_T = 0;
    // TGP_OPT_DEFER_STATUS: un-synchronised calls fold their status into a sticky device block {min failing step, #not converged}
    // instead of being resolved by the next call, so consecutive sharded calls queue back to back; tgp_synchronize reads it.
    bool defer_status = false;
    unsigned long long* sticky = nullptr;
    void* xchg = nullptr;                               // tgp::XchgState: peer-memory exchange of the time-sharded path (tgp_xchg.cu)
    tgp_fir_state fir;
    bool shard_overlap = false;                         // TGP_OPT_SHARD_OVERLAP
    int shard_world = 1;                                // world size of the last tgp_shard_logpdf
};

namespace tgp {

inline int fail(tgp_ctx* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define TGP_CUDA(h, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return tgp::fail(h, TGP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Input array -> device pointer (copy to the arena if it lives on the host).
inline int stage_in(tgp_ctx* h, const double* p, size_t n, const double** out) {
    if (!p) { *out = nullptr; return TGP_OK; }
    if (is_device_ptr(p)) { *out = p; return TGP_OK; }
    double* d = (double*)h->arena.alloc(n * sizeof(double));
    if (!d) return fail(h, TGP_ENOMEM, "device workspace allocation of %zu bytes failed", n * sizeof(double));
    TGP_CUDA(h, cudaMemcpyAsync(d, p, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->h2d += (int64_t)(n * sizeof(double));
    *out = d;
    return TGP_OK;
}

// Strided per-step input: `inner` doubles per step, stride (elements, may be 0) between steps.
inline int stage_steps(tgp_ctx* h, const double* p, int64_t stride, int64_t T, size_t inner, const double** out) {
    const size_t n = stride == 0 ? inner : (size_t)((T - 1) * (stride < 0 ? -stride : stride)) + inner;
    return stage_in(h, p, n, out);
}

// Output array: returns the device pointer (and device stride) kernels should write. Host
// destinations get a PACKED arena buffer and a pending (possibly strided) copy-back.
inline int stage_out(tgp_ctx* h, double* p, size_t inner, int64_t stride, int64_t T, double** out, int64_t* dstride) {
    *dstride = stride;
    if (!p) { *out = nullptr; return TGP_OK; }
    if (is_device_ptr(p)) { *out = p; return TGP_OK; }
    const bool contiguous = (size_t)stride == inner;
    const size_t n = inner * (size_t)T;
    double* d = (double*)h->arena.alloc(n * sizeof(double));
    if (!d) return fail(h, TGP_ENOMEM, "device workspace allocation of %zu bytes failed", n * sizeof(double));
    tgp_ctx::Pending pd;
    pd.host = p; pd.dev = d; pd.bytes = n * sizeof(double);
    pd.width = inner * sizeof(double); pd.hpitch = (size_t)stride * sizeof(double); pd.dpitch = inner * sizeof(double);
    pd.rows = contiguous ? 0 : (size_t)T;
    h->pending.push_back(pd);
    *out = d;
    *dstride = (int64_t)inner;
    return TGP_OK;
}

inline int flush_outputs(tgp_ctx* h) {
    for (auto& pd : h->pending) {
        if (pd.rows == 0) {
            TGP_CUDA(h, cudaMemcpyAsync(pd.host, pd.dev, pd.bytes, cudaMemcpyDeviceToHost, h->stream));
        } else {
            TGP_CUDA(h, cudaMemcpy2DAsync(pd.host, pd.hpitch, pd.dev, pd.dpitch, pd.width, pd.rows, cudaMemcpyDeviceToHost, h->stream));
        }
        h->d2h += (int64_t)pd.bytes;
    }
    h->pending.clear();
    return TGP_OK;
}

#define TGP_TRY(expr)                 \
    do {                              \
        int rc_ = (expr);             \
        if (rc_ != TGP_OK) return rc_; \
    } while (0)

inline cudaEvent_t prof_event(tgp_ctx* h) {
    cudaEvent_t e = nullptr;
    if (!h->ev_pool.empty()) { e = h->ev_pool.back(); h->ev_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
// before a launch: remember the kernel's name and (timing mode) record a start event
inline void prof_begin(tgp_ctx* h, const char* name) {
    h->cur_name = name;
    if (h->timing) { h->cur_t0 = prof_event(h); cudaEventRecord(h->cur_t0, h->stream); }
}
inline void prof_end(tgp_ctx* h) {
    if (h->timing && h->cur_t0) {
        cudaEvent_t t1 = prof_event(h);
        cudaEventRecord(t1, h->stream);
        h->spans.push_back({h->cur_name, h->cur_t0, t1});
        h->cur_t0 = nullptr;
    }
}
#define TGP_K(h, name) tgp::prof_begin(h, name)

#define TGP_LAUNCH_CHECK(h)                                                                     \
    do {                                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                    \
        if (e_ != cudaSuccess)                                                                  \
            return tgp::fail(h, TGP_ECUDA, "launch of %s failed: %s (%s:%d)", (h)->cur_name ? (h)->cur_name : "kernel", \
                             cudaGetErrorString(e_), __FILE__, __LINE__);                       \
        ++(h)->launches;                                                                        \
        tgp::prof_end(h);                                                                       \
    } while (0)

// Side stream: aux_begin makes it wait for everything enqueued on h->stream so far; aux_end makes h->stream wait for it.
inline int aux_begin(tgp_ctx* h) {
    if (!h->aux_stream) {
        TGP_CUDA(h, cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
        TGP_CUDA(h, cudaEventCreateWithFlags(&h->aux_fork, cudaEventDisableTiming));
        TGP_CUDA(h, cudaEventCreateWithFlags(&h->aux_join, cudaEventDisableTiming));
    }
    TGP_CUDA(h, cudaEventRecord(h->aux_fork, h->stream));
    TGP_CUDA(h, cudaStreamWaitEvent(h->aux_stream, h->aux_fork, 0));
    return TGP_OK;
}
inline int aux_end(tgp_ctx* h) {
    TGP_CUDA(h, cudaEventRecord(h->aux_join, h->aux_stream));
    TGP_CUDA(h, cudaStreamWaitEvent(h->stream, h->aux_join, 0));
    return TGP_OK;
}

template <class T>
inline int dalloc(tgp_ctx* h, size_t n, T** out) {
    *out = (T*)h->arena.alloc(n * sizeof(T));
    if (!*out) return fail(h, TGP_ENOMEM, "device workspace allocation of %zu bytes failed", n * sizeof(T));
    return TGP_OK;
}


// Scalar result (device) -> caller's pointer (host or device), stream-ordered.
inline int deliver_scalar(tgp_ctx* h, const double* dev, double* user) {
    if (!user) return TGP_OK;
    if (is_device_ptr(user)) {
        TGP_CUDA(h, cudaMemcpyAsync(user, dev, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    } else {
        tgp_ctx::Pending pd{};
        pd.host = user; pd.dev = dev; pd.bytes = sizeof(double); pd.rows = 0;
        h->pending.push_back(pd);
    }
    return TGP_OK;
}


// Request / result block of one filtering pass (shared by the general and steady-state drivers).
struct FilterReq {
    double* lml_out = nullptr;    // user pointer (host / device), nullable
    double* lml_steps = nullptr;  // device pointer in MEMORY order, nullable
    double* m_f = nullptr;  int64_t s_m = 0;   // device pointers / strides, nullable
    double* P_f = nullptr;  int64_t s_P = 0;
    bool keep_ws = false;         // keep the SoA filtering distributions (smoother)
    // produced
    double* ws = nullptr;         // (D + SymN) x T SoA, index = time
    double* xT = nullptr;         // packed final filtering distribution
    double* x0buf = nullptr;      // packed x0
    double* lml_dev = nullptr;
    unsigned long long* err = nullptr;
    bool packed_result = false;   // err points at a result block {u64 err_step, double lml, int converged}
};


}  // namespace tgp

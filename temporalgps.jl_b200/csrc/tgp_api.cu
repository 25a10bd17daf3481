// tgp_api.cu — the C ABI of libtgpb200.so (include/tgp_b200.h): argument checking, staging of
// host buffers, and the launch sequences of the small-state scan kernels. No CPU fallback exists:
// every entry point either runs the sm_100a kernels or returns an error code.
#include <algorithm>
#include <new>

#include "tgp_ctx.cuh"
#include "tgp_dispatch.h"

using namespace tgp;

namespace {

const char* g_create_err = "";
std::string g_create_err_buf;

// ---- argument checking (the reference throws at lgssm.jl:202-208 / MethodErrors) ---------------
int validate(tgp_ctx* h, const tgp_lgssm* m, bool need_y, const void* y) {
    if (!h) return TGP_EINVAL;
    if (!m) return fail(h, TGP_EINVAL, "model descriptor is NULL");
    if (m->D < 1 || m->D > TGP_MAX_D) return fail(h, TGP_EUNSUPPORTED, "latent dimension D=%d outside 1..%d", m->D, TGP_MAX_D);
    if (m->M > TGP_MAX_D) return fail(h, TGP_EUNSUPPORTED, "observation dimension M=%d outside 1..%d", m->M, TGP_MAX_D);
    if (m->M < 1) return fail(h, TGP_EINVAL, "observation dimension M=%d must be >= 1", m->M);
    if (m->T < 1) return fail(h, TGP_EINVAL, "Dimension mismatch. length(prior) is %lld", (long long)m->T);
    if (m->ordering != TGP_FORWARD && m->ordering != TGP_REVERSE) return fail(h, TGP_EINVAL, "ordering must be TGP_FORWARD or TGP_REVERSE");
    if (m->R_kind < TGP_R_SCALAR || m->R_kind > TGP_R_DENSE) return fail(h, TGP_EINVAL, "R_kind must be one of TGP_R_*");
    if (m->R_kind == TGP_R_SCALAR && m->M != 1) return fail(h, TGP_EINVAL, "TGP_R_SCALAR requires M == 1");
    if (!m->A || !m->a || !m->Q || !m->H || !m->h || !m->R || !m->m0 || !m->P0) return fail(h, TGP_EINVAL, "model descriptor has a NULL array");
    const int64_t D = m->D, M = m->M;
    const int64_t rin = m->R_kind == TGP_R_SCALAR ? 1 : (m->R_kind == TGP_R_DIAG ? M : M * M);
    struct { const char* name; int64_t s, inner; } chk[] = {{"sA", m->sA, D * D}, {"sa", m->sa, D},     {"sQ", m->sQ, D * D},
                                                              {"sH", m->sH, M * D}, {"sh", m->sh, M}, {"sR", m->sR, rin}};
    for (auto& c : chk)
        if (c.s != 0 && c.s < c.inner) return fail(h, TGP_EINVAL, "stride %s=%lld must be 0 or >= %lld", c.name, (long long)c.s, (long long)c.inner);
    if (need_y && !y) return fail(h, TGP_EINVAL, "Dimension mismatch. y is NULL but length(prior) is %lld", (long long)m->T);
    return TGP_OK;
}

// Shapes the small-state scan kernels are instantiated for; everything else takes the dense sequential path.
bool scan_shape(const tgp_lgssm* m) {
    if (m->M != 1 || m->R_kind != TGP_R_SCALAR) return false;
    switch (m->D) {
#define TGP_SHAPE_CASE(Dv) case Dv:
        TGP_FOR_EACH_D(TGP_SHAPE_CASE)
#undef TGP_SHAPE_CASE
        return true;
        default: return false;
    }
}

int require_scalar_obs(tgp_ctx* h, const tgp_lgssm* m) {
    if (m->M != 1 || m->R_kind != TGP_R_SCALAR)
        return fail(h, TGP_EUNSUPPORTED, "this build runs scalar observations (M == 1, TGP_R_SCALAR) on the scan path; got M=%d R_kind=%d", m->M, m->R_kind);
    return TGP_OK;
}

// Collect the status of a call that returned without synchronising (tgp_shard_phase2).
int resolve_deferred(tgp_ctx* h) {
    if (!h->deferred_res) return TGP_OK;
    const unsigned long long* res = h->deferred_res;
    h->deferred_res = nullptr;
    unsigned long long* perr = (unsigned long long*)h->pinned;
    int* pflag = (int*)(h->pinned + 2);
    TGP_CUDA(h, cudaMemcpyAsync(perr, res, 24, cudaMemcpyDeviceToHost, h->stream));
    h->d2h += 24;
    TGP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (*perr != ~0ull) return fail(h, TGP_ENOTPD, "covariance not positive definite at time index %lld (0-based)", (long long)*perr);
    if (*pflag == 0)
        return fail(h, TGP_EUNSUPPORTED, "the filtering covariance did not converge within the transient budget on this shard; "
                                         "use the general sharded path (tgp_shard_reduce / tgp_shard_prefix)");
    return TGP_OK;
}

int begin_call(tgp_ctx* h) {
    h->err.clear();
    h->pending.clear();
    TGP_CUDA(h, cudaSetDevice(h->device));
    TGP_TRY(resolve_deferred(h));
    if (h->aux_stream) TGP_CUDA(h, cudaStreamSynchronize(h->aux_stream));   // nothing of an earlier (failed) call may still use the arena
    TGP_CUDA(h, h->arena.reset());
    return TGP_OK;
}

#define TGP_CASE_D(Dv) case Dv: return CALL(Dv);
#define TGP_DISPATCH_D(h, Dv)                                                                               \
    switch (Dv) {                                                                                           \
        TGP_FOR_EACH_D(TGP_CASE_D)                                                                          \
        default: return fail(h, TGP_EUNSUPPORTED, "latent dimension D=%d has no kernel instantiation in this build", Dv); \
    }

}  // namespace

// ================================================================================================
extern "C" {

const char* tgp_version(void) { return "tgp-b200 0.1.0 (sm_100a)"; }

int tgp_create(tgp_handle* out, int device) {
    if (!out) return TGP_EINVAL;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err_buf = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                           " (libtgpb200 has no CPU path)";
        g_create_err = g_create_err_buf.c_str();
        cudaGetLastError();
        return TGP_ECUDA;
    }
    if (device < 0 || device >= n) {
        g_create_err_buf = "device index out of range";
        g_create_err = g_create_err_buf.c_str();
        return TGP_EINVAL;
    }
    tgp_ctx* h = new (std::nothrow) tgp_ctx();
    if (!h) return TGP_ENOMEM;
    h->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMallocHost((void**)&h->pinned, 4096)) != cudaSuccess) {
        g_create_err_buf = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
        g_create_err = g_create_err_buf.c_str();
        delete h;
        return TGP_ECUDA;
    }
    h->stream = h->own_stream;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = h;
    return TGP_OK;
}

void tgp_destroy(tgp_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    xchg_destroy(h);
    dense_release(h);
    h->fir.release();
    h->arena.release();
    for (auto& sp : h->spans) { cudaEventDestroy(sp.t0); cudaEventDestroy(sp.t1); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->sticky) cudaFree(h->sticky);
    if (h->aux_stream) { cudaStreamDestroy(h->aux_stream); cudaEventDestroy(h->aux_fork); cudaEventDestroy(h->aux_join); }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

const char* tgp_last_error(tgp_handle h) { return h ? h->err.c_str() : g_create_err; }

int tgp_set_option(tgp_handle h, int option, int64_t value) {
    if (!h) return TGP_EINVAL;
    switch (option) {
        case TGP_OPT_ALGO:
            if (value != TGP_ALGO_AUTO && value != TGP_ALGO_SCAN) return fail(h, TGP_EINVAL, "unknown algorithm %lld", (long long)value);
            h->algo = (int)value;
            return TGP_OK;
        case TGP_OPT_CHUNK:
            if (value < 0 || value > 4096) return fail(h, TGP_EINVAL, "chunk must be in 0..4096");
            h->chunk = (int)value;
            return TGP_OK;
        case TGP_OPT_SS_TOL: {
            double v;
            memcpy(&v, &value, sizeof v);
            if (!(v >= 0.0) || !(v < 1e-3)) return fail(h, TGP_EINVAL, "steady-state tolerance must be in [0, 1e-3)");
            h->ss_tol = v;
            return TGP_OK;
        }
        case TGP_OPT_TIMING:
            h->timing = value != 0;
            if (!h->timing) {
                cudaStreamSynchronize(h->stream);
                for (auto& sp : h->spans) { h->ev_pool.push_back(sp.t0); h->ev_pool.push_back(sp.t1); }
                h->spans.clear();
            }
            return TGP_OK;
        case TGP_OPT_SS_PREFIX:
            if (value < 0 || value > (int64_t(1) << 24)) return fail(h, TGP_EINVAL, "steady-state prefix must be in 0..2^24");
            h->ss_prefix = value;
            return TGP_OK;
        case TGP_OPT_SHARD_OVERLAP:
            h->shard_overlap = value != 0;
            return TGP_OK;
        case TGP_OPT_DEFER_STATUS:
            h->defer_status = value != 0;
            if (h->defer_status && !h->sticky) {
                const unsigned long long init[2] = {~0ull, 0ull};
                TGP_CUDA(h, cudaSetDevice(h->device));
                TGP_CUDA(h, cudaMalloc((void**)&h->sticky, sizeof init));
                TGP_CUDA(h, cudaMemcpy(h->sticky, init, sizeof init, cudaMemcpyHostToDevice));
            }
            return TGP_OK;
        case TGP_OPT_DENSE_MATH:
            if (value != TGP_DENSE_F64 && value != TGP_DENSE_TF32X3) return fail(h, TGP_EINVAL, "unknown dense arithmetic %lld", (long long)value);
            h->dense_math = (int)value;
            return TGP_OK;
        default: return fail(h, TGP_EINVAL, "unknown option %d", option);
    }
}

int tgp_get_counters(tgp_handle h, int64_t* launches, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!h) return TGP_EINVAL;
    if (launches) *launches = h->launches;
    if (h2d_bytes) *h2d_bytes = h->h2d;
    if (d2h_bytes) *d2h_bytes = h->d2h;
    return TGP_OK;
}

int tgp_get_timing(tgp_handle h, int cap, const char** names, double* total_ms, int64_t* calls) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    struct Acc { const char* name; double ms; int64_t n; };
    std::vector<Acc> acc;
    for (auto& sp : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.t0, sp.t1) != cudaSuccess) { cudaGetLastError(); continue; }
        bool found = false;
        for (auto& a : acc)
            if (!strcmp(a.name, sp.name)) { a.ms += ms; ++a.n; found = true; break; }
        if (!found) acc.push_back({sp.name, (double)ms, 1});
    }
    std::sort(acc.begin(), acc.end(), [](const Acc& x, const Acc& y) { return x.ms > y.ms; });
    for (int i = 0; i < (int)acc.size() && i < cap; ++i) {
        if (names) names[i] = acc[i].name;
        if (total_ms) total_ms[i] = acc[i].ms;
        if (calls) calls[i] = acc[i].n;
    }
    return (int)acc.size();
}

int tgp_set_stream(tgp_handle h, void* cuda_stream) {
    if (!h) return TGP_EINVAL;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return TGP_OK;
}

int tgp_logpdf(tgp_handle h, const tgp_lgssm* model, const double* y, double* lml_out, double* lml_per_step) {
    TGP_TRY(validate(h, model, true, y));
    TGP_TRY(begin_call(h));
    if (!scan_shape(model)) return dense_filter(h, model, y, lml_out, lml_per_step, nullptr, 0, nullptr, 0);
#define CALL(Dv) do_filter<Dv>(h, model, y, nullptr, 0, nullptr, 0, lml_out, lml_per_step)
    TGP_DISPATCH_D(h, model->D)
#undef CALL
}

int tgp_filter(tgp_handle h, const tgp_lgssm* model, const double* y, double* m_f, int64_t s_m, double* P_f, int64_t s_P,
               double* lml_out) {
    TGP_TRY(validate(h, model, true, y));
    if (m_f && s_m < model->D) return fail(h, TGP_EINVAL, "s_m=%lld must be >= D", (long long)s_m);
    if (P_f && s_P < (int64_t)model->D * model->D) return fail(h, TGP_EINVAL, "s_P=%lld must be >= D*D", (long long)s_P);
    TGP_TRY(begin_call(h));
    if (!scan_shape(model)) return dense_filter(h, model, y, lml_out, nullptr, m_f, s_m, P_f, s_P);
#define CALL(Dv) do_filter<Dv>(h, model, y, m_f, s_m, P_f, s_P, lml_out, nullptr)
    TGP_DISPATCH_D(h, model->D)
#undef CALL
}

int tgp_posterior(tgp_handle h, const tgp_lgssm* model, const double* y, double* G, double* g, double* Sig, double* m_T,
                  double* P_T) {
    TGP_TRY(validate(h, model, true, y));
    TGP_TRY(begin_call(h));
    if (!scan_shape(model) || model->ordering != TGP_FORWARD) return seq_posterior(h, model, y, G, g, Sig, m_T, P_T, nullptr);
#define CALL(Dv) do_posterior<Dv>(h, model, y, G, g, Sig, m_T, P_T)
    TGP_DISPATCH_D(h, model->D)
#undef CALL
}

int tgp_marginals(tgp_handle h, const tgp_lgssm* model, double* mean_out, double* cov_out) {
    TGP_TRY(validate(h, model, false, nullptr));
    if (!mean_out || !cov_out) return fail(h, TGP_EINVAL, "mean_out and cov_out must be non-NULL");
    TGP_TRY(begin_call(h));
    if (!scan_shape(model)) return seq_marginals(h, model, mean_out, cov_out, 0);
#define CALL(Dv) do_marginals<Dv>(h, model, mean_out, cov_out)
    TGP_DISPATCH_D(h, model->D)
#undef CALL
}

int tgp_marginals_diag(tgp_handle h, const tgp_lgssm* model, double* mean_out, double* var_out) {
    TGP_TRY(validate(h, model, false, nullptr));
    if (!mean_out || !var_out) return fail(h, TGP_EINVAL, "mean_out and var_out must be non-NULL");
    TGP_TRY(begin_call(h));
    if (!scan_shape(model)) return seq_marginals(h, model, mean_out, var_out, 1);
#define CALL(Dv) do_marginals<Dv>(h, model, mean_out, var_out)      /* M == 1: the marginal covariance IS its diagonal */
    TGP_DISPATCH_D(h, model->D)
#undef CALL
}

int tgp_lti_components(tgp_handle h, int32_t D, int64_t T, const double* F, const double* F0, const double* P, const double* t,
                       double* A_out, double* Q_out) {
    if (!h) return TGP_EINVAL;
    if (D < 1 || T < 1) return fail(h, TGP_EINVAL, "D and T must be >= 1 (got D=%d, T=%lld)", D, (long long)T);
    if (!F || !P || !t || !A_out || !Q_out) return fail(h, TGP_EINVAL, "F, P, t, A_out and Q_out must be non-NULL");
    if (is_device_ptr(F) || is_device_ptr(P) || (F0 && is_device_ptr(F0)))
        return fail(h, TGP_EINVAL, "F, F0 and P are read on the host (D x D doubles each)");
    TGP_TRY(begin_call(h));
    return lti_components(h, D, T, F, F0, P, t, A_out, Q_out);
}

int tgp_posterior_marginals(tgp_handle h, const tgp_lgssm* model, const double* y, const double* R_new, int64_t sRnew,
                            double* mean_out, double* var_out, double* lml_out) {
    TGP_TRY(validate(h, model, true, y));
    if (!R_new || !mean_out || !var_out) return fail(h, TGP_EINVAL, "R_new, mean_out and var_out must be non-NULL");
    if (sRnew != 0 && sRnew < 1) return fail(h, TGP_EINVAL, "sRnew must be 0 or >= 1");
    TGP_TRY(begin_call(h));
    if (!scan_shape(model) || model->ordering != TGP_FORWARD)
        return seq_posterior_marginals(h, model, y, R_new, sRnew, mean_out, var_out, lml_out);
#define CALL(Dv) do_posterior_marginals<Dv>(h, model, y, R_new, sRnew, mean_out, var_out, lml_out)
    TGP_DISPATCH_D(h, model->D)
#undef CALL
}

int tgp_debug_tc_gemm(tgp_handle h, int K, int Mx, int N, const float* X, const float* Y, float* C, int symmetric) {
    if (!h || !X || !Y || !C || K < 1 || Mx < 1 || N < 1) return TGP_EINVAL;
    TGP_TRY(begin_call(h));
    return tc_gemm_selftest(h, K, Mx, N, X, Y, C, symmetric);
}

int tgp_elem_size(int D) { return 3 * D * D + 2 * D; }

int tgp_shard_reduce(tgp_handle h, const tgp_lgssm* shard, const double* y, double* elem_out) {
    TGP_TRY(validate(h, shard, true, y));
    TGP_TRY(require_scalar_obs(h, shard));
    if (!elem_out) return fail(h, TGP_EINVAL, "elem_out must be non-NULL");
    TGP_TRY(begin_call(h));
#define CALL(Dv) do_shard_reduce<Dv>(h, shard, y, elem_out)
    TGP_DISPATCH_D(h, shard->D)
#undef CALL
}

int tgp_shard_phase1(tgp_handle h, const tgp_lgssm* shard, const double* y, int rank, int world, double* xchg_out) {
    TGP_TRY(validate(h, shard, true, y));
    TGP_TRY(require_scalar_obs(h, shard));
    if (rank < 0 || world < 1 || rank >= world || !xchg_out) return fail(h, TGP_EINVAL, "bad rank / world / xchg_out");
    TGP_TRY(begin_call(h));
#define CALL(Dv) do_shard_phase1<Dv>(h, shard, y, rank, world, xchg_out)
    TGP_DISPATCH_D(h, shard->D)
#undef CALL
}

int tgp_shard_logpdf(tgp_handle h, const tgp_lgssm* shard, const double* y, int rank, int world) {
    TGP_TRY(validate(h, shard, true, y));
    TGP_TRY(require_scalar_obs(h, shard));
    if (rank < 0 || world < 1 || rank >= world) return fail(h, TGP_EINVAL, "bad rank / world");
    TGP_TRY(begin_call(h));
    bool handled = false;
    int rc = TGP_EUNSUPPORTED;
    switch (shard->D) {
#define TGP_SL_CASE(Dv) case Dv: rc = do_shard_logpdf<Dv>(h, shard, y, rank, world, &handled); break;
        TGP_FOR_EACH_D(TGP_SL_CASE)
#undef TGP_SL_CASE
        default: break;
    }
    if (rc != TGP_OK) return rc;
    if (!handled)
        return fail(h, TGP_EUNSUPPORTED, "tgp_shard_logpdf runs Forward, time-invariant, scalar-observation models with D <= 4 whose filter reaches its "
                                         "fixed point and forgets its start within 3072 steps; use tgp_shard_phase1/2 or tgp_shard_reduce/prefix");
    h->shard_world = world;
    return TGP_OK;
}

int tgp_shard_partial(tgp_handle h, double* lml_shard) {
    if (!h || !lml_shard) return TGP_EINVAL;
    TGP_CUDA(h, cudaSetDevice(h->device));
    if (!h->fir.result || h->fir.epoch == 0) return fail(h, TGP_EINVAL, "tgp_shard_partial without a tgp_shard_logpdf");
    const double* src = h->fir.result + 4 * (h->fir.epoch & 1ull);
    if (is_device_ptr(lml_shard)) {
        TGP_CUDA(h, cudaMemcpyAsync(lml_shard, src, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    } else {
        TGP_CUDA(h, cudaMemcpyAsync(h->pinned + 8, src, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        h->d2h += 8;
        *lml_shard = h->pinned[8];
    }
    return TGP_OK;
}

int tgp_shard_result(tgp_handle h, double* lml_total) {
    if (!h || !lml_total) return TGP_EINVAL;
    TGP_CUDA(h, cudaSetDevice(h->device));
    if (!h->fir.result || h->fir.epoch == 0) return fail(h, TGP_EINVAL, "tgp_shard_result without a tgp_shard_logpdf");
    const bool dev = is_device_ptr(lml_total);
    double* dst = dev ? lml_total : h->fir.result + 2;
    if (h->shard_world > 1) {
        XchgFirView v;
        if (!xchg_fir_view(h, &v)) return fail(h, TGP_EINVAL, "exchange not opened");
        TGP_TRY(xchg_fir_total(h, *v.epoch, dst));
    } else {
        TGP_CUDA(h, cudaMemcpyAsync(dst, h->fir.result + 4 * (h->fir.epoch & 1ull), sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    if (!dev) {
        TGP_CUDA(h, cudaMemcpyAsync(h->pinned + 8, dst, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        h->d2h += 8;
        *lml_total = h->pinned[8];
    }
    return TGP_OK;
}

int tgp_shard_phase2(tgp_handle h, const double* xchg_all, double* lml_partial) {
    if (!h) return TGP_EINVAL;
    if (!xchg_all || !lml_partial) return fail(h, TGP_EINVAL, "xchg_all and lml_partial must be non-NULL");
    TGP_CUDA(h, cudaSetDevice(h->device));
#define CALL(Dv) do_shard_phase2<Dv>(h, xchg_all, lml_partial)
    TGP_DISPATCH_D(h, h->shard.D)
#undef CALL
}

int tgp_shard_xchg_size(int D) { return D * D + D; }

int tgp_xchg_create(tgp_handle h, int rank, int world, int slot_doubles, void* ipc_handle_out) {
    if (!h || !ipc_handle_out) return TGP_EINVAL;
    TGP_CUDA(h, cudaSetDevice(h->device));
    return xchg_create(h, rank, world, slot_doubles, ipc_handle_out);
}
int tgp_xchg_open(tgp_handle h, const void* ipc_handles_all) {
    if (!h || !ipc_handles_all) return TGP_EINVAL;
    TGP_CUDA(h, cudaSetDevice(h->device));
    return xchg_open(h, ipc_handles_all);
}

int tgp_synchronize(tgp_handle h) {
    if (!h) return TGP_EINVAL;
    TGP_CUDA(h, cudaSetDevice(h->device));
    if (h->deferred_res) return resolve_deferred(h);
    if (h->sticky) {    // status accumulated by un-synchronised calls (TGP_OPT_DEFER_STATUS): report and clear
        unsigned long long* p = (unsigned long long*)(h->pinned + 32);
        const unsigned long long init[2] = {~0ull, 0ull};
        TGP_CUDA(h, cudaMemcpyAsync(p, h->sticky, 16, cudaMemcpyDeviceToHost, h->stream));
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        h->d2h += 16;
        const unsigned long long err = p[0], notconv = p[1];
        if (err != ~0ull || notconv) TGP_CUDA(h, cudaMemcpy(h->sticky, init, sizeof init, cudaMemcpyHostToDevice));
        if (err != ~0ull) return fail(h, TGP_ENOTPD, "covariance not positive definite at time index %lld (0-based) of a shard", (long long)err);
        if (notconv)
            return fail(h, TGP_EUNSUPPORTED, "the filtering covariance did not converge within the transient budget on this shard in %llu call(s); "
                                             "use the general sharded path (tgp_shard_reduce / tgp_shard_prefix)", notconv);
        return TGP_OK;
    }
    TGP_CUDA(h, cudaStreamSynchronize(h->stream));
    return TGP_OK;
}

int tgp_shard_prefix(tgp_handle h, int D, int n_elems, const double* elems, const double* m0, const double* P0, double* m_in,
                     double* P_in) {
    // pure host arithmetic (no device work): h may be NULL
    if (n_elems < 0 || (n_elems > 0 && !elems) || !m0 || !P0 || !m_in || !P_in) return fail(h, TGP_EINVAL, "bad argument to tgp_shard_prefix");
#define CALL(Dv) do_shard_prefix<Dv>(n_elems, elems, m0, P0, m_in, P_in)
    TGP_DISPATCH_D(h, D)
#undef CALL
}

}  // extern "C"

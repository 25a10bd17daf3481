// tgp_drivers.cuh — launch sequences of the small-state scan kernels, templated on the latent
// dimension D. Included only by tgp_inst.cu (one translation unit per D, -DTGP_D=<D>) so the
// instantiations compile in parallel; tgp_api.cu sees the declarations in tgp_dispatch.h.
#pragma once
#include "tgp_ctx.cuh"
#include "tgp_dispatch.h"
#include "tgp_scan_small.cuh"
#include "tgp_steady.cuh"
#include "tgp_fir.cuh"
#include "tgp_steady_smooth.cuh"

namespace tgp {

// Model + observations -> device-resident descriptor.
inline int stage_model(tgp_ctx* h, const tgp_lgssm* m, const double* y, tgp_lgssm* d, const double** dy) {
    *d = *m;
    const size_t D = m->D, M = m->M;
    const size_t rin = m->R_kind == TGP_R_SCALAR ? 1 : (m->R_kind == TGP_R_DIAG ? M : M * M);
    TGP_TRY(stage_steps(h, m->A, m->sA, m->T, D * D, &d->A));
    TGP_TRY(stage_steps(h, m->a, m->sa, m->T, D, &d->a));
    TGP_TRY(stage_steps(h, m->Q, m->sQ, m->T, D * D, &d->Q));
    TGP_TRY(stage_steps(h, m->H, m->sH, m->T, M * D, &d->H));
    TGP_TRY(stage_steps(h, m->h, m->sh, m->T, M, &d->h));
    TGP_TRY(stage_steps(h, m->R, m->sR, m->T, rin, &d->R));
    TGP_TRY(stage_in(h, m->m0, D, &d->m0));
    TGP_TRY(stage_in(h, m->P0, D * D, &d->P0));
    if (dy) TGP_TRY(stage_in(h, y, (size_t)m->T * M, dy));
    return TGP_OK;
}

inline bool time_invariant(const tgp_lgssm& m) { return !(m.sA | m.sa | m.sQ | m.sH | m.sh | m.sR); }

// Folds the result block {err step, lml, converged} of an un-synchronised call into the handle's sticky status.
static __global__ void k_sticky_status(const unsigned long long* __restrict__ res, unsigned long long* __restrict__ sticky) {
    if (res[0] < sticky[0]) sticky[0] = res[0];
    if (*reinterpret_cast<const int*>(res + 2) == 0) sticky[1] += 1ull;
}

// End of a call: copy back host outputs, fetch the failing-step word (and the steady-state
// convergence word, if any), wait. *ss_converged is left untouched when ss_flag is NULL.
inline int end_call(tgp_ctx* h, const unsigned long long* err_step, int64_t T, bool reverse_t, const int* ss_flag = nullptr,
                    bool* ss_converged = nullptr, const FilterReq* packed = nullptr) {
    unsigned long long* perr = (unsigned long long*)h->pinned;
    double* plml = h->pinned + 1;
    int* pflag = (int*)(h->pinned + 2);
    *perr = ~0ull;
    *pflag = 1;
    if (packed && packed->packed_result) {   // {err, lml, flag} contiguous on the device: one copy
        TGP_CUDA(h, cudaMemcpyAsync(perr, packed->err, 24, cudaMemcpyDeviceToHost, h->stream));
        h->d2h += 24;
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        if (ss_converged) *ss_converged = *pflag != 0;
        if (*pflag == 0 && *perr == ~0ull) return TGP_OK;
        if (packed->lml_out && !is_device_ptr(packed->lml_out)) *packed->lml_out = *plml;
        if (!h->pending.empty()) {
            TGP_TRY(flush_outputs(h));
            TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        }
    } else {
        if (ss_flag) {
            TGP_CUDA(h, cudaMemcpyAsync(pflag, ss_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            h->d2h += 4;
        }
        if (err_step) {
            TGP_CUDA(h, cudaMemcpyAsync(perr, err_step, sizeof(*perr), cudaMemcpyDeviceToHost, h->stream));
            h->d2h += 8;
        }
        if (ss_flag) {  // outputs are only valid if the steady-state test passed: look before copying back
            TGP_CUDA(h, cudaStreamSynchronize(h->stream));
            if (ss_converged) *ss_converged = *pflag != 0;
            if (*pflag == 0 && *perr == ~0ull) return TGP_OK;
        }
        TGP_TRY(flush_outputs(h));
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    if (*perr != ~0ull) {
        const long long n = (long long)*perr;
        const long long t = reverse_t ? (long long)T - 1 - n : n;
        return fail(h, TGP_ENOTPD, "covariance not positive definite at time index %lld (0-based)", t);
    }
    return TGP_OK;
}

// ---- general (time-varying) filter: reduce -> mid -> apply ----------------------------------------
template <int D>
int filter_general(tgp_ctx* h, const tgp_lgssm& d, const double* dy, FilterReq& rq) {
    constexpr int SN = D + Sym<D>::N;
    const int64_t T = d.T;
    const bool rev = d.ordering == TGP_REVERSE;
    cudaStream_t st = h->stream;
    TGP_TRY(dalloc(h, 1, &rq.err));
    TGP_CUDA(h, cudaMemsetAsync(rq.err, 0xFF, sizeof(unsigned long long), st));
    TGP_TRY(dalloc(h, SN, &rq.x0buf));
    TGP_TRY(dalloc(h, SN, &rq.xT));
    TGP_TRY(dalloc(h, 1, &rq.lml_dev));
    double* lml_extra = nullptr;
    TGP_TRY(dalloc(h, 1, &lml_extra));
    if (rq.keep_ws) TGP_TRY(dalloc(h, (size_t)SN * T, &rq.ws));

    // x0 (and, for Reverse, the leading update of the last memory index; lgssm.jl:161-165)
    const int64_t tl = T - 1;
    TGP_K(h, "k_init_state");
    k_init_state<D><<<1, 32, 0, st>>>(d.m0, d.P0, rq.x0buf, rev ? 1 : 0, d.H + tl * d.sH, d.h + tl * d.sh, d.R + tl * d.sR,
                                      dy + tl, lml_extra, rq.lml_steps ? rq.lml_steps + tl : nullptr,
                                      rq.m_f ? rq.m_f + tl * rq.s_m : nullptr, rq.P_f ? rq.P_f + tl * rq.s_P : nullptr,
                                      rq.ws, T, tl, rq.err);
    TGP_LAUNCH_CHECK(h);

    const int64_t Ts = rev ? T - 1 : T;  // scan steps: (predict, update) pairs
    if (Ts == 0) {
        TGP_CUDA(h, cudaMemcpyAsync(rq.xT, rq.x0buf, SN * sizeof(double), cudaMemcpyDeviceToDevice, st));
        TGP_CUDA(h, cudaMemcpyAsync(rq.lml_dev, lml_extra, sizeof(double), cudaMemcpyDeviceToDevice, st));
        return deliver_scalar(h, rq.lml_dev, rq.lml_out);
    }
    DevModel dm;
    if (!rev) {
        dm = DevModel{d.A, d.a, d.Q, d.H, d.h, d.R, d.sA, d.sa, d.sQ, d.sH, d.sh, d.sR, dy, 1, Ts};
    } else {  // scan step j: transition of memory index T-1-j, emission / observation of index T-2-j
        dm = DevModel{d.A + tl * d.sA, d.a + tl * d.sa, d.Q + tl * d.sQ, d.H + (tl - 1) * d.sH, d.h + (tl - 1) * d.sh,
                      d.R + (tl - 1) * d.sR, -d.sA, -d.sa, -d.sQ, -d.sH, -d.sh, -d.sR, dy + (tl - 1), -1, Ts};
    }
    const bool tv = !time_invariant(d);
    ConstModel<D>* cm = nullptr;
    if (!tv) {
        TGP_TRY(dalloc(h, 1, &cm));
        TGP_K(h, "k_const_model");
        k_const_model<D><<<1, 32, 0, st>>>(dm, cm);
        TGP_LAUNCH_CHECK(h);
    }
    const int L = h->chunk > 0 ? h->chunk : 16;
    const int64_t nchunk = (Ts + L - 1) / L;
    const int64_t grid = (nchunk + kBlock - 1) / kBlock;
    const int64_t nthreads = grid * kBlock, nwarps = nthreads / 32;
    double *excl, *wagg, *wstate, *partials;
    TGP_TRY(dalloc(h, (size_t)Elem<D>::N * nthreads, &excl));
    TGP_TRY(dalloc(h, (size_t)Elem<D>::N * nwarps, &wagg));
    TGP_TRY(dalloc(h, (size_t)SN * nwarps, &wstate));
    TGP_TRY(dalloc(h, (size_t)grid, &partials));
    TGP_K(h, "k_filter_reduce");
    if (tv) k_filter_reduce<D, true><<<(unsigned)grid, kBlock, 0, st>>>(dm, cm, L, nthreads, excl, wagg, nwarps);
    else    k_filter_reduce<D, false><<<(unsigned)grid, kBlock, 0, st>>>(dm, cm, L, nthreads, excl, wagg, nwarps);
    TGP_LAUNCH_CHECK(h);
    if (nwarps > 4 * kMidThreads) {     // many aggregates: scan them in groups of 32 on the whole GPU, chain only the group aggregates
        const int64_t ngroups = (nwarps + 31) / 32;
        double *gexcl, *gagg, *gstate;
        TGP_TRY(dalloc(h, (size_t)Elem<D>::N * nwarps, &gexcl));
        TGP_TRY(dalloc(h, (size_t)Elem<D>::N * ngroups, &gagg));
        TGP_TRY(dalloc(h, (size_t)SN * ngroups, &gstate));
        const unsigned mg = (unsigned)((ngroups * 32 + kBlock - 1) / kBlock);
        TGP_K(h, "k_mid_scan");
        k_mid_scan<D><<<mg, kBlock, 0, st>>>(wagg, nwarps, gexcl, gagg, ngroups);
        TGP_LAUNCH_CHECK(h);
        TGP_K(h, "k_filter_mid");
        k_filter_mid<D><<<1, kMidThreads, 0, st>>>(gagg, ngroups, rq.x0buf, gstate, rq.xT);
        TGP_LAUNCH_CHECK(h);
        TGP_K(h, "k_mid_apply");
        k_mid_apply<D><<<(unsigned)((nwarps + kBlock - 1) / kBlock), kBlock, 0, st>>>(gexcl, gstate, nwarps, ngroups, wstate);
        TGP_LAUNCH_CHECK(h);
    } else {
        TGP_K(h, "k_filter_mid");
        k_filter_mid<D><<<1, kMidThreads, 0, st>>>(wagg, nwarps, rq.x0buf, wstate, rq.xT);
        TGP_LAUNCH_CHECK(h);
    }
    FilterOut fo;
    const int64_t o0 = rev ? tl - 1 : 0;  // memory index of scan step 0
    const int64_t sg = rev ? -1 : 1;
    fo.lml_steps = rq.lml_steps ? rq.lml_steps + o0 : nullptr;
    fo.s_l = sg;
    fo.m_f = rq.m_f ? rq.m_f + o0 * rq.s_m : nullptr;
    fo.s_m = sg * rq.s_m;
    fo.P_f = rq.P_f ? rq.P_f + o0 * rq.s_P : nullptr;
    fo.s_P = sg * rq.s_P;
    fo.ws_m = rq.ws;   // forward only (checked by callers): index = scan step = time, stride Ts == T
    fo.partials = partials;
    fo.err_step = rq.err;
    TGP_K(h, "k_filter_apply");
    if (tv) k_filter_apply<D, true><<<(unsigned)grid, kBlock, 0, st>>>(dm, cm, L, nthreads, excl, wstate, nwarps, fo);
    else    k_filter_apply<D, false><<<(unsigned)grid, kBlock, 0, st>>>(dm, cm, L, nthreads, excl, wstate, nwarps, fo);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_sum_partials");
    k_sum_partials<<<1, 256, 0, st>>>(partials, grid, lml_extra, rq.lml_dev);
    TGP_LAUNCH_CHECK(h);
    return deliver_scalar(h, rq.lml_dev, rq.lml_out);
}

// Reverse scan steps are offset by one against memory (the leading update is step "-1"):
// err_step holds a scan step; map it to a memory index for the message.
inline int64_t err_T(const tgp_lgssm& d) { return d.ordering == TGP_REVERSE ? d.T - 1 : d.T; }

constexpr int64_t kSmMaxEnd = 8192;      // budget (steps) for each covariance recursion of tgp_steady_smooth.cuh to stop moving

// ---- logpdf of a time-invariant model whose D has no register-resident steady kernel (D = 8, 10): the head by the sequential
// single-CTA filter, the rest by ONE constant-coefficient forward scan over vectors (tgp_steady_smooth.cuh). Log-likelihood only.
template <int D>
int logpdf_steady_vec(tgp_ctx* h, const tgp_lgssm& d, const double* dy, double* lml_out, bool* converged) {
    const int64_t T = d.T;
    cudaStream_t st = h->stream;
    *converged = false;
    SmConst<D>* cst;
    unsigned long long* err;
    TGP_TRY(dalloc(h, 1, &cst));
    TGP_TRY(dalloc(h, 1, &err));
    TGP_CUDA(h, cudaMemsetAsync(err, 0xFF, sizeof(unsigned long long), st));
    const DevModel dm{d.A, d.a, d.Q, d.H, d.h, d.R, 0, 0, 0, 0, 0, 0, dy, 1, T};
    TGP_K(h, "k_sm_head_fwd");
    k_sm_head_fwd<D><<<1, 128, 0, st>>>(dm, d.m0, d.P0, kSmMaxEnd, h->ss_tol, nullptr, nullptr, cst, err);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_sm_setup");
    k_sm_setup<D><<<1, 128, 0, st>>>(dm, cst);
    TGP_LAUNCH_CHECK(h);
    long long* pN0 = (long long*)(h->pinned + 24);
    int* pconv = (int*)(h->pinned + 25);
    TGP_CUDA(h, cudaMemcpyAsync(pN0, &cst->N0, sizeof(long long), cudaMemcpyDeviceToHost, st));
    TGP_CUDA(h, cudaMemcpyAsync(pconv, &cst->conv_f, sizeof(int), cudaMemcpyDeviceToHost, st));
    TGP_CUDA(h, cudaStreamSynchronize(st));
    h->d2h += 12;
    const int64_t N0 = *pN0;
    if (!*pconv || T < N0 + 2 * kCsL) return TGP_OK;
    const int64_t nf = T - N0;
    double *partials, *lml_dev;
    const int64_t nblk = ((nf + kCsL - 1) / kCsL + kCsThreads - 1) / kCsThreads;
    TGP_TRY(dalloc(h, (size_t)nblk, &partials));
    TGP_TRY(dalloc(h, 1, &lml_dev));
    FwdItems<D> fi{dy, N0, nullptr, partials};
    BwdItems<D> bi{};
    TGP_TRY((cs_scan<D, true>(h, cst, fi, bi, nf, cst->mstart, nullptr)));
    TGP_K(h, "k_sm_lml");
    k_sm_lml<D><<<1, 256, 0, st>>>(cst, partials, nblk, nf, lml_dev);
    TGP_LAUNCH_CHECK(h);
    TGP_TRY(deliver_scalar(h, lml_dev, lml_out));
    return end_call(h, err, T, false, &cst->conv_f, converged);
}

template <int D>
int do_filter(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* m_f, int64_t s_m, double* P_f, int64_t s_P,
              double* lml_out, double* lml_steps) {
    if constexpr (D <= kFirMaxD) {   // log-likelihood only, time-invariant: one launch, one pass over y (tgp_fir.cuh)
        if (h->algo == TGP_ALGO_AUTO && m->ordering == TGP_FORWARD && time_invariant(*m) && !m_f && !P_f && !lml_steps) {
            bool handled = false;
            TGP_TRY(logpdf_fir<D>(h, m, y, lml_out, &handled));
            if (handled) return TGP_OK;
            h->pending.clear();
            TGP_CUDA(h, h->arena.reset());
        }
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        tgp_lgssm d;
        const double* dy;
        TGP_TRY(stage_model(h, m, y, &d, &dy));
        FilterReq rq;
        rq.lml_out = lml_out;
        int64_t ds;
        TGP_TRY(stage_out(h, lml_steps, 1, 1, m->T, &rq.lml_steps, &ds));
        TGP_TRY(stage_out(h, m_f, D, s_m, m->T, &rq.m_f, &rq.s_m));
        TGP_TRY(stage_out(h, P_f, D * D, s_P, m->T, &rq.P_f, &rq.s_P));
        bool handled = false;
        const int* flag = nullptr;
        if constexpr (D <= TGP_REG_D) {
            if (attempt == 0 && h->algo == TGP_ALGO_AUTO && m->ordering == TGP_FORWARD && time_invariant(*m))
                TGP_TRY(filter_steady<D>(h, d, dy, rq, &handled, &flag));
        }
        if constexpr (D > TGP_REG_D) {
            if (attempt == 0 && h->algo == TGP_ALGO_AUTO && m->ordering == TGP_FORWARD && time_invariant(*m) && !m_f && !P_f && !lml_steps &&
                m->T >= 4 * kSmMaxEnd) {
                bool conv = false;
                TGP_TRY(logpdf_steady_vec<D>(h, d, dy, lml_out, &conv));
                if (conv) return TGP_OK;
                h->pending.clear();               // P did not settle within the head budget: the general scan below
                TGP_CUDA(h, h->arena.reset());
                continue;
            }
        }
        if (!handled) TGP_TRY(filter_general<D>(h, d, dy, rq));
        bool converged = true;
        TGP_TRY(end_call(h, rq.err, err_T(d), m->ordering == TGP_REVERSE, flag, &converged, &rq));
        if (converged) return TGP_OK;
        // P had not reached its fixed point after the transient: redo the series with the general scan
        h->pending.clear();
        TGP_CUDA(h, h->arena.reset());
    }
    return fail(h, TGP_ECUDA, "internal: steady-state fallback did not terminate");
}

// ---- backward pass over the stored filtering distributions (posterior marginals) -----------------
// ---- posterior marginals of a time-invariant model: head / tail by the general kernels, the rest by constant-coefficient
// scans over vectors (tgp_steady_smooth.cuh). *converged = false: nothing was delivered, the caller redoes the call generally.

template <int D>
int posterior_marginals_steady(tgp_ctx* h, const tgp_lgssm& d, const double* dy, const double* dRn, int64_t sRnew, double* dmean, double* dvar,
                               double* lml_out, bool* converged) {
    const int64_t T = d.T;
    cudaStream_t st = h->stream;
    *converged = false;
    SmConst<D>* cst;
    double *wsh, *x0buf, *MFh, *GG, *gg, *SS;
    unsigned long long* err;
    constexpr int SN = D + Sym<D>::N;
    TGP_TRY(dalloc(h, 1, &cst));
    TGP_TRY(dalloc(h, (size_t)kSmMaxEnd * SN, &wsh));
    TGP_TRY(dalloc(h, SN, &x0buf));
    TGP_TRY(dalloc(h, (size_t)kSmMaxEnd * D, &MFh));
    TGP_TRY(dalloc(h, 1, &err));
    TGP_CUDA(h, cudaMemsetAsync(err, 0xFF, sizeof(unsigned long long), st));
    const DevModel dm{d.A, d.a, d.Q, d.H, d.h, d.R, 0, 0, 0, 0, 0, 0, dy, 1, T};
    // head, forward (one CTA, sequential) and the constants of the steady phase
    TGP_K(h, "k_sm_head_fwd");
    k_sm_head_fwd<D><<<1, 128, 0, st>>>(dm, d.m0, d.P0, kSmMaxEnd, h->ss_tol, wsh, MFh, cst, err);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_sm_setup");
    k_sm_setup<D><<<1, 128, 0, st>>>(dm, cst);
    TGP_LAUNCH_CHECK(h);
    // the host needs N0 to size the scans
    long long* pN0 = (long long*)(h->pinned + 24);
    int* pconv = (int*)(h->pinned + 25);
    TGP_CUDA(h, cudaMemcpyAsync(pN0, &cst->N0, sizeof(long long), cudaMemcpyDeviceToHost, st));
    TGP_CUDA(h, cudaMemcpyAsync(pconv, &cst->conv_f, sizeof(int), cudaMemcpyDeviceToHost, st));
    TGP_CUDA(h, cudaStreamSynchronize(st));
    h->d2h += 12;
    const int64_t N0 = *pN0;
    if (!*pconv || T < N0 + 2 * kSmMaxEnd) return TGP_OK;          // not converged (or too short): general path
    // side stream: the tail's covariance recursion and the head's reverse-time dynamics need only the set-up; they overlap the scans
    double* tvar;
    TGP_TRY(dalloc(h, (size_t)kSmMaxEnd, &tvar));
    TGP_TRY(dalloc(h, (size_t)N0 * D * D, &GG));
    TGP_TRY(dalloc(h, (size_t)N0 * D * D, &SS));
    TGP_TRY(dalloc(h, (size_t)N0 * D, &gg));
    const bool overlap = !h->timing;                         // per-kernel timing brackets launches on h->stream only
    cudaStream_t sx = st;
    if (overlap) { TGP_TRY(aux_begin(h)); sx = h->aux_stream; }
    TGP_K(h, "k_sm_tail_var");
    k_sm_tail_var<D><<<1, 128, 0, sx>>>(cst, kSmMaxEnd, h->ss_tol, tvar);
    TGP_LAUNCH_CHECK(h);
    if (N0 > 1) {
        SmootherProvider<D> ph;     // stored states of the head, SoA with stride kSmMaxEnd (the provider's dm.T is that stride)
        ph.dm = DevModel{d.A, d.a, d.Q, d.H, d.h, d.R, 0, 0, 0, 0, 0, 0, nullptr, 1, kSmMaxEnd};
        ph.ws = wsh;
        ph.x0buf = x0buf;           // never read: t >= 1
        ph.err_step = err;
        TGP_K(h, "k_sm_head_dyn");
        k_sm_head_dyn<D><<<(unsigned)((N0 + kBlock - 1) / kBlock), kBlock, 0, sx>>>(ph, N0, GG, gg, SS);
        TGP_LAUNCH_CHECK(h);
    }
    // forward means of [N0, T): MF[j] = m_f[N0 - 1 + j]
    const int64_t nf = T - N0;
    double *MF, *partials, *lml_dev, *xlast;
    TGP_TRY(dalloc(h, (size_t)(nf + 1) * D, &MF));
    const int64_t nblk = ((nf + kCsL - 1) / kCsL + kCsThreads - 1) / kCsThreads;
    TGP_TRY(dalloc(h, (size_t)nblk, &partials));
    TGP_TRY(dalloc(h, 1, &lml_dev));
    TGP_TRY(dalloc(h, D, &xlast));
    TGP_K(h, "k_sm_set_mstart");
    k_sm_set_mstart<D><<<1, 32, 0, st>>>(cst, MF);
    TGP_LAUNCH_CHECK(h);
    FwdItems<D> fi{dy, N0, MF, partials};
    BwdItems<D> bi{};
    TGP_TRY((cs_scan<D, true>(h, cst, fi, bi, nf, cst->mstart, nullptr)));
    TGP_K(h, "k_sm_lml");
    k_sm_lml<D><<<1, 256, 0, st>>>(cst, partials, nblk, nf, lml_dev);
    TGP_LAUNCH_CHECK(h);
    // backward means of t = T-1 .. N0-1 (item 0 = the filtered mean at T-1), var = vss + R_new; the tail variances are overwritten below
    TGP_K(h, "k_sm_set_xfirst");
    k_sm_set_xfirst<D><<<1, 32, 0, st>>>(cst, MF + (size_t)nf * D);
    TGP_LAUNCH_CHECK(h);
    const int64_t nbk = T - N0 + 1;
    bi = BwdItems<D>{MF, nf, T - 1, dRn, sRnew, dmean, dvar};
    TGP_TRY((cs_scan<D, false>(h, cst, fi, bi, nbk, nullptr, xlast)));
    if (overlap) TGP_TRY(aux_end(h));
    TGP_K(h, "k_sm_tail_copy");
    k_sm_tail_copy<D><<<8, 256, 0, st>>>(cst, tvar, T, dRn, sRnew, dvar);
    TGP_LAUNCH_CHECK(h);
    // head, backward
    int* flag;
    TGP_TRY(dalloc(h, 1, &flag));
    TGP_K(h, "k_sm_head_bwd");
    k_sm_head_bwd<D><<<1, 128, 0, st>>>(dm, cst, xlast, GG, gg, SS, dRn, sRnew, dmean, dvar, flag);
    TGP_LAUNCH_CHECK(h);
    TGP_TRY(deliver_scalar(h, lml_dev, lml_out));
    return end_call(h, err, T, false, flag, converged);
}

template <int D>
int do_posterior_marginals(tgp_ctx* h, const tgp_lgssm* m, const double* y, const double* R_new, int64_t sRnew,
                           double* mean_out, double* var_out, double* lml_out) {
    if (m->ordering != TGP_FORWARD) return fail(h, TGP_EUNSUPPORTED, "posterior of a Reverse-ordered model is not built yet");
    constexpr int SN = D + Sym<D>::N;
    const int64_t T = m->T;
    cudaStream_t st = h->stream;
    if (h->algo == TGP_ALGO_AUTO && time_invariant(*m) && T >= 4 * kSmMaxEnd) {
        tgp_lgssm d;
        const double* dy;
        TGP_TRY(stage_model(h, m, y, &d, &dy));
        const double* dRn;
        TGP_TRY(stage_steps(h, R_new, sRnew, T, 1, &dRn));
        double *dmean, *dvar;
        int64_t s1;
        TGP_TRY(stage_out(h, mean_out, 1, 1, T, &dmean, &s1));
        TGP_TRY(stage_out(h, var_out, 1, 1, T, &dvar, &s1));
        bool converged = true;
        TGP_TRY(posterior_marginals_steady<D>(h, d, dy, dRn, sRnew, dmean, dvar, lml_out, &converged));
        if (converged) return TGP_OK;
        h->pending.clear();                  // a covariance had not reached its fixed point: redo with the general kernels
        TGP_CUDA(h, h->arena.reset());
    }
    tgp_lgssm d;
    const double* dy;
    TGP_TRY(stage_model(h, m, y, &d, &dy));
    const double* dRn;
    TGP_TRY(stage_steps(h, R_new, sRnew, T, 1, &dRn));
    FilterReq rq;
    rq.lml_out = lml_out;
    rq.keep_ws = true;
    TGP_TRY(filter_general<D>(h, d, dy, rq));
    double *dmean, *dvar;
    int64_t s1;
    TGP_TRY(stage_out(h, mean_out, 1, 1, T, &dmean, &s1));
    TGP_TRY(stage_out(h, var_out, 1, 1, T, &dvar, &s1));

    SmootherProvider<D> prov;
    prov.dm = DevModel{d.A, d.a, d.Q, d.H, d.h, d.R, d.sA, d.sa, d.sQ, d.sH, d.sh, d.sR, dy, 1, T};
    prov.ws = rq.ws;
    prov.x0buf = rq.x0buf;
    prov.err_step = rq.err;
    const int L = h->chunk > 0 ? h->chunk : 16;
    const int64_t nchunk = (T + L - 1) / L;
    const int64_t grid = (nchunk + kBlock - 1) / kBlock;
    const int64_t nthreads = grid * kBlock, nwarps = nthreads / 32;
    double *excl, *wagg, *wstate;
    TGP_TRY(dalloc(h, (size_t)Aff<D>::N * nthreads, &excl));
    TGP_TRY(dalloc(h, (size_t)Aff<D>::N * nwarps, &wagg));
    TGP_TRY(dalloc(h, (size_t)SN * nwarps, &wstate));
    TGP_K(h, "k_aff_reduce");
    k_aff_reduce<D, SmootherProvider<D>><<<(unsigned)grid, kBlock, 0, st>>>(prov, T, L, nthreads, excl, wagg, nwarps);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_aff_mid");
    k_aff_mid<D><<<1, kMidThreads, 0, st>>>(wagg, nwarps, rq.xT, wstate, nullptr);
    TGP_LAUNCH_CHECK(h);
    const int64_t tl = T - 1;
    EmitOut eo{d.H + tl * d.sH, d.h + tl * d.sh, dRn + tl * sRnew, -d.sH, -d.sh, -sRnew, dmean + tl, dvar + tl, -1};
    TGP_K(h, "k_aff_apply");
    k_aff_apply<D, SmootherProvider<D>><<<(unsigned)grid, kBlock, 0, st>>>(prov, T, L, nthreads, excl, wstate, nwarps, eo, 1);
    TGP_LAUNCH_CHECK(h);
    return end_call(h, rq.err, T, false);
}

template <int D>
int do_posterior(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* G, double* g, double* Sig, double* m_T, double* P_T) {
    if (m->ordering != TGP_FORWARD) return fail(h, TGP_EUNSUPPORTED, "posterior of a Reverse-ordered model is not built yet");
    const int64_t T = m->T;
    cudaStream_t st = h->stream;
    tgp_lgssm d;
    const double* dy;
    TGP_TRY(stage_model(h, m, y, &d, &dy));
    FilterReq rq;
    rq.keep_ws = true;
    TGP_TRY(filter_general<D>(h, d, dy, rq));
    double *dG, *dg, *dS, *dmT, *dPT;
    int64_t s1;
    TGP_TRY(stage_out(h, G, D * D, D * D, T, &dG, &s1));
    TGP_TRY(stage_out(h, g, D, D, T, &dg, &s1));
    TGP_TRY(stage_out(h, Sig, D * D, D * D, T, &dS, &s1));
    TGP_TRY(stage_out(h, m_T, D, D, 1, &dmT, &s1));
    TGP_TRY(stage_out(h, P_T, D * D, D * D, 1, &dPT, &s1));
    SmootherProvider<D> prov;
    prov.dm = DevModel{d.A, d.a, d.Q, d.H, d.h, d.R, d.sA, d.sa, d.sQ, d.sH, d.sh, d.sR, dy, 1, T};
    prov.ws = rq.ws;
    prov.x0buf = rq.x0buf;
    prov.err_step = rq.err;
    if (dG && dg && dS) {
        TGP_K(h, "k_posterior_dynamics");
        k_posterior_dynamics<D><<<(unsigned)((T + kBlock - 1) / kBlock), kBlock, 0, st>>>(prov, dG, dg, dS);
        TGP_LAUNCH_CHECK(h);
    } else if (dG || dg || dS) {
        return fail(h, TGP_EINVAL, "G, g and Sig must be all NULL or all non-NULL");
    }
    TGP_K(h, "k_unpack_state");
    k_unpack_state<D><<<1, 32, 0, st>>>(rq.xT, dmT, dPT);
    TGP_LAUNCH_CHECK(h);
    return end_call(h, rq.err, T, false);
}

template <int D>
int do_marginals(tgp_ctx* h, const tgp_lgssm* m, double* mean_out, double* var_out) {
    constexpr int SN = D + Sym<D>::N;
    const int64_t T = m->T;
    const bool rev = m->ordering == TGP_REVERSE;
    cudaStream_t st = h->stream;
    tgp_lgssm d;
    TGP_TRY(stage_model(h, m, nullptr, &d, nullptr));
    double *dmean, *dvar, *x0buf;
    int64_t s1;
    TGP_TRY(stage_out(h, mean_out, 1, 1, T, &dmean, &s1));
    TGP_TRY(stage_out(h, var_out, 1, 1, T, &dvar, &s1));
    TGP_TRY(dalloc(h, SN, &x0buf));
    TGP_K(h, "k_init_state");
    k_init_state<D><<<1, 32, 0, st>>>(d.m0, d.P0, x0buf, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      nullptr, 0, 0, nullptr);
    TGP_LAUNCH_CHECK(h);
    const int64_t tl = T - 1;
    ModelAffProvider<D> prov;
    EmitOut eo;
    if (!rev) {
        prov.dm = DevModel{d.A, d.a, d.Q, d.H, d.h, d.R, d.sA, d.sa, d.sQ, d.sH, d.sh, d.sR, nullptr, 1, T};
        eo = EmitOut{d.H, d.h, d.R, d.sH, d.sh, d.sR, dmean, dvar, 1};
    } else {
        prov.dm = DevModel{d.A + tl * d.sA, d.a + tl * d.sa, d.Q + tl * d.sQ, d.H, d.h, d.R, -d.sA, -d.sa, -d.sQ, d.sH, d.sh, d.sR, nullptr, 1, T};
        eo = EmitOut{d.H + tl * d.sH, d.h + tl * d.sh, d.R + tl * d.sR, -d.sH, -d.sh, -d.sR, dmean + tl, dvar + tl, -1};
    }
    const int L = h->chunk > 0 ? h->chunk : 16;
    const int64_t nchunk = (T + L - 1) / L;
    const int64_t grid = (nchunk + kBlock - 1) / kBlock;
    const int64_t nthreads = grid * kBlock, nwarps = nthreads / 32;
    double *excl, *wagg, *wstate;
    TGP_TRY(dalloc(h, (size_t)Aff<D>::N * nthreads, &excl));
    TGP_TRY(dalloc(h, (size_t)Aff<D>::N * nwarps, &wagg));
    TGP_TRY(dalloc(h, (size_t)SN * nwarps, &wstate));
    TGP_K(h, "k_aff_reduce");
    k_aff_reduce<D, ModelAffProvider<D>><<<(unsigned)grid, kBlock, 0, st>>>(prov, T, L, nthreads, excl, wagg, nwarps);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_aff_mid");
    k_aff_mid<D><<<1, kMidThreads, 0, st>>>(wagg, nwarps, x0buf, wstate, nullptr);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_aff_apply");
    k_aff_apply<D, ModelAffProvider<D>><<<(unsigned)grid, kBlock, 0, st>>>(prov, T, L, nthreads, excl, wstate, nwarps, eo, rev ? 1 : 0);
    TGP_LAUNCH_CHECK(h);
    return end_call(h, nullptr, T, false);
}

// ---- time-sharded path -----------------------------------------------------------------------------
template <int D>
int do_shard_reduce(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* elem_out) {
    if (m->ordering != TGP_FORWARD) return fail(h, TGP_EUNSUPPORTED, "time sharding runs Forward-ordered models");
    const int64_t T = m->T;
    cudaStream_t st = h->stream;
    tgp_lgssm d;
    const double* dy;
    TGP_TRY(stage_model(h, m, y, &d, &dy));
    double* de;
    int64_t s1;
    TGP_TRY(stage_out(h, elem_out, 3 * D * D + 2 * D, 3 * D * D + 2 * D, 1, &de, &s1));
    DevModel dm{d.A, d.a, d.Q, d.H, d.h, d.R, d.sA, d.sa, d.sQ, d.sH, d.sh, d.sR, dy, 1, T};
    const bool tv = !time_invariant(d);
    ConstModel<D>* cm = nullptr;
    if (!tv) {
        TGP_TRY(dalloc(h, 1, &cm));
        TGP_K(h, "k_const_model");
        k_const_model<D><<<1, 32, 0, st>>>(dm, cm);
        TGP_LAUNCH_CHECK(h);
    }
    const int L = h->chunk > 0 ? h->chunk : 16;
    const int64_t nchunk = (T + L - 1) / L;
    const int64_t grid = (nchunk + kBlock - 1) / kBlock;
    const int64_t nthreads = grid * kBlock, nwarps = nthreads / 32;
    double *excl, *wagg;
    TGP_TRY(dalloc(h, (size_t)Elem<D>::N * nthreads, &excl));
    TGP_TRY(dalloc(h, (size_t)Elem<D>::N * nwarps, &wagg));
    TGP_K(h, "k_filter_reduce");
    if (tv) k_filter_reduce<D, true><<<(unsigned)grid, kBlock, 0, st>>>(dm, cm, L, nthreads, excl, wagg, nwarps);
    else    k_filter_reduce<D, false><<<(unsigned)grid, kBlock, 0, st>>>(dm, cm, L, nthreads, excl, wagg, nwarps);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_elem_total");
    k_elem_total<D><<<1, kMidThreads, 0, st>>>(wagg, nwarps, de);
    TGP_LAUNCH_CHECK(h);
    return end_call(h, nullptr, T, false);
}

// Time-sharded steady-state logpdf (time-invariant models): phase 1 leaves this rank's record (Phi, Z) in xchg_out
// (device, D*D + D doubles) without synchronising; the caller all-gathers the records; phase 2 filters the shard from
// the folded incoming mean and writes the shard's log-likelihood to lml_partial (device).
template <int D>
int do_shard_phase1(tgp_ctx* h, const tgp_lgssm* m, const double* y, int rank, int world, double* xchg_out) {
    if (m->ordering != TGP_FORWARD || !time_invariant(*m))
        return fail(h, TGP_EUNSUPPORTED, "the steady-state sharded path runs Forward, time-invariant models (use tgp_shard_reduce otherwise)");
    if (!is_device_ptr(xchg_out)) return fail(h, TGP_EINVAL, "xchg_out must be a device pointer");
    if constexpr (D <= TGP_REG_D) {
        tgp_lgssm d;
        const double* dy;
        TGP_TRY(stage_model(h, m, y, &d, &dy));
        return shard_phase1<D>(h, &h->shard, d, dy, rank, world, xchg_out);
    } else {
        return fail(h, TGP_EUNSUPPORTED, "the steady-state sharded path is instantiated for D <= %d", TGP_REG_D);
    }
}

template <int D>
int do_shard_phase2(tgp_ctx* h, const double* xchg_all, double* lml_partial) {
    if (!h->shard.active || h->shard.D != D) return fail(h, TGP_EINVAL, "tgp_shard_phase2 without a matching tgp_shard_phase1");
    if (!is_device_ptr(xchg_all) || !is_device_ptr(lml_partial)) return fail(h, TGP_EINVAL, "xchg_all and lml_partial must be device pointers");
    if constexpr (D <= TGP_REG_D) {
        const int64_t T = h->shard.T;
        SSWork<D>& w = *reinterpret_cast<SSWork<D>*>(h->shard.work);
        TGP_TRY(shard_phase2<D>(h, &h->shard, xchg_all, lml_partial));
        // no synchronisation here: the caller enqueues its all-reduce right behind this kernel; status (convergence,
        // positive-definiteness) is collected by tgp_synchronize() or by the next call on this handle.
        if (h->defer_status && h->sticky) {
            TGP_K(h, "k_sticky_status");
            k_sticky_status<<<1, 1, 0, h->stream>>>(w.resblk, h->sticky);
            TGP_LAUNCH_CHECK(h);
        } else {
            h->deferred_res = w.resblk;
            h->deferred_T = T;
        }
        return TGP_OK;
    } else {
        return fail(h, TGP_EUNSUPPORTED, "the steady-state sharded path is instantiated for D <= %d", TGP_REG_D);
    }
}

// One-launch sharded logpdf (tgp_fir.cuh): this rank's shard of a Forward, time-invariant series in ONE un-synchronised launch.
// Shards never wait for each other's results: rank r > 0 takes the observations that precede its shard (the halo) from the ring its
// predecessor's kernel fills over NVLink at the start of its run, and every rank ships its shard's log-likelihood to all peers'
// buffers; tgp_shard_result forms the total when somebody wants it. *handled = false: the plan declined (see logpdf_fir).
template <int D>
int do_shard_logpdf(tgp_ctx* h, const tgp_lgssm* m, const double* y, int rank, int world, bool* handled) {
    *handled = false;
    if (m->ordering != TGP_FORWARD || !time_invariant(*m)) return TGP_OK;
    if constexpr (D <= kFirMaxD) {
        if (world == 1) return logpdf_fir<D>(h, m, y, nullptr, handled, true, nullptr);
        XchgFirView v;
        if (!xchg_fir_view(h, &v) || v.world != world || v.rank != rank)
            return fail(h, TGP_EINVAL, "tgp_shard_logpdf needs an opened exchange (tgp_xchg_create / tgp_xchg_open) of matching rank / world");
        using L = FirXchgLayout;
        const unsigned long long ep = *v.epoch + 1;
        FirXchg x{};
        char* mine = v.self + v.fir_off;
        const bool overlap = h->shard_overlap;
        if (overlap && rank > 0) {
            if (!is_device_ptr(y))
                return fail(h, TGP_EINVAL, "TGP_OPT_SHARD_OVERLAP: the shard (and the TGP_SHARD_HALO observations before it) must be device-resident");
            x.local_halo = 1;
            // the ack word keeps advancing, so a later call with the exchange layout finds the ring slot free
            x.ack_out = reinterpret_cast<unsigned long long*>(v.prev + v.fir_off + L::ack_off());
        } else if (rank > 0) {
            x.halo = reinterpret_cast<const double*>(mine + L::halo_off(ep));
            x.halo_flag = reinterpret_cast<const unsigned long long*>(mine + L::halo_flag_off(ep));
            x.ack_out = reinterpret_cast<unsigned long long*>(v.prev + v.fir_off + L::ack_off());
        }
        if (v.next && !overlap) {
            x.push_dst = reinterpret_cast<double*>(v.next + v.fir_off + L::halo_off(ep));
            x.push_flag = reinterpret_cast<unsigned long long*>(v.next + v.fir_off + L::halo_flag_off(ep));
            x.ack_in = reinterpret_cast<const unsigned long long*>(mine + L::ack_off());
            x.ring = L::kRing;
        }
        x.peers = v.peers;
        x.lml_off = v.fir_off + L::lml_off(world, ep, rank);
        x.epoch = ep;
        x.rank = rank;
        x.world = world;
        TGP_TRY(logpdf_fir<D>(h, m, y, nullptr, handled, rank == 0, &x));
        if (*handled) *v.epoch = ep;
        return TGP_OK;
    } else {
        return TGP_OK;
    }
}

template <int D>
int do_shard_prefix(int n, const double* elems, const double* m0, const double* P0, double* m_in, double* P_in) {
    Vec<D> m;
    Sym<D> P;
    for (int i = 0; i < D; ++i) m[i] = m0[i];
    for (int j = 0; j < D; ++j)
        for (int i = 0; i <= j; ++i) P(i, j) = P0[i + D * j];
    const int ES = 3 * D * D + 2 * D;
    for (int k = 0; k < n; ++k) {
        const double* p = elems + (size_t)k * ES;
        Elem<D> E;
        for (int i = 0; i < D * D; ++i) E.A.v[i] = p[i];
        for (int i = 0; i < D; ++i) E.b[i] = p[D * D + i];
        const double* C = p + D * D + D;
        const double* et = C + D * D;
        const double* J = et + D;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i <= j; ++i) { E.C(i, j) = C[i + D * j]; E.J(i, j) = J[i + D * j]; }
        for (int i = 0; i < D; ++i) E.eta[i] = et[i];
        apply_elem(E, m, P);
    }
    for (int i = 0; i < D; ++i) m_in[i] = m[i];
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < D; ++i) P_in[i + D * j] = P(i, j);
    return TGP_OK;
}


}  // namespace tgp

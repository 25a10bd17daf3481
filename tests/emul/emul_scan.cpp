// emul_scan.cpp — g++-compiled, single-threaded emulation of the chunked scan that the CUDA
// kernels run, built from the SAME tgp_math.cuh. It exists so the scan algebra (fold_step /
// combine / apply_elem / aff_combine / invert_dynamics) can be tested without a GPU.
// TEST HARNESS ONLY: never loaded by the product package, never a fallback.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/tgp_b200.h"
#include "../../temporalgps.jl_b200/csrc/tgp_math.cuh"

using namespace tgp;

template <int D> static Mat<D> ld_mat(const double* p) { Mat<D> m; for (int i = 0; i < D * D; ++i) m.v[i] = p[i]; return m; }
template <int D> static Vec<D> ld_vec(const double* p) { Vec<D> m; for (int i = 0; i < D; ++i) m.v[i] = p[i]; return m; }
template <int D> static Sym<D> ld_sym(const double* p) { Sym<D> s; for (int j = 0; j < D; ++j) for (int i = 0; i <= j; ++i) s(i, j) = p[i + D * j]; return s; }

template <int D>
static StepConst<D> step_const(const tgp_lgssm* md, int64_t t) {
    return make_step_const<D>(ld_mat<D>(md->A + t * md->sA), ld_vec<D>(md->a + t * md->sa), ld_sym<D>(md->Q + t * md->sQ),
                              ld_vec<D>(md->H + t * md->sH), md->h[t * md->sh], md->R[t * md->sR]);
}

// Forward filter by chunked scan. L = steps per chunk, W = chunks per "warp" (Kogge–Stone width).
template <int D>
static int filter_scan(const tgp_lgssm* md, const double* y, int L, int W, double* lml_steps, double* m_f, double* P_f) {
    const int64_t T = md->T;
    const int64_t nchunk = (T + L - 1) / L;
    const int64_t nwarp = (nchunk + W - 1) / W;
    std::vector<Elem<D>> excl(nwarp * W), wagg(nwarp);
    // phase 1: fold + warp-level Kogge–Stone inclusive scan, exclusive = shifted
    for (int64_t w = 0; w < nwarp; ++w) {
        std::vector<Elem<D>> e(W);
        for (int l = 0; l < W; ++l) {
            e[l] = elem_identity<D>();
            const int64_t c = w * W + l;
            for (int64_t t = c * L; t < std::min<int64_t>(T, (c + 1) * L); ++t) fold_step(e[l], step_const<D>(md, t), y[t]);
        }
        for (int off = 1; off < W; off <<= 1) {
            std::vector<Elem<D>> n = e;
            for (int l = off; l < W; ++l) n[l] = combine(e[l - off], e[l]);
            e = n;
        }
        for (int l = 0; l < W; ++l) excl[w * W + l] = l ? e[l - 1] : elem_identity<D>();
        wagg[w] = e[W - 1];
    }
    // mid: states entering each warp
    std::vector<Vec<D>> wm(nwarp); std::vector<Sym<D>> wP(nwarp);
    Vec<D> m = ld_vec<D>(md->m0); Sym<D> P = ld_sym<D>(md->P0);
    for (int64_t w = 0; w < nwarp; ++w) { wm[w] = m; wP[w] = P; apply_elem(wagg[w], m, P); }
    // phase 2
    for (int64_t c = 0; c < nchunk; ++c) {
        Vec<D> mm = wm[c / W]; Sym<D> PP = wP[c / W];
        apply_elem(excl[c], mm, PP);
        for (int64_t t = c * L; t < std::min<int64_t>(T, (c + 1) * L); ++t) {
            predict(mm, PP, ld_mat<D>(md->A + t * md->sA), ld_vec<D>(md->a + t * md->sa), ld_sym<D>(md->Q + t * md->sQ));
            double quad;
            const double S = update_scalar(mm, PP, ld_vec<D>(md->H + t * md->sH), md->h[t * md->sh], md->R[t * md->sR], y[t], &quad);
            if (!(S > 0.0)) return TGP_ENOTPD;
            lml_steps[t] = lml_from(S, quad);
            if (m_f) for (int i = 0; i < D; ++i) m_f[t * D + i] = mm[i];
            if (P_f) for (int j = 0; j < D; ++j) for (int i = 0; i < D; ++i) P_f[t * D * D + i + D * j] = PP(i, j);
        }
    }
    return TGP_OK;
}

// Reverse-time marginals of the posterior model by a chunked (A,b,C) scan over invert_dynamics
// elements built from the filtering distributions (m_f, P_f): posterior_lti_sde.jl:27-36 chain.
template <int D>
static int smooth_scan(const tgp_lgssm* md, const double* m_f, const double* P_f, const double* Rn, int64_t sRn, int L,
                       double* mean, double* var) {
    const int64_t T = md->T;
    const int64_t nchunk = (T + L - 1) / L;
    auto elem_at = [&](int64_t t, Aff<D>& e) {  // maps smoothed x_t -> x_{t-1}; filtered state at t-1 (x0 for t = 0)
        Vec<D> mf = t ? ld_vec<D>(m_f + (t - 1) * D) : ld_vec<D>(md->m0);
        Sym<D> Pf = t ? ld_sym<D>(P_f + (t - 1) * D * D) : ld_sym<D>(md->P0);
        Vec<D> mp = mf; Sym<D> Pp = Pf;
        Mat<D> A = ld_mat<D>(md->A + t * md->sA);
        predict(mp, Pp, A, ld_vec<D>(md->a + t * md->sa), ld_sym<D>(md->Q + t * md->sQ));
        return invert_dynamics(mf, Pf, mp, Pp, A, e);
    };
    std::vector<Aff<D>> agg(nchunk);
    for (int64_t c = 0; c < nchunk; ++c) {
        Aff<D> a = aff_identity<D>();
        for (int64_t t = std::min<int64_t>(T, (c + 1) * L) - 1; t >= c * L; --t) { Aff<D> e; if (!elem_at(t, e)) return TGP_ENOTPD; a = aff_combine(a, e); }
        agg[c] = a;
    }
    Vec<D> m = ld_vec<D>(m_f + (T - 1) * D); Sym<D> P = ld_sym<D>(P_f + (T - 1) * D * D);
    std::vector<Vec<D>> cm(nchunk); std::vector<Sym<D>> cP(nchunk);
    for (int64_t c = nchunk - 1; c >= 0; --c) { cm[c] = m; cP[c] = P; aff_apply(agg[c], m, P); }
    for (int64_t c = 0; c < nchunk; ++c) {
        Vec<D> mm = cm[c]; Sym<D> PP = cP[c];
        for (int64_t t = std::min<int64_t>(T, (c + 1) * L) - 1; t >= c * L; --t) {
            emit_scalar(mm, PP, ld_vec<D>(md->H + t * md->sH), md->h[t * md->sh], Rn[t * sRn], mean + t, var + t);
            Aff<D> e; elem_at(t, e); aff_apply(e, mm, PP);
        }
    }
    return TGP_OK;
}

#define DISPATCH(D_, call) switch (D_) { case 1: return call<1>; case 2: return call<2>; case 3: return call<3>; case 4: return call<4>; default: return TGP_EUNSUPPORTED; }

extern "C" int emul_filter_scan(const tgp_lgssm* md, const double* y, int L, int W, double* lml_steps, double* m_f, double* P_f) {
    switch (md->D) {
        case 1: return filter_scan<1>(md, y, L, W, lml_steps, m_f, P_f);
        case 2: return filter_scan<2>(md, y, L, W, lml_steps, m_f, P_f);
        case 3: return filter_scan<3>(md, y, L, W, lml_steps, m_f, P_f);
        case 4: return filter_scan<4>(md, y, L, W, lml_steps, m_f, P_f);
        case 6: return filter_scan<6>(md, y, L, W, lml_steps, m_f, P_f);
        case 8: return filter_scan<8>(md, y, L, W, lml_steps, m_f, P_f);
        case 10: return filter_scan<10>(md, y, L, W, lml_steps, m_f, P_f);
        default: return TGP_EUNSUPPORTED;
    }
}
extern "C" int emul_smooth_scan(const tgp_lgssm* md, const double* m_f, const double* P_f, const double* Rn, int64_t sRn, int L,
                                double* mean, double* var) {
    switch (md->D) {
        case 1: return smooth_scan<1>(md, m_f, P_f, Rn, sRn, L, mean, var);
        case 2: return smooth_scan<2>(md, m_f, P_f, Rn, sRn, L, mean, var);
        case 3: return smooth_scan<3>(md, m_f, P_f, Rn, sRn, L, mean, var);
        case 4: return smooth_scan<4>(md, m_f, P_f, Rn, sRn, L, mean, var);
        case 6: return smooth_scan<6>(md, m_f, P_f, Rn, sRn, L, mean, var);
        case 8: return smooth_scan<8>(md, m_f, P_f, Rn, sRn, L, mean, var);
        case 10: return smooth_scan<10>(md, m_f, P_f, Rn, sRn, L, mean, var);
        default: return TGP_EUNSUPPORTED;
    }
}

// Phase 1 of the time-sharded path on the CPU (test stand-in for tgp_shard_reduce): folds a whole shard into ONE
// element in the ABI's shard format (A, b, C, eta, J; full column-major matrices).
template <int D>
static int shard_reduce(const tgp_lgssm* md, const double* y, double* out) {
    Elem<D> E = elem_identity<D>();
    for (int64_t t = 0; t < md->T; ++t) fold_step(E, step_const<D>(md, t), y[t]);
    double* p = out;
    for (int k = 0; k < D * D; ++k) p[k] = E.A.v[k];
    p += D * D;
    for (int k = 0; k < D; ++k) p[k] = E.b.v[k];
    p += D;
    for (int j = 0; j < D; ++j) for (int i = 0; i < D; ++i) p[i + D * j] = E.C(i, j);
    p += D * D;
    for (int k = 0; k < D; ++k) p[k] = E.eta.v[k];
    p += D;
    for (int j = 0; j < D; ++j) for (int i = 0; i < D; ++i) p[i + D * j] = E.J(i, j);
    return TGP_OK;
}
extern "C" int emul_shard_reduce(const tgp_lgssm* md, const double* y, double* out) {
    switch (md->D) {
        case 1: return shard_reduce<1>(md, y, out);
        case 2: return shard_reduce<2>(md, y, out);
        case 3: return shard_reduce<3>(md, y, out);
        case 4: return shard_reduce<4>(md, y, out);
        default: return TGP_EUNSUPPORTED;
    }
}

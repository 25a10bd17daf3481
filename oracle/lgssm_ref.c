/*
 * lgssm_ref.c — plain-C restatement of TemporalGPs.jl's sequential LGSSM recursions.
 *
 * TEST INFRASTRUCTURE ONLY: this is the oracle the CUDA path is checked against, and the CPU
 * baseline timed beside it (bench.py cpu_baseline / --impl reference). Nothing under
 * temporalgps.jl_b200/ links, loads or calls it.
 *
 * What it follows (paths relative to /root/reference):
 *   src/util/scan.jl:15-28                                   the sequential scan_emit loop
 *   src/models/linear_gaussian_conditionals.jl:46-52         predict
 *   src/models/linear_gaussian_conditionals.jl:247-257       posterior_and_lml (ScalarOutputLGC)
 *   src/models/linear_gaussian_conditionals.jl:129-141       posterior_and_lml (SmallOutputLGC)
 *   src/models/lgssm.jl:99-115, 147-187, 193-240             marginals / logpdf / _filter / posterior
 *   src/models/gauss_markov_model.jl:36-46                   index order for Forward / Reverse
 * The model descriptor is the tgp_lgssm struct of include/tgp_b200.h (shared so that tests hand
 * identical inputs to both sides).
 *
 * The `_sD` entry points (D = 1, 2, 3, 4) are compile-time-sized, fully unrolled instantiations —
 * the analogue of the reference's SArrayStorage(Float64) path; the generic ones use run-time
 * dimensions like ArrayStorage (without Julia's per-step heap allocation). Single-threaded,
 * like the reference.
 *
 * Pinning: see oracle/tgp_oracle.py header — no golden vectors exist in the reference and Julia
 * is absent, so the pins are the reference's own dense-GP equivalence tests, re-run in
 * tests/test_oracle_pins.py against both this file and the NumPy restatement.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tgp_b200.h"

#define REF_LOG2PI 1.8378770664093454835606594728112

/* Julia's sum(::Vector{Float64}) is pairwise with <=1024-element sequential leaves
 * (Base.mapreduce_impl); lgssm.jl:150 sums the emitted lmls that way. */
double oracle_pairwise_sum(const double* x, int64_t n) {
    if (n <= 1024) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += x[i];
        return s;
    }
    const int64_t h = n / 2;
    return oracle_pairwise_sum(x, h) + oracle_pairwise_sum(x + h, n - h);
}

/* ---- instantiations ----------------------------------------------------------------------- */
#define SUFFIX _g
#include "lgssm_ref_steps.inc"
#define SUFFIX _s1
#define FIXED_D 1
#include "lgssm_ref_steps.inc"
#define SUFFIX _s2
#define FIXED_D 2
#include "lgssm_ref_steps.inc"
#define SUFFIX _s3
#define FIXED_D 3
#include "lgssm_ref_steps.inc"
#define SUFFIX _s4
#define FIXED_D 4
#include "lgssm_ref_steps.inc"

/* ---- vector emissions (SmallOutputLGC), run-time D and M ------------------------------------- */
static void load_R(const tgp_lgssm* md, const double* Rp, double* R /* M x M */) {
    const int M = md->M;
    memset(R, 0, sizeof(double) * M * M);
    if (md->R_kind == TGP_R_DENSE) memcpy(R, Rp, sizeof(double) * M * M);
    else if (md->R_kind == TGP_R_DIAG) for (int i = 0; i < M; ++i) R[i + M * i] = Rp[i];
    else for (int i = 0; i < M; ++i) R[i + M * i] = Rp[0];
}

/* posterior_and_lml(x, f::SmallOutputLGC, y) — LGC:129-141. */
static int update_small(int D, int M, double* m, double* P, const double* H, const double* h,
                        const double* R, const double* y, double* lml) {
    double* V = malloc(sizeof(double) * (M * D + M * M + M * D + 2 * M));
    double* S = V + M * D;
    double* B = S + M * M;
    double* al = B + M * D;
    double* r = al + M;
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += H[i + M * k] * P[k + D * j];
            V[i + M * j] = s;
        }
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += V[i + M * k] * H[j + M * k];
            S[i + M * j] = s + R[i + M * j];
        }
    if (ref_chol_upper_g(M, S)) { free(V); return 1; }
    /* B = U' \ V ; alpha = U' \ (y - (H m + h)) */
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < M; ++i) {
            double s = V[i + M * j];
            for (int k = 0; k < i; ++k) s -= S[k + M * i] * B[k + M * j];
            B[i + M * j] = s / S[i + M * i];
        }
    for (int i = 0; i < M; ++i) {
        double s = 0.0;
        for (int k = 0; k < D; ++k) s += H[i + M * k] * m[k];
        r[i] = y[i] - (s + h[i]);
    }
    double logdet = 0.0, aa = 0.0;
    for (int i = 0; i < M; ++i) {
        double s = r[i];
        for (int k = 0; k < i; ++k) s -= S[k + M * i] * al[k];
        al[i] = s / S[i + M * i];
        aa += al[i] * al[i];
        logdet += 2.0 * log(S[i + M * i]);
    }
    *lml = -(M * REF_LOG2PI + logdet + aa) / 2.0;
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int k = 0; k < M; ++k) s += B[k + M * i] * al[k];
        m[i] += s;
    }
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            for (int k = 0; k < M; ++k) s += B[k + M * i] * B[k + M * j];
            P[i + D * j] -= s;
        }
    free(V);
    return 0;
}

static int64_t filter_small(const tgp_lgssm* md, const double* y, double* lml_steps, double* m_f,
                            int64_t s_m, double* P_f, int64_t s_P, double* m_T, double* P_T) {
    const int D = md->D, M = md->M;
    double* m = malloc(sizeof(double) * (D + D * D + M * M));
    double* P = m + D;
    double* R = P + D * D;
    memcpy(m, md->m0, sizeof(double) * D);
    memcpy(P, md->P0, sizeof(double) * D * D);
    const int rev = md->ordering == TGP_REVERSE;
    int64_t rc = 0;
    for (int64_t n = 0; n < md->T && !rc; ++n) {
        const int64_t t = rev ? md->T - 1 - n : n;
        const double* A = md->A + t * md->sA;
        const double* a = md->a + t * md->sa;
        const double* Q = md->Q + t * md->sQ;
        double lml = 0.0;
        load_R(md, md->R + t * md->sR, R);
        if (!rev) ref_predict_g(D, m, P, A, a, Q);
        if (update_small(D, M, m, P, md->H + t * md->sH, md->h + t * md->sh, R, y + t * M, &lml)) rc = 1 + t;
        if (lml_steps) lml_steps[t] = lml;
        if (m_f) memcpy(m_f + t * s_m, m, sizeof(double) * D);
        if (P_f) memcpy(P_f + t * s_P, P, sizeof(double) * D * D);
        if (rev) ref_predict_g(D, m, P, A, a, Q);
    }
    if (m_T) memcpy(m_T, m, sizeof(double) * D);
    if (P_T) memcpy(P_T, P, sizeof(double) * D * D);
    free(m);
    return rc;
}

/* ---- public oracle entry points (mirror the tgp_* ABI; same status codes) -------------------- */
#define DISPATCH_D(call_s1, call_s2, call_s3, call_s4, call_g) \
    switch (specialise ? md->D : 0) {                         \
        case 1: rc = call_s1; break;                          \
        case 2: rc = call_s2; break;                          \
        case 3: rc = call_s3; break;                          \
        case 4: rc = call_s4; break;                          \
        default: rc = call_g; break;                          \
    }

static int g_specialise = 1;
/* 0: always use the run-time-sized (ArrayStorage-like) code; 1: use the static instantiations */
void oracle_set_static(int on) { g_specialise = on; }

int oracle_filter(const tgp_lgssm* md, const double* y, double* m_f, int64_t s_m, double* P_f,
                  int64_t s_P, double* lml_out, double* lml_per_step, int64_t* fail_t) {
    if (md->T <= 0 || md->D <= 0 || md->D > 64 || md->M <= 0) return TGP_EINVAL;
    const int specialise = g_specialise;
    double* steps = lml_per_step;
    if (!steps && lml_out) {
        steps = malloc(sizeof(double) * md->T);
        if (!steps) return TGP_ENOMEM;
    }
    int64_t rc;
    if (md->M == 1 && md->R_kind == TGP_R_SCALAR) {
        DISPATCH_D(ref_filter_scalar_s1(1, md, y, steps, m_f, s_m, P_f, s_P, 0, 0),
                   ref_filter_scalar_s2(2, md, y, steps, m_f, s_m, P_f, s_P, 0, 0),
                   ref_filter_scalar_s3(3, md, y, steps, m_f, s_m, P_f, s_P, 0, 0),
                   ref_filter_scalar_s4(4, md, y, steps, m_f, s_m, P_f, s_P, 0, 0),
                   ref_filter_scalar_g(md->D, md, y, steps, m_f, s_m, P_f, s_P, 0, 0))
    } else {
        rc = filter_small(md, y, steps, m_f, s_m, P_f, s_P, 0, 0);
    }
    if (rc == 0 && lml_out) *lml_out = oracle_pairwise_sum(steps, md->T);
    if (steps != lml_per_step) free(steps);
    if (rc) { if (fail_t) *fail_t = rc - 1; return TGP_ENOTPD; }
    return TGP_OK;
}

int oracle_logpdf(const tgp_lgssm* md, const double* y, double* lml_out, double* lml_per_step) {
    return oracle_filter(md, y, 0, 0, 0, 0, lml_out, lml_per_step, 0);
}

/* All host cores, the only way the single-threaded reference can use them for this path: `n` independent logpdf evaluations of the
 * same model on `n` series (y + r * ystride), one POSIX thread each (this toolchain ships no libgomp). NOT what the reference does
 * for one series — reported beside the 1-thread number as "replicas" (BASELINE.md section 2). lml_out[n]. Returns the threads used. */
typedef struct { const tgp_lgssm* md; const double* y; double* out; } replica_arg;
static void* replica_main(void* p) {
    replica_arg* a = (replica_arg*)p;
    oracle_logpdf(a->md, a->y, a->out, NULL);
    return NULL;
}
int oracle_logpdf_replicas(const tgp_lgssm* md, const double* y, int64_t ystride, int n, double* lml_out) {
    if (n < 1 || n > 1024) return 0;
    pthread_t th[1024];
    replica_arg args[1024];
    for (int r = 0; r < n; ++r) {
        args[r].md = md; args[r].y = y + (int64_t)r * ystride; args[r].out = lml_out + r;
        if (pthread_create(&th[r], NULL, replica_main, &args[r]) != 0) { replica_main(&args[r]); th[r] = 0; }
    }
    for (int r = 0; r < n; ++r) if (th[r]) pthread_join(th[r], NULL);
    return n;
}

int oracle_posterior(const tgp_lgssm* md, const double* y, double* G, double* g, double* Sig,
                     double* m_T, double* P_T, double* lml_out) {
    if (md->M != 1 || md->R_kind != TGP_R_SCALAR) return TGP_EUNSUPPORTED;
    const int specialise = g_specialise;
    double* steps = lml_out ? malloc(sizeof(double) * md->T) : 0;
    int64_t rc;
    DISPATCH_D(ref_posterior_scalar_s1(1, md, y, G, g, Sig, m_T, P_T, steps),
               ref_posterior_scalar_s2(2, md, y, G, g, Sig, m_T, P_T, steps),
               ref_posterior_scalar_s3(3, md, y, G, g, Sig, m_T, P_T, steps),
               ref_posterior_scalar_s4(4, md, y, G, g, Sig, m_T, P_T, steps),
               ref_posterior_scalar_g(md->D, md, y, G, g, Sig, m_T, P_T, steps))
    if (lml_out) { *lml_out = oracle_pairwise_sum(steps, md->T); free(steps); }
    return rc ? TGP_ENOTPD : TGP_OK;
}

int oracle_marginals(const tgp_lgssm* md, double* mean_out, double* var_out) {
    if (md->M != 1 || md->R_kind != TGP_R_SCALAR) return TGP_EUNSUPPORTED;
    const int specialise = g_specialise;
    int64_t rc = 0;
    DISPATCH_D((ref_marginals_scalar_s1(1, md, 0, 0, mean_out, var_out), 0),
               (ref_marginals_scalar_s2(2, md, 0, 0, mean_out, var_out), 0),
               (ref_marginals_scalar_s3(3, md, 0, 0, mean_out, var_out), 0),
               (ref_marginals_scalar_s4(4, md, 0, 0, mean_out, var_out), 0),
               (ref_marginals_scalar_g(md->D, md, 0, 0, mean_out, var_out), 0))
    (void)rc;
    return TGP_OK;
}

/* marginals(replace_observation_noise_cov(posterior(model, y), R_new)) —
 * src/gp/posterior_lti_sde.jl:27-36 at the training inputs. Materialises (G, g, Sigma) exactly like
 * the reference (lgssm.jl:196-199), then runs the reverse-ordered marginals recursion. */
int oracle_posterior_marginals(const tgp_lgssm* md, const double* y, const double* R_new,
                               int64_t sRnew, double* mean_out, double* var_out, double* lml_out) {
    if (md->M != 1 || md->R_kind != TGP_R_SCALAR) return TGP_EUNSUPPORTED;
    const int D = md->D;
    const int64_t T = md->T;
    double* G = malloc(sizeof(double) * T * (2 * D * D + D));
    if (!G) return TGP_ENOMEM;
    double* Sig = G + T * D * D;
    double* g = Sig + T * D * D;
    double mT[64], PT[64 * 64];
    int rc = oracle_posterior(md, y, G, g, Sig, mT, PT, lml_out);
    if (rc == TGP_OK) {
        tgp_lgssm post = *md;
        post.ordering = md->ordering == TGP_FORWARD ? TGP_REVERSE : TGP_FORWARD;
        post.A = G; post.sA = D * D;
        post.a = g; post.sa = D;
        post.Q = Sig; post.sQ = D * D;
        post.m0 = mT; post.P0 = PT;
        post.R = R_new; post.sR = sRnew;
        rc = oracle_marginals(&post, mean_out, var_out);
    }
    free(G);
    return rc;
}

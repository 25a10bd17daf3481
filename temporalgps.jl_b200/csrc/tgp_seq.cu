// tgp_seq.cu — the smoother-side entry points for EVERY shape the scan kernels do not cover: vector observations
// (SmallOutputLGC, LGC:129-141: space-time models), Reverse-ordered models (the posterior of a posterior), any D, M <= 64.
//     posterior(::LGSSM, y)            lgssm.jl:193-228 (step_posterior, both orderings; invert_dynamics :231-240, jitter 1e-10)
//     marginals / marginals_diag       lgssm.jl:99-141  (step_marginals, step_marginals_diag; predict LGC:46-52,
//                                      predict_marginals LGC:63-68)
//     marginals(replace_observation_noise_cov(posterior(model, y), R_new))     posterior_lti_sde.jl:27-36
// These calls are the small-T / moderate-D corner of the reference's API (its own tests run them at T <= 50); they are executed
// the way the reference executes them — step by step — by ONE CTA whose threads share every matrix product, Cholesky and
// triangular solve of a step through shared memory. Hand-written FP64, no library call. (The throughput paths are the scan /
// steady kernels for scalar observations and tgp_dense*.cu for the large-state filter.)
#include <algorithm>

#include "tgp_ctx.cuh"
#include "tgp_dispatch.h"

namespace tgp {

constexpr int kSeqThreads = 256;
constexpr int kSeqMaxDim = 64;
constexpr double kSeqLog2Pi = 1.8378770664093454835606594728112;
constexpr double kSeqJitter = 1e-10;      // lgssm.jl:235

struct SeqModel {
    int D, M, rkind, reverse;
    long long T;
    const double *A, *a, *Q, *H, *h, *R, *m0, *P0, *y;
    long long sA, sa, sQ, sH, sh, sR;
};

struct SeqOut {
    // filter / posterior
    double *G, *g, *Sig;          // T x D x D, T x D, T x D x D (memory order = model index), nullable
    double *mT, *PT;              // final state (x0 of the posterior model)
    double *lml;                  // total log-likelihood, nullable
    // marginals
    double *mean, *cov;           // T x M, T x M x M (diag: T x M)
    int diag;
    const double* Rnew;           // posterior marginals: replacement observation noise, stride sRnew
    long long sRnew;
    unsigned long long* err;      // failing step (memory index) of a Cholesky, ~0 if none
};

// ---- CTA-cooperative column-major helpers (all threads call; no barrier inside unless stated) --------------------------------------
__device__ __forceinline__ int seq_tid() { return threadIdx.x; }

// C (m x n) = op(A) (m x k) * op(B) (k x n) [+ C0]; ta / tb: operand stored transposed. lda etc. = leading dimensions.
__device__ void seq_mm(double* C, int ldc, const double* A, int lda, bool ta, const double* B, int ldb, bool tb, int m, int n, int k,
                       const double* C0 = nullptr, int ldc0 = 0, double beta = 1.0) {
    for (int e = seq_tid(); e < m * n; e += kSeqThreads) {
        const int i = e % m, j = e / m;
        double s = C0 ? beta * C0[i + (size_t)ldc0 * j] : 0.0;
        for (int l = 0; l < k; ++l) {
            const double x = ta ? A[l + (size_t)lda * i] : A[i + (size_t)lda * l];
            const double z = tb ? B[j + (size_t)ldb * l] : B[l + (size_t)ldb * j];
            s = fma(x, z, s);
        }
        C[i + (size_t)ldc * j] = s;
    }
}
// Symmetric(P): mirror the upper triangle into the lower one (the reference reads P through Symmetric, LGC:51).
__device__ void seq_symmetrize(double* P, int n) {
    for (int e = seq_tid(); e < n * n; e += kSeqThreads) {
        const int i = e % n, j = e / n;
        if (i > j) P[i + (size_t)n * j] = P[j + (size_t)n * i];
    }
}
// In-place upper Cholesky S = U'U (upper triangle read; strict lower zeroed). Contains barriers. Returns false on a non-positive pivot.
__device__ bool seq_chol(double* S, int n, int* flag) {
    if (seq_tid() == 0) *flag = 1;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        if (seq_tid() == 0) {
            double d = S[j + (size_t)n * j];
            for (int k = 0; k < j; ++k) d = fma(-S[k + (size_t)n * j], S[k + (size_t)n * j], d);
            if (!(d > 0.0)) { *flag = 0; d = 1.0; }
            S[j + (size_t)n * j] = sqrt(d);
        }
        __syncthreads();
        const double inv = 1.0 / S[j + (size_t)n * j];
        for (int c = j + 1 + seq_tid(); c < n; c += kSeqThreads) {
            double s = S[j + (size_t)n * c];
            for (int k = 0; k < j; ++k) s = fma(-S[k + (size_t)n * j], S[k + (size_t)n * c], s);
            S[j + (size_t)n * c] = s * inv;
        }
        __syncthreads();
    }
    for (int e = seq_tid(); e < n * n; e += kSeqThreads) {
        const int i = e % n, j = e / n;
        if (i > j) S[i + (size_t)n * j] = 0.0;
    }
    __syncthreads();
    return *flag != 0;
}
// X (n x nrhs) <- U' \ X (forward substitution) or U \ X (back substitution), one thread per right-hand side.
__device__ void seq_trsm(const double* U, int n, double* X, int ldx, int nrhs, bool transposed) {
    for (int c = seq_tid(); c < nrhs; c += kSeqThreads) {
        double* x = X + (size_t)ldx * c;
        if (transposed) {
            for (int i = 0; i < n; ++i) {
                double s = x[i];
                for (int k = 0; k < i; ++k) s = fma(-U[k + (size_t)n * i], x[k], s);
                x[i] = s / U[i + (size_t)n * i];
            }
        } else {
            for (int i = n - 1; i >= 0; --i) {
                double s = x[i];
                for (int k = i + 1; k < n; ++k) s = fma(-U[i + (size_t)n * k], x[k], s);
                x[i] = s / U[i + (size_t)n * i];
            }
        }
    }
}

struct SeqSmem {
    double *m, *P, *mp, *Pp, *W, *V, *S, *B, *al, *A, *Q, *H, *R, *t1;
};
__device__ SeqSmem seq_carve(double* s, int D, int M) {
    SeqSmem w;
    w.m = s; s += D;
    w.mp = s; s += D;
    w.al = s; s += (M > D ? M : D);
    w.P = s; s += D * D;
    w.Pp = s; s += D * D;
    w.W = s; s += D * D;
    w.t1 = s; s += D * D;
    w.V = s; s += M * D;
    w.B = s; s += M * D;
    w.S = s; s += M * M;
    w.A = s; s += D * D;
    w.Q = s; s += D * D;
    w.H = s; s += M * D;
    w.R = s; s += M * M;
    return w;
}
static size_t seq_smem_bytes(int D, int M) {
    return sizeof(double) * (size_t)(2 * D + std::max(M, D) + 6 * D * D + 3 * M * D + 2 * M * M) + 64;
}

// Step parameters of memory index t -> shared memory (R expanded to a dense M x M matrix).
__device__ void seq_load_step(const SeqModel& md, long long t, const SeqSmem& w) {
    const int D = md.D, M = md.M;
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) { w.A[e] = md.A[t * md.sA + e]; w.Q[e] = md.Q[t * md.sQ + e]; }
    for (int e = seq_tid(); e < M * D; e += kSeqThreads) w.H[e] = md.H[t * md.sH + e];
    for (int e = seq_tid(); e < M * M; e += kSeqThreads) {
        const int i = e % M, j = e / M;
        double r;
        if (md.rkind == TGP_R_DENSE) r = md.R[t * md.sR + e];
        else if (md.rkind == TGP_R_DIAG) r = i == j ? md.R[t * md.sR + i] : 0.0;
        else r = md.R[t * md.sR];
        w.R[e] = r;
    }
}
// predict (LGC:46-52): (m, P) -> (mp, Pp) with the step's (A, a, Q) in shared memory. Contains barriers.
__device__ void seq_predict(const SeqModel& md, long long t, const SeqSmem& w, const double* m, double* P, double* mp, double* Pp) {
    const int D = md.D;
    seq_symmetrize(P, D);
    for (int i = seq_tid(); i < D; i += kSeqThreads) {
        double s = md.a[t * md.sa + i];
        for (int l = 0; l < D; ++l) s = fma(w.A[i + (size_t)D * l], m[l], s);
        mp[i] = s;
    }
    __syncthreads();
    seq_mm(w.W, D, w.A, D, false, P, D, false, D, D, D);
    __syncthreads();
    seq_mm(Pp, D, w.W, D, false, w.A, D, true, D, D, D, w.Q, D);
    __syncthreads();
}
// posterior_and_lml(x, SmallOutputLGC(H, h, R), y) (LGC:129-141; M = 1 reproduces ScalarOutputLGC :247-257): (m, P) updated in place.
// Returns lml through *lml (thread 0 valid). Contains barriers. ok = false on a failed Cholesky.
__device__ bool seq_update(const SeqModel& md, long long t, const SeqSmem& w, double* m, double* P, double* lml, int* flag) {
    const int D = md.D, M = md.M;
    seq_symmetrize(P, D);
    __syncthreads();
    seq_mm(w.V, M, w.H, M, false, P, D, false, M, D, D);                 // V = H P
    __syncthreads();
    seq_mm(w.S, M, w.V, M, false, w.H, M, true, M, M, D, w.R, M);        // S = V H' + R
    for (int i = seq_tid(); i < M; i += kSeqThreads) {                   // residual y - H m - h
        double s = md.y[t * M + i] - md.h[t * md.sh + i];
        for (int l = 0; l < D; ++l) s = fma(-w.H[i + (size_t)M * l], m[l], s);
        w.al[i] = s;
    }
    __syncthreads();
    const bool ok = seq_chol(w.S, M, flag);
    for (int e = seq_tid(); e < M * D; e += kSeqThreads) w.B[e] = w.V[e];
    __syncthreads();
    seq_trsm(w.S, M, w.B, M, D, true);                                   // B = U' \ V
    if (seq_tid() == kSeqThreads - 1) {      // alpha = U' \ r by the last thread (a single right-hand side)
        for (int i = 0; i < M; ++i) {
            double s = w.al[i];
            for (int k = 0; k < i; ++k) s = fma(-w.S[k + (size_t)M * i], w.al[k], s);
            w.al[i] = s / w.S[i + (size_t)M * i];
        }
    }
    __syncthreads();
    if (seq_tid() == 0) {
        double q = 0.0, ld = 0.0;
        for (int i = 0; i < M; ++i) { q = fma(w.al[i], w.al[i], q); ld += log(w.S[i + (size_t)M * i]); }
        *lml = -0.5 * ((double)M * kSeqLog2Pi + 2.0 * ld + q);
    }
    for (int i = seq_tid(); i < D; i += kSeqThreads) {                   // m += B' alpha
        double s = m[i];
        for (int l = 0; l < M; ++l) s = fma(w.B[l + (size_t)M * i], w.al[l], s);
        w.mp[i] = s;
    }
    seq_mm(w.W, D, w.B, M, true, w.B, M, false, D, D, M);                // B'B
    __syncthreads();
    for (int i = seq_tid(); i < D; i += kSeqThreads) m[i] = w.mp[i];
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) P[e] -= w.W[e];
    __syncthreads();
    return ok;
}
// invert_dynamics(xf, xp, A) (lgssm.jl:231-240) -> G, g, Sig in global memory (column-major G, Sig). Uses W, t1, S-free scratch. Barriers.
__device__ bool seq_invert(const SeqModel& md, const SeqSmem& w, const double* mf, const double* Pf, const double* mp, const double* Pp,
                           double* G, double* g, double* Sig, int* flag) {
    const int D = md.D;
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) {
        const int i = e % D, j = e / D;
        w.W[e] = (i <= j ? Pp[e] : Pp[j + (size_t)D * i]) + (i == j ? kSeqJitter : 0.0);      // Symmetric(Pp + eps I)
    }
    __syncthreads();
    const bool ok = seq_chol(w.W, D, flag);                                                    // U
    seq_mm(w.t1, D, w.A, D, false, Pf, D, false, D, D, D);                                     // X = A Pf
    __syncthreads();
    seq_trsm(w.W, D, w.t1, D, D, true);                                                        // B = U' \ X   (= U Gt)
    __syncthreads();
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) {                                     // Sig = Pf - B'B
        const int i = e % D, j = e / D;
        double s = Pf[e];
        for (int k = 0; k < D; ++k) s = fma(-w.t1[k + (size_t)D * i], w.t1[k + (size_t)D * j], s);
        Sig[e] = s;
    }
    __syncthreads();
    seq_trsm(w.W, D, w.t1, D, D, false);                                                       // Gt = U \ B
    __syncthreads();
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) {
        const int i = e % D, j = e / D;
        G[e] = w.t1[j + (size_t)D * i];                                                        // G = Gt'
    }
    for (int i = seq_tid(); i < D; i += kSeqThreads) {
        double s = mf[i];
        for (int l = 0; l < D; ++l) s = fma(-w.t1[l + (size_t)D * i], mp[l], s);               // g = mf - Gt' mp
        g[i] = s;
    }
    __syncthreads();
    return ok;
}

// posterior(model, y): emits (G, g, Sig) per memory index and the final state. One CTA.
__global__ void __launch_bounds__(kSeqThreads) k_seq_posterior(const SeqModel md, const SeqOut out) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int flag;
    __shared__ double lml_step;
    const int D = md.D;
    SeqSmem w = seq_carve(smem, D, md.M);
    for (int i = seq_tid(); i < D; i += kSeqThreads) w.m[i] = md.m0[i];
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) w.P[e] = md.P0[e];
    double lml = 0.0;
    __syncthreads();
    for (long long k = 0; k < md.T; ++k) {
        const long long t = md.reverse ? md.T - 1 - k : k;
        seq_load_step(md, t, w);
        __syncthreads();
        double* G = out.G ? out.G + t * D * D : nullptr;
        double* g = out.g ? out.g + t * D : nullptr;
        double* Sg = out.Sig ? out.Sig + t * D * D : nullptr;
        bool ok = true;
        if (!md.reverse) {       // step_posterior(::Forward): predict, invert_dynamics(xf, xp), update
            seq_predict(md, t, w, w.m, w.P, w.mp, w.Pp);
            if (G) ok = seq_invert(md, w, w.m, w.P, w.mp, w.Pp, G, g, Sg, &flag) && ok;
            for (int i = seq_tid(); i < D; i += kSeqThreads) w.m[i] = w.mp[i];
            for (int e = seq_tid(); e < D * D; e += kSeqThreads) w.P[e] = w.Pp[e];
            __syncthreads();
            ok = seq_update(md, t, w, w.m, w.P, &lml_step, &flag) && ok;
        } else {                 // step_posterior(::Reverse): update, predict, invert_dynamics(xp, xf)  (arguments swapped, lgssm.jl:227)
            ok = seq_update(md, t, w, w.m, w.P, &lml_step, &flag) && ok;
            seq_predict(md, t, w, w.m, w.P, w.mp, w.Pp);
            if (G) ok = seq_invert(md, w, w.mp, w.Pp, w.m, w.P, G, g, Sg, &flag) && ok;
            for (int i = seq_tid(); i < D; i += kSeqThreads) w.m[i] = w.mp[i];
            for (int e = seq_tid(); e < D * D; e += kSeqThreads) w.P[e] = w.Pp[e];
            __syncthreads();
        }
        if (seq_tid() == 0) {
            lml += lml_step;
            if (!ok) atomicMin(out.err, (unsigned long long)t);
        }
        __syncthreads();
    }
    for (int i = seq_tid(); i < D; i += kSeqThreads) if (out.mT) out.mT[i] = w.m[i];
    seq_symmetrize(w.P, D);
    __syncthreads();
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) if (out.PT) out.PT[e] = w.P[e];
    if (seq_tid() == 0 && out.lml) *out.lml = lml;
}

// marginals / marginals_diag: data-free predict recursion, emission-space Gaussian per memory index. One CTA.
// With Rnew != null the emission noise is replaced (replace_observation_noise_cov, missings.jl:35-41).
__global__ void __launch_bounds__(kSeqThreads) k_seq_marginals(const SeqModel md, const SeqOut out) {
    extern __shared__ __align__(16) double smem[];
    const int D = md.D, M = md.M;
    SeqSmem w = seq_carve(smem, D, M);
    for (int i = seq_tid(); i < D; i += kSeqThreads) w.m[i] = md.m0[i];
    for (int e = seq_tid(); e < D * D; e += kSeqThreads) w.P[e] = md.P0[e];
    __syncthreads();
    for (long long k = 0; k < md.T; ++k) {
        const long long t = md.reverse ? md.T - 1 - k : k;
        seq_load_step(md, t, w);
        if (out.Rnew) {
            __syncthreads();
            for (int e = seq_tid(); e < M * M; e += kSeqThreads) {
                const int i = e % M, j = e / M;
                double r;
                if (md.rkind == TGP_R_DENSE) r = out.Rnew[t * out.sRnew + e];
                else if (md.rkind == TGP_R_DIAG) r = i == j ? out.Rnew[t * out.sRnew + i] : 0.0;
                else r = out.Rnew[t * out.sRnew];
                w.R[e] = r;
            }
        }
        __syncthreads();
        if (!md.reverse) {
            seq_predict(md, t, w, w.m, w.P, w.mp, w.Pp);
            for (int i = seq_tid(); i < D; i += kSeqThreads) w.m[i] = w.mp[i];
            for (int e = seq_tid(); e < D * D; e += kSeqThreads) w.P[e] = w.Pp[e];
            __syncthreads();
        }
        // emit predict(x, emission): mean = H m + h, cov = (H Symmetric(P)) H' + R
        seq_symmetrize(w.P, D);
        __syncthreads();
        for (int i = seq_tid(); i < M; i += kSeqThreads) {
            double s = md.h[t * md.sh + i];
            for (int l = 0; l < D; ++l) s = fma(w.H[i + (size_t)M * l], w.m[l], s);
            out.mean[t * M + i] = s;
        }
        seq_mm(w.V, M, w.H, M, false, w.P, D, false, M, D, D);
        __syncthreads();
        if (out.diag) {
            for (int i = seq_tid(); i < M; i += kSeqThreads) {
                double s = w.R[i + (size_t)M * i];
                for (int l = 0; l < D; ++l) s = fma(w.V[i + (size_t)M * l], w.H[i + (size_t)M * l], s);
                out.cov[t * M + i] = s;
            }
        } else {
            seq_mm(out.cov + t * M * M, M, w.V, M, false, w.H, M, true, M, M, D, w.R, M);
        }
        __syncthreads();
        if (md.reverse) {
            seq_predict(md, t, w, w.m, w.P, w.mp, w.Pp);
            for (int i = seq_tid(); i < D; i += kSeqThreads) w.m[i] = w.mp[i];
            for (int e = seq_tid(); e < D * D; e += kSeqThreads) w.P[e] = w.Pp[e];
            __syncthreads();
        }
    }
}

// ---- host drivers ----------------------------------------------------------------------------------------------------------------
static int seq_stage(tgp_ctx* h, const tgp_lgssm* m, const double* y, SeqModel* sm) {
    const size_t D = m->D, M = m->M;
    const size_t rin = m->R_kind == TGP_R_SCALAR ? 1 : (m->R_kind == TGP_R_DIAG ? M : M * M);
    sm->D = m->D; sm->M = m->M; sm->rkind = m->R_kind; sm->reverse = m->ordering == TGP_REVERSE; sm->T = m->T;
    sm->sA = m->sA; sm->sa = m->sa; sm->sQ = m->sQ; sm->sH = m->sH; sm->sh = m->sh; sm->sR = m->sR;
    TGP_TRY(stage_steps(h, m->A, m->sA, m->T, D * D, &sm->A));
    TGP_TRY(stage_steps(h, m->a, m->sa, m->T, D, &sm->a));
    TGP_TRY(stage_steps(h, m->Q, m->sQ, m->T, D * D, &sm->Q));
    TGP_TRY(stage_steps(h, m->H, m->sH, m->T, M * D, &sm->H));
    TGP_TRY(stage_steps(h, m->h, m->sh, m->T, M, &sm->h));
    TGP_TRY(stage_steps(h, m->R, m->sR, m->T, rin, &sm->R));
    TGP_TRY(stage_in(h, m->m0, D, &sm->m0));
    TGP_TRY(stage_in(h, m->P0, D * D, &sm->P0));
    sm->y = nullptr;
    if (y) TGP_TRY(stage_in(h, y, (size_t)m->T * M, &sm->y));
    return TGP_OK;
}

template <class K>
static int seq_launch(tgp_ctx* h, K kernel, const char* name, const SeqModel& sm, const SeqOut& so) {
    const size_t smem = seq_smem_bytes(sm.D, sm.M);
    if (smem > 200 * 1024) return fail(h, TGP_EUNSUPPORTED, "D=%d, M=%d need %zu bytes of shared memory per step: beyond this path", sm.D, sm.M, smem);
    TGP_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TGP_K(h, name);
    kernel<<<1, kSeqThreads, smem, h->stream>>>(sm, so);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

static int seq_finish(tgp_ctx* h, unsigned long long* derr, int64_t T) {
    unsigned long long* perr = (unsigned long long*)h->pinned;
    TGP_CUDA(h, cudaMemcpyAsync(perr, derr, sizeof(*perr), cudaMemcpyDeviceToHost, h->stream));
    h->d2h += 8;
    TGP_TRY(flush_outputs(h));
    TGP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (*perr != ~0ull) return fail(h, TGP_ENOTPD, "covariance not positive definite at time index %lld (0-based)", (long long)*perr);
    (void)T;
    return TGP_OK;
}

static int seq_check(tgp_ctx* h, const tgp_lgssm* m) {
    if (m->D > kSeqMaxDim || m->M > kSeqMaxDim)
        return fail(h, TGP_EUNSUPPORTED, "posterior / marginals of a model with D=%d, M=%d: the step-by-step path takes D, M <= %d", m->D, m->M, kSeqMaxDim);
    return TGP_OK;
}

int seq_posterior(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* G, double* g, double* Sig, double* m_T, double* P_T, double* lml_out) {
    TGP_TRY(seq_check(h, m));
    if ((G || g || Sig) && !(G && g && Sig)) return fail(h, TGP_EINVAL, "G, g and Sig must be all NULL or all non-NULL");
    const int64_t T = m->T, D = m->D;
    SeqModel sm;
    TGP_TRY(seq_stage(h, m, y, &sm));
    SeqOut so{};
    int64_t s1;
    TGP_TRY(stage_out(h, G, D * D, D * D, T, &so.G, &s1));
    TGP_TRY(stage_out(h, g, D, D, T, &so.g, &s1));
    TGP_TRY(stage_out(h, Sig, D * D, D * D, T, &so.Sig, &s1));
    TGP_TRY(stage_out(h, m_T, D, D, 1, &so.mT, &s1));
    TGP_TRY(stage_out(h, P_T, D * D, D * D, 1, &so.PT, &s1));
    TGP_TRY(stage_out(h, lml_out, 1, 1, 1, &so.lml, &s1));
    TGP_TRY(dalloc(h, 1, &so.err));
    TGP_CUDA(h, cudaMemsetAsync(so.err, 0xFF, sizeof(unsigned long long), h->stream));
    TGP_TRY(seq_launch(h, k_seq_posterior, "k_seq_posterior", sm, so));
    return seq_finish(h, so.err, T);
}

int seq_marginals(tgp_ctx* h, const tgp_lgssm* m, double* mean_out, double* cov_out, int diag) {
    TGP_TRY(seq_check(h, m));
    const int64_t T = m->T, M = m->M;
    SeqModel sm;
    TGP_TRY(seq_stage(h, m, nullptr, &sm));
    SeqOut so{};
    int64_t s1;
    TGP_TRY(stage_out(h, mean_out, M, M, T, &so.mean, &s1));
    TGP_TRY(stage_out(h, cov_out, diag ? M : M * M, diag ? M : M * M, T, &so.cov, &s1));
    so.diag = diag;
    TGP_TRY(dalloc(h, 1, &so.err));
    TGP_CUDA(h, cudaMemsetAsync(so.err, 0xFF, sizeof(unsigned long long), h->stream));
    TGP_TRY(seq_launch(h, k_seq_marginals, "k_seq_marginals", sm, so));
    return seq_finish(h, so.err, T);
}

// marginals_diag(replace_observation_noise_cov(posterior(model, y), R_new)): the posterior's (G, g, Sig) go to device scratch, the
// marginals of the reversed model are then emitted from them.
int seq_posterior_marginals(tgp_ctx* h, const tgp_lgssm* m, const double* y, const double* R_new, int64_t sRnew, double* mean_out,
                            double* var_out, double* lml_out) {
    TGP_TRY(seq_check(h, m));
    const int64_t T = m->T, D = m->D, M = m->M;
    const size_t rin = m->R_kind == TGP_R_SCALAR ? 1 : (m->R_kind == TGP_R_DIAG ? M : M * M);
    SeqModel sm;
    TGP_TRY(seq_stage(h, m, y, &sm));
    SeqOut so{};
    int64_t s1;
    TGP_TRY(dalloc(h, (size_t)T * D * D, &so.G));
    TGP_TRY(dalloc(h, (size_t)T * D, &so.g));
    TGP_TRY(dalloc(h, (size_t)T * D * D, &so.Sig));
    TGP_TRY(dalloc(h, (size_t)D, &so.mT));
    TGP_TRY(dalloc(h, (size_t)D * D, &so.PT));
    TGP_TRY(stage_out(h, lml_out, 1, 1, 1, &so.lml, &s1));
    TGP_TRY(dalloc(h, 1, &so.err));
    TGP_CUDA(h, cudaMemsetAsync(so.err, 0xFF, sizeof(unsigned long long), h->stream));
    TGP_TRY(seq_launch(h, k_seq_posterior, "k_seq_posterior", sm, so));
    // the posterior model: reversed ordering, transitions (G, g, Sig), x0 = final filtering state, same emissions with R_new
    SeqModel pm = sm;
    pm.reverse = !sm.reverse;
    pm.A = so.G; pm.sA = D * D;
    pm.a = so.g; pm.sa = D;
    pm.Q = so.Sig; pm.sQ = D * D;
    pm.m0 = so.mT; pm.P0 = so.PT;
    SeqOut mo{};
    TGP_TRY(stage_out(h, mean_out, M, M, T, &mo.mean, &s1));
    TGP_TRY(stage_out(h, var_out, M, M, T, &mo.cov, &s1));
    mo.diag = 1;
    TGP_TRY(stage_steps(h, R_new, sRnew, T, rin, &mo.Rnew));
    mo.sRnew = sRnew;
    mo.err = so.err;
    TGP_TRY(seq_launch(h, k_seq_marginals, "k_seq_marginals", pm, mo));
    return seq_finish(h, so.err, T);
}

}  // namespace tgp

/*
 * tgp_b200.h — C ABI of libtgpb200.so: B200-native (sm_100a) replacement for the LGSSM
 * inference hot path of TemporalGPs.jl (Kalman filter / RTS smoother recursions).
 *
 * The reference has no FFI; its seam is Julia dispatch on the five L2 entry points called by
 * the GP layer. Each function below names the reference function it replaces
 * (file:line relative to the reference tree). A Julia glue module binds these with `ccall`
 * (see INTEGRATION.md and temporalgps.jl_b200/julia/TemporalGPsB200.jl).
 *
 * Conventions
 *  - Plain pointers and sizes only. All matrices are COLUMN-MAJOR (Julia layout).
 *  - Every data pointer may be a HOST or a DEVICE pointer (detected with
 *    cudaPointerGetAttributes). Host inputs are staged to the device, host outputs are copied
 *    back before the call returns. Device outputs are complete when the call returns.
 *  - Per-step arrays use a stride in ELEMENTS between consecutive time steps; stride 0 means
 *    "time-invariant" (Julia `Fill`, src/gp/lti_sde.jl:148-160).
 *  - Return value: 0 (TGP_OK) on success, else a TGP_E* code; tgp_last_error(h) has the text.
 *    No exception ever crosses the boundary (reference throws ErrorException at
 *    src/models/lgssm.jl:202-208 and PosDefException from `cholesky`).
 *  - Calls on one handle are serialised by the caller; distinct handles are independent.
 *  - Missing data never crosses the ABI: the caller applies the reference's transform
 *    (y := 0, R := 1e15, src/models/missings.jl:25-53) and adds the volume compensation.
 *  - There is NO CPU fallback: without a CUDA device tgp_create fails with TGP_ECUDA.
 */
#ifndef TGP_B200_H
#define TGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGP_OK            0
#define TGP_EINVAL        1  /* bad argument / dimension mismatch (lgssm.jl:202-208)          */
#define TGP_ENOTPD        2  /* a covariance that must be factorised is not positive definite  */
#define TGP_ECUDA         3  /* CUDA runtime failure or no device                              */
#define TGP_ENOMEM        4
#define TGP_EUNSUPPORTED  5  /* shape outside what this build supports                         */

#define TGP_FORWARD 0        /* gauss_markov_model.jl:1  — step = predict, then update         */
#define TGP_REVERSE 1        /* gauss_markov_model.jl:3  — step = update, then predict; t=T..1  */

#define TGP_R_SCALAR 0       /* M == 1: R is one variance per step (ScalarOutputLGC, LGC:225)  */
#define TGP_R_DIAG   1       /* R is the M diagonal entries per step                           */
#define TGP_R_DENSE  2       /* R is a dense M x M matrix per step                             */

#define TGP_MAX_D 4096       /* largest latent / observation dimension accepted (dense path)    */

/* algorithm selection for tgp_set_option(h, TGP_OPT_ALGO, v) */
#define TGP_OPT_ALGO         1
#define TGP_ALGO_AUTO        0  /* steady-state path when the model is time-invariant, else scan  */
#define TGP_ALGO_SCAN        1  /* always the general 5-tuple associative scan                 */
#define TGP_OPT_CHUNK        2  /* steps per thread in the scan kernels (0 = auto)              */
#define TGP_OPT_SS_TOL       3  /* (double bits) relative tolerance for steady-state detection  */
#define TGP_OPT_TIMING       4  /* 1: bracket every kernel launch with CUDA events (tgp_get_timing) */
#define TGP_OPT_SS_PREFIX    5  /* steps filtered by the general scan before the steady-state test */
#define TGP_OPT_DENSE_MATH   6  /* arithmetic of the large-state / vector-observation path (tgp_dense.cu)      */
#define TGP_OPT_DEFER_STATUS 7  /* 1: tgp_shard_phase2 calls fold their status into a sticky device block and the next call does
                                 * NOT wait for it (consecutive sharded calls queue back to back on the stream);
                                 * tgp_synchronize() reports and clears the accumulated status                  */
#define TGP_OPT_SHARD_OVERLAP 8  /* 1: the series was scattered WITH AN OVERLAP: on every rank > 0 the TGP_SHARD_HALO observations
                                 * that precede the shard sit in device memory directly before y (y[-TGP_SHARD_HALO .. -1]).
                                 * tgp_shard_logpdf then reads them in place instead of receiving them from rank - 1 over the
                                 * exchange: the step has no inter-GPU dependency at all (the partial results still go to every
                                 * peer's buffer). Set the same value on every rank.                              */
#define TGP_SHARD_HALO       3072
#define TGP_DENSE_F64        0  /* FP64 throughout (reference ArrayStorage(Float64)); own DFMA kernels           */
#define TGP_DENSE_TF32X3     1  /* FP32 storage (reference ArrayStorage(Float32)): covariance algebra on the tcgen05
                                 * tensor cores as 3xTF32 split products with FP32 accumulation in TMEM; innovation
                                 * Cholesky, means and the log-likelihood stay FP64                               */

typedef struct tgp_ctx* tgp_handle;

/*
 * LGSSM descriptor = GaussMarkovModel (gauss_markov_model.jl:20-32) + emissions
 * (lgssm.jl:9-12; StructArray fields .A/.a/.Q of lti_sde.jl:88-109).
 *   x[t] = A[t] x[t-1] + a[t] + N(0, Q[t]);   y[t] = H[t] x[t] + h[t] + N(0, R[t]).
 * Array index t = 0..T-1 is the reference's index 1..T in MEMORY order for both orderings;
 * for TGP_REVERSE the recursion visits t = T-1 down to 0 (gauss_markov_model.jl:38-40).
 */
typedef struct {
    int32_t D;             /* latent dimension                                                  */
    int32_t M;             /* observation dimension per step (1 = ScalarOutputLGC)               */
    int64_t T;             /* number of time steps                                              */
    int32_t ordering;      /* TGP_FORWARD / TGP_REVERSE                                          */
    int32_t R_kind;        /* TGP_R_*                                                            */
    const double* A;  int64_t sA;   /* D x D, stride D*D or 0                                    */
    const double* a;  int64_t sa;   /* D,     stride D   or 0                                    */
    const double* Q;  int64_t sQ;   /* D x D                                                     */
    const double* H;  int64_t sH;   /* M x D  (M == 1: the D entries of the adjoint vector)      */
    const double* h;  int64_t sh;   /* M                                                         */
    const double* R;  int64_t sR;   /* 1 | M | M x M according to R_kind                         */
    const double* m0;               /* D      — x0.m (gaussian.jl:16-19)                         */
    const double* P0;               /* D x D  — x0.P                                             */
} tgp_lgssm;

/* ---- lifetime ------------------------------------------------------------------------- */
int         tgp_create(tgp_handle* out, int device);      /* one handle = one GPU + stream + workspace */
void        tgp_destroy(tgp_handle h);
const char* tgp_last_error(tgp_handle h);                 /* h may be NULL: last create error */
const char* tgp_version(void);
int         tgp_set_option(tgp_handle h, int option, int64_t value);
/* counters since create: kernels launched by this library, bytes copied H2D / D2H */
int         tgp_get_counters(tgp_handle h, int64_t* launches, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* Per-kernel device time accumulated since TGP_OPT_TIMING was switched on (measurement aid for
 * bench.py's roofline leg; synchronises the stream). Writes up to `cap` records sorted by total
 * time, returns the number of distinct kernels. names[i] points to a static string. */
int         tgp_get_timing(tgp_handle h, int cap, const char** names, double* total_ms, int64_t* calls);
/* run subsequent work of this handle on a caller-owned cudaStream_t (0 = handle's own) */
int         tgp_set_stream(tgp_handle h, void* cuda_stream);

/* ---- hot path --------------------------------------------------------------------------
 * tgp_logpdf   replaces logpdf(::LGSSM, y)            src/models/lgssm.jl:147-165
 *              (scan_emit + step_logpdf: predict LGC:46-52, posterior_and_lml LGC:247-257 /
 *              129-141). lml_per_step (T doubles, may be NULL) receives the emitted `lmls`.
 *              Scalar observations with D in {1..6, 8, 10} run the parallel-in-time scan kernels; every other
 *              shape (vector observations M > 1 = SmallOutputLGC, any R_kind; larger D) runs the dense
 *              step-by-step path (tgp_dense.cu), y then being T x M (M fastest).
 */
int tgp_logpdf(tgp_handle h, const tgp_lgssm* model, const double* y,
               double* lml_out, double* lml_per_step);

/* tgp_filter   replaces _filter(::LGSSM, y)           src/models/lgssm.jl:171-187
 *              m_f[t*s_m + i], P_f[t*s_P + i + D*j]: strides in elements (s_m >= D, s_P >= D*D),
 *              so a Julia Vector{Gaussian{SVector{D},SMatrix{D,D}}} (records of D+D*D doubles)
 *              is written in place with m_f = base, P_f = base + D, s_m = s_P = D + D*D.
 *              Either output may be NULL. lml_out may be NULL.
 */
int tgp_filter(tgp_handle h, const tgp_lgssm* model, const double* y,
               double* m_f, int64_t s_m, double* P_f, int64_t s_P, double* lml_out);

/* tgp_posterior replaces posterior(::LGSSM, y)        src/models/lgssm.jl:193-238
 *              (step_posterior + invert_dynamics, jitter 1e-10). Emits the reverse-time
 *              dynamics G (T x D x D), g (T x D), Sig (T x D x D), contiguous per step, and the
 *              final filtering distribution (m_T, P_T) = x0 of the returned Reverse model.
 */
int tgp_posterior(tgp_handle h, const tgp_lgssm* model, const double* y,
                  double* G, double* g, double* Sig, double* m_T, double* P_T);

/* tgp_marginals replaces marginals(::LGSSM)           src/models/lgssm.jl:99-115
 *              data-free predict recursion; emits the emission-space marginal per step:
 *              mean_out (T x M) and cov_out (T x M x M) (M == 1: the variance).
 */
int tgp_marginals(tgp_handle h, const tgp_lgssm* model, double* mean_out, double* cov_out);

/* tgp_marginals_diag replaces marginals_diag(::LGSSM)  src/models/lgssm.jl:125-141
 *              (step_marginals_diag with predict_marginals, LGC:63-68): the same recursion emitting only the diagonal of the
 *              emission-space covariance: mean_out, var_out are T x M.
 */
int tgp_marginals_diag(tgp_handle h, const tgp_lgssm* model, double* mean_out, double* var_out);

/* tgp_lti_components replaces broadcast_components((F, q, H), x0, t::AbstractVector, storage)   src/gp/lti_sde.jl:136-147
 *              — the model construction of an IRREGULAR time grid, on the device: with t' = vcat(t[0] - 1, t),
 *                  A[i] = exp(F * (t'[i+1] - t'[i])),   Q[i] = P - A[i] P A[i]'      (P read as Symmetric: upper triangle)
 *              written as T x D x D arrays, column-major per step — exactly what tgp_lgssm.A / .Q take with sA = sQ = D * D, so a
 *              caller holding (F, P, t) ships 8 B/step (t) instead of 16 D^2 B/step and spends no host matrix exponentials.
 *              F, P (and F0) are D x D column-major on the HOST; t (T doubles) and the outputs may be host or device. F0, if
 *              non-NULL, is the drift of the FIRST transition (dt = 1): a kernel stretched in time (lti_sde.jl:361-373) passes
 *              its stretched drift as F and the unstretched one as F0, block by block for a sum of kernels (lti_sde.jl:386-418).
 *              Device outputs are stream-ordered (usable by the next call on this handle without a synchronisation).
 *              Matrix exponential: scaling and squaring with a degree-12 Taylor polynomial at |F dt| / 2^s <= 1/4.
 */
int tgp_lti_components(tgp_handle h, int32_t D, int64_t T, const double* F, const double* F0, const double* P, const double* t,
                       double* A_out, double* Q_out);

/* tgp_posterior_marginals fuses the chain used by marginals(::FinitePosteriorLTISDE) at the
 *              training inputs (src/gp/posterior_lti_sde.jl:27-36):
 *              posterior(model, y) -> replace_observation_noise_cov(., R_new) -> marginals(.)
 *              -> diagonal. The (G,g,Sig) dynamics are never materialised (scalar observations). R_new is per step with
 *              stride sRnew (0 = constant) and the model's R_kind. Outputs mean_out, var_out: T x M.
 *              lml_out (nullable) receives logpdf(model, y) as a by-product of the forward pass.
 */
int tgp_posterior_marginals(tgp_handle h, const tgp_lgssm* model, const double* y,
                            const double* R_new, int64_t sRnew,
                            double* mean_out, double* var_out, double* lml_out);

/* Test hook (no reference analogue): C (Mx x N) = X' Y for HOST column-major float matrices X (K x Mx), Y (K x N),
 * computed by the tcgen05 3xTF32 contraction kernel alone; symmetric != 0 exercises the mirrored upper-triangle epilogue. */
int tgp_debug_tc_gemm(tgp_handle h, int K, int Mx, int N, const float* X, const float* Y, float* C, int symmetric);

/* ---- time-sharded multi-GPU path (no reference analogue) -------------------
 * A scan element is (A, b, C, eta, J): 3*D*D + 2*D doubles, matrices column-major, in that
 * order. tgp_shard_reduce folds a whole shard of steps into ONE element (phase 1). The caller
 * all-gathers the elements of all ranks (NCCL, 264 B per rank at D = 3), then
 * tgp_shard_prefix applies elements 0..rank-1 to x0 to obtain the filtering distribution
 * entering this rank's shard; phase 2 is the ordinary tgp_logpdf / tgp_filter call on the shard
 * with (m0, P0) := that state.
 */
int tgp_elem_size(int D);                                  /* 3*D*D + 2*D */
int tgp_shard_reduce(tgp_handle h, const tgp_lgssm* shard, const double* y, double* elem_out);
/* tgp_shard_prefix is host-side arithmetic only; h may be NULL. */
int tgp_shard_prefix(tgp_handle h, int D, int n_elems, const double* elems /*host*/,
                     const double* m0, const double* P0, double* m_in, double* P_in /*host*/);

/* Steady-state variant for TIME-INVARIANT shards (RegularSpacing + homoscedastic noise; every shard >= 65536 steps):
 * the exchange shrinks to one affine record (Phi_shard, Z_shard) = D*D + D doubles per rank (96 B at D = 3) and no
 * host round trip is needed: tgp_shard_phase1 enqueues the zero-state pass and leaves this rank's record in
 * xchg_out (DEVICE memory) without synchronising; the caller all-gathers the records on the same stream
 * (ncclAllGather); tgp_shard_phase2 filters the shard from the mean folded out of the records of the ranks before it
 * and writes the shard's log-likelihood to lml_partial (DEVICE); the caller sums the partials (ncclAllReduce).
 * y must stay valid (and, if it is a host pointer, unchanged) between the two calls. */
int tgp_shard_xchg_size(int D);                            /* D*D + D */
int tgp_shard_phase1(tgp_handle h, const tgp_lgssm* shard, const double* y, int rank, int world, double* xchg_out);
int tgp_shard_phase2(tgp_handle h, const double* xchg_all, double* lml_partial);   /* returns WITHOUT synchronising */
/* Waits for the handle's stream and returns the status of work that was enqueued without synchronisation
 * (tgp_shard_phase2: TGP_ENOTPD / TGP_EUNSUPPORTED "not converged"). The next call on the handle does the same. */
int tgp_synchronize(tgp_handle h);

/* Peer-memory exchange buffers for the sharded path (one process per GPU, NVLink / NVSwitch P2P): tgp_shard_logpdf's kernel stores
 * straight into its peers' buffers (no NCCL call, no host round trip). tgp_xchg_create allocates this rank's buffer and returns its
 * 64-byte CUDA IPC handle; the caller gathers the handles of all ranks (any transport) and passes them, rank-ordered, to
 * tgp_xchg_open. slot_doubles: size of the general-purpose record slots kept in the buffer (16 is enough). */
int tgp_xchg_create(tgp_handle h, int rank, int world, int slot_doubles, void* ipc_handle_out);
int tgp_xchg_open(tgp_handle h, const void* ipc_handles_all);

/* ONE-LAUNCH SHARDED LOGPDF (the path BASELINE config 4 runs; needs an opened exchange when world > 1). `shard` describes this
 * rank's steps [rank*T, ...) of ONE Forward, time-invariant, scalar-observation series (D <= 4); its (m0, P0) is the prior of the
 * whole series. The call ENQUEUES one kernel and returns: rank 0 runs the covariance transient, every other rank starts from pass A
 * over the <= 3072 observations that precede its shard, which its predecessor's kernel stores into this rank's exchange buffer over
 * NVLink at the START of its own run — shards never wait for each other's results. With TGP_OPT_SHARD_OVERLAP the caller has stored
 * those TGP_SHARD_HALO observations directly before y instead (y[-3072 .. -1], device memory): nothing is pushed, no rank waits for
 * another at all. Each kernel ships its shard's log-likelihood
 * into every rank's buffer. tgp_shard_result enqueues the fixed-order sum over ranks for the LAST tgp_shard_logpdf into lml_total
 * (device pointer: stream-ordered; host pointer: synchronises) — call it when the value is wanted, not necessarily per step.
 * Every rank must issue the same sequence of tgp_shard_logpdf calls. Returns TGP_EUNSUPPORTED (nothing enqueued, the same
 * decision on every rank) for models outside the path's range: use tgp_shard_phase1/2 or tgp_shard_reduce/prefix then. */
int tgp_shard_logpdf(tgp_handle h, const tgp_lgssm* shard, const double* y, int rank, int world);
int tgp_shard_result(tgp_handle h, double* lml_total);
/* This rank's own term of that sum: log p(y_shard | everything before the shard) of the LAST tgp_shard_logpdf (device or host pointer). */
int tgp_shard_partial(tgp_handle h, double* lml_shard);

#ifdef __cplusplus
}
#endif
#endif /* TGP_B200_H */

// tgp_dense.cu — large-state / vector-observation path (SmallOutputLGC, LGC:129-141; BASELINE config 5: separable
// space-time GPs, D = 768, M = 256). The state is too large for a register-resident scan element, and every step is
// a genuine dense contraction, so the recursion runs step by step on ONE stream exactly as the reference executes it:
//     predict   m <- A m + a;  P <- (A * Symmetric(P)) * A' + Q                                  (LGC:46-52)
//     update    V = H P;  S = chol(Symmetric(V H' + R));  B = U' \ V;  alpha = U' \ (y - H m - h)
//               lml = -(M log 2pi + logdet S + alpha'alpha)/2;  m += B' alpha;  P -= B'B          (LGC:129-141)
// Every kernel on this path is ours: the FP64 products are tgp_dense_f64.cuh (tiled DFMA GEMM, GEMV, triangular solve), the
// Cholesky, the residual / likelihood and the bookkeeping kernels are below; no library call is left (round 1 used cuBLAS here). For time-invariant models one step is captured into a CUDA graph (a device-side
// step counter indexes y / outputs) and replayed T times, so the host issues one launch per step instead of ~14.
#include <algorithm>
#include <cstdlib>

#include "tgp_ctx.cuh"
#include "tgp_dense_f64.cuh"

namespace tgp {

// one of the FP64 product kernels: name it for the timing table, launch, count
#define TGP_F64(h, name, call)                                                                                     \
    do {                                                                                                           \
        TGP_K(h, name);                                                                                            \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return fail(h, TGP_ECUDA, "%s failed: %s (%s:%d)", name, cudaGetErrorString(e_), __FILE__, __LINE__); \
        ++(h)->launches;                                                                                           \
        tgp::prof_end(h);                                                                                          \
    } while (0)

constexpr double kLog2PiD = 1.8378770664093454835606594728112;

// Upper Cholesky factor in place (S = U'U, column-major, upper triangle read), blocked right-looking with NB = 32:
//   k_chol_panel2 (1 CTA)    diagonal block factored by one warp, its inverse, then U_kk^-T applied to the block row;
//   k_chol_trail  (many CTAs) S[i, c] -= sum_r U[k0+r, i] U[k0+r, c] for k0+nb <= i <= c < M.
// 2*ceil(M/32) - 1 launches per factorisation (inside the captured step graph). err gets the time step on failure.
constexpr int kCholNB = 32;

// ---- faster panel: the only serial part is the 32 x 32 diagonal block ------------------------------------------------------
// k_chol_panel2 (1 CTA): diagonal block S_kk -> U_kk by ONE warp (lane = column, matrix in shared memory, warp-synchronous),
// its inverse X = U_kk^-1 by the same warp (lane = column, back substitution), then every warp applies X' to the block row:
// U[k-block, c] = X' S[k-block, c] for the columns right of the block (a 32-long dot per entry, no serial dependency).
// X is also stored (Dinv, 32 x 32 per block): k_tri_inv2 builds U^-1 from these.
__global__ void __launch_bounds__(256) k_chol_panel2(double* __restrict__ S, int M, int k0, const long long* __restrict__ step,
                                                      unsigned long long* __restrict__ err, double* __restrict__ Dinv) {
    __shared__ double A[32][33];
    __shared__ double X[32][33];
    __shared__ double rinv[32];
    __shared__ double colbuf[8][4][32];
    const int nb = min(32, M - k0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < 1024; e += 256) {
        const int r = e & 31, c = e >> 5;
        A[r][c] = (r < nb && c < nb) ? S[(size_t)(k0 + r) + (size_t)M * (k0 + c)] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
        // lane c keeps column c of the block in REGISTERS (all indices compile-time after unrolling); only row j of U is
        // exchanged through shared memory at step j. No shared-memory read-modify-write in the inner loop.
        const int c = lane;
        double col[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) col[r] = A[r][c];
        double* rowbuf = colbuf[0][0];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            double d = __shfl_sync(0xffffffffu, col[j], j);           // pivot S_jj - sum_k U_kj^2, held by lane j
            if (!(d > 0.0)) { if (lane == 0) atomicMin(err, (unsigned long long)*step); d = 1.0; }
            const double inv = rsqrt(d);
            const double ujc = c == j ? d * inv : col[j] * inv;       // U[j][c] for c >= j
            col[j] = ujc;
            rowbuf[c] = ujc;
            if (c == j) rinv[j] = inv;
            __syncwarp();
#pragma unroll
            for (int r = j + 1; r < 32; ++r)
                if (r <= c) col[r] = fma(-rowbuf[r], ujc, col[r]);    // S[r][c] -= U[j][r] U[j][c]
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) A[r][c] = r <= c ? col[r] : 0.0;  // U_kk (upper)
        __syncwarp();
        // X = U^-1, column c in registers: X[c][c] = 1/U[c][c]; X[i][c] = -(sum_{i < l <= c} U[i][l] X[l][c]) / U[i][i]
        double xc[32];
#pragma unroll
        for (int i = 31; i >= 0; --i) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int l = i + 1; l < 32; ++l) {
                const double u = A[i][l];                              // same address for every lane: broadcast
                if (l <= c) { if (l & 1) s1 = fma(u, xc[l], s1); else s0 = fma(u, xc[l], s0); }
            }
            xc[i] = i == c ? rinv[i] : (i < c ? -(s0 + s1) * rinv[i] : 0.0);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) X[i][c] = xc[i];
    }
    __syncthreads();
    for (int e = tid; e < 1024; e += 256) {
        const int r = e & 31, c = e >> 5;
        if (r < nb && c < nb) S[(size_t)(k0 + r) + (size_t)M * (k0 + c)] = (c >= r) ? A[r][c] : 0.0;
        Dinv[(size_t)(k0 / 32) * 1024 + r + 32 * c] = X[r][c];
    }
    // block row: a warp takes 4 columns at a time (4 independent FMA chains), lane = row i of the block:
    // U[k0 + i, col] = sum_{r <= i} X[r][i] S[k0 + r, col]
    const int W = M - k0 - nb;
    for (int c4 = warp * 4; c4 < W; c4 += 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int cc = c4 + q;
            colbuf[warp][q][lane] = (cc < W && lane < nb) ? S[(size_t)(k0 + lane) + (size_t)M * (k0 + nb + cc)] : 0.0;
        }
        __syncwarp();
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = 0; r <= lane; ++r) {
            const double xr = X[r][lane];
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] = fma(xr, colbuf[warp][q][r], s[q]);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int cc = c4 + q;
            if (cc < W && lane < nb) S[(size_t)(k0 + lane) + (size_t)M * (k0 + nb + cc)] = s[q];
        }
    }
}

__global__ void __launch_bounds__(256) k_chol_trail(double* __restrict__ S, int M, int k0, int nb) {
    const int n = M - k0 - nb;                 // trailing size
    const int base = k0 + nb;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)n * n; e += (long long)gridDim.x * blockDim.x) {
        const int i = base + (int)(e % n), c = base + (int)(e / n);
        if (i > c) { continue; }
        const double* ui = S + (size_t)k0 + (size_t)M * i;
        const double* uc = S + (size_t)k0 + (size_t)M * c;
        double acc = 0.0;
#pragma unroll 8
        for (int r = 0; r < nb; ++r) acc = fma(ui[r], uc[r], acc);
        S[(size_t)i + (size_t)M * c] -= acc;
    }
}

__global__ void k_clear_lower(double* __restrict__ S, int M) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)M * M; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % M), c = (int)(e / M);
        if (r > c) S[e] = 0.0;
    }
}

// S <- R_t expanded to a dense M x M matrix (R_kind: 0 scalar, 1 diag, 2 dense).
__global__ void k_expand_R(double* __restrict__ S, int M, const double* __restrict__ R, long long sR, int R_kind,
                           const long long* __restrict__ step) {
    const double* Rt = R + *step * sR;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)M * M; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % M), c = (int)(e / M);
        double v = 0.0;
        if (R_kind == 2) v = Rt[e];
        else if (r == c) v = R_kind == 1 ? Rt[r] : Rt[0];
        S[e] = v;
    }
}

// r <- y_t - h_t   (H m is subtracted by a gemv afterwards)
__global__ void k_residual0(double* __restrict__ r, int M, const double* __restrict__ y, const double* __restrict__ hh, long long sh,
                            const long long* __restrict__ step) {
    const long long t = *step;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) r[i] = y[t * M + i] - hh[t * sh + i];
}

// lml_t = -(M log 2pi + 2 sum log U_ii + alpha'alpha)/2 ; accumulates the total; advances nothing.
__global__ void __launch_bounds__(256) k_lml(const double* __restrict__ U, const double* __restrict__ alpha, int M, double* __restrict__ lml_steps,
                                             double* __restrict__ lml_total, const long long* __restrict__ step) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < M; i += 256) s += 2.0 * log(U[i + (size_t)M * i]) + alpha[i] * alpha[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double l = -0.5 * (M * kLog2PiD + sh[0]);
        if (lml_steps) lml_steps[*step] = l;
        *lml_total += l;
    }
}

// per-step transition vector: mt <- a_t (before the gemv accumulates A m into it); Pn <- Q_t
__global__ void k_load_aQ(double* __restrict__ mt, double* __restrict__ Pn, int D, const double* __restrict__ a, long long sa,
                          const double* __restrict__ Q, long long sQ, const long long* __restrict__ step) {
    const long long t = *step;
    const long long n = (long long)D * D;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n + D; e += (long long)gridDim.x * blockDim.x) {
        if (e < n) { if (Pn) Pn[e] = Q[t * sQ + e]; }
        else mt[e - n] = a[t * sa + (e - n)];
    }
}

__global__ void k_emit_state(const double* __restrict__ m, const double* __restrict__ P, int D, double* __restrict__ m_f, long long s_m,
                             double* __restrict__ P_f, long long s_P, const long long* __restrict__ step) {
    const long long t = *step;
    const long long n = (long long)D * D;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n + D; e += (long long)gridDim.x * blockDim.x) {
        if (e < n) { if (P_f) P_f[t * s_P + e] = P[e]; }
        else if (m_f) m_f[t * s_m + (e - n)] = m[e - n];
    }
}

__global__ void k_advance(long long* step, long long delta) { *step += delta; }

// Steady-state detection for time-invariant models: conv[0] = max |P - Pprev|, conv[1] = max |P| (bit patterns of
// non-negative doubles order like integers), and Pprev <- P for the next step.
__global__ void k_conv_check(const double* __restrict__ P, double* __restrict__ Pprev, long long n, unsigned long long* __restrict__ conv) {
    double md = 0.0, ma = 0.0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const double x = P[e];
        md = fmax(md, fabs(x - Pprev[e]));
        ma = fmax(ma, fabs(x));
        Pprev[e] = x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        md = fmax(md, __shfl_xor_sync(0xffffffffu, md, off));
        ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(conv, (unsigned long long)__double_as_longlong(md));
        atomicMax(conv + 1, (unsigned long long)__double_as_longlong(ma));
    }
}

__global__ void k_advance_conv64(long long* step, long long delta, unsigned long long* conv, double tol, long long* ss_at) {
    if (conv) {
        const double d = __longlong_as_double((long long)conv[0]), a = __longlong_as_double((long long)conv[1]);
        conv[2] = conv[0];   // kept for diagnostics (TGP_DEBUG): the last measured |dP|, |P|
        conv[3] = conv[1];
        if (*ss_at < 0 && a > 0.0 && d <= tol * a) *ss_at = *step;
        conv[0] = 0ull;
        conv[1] = 0ull;
    }
    *step += delta;
}

struct DenseWs {
    double *m, *mt, *P, *Pn, *T1, *V, *S, *B, *r;
    long long* step;
    double* lml;
    unsigned long long* err;
    double* Dinv = nullptr;              // inverses of the 32 x 32 diagonal blocks of U (k_chol_panel2)
    double* W = nullptr;                 // U^-1 (k_tri_inv_f64): B = W'V, alpha = W'r. nullptr: M too large, triangular solves instead
    double* al = nullptr;                // alpha (M)
    double* gws = nullptr;               // split-K workspace of the products
    size_t gws_doubles = 0;
    double* Pprev = nullptr;             // steady-state detection (time-invariant models only)
    unsigned long long* conv = nullptr;
    long long* ss_at = nullptr;
};

// frozen: the covariance recursion has reached its fixed point; only the mean / likelihood part of the step runs.
static int dense_step(tgp_ctx* h, const tgp_lgssm& d, const double* dy, const DenseWs& w, long long t /* host index or -1: use device counter */,
                      bool graph_mode, double* lml_steps, double* m_f, int64_t s_m, double* P_f, int64_t s_P, bool frozen = false) {
    const int D = d.D, M = d.M;
    cudaStream_t st = h->stream;
    // In graph mode the model is time-invariant (all strides 0), so parameter pointers are fixed and only y / outputs
    // are indexed through the device counter; otherwise t is known on the host.
    const long long tt = graph_mode ? 0 : t;
    const double* A = d.A + tt * d.sA;
    const double* H = d.H + tt * d.sH;
    const bool rev = d.ordering == TGP_REVERSE;
    const int nb = (int)std::min<long long>(((long long)D * D + D + 255) / 256, 1184);

    auto predict = [&]() -> int {
        TGP_K(h, "dense:k_load_aQ");
        k_load_aQ<<<nb, 256, 0, st>>>(w.mt, frozen ? nullptr : w.Pn, D, d.a, d.sa, d.Q, d.sQ, w.step);
        TGP_LAUNCH_CHECK(h);
        TGP_F64(h, "dense:k_dgemv_n", dgemv(st, Op::N, D, D, 1.0, A, D, w.m, 1.0, w.mt));                          // mt = A m + a
        if (frozen) {
            TGP_CUDA(h, cudaMemcpyAsync(w.m, w.mt, sizeof(double) * D, cudaMemcpyDeviceToDevice, st));
            h->launches += 1;
            return TGP_OK;
        }
        TGP_F64(h, "dense:k_dgemm", dgemm(st, Op::N, Op::N, true, D, D, D, 1.0, A, D, w.P, D, 0.0, w.T1, D, w.gws, w.gws_doubles));          // A * Symmetric(P)
        TGP_F64(h, "dense:k_dgemm", dgemm(st, Op::N, Op::T, false, D, D, D, 1.0, w.T1, D, A, D, 1.0, w.Pn, D, w.gws, w.gws_doubles));        // Pn = T1 A' + Q
        TGP_CUDA(h, cudaMemcpyAsync(w.m, w.mt, sizeof(double) * D, cudaMemcpyDeviceToDevice, st));
        TGP_CUDA(h, cudaMemcpyAsync(w.P, w.Pn, sizeof(double) * D * D, cudaMemcpyDeviceToDevice, st));
        h->launches += 2;
        return TGP_OK;
    };
    auto update = [&]() -> int {
      if (!frozen) {
        TGP_F64(h, "dense:k_dgemm", dgemm(st, Op::N, Op::N, false, M, D, D, 1.0, H, M, w.P, D, 0.0, w.V, M, w.gws, w.gws_doubles));          // V = H P
        TGP_K(h, "dense:k_expand_R");
        k_expand_R<<<std::min((M * M + 255) / 256, 1184), 256, 0, st>>>(w.S, M, d.R, d.sR, d.R_kind, w.step);
        TGP_LAUNCH_CHECK(h);
        TGP_F64(h, "dense:k_dgemm", dgemm(st, Op::N, Op::T, false, M, M, D, 1.0, w.V, M, H, M, 1.0, w.S, M, w.gws, w.gws_doubles));          // S = V H' + R
        for (int k0 = 0; k0 < M; k0 += kCholNB) {
            const int nbk = std::min(kCholNB, M - k0);
            TGP_K(h, "dense:k_chol_panel2");
            k_chol_panel2<<<1, 256, 0, st>>>(w.S, M, k0, w.step, w.err, w.Dinv);
            TGP_LAUNCH_CHECK(h);
            const int n = M - k0 - nbk;
            if (n > 0) {
                TGP_K(h, "dense:k_chol_trail");
                k_chol_trail<<<(int)std::min<long long>(((long long)n * n + 255) / 256, 1184), 256, 0, st>>>(w.S, M, k0, nbk);
                TGP_LAUNCH_CHECK(h);
            }
        }
        TGP_K(h, "dense:k_clear_lower");
        k_clear_lower<<<std::min((M * M + 255) / 256, 1184), 256, 0, st>>>(w.S, M);
        TGP_LAUNCH_CHECK(h);
        if (w.W) {
            TGP_K(h, "dense:k_tri_inv_f64");
            k_tri_inv_f64<<<(M + 7) / 8, 256, tri_inv_smem(M), st>>>(w.S, M, w.Dinv, w.W);
            TGP_LAUNCH_CHECK(h);
            TGP_F64(h, "dense:k_dgemm", dgemm(st, Op::T, Op::N, false, M, D, M, 1.0, w.W, M, w.V, M, 0.0, w.B, M, w.gws, w.gws_doubles));  // B = U' \ V = W' V
        } else {
            TGP_CUDA(h, cudaMemcpyAsync(w.B, w.V, sizeof(double) * M * D, cudaMemcpyDeviceToDevice, st));
            ++h->launches;
            TGP_F64(h, "dense:k_trsm_ut", trsm_ut(st, M, D, w.S, M, w.B, M));                                          // B = U' \ V
        }
      }
        TGP_K(h, "dense:k_residual0");
        k_residual0<<<(M + 255) / 256, 256, 0, st>>>(w.r, M, dy, d.h, d.sh, w.step);
        TGP_LAUNCH_CHECK(h);
        TGP_F64(h, "dense:k_dgemv_n", dgemv(st, Op::N, M, D, -1.0, H, M, w.m, 1.0, w.r));                              // r = y - h - H m
        const double* alpha = w.r;
        if (w.W) {
            TGP_F64(h, "dense:k_dgemv_t", dgemv(st, Op::T, M, M, 1.0, w.W, M, w.r, 0.0, w.al));                        // alpha = U' \ r = W' r
            alpha = w.al;
        } else {
            TGP_F64(h, "dense:k_trsm_ut", trsm_ut(st, M, 1, w.S, M, w.r, M));                                          // alpha = U' \ r
        }
        TGP_K(h, "dense:k_lml");
        k_lml<<<1, 256, 0, st>>>(w.S, alpha, M, lml_steps, w.lml, w.step);
        TGP_LAUNCH_CHECK(h);
        TGP_F64(h, "dense:k_dgemv_t", dgemv(st, Op::T, M, D, 1.0, w.B, M, alpha, 1.0, w.m));                           // m += B' alpha
        if (!frozen) {
            TGP_F64(h, "dense:k_dgemm", dgemm(st, Op::T, Op::N, false, D, D, M, -1.0, w.B, M, w.B, M, 1.0, w.P, D, w.gws, w.gws_doubles));   // P -= B'B
            if (w.conv) {
                TGP_K(h, "dense:k_conv_check");
                k_conv_check<<<nb, 256, 0, st>>>(w.P, w.Pprev, (long long)D * D, w.conv);
                TGP_LAUNCH_CHECK(h);
            }
        }
        if (m_f || P_f) {
            TGP_K(h, "dense:k_emit_state");
            k_emit_state<<<nb, 256, 0, st>>>(w.m, w.P, D, m_f, s_m, P_f, s_P, w.step);
            TGP_LAUNCH_CHECK(h);
        }
        return TGP_OK;
    };
    if (!rev) { TGP_TRY(predict()); TGP_TRY(update()); }
    else      { TGP_TRY(update()); TGP_TRY(predict()); }
    TGP_K(h, "dense:k_advance");
    k_advance_conv64<<<1, 1, 0, st>>>(w.step, rev ? -1 : 1, frozen ? nullptr : w.conv, h->ss_tol, w.ss_at);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

int dense_filter_tc(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* lml_out, double* lml_steps_user, double* m_f_user, int64_t s_m,
                    double* P_f_user, int64_t s_P);

// Entry point used by tgp_logpdf / tgp_filter for shapes outside the small-state scan kernels.
int dense_filter(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* lml_out, double* lml_steps_user, double* m_f_user, int64_t s_m,
                 double* P_f_user, int64_t s_P) {
    if (h->dense_math == TGP_DENSE_TF32X3) return dense_filter_tc(h, m, y, lml_out, lml_steps_user, m_f_user, s_m, P_f_user, s_P);
    const int D = m->D, M = m->M;
    const int64_t T = m->T;
    cudaStream_t st = h->stream;
    if ((size_t)M * sizeof(double) > 48 * 1024)
        return fail(h, TGP_EUNSUPPORTED, "observation dimension M=%d too large for the FP64 triangular solve (M <= 6144)", M);

    // stage the model
    tgp_lgssm d = *m;
    const size_t rin = m->R_kind == TGP_R_SCALAR ? 1 : (m->R_kind == TGP_R_DIAG ? (size_t)M : (size_t)M * M);
    TGP_TRY(stage_steps(h, m->A, m->sA, T, (size_t)D * D, &d.A));
    TGP_TRY(stage_steps(h, m->a, m->sa, T, D, &d.a));
    TGP_TRY(stage_steps(h, m->Q, m->sQ, T, (size_t)D * D, &d.Q));
    TGP_TRY(stage_steps(h, m->H, m->sH, T, (size_t)M * D, &d.H));
    TGP_TRY(stage_steps(h, m->h, m->sh, T, M, &d.h));
    TGP_TRY(stage_steps(h, m->R, m->sR, T, rin, &d.R));
    TGP_TRY(stage_in(h, m->m0, D, &d.m0));
    TGP_TRY(stage_in(h, m->P0, (size_t)D * D, &d.P0));
    const double* dy;
    TGP_TRY(stage_in(h, y, (size_t)T * M, &dy));
    double *lml_steps, *m_f, *P_f;
    int64_t ds, dsm, dsP;
    TGP_TRY(stage_out(h, lml_steps_user, 1, 1, T, &lml_steps, &ds));
    TGP_TRY(stage_out(h, m_f_user, D, s_m, T, &m_f, &dsm));
    TGP_TRY(stage_out(h, P_f_user, (size_t)D * D, s_P, T, &P_f, &dsP));

    DenseWs w;
    TGP_TRY(dalloc(h, D, &w.m));
    TGP_TRY(dalloc(h, D, &w.mt));
    TGP_TRY(dalloc(h, (size_t)D * D, &w.P));
    TGP_TRY(dalloc(h, (size_t)D * D, &w.Pn));
    TGP_TRY(dalloc(h, (size_t)D * D, &w.T1));
    TGP_TRY(dalloc(h, (size_t)M * D, &w.V));
    TGP_TRY(dalloc(h, (size_t)M * M, &w.S));
    TGP_TRY(dalloc(h, (size_t)M * D, &w.B));
    TGP_TRY(dalloc(h, M, &w.r));
    TGP_TRY(dalloc(h, 1, &w.step));
    TGP_TRY(dalloc(h, 1, &w.lml));
    TGP_TRY(dalloc(h, 1, &w.err));
    TGP_TRY(dalloc(h, (size_t)((M + 31) / 32) * 1024, &w.Dinv));
    TGP_TRY(dalloc(h, M, &w.al));
    if (tri_inv_smem(M) <= 200 * 1024) {
        TGP_TRY(dalloc(h, (size_t)M * M, &w.W));
        if (tri_inv_smem(M) > 48 * 1024)
            TGP_CUDA(h, cudaFuncSetAttribute(k_tri_inv_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_inv_smem(M)));
    }
    w.gws_doubles = std::min<size_t>((size_t)8 * D * D, (size_t)1 << 25);       // split-K partial products (<= 256 MB)
    TGP_TRY(dalloc(h, w.gws_doubles, &w.gws));
    const bool ti = !(m->sA | m->sa | m->sQ | m->sH | m->sh | m->sR);
    // Reverse ordering ends a step with predict(): after the freeze w.P would hold the PREDICTED covariance, which is not what
    // P_f emits — so a Reverse model that emits P_f runs every step in full.
    const bool may_freeze = !(m->ordering == TGP_REVERSE && P_f != nullptr);
    if (ti && h->algo == TGP_ALGO_AUTO && may_freeze) {
        TGP_TRY(dalloc(h, (size_t)D * D, &w.Pprev));
        TGP_TRY(dalloc(h, 4, &w.conv));
        TGP_TRY(dalloc(h, 1, &w.ss_at));
        TGP_CUDA(h, cudaMemcpyAsync(w.Pprev, d.P0, sizeof(double) * D * D, cudaMemcpyDeviceToDevice, st));
        TGP_CUDA(h, cudaMemsetAsync(w.conv, 0, 4 * sizeof(unsigned long long), st));
        TGP_CUDA(h, cudaMemsetAsync(w.ss_at, 0xFF, sizeof(long long), st));
    }
    TGP_CUDA(h, cudaMemcpyAsync(w.m, d.m0, sizeof(double) * D, cudaMemcpyDeviceToDevice, st));
    TGP_CUDA(h, cudaMemcpyAsync(w.P, d.P0, sizeof(double) * D * D, cudaMemcpyDeviceToDevice, st));
    TGP_CUDA(h, cudaMemsetAsync(w.lml, 0, sizeof(double), st));
    TGP_CUDA(h, cudaMemsetAsync(w.err, 0xFF, sizeof(unsigned long long), st));
    const bool rev = m->ordering == TGP_REVERSE;
    const long long t0 = rev ? T - 1 : 0;
    TGP_CUDA(h, cudaMemcpyAsync(w.step, &t0, sizeof(long long), cudaMemcpyHostToDevice, st));
    TGP_CUDA(h, cudaStreamSynchronize(st));   // t0 is a stack variable

    if (ti && T >= 8 && !h->timing) {
        // Capture ONE step, replay it. Every kPoll (4) steps the host looks at the steady-state word; once the covariance
        // recursion has converged (max |P_t - P_{t-1}| <= ss_tol max |P_t|, tested on the device) the remaining steps replay
        // the mean-only graph: same arithmetic with P, S, U, B frozen at their limit.
        cudaGraph_t graph[2] = {nullptr, nullptr};
        cudaGraphExec_t exec[2] = {nullptr, nullptr};
        int64_t per_replay[2] = {0, 0};
        for (int fz = 0; fz < (w.conv ? 2 : 1); ++fz) {
            TGP_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int64_t l0 = h->launches;
            int rc = dense_step(h, d, dy, w, -1, true, lml_steps, m_f, dsm, P_f, dsP, fz == 1);
            per_replay[fz] = h->launches - l0;    // kernels of one replay; nothing ran during the capture itself
            h->launches = l0;
            cudaError_t ce = cudaStreamEndCapture(st, &graph[fz]);
            if (rc != TGP_OK) return rc;
            TGP_CUDA(h, ce);
            TGP_CUDA(h, cudaGraphInstantiate(&exec[fz], graph[fz], 0));
        }
        constexpr int64_t kPoll = 4;       // a full step costs ~0.5 ms at D = 768: looking every 4 steps costs one ~20 us sync per 2 ms
        long long* pss = (long long*)(h->pinned + 16);
        *pss = -1;
        bool frozen = false;
        for (int64_t t = 0; t < T; ++t) {
            TGP_CUDA(h, cudaGraphLaunch(exec[frozen ? 1 : 0], st));
            h->launches += per_replay[frozen ? 1 : 0];
            if (w.conv && !frozen && (t + 1) % kPoll == 0) {
                TGP_CUDA(h, cudaMemcpyAsync(pss, w.ss_at, sizeof(long long), cudaMemcpyDeviceToHost, st));
                TGP_CUDA(h, cudaStreamSynchronize(st));
                h->d2h += 8;
                frozen = *pss >= 0;
            }
        }
        TGP_CUDA(h, cudaStreamSynchronize(st));
        for (int fz = 0; fz < 2; ++fz) {
            if (exec[fz]) cudaGraphExecDestroy(exec[fz]);
            if (graph[fz]) cudaGraphDestroy(graph[fz]);
        }
    } else {
        // time-varying parameters: pointers depend on t, issue the step's launches directly. The custom kernels index
        // through the device counter, the products through host-computed pointers.
        for (int64_t n = 0; n < T; ++n) {
            const long long t = rev ? T - 1 - n : n;
            TGP_TRY(dense_step(h, d, dy, w, t, false, lml_steps, m_f, dsm, P_f, dsP));
        }
    }
    if (w.conv && getenv("TGP_DEBUG")) {
        double cv[4];
        long long ss;
        cudaMemcpy(cv, w.conv, sizeof cv, cudaMemcpyDeviceToHost);
        cudaMemcpy(&ss, w.ss_at, sizeof ss, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[tgp] dense f64: last max|dP| = %.3e, max|P| = %.3e, ratio %.3e, tol %.1e, converged at step %lld\n", cv[2], cv[3],
                cv[3] > 0 ? cv[2] / cv[3] : 0.0, h->ss_tol, ss);
    }
    // results
    unsigned long long* perr = (unsigned long long*)h->pinned;
    TGP_CUDA(h, cudaMemcpyAsync(perr, w.err, 8, cudaMemcpyDeviceToHost, st));
    h->d2h += 8;
    TGP_TRY(deliver_scalar(h, w.lml, lml_out));
    TGP_TRY(flush_outputs(h));
    TGP_CUDA(h, cudaStreamSynchronize(st));
    if (*perr != ~0ull) return fail(h, TGP_ENOTPD, "innovation covariance not positive definite at time index %lld (0-based)", (long long)*perr);
    return TGP_OK;
}

}  // namespace tgp

namespace tgp {
void dense_release(tgp_ctx* h) {
    (void)h;   // nothing to release: the FP64 path owns no library handle any more
}
}  // namespace tgp

#include "tgp_dense_tc.cuh"

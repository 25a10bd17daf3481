// tgp_steady_smooth.cuh — posterior marginals of TIME-INVARIANT scalar-observation models (BASELINE config 3; the chain
// posterior(model, y) -> replace_observation_noise_cov -> marginals of posterior_lti_sde.jl:27-36, i.e. step_posterior /
// invert_dynamics lgssm.jl:215-240 followed by step_marginals lgssm.jl:111-115) without a D x D matrix product per step.
//
// In a time-invariant model both covariance recursions are data-free and converge: the filtering covariance P_f (forward)
// and, once P_f is constant, the smoothing covariance P_s (backward, P_s <- G P_s G' + Sigma with constant G, Sigma from
// invert_dynamics). Away from the two ends of the series only the MEANS move, by constant-coefficient affine maps:
//     forward   m_f[t]   = Abar m_f[t-1] + K y_t + c          v_t = y_t - w'm_f[t-1] - hh,  lml_t = -(log 2pi + log S + v_t^2/S)/2
//     backward  m_s[t-1] = G m_s[t] + E m_f[t-1] + e0          E = I - G A,  e0 = -G a        (lgssm.jl:231-240 with frozen P)
//     output    mean_t = H m_s[t] + h,   var_t = H P_s^inf H' + R_new[t]
// so only the two ENDS of the series need matrix work, and there it is data-free except for the means:
//     head  [0, N0)        k_sm_head_fwd: ONE CTA runs the filter sequentially (block-cooperative D x D products in shared memory)
//                          until |P_f[t] - P_f[t-1]| <= tol |P_f| (N0 ~ 10^3), keeping P_f[t], P_p[t], m_f[t];
//                          k_sm_head_dyn: (G_t, Sigma_t) of those steps, one thread per step (invert_dynamics, lgssm.jl:231-240);
//                          k_sm_head_bwd: ONE CTA runs the RTS recursion back from (m_s[N0-1], P_s^inf) and emits the marginals;
//     tail  (T - n1, T)    k_sm_tail_var (on a side stream, overlapping the scans): ONE CTA iterates P_s <- G P_s G' + Sigma back from P_s[T-1] = P_f^inf until it stops moving,
//                          emitting var_t (the tail MEANS follow the constant-coefficient recursion: G is constant there);
//     means [N0, T)        two constant-coefficient scans over vectors (forward filter means, backward smoother means).
// Convergence is CHECKED on the device against TGP_OPT_SS_TOL (the forward test inside k_sm_head_fwd, the backward one inside
// k_sm_tail_var, cross-checked against P_s^inf obtained independently by doubling); if either fails the caller redoes the call
// with the general scan kernels.
//
// The constant-coefficient scan x_i = Phi x_{i-1} + u_i is three levels of chunks of kCsL items (Phi^(L^k) precomputed): fold every
// chunk from zero (reduce), recurse on the chunk aggregates, a short sequential pass on top, then re-run every chunk from its true
// incoming state (apply). One thread per chunk, Phi in shared memory (broadcast reads), the state vector in registers.
#pragma once
#include "tgp_ctx.cuh"
#include "tgp_scan_small.cuh"

namespace tgp {

constexpr int kCsL = 64;          // items per chunk at level 0
constexpr int kCsLu = 16;         // items per chunk at the upper levels (few items there: shorter chunks = more threads)
constexpr int kCsThreads = 128;

template <int D>
struct SmConst {
    double PhiF[4][D * D];   // row-major Abar^e_k, e = 1, L, L Lu, L Lu^2: the transfer of one item at level k
    double PhiB[4][D * D];   // row-major G^e_k
    double K[D], c[D], w[D];
    double E[D * D], e0[D];  // row-major E
    double H[D];
    double hh, h0, S, invS, logS, vss;
    double Psinf[Sym<D>::N];
    double xfirst[D];        // smoothed mean at the first backward item (from the tail)
    double mstart[D];        // filtered mean entering the forward scan (m_f[Nh - 1])
    double sback[D + Sym<D>::N];   // packed (m_s[Nh - 1], P_s^inf): initial state of the head's backward pass
    double Sig[D * D];       // row-major Sigma of the steady reverse-time dynamics
    double Pfinf[D * D];     // filtering covariance at its fixed point (full)
    double lml_head;         // log marginal likelihood of the head steps
    long long N0, n1;        // head length; tail length (steps until P_s stopped moving)
    int conv_f, conv_b;
    double err_f, err_b;
};

// ---- block-cooperative D x D helpers (row-major matrices in shared memory, one element per thread) -------------------------------
template <int D> __device__ __forceinline__ void bk_mm(const double* A, const double* B, double* C) {        // C = A B
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(A[i * D + k], B[k * D + j], s);
        C[e] = s;
    }
    __syncthreads();
}
template <int D> __device__ __forceinline__ void bk_mmT_add(const double* A, const double* B, const double* Q, double* C) {   // C = A B' + Q
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        double s = Q ? Q[e] : 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(A[i * D + k], B[j * D + k], s);
        C[e] = s;
    }
    __syncthreads();
}
// block max of a per-thread value (blockDim = 128)
__device__ __forceinline__ double bk_max(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
    __syncthreads();
    return v;
}

// ---- head, forward: sequential filter in one CTA until the covariance stops moving ---------------------------------------------------
// Keeps the filtering distributions of t < N0 in the smoother's SoA layout (ws[k * Nmax + t]: mean, then the packed upper triangle)
// and MFh[t] = m_f[t]. Same arithmetic as predict (LGC:46-52) and posterior_and_lml(::ScalarOutputLGC) (LGC:247-257).
template <int D>
__global__ void __launch_bounds__(128) k_sm_head_fwd(const DevModel dm, const double* __restrict__ m0, const double* __restrict__ P0, long long Nmax,
                                                     double tol, double* __restrict__ ws, double* __restrict__ MFh,
                                                     SmConst<D>* __restrict__ cst, unsigned long long* __restrict__ err_step) {
    __shared__ double sA[D * D], sQ[D * D], sP[D * D], sT[D * D], sPp[D * D];
    __shared__ double sh[D], sa[D], sm[D], smp[D], sV[D], red[4];
    __shared__ double sSq[2048];                                    // (S_t, v_t^2 / S_t) of up to 1024 steps awaiting their log
    const int tid = threadIdx.x;
    for (int e = tid; e < D * D; e += 128) {
        const int i = e / D, j = e % D;
        sA[e] = dm.A[i + D * j]; sQ[e] = dm.Q[i + D * j]; sP[e] = P0[i + D * j];
    }
    if (tid < D) { sh[tid] = dm.H[tid]; sa[tid] = dm.a[tid]; sm[tid] = m0[tid]; }
    __syncthreads();
    const double h0 = *dm.h, R = *dm.R;
    double lml = 0.0;
    long long N0 = Nmax;
    int conv = 0;
    double y_next = __ldg(dm.y);                                     // the observation is fetched one step ahead: its latency is off the chain
    for (long long t = 0; t < Nmax; ++t) {
        const double y_t = y_next;
        if (t + 1 < Nmax) y_next = __ldg(dm.y + t + 1);
        bk_mm<D>(sA, sP, sT);
        bk_mmT_add<D>(sT, sA, sQ, sPp);                              // P_p = A P A' + Q
        if (tid < D) {
            double v = 0.0, mp = sa[tid];
#pragma unroll
            for (int k = 0; k < D; ++k) { v = fma(sPp[tid * D + k], sh[k], v); mp = fma(sA[tid * D + k], sm[k], mp); }
            sV[tid] = v; smp[tid] = mp;
        }
        __syncthreads();
        double S = R, pred = h0;
#pragma unroll
        for (int k = 0; k < D; ++k) { S = fma(sh[k], sV[k], S); pred = fma(sh[k], smp[k], pred); }
        if (!(S > 1e-300) || !(S < 1e300)) { if (tid == 0) atomicMin(err_step, (unsigned long long)t); S = 1.0; }
        const double invS = 1.0 / S, v = y_t - pred;
        if (tid == 0) { sSq[(t & 1023) * 2] = S; sSq[(t & 1023) * 2 + 1] = v * v * invS; }
        if ((t & 1023) == 1023) {                                   // drain: the logs of 1024 steps, in parallel
            __syncthreads();
            for (int q = tid; q < 1024; q += 128) lml -= 0.5 * (kLog2Pi + log(sSq[2 * q]) + sSq[2 * q + 1]);
            __syncthreads();
        }
        double dmax = 0.0, amax = 0.0;
        for (int e = tid; e < D * D; e += 128) {
            const int i = e / D, j = e % D;
            const double pf = fma(-sV[i] * invS, sV[j], sPp[e]);
            if (ws && i <= j) ws[(size_t)(D + Sym<D>::idx(i, j)) * Nmax + t] = pf;
            dmax = fmax(dmax, fabs(pf - sP[e]));
            amax = fmax(amax, fabs(pf));
            sP[e] = pf;
        }
        if (tid < D) {
            const double mf = fma(sV[tid] * invS, v, smp[tid]);
            sm[tid] = mf;
            if (MFh) MFh[t * D + tid] = mf;
            if (ws) ws[(size_t)tid * Nmax + t] = mf;
        }
        __syncthreads();
        if ((t & 15) == 15 && t >= 31) {                             // test every 16 steps: a block reduction costs two barriers
            const double d = bk_max(dmax, red), a = bk_max(amax, red);
            if (d <= tol * a) { N0 = t + 1; conv = 1; if (tid == 0) cst->err_f = a > 0.0 ? d / a : 0.0; break; }
        }
    }
    __syncthreads();
    {   // the steps since the last drain, then the block total
        const long long done = N0 & ~1023LL;
        for (long long q = done + tid; q < N0; q += 128) lml -= 0.5 * (kLog2Pi + log(sSq[2 * (q & 1023)]) + sSq[2 * (q & 1023) + 1]);
        __shared__ double lred[128];
        lred[tid] = lml;
        __syncthreads();
        for (int off = 64; off > 0; off >>= 1) { if (tid < off) lred[tid] += lred[tid + off]; __syncthreads(); }
        if (tid == 0) cst->lml_head = lred[0];
    }
    for (int e = tid; e < D * D; e += 128) cst->Pfinf[e] = sP[e];
    if (tid < D) cst->mstart[tid] = sm[tid];
    if (tid == 0) { cst->N0 = N0; cst->conv_f = conv; }
}


// y = M x (M row-major in shared memory, x in registers)
template <int D> __device__ __forceinline__ void cs_matvec(const double* __restrict__ M, const double (&x)[D], double (&y)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(M[i * D + j], x[j], s);
        y[i] = s;
    }
}

// ---- set-up: one CTA ----------------------------------------------------------------------------------------------------
// Row-major product C = A B of D x D matrices in shared memory, one element per thread.
template <int D> __device__ __forceinline__ void cs_matmul(const double* A, const double* B, double* C) {
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        double s = 0.0;
        for (int k = 0; k < D; ++k) s = fma(A[i * D + k], B[k * D + j], s);
        C[e] = s;
    }
    __syncthreads();
}

template <int D>
__global__ void __launch_bounds__(128) k_sm_setup(const DevModel dm, SmConst<D>* __restrict__ cst) {
    __shared__ double sA[D * D], sB[D * D], sC[D * D], sG[D * D], sS[D * D], sT[D * D];
    const int tid = threadIdx.x;
    if (tid == 0) {
        Sym<D> Pf;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i <= j; ++i) Pf(i, j) = 0.5 * (cst->Pfinf[i * D + j] + cst->Pfinf[j * D + i]);
        const Mat<D> A = ldg_mat<D>(dm.A);
        const Vec<D> a = ldg_vec<D>(dm.a);
        const Sym<D> Q = ldg_sym_full<D>(dm.Q);
        const Vec<D> H = ldg_vec<D>(dm.H);
        const double h0 = *dm.h, R = *dm.R;
        Vec<D> mz = vzero<D>(), mp = vzero<D>();
        Sym<D> Pp = Pf;
        predict(mp, Pp, A, vzero<D>(), Q);                        // P_p = A P_f A' + Q
        const Vec<D> V = symvec(Pp, H);
        const double S = dot(V, H) + R;
        for (int i = 0; i < D; ++i) { cst->K[i] = V[i] / S; cst->H[i] = H[i]; }
        cst->S = S; cst->invS = 1.0 / S; cst->logS = log(S); cst->h0 = h0;
        cst->hh = dot(H, a) + h0;
        const Vec<D> w = matTvec(A, H);
        for (int i = 0; i < D; ++i) cst->w[i] = w[i];
        // Abar = (I - K H') A (row-major), c = (I - K H') a - K h0
        for (int i = 0; i < D; ++i) {
            for (int j = 0; j < D; ++j) sA[i * D + j] = A(i, j) - cst->K[i] * w[j];
            cst->c[i] = a[i] - cst->K[i] * cst->hh;
        }
        Aff<D> inv;
        invert_dynamics(mz, Pf, mp, Pp, A, inv);                  // G = inv.A, Sigma = inv.C (lgssm.jl:231-240)
        for (int i = 0; i < D; ++i) {
            double ge = 0.0;
            for (int j = 0; j < D; ++j) {
                sG[i * D + j] = inv.A(i, j);
                double ga = 0.0;
                for (int k = 0; k < D; ++k) ga = fma(inv.A(i, k), A(k, j), ga);
                cst->E[i * D + j] = (i == j ? 1.0 : 0.0) - ga;
                ge = fma(inv.A(i, j), a[j], ge);
                sS[i * D + j] = inv.C(i, j);
                cst->Sig[i * D + j] = inv.C(i, j);
            }
            cst->e0[i] = -ge;
        }
    }
    __syncthreads();
    // powers: level-1 items span L = 2^6 steps, every further level Lu = 2^4 items of the previous one
    for (int which = 0; which < 2; ++which) {
        double* base = which == 0 ? sA : sG;
        for (int e = tid; e < D * D; e += blockDim.x) { sB[e] = base[e]; (which == 0 ? cst->PhiF[0] : cst->PhiB[0])[e] = base[e]; }
        __syncthreads();
        for (int lev = 1; lev < 4; ++lev) {
            for (int sq = 0; sq < (lev == 1 ? 6 : 4); ++sq) {
                cs_matmul<D>(sB, sB, sC);
                for (int e = tid; e < D * D; e += blockDim.x) sB[e] = sC[e];
                __syncthreads();
            }
            for (int e = tid; e < D * D; e += blockDim.x) (which == 0 ? cst->PhiF[lev] : cst->PhiB[lev])[e] = sB[e];
            __syncthreads();
        }
    }
    // P_s^inf = sum_k G^k Sigma G^k' by doubling: (S, G) <- (S + G S G', G G)
    for (int e = tid; e < D * D; e += blockDim.x) sB[e] = sG[e];
    __syncthreads();
    for (int it = 0; it < 48; ++it) {
        cs_matmul<D>(sB, sS, sC);                                 // C = G S
        for (int e = tid; e < D * D; e += blockDim.x) {           // T = C G' ; S += T
            const int i = e / D, j = e % D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s = fma(sC[i * D + k], sB[j * D + k], s);
            sT[e] = s;
        }
        __syncthreads();
        for (int e = tid; e < D * D; e += blockDim.x) sS[e] += sT[e];
        cs_matmul<D>(sB, sB, sC);
        for (int e = tid; e < D * D; e += blockDim.x) sB[e] = sC[e];
        __syncthreads();
    }
    if (tid == 0) {
        double v = 0.0;
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) v = fma(cst->H[i] * cst->H[j], 0.5 * (sS[i * D + j] + sS[j * D + i]), v);
        cst->vss = v;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i <= j; ++i) cst->Psinf[Sym<D>::idx(i, j)] = 0.5 * (sS[i * D + j] + sS[j * D + i]);
    }
}

// ---- tail: smoothing covariance back from P_s[T-1] = P_f^inf until it stops moving; emits var_t of those steps ------------------------
template <int D>
__global__ void __launch_bounds__(128) k_sm_tail_var(SmConst<D>* __restrict__ cst, long long nmax, double tol, double* __restrict__ tvar) {
    __shared__ double sG[D * D], sSig[D * D], sP[D * D], sT[D * D], sN[D * D], sh[D], sV[D], red[4];
    const int tid = threadIdx.x;
    for (int e = tid; e < D * D; e += 128) { sG[e] = cst->PhiB[0][e]; sSig[e] = cst->Sig[e]; sP[e] = cst->Pfinf[e]; }
    if (tid < D) sh[tid] = cst->H[tid];
    __syncthreads();
    long long n1 = nmax;
    int conv = 0;
    for (long long n = 0; n < nmax; ++n) {
        if (tid >= 128 - D) {                                        // the last warp is idle in the products: V = P h there
            const int i = tid - (128 - D);
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) v = fma(sP[i * D + j], sh[j], v);
            sV[i] = v;
        }
        bk_mm<D>(sG, sP, sT);
        if (tid == 127) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) v = fma(sh[i], sV[i], v);
            tvar[n] = v;                                              // H P_s[T-1-n] H' (the caller adds R_new)
        }
        bk_mmT_add<D>(sT, sG, sSig, sN);                             // P_s <- G P_s G' + Sigma
        double dmax = 0.0, amax = 0.0;
        for (int e = tid; e < D * D; e += 128) {
            dmax = fmax(dmax, fabs(sN[e] - sP[e]));
            amax = fmax(amax, fabs(sN[e]));
            sP[e] = sN[e];
        }
        __syncthreads();
        if ((n & 15) == 15) {
            const double d = bk_max(dmax, red), a = bk_max(amax, red);
            if (d <= tol * a) { n1 = n + 1; conv = 1; break; }
        }
    }
    // cross-check against the fixed point obtained by doubling in k_sm_setup (different summation order: allow roundoff)
    double dmax = 0.0, amax = 0.0;
    for (int e = tid; e < D * D; e += 128) {
        const int i = e / D, j = e % D;
        const double ps = cst->Psinf[Sym<D>::idx(i, j)];
        dmax = fmax(dmax, fabs(sP[e] - ps));
        amax = fmax(amax, fabs(ps));
    }
    const double d = bk_max(dmax, red), a = bk_max(amax, red);
    if (tid == 0) { cst->n1 = n1; cst->err_b = a > 0.0 ? d / a : 0.0; cst->conv_b = conv && d <= 1000.0 * tol * a; }
}

// var[T-1-n] = tvar[n] + R_new for the n1 tail steps (runs after both the tail recursion and the backward scan)
template <int D>
__global__ void __launch_bounds__(256) k_sm_tail_copy(const SmConst<D>* __restrict__ cst, const double* __restrict__ tvar, long long T,
                                                      const double* __restrict__ Rn, long long sR, double* __restrict__ var) {
    const long long n1 = cst->n1;
    for (long long n = (long long)blockIdx.x * 256 + threadIdx.x; n < n1; n += (long long)gridDim.x * 256)
        var[T - 1 - n] = tvar[n] + Rn[(T - 1 - n) * sR];
}

// ---- head, backward ------------------------------------------------------------------------------------------------------------------
// Reverse-time dynamics (G_t, g_t, Sigma_t) mapping x_t -> x_{t-1}, t = 1 .. N0-1, one thread per t: the same provider the general
// backward kernels use (SmootherProvider: predict + invert_dynamics, lgssm.jl:215-240, jitter 1e-10) on the head's stored states.
template <int D>
__global__ void __launch_bounds__(kBlock) k_sm_head_dyn(const SmootherProvider<D> prov, long long N0, double* __restrict__ GG, double* __restrict__ gg,
                                                        double* __restrict__ SS) {
    const long long t = (long long)blockIdx.x * kBlock + threadIdx.x + 1;
    if (t >= N0) return;
    Aff<D> e;
    prov.get(prov.dm.T - 1 - t, e);
    for (int i = 0; i < D; ++i) {
        gg[t * D + i] = e.b[i];
        for (int j = 0; j < D; ++j) { GG[t * D * D + i * D + j] = e.A(i, j); SS[t * D * D + i * D + j] = e.C(i, j); }
    }
}
// RTS recursion back from (m_s[N0-1] = xlast, P_s^inf): m_s[t-1] = G_t m_s[t] + g_t, P_s[t-1] = G_t P_s[t] G_t' + Sigma_t;
// emits mean / var of t = N0-2 .. 0 (t = N0-1 was emitted by the backward scan). Also sets the combined convergence word.
template <int D>
__global__ void __launch_bounds__(128) k_sm_head_bwd(const DevModel dm, const SmConst<D>* __restrict__ cst, const double* __restrict__ xlast,
                                                     const double* __restrict__ GG, const double* __restrict__ gg, const double* __restrict__ SS,
                                                     const double* __restrict__ Rn, long long sR, double* __restrict__ mean, double* __restrict__ var,
                                                     int* __restrict__ flag) {
    __shared__ double sG[D * D], sSig[D * D], sP[D * D], sT[D * D], sN[D * D];
    __shared__ double sh[D], sm[D], sd[D], sg[D], sV[D];
    const int tid = threadIdx.x;
    const long long N0 = cst->N0;
    for (int e = tid; e < D * D; e += 128) {
        const int i = e / D, j = e % D;
        sP[e] = cst->Psinf[Sym<D>::idx(i, j)];
    }
    if (tid < D) { sh[tid] = cst->H[tid]; sm[tid] = xlast[tid]; }
    if (tid == 0) *flag = cst->conv_f && cst->conv_b;
    __syncthreads();
    const double h0 = cst->h0;
    // (G_t, g_t, Sigma_t) travel one step ahead in registers (D * D <= 128: one entry per thread), so the global-memory latency of
    // step t - 1 is covered by the products of step t
    static_assert(D * D <= 128, "one entry of G and Sigma per thread");
    double pG = 0.0, pS = 0.0, pg = 0.0, pR = 0.0;
    if (N0 - 1 >= 1) {
        if (tid < D * D) { pG = GG[(N0 - 1) * D * D + tid]; pS = SS[(N0 - 1) * D * D + tid]; }
        if (tid < D) pg = gg[(N0 - 1) * D + tid];
        if (tid == 127) pR = Rn[(N0 - 2) * sR];
    }
    for (long long t = N0 - 1; t >= 1; --t) {
        if (tid < D * D) { sG[tid] = pG; sSig[tid] = pS; }
        if (tid < D) { sg[tid] = pg; sd[tid] = sm[tid]; }
        const double R_t = pR;
        if (t - 1 >= 1) {
            if (tid < D * D) { pG = GG[(t - 1) * D * D + tid]; pS = SS[(t - 1) * D * D + tid]; }
            if (tid < D) pg = gg[(t - 1) * D + tid];
            if (tid == 127) pR = Rn[(t - 2) * sR];
        }
        __syncthreads();
        bk_mm<D>(sG, sP, sT);
        bk_mmT_add<D>(sT, sG, sSig, sN);
        if (tid < D) {
            double ms = sg[tid];
#pragma unroll
            for (int k = 0; k < D; ++k) ms = fma(sG[tid * D + k], sd[k], ms);
            sm[tid] = ms;
        }
        for (int e = tid; e < D * D; e += 128) sP[e] = sN[e];
        if (tid >= 128 - D) {                                        // V = P_s h from the fresh sN (the idle last warp)
            const int i = tid - (128 - D);
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) v = fma(sN[i * D + j], sh[j], v);
            sV[i] = v;
        }
        __syncthreads();
        if (tid == 127) {
            double mu = h0, v = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) { mu = fma(sh[i], sm[i], mu); v = fma(sh[i], sV[i], v); }
            mean[t - 1] = mu;
            var[t - 1] = v + R_t;
        }
    }
}

// ---- item sources / sinks of level 0 -------------------------------------------------------------------------------------------
template <int D>
struct FwdItems {                 // item i <-> time t0 + i;  u = K y + c
    const double* y;
    long long t0;
    double* MF;                   // MF[(i + 1) D ..] = m_f[t0 + i];  MF[0 ..] = m_f[t0 - 1]
    double* partials;             // per CTA: sum of v^2 / S
    __device__ __forceinline__ void input(const SmConst<D>* c, long long i, double (&u)[D]) const {
        const double yy = __ldg(y + t0 + i);
#pragma unroll
        for (int k = 0; k < D; ++k) u[k] = fma(c->K[k], yy, c->c[k]);
    }
};
template <int D>
struct BwdItems {                 // item i <-> time t_hi - i;  u = E m_f[t] + e0  (item 0: the smoothed mean handed over by the tail)
    const double* MF;             // MF[j D ..] = m_f[tbase + j]
    long long j_hi;               // MF index of item 0
    long long t_hi;
    const double* Rn; long long sR;
    double* mean; double* var;
};

// ---- kernels ---------------------------------------------------------------------------------------------------------------------
// fold chunk `ch` of level-0 items from the zero state
template <int D, bool FWD>
__global__ void __launch_bounds__(kCsThreads) k_cs_reduce0(const SmConst<D>* __restrict__ cst, FwdItems<D> fi, BwdItems<D> bi, long long n,
                                                           double* __restrict__ z1) {
    __shared__ double Phi[D * D];
    __shared__ double Em[D * D];
    for (int e = threadIdx.x; e < D * D; e += kCsThreads) { Phi[e] = FWD ? cst->PhiF[0][e] : cst->PhiB[0][e]; Em[e] = cst->E[e]; }
    __syncthreads();
    const long long ch = (long long)blockIdx.x * kCsThreads + threadIdx.x;
    const long long s = ch * kCsL, e = min(s + (long long)kCsL, n);
    if (s >= n) return;
    double x[D], u[D], t[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = 0.0;
    for (long long i = s; i < e; ++i) {
        if (FWD) fi.input(cst, i, u);
        else {
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < D; ++k) u[k] = cst->xfirst[k];
            } else {
                double mf[D];
                const double* p = bi.MF + (bi.j_hi - i) * D;
#pragma unroll
                for (int k = 0; k < D; ++k) mf[k] = __ldg(p + k);
                cs_matvec<D>(Em, mf, u);
#pragma unroll
                for (int k = 0; k < D; ++k) u[k] += cst->e0[k];
            }
        }
        cs_matvec<D>(Phi, x, t);
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = t[k] + u[k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) z1[ch * D + k] = x[k];
}

// fold chunk `ch` of vector items z (level lev >= 1) from the zero state
template <int D>
__global__ void __launch_bounds__(kCsThreads) k_cs_reduce(const double* __restrict__ PhiG, const double* __restrict__ z, long long n,
                                                          double* __restrict__ zn) {
    __shared__ double Phi[D * D];
    for (int e = threadIdx.x; e < D * D; e += kCsThreads) Phi[e] = PhiG[e];
    __syncthreads();
    const long long ch = (long long)blockIdx.x * kCsThreads + threadIdx.x;
    const long long s = ch * kCsLu, e = min(s + (long long)kCsLu, n);
    if (s >= n) return;
    double x[D], t[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = 0.0;
    for (long long i = s; i < e; ++i) {
        cs_matvec<D>(Phi, x, t);
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = t[k] + __ldg(z + i * D + k);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) zn[ch * D + k] = x[k];
}

// top: sequential over the n items of the highest level; X[i] = state BEFORE item i. xinit nullable (zero).
template <int D>
__global__ void k_cs_top(const double* __restrict__ PhiG, const double* __restrict__ z, long long n, const double* __restrict__ xinit,
                         double* __restrict__ X) {
    if (threadIdx.x || blockIdx.x) return;
    double x[D], t[D];
    for (int k = 0; k < D; ++k) x[k] = xinit ? xinit[k] : 0.0;
    for (long long i = 0; i < n; ++i) {
        for (int k = 0; k < D; ++k) X[i * D + k] = x[k];
        for (int r = 0; r < D; ++r) {
            double s = 0.0;
            for (int j = 0; j < D; ++j) s = fma(PhiG[r * D + j], x[j], s);
            t[r] = s + z[i * D + r];
        }
        for (int k = 0; k < D; ++k) x[k] = t[k];
    }
}

// apply (levels >= 1): chunk `ch` of vector items starts from Xin[ch]; writes X[i] = state BEFORE item i
template <int D>
__global__ void __launch_bounds__(kCsThreads) k_cs_apply(const double* __restrict__ PhiG, const double* __restrict__ z, long long n,
                                                         const double* __restrict__ Xin, double* __restrict__ X) {
    __shared__ double Phi[D * D];
    for (int e = threadIdx.x; e < D * D; e += kCsThreads) Phi[e] = PhiG[e];
    __syncthreads();
    const long long ch = (long long)blockIdx.x * kCsThreads + threadIdx.x;
    const long long s = ch * kCsLu, e = min(s + (long long)kCsLu, n);
    if (s >= n) return;
    double x[D], t[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = Xin[ch * D + k];
    for (long long i = s; i < e; ++i) {
#pragma unroll
        for (int k = 0; k < D; ++k) X[i * D + k] = x[k];
        cs_matvec<D>(Phi, x, t);
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = t[k] + __ldg(z + i * D + k);
    }
}

// apply (level 0): re-run every chunk from its true incoming state and emit.
template <int D, bool FWD>
__global__ void __launch_bounds__(kCsThreads) k_cs_apply0(const SmConst<D>* __restrict__ cst, FwdItems<D> fi, BwdItems<D> bi, long long n,
                                                          const double* __restrict__ Xin, double* __restrict__ xlast) {
    __shared__ double Phi[D * D];
    __shared__ double Em[D * D];
    __shared__ double red[kCsThreads];
    for (int e = threadIdx.x; e < D * D; e += kCsThreads) { Phi[e] = FWD ? cst->PhiF[0][e] : cst->PhiB[0][e]; Em[e] = cst->E[e]; }
    __syncthreads();
    const long long ch = (long long)blockIdx.x * kCsThreads + threadIdx.x;
    const long long s = ch * kCsL, e = min(s + (long long)kCsL, n);
    double quad = 0.0;
    if (s < n) {
        double x[D], u[D], t[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = Xin[ch * D + k];
        for (long long i = s; i < e; ++i) {
            if (FWD) {
                const double yy = __ldg(fi.y + fi.t0 + i);
                double pred = cst->hh;
#pragma unroll
                for (int k = 0; k < D; ++k) pred = fma(cst->w[k], x[k], pred);
                const double v = yy - pred;
                quad = fma(v, v, quad);
#pragma unroll
                for (int k = 0; k < D; ++k) u[k] = fma(cst->K[k], yy, cst->c[k]);
            } else {
                if (i == 0) {
#pragma unroll
                    for (int k = 0; k < D; ++k) u[k] = cst->xfirst[k];
                } else {
                    double mf[D];
                    const double* p = bi.MF + (bi.j_hi - i) * D;
#pragma unroll
                    for (int k = 0; k < D; ++k) mf[k] = __ldg(p + k);
                    cs_matvec<D>(Em, mf, u);
#pragma unroll
                    for (int k = 0; k < D; ++k) u[k] += cst->e0[k];
                }
            }
            cs_matvec<D>(Phi, x, t);
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = t[k] + u[k];
            if (FWD) {
                if (fi.MF) {                                   // nullable: the log-likelihood alone needs no stored means
                    double* o = fi.MF + (i + 1) * D;
#pragma unroll
                    for (int k = 0; k < D; ++k) o[k] = x[k];
                }
            } else {
                const long long tt = bi.t_hi - i;
                double mu = cst->h0;
#pragma unroll
                for (int k = 0; k < D; ++k) mu = fma(cst->H[k], x[k], mu);
                bi.mean[tt] = mu;
                bi.var[tt] = cst->vss + __ldg(bi.Rn + tt * bi.sR);
            }
        }
        if (e == n && xlast) {
#pragma unroll
            for (int k = 0; k < D; ++k) xlast[k] = x[k];
        }
    }
    if (FWD) {
        red[threadIdx.x] = quad;
        __syncthreads();
        for (int off = kCsThreads / 2; off > 0; off >>= 1) {
            if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) fi.partials[blockIdx.x] = red[0];
    }
}

// lml of the steady steps from the per-CTA sums of v^2, added to the head's lml
template <int D>
__global__ void __launch_bounds__(256) k_sm_lml(const SmConst<D>* __restrict__ cst, const double* __restrict__ partials, long long np, long long n,
                                                double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (long long i = threadIdx.x; i < np; i += 256) s += partials[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = cst->lml_head - 0.5 * ((double)n * (kLog2Pi + cst->logS) + sm[0] * cst->invS);
}

template <int D>
__global__ void k_sm_set_mstart(const SmConst<D>* __restrict__ cst, double* __restrict__ MF) {
    if (threadIdx.x || blockIdx.x) return;
    for (int k = 0; k < D; ++k) MF[k] = cst->mstart[k];
}
// the backward scan starts from the smoothed mean at T-1 = the filtered mean there (last entry of MF)
template <int D>
__global__ void k_sm_set_xfirst(SmConst<D>* __restrict__ cst, const double* __restrict__ mf_last) {
    if (threadIdx.x || blockIdx.x) return;
    for (int k = 0; k < D; ++k) cst->xfirst[k] = mf_last[k];
}

// Constant-coefficient scan driver. Level sizes n1 = ceil(n / L), n2 = ceil(n1 / Lu), n3 = ceil(n2 / Lu); three chunk levels and a
// sequential top over n3 items.
template <int D, bool FWD>
int cs_scan(tgp_ctx* h, const SmConst<D>* cst, const FwdItems<D>& fi, const BwdItems<D>& bi, long long n, const double* xinit, double* xlast) {
    cudaStream_t st = h->stream;
    const long long n1 = (n + kCsL - 1) / kCsL, n2 = (n1 + kCsLu - 1) / kCsLu, n3 = (n2 + kCsLu - 1) / kCsLu;
    double *z1, *z2, *z3, *X1, *X2, *X3;
    TGP_TRY(dalloc(h, (size_t)n1 * D, &z1));
    TGP_TRY(dalloc(h, (size_t)n2 * D, &z2));
    TGP_TRY(dalloc(h, (size_t)n3 * D, &z3));
    TGP_TRY(dalloc(h, (size_t)n1 * D, &X1));
    TGP_TRY(dalloc(h, (size_t)n2 * D, &X2));
    TGP_TRY(dalloc(h, (size_t)n3 * D, &X3));
    const double* Phi1 = FWD ? cst->PhiF[1] : cst->PhiB[1];
    const double* Phi2 = FWD ? cst->PhiF[2] : cst->PhiB[2];
    const double* Phi3 = FWD ? cst->PhiF[3] : cst->PhiB[3];
    auto blocks = [](long long chunks) { return (unsigned)((chunks + kCsThreads - 1) / kCsThreads); };
    TGP_K(h, FWD ? "k_cs_reduce0(fwd)" : "k_cs_reduce0(bwd)");
    k_cs_reduce0<D, FWD><<<blocks(n1), kCsThreads, 0, st>>>(cst, fi, bi, n, z1);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_cs_reduce");
    k_cs_reduce<D><<<blocks(n2), kCsThreads, 0, st>>>(Phi1, z1, n1, z2);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_cs_reduce");
    k_cs_reduce<D><<<blocks(n3), kCsThreads, 0, st>>>(Phi2, z2, n2, z3);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_cs_top");
    k_cs_top<D><<<1, 32, 0, st>>>(Phi3, z3, n3, xinit, X3);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_cs_apply");
    k_cs_apply<D><<<blocks(n3), kCsThreads, 0, st>>>(Phi2, z2, n2, X3, X2);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "k_cs_apply");
    k_cs_apply<D><<<blocks(n2), kCsThreads, 0, st>>>(Phi1, z1, n1, X2, X1);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, FWD ? "k_cs_apply0(fwd)" : "k_cs_apply0(bwd)");
    k_cs_apply0<D, FWD><<<blocks(n1), kCsThreads, 0, st>>>(cst, fi, bi, n, X1, xlast);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

}  // namespace tgp

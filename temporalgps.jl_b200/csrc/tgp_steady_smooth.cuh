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
// so the series is cut in three:
//     head  [0, Nh)        the general scan kernels (tgp_scan_small.cuh), forward and backward: time-varying gains;
//     tail  [T - Nt, T)    the general backward kernels on (m_f from the steady pass, P_f^inf): P_s still moving;
//     rest                 two constant-coefficient scans over vectors (this file): a D x D mat-VEC per step and pass.
// Convergence is CHECKED on the device (|P_f[Nh-1] - P_f[Nh-2]|, |P_s[T-Nt-1] - P_s^inf| against TGP_OPT_SS_TOL); if either test
// fails the caller redoes the call with the general path.
//
// The constant-coefficient scan x_i = Phi x_{i-1} + u_i is three levels of chunks of kCsL items (Phi^(L^k) precomputed): fold every
// chunk from zero (reduce), recurse on the chunk aggregates, a short sequential pass on top, then re-run every chunk from its true
// incoming state (apply). One thread per chunk, Phi in shared memory (broadcast reads), the state vector in registers.
#pragma once
#include "tgp_ctx.cuh"
#include "tgp_scan_small.cuh"

namespace tgp {

constexpr int kCsL = 64;
constexpr int kCsThreads = 128;

template <int D>
struct SmConst {
    double PhiF[4][D * D];   // row-major Abar^(L^k), k = 0..3
    double PhiB[4][D * D];   // row-major G^(L^k)
    double K[D], c[D], w[D];
    double E[D * D], e0[D];  // row-major E
    double H[D];
    double hh, h0, S, invS, logS, vss;
    double Psinf[Sym<D>::N];
    double xfirst[D];        // smoothed mean at the first backward item (from the tail)
    double mstart[D];        // filtered mean entering the forward scan (m_f[Nh - 1])
    double sback[D + Sym<D>::N];   // packed (m_s[Nh - 1], P_s^inf): initial state of the head's backward pass
    int conv_f, conv_b;
    double err_f, err_b;
};

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
__global__ void __launch_bounds__(128) k_sm_setup(const DevModel dm, const double* __restrict__ ws_head, long long Nh, double tol,
                                                  SmConst<D>* __restrict__ cst) {
    __shared__ double sA[D * D], sB[D * D], sC[D * D], sG[D * D], sS[D * D], sT[D * D];
    const int tid = threadIdx.x;
    if (tid == 0) {
        Vec<D> m1, m0;
        Sym<D> Pf, P0;
        load_state<D>(ws_head, Nh, Nh - 1, m1, Pf);
        load_state<D>(ws_head, Nh, Nh - 2, m0, P0);
        double md = 0.0, ma = 0.0;
        for (int k = 0; k < Sym<D>::N; ++k) { md = fmax(md, fabs(Pf.v[k] - P0.v[k])); ma = fmax(ma, fabs(Pf.v[k])); }
        cst->err_f = ma > 0.0 ? md / ma : 0.0;
        cst->conv_f = md <= tol * ma;
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
        for (int i = 0; i < D; ++i) { cst->K[i] = V[i] / S; cst->H[i] = H[i]; cst->mstart[i] = m1[i]; }
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
            }
            cst->e0[i] = -ge;
        }
    }
    __syncthreads();
    // powers Phi^(L^k): L = 64 = 2^6
    for (int which = 0; which < 2; ++which) {
        double* base = which == 0 ? sA : sG;
        for (int e = tid; e < D * D; e += blockDim.x) { sB[e] = base[e]; (which == 0 ? cst->PhiF[0] : cst->PhiB[0])[e] = base[e]; }
        __syncthreads();
        for (int lev = 1; lev < 4; ++lev) {
            for (int sq = 0; sq < 6; ++sq) {
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

// After the tail: test the backward covariance, hand over the first backward mean. xtail = packed (m_s, P_s) at T - Nt - 1.
template <int D>
__global__ void k_sm_after_tail(const double* __restrict__ xtail, double tol, SmConst<D>* __restrict__ cst) {
    if (threadIdx.x || blockIdx.x) return;
    double md = 0.0, ma = 0.0;
    for (int k = 0; k < Sym<D>::N; ++k) { md = fmax(md, fabs(xtail[D + k] - cst->Psinf[k])); ma = fmax(ma, fabs(cst->Psinf[k])); }
    cst->err_b = ma > 0.0 ? md / ma : 0.0;
    cst->conv_b = md <= 100.0 * tol * ma;      // P_s^inf comes from a different summation order than the recursion: allow roundoff
    for (int i = 0; i < D; ++i) cst->xfirst[i] = xtail[i];
}
// After the backward scan: packed initial state of the head's backward pass, and the combined convergence word.
template <int D>
__global__ void k_sm_finish(const double* __restrict__ xlast, SmConst<D>* __restrict__ cst, int* __restrict__ flag) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < D; ++i) cst->sback[i] = xlast[i];
    for (int k = 0; k < Sym<D>::N; ++k) cst->sback[D + k] = cst->Psinf[k];
    *flag = cst->conv_f && cst->conv_b;
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
    const long long s = ch * kCsL, e = min(s + (long long)kCsL, n);
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
    const long long s = ch * kCsL, e = min(s + (long long)kCsL, n);
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
                double* o = fi.MF + (i + 1) * D;
#pragma unroll
                for (int k = 0; k < D; ++k) o[k] = x[k];
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
                                                const double* __restrict__ lml_head, double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (long long i = threadIdx.x; i < np; i += 256) s += partials[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = *lml_head - 0.5 * ((double)n * (kLog2Pi + cst->logS) + sm[0] * cst->invS);
}

// filtering distributions of the tail in the smoother's SoA layout: means from MF, covariance P_f^inf (= the head's last one)
template <int D>
__global__ void __launch_bounds__(256) k_sm_fill_tail(const double* __restrict__ MF, long long j0, long long Nt, const double* __restrict__ xT_head,
                                                      double* __restrict__ ws_tail, double* __restrict__ x0_tail, double* __restrict__ xT_tail) {
    constexpr int SN = D + Sym<D>::N;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // local tail time
    if (i < Nt) {
        for (int k = 0; k < D; ++k) ws_tail[(size_t)k * Nt + i] = MF[(j0 + i) * D + k];
        for (int k = D; k < SN; ++k) ws_tail[(size_t)k * Nt + i] = xT_head[k];
    }
    if (i == 0) {
        for (int k = 0; k < D; ++k) { x0_tail[k] = MF[(j0 - 1) * D + k]; xT_tail[k] = MF[(j0 + Nt - 1) * D + k]; }
        for (int k = D; k < SN; ++k) { x0_tail[k] = xT_head[k]; xT_tail[k] = xT_head[k]; }
    }
}

template <int D>
__global__ void k_sm_set_mstart(const SmConst<D>* __restrict__ cst, double* __restrict__ MF) {
    if (threadIdx.x || blockIdx.x) return;
    for (int k = 0; k < D; ++k) MF[k] = cst->mstart[k];
}

// Constant-coefficient scan driver. Level sizes n0 = n, n_{k+1} = ceil(n_k / L); three chunk levels and a sequential top.
template <int D, bool FWD>
int cs_scan(tgp_ctx* h, const SmConst<D>* cst, const FwdItems<D>& fi, const BwdItems<D>& bi, long long n, const double* xinit, double* xlast) {
    cudaStream_t st = h->stream;
    const long long n1 = (n + kCsL - 1) / kCsL, n2 = (n1 + kCsL - 1) / kCsL, n3 = (n2 + kCsL - 1) / kCsL;
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

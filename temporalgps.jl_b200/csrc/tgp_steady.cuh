// tgp_steady.cuh — steady-state fast path of the forward filter for TIME-INVARIANT models with scalar
// observations (RegularSpacing + homoscedastic noise: BASELINE configs 1, 2, 4).
//
// In a time-invariant LGSSM the covariance side of the Kalman recursion (predict LGC:46-52, update
// LGC:247-257) does not depend on y and converges geometrically to a fixed point P∞. Once
// |P_t - P_{t-1}| <= tol*|P_t| (checked on the device, tol = TGP_OPT_SS_TOL) the gain K, the
// innovation variance S and the filtering covariance are constants and the mean recursion is the
// constant-coefficient affine map
//      m_t = Abar m_{t-1} + K y_t + c,   Abar = A - K w',  w = A'H',  c = a - K (H a + h)
//      v_t = y_t - w'm_{t-1} - (H a + h),    lml_t = -(log 2π + log S + v_t²/S)/2
// i.e. exactly the arithmetic step_logpdf performs (lgssm.jl:155-159) with P frozen at its limit.
// The first N0 steps (the transient) run through the general 5-tuple scan (tgp_scan_small.cuh); the
// remaining T - N0 steps run here as ONE persistent kernel:
//   phase 1  every CTA reduces its contiguous range of steps to the affine aggregate z (zero-state
//            response), reading y once, fully coalesced (thread-strided recurrence with Abar^NT);
//   barrier  one grid-wide flag barrier; every CTA folds the aggregates of the CTAs before it;
//   phase 2  tile by tile: y tile -> shared memory, per-thread chunk fold, warp scan by shuffles,
//            start state per thread, then the sequential (predict, update) per step emitting
//            v_t² (and optionally lml_t, m_t, P∞).
// HBM traffic: y once (+ once more from L2) and the requested outputs; no per-step workspace.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>

#include "tgp_ctx.cuh"
#include "tgp_scan_small.cuh"

namespace tgp {

constexpr int kSSThreads = 256;
constexpr int kSSWarps = kSSThreads / 32;

template <int D>
struct SSConst {
    int converged;
    int pad_;
    double S, invS, logS, hh, conv_err;
    Vec<D> K, w, a, c;
    Mat<D> A, Abar;
    double PfFull[D * D];  // P∞, full column-major
    Mat<D> P1[5];          // Abar^(2^k)
    Mat<D> P1w;            // Abar^32
    Mat<D> P1t;            // Abar^NT
    Mat<D> P2[5];          // Abar^(L 2^k)
    Mat<D> P2w;            // Abar^(32 L)
    Mat<D> PR[5];          // PhiR^(2^k), PhiR = Abar^R
    Mat<D> PRw;            // PhiR^32
    Mat<D> PRt;            // PhiR^NT
    Mat<D> Plane[32];      // Abar^(L lane)
};

template <int D> TGP_HD Mat<D> mat_pow(Mat<D> B, long long e) {
    Mat<D> R = meye<D>();
    while (e > 0) {
        if (e & 1) R = matmul(B, R);
        B = matmul(B, B);
        e >>= 1;
    }
    return R;
}

// One warp; every lane computes the shared constants redundantly, lane i also Plane[i].
template <int D>
__global__ void __launch_bounds__(32)
k_ss_setup(const double* __restrict__ A_, const double* __restrict__ a_, const double* __restrict__ Q_,
           const double* __restrict__ H_, const double* __restrict__ h_, const double* __restrict__ R_,
           const double* __restrict__ x_in /* packed (m, P) after the transient */, double tol, int L, long long Rsteps,
           SSConst<D>* __restrict__ out, unsigned* __restrict__ counters) {
    const int lane = threadIdx.x;
    const Mat<D> A = ldg_mat<D>(A_);
    const Vec<D> a = ldg_vec<D>(a_);
    const Sym<D> Q = ldg_sym_full<D>(Q_);
    const Vec<D> H = ldg_vec<D>(H_);
    const double h = *h_, R = *R_;
    Vec<D> m;
    Sym<D> P;
    load_state<D>(x_in, 1, 0, m, P);
    const Sym<D> Pp = congruence(A, P, Q);
    const Vec<D> V = symvec(Pp, H);
    const double S = dot(V, H) + R;
    const double invS = 1.0 / S;
    Vec<D> K;
#pragma unroll
    for (int i = 0; i < D; ++i) K[i] = V[i] * invS;
    Sym<D> Pf;
    double err = 0.0, nrm = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i <= j; ++i) {
            Pf(i, j) = fma(-V[i], K[j], Pp(i, j));
            err = fmax(err, fabs(Pf(i, j) - P(i, j)));
            nrm = fmax(nrm, fabs(Pf(i, j)));
        }
    const Vec<D> w = matTvec(A, H);
    const double hh = dot(H, a) + h;
    Mat<D> Abar;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) Abar(i, j) = fma(-K[i], w[j], A(i, j));
    // powers
    Mat<D> p = Abar, P1[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { P1[k] = p; p = matmul(p, p); }
    const Mat<D> P1w = p;                       // ^32
    const Mat<D> P1t = mat_pow(P1w, kSSThreads / 32);
    Mat<D> P2[5];
    p = mat_pow(Abar, L);
#pragma unroll
    for (int k = 0; k < 5; ++k) { P2[k] = p; p = matmul(p, p); }
    const Mat<D> P2w = p;                       // ^(32 L)
    const long long tile = (long long)kSSThreads * L;
    const Mat<D> PhiTile = mat_pow(P2w, kSSWarps);
    Mat<D> PR[5];
    p = mat_pow(PhiTile, Rsteps / tile);
#pragma unroll
    for (int k = 0; k < 5; ++k) { PR[k] = p; p = matmul(p, p); }
    const Mat<D> PRw = p;
    const Mat<D> PRt = mat_pow(PRw, kSSThreads / 32);
    out->Plane[lane] = mat_pow(P2[0], lane);
    if (lane == 0) {
        out->converged = (S > 0.0 && err <= tol * nrm) ? 1 : 0;
        out->conv_err = nrm > 0.0 ? err / nrm : 0.0;
        out->S = S; out->invS = invS; out->logS = log(S); out->hh = hh;
        out->K = K; out->w = w; out->a = a;
#pragma unroll
        for (int i = 0; i < D; ++i) out->c[i] = fma(-K[i], hh, a[i]);
        out->A = A; out->Abar = Abar;
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i) out->PfFull[i + D * j] = Pf(i, j);
#pragma unroll
        for (int k = 0; k < 5; ++k) { out->P1[k] = P1[k]; out->P2[k] = P2[k]; out->PR[k] = PR[k]; }
        out->P1w = P1w; out->P1t = P1t; out->P2w = P2w; out->PRw = PRw; out->PRt = PRt;
        counters[0] = 0u;
        counters[1] = 0u;
    }
}

struct SSOut {
    double* lml_steps;  // at SS step 0, contiguous; nullable
    double* m_f;        // at SS step 0; nullable
    long long s_m;
    double* P_f;        // at SS step 0; nullable
    long long s_P;
    double* xT;         // packed final filtering distribution
    double* partials;   // one per CTA
    const double* lml_prefix;  // lml of the transient (device), nullable
    double* lml_out;    // total
};

template <int D> __device__ __forceinline__ Vec<D> shfl_up_vec(const Vec<D>& v, int off) {
    Vec<D> r;
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = __shfl_up_sync(0xffffffffu, v[i], off);
    return r;
}
template <int D> __device__ __forceinline__ Vec<D> affine(const Mat<D>& B, const Vec<D>& z, const Vec<D>& u) {
    Vec<D> r;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = u[i];
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(B(i, j), z[j], s);
        r[i] = s;
    }
    return r;
}

// Decayed sum over the CTA: every thread holds z (aligned at its own position, positions = thread
// index); returns at thread 0.. the value  sum_i B^(NT-1-i) z_i  in `total` (valid in ALL threads via
// shared memory). Bk = B^(2^k), Bw = B^32.
template <int D>
__device__ __forceinline__ Vec<D> cta_decayed_sum(Vec<D> z, const Mat<D>* Bk, const Mat<D>& Bw, double* sh /* (kSSWarps + 1) * D */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const Vec<D> zu = shfl_up_vec(z, 1 << k);
        if (lane >= (1 << k)) z = affine(Bk[k], zu, z);
    }
    __syncthreads();
    if (lane == 31) {
#pragma unroll
        for (int i = 0; i < D; ++i) sh[w * D + i] = z[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        Vec<D> acc = vzero<D>();
        for (int ww = 0; ww < kSSWarps; ++ww) {
            Vec<D> t;
#pragma unroll
            for (int i = 0; i < D; ++i) t[i] = sh[ww * D + i];
            acc = affine(Bw, acc, t);
        }
#pragma unroll
        for (int i = 0; i < D; ++i) sh[kSSWarps * D + i] = acc[i];
    }
    __syncthreads();
    Vec<D> r;
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = sh[kSSWarps * D + i];
    return r;
}

// Dynamic shared memory layout (doubles): [SSConst<D>] [red: (kSSWarps+1)*D] [win: kSSWarps*D] [mt: D]
//                                         [ytile: NT*(L+1)] [mstage: NT*(L*D+1) if m_f]
template <int D>
__global__ void __launch_bounds__(kSSThreads)
k_ss_main(const SSConst<D>* __restrict__ cg, const double* __restrict__ y, long long Ts, int L, long long Rsteps,
          const double* __restrict__ x_in, double* __restrict__ agg, unsigned* __restrict__ counters, const SSOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SSConst<D>& c = *reinterpret_cast<SSConst<D>*>(smem_raw);
    constexpr int CW = (sizeof(SSConst<D>) + 7) / 8;
    double* red = reinterpret_cast<double*>(smem_raw) + CW;
    double* win = red + (kSSWarps + 1) * D;
    double* mt = win + kSSWarps * D;
    double* ytile = mt + D;
    double* mstage = ytile + kSSThreads * (L + 1);
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    {
        const double* src = reinterpret_cast<const double*>(cg);
        double* dst = reinterpret_cast<double*>(smem_raw);
        for (int i = tid; i < CW; i += kSSThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (!c.converged) return;  // uniform: the host falls back to the general scan

    const long long r0 = (long long)b * Rsteps;
    const long long r1 = min(r0 + Rsteps, Ts);
    const Vec<D> K = c.K;

    // ---- phase 1: zero-state response of the whole range, thread-strided (coalesced) ----------
    {
        const Mat<D> Bt = c.P1t;
        const Vec<D> cc = c.c;
        Vec<D> z = vzero<D>();
        const long long jend = Rsteps / kSSThreads;
        const double* yp = y + r0 + tid;
        const long long nfull = (r1 - r0) / kSSThreads;  // iterations where every thread is in range
        long long j = 0;
#pragma unroll 4
        for (; j < nfull; ++j) {
            const double yv = __ldg(yp + j * kSSThreads);
            Vec<D> u;
#pragma unroll
            for (int i = 0; i < D; ++i) u[i] = fma(K[i], yv, cc[i]);
            z = affine(Bt, z, u);
        }
        for (; j < jend; ++j) {
            const long long t = r0 + j * kSSThreads + tid;
            Vec<D> u = vzero<D>();
            if (t < r1) {
                const double yv = __ldg(y + t);
#pragma unroll
                for (int i = 0; i < D; ++i) u[i] = fma(K[i], yv, cc[i]);
            }
            z = affine(Bt, z, u);
        }
        const Vec<D> Z = cta_decayed_sum<D>(z, c.P1, c.P1w, red);
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) __stcg(agg + (size_t)(b + 1) * D + i, Z[i]);
            if (b == 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) __stcg(agg + i, x_in[i]);  // "aggregate" of everything before the SS region
            }
            __threadfence();
            atomicAdd(counters, 1u);
            while (*reinterpret_cast<volatile unsigned*>(counters) < (unsigned)G) { __nanosleep(64); }
            __threadfence();
        }
        __syncthreads();
    }

    // ---- incoming mean of this CTA: sum_{e=0..b} PhiR^(b-e) V[e] ---------------------------------
    {
        const long long n = (long long)b + 1;
        const long long J = (n + kSSThreads - 1) / kSSThreads;
        const long long pad = J * kSSThreads - n;
        const Mat<D> Bt = c.PRt;
        Vec<D> z = vzero<D>();
        for (long long j = 0; j < J; ++j) {
            const long long e = j * kSSThreads + tid - pad;
            Vec<D> u = vzero<D>();
            if (e >= 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) u[i] = __ldcg(agg + (size_t)e * D + i);
            }
            z = affine(Bt, z, u);
        }
        const Vec<D> m_in = cta_decayed_sum<D>(z, c.PR, c.PRw, red);
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) mt[i] = m_in[i];
        }
        __syncthreads();
    }

    // ---- phase 2 ------------------------------------------------------------------------------------
    const long long tile = (long long)kSSThreads * L;
    const Mat<D> A = c.A;
    const Vec<D> av = c.a, wv = c.w;
    const double hh = c.hh, invS = c.invS;
    const double lc = -0.5 * (kLog2Pi + c.logS);
    double q = 0.0;
    const int ys = L + 1;         // padded chunk stride in ytile
    const int msd = L * D + 1;    // padded chunk stride in mstage
    const bool m_contig = out.m_f && out.s_m == D;
    for (long long ts = r0; ts < r1; ts += tile) {
        const long long nt = min(tile, r1 - ts);
        for (long long e = tid; e < tile; e += kSSThreads) ytile[e + e / L] = e < nt ? __ldg(y + ts + e) : 0.0;
        __syncthreads();
        // chunk fold (zero-state)
        Vec<D> z = vzero<D>();
        {
            const Mat<D> Ab = c.Abar;
            const Vec<D> cc = c.c;
            const double* yc = ytile + tid * ys;
            const long long c0 = (long long)tid * L;
            for (int j = 0; j < L; ++j) {
                Vec<D> u = vzero<D>();
                if (c0 + j < nt) {
                    const double yv = yc[j];
#pragma unroll
                    for (int i = 0; i < D; ++i) u[i] = fma(K[i], yv, cc[i]);
                }
                z = affine(Ab, z, u);
            }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const Vec<D> zu = shfl_up_vec(z, 1 << k);
            if (lane >= (1 << k)) z = affine(c.P2[k], zu, z);
        }
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < D; ++i) red[wp * D + i] = z[i];
        }
        __syncthreads();
        if (tid == 0) {
            Vec<D> mw;
#pragma unroll
            for (int i = 0; i < D; ++i) mw[i] = mt[i];
            for (int ww = 0; ww < kSSWarps; ++ww) {
                Vec<D> t;
#pragma unroll
                for (int i = 0; i < D; ++i) { win[ww * D + i] = mw[i]; t[i] = red[ww * D + i]; }
                mw = affine(c.P2w, mw, t);
            }
#pragma unroll
            for (int i = 0; i < D; ++i) mt[i] = mw[i];  // state entering the next tile
        }
        __syncthreads();
        Vec<D> m;
        {
            Vec<D> ze = shfl_up_vec(z, 1);
            if (lane == 0) ze = vzero<D>();
            Vec<D> mw;
#pragma unroll
            for (int i = 0; i < D; ++i) mw[i] = win[wp * D + i];
            m = affine(c.Plane[lane], mw, ze);
        }
        // sequential (predict, update) over the chunk
        {
            double* yc = ytile + tid * ys;
            double* mc = mstage + tid * msd;
            const long long c0 = (long long)tid * L;
            for (int j = 0; j < L; ++j) {
                if (c0 + j >= nt) break;
                const double yv = yc[j];
                const double v = yv - hh - dot(wv, m);
                q = fma(v, v, q);
                Vec<D> kv;
#pragma unroll
                for (int i = 0; i < D; ++i) kv[i] = fma(K[i], v, av[i]);
                m = affine(A, m, kv);
                if (out.lml_steps) yc[j] = fma(-0.5 * invS * v, v, lc);
                if (out.m_f) {
                    if (m_contig) {
#pragma unroll
                        for (int i = 0; i < D; ++i) mc[j * D + i] = m[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < D; ++i) out.m_f[(ts + c0 + j) * out.s_m + i] = m[i];
                    }
                }
                if (ts + c0 + j == Ts - 1) {
#pragma unroll
                    for (int i = 0; i < D; ++i) out.xT[i] = m[i];
#pragma unroll
                    for (int jj = 0; jj < D; ++jj)
#pragma unroll
                        for (int ii = 0; ii <= jj; ++ii) out.xT[D + Sym<D>::idx(ii, jj)] = c.PfFull[ii + D * jj];
                }
            }
        }
        __syncthreads();
        if (out.lml_steps)
            for (long long e = tid; e < nt; e += kSSThreads) out.lml_steps[ts + e] = ytile[e + e / L];
        if (m_contig) {
            const long long ne = nt * D;
            double* dst = out.m_f + ts * D;
            for (long long g = tid; g < ne; g += kSSThreads) dst[g] = mstage[g + g / (L * D)];
        }
        if (out.P_f) {
            if (out.s_P == D * D) {
                const long long ne = nt * D * D;
                double* dst = out.P_f + ts * D * D;
                for (long long g = tid; g < ne; g += kSSThreads) dst[g] = c.PfFull[g % (D * D)];
            } else {
                for (long long e = tid; e < nt; e += kSSThreads)
#pragma unroll
                    for (int k = 0; k < D * D; ++k) out.P_f[(ts + e) * out.s_P + k] = c.PfFull[k];
            }
        }
        __syncthreads();
    }

    // ---- log-likelihood: fixed-order reduction, last CTA finishes ----------------------------------
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_down_sync(0xffffffffu, q, off);
    if (lane == 0) red[wp] = q;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < kSSWarps; ++i) t += red[i];
        __stcg(out.partials + b, t);
        __threadfence();
        const unsigned done = atomicAdd(counters + 1, 1u);
        if (done == (unsigned)G - 1) {
            __threadfence();
            double s = 0.0;
            for (int i = 0; i < G; ++i) s += __ldcg(out.partials + i);
            const double pre = out.lml_prefix ? *out.lml_prefix : 0.0;
            *out.lml_out = pre + (double)Ts * lc - 0.5 * invS * s;
        }
    }
}

// Forward declaration: the transient runs through the general scan driver (tgp_drivers.cuh).
template <int D> int filter_general(tgp_ctx* h, const tgp_lgssm& d, const double* dy, FilterReq& rq);

template <int D>
size_t ss_smem_bytes(int L, bool stage_m) {
    size_t n = (sizeof(SSConst<D>) + 7) / 8 + (kSSWarps + 1) * D + kSSWarps * D + D + (size_t)kSSThreads * (L + 1);
    if (stage_m) n += (size_t)kSSThreads * (L * D + 1);
    return n * sizeof(double);
}

// Host driver. *flag receives the device address of the convergence word (the caller reads it with
// its end-of-call copies; 0 there means "P had not converged after the transient": rerun with the
// general scan). *handled = false when the path does not apply (short series, smoother workspace).
template <int D>
int filter_steady(tgp_ctx* h, const tgp_lgssm& d, const double* dy, FilterReq& rq, bool* handled, const int** flag) {
    *handled = false;
    *flag = nullptr;
    const int64_t T = d.T;
    const int64_t N0 = h->ss_prefix > 0 ? h->ss_prefix : 4096;
    if (rq.keep_ws || T < 4 * N0) return TGP_OK;
    cudaStream_t st = h->stream;
    // transient: first N0 steps through the general scan, same output arrays
    tgp_lgssm dp = d;
    dp.T = N0;
    FilterReq rp = rq;
    rp.lml_out = nullptr;
    TGP_TRY(filter_general<D>(h, dp, dy, rp));
    rq.err = rp.err;
    rq.x0buf = rp.x0buf;

    const bool stage_m = rq.m_f && rq.s_m == D;
    int L = h->chunk > 0 ? h->chunk : (rq.m_f ? 8 : 16);
    if (L > 64) L = 64;
    const size_t smem = ss_smem_bytes<D>(L, stage_m);
    static bool attr_set = false;
    if (!attr_set) {
        TGP_CUDA(h, cudaFuncSetAttribute(k_ss_main<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    int occ = 0;
    TGP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ss_main<D>, kSSThreads, smem));
    if (occ < 1) return fail(h, TGP_ECUDA, "steady-state kernel does not fit on an SM (smem %zu B)", smem);
    const int64_t Ts = T - N0;
    const int64_t tile = (int64_t)kSSThreads * L;
    int64_t G = std::min<int64_t>((int64_t)occ * h->sm_count, (Ts + tile - 1) / tile);
    int64_t R = (Ts + G - 1) / G;
    R = (R + tile - 1) / tile * tile;
    G = (Ts + R - 1) / R;

    SSConst<D>* cst;
    double *agg, *partials, *xT;
    unsigned* counters;
    TGP_TRY(dalloc(h, 1, &cst));
    TGP_TRY(dalloc(h, (size_t)(G + 1) * D, &agg));
    TGP_TRY(dalloc(h, (size_t)G, &partials));
    TGP_TRY(dalloc(h, D + Sym<D>::N, &xT));
    TGP_TRY(dalloc(h, 2, &counters));
    TGP_TRY(dalloc(h, 1, &rq.lml_dev));
    TGP_K(h, "k_ss_setup");
    k_ss_setup<D><<<1, 32, 0, st>>>(d.A, d.a, d.Q, d.H, d.h, d.R, rp.xT, h->ss_tol, L, R, cst, counters);
    TGP_LAUNCH_CHECK(h);
    SSOut so;
    so.lml_steps = rq.lml_steps ? rq.lml_steps + N0 : nullptr;
    so.m_f = rq.m_f ? rq.m_f + N0 * rq.s_m : nullptr;
    so.s_m = rq.s_m;
    so.P_f = rq.P_f ? rq.P_f + N0 * rq.s_P : nullptr;
    so.s_P = rq.s_P;
    so.xT = xT;
    so.partials = partials;
    so.lml_prefix = rp.lml_dev;
    so.lml_out = rq.lml_dev;
    const SSConst<D>* cst_c = cst;
    const double* ysp = dy + N0;
    long long Ts_ll = Ts, R_ll = R;
    const double* xin = rp.xT;
    void* args[] = {(void*)&cst_c, (void*)&ysp, (void*)&Ts_ll, (void*)&L, (void*)&R_ll, (void*)&xin, (void*)&agg, (void*)&counters, (void*)&so};
    TGP_K(h, "k_ss_main");
    TGP_CUDA(h, cudaLaunchCooperativeKernel((const void*)k_ss_main<D>, dim3((unsigned)G), dim3(kSSThreads), args, smem, st));
    TGP_LAUNCH_CHECK(h);
    rq.xT = xT;
    *flag = &cst->converged;
    *handled = true;
    return deliver_scalar(h, rq.lml_dev, rq.lml_out);
}

}  // namespace tgp

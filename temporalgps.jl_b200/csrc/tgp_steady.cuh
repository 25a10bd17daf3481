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
//
// Two launches:
//   k_transient  ONE CTA runs the general 5-tuple scan (tgp_math.cuh) over blocks of 2048 steps until
//                the covariance has converged (usually the first block), emits the requested outputs
//                of those steps, then derives every constant of the steady phase (gain, powers of
//                Abar for the scans) on the device.
//   k_ss_main    persistent cooperative kernel over the remaining steps, CTA b owning a contiguous
//                range of R steps:
//       phase 1  tile by tile (cp.async double-buffered into padded shared memory): per-thread
//                chunk fold of the zero-state response, warp scan by shuffles, tile-exclusive
//                prefix per thread -> 24 B / chunk scratch (stays in L2); range aggregate;
//       barrier  one grid-wide counter barrier; every CTA folds the aggregates before it;
//       phase 2  tile by tile: start state per thread = power * tile-in state + its prefix, then
//                the sequential (predict, update) per step: v_t² (and optionally lml_t, m_t, P∞).
// HBM traffic: y once (second pass is served by L2) + the requested outputs.
#pragma once
#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "tgp_ctx.cuh"
#include "tgp_scan_small.cuh"

namespace tgp {

constexpr int kSSThreads = 256;
constexpr int kSSWarps = kSSThreads / 32;
constexpr int kTrThreads = 128;              // transient CTA
constexpr int kTrWarps = kTrThreads / 32;
constexpr int kTrL = 16;
constexpr int kTrBlock = kTrThreads * kTrL;  // 2048 steps per transient block
constexpr int kSqN = 44;                     // squares table Abar^(2^k), k < 44

template <int D>
struct SSConst {
    int converged;
    int n_blocks;
    long long N0, Ts, R;  // transient length used, steady steps, steps per CTA range
    double S, invS, logS, hh, conv_err;
    Vec<D> K, w, a, c, x_in;  // x_in: filtered mean after the transient
    Mat<D> A, Abar;
    double PfFull[D * D];     // P∞, full column-major
    Mat<D> P2[5];             // Abar^(L 2^k)
    Mat<D> Pw[kSSWarps];      // Abar^(32 L w)
    Mat<D> PhiTile;           // Abar^(NT L)
    Mat<D> PR[5];             // PhiR^(2^k), PhiR = Abar^R
    Mat<D> PRw;               // PhiR^32
    Mat<D> PRt;               // PhiR^NT
    Mat<D> Plane[32];         // Abar^(L lane)
};

struct SSOut {
    double* lml_steps;  // memory order, index = time; nullable
    double* m_f;        // nullable
    long long s_m;
    double* P_f;        // nullable
    long long s_P;
    double* xT;         // packed final filtering distribution
    double* partials;   // one per CTA
    const double* lml_prefix;  // lml of the transient (device)
    double* lml_out;    // total
};

template <int D> __device__ __forceinline__ Vec<D> shfl_up_vec(const Vec<D>& v, int off) {
    Vec<D> r;
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = __shfl_up_sync(0xffffffffu, v[i], off);
    return r;
}
// B z + u
template <int D> TGP_HD Vec<D> affine(const Mat<D>& B, const Vec<D>& z, const Vec<D>& u) {
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

// B^e from the table of squares sq[k] = B^(2^k) (shared memory); lanes may pass different e.
template <int D> __device__ __forceinline__ Mat<D> pow_from_squares(const Mat<D>* sq, unsigned long long e) {
    Mat<D> R = meye<D>();
    for (int k = 0; k < kSqN && (e >> k); ++k)
        if ((e >> k) & 1ull) R = matmul(sq[k], R);
    return R;
}

// =============================================================================================
// Transient + set-up: one CTA.
// =============================================================================================
template <int D>
__global__ void __launch_bounds__(kTrThreads)
k_transient(const DevModel dm, const double* __restrict__ m0, const double* __restrict__ P0, int max_blocks, double tol, int ssL,
            int G, const FilterOut out, SSConst<D>* __restrict__ cst, unsigned* __restrict__ counters, double* __restrict__ lml_prefix) {
    __shared__ Elem<D> tot[kTrWarps];
    __shared__ double blk_state[2][D + Sym<D>::N];
    __shared__ double red[kTrWarps];
    __shared__ int s_conv;
    __shared__ Mat<D> sq[kSqN];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const Mat<D> A = ldg_mat<D>(dm.A);
    const Vec<D> a = ldg_vec<D>(dm.a);
    const Sym<D> Q = ldg_sym_full<D>(dm.Q);
    const Vec<D> H = ldg_vec<D>(dm.H);
    const double h = __ldg(dm.h), R = __ldg(dm.R);
    const StepConst<D> sc = make_step_const<D>(A, a, Q, H, h, R);
    if (tid == 0) {
        store_state<D>(blk_state[0], 1, 0, ldg_vec<D>(m0), ldg_sym_full<D>(P0));
        s_conv = 0;
    }
    __syncthreads();
    double quad_sum = 0.0, lml_direct = 0.0;
    LogAcc la;
    long long n_steps = 0;
    int nb = 0;
    double conv_err = 0.0;
    for (int b = 0; b < max_blocks; ++b) {
        const long long s = (long long)b * kTrBlock + (long long)tid * kTrL;
        // phase 1: chunk fold + warp scan
        Elem<D> E = elem_identity<D>();
#pragma unroll 1
        for (int j = 0; j < kTrL; ++j) fold_step(E, sc, __ldg(dm.y + s + j));
#pragma unroll 1
        for (int off = 1; off < 32; off <<= 1) {
            const Elem<D> O = shfl_up_elem(E, off);
            if (lane >= off) E = combine(O, E);
        }
        if (lane == 31) tot[wp] = E;
        const Elem<D> X = shfl_up_elem(E, 1);
        __syncthreads();
        // state entering this thread's chunk
        Vec<D> m;
        Sym<D> P;
        load_state<D>(blk_state[b & 1], 1, 0, m, P);
        for (int ww = 0; ww < wp; ++ww) apply_elem(tot[ww], m, P);
        if (lane > 0) apply_elem(X, m, P);
        // phase 2: the ordinary (predict, update) over the chunk
#pragma unroll 1
        for (int j = 0; j < kTrL; ++j) {
            const long long n = s + j;
            const double y = __ldg(dm.y + n);
            predict(m, P, A, a, Q);
            double quad;
            double S = update_scalar(m, P, H, h, R, y, &quad);
            if (!(S > 1e-300) || !(S < 1e300)) { atomicMin(out.err_step, (unsigned long long)n); S = 1.0; }
            if (out.lml_steps) {
                const double l = lml_from(S, quad);
                out.lml_steps[n] = l;
                lml_direct += l;
            } else {
                la.add(S);
                quad_sum += quad;
            }
            if (out.m_f) {
#pragma unroll
                for (int i = 0; i < D; ++i) out.m_f[n * out.s_m + i] = m[i];
            }
            if (out.P_f) {
#pragma unroll
                for (int jj = 0; jj < D; ++jj)
#pragma unroll
                    for (int ii = 0; ii < D; ++ii) out.P_f[n * out.s_P + ii + D * jj] = P(ii, jj);
            }
        }
        n_steps += kTrL;
        if (tid == kTrThreads - 1) store_state<D>(blk_state[(b + 1) & 1], 1, 0, m, P);
        __syncthreads();
        nb = b + 1;
        if (tid == 0) {  // one more covariance step: has P reached its fixed point?
            Vec<D> mm;
            Sym<D> PP;
            load_state<D>(blk_state[nb & 1], 1, 0, mm, PP);
            const Sym<D> Pp = congruence(A, PP, Q);
            const Vec<D> V = symvec(Pp, H);
            const double S = dot(V, H) + R;
            double err = 0.0, nrm = 0.0;
#pragma unroll
            for (int jj = 0; jj < D; ++jj)
#pragma unroll
                for (int ii = 0; ii <= jj; ++ii) {
                    const double pf = fma(-V[ii], V[jj] / S, Pp(ii, jj));
                    err = fmax(err, fabs(pf - PP(ii, jj)));
                    nrm = fmax(nrm, fabs(pf));
                }
            s_conv = (S > 0.0 && err <= tol * nrm) ? 1 : 0;
            red[0] = nrm > 0.0 ? err / nrm : 0.0;
        }
        __syncthreads();
        conv_err = red[0];
        if (s_conv) break;
        __syncthreads();
    }
    // ---- lml of the transient: fixed-order CTA reduction ------------------------------------------
    double part = out.lml_steps ? lml_direct : -0.5 * ((double)n_steps * kLog2Pi + la.total() + quad_sum);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
    __syncthreads();
    if (lane == 0) red[wp] = part;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kTrWarps; ++i) t += red[i];
        *lml_prefix = t;
    }
    if (wp != 0) return;
    // ---- constants of the steady phase (warp 0) ---------------------------------------------------
    Vec<D> mT;
    Sym<D> PT;
    load_state<D>(blk_state[nb & 1], 1, 0, mT, PT);
    const Sym<D> Pp = congruence(A, PT, Q);
    const Vec<D> V = symvec(Pp, H);
    const double S = dot(V, H) + R;
    const double invS = 1.0 / S;
    Vec<D> K;
#pragma unroll
    for (int i = 0; i < D; ++i) K[i] = V[i] * invS;
    const Vec<D> w = matTvec(A, H);
    const double hh = dot(H, a) + h;
    Mat<D> Abar;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) Abar(i, j) = fma(-K[i], w[j], A(i, j));
    {   // squares table (every lane computes it; lane 0 publishes)
        Mat<D> p = Abar;
        for (int k = 0; k < kSqN; ++k) {
            if (lane == 0) sq[k] = p;
            p = matmul(p, p);
        }
    }
    __syncwarp();
    const long long N0 = (long long)nb * kTrBlock;
    const long long Ts = dm.T - N0;
    const long long tile = (long long)kSSThreads * ssL;
    long long Rr = (Ts + G - 1) / G;
    Rr = (Rr + tile - 1) / tile * tile;
    if (Rr < tile) Rr = tile;
    cst->Plane[lane] = pow_from_squares<D>(sq, (unsigned long long)ssL * lane);
    if (lane < 5) cst->P2[lane] = pow_from_squares<D>(sq, (unsigned long long)ssL << lane);
    else if (lane < 5 + kSSWarps) cst->Pw[lane - 5] = pow_from_squares<D>(sq, 32ull * ssL * (lane - 5));
    else if (lane == 13) cst->PhiTile = pow_from_squares<D>(sq, (unsigned long long)tile);
    else if (lane == 14) {
        Mat<D> p = pow_from_squares<D>(sq, (unsigned long long)Rr);
        for (int k = 0; k < 5; ++k) { cst->PR[k] = p; p = matmul(p, p); }
        cst->PRw = p;                               // PhiR^32
        for (int k = 0; k < 3; ++k) p = matmul(p, p);
        cst->PRt = p;                               // PhiR^256
    } else if (lane == 15) {
        cst->converged = s_conv;
        cst->n_blocks = nb;
        cst->N0 = N0; cst->Ts = Ts; cst->R = Rr;
        cst->S = S; cst->invS = invS; cst->logS = log(S); cst->hh = hh; cst->conv_err = conv_err;
        cst->K = K; cst->w = w; cst->a = a; cst->x_in = mT;
#pragma unroll
        for (int i = 0; i < D; ++i) cst->c[i] = fma(-K[i], hh, a[i]);
        cst->A = A; cst->Abar = Abar;
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i) cst->PfFull[i + D * j] = fma(-V[i], K[j], Pp(i, j));
        counters[0] = 0u;
        counters[1] = 0u;
    }
    static_assert(kSSThreads == 256, "PRt assumes 256 threads");
}

// =============================================================================================
// Steady phase.
// =============================================================================================
// Decayed sum over the CTA: thread i holds z_i; returns sum_i B^(NT-1-i) z_i in every thread.
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

template <int D, int L>
struct SSLayout {
    static constexpr int CW = (sizeof(SSConst<D>) + 7) / 8;
    static constexpr int YS = L + 1;                 // padded chunk stride (odd: conflict-free LDS.64)
    static constexpr int YB = kSSThreads * YS;       // one tile buffer
    static constexpr int MS = L * D + 1;             // padded chunk stride of the m_f staging tile
    static constexpr int o_red = CW;
    static constexpr int o_win = o_red + (kSSWarps + 1) * D;
    static constexpr int o_mt = o_win + kSSWarps * D;
    static constexpr int o_y = o_mt + D + ((o_mt + D) & 1);
    static constexpr int o_ms = o_y + 2 * YB;
    static size_t bytes(bool stage_m) { return (size_t)(o_ms + (stage_m ? kSSThreads * MS : 0)) * sizeof(double); }
};

template <int D, int L>
__global__ void __launch_bounds__(kSSThreads, 2)
k_ss_main(const SSConst<D>* __restrict__ cg, const double* __restrict__ y_all, double* __restrict__ zbuf, long long zstride,
          double* __restrict__ agg, unsigned* __restrict__ counters, const SSOut out) {
    using LY = SSLayout<D, L>;
    extern __shared__ __align__(16) double smem[];
    SSConst<D>& c = *reinterpret_cast<SSConst<D>*>(smem);
    double* red = smem + LY::o_red;
    double* win = smem + LY::o_win;
    double* mt = smem + LY::o_mt;
    double* ybuf = smem + LY::o_y;
    double* mstage = smem + LY::o_ms;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    {
        const double* src = reinterpret_cast<const double*>(cg);
        for (int i = tid; i < LY::CW; i += kSSThreads) smem[i] = __ldg(src + i);
    }
    __syncthreads();
    if (!c.converged) return;  // uniform: the host reruns the series with the general scan

    constexpr long long tile = (long long)kSSThreads * L;
    const long long N0 = c.N0, Ts = c.Ts, Rr = c.R;
    const double* __restrict__ y = y_all + N0;
    const long long r0 = min((long long)b * Rr, Ts);
    const long long r1 = min(r0 + Rr, Ts);
    const Vec<D> K = c.K;

    auto issue_tile = [&](long long ts, int buf) {
        double* dst = ybuf + buf * LY::YB;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const int e = tid + k * kSSThreads;
            const long long t = ts + e;
            const bool ok = t < r1;
            __pipeline_memcpy_async(dst + e + e / L, y + (ok ? t : r1 - 1), 8, ok ? 0 : 8);
        }
    };

    // ---- phase 1: zero-state responses ------------------------------------------------------------
    {
        Vec<D> Zr = vzero<D>();
        if (r0 < r1) issue_tile(r0, 0);
        __pipeline_commit();
        int it = 0;
        for (long long ts = r0; ts < r1; ts += tile, ++it) {
            if (ts + tile < r1) issue_tile(ts + tile, (it + 1) & 1);
            __pipeline_commit();
            __pipeline_wait_prior(1);
            __syncthreads();
            const double* yc = ybuf + (it & 1) * LY::YB + tid * LY::YS;
            Vec<D> z = vzero<D>();
            {
                const Mat<D> Ab = c.Abar;
                const Vec<D> cc = c.c;
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    const double yv = yc[j];
                    Vec<D> u;
#pragma unroll
                    for (int i = 0; i < D; ++i) u[i] = fma(K[i], yv, cc[i]);
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
                Vec<D> acc = vzero<D>();
                for (int ww = 0; ww < kSSWarps; ++ww) {
                    Vec<D> t;
#pragma unroll
                    for (int i = 0; i < D; ++i) { win[ww * D + i] = acc[i]; t[i] = red[ww * D + i]; }
                    acc = affine(c.Pw[1], acc, t);
                }
                Zr = affine(c.PhiTile, Zr, acc);
            }
            __syncthreads();
            Vec<D> ze = shfl_up_vec(z, 1);
            if (lane == 0) ze = vzero<D>();
            Vec<D> mw;
#pragma unroll
            for (int i = 0; i < D; ++i) mw[i] = win[wp * D + i];
            const Vec<D> zx = affine(c.Plane[lane], mw, ze);
            const long long cidx = ts / L + tid;
#pragma unroll
            for (int i = 0; i < D; ++i) __stcg(zbuf + i * zstride + cidx, zx[i]);
        }
        // a range shorter than Rr (the last one) is never consumed, so no alignment fix-up is needed
        if (r0 < r1) issue_tile(r0, 0);  // prefetch the first tile of phase 2 across the barrier
        __pipeline_commit();
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) __stcg(agg + (size_t)b * D + i, Zr[i]);
            __threadfence();
            atomicAdd(counters, 1u);
            while (*reinterpret_cast<volatile unsigned*>(counters) < (unsigned)G) { __nanosleep(32); }
            __threadfence();
        }
        __syncthreads();
    }

    // ---- mean entering this CTA's range: PhiR^b x_in + sum_{e<b} PhiR^(b-1-e) Z_e ------------------
    {
        const long long n = (long long)b + 1;   // virtual elements V[0] = x_in, V[e] = Z_{e-1}
        const long long J = (n + kSSThreads - 1) / kSSThreads;
        const long long pad = J * kSSThreads - n;
        const Mat<D> Bt = c.PRt;
        Vec<D> z = vzero<D>();
        for (long long j = 0; j < J; ++j) {
            const long long e = j * kSSThreads + tid - pad;
            Vec<D> u = vzero<D>();
            if (e == 0) u = c.x_in;
            else if (e > 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) u[i] = __ldcg(agg + (size_t)(e - 1) * D + i);
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
    double q = 0.0;
    {
        const Mat<D> A = c.A;
        const Vec<D> av = c.a, wv = c.w;
        const double hh = c.hh, invS = c.invS;
        const double lc = -0.5 * (kLog2Pi + c.logS);
        const bool m_contig = out.m_f && out.s_m == D;
        int it = 0;
        for (long long ts = r0; ts < r1; ts += tile, ++it) {
            if (ts + tile < r1) issue_tile(ts + tile, (it + 1) & 1);
            __pipeline_commit();
            const long long cidx = ts / L + tid;
            Vec<D> zx;
#pragma unroll
            for (int i = 0; i < D; ++i) zx[i] = __ldcg(zbuf + i * zstride + cidx);
            __pipeline_wait_prior(1);
            __syncthreads();
            Vec<D> m;
            {
                Vec<D> mti;
#pragma unroll
                for (int i = 0; i < D; ++i) mti[i] = mt[i];
                const Vec<D> mw = matvec(c.Pw[wp], mti);
                m = affine(c.Plane[lane], mw, zx);
            }
            double* yc = ybuf + (it & 1) * LY::YB + tid * LY::YS;
            double* mc = mstage + tid * LY::MS;
            const long long t0 = ts + (long long)tid * L;
            const int nv = (int)max(0ll, min((long long)L, r1 - t0));  // valid steps of this chunk
            if (nv == L) {
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    const double v = yc[j] - hh - dot(wv, m);
                    q = fma(v, v, q);
                    Vec<D> kv;
#pragma unroll
                    for (int i = 0; i < D; ++i) kv[i] = fma(K[i], v, av[i]);
                    m = affine(A, m, kv);
                    if (out.lml_steps) yc[j] = fma(-0.5 * invS * v, v, lc);
                    if (m_contig) {
#pragma unroll
                        for (int i = 0; i < D; ++i) mc[j * D + i] = m[i];
                    } else if (out.m_f) {
#pragma unroll
                        for (int i = 0; i < D; ++i) out.m_f[(N0 + t0 + j) * out.s_m + i] = m[i];
                    }
                }
            } else {
                for (int j = 0; j < nv; ++j) {
                    const double v = yc[j] - hh - dot(wv, m);
                    q = fma(v, v, q);
                    Vec<D> kv;
#pragma unroll
                    for (int i = 0; i < D; ++i) kv[i] = fma(K[i], v, av[i]);
                    m = affine(A, m, kv);
                    if (out.lml_steps) yc[j] = fma(-0.5 * invS * v, v, lc);
                    if (m_contig) {
#pragma unroll
                        for (int i = 0; i < D; ++i) mc[j * D + i] = m[i];
                    } else if (out.m_f) {
#pragma unroll
                        for (int i = 0; i < D; ++i) out.m_f[(N0 + t0 + j) * out.s_m + i] = m[i];
                    }
                }
            }
            if (nv > 0 && t0 + nv == Ts) {  // owner of the last step of the series
#pragma unroll
                for (int i = 0; i < D; ++i) out.xT[i] = m[i];
#pragma unroll
                for (int jj = 0; jj < D; ++jj)
#pragma unroll
                    for (int ii = 0; ii <= jj; ++ii) out.xT[D + Sym<D>::idx(ii, jj)] = c.PfFull[ii + D * jj];
            }
            __syncthreads();  // every thread has read mt
            if (tid == kSSThreads - 1) {
#pragma unroll
                for (int i = 0; i < D; ++i) mt[i] = m[i];  // state entering the next tile
            }
            const long long nt = min(tile, r1 - ts);
            if (out.lml_steps) {
                const double* yb = ybuf + (it & 1) * LY::YB;
                for (int e = tid; e < nt; e += kSSThreads) out.lml_steps[N0 + ts + e] = yb[e + e / L];
            }
            if (m_contig) {
                const int ne = (int)nt * D;
                double* dst = out.m_f + (N0 + ts) * D;
                for (int g = tid; g < ne; g += kSSThreads) dst[g] = mstage[g + g / (L * D)];
            }
            if (out.P_f) {
                if (out.s_P == D * D) {
                    const int ne = (int)nt * D * D;
                    double* dst = out.P_f + (N0 + ts) * D * D;
                    for (int g = tid; g < ne; g += kSSThreads) dst[g] = c.PfFull[g % (D * D)];
                } else {
                    for (int e = tid; e < nt; e += kSSThreads)
#pragma unroll
                        for (int k = 0; k < D * D; ++k) out.P_f[(N0 + ts + e) * out.s_P + k] = c.PfFull[k];
                }
            }
            __syncthreads();
        }
        __pipeline_wait_prior(0);
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
            *out.lml_out = *out.lml_prefix + (double)Ts * (-0.5 * (kLog2Pi + c.logS)) - 0.5 * c.invS * s;
        }
    }
}

// Forward declaration: a non-converged series is redone by the general scan driver (tgp_drivers.cuh).
template <int D> int filter_general(tgp_ctx* h, const tgp_lgssm& d, const double* dy, FilterReq& rq);

template <int D, int L>
int launch_ss_main(tgp_ctx* h, bool stage_m, const SSConst<D>* cst, const double* dy, int64_t T, int G, double* agg,
                   unsigned* counters, const SSOut& so) {
    using LY = SSLayout<D, L>;
    const size_t smem = LY::bytes(stage_m);
    double* zbuf;
    const long long zstride = (T + (long long)G * kSSThreads * L) / L + kSSThreads;
    TGP_TRY(dalloc(h, (size_t)zstride * D, &zbuf));
    void* args[] = {(void*)&cst, (void*)&dy, (void*)&zbuf, (void*)&zstride, (void*)&agg, (void*)&counters, (void*)&so};
    TGP_K(h, "k_ss_main");
    TGP_CUDA(h, cudaLaunchCooperativeKernel((const void*)k_ss_main<D, L>, dim3((unsigned)G), dim3(kSSThreads), args, smem, h->stream));
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

template <int D, int L>
int ss_grid(tgp_ctx* h, bool stage_m, int* G) {
    using LY = SSLayout<D, L>;
    const size_t smem = LY::bytes(stage_m);
    static bool attr_set = false;
    if (!attr_set) {
        TGP_CUDA(h, cudaFuncSetAttribute(k_ss_main<D, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    int occ = 0;
    TGP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ss_main<D, L>, kSSThreads, smem));
    if (occ < 1) return fail(h, TGP_ECUDA, "steady-state kernel does not fit on an SM (%zu B of shared memory)", smem);
    *G = std::min(occ, 2) * h->sm_count;  // a whole number of CTAs per SM: the FP64 pipe is the bound, keep SMs balanced
    return TGP_OK;
}

// Host driver. *flag receives the device address of the convergence word (the caller reads it with
// its end-of-call copies; 0 there means "P had not converged within the transient budget": rerun with
// the general scan). *handled = false when the path does not apply (short series, smoother workspace).
template <int D>
int filter_steady(tgp_ctx* h, const tgp_lgssm& d, const double* dy, FilterReq& rq, bool* handled, const int** flag) {
    *handled = false;
    *flag = nullptr;
    const int64_t T = d.T;
    if (rq.keep_ws || T < 65536) return TGP_OK;
    int64_t max_blocks = h->ss_prefix > 0 ? (h->ss_prefix + kTrBlock - 1) / kTrBlock : 8;
    max_blocks = std::max<int64_t>(1, std::min<int64_t>(max_blocks, T / (2 * kTrBlock)));
    cudaStream_t st = h->stream;
    const bool stage_m = rq.m_f && rq.s_m == D;
    const bool small_L = rq.m_f != nullptr;   // m_f staging tile must fit next to the y tiles
    int G = 0;
    if (small_L) TGP_TRY((ss_grid<D, 8>(h, stage_m, &G)));
    else TGP_TRY((ss_grid<D, 16>(h, stage_m, &G)));
    const int L = small_L ? 8 : 16;

    SSConst<D>* cst;
    double *agg, *partials, *xT, *lml_prefix;
    unsigned* counters;
    TGP_TRY(dalloc(h, 1, &rq.err));
    TGP_CUDA(h, cudaMemsetAsync(rq.err, 0xFF, sizeof(unsigned long long), st));
    TGP_TRY(dalloc(h, 1, &cst));
    TGP_TRY(dalloc(h, (size_t)(G + 1) * D, &agg));
    TGP_TRY(dalloc(h, (size_t)G, &partials));
    TGP_TRY(dalloc(h, D + Sym<D>::N, &xT));
    TGP_TRY(dalloc(h, 2, &counters));
    TGP_TRY(dalloc(h, 1, &lml_prefix));
    TGP_TRY(dalloc(h, 1, &rq.lml_dev));
    DevModel dm{d.A, d.a, d.Q, d.H, d.h, d.R, 0, 0, 0, 0, 0, 0, dy, 1, T};
    FilterOut fo;
    fo.lml_steps = rq.lml_steps; fo.s_l = 1;
    fo.m_f = rq.m_f; fo.s_m = rq.s_m;
    fo.P_f = rq.P_f; fo.s_P = rq.s_P;
    fo.ws_m = nullptr; fo.partials = nullptr;
    fo.err_step = rq.err;
    TGP_K(h, "k_transient");
    k_transient<D><<<1, kTrThreads, 0, st>>>(dm, d.m0, d.P0, (int)max_blocks, h->ss_tol, L, G, fo, cst, counters, lml_prefix);
    TGP_LAUNCH_CHECK(h);
    SSOut so;
    so.lml_steps = rq.lml_steps;
    so.m_f = rq.m_f; so.s_m = rq.s_m;
    so.P_f = rq.P_f; so.s_P = rq.s_P;
    so.xT = xT;
    so.partials = partials;
    so.lml_prefix = lml_prefix;
    so.lml_out = rq.lml_dev;
    if (small_L) TGP_TRY((launch_ss_main<D, 8>(h, stage_m, cst, dy, T, G, agg, counters, so)));
    else TGP_TRY((launch_ss_main<D, 16>(h, stage_m, cst, dy, T, G, agg, counters, so)));
    rq.xT = xT;
    rq.x0buf = nullptr;
    *flag = &cst->converged;
    *handled = true;
    return deliver_scalar(h, rq.lml_dev, rq.lml_out);
}

}  // namespace tgp

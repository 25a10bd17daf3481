// tgp_scan_small.cuh — sm_100a kernels of the general (time-varying) path for small latent
// dimension: one thread owns a contiguous chunk of L time steps and keeps its scan element in
// registers (FP64 SIMT; no tensor cores — SURVEY.md §2 kernel inventory K1/K2/K3).
//
// Filtering (replaces scan_emit + step_logpdf/step_filter, scan.jl:22-25, lgssm.jl:153-187):
//   k_filter_reduce   fold each chunk into a 5-tuple, Kogge–Stone scan inside each warp by
//                     shuffles; writes the lane-exclusive prefix per thread and one aggregate
//                     per warp.
//   k_filter_mid      one CTA scans the warp aggregates and emits the filtering distribution
//                     entering every warp.
//   k_filter_apply    every thread applies its exclusive prefix to its warp's incoming state and
//                     re-runs the ordinary Kalman step over its chunk, emitting lml / (m, P).
// Smoothing (replaces step_posterior + invert_dynamics + reverse step_marginals,
// lgssm.jl:111-115, 215-240): the same three-kernel shape over the (A, b, C) semigroup, in
// reverse time, with elements rebuilt from the stored filtering distributions.
#pragma once
#include <cuda_runtime.h>

#include "tgp_math.cuh"

namespace tgp {

constexpr int kBlock = 128;      // threads per CTA in the chunk kernels
constexpr int kMidThreads = 256; // threads of the single mid CTA

// Model arrays positioned at processing step 0, signed strides per processing step.
struct DevModel {
    const double *A, *a, *Q, *H, *h, *R;
    long long sA, sa, sQ, sH, sh, sR;
    const double* y;
    long long sy;
    long long T;   // number of scan steps
};

// Time-invariant model, built on the device by k_const_model (the model arrays may be device
// pointers, so no host arithmetic is involved) and read through the read-only path.
template <int D>
struct ConstModel {
    Mat<D> A;
    Vec<D> a;
    Sym<D> Q;
    Vec<D> H;
    double h, R;
    StepConst<D> sc;
};

template <int D> __device__ __forceinline__ Mat<D> ldg_mat(const double* p) {
    Mat<D> m;
#pragma unroll
    for (int i = 0; i < D * D; ++i) m.v[i] = __ldg(p + i);
    return m;
}
template <int D> __device__ __forceinline__ Vec<D> ldg_vec(const double* p) {
    Vec<D> m;
#pragma unroll
    for (int i = 0; i < D; ++i) m.v[i] = __ldg(p + i);
    return m;
}
template <int D> __device__ __forceinline__ Sym<D> ldg_sym_full(const double* p) {  // from full col-major
    Sym<D> s;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i <= j; ++i) s(i, j) = __ldg(p + i + D * j);
    return s;
}

// ---- SoA element I/O: component c of item i lives at base[c * stride + i] -------------------
template <int D> __device__ __forceinline__ void store_elem(double* base, long long stride, long long i, const Elem<D>& e) {
    double* p = base + i;
#pragma unroll
    for (int k = 0; k < D * D; ++k) { *p = e.A.v[k]; p += stride; }
#pragma unroll
    for (int k = 0; k < D; ++k) { *p = e.b.v[k]; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { *p = e.C.v[k]; p += stride; }
#pragma unroll
    for (int k = 0; k < D; ++k) { *p = e.eta.v[k]; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { *p = e.J.v[k]; p += stride; }
}
template <int D> __device__ __forceinline__ Elem<D> load_elem(const double* base, long long stride, long long i) {
    Elem<D> e;
    const double* p = base + i;
#pragma unroll
    for (int k = 0; k < D * D; ++k) { e.A.v[k] = *p; p += stride; }
#pragma unroll
    for (int k = 0; k < D; ++k) { e.b.v[k] = *p; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { e.C.v[k] = *p; p += stride; }
#pragma unroll
    for (int k = 0; k < D; ++k) { e.eta.v[k] = *p; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { e.J.v[k] = *p; p += stride; }
    return e;
}
template <int D> __device__ __forceinline__ Elem<D> shfl_up_elem(const Elem<D>& e, int off) {
    Elem<D> r;
#pragma unroll
    for (int k = 0; k < D * D; ++k) r.A.v[k] = __shfl_up_sync(0xffffffffu, e.A.v[k], off);
#pragma unroll
    for (int k = 0; k < D; ++k) r.b.v[k] = __shfl_up_sync(0xffffffffu, e.b.v[k], off);
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) r.C.v[k] = __shfl_up_sync(0xffffffffu, e.C.v[k], off);
#pragma unroll
    for (int k = 0; k < D; ++k) r.eta.v[k] = __shfl_up_sync(0xffffffffu, e.eta.v[k], off);
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) r.J.v[k] = __shfl_up_sync(0xffffffffu, e.J.v[k], off);
    return r;
}
template <int D> __device__ __forceinline__ void store_state(double* base, long long stride, long long i, const Vec<D>& m, const Sym<D>& P) {
    double* p = base + i;
#pragma unroll
    for (int k = 0; k < D; ++k) { *p = m.v[k]; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { *p = P.v[k]; p += stride; }
}
template <int D> __device__ __forceinline__ void load_state(const double* base, long long stride, long long i, Vec<D>& m, Sym<D>& P) {
    const double* p = base + i;
#pragma unroll
    for (int k = 0; k < D; ++k) { m.v[k] = *p; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { P.v[k] = *p; p += stride; }
}

template <int D> __device__ __forceinline__ StepConst<D> step_const_at(const DevModel& dm, long long n) {
    return make_step_const<D>(ldg_mat<D>(dm.A + n * dm.sA), ldg_vec<D>(dm.a + n * dm.sa), ldg_sym_full<D>(dm.Q + n * dm.sQ),
                              ldg_vec<D>(dm.H + n * dm.sH), __ldg(dm.h + n * dm.sh), __ldg(dm.R + n * dm.sR));
}

template <int D>
__global__ void k_const_model(const DevModel dm, ConstModel<D>* __restrict__ cm) {
    if (threadIdx.x || blockIdx.x) return;
    cm->A = ldg_mat<D>(dm.A);
    cm->a = ldg_vec<D>(dm.a);
    cm->Q = ldg_sym_full<D>(dm.Q);
    cm->H = ldg_vec<D>(dm.H);
    cm->h = *dm.h;
    cm->R = *dm.R;
    cm->sc = step_const_at<D>(dm, 0);
}

// =============================================================================================
// Filtering, phase 1
// =============================================================================================
template <int D, bool TV>
__global__ void __launch_bounds__(kBlock)
k_filter_reduce(const DevModel dm, const ConstModel<D>* __restrict__ cmp, int L, long long nthreads,
                double* __restrict__ excl, double* __restrict__ wagg, long long nwarps) {
    const long long tid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long s = tid * L;
    const long long e = min(s + (long long)L, dm.T);
    Elem<D> E = elem_identity<D>();
    StepConst<D> sc;
    if (!TV) sc = cmp->sc;
    for (long long n = s; n < e; ++n) {
        const double y = __ldg(dm.y + n * dm.sy);
        if (TV) sc = step_const_at<D>(dm, n);
        fold_step(E, sc, y);
    }
    // inclusive Kogge–Stone scan over the 32 chunk aggregates of this warp
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        const Elem<D> O = shfl_up_elem(E, off);
        if (lane >= off) E = combine(O, E);
    }
    Elem<D> X = shfl_up_elem(E, 1);
    if (lane == 0) X = elem_identity<D>();
    if (tid < nthreads) store_elem(excl, nthreads, tid, X);
    if (lane == 31) store_elem(wagg, nwarps, tid >> 5, E);
}

// =============================================================================================
// Filtering, mid: scan of warp aggregates by one CTA -> incoming state of every warp
//   x0buf: [m (D), P packed (SymN)];  wstate: SoA (D + SymN) x nwarps;  xT: final state
// =============================================================================================
template <int D>
__global__ void __launch_bounds__(kMidThreads)
k_filter_mid(const double* __restrict__ wagg, long long nwarps, const double* __restrict__ x0buf,
             double* __restrict__ wstate, double* __restrict__ xT) {
    __shared__ Elem<D> tot[kMidThreads / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long q = (nwarps + kMidThreads - 1) / kMidThreads;
    const long long s = (long long)t * q, e = min(s + q, nwarps);
    Elem<D> E = elem_identity<D>();
    for (long long j = s; j < e; ++j) E = combine(E, load_elem<D>(wagg, nwarps, j));
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        const Elem<D> O = shfl_up_elem(E, off);
        if (lane >= off) E = combine(O, E);
    }
    if (lane == 31) tot[w] = E;
    Elem<D> X = shfl_up_elem(E, 1);
    if (lane == 0) X = elem_identity<D>();
    __syncthreads();
    Vec<D> m;
    Sym<D> P;
#pragma unroll
    for (int k = 0; k < D; ++k) m.v[k] = x0buf[k];
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) P.v[k] = x0buf[D + k];
    for (int ww = 0; ww < w; ++ww) apply_elem(tot[ww], m, P);
    apply_elem(X, m, P);
    for (long long j = s; j < e; ++j) {
        store_state<D>(wstate, nwarps, j, m, P);
        apply_elem(load_elem<D>(wagg, nwarps, j), m, P);
        if (j == nwarps - 1) store_state<D>(xT, 1, 0, m, P);
    }
}

// Middle level in parallel: with T = 1e7 there are ~2e4 warp aggregates, too many for one CTA to chain (651 us in round 1).
//   k_mid_scan   every warp (of many CTAs) scans 32 consecutive aggregates: group-exclusive prefixes + one aggregate per group;
//   k_filter_mid (above, one CTA) then only sees the ~600 group aggregates and emits the state entering every group;
//   k_mid_apply  one thread per warp aggregate: group state, then its group-exclusive prefix -> the state entering the warp.
template <int D>
__global__ void __launch_bounds__(kBlock)
k_mid_scan(const double* __restrict__ wagg, long long nwarps, double* __restrict__ gexcl, double* __restrict__ gagg, long long ngroups) {
    const long long j = (long long)blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    Elem<D> E = j < nwarps ? load_elem<D>(wagg, nwarps, j) : elem_identity<D>();
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        const Elem<D> O = shfl_up_elem(E, off);
        if (lane >= off) E = combine(O, E);
    }
    Elem<D> X = shfl_up_elem(E, 1);
    if (lane == 0) X = elem_identity<D>();
    if (j < nwarps) store_elem(gexcl, nwarps, j, X);
    if (lane == 31 && (j >> 5) < ngroups) store_elem(gagg, ngroups, j >> 5, E);
}
template <int D>
__global__ void __launch_bounds__(kBlock)
k_mid_apply(const double* __restrict__ gexcl, const double* __restrict__ gstate, long long nwarps, long long ngroups, double* __restrict__ wstate) {
    const long long j = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (j >= nwarps) return;
    Vec<D> m;
    Sym<D> P;
    load_state<D>(gstate, ngroups, j >> 5, m, P);
    apply_elem(load_elem<D>(gexcl, nwarps, j), m, P);
    store_state<D>(wstate, nwarps, j, m, P);
}

// Total of all warp aggregates as ONE element in the ABI's shard format (A, b, C, eta, J with full
// column-major matrices): phase 1 of the time-sharded path (include/tgp_b200.h, tgp_shard_reduce).
template <int D>
__global__ void __launch_bounds__(kMidThreads)
k_elem_total(const double* __restrict__ wagg, long long nwarps, double* __restrict__ out) {
    __shared__ Elem<D> tot[kMidThreads / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long q = (nwarps + kMidThreads - 1) / kMidThreads;
    const long long s = (long long)t * q, e = min(s + q, nwarps);
    Elem<D> E = elem_identity<D>();
    for (long long j = s; j < e; ++j) E = combine(E, load_elem<D>(wagg, nwarps, j));
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        const Elem<D> O = shfl_up_elem(E, off);
        if (lane >= off) E = combine(O, E);
    }
    if (lane == 31) tot[w] = E;
    __syncthreads();
    if (t == 0) {
        Elem<D> R = tot[0];
        for (int ww = 1; ww < kMidThreads / 32; ++ww) R = combine(R, tot[ww]);
        double* p = out;
        for (int k = 0; k < D * D; ++k) p[k] = R.A.v[k];
        p += D * D;
        for (int k = 0; k < D; ++k) p[k] = R.b.v[k];
        p += D;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i < D; ++i) p[i + D * j] = R.C(i, j);
        p += D * D;
        for (int k = 0; k < D; ++k) p[k] = R.eta.v[k];
        p += D;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i < D; ++i) p[i + D * j] = R.J(i, j);
    }
}

// =============================================================================================
// Filtering, phase 2
// =============================================================================================
struct FilterOut {
    double* lml_steps;  // positioned at processing step 0, stride s_l (±1); nullable
    long long s_l;
    double* m_f;        // nullable
    long long s_m;
    double* P_f;        // nullable; full D x D column-major per step
    long long s_P;
    double* ws_m;       // workspace copy of the filtering distributions (SoA, for the smoother); nullable
    double* partials;   // one lml partial per CTA
    unsigned long long* err_step;  // min failing processing step (init ~0ull)
};

// log(prod S) accumulated as (mantissa product, exponent sum): one log() per 32 steps.
struct LogAcc {
    double pm = 1.0, acc = 0.0;
    long long esum = 0;
    int cnt = 0;
    __device__ __forceinline__ void add(double S) {
        const int hi = __double2hiint(S), lo = __double2loint(S);
        esum += ((hi >> 20) & 0x7ff) - 1022;
        pm *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, lo);
        if (++cnt == 32) { acc += log(pm); pm = 1.0; cnt = 0; }
    }
    __device__ __forceinline__ double total() const { return acc + log(pm) + (double)esum * 0.69314718055994530942; }
};

template <int D, bool TV>
__global__ void __launch_bounds__(kBlock)
k_filter_apply(const DevModel dm, const ConstModel<D>* __restrict__ cmp, int L, long long nthreads,
               const double* __restrict__ excl, const double* __restrict__ wstate, long long nwarps, const FilterOut out) {
    const long long tid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long s = tid * L;
    const long long e = min(s + (long long)L, dm.T);
    double quad_sum = 0.0, lml_direct = 0.0;
    LogAcc la;
    if (s < e) {
        Vec<D> m;
        Sym<D> P;
        load_state<D>(wstate, nwarps, tid >> 5, m, P);
        {
            const Elem<D> X = load_elem<D>(excl, nthreads, tid);
            apply_elem(X, m, P);
        }
        Mat<D> cA; Vec<D> ca, cH; Sym<D> cQ; double ch = 0.0, cR = 0.0;
        if (!TV) { cA = cmp->A; ca = cmp->a; cQ = cmp->Q; cH = cmp->H; ch = cmp->h; cR = cmp->R; }
        for (long long n = s; n < e; ++n) {
            const double y = __ldg(dm.y + n * dm.sy);
            double S, quad;
            if (TV) {
                predict(m, P, ldg_mat<D>(dm.A + n * dm.sA), ldg_vec<D>(dm.a + n * dm.sa), ldg_sym_full<D>(dm.Q + n * dm.sQ));
                S = update_scalar(m, P, ldg_vec<D>(dm.H + n * dm.sH), __ldg(dm.h + n * dm.sh), __ldg(dm.R + n * dm.sR), y, &quad);
            } else {
                predict(m, P, cA, ca, cQ);
                S = update_scalar(m, P, cH, ch, cR, y, &quad);
            }
            if (!(S > 1e-300) || !(S < 1e300)) { atomicMin(out.err_step, (unsigned long long)n); S = 1.0; }
            if (out.lml_steps) {
                const double l = lml_from(S, quad);
                out.lml_steps[n * out.s_l] = l;
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
                for (int j = 0; j < D; ++j)
#pragma unroll
                    for (int i = 0; i < D; ++i) out.P_f[n * out.s_P + i + D * j] = P(i, j);
            }
            if (out.ws_m) store_state<D>(out.ws_m, dm.T, n, m, P);
        }
    }
    double part = out.lml_steps ? lml_direct
                                : -0.5 * ((double)(e > s ? e - s : 0) * kLog2Pi + la.total() + quad_sum);
    // CTA reduction in a fixed order (deterministic)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
    __shared__ double wsum[kBlock / 32];
    if (lane == 0) wsum[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kBlock / 32; ++i) t += wsum[i];
        out.partials[blockIdx.x] = t;
    }
}

// Sum of CTA partials (+ an optional extra term) -> *out. One CTA, fixed order.
static __global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ partials, long long n,
                                                     const double* __restrict__ extra, double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s += partials[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0] + (extra ? *extra : 0.0);
}

// Initial state: x0buf <- (m0, upper(P0)). For TGP_REVERSE the first visited step is an update
// without a preceding predict (lgssm.jl:161-165): fold it here and emit its lml / outputs.
template <int D>
__global__ void k_init_state(const double* __restrict__ m0, const double* __restrict__ P0, double* __restrict__ x0buf,
                             int pre_update, const double* H, const double* h, const double* R, const double* y,
                             double* lml_extra, double* lml_step_out, double* m_out, double* P_out,
                             double* ws_state, long long ws_stride, long long ws_idx, unsigned long long* err_step) {
    if (threadIdx.x || blockIdx.x) return;
    Vec<D> m = ldg_vec<D>(m0);
    Sym<D> P = ldg_sym_full<D>(P0);
    double l = 0.0;
    if (pre_update) {
        double quad;
        double S = update_scalar(m, P, ldg_vec<D>(H), *h, *R, *y, &quad);
        if (!(S > 1e-300) || !(S < 1e300)) { atomicMin(err_step, 0ull); S = 1.0; }
        l = lml_from(S, quad);
        if (lml_step_out) *lml_step_out = l;
        if (m_out)
            for (int i = 0; i < D; ++i) m_out[i] = m[i];
        if (P_out)
            for (int j = 0; j < D; ++j)
                for (int i = 0; i < D; ++i) P_out[i + D * j] = P(i, j);
        if (ws_state) store_state<D>(ws_state, ws_stride, ws_idx, m, P);
    }
    if (lml_extra) *lml_extra = l;
    store_state<D>(x0buf, 1, 0, m, P);
}

// =============================================================================================
// Affine (A, b, C) scans: data-free marginals (lgssm.jl:99-115) and the backward pass of the
// posterior (lgssm.jl:215-240 + :111-115). Processing step n = 0..T-1; provider builds the
// element that maps the running state across step n.
// =============================================================================================
template <int D> __device__ __forceinline__ void store_aff(double* base, long long stride, long long i, const Aff<D>& e) {
    double* p = base + i;
#pragma unroll
    for (int k = 0; k < D * D; ++k) { *p = e.A.v[k]; p += stride; }
#pragma unroll
    for (int k = 0; k < D; ++k) { *p = e.b.v[k]; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { *p = e.C.v[k]; p += stride; }
}
template <int D> __device__ __forceinline__ Aff<D> load_aff(const double* base, long long stride, long long i) {
    Aff<D> e;
    const double* p = base + i;
#pragma unroll
    for (int k = 0; k < D * D; ++k) { e.A.v[k] = *p; p += stride; }
#pragma unroll
    for (int k = 0; k < D; ++k) { e.b.v[k] = *p; p += stride; }
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) { e.C.v[k] = *p; p += stride; }
    return e;
}
template <int D> __device__ __forceinline__ Aff<D> shfl_up_aff(const Aff<D>& e, int off) {
    Aff<D> r;
#pragma unroll
    for (int k = 0; k < D * D; ++k) r.A.v[k] = __shfl_up_sync(0xffffffffu, e.A.v[k], off);
#pragma unroll
    for (int k = 0; k < D; ++k) r.b.v[k] = __shfl_up_sync(0xffffffffu, e.b.v[k], off);
#pragma unroll
    for (int k = 0; k < Sym<D>::N; ++k) r.C.v[k] = __shfl_up_sync(0xffffffffu, e.C.v[k], off);
    return r;
}

// Element providers -------------------------------------------------------------------------
// (1) transitions straight from the model arrays (marginals of a Forward / Reverse model).
template <int D>
struct ModelAffProvider {
    DevModel dm;
    __device__ __forceinline__ bool get(long long n, Aff<D>& e) const {
        e.A = ldg_mat<D>(dm.A + n * dm.sA);
        e.b = ldg_vec<D>(dm.a + n * dm.sa);
        e.C = ldg_sym_full<D>(dm.Q + n * dm.sQ);
        return true;
    }
};
// (2) reverse-time dynamics rebuilt from stored filtering distributions (SoA ws, stride T).
//     Processing step n corresponds to forward time t = T-1-n; the element maps x_t -> x_{t-1}
//     and needs the filtering distribution at t-1 (x0 when t = 0).
template <int D>
struct SmootherProvider {
    DevModel dm;              // forward-time model arrays (A at forward step t: dm.A + t*sA)
    const double* ws;         // filtered (m, P packed) SoA, stride dm.T
    const double* x0buf;      // prior x0
    unsigned long long* err_step;
    __device__ __forceinline__ bool get(long long n, Aff<D>& e) const {
        const long long t = dm.T - 1 - n;
        Vec<D> mf;
        Sym<D> Pf;
        if (t > 0) load_state<D>(ws, dm.T, t - 1, mf, Pf);
        else load_state<D>(x0buf, 1, 0, mf, Pf);
        Vec<D> mp = mf;
        Sym<D> Pp = Pf;
        const Mat<D> A = ldg_mat<D>(dm.A + t * dm.sA);
        predict(mp, Pp, A, ldg_vec<D>(dm.a + t * dm.sa), ldg_sym_full<D>(dm.Q + t * dm.sQ));
        const bool ok = invert_dynamics(mf, Pf, mp, Pp, A, e);
        if (!ok) atomicMin(err_step, (unsigned long long)t);
        return ok;
    }
};

template <int D, class Prov>
__global__ void __launch_bounds__(kBlock)
k_aff_reduce(const Prov prov, long long T, int L, long long nthreads, double* __restrict__ excl,
             double* __restrict__ wagg, long long nwarps) {
    const long long tid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long s = tid * L, e = min(s + (long long)L, T);
    Aff<D> E = aff_identity<D>();
    for (long long n = s; n < e; ++n) {
        Aff<D> el;
        prov.get(n, el);
        E = aff_combine(E, el);
    }
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        const Aff<D> O = shfl_up_aff(E, off);
        if (lane >= off) E = aff_combine(O, E);
    }
    Aff<D> X = shfl_up_aff(E, 1);
    if (lane == 0) X = aff_identity<D>();
    if (tid < nthreads) store_aff(excl, nthreads, tid, X);
    if (lane == 31) store_aff(wagg, nwarps, tid >> 5, E);
}

template <int D>
__global__ void __launch_bounds__(kMidThreads)
k_aff_mid(const double* __restrict__ wagg, long long nwarps, const double* __restrict__ x0buf,
          double* __restrict__ wstate, double* __restrict__ xT) {
    __shared__ Aff<D> tot[kMidThreads / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long q = (nwarps + kMidThreads - 1) / kMidThreads;
    const long long s = (long long)t * q, e = min(s + q, nwarps);
    Aff<D> E = aff_identity<D>();
    for (long long j = s; j < e; ++j) E = aff_combine(E, load_aff<D>(wagg, nwarps, j));
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        const Aff<D> O = shfl_up_aff(E, off);
        if (lane >= off) E = aff_combine(O, E);
    }
    if (lane == 31) tot[w] = E;
    Aff<D> X = shfl_up_aff(E, 1);
    if (lane == 0) X = aff_identity<D>();
    __syncthreads();
    Vec<D> m;
    Sym<D> P;
    load_state<D>(x0buf, 1, 0, m, P);
    for (int ww = 0; ww < w; ++ww) aff_apply(tot[ww], m, P);
    aff_apply(X, m, P);
    for (long long j = s; j < e; ++j) {
        store_state<D>(wstate, nwarps, j, m, P);
        aff_apply(load_aff<D>(wagg, nwarps, j), m, P);
        if (j == nwarps - 1 && xT) store_state<D>(xT, 1, 0, m, P);
    }
}

// Emission of the per-step marginal. emit_before: emit from the state BEFORE crossing step n
// (Reverse ordering, lgssm.jl:111-115), else after (Forward, :105-109).
struct EmitOut {
    const double *H, *h, *R;      // positioned at processing step 0
    long long sH, sh, sR;
    double *mean, *var;           // positioned at processing step 0
    long long s_o;                // ±1
};

template <int D, class Prov>
__global__ void __launch_bounds__(kBlock)
k_aff_apply(const Prov prov, long long T, int L, long long nthreads, const double* __restrict__ excl,
            const double* __restrict__ wstate, long long nwarps, const EmitOut out, int emit_before) {
    const long long tid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const long long s = tid * L, e = min(s + (long long)L, T);
    if (s >= e) return;
    Vec<D> m;
    Sym<D> P;
    load_state<D>(wstate, nwarps, tid >> 5, m, P);
    {
        const Aff<D> X = load_aff<D>(excl, nthreads, tid);
        aff_apply(X, m, P);
    }
    for (long long n = s; n < e; ++n) {
        double mu, var;
        const Vec<D> H = ldg_vec<D>(out.H + n * out.sH);
        const double h = __ldg(out.h + n * out.sh), R = __ldg(out.R + n * out.sR);
        if (emit_before) {
            emit_scalar(m, P, H, h, R, &mu, &var);
            out.mean[n * out.s_o] = mu;
            out.var[n * out.s_o] = var;
        }
        Aff<D> el;
        prov.get(n, el);
        aff_apply(el, m, P);
        if (!emit_before) {
            emit_scalar(m, P, H, h, R, &mu, &var);
            out.mean[n * out.s_o] = mu;
            out.var[n * out.s_o] = var;
        }
    }
}

// posterior(::LGSSM, y) materialised (lgssm.jl:193-200): G, g, Sigma per forward step from the stored
// filtering distributions. One thread per step; parity entry point tgp_posterior.
template <int D>
__global__ void __launch_bounds__(kBlock)
k_posterior_dynamics(const SmootherProvider<D> prov, double* __restrict__ G, double* __restrict__ g, double* __restrict__ Sig) {
    const long long t = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (t >= prov.dm.T) return;
    Aff<D> e;
    prov.get(prov.dm.T - 1 - t, e);
#pragma unroll
    for (int k = 0; k < D * D; ++k) G[t * D * D + k] = e.A.v[k];
#pragma unroll
    for (int k = 0; k < D; ++k) g[t * D + k] = e.b.v[k];
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) Sig[t * D * D + i + D * j] = e.C(i, j);
}

// unpack a packed state (m, P upper) to (m, full col-major P)
template <int D>
__global__ void k_unpack_state(const double* __restrict__ st, double* __restrict__ m, double* __restrict__ P) {
    if (threadIdx.x || blockIdx.x) return;
    Vec<D> mm;
    Sym<D> PP;
    load_state<D>(st, 1, 0, mm, PP);
    if (m)
        for (int i = 0; i < D; ++i) m[i] = mm[i];
    if (P)
        for (int j = 0; j < D; ++j)
            for (int i = 0; i < D; ++i) P[i + D * j] = PP(i, j);
}

}  // namespace tgp

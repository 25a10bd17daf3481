// tgp_fir.cuh — log-likelihood of a TIME-INVARIANT scalar-observation LGSSM in ONE launch, one pass over y
// (RegularSpacing + homoscedastic noise: BASELINE configs 1, 2, 4; replaces scan_emit + step_logpdf, scan.jl:15-28,
// lgssm.jl:147-159, with predict LGC:46-52 and posterior_and_lml LGC:247-257 as the per-step arithmetic).
//
// Past the covariance transient (host plan, tgp_fir_plan.h) the filter is the constant-coefficient recursion
//      m_t = Abar m_{t-1} + K y_t + c,    v_t = y_t - w'm_{t-1} - hh,    lml_t = -(log 2pi + log S + v_t^2 / S) / 2.
// The steady steps are cut into tiles of 1024 = 32 lanes x 32 steps; a lane keeps its 32 observations in registers
// (8 x 256-bit loads) and works on blocks of 8 steps:
//   pass A   u_b = zc + sum_j (Abar^(7-j) K) y_j ;  z <- Abar^8 z + u_b          zero-state response of the lane's run
//   scan     5 shuffle levels with Abar^(32 2^k): lane-exclusive prefix; lane 31 PUBLISHES the tile's zero-state
//            response (3 x 16-byte {value, epoch} words, no fence)
//   carry    the state entering the tile = sum_{k=1..nb} Abar^(1024 (k-1)) tot_{t-k}: a stable filter forgets its start,
//            so nb (1..3, from |Abar^1024|, decided by the plan) PREDECESSOR tiles suffice to double precision. Those
//            tiles are being processed by neighbouring warps at the same moment — there is no serial chain through the
//            series, no second pass over y, no grid barrier and no scratch besides 48 B per tile.
//   pass B   from the true block-start state m: v_j = y_j - kap_j - (w'Abar^j) m - sum_{i<j} g_{j-1-i} y_i (independent FMAs,
//            coefficients are kernel parameters = constant-bank operands), q += v_j^2 ;  m <- Abar^8 m + u_b.
// 14.4 DFMA per step instead of ~25 for the two-phase kernel (tgp_steady.cuh), the loop-carried chain is 3 DFMA per 8 steps.
// The transient (the first N0 steps, where P_t still moves) is one warp of a service CTA running the time-varying affine
// recursion with the gains K_t, 1/S_t tabulated by the plan. Time shards (multi-GPU) use the same kernel: a rank > 0 replaces the
// transient by pass A over the last nb tiles of its predecessor's shard (the "halo": 8 KB per tile, pushed over NVLink by the
// predecessor's service CTA at the START of its kernel), so shards never wait for each other's results.
// Work assignment is static and deterministic: CTAs take a virtual index in start order (atomic), warp gw owns tiles
// gw, gw + NW, ...; every reduction has a fixed order, so the result is bit-reproducible.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "tgp_ctx.cuh"
#include "tgp_fir_plan.h"

namespace tgp {

struct FirXchg {                  // time-sharded use (all null / 0 on a single GPU)
    const double* halo;           // nb * 1024 observations preceding this shard's first step, oldest first (rank > 0)
    const unsigned long long* halo_flag;   // wait for *halo_flag >= epoch before reading the halo (null: no wait)
    unsigned long long* ack_out;  // predecessor's ack word (peer memory): set to epoch once the halo is in registers
    double* push_dst;             // successor's halo buffer for this epoch (peer memory), null on the last rank
    unsigned long long* push_flag;         // successor's halo flag (peer memory)
    const unsigned long long* ack_in;      // own ack word: wait for >= epoch - ring before overwriting the ring slot
    unsigned long long ring;      // halo ring depth
    char* const* peers;           // mapped exchange buffers of all ranks: the partial log-likelihood goes to every peer
    unsigned long long lml_off, lml_flag_off;   // byte offsets of this rank's lml slot (for this epoch) / flag inside a peer buffer
    int rank, world;
};

struct FirArgs {
    const double* y;              // observations of this shard, index = time
    const double* tab;            // transient table [N0][D + 1]: K_t, 1/S_t
    const double* plane;          // [D*D][32]: Abar^(32 lane), entry-major
    double2* agg;                 // (ntiles + kFirNbMax) * D words {value, epoch bits}; slot(t) = t + kFirNbMax
    unsigned long long epoch;     // tag of this call's words
    unsigned* counters;           // [0] virtual CTA index claim, [1] finished CTAs (both left at 0 by the last CTA)
    double* partials;             // gridDim.x per-CTA sums of v^2 (+ [gridDim.x]: transient's sum of v^2 / S_t)
    double* result;               // lml of this shard (device)
    double* lml_user;             // caller's device destination, nullable
    FirXchg x;
};

__device__ __forceinline__ void fir_ld4(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ unsigned long long fir_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fir_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fir_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fir_put_word(double2* p, double v, unsigned long long tag) {
    asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v), "d"(__longlong_as_double((long long)tag)) : "memory");
}
__device__ __forceinline__ void fir_get_word(const double2* p, double& v, unsigned long long& tag) {
    double t;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(t) : "l"(p) : "memory");
    tag = (unsigned long long)__double_as_longlong(t);
}
__device__ __forceinline__ unsigned long long fir_ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int D> __device__ __forceinline__ Vec<D> fir_shfl_up(const Vec<D>& v, int off) {
    Vec<D> r;
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = __shfl_up_sync(0xffffffffu, v[i], off);
    return r;
}

// ---- staging of a full, aligned tile: coalesced 16-byte cp.async into a per-warp buffer of 32 padded rows (one row = one lane's run,
// 34 doubles apart so the lanes' 128-bit reads are conflict-free). The NEXT tile of the warp streams in while the current one is
// being computed from registers.
constexpr int kFirRow = kFirL + 2;                    // doubles between rows
constexpr int kFirBufDoubles = 32 * kFirRow;          // per warp
__device__ __forceinline__ void fir_cp16(double* smem_dst, const double* g, unsigned long long pol) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "l"(pol) : "memory");
}
__device__ __forceinline__ void fir_issue_tile(double* buf, const double* __restrict__ ys, int lane, unsigned long long pol) {
#pragma unroll
    for (int k = 0; k < kFirL / 2; ++k) {
        const int c = k * 32 + lane;                  // 16-byte chunk of the tile: consecutive lanes, consecutive addresses
        fir_cp16(buf + (c >> 4) * kFirRow + (c & 15) * 2, ys + 2 * c, pol);
    }
}
__device__ __forceinline__ void fir_read_tile(const double* buf, int lane, double (&yv)[kFirL]) {
    const double2* row = reinterpret_cast<const double2*>(buf + lane * kFirRow);
#pragma unroll
    for (int i = 0; i < kFirL / 2; ++i) {
        const double2 t = row[i];
        yv[2 * i] = t.x;
        yv[2 * i + 1] = t.y;
    }
}

// One tile from the lane's 32 observations in registers. slot: index of the tile's word group in agg. nvalid: steps of the tile that
// exist (TAIL only). pub: publish the zero-state response. full: wait for the carry and run pass B. Returns the lane's sum of v^2.
template <int D, bool TAIL>
__device__ __forceinline__ double fir_tile_compute(const FirPlan<D>& pl, const FirArgs& ar, double (&yv)[kFirL], long long slot, int nvalid,
                                                   bool pub, bool full, const double* __restrict__ splane, int lane) {
    // ---- pass A: zero-state response of the lane's run --------------------------------------------------------------
    Vec<D> u[kFirNBlk], z = vzero<D>();
    fir_pass_a<D>(pl, yv, u, z);
    // ---- warp scan: state at the end of every lane's run, the tile starting from zero -------------------------------------
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const Vec<D> zu = fir_shfl_up(z, 1 << k);
        if (lane >= (1 << k)) z = fir_scan_level<D>(pl, k, z, zu);
    }
    if (pub && lane == 31) {
#pragma unroll
        for (int i = 0; i < D; ++i) fir_put_word(ar.agg + slot * D + i, z[i], ar.epoch);
    }
    if (!full) return 0.0;
    Vec<D> m = fir_shfl_up(z, 1);
    if (lane == 0) m = vzero<D>();
    // ---- carry: state entering the tile from the nb tiles before it. First look right away (the load is in flight during the
    // data-only half of pass B), then spin if a predecessor had not published yet.
    const int npoll = pl.nb * D;
    double val = 0.0;
    unsigned long long tag = ar.epoch;
    const double2* wp = ar.agg + (slot - 1 - lane / D) * D + (lane % D);
    if (lane < npoll) fir_get_word(wp, val, tag);
    fir_pass_b1<D>(pl, yv);
    if (lane < npoll) {
        unsigned spins = 0;
        while (tag != ar.epoch) {
            if (++spins > (1u << 28)) __trap();    // a predecessor that never arrives (a lost peer rank): fail loudly
            __nanosleep(32);
            fir_get_word(wp, val, tag);
        }
    }
    __syncwarp();
    {
        Vec<D> c;
#pragma unroll
        for (int i = 0; i < D; ++i) c[i] = __shfl_sync(0xffffffffu, val, i);
        for (int k = 1; k < pl.nb; ++k) {
            Vec<D> tk;
#pragma unroll
            for (int i = 0; i < D; ++i) tk[i] = __shfl_sync(0xffffffffu, val, k * D + i);
            fir_carry_add<D>(pl, k, tk, c);
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) m[i] = fma(splane[(i * D + j) * 32 + lane], c[j], m[i]);
    }
    // ---- pass B, state half -------------------------------------------------------------------------------------------------
    return fir_pass_b2<D, TAIL>(pl, yv, u, m, TAIL ? nvalid - lane * kFirL : kFirL);
}

// Partial / unaligned / halo tiles: guarded scalar loads straight from global memory, one out-of-line copy.
template <int D>
__device__ __noinline__ double fir_tile_guarded(const FirPlan<D>& pl, const FirArgs& ar, const double* __restrict__ ys, long long slot, int nvalid,
                                                bool pub, bool full, const double* __restrict__ splane, int lane) {
    double yv[kFirL];
#pragma unroll
    for (int j = 0; j < kFirL; ++j) {
        const int e = lane * kFirL + j;
        yv[j] = e < nvalid ? __ldg(ys + e) : 0.0;
    }
    return fir_tile_compute<D, true>(pl, ar, yv, slot, nvalid, pub, full, splane, lane);
}

// The transient: steps [0, N0) with the tabulated gains, one warp. Lane l owns a run of ceil(N0 / 32) steps: it composes the run's
// affine map, a shuffle scan over (M, b) gives its start state, a second sweep forms the innovations. Publishes m_{N0-1} as the
// word group of "tile -1" (zeros for the tiles before it) and returns sum_t v_t^2 / S_t in lane 0.
template <int D>
__device__ __noinline__ double fir_head(const FirPlan<D>& pl, const FirArgs& ar, int lane) {
    const long long N0 = pl.N0;
    const long long c = (N0 + 31) / 32;
    const long long t0 = min((long long)lane * c, N0), t1 = min(t0 + c, N0);
    Mat<D> A;
    Vec<D> a, w;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        a[i] = pl.a[i];
        w[i] = pl.w[i];
#pragma unroll
        for (int j = 0; j < D; ++j) A(i, j) = pl.A[i][j];
    }
    const double hh = pl.hh;
    Mat<D> M = meye<D>();
    Vec<D> b = vzero<D>();
#pragma unroll 1
    for (long long t = t0; t < t1; ++t) {
        Vec<D> K;
#pragma unroll
        for (int i = 0; i < D; ++i) K[i] = __ldg(ar.tab + t * (D + 1) + i);
        const double r = __ldg(ar.y + t) - hh - dot(w, b);
        Vec<D> bn = matvec(A, b);
#pragma unroll
        for (int i = 0; i < D; ++i) b[i] = fma(K[i], r, bn[i] + a[i]);
        const Vec<D> s = matTvec(M, w);          // w'M, per column
        Mat<D> Mn = matmul(A, M);
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i) Mn(i, j) = fma(-K[i], s[j], Mn(i, j));
        M = Mn;
    }
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        Mat<D> Mu;
        Vec<D> bu;
#pragma unroll
        for (int i = 0; i < D * D; ++i) Mu.v[i] = __shfl_up_sync(0xffffffffu, M.v[i], off);
#pragma unroll
        for (int i = 0; i < D; ++i) bu[i] = __shfl_up_sync(0xffffffffu, b[i], off);
        if (lane >= off) {
            const Vec<D> t = matvec(M, bu);
#pragma unroll
            for (int i = 0; i < D; ++i) b[i] += t[i];
            M = matmul(M, Mu);
        }
    }
    Vec<D> m0;
#pragma unroll
    for (int i = 0; i < D; ++i) m0[i] = pl.m0[i];
    {   // lane 31's inclusive map covers [0, N0): the filtered mean after the transient enters tile 0
        const Vec<D> t = matvec(M, m0);
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < D; ++i) fir_put_word(ar.agg + (kFirNbMax - 1) * D + i, t[i] + b[i], ar.epoch);
            for (int k = 2; k <= kFirNbMax; ++k)
#pragma unroll
                for (int i = 0; i < D; ++i) fir_put_word(ar.agg + (kFirNbMax - k) * D + i, 0.0, ar.epoch);
        }
    }
    Mat<D> Me;
    Vec<D> be;
#pragma unroll
    for (int i = 0; i < D * D; ++i) Me.v[i] = __shfl_up_sync(0xffffffffu, M.v[i], 1);
#pragma unroll
    for (int i = 0; i < D; ++i) be[i] = __shfl_up_sync(0xffffffffu, b[i], 1);
    Vec<D> m = m0;
    if (lane > 0) {
        const Vec<D> t = matvec(Me, m0);
#pragma unroll
        for (int i = 0; i < D; ++i) m[i] = t[i] + be[i];
    }
    double q = 0.0;
#pragma unroll 1
    for (long long t = t0; t < t1; ++t) {
        Vec<D> K;
#pragma unroll
        for (int i = 0; i < D; ++i) K[i] = __ldg(ar.tab + t * (D + 1) + i);
        const double is = __ldg(ar.tab + t * (D + 1) + D);
        const double v = __ldg(ar.y + t) - hh - dot(w, m);
        q = fma(v * v, is, q);
        const Vec<D> mn = matvec(A, m);
#pragma unroll
        for (int i = 0; i < D; ++i) m[i] = fma(K[i], v, mn[i] + a[i]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_down_sync(0xffffffffu, q, off);
    return q;
}

template <int D>
__global__ void __launch_bounds__(kFirThreads, kFirCtasPerSm)
k_fir_logpdf(const __grid_constant__ FirPlan<D> pl, const __grid_constant__ FirArgs ar) {
    extern __shared__ __align__(16) double sbuf[];   // kFirWarps staging buffers
    __shared__ double splane[D * D * 32];
    __shared__ double sred[kFirWarps];
    __shared__ unsigned s_vb;
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const unsigned G = gridDim.x;
    if (tid == 0) s_vb = atomicAdd(ar.counters, 1u);
    for (int i = tid; i < D * D * 32; i += kFirThreads) splane[i] = __ldg(ar.plane + i);
    __syncthreads();
    const unsigned vb = s_vb;
    const double* __restrict__ ys = ar.y + pl.N0;
    const long long Ts = pl.T - pl.N0;
    double q = 0.0;
    if (vb == 0) {
        // ---- service CTA: the transient (or, on a shard with rank > 0, pass A over the halo) and the halo push -------------
        if (wp == 0) {
            if (ar.x.halo == nullptr) {
                const double qh = fir_head<D>(pl, ar, lane);
                if (lane == 0) ar.partials[G] = qh;
            } else {
                if (ar.x.halo_flag) {
                    if (lane == 0) {
                        unsigned spins = 0;
                        while (fir_ld_sys(ar.x.halo_flag) < ar.epoch) {
                            if (++spins > (1u << 28)) __trap();
                            __nanosleep(100);
                        }
                    }
                    __syncwarp();
                    __threadfence_system();
                }
                for (int k = pl.nb; k >= 1; --k)     // tile -k: its zero-state response goes to slot kFirNbMax - k
                    fir_tile_guarded<D>(pl, ar, ar.x.halo + (size_t)(pl.nb - k) * kFirTile, kFirNbMax - k, kFirTile, true, false, splane, lane);
                if (ar.x.ack_out && lane == 0) {
                    __threadfence_system();
                    *reinterpret_cast<volatile unsigned long long*>(ar.x.ack_out) = ar.epoch;
                }
                if (lane == 0) ar.partials[G] = 0.0;
            }
        } else if (wp == 1 && ar.x.push_dst) {
            // the last nb tiles of this shard -> the successor's halo ring slot, then its flag
            if (lane == 0 && ar.x.ack_in && ar.epoch > ar.x.ring) {
                unsigned spins = 0;
                while (fir_ld_sys(ar.x.ack_in) + ar.x.ring < ar.epoch) {
                    if (++spins > (1u << 28)) __trap();
                    __nanosleep(100);
                }
            }
            __syncwarp();
            const long long n = (long long)pl.nb * kFirTile;
            const double* src = ar.y + pl.T - n;
            for (long long e = lane; e < n; e += 32) ar.x.push_dst[e] = __ldg(src + e);
            __threadfence_system();
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(ar.x.push_flag) = ar.epoch;
        }
    } else {
        const long long NW = (long long)(G - 1) * kFirWarps;
        const long long gw = (long long)(vb - 1) * kFirWarps + wp;
        const long long ntiles = pl.ntiles;
        const bool deferred = gw < pl.nb && gw < ntiles;   // this warp's first tile waits for the transient / halo: publish now, finish last
        // items of this warp: [its deferred tile, pass A only] tiles first, first + NW, ... [the deferred tile, pass B]
        const long long first = deferred ? gw + NW : gw;
        const long long n_main = ntiles > first ? (ntiles - first + NW - 1) / NW : 0;
        const long long it0 = deferred ? -1 : 0, it1 = n_main + (deferred ? 1 : 0);
        double* buf = sbuf + wp * kFirBufDoubles;
        const unsigned long long pol = fir_policy_evict_first();
        auto tile_of = [&](long long it) { return (it >= 0 && it < n_main) ? first + it * NW : gw; };
        auto staged = [&](long long it) { return it < it1 && pl.aligned && Ts - tile_of(it) * kFirTile >= kFirTile; };
        if (staged(it0)) fir_issue_tile(buf, ys + tile_of(it0) * kFirTile, lane, pol);
        fir_cp_commit();
        for (long long it = it0; it < it1; ++it) {
            const long long t = tile_of(it);
            const bool pub = it < n_main, full = it >= 0;
            const long long s0 = t * kFirTile;
            if (staged(it)) {
                double yv[kFirL];
                fir_cp_wait_all();
                __syncwarp();
                fir_read_tile(buf, lane, yv);
                __syncwarp();
                if (staged(it + 1)) fir_issue_tile(buf, ys + tile_of(it + 1) * kFirTile, lane, pol);
                fir_cp_commit();
                q += fir_tile_compute<D, false>(pl, ar, yv, t + kFirNbMax, kFirTile, pub, full, splane, lane);
            } else {
                if (staged(it + 1)) fir_issue_tile(buf, ys + tile_of(it + 1) * kFirTile, lane, pol);
                fir_cp_commit();
                q += fir_tile_guarded<D>(pl, ar, ys + s0, t + kFirNbMax, (int)min(Ts - s0, (long long)kFirTile), pub, full, splane, lane);
            }
        }
        fir_cp_wait_all();
    }
    // ---- fixed-order reductions; the last CTA to finish forms the log-likelihood ------------------------------------------
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_down_sync(0xffffffffu, q, off);
    if (lane == 0) sred[wp] = q;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kFirWarps; ++i) t += sred[i];
        __stcg(ar.partials + vb, vb == 0 ? 0.0 : t);
        __threadfence();
        s_last = atomicAdd(ar.counters + 1, 1u) == G - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned i = tid; i < G; i += kFirThreads) s += __ldcg(ar.partials + i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    __syncthreads();
    if (lane == 0) sred[wp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kFirWarps; ++i) t += sred[i];
        const double lml = pl.c0 - 0.5 * (__ldcg(ar.partials + G) + pl.invS * t);
        *ar.result = lml;
        if (ar.lml_user) *ar.lml_user = lml;
        ar.counters[0] = 0u;
        ar.counters[1] = 0u;
        if (ar.x.peers) {      // partial log-likelihood of this shard -> every rank's buffer, then the flags
            for (int p = 0; p < ar.x.world; ++p) *reinterpret_cast<volatile double*>(ar.x.peers[p] + ar.x.lml_off) = lml;
            __threadfence_system();
            for (int p = 0; p < ar.x.world; ++p) *reinterpret_cast<volatile unsigned long long*>(ar.x.peers[p] + ar.x.lml_flag_off) = ar.epoch;
        }
    }
}

// Model arrays of a time-invariant descriptor -> host doubles (a tiny synchronous copy if the caller keeps them on the device).
inline int fir_fetch(tgp_ctx* h, const double* p, size_t n, double* dst) {
    if (is_device_ptr(p)) {
        TGP_CUDA(h, cudaMemcpy(dst, p, n * sizeof(double), cudaMemcpyDeviceToHost));
        h->d2h += (int64_t)(n * sizeof(double));
    } else {
        memcpy(dst, p, n * sizeof(double));
    }
    return TGP_OK;
}

template <class T>
inline int fir_grow(tgp_ctx* h, T** p, size_t* cap, size_t bytes, bool zero) {
    if (*p && *cap >= bytes) return TGP_OK;
    if (*p) { TGP_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFree(*p); *p = nullptr; *cap = 0; }
    const size_t want = bytes + (bytes >> 2) + 4096;
    TGP_CUDA(h, cudaMalloc((void**)p, want));
    if (zero) TGP_CUDA(h, cudaMemsetAsync(*p, 0, want, h->stream));
    *cap = want;
    return TGP_OK;
}

// logpdf of a Forward, time-invariant, scalar-observation model (or of one time shard of it: rank / x) in one launch.
// *handled = false: the plan says the path does not apply (no covariance fixed point within the budget, a filter that forgets
// too slowly, a short series) and nothing was enqueued. lml_out: host pointer (synchronous) or device pointer (enqueued only).
template <int D>
int logpdf_fir(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* lml_out, bool* handled, bool first_shard = true,
               const FirXchg* xc = nullptr) {
    *handled = false;
    tgp_fir_state& st = h->fir;
    const int64_t T = m->T;
    if (T < 4 * kFirTile) return TGP_OK;
    const double* dy;
    TGP_TRY(stage_in(h, y, (size_t)T, &dy));
    // ---- plan: cached on the bytes it is a function of ---------------------------------------------------------------
    struct Key { int dim, first; long long T; unsigned long long align; double tol; double v[3 * D * D + 3 * D + 2]; } key;
    memset(&key, 0, sizeof key);
    key.dim = D; key.first = first_shard ? 1 : 0; key.T = T; key.align = reinterpret_cast<unsigned long long>(dy) & 31ull; key.tol = h->ss_tol;
    double* v = key.v;
    TGP_TRY(fir_fetch(h, m->A, D * D, v));
    TGP_TRY(fir_fetch(h, m->a, D, v + D * D));
    TGP_TRY(fir_fetch(h, m->Q, D * D, v + D * D + D));
    TGP_TRY(fir_fetch(h, m->H, D, v + 2 * D * D + D));
    TGP_TRY(fir_fetch(h, m->h, 1, v + 2 * D * D + 2 * D));
    TGP_TRY(fir_fetch(h, m->R, 1, v + 2 * D * D + 2 * D + 1));
    TGP_TRY(fir_fetch(h, m->m0, D, v + 2 * D * D + 2 * D + 2));
    TGP_TRY(fir_fetch(h, m->P0, D * D, v + 2 * D * D + 3 * D + 2));
    if (st.key.size() != sizeof key || memcmp(st.key.data(), &key, sizeof key) != 0) {
        FirHostPlan<D> hp;
        fir_build_plan<D>(v, v + D * D, v + D * D + D, v + 2 * D * D + D, v[2 * D * D + 2 * D], v[2 * D * D + 2 * D + 1], v + 2 * D * D + 2 * D + 2,
                          v + 2 * D * D + 3 * D + 2, T, h->ss_tol, first_shard, reinterpret_cast<unsigned long long>(dy), &hp);
        st.status = hp.status;
        st.bad_step = hp.bad_step;
        st.plan.assign(reinterpret_cast<unsigned char*>(&hp.dev), reinterpret_cast<unsigned char*>(&hp.dev) + sizeof hp.dev);
        if (hp.status == 0) {
            st.upload.swap(hp.upload);
            TGP_TRY(fir_grow(h, &st.dev, &st.dev_cap, st.upload.size() * sizeof(double), false));
            TGP_CUDA(h, cudaMemcpyAsync(st.dev, st.upload.data(), st.upload.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            h->h2d += (int64_t)(st.upload.size() * sizeof(double));
        }
        st.key.assign(reinterpret_cast<unsigned char*>(&key), reinterpret_cast<unsigned char*>(&key) + sizeof key);
    }
    if (st.status == 2) return fail(h, TGP_ENOTPD, "covariance not positive definite at time index %lld (0-based)", st.bad_step);
    if (st.status != 0) return TGP_OK;
    FirPlan<D> pl;
    memcpy(&pl, st.plan.data(), sizeof pl);
    // ---- workspace ---------------------------------------------------------------------------------------------------
    const long long ntiles = pl.ntiles;
    const long long want_ctas = (ntiles + kFirWarps - 1) / kFirWarps;
    const unsigned G = 1u + (unsigned)std::min<long long>(want_ctas, (long long)h->sm_count * kFirCtasPerSm - 1);
    size_t agg_cap = st.agg_cap;
    double2* agg = (double2*)st.agg;
    TGP_TRY(fir_grow(h, &agg, &agg_cap, (size_t)(ntiles + kFirNbMax) * D * sizeof(double2), true));
    st.agg = agg; st.agg_cap = agg_cap;
    TGP_TRY(fir_grow(h, &st.partials, &st.partials_cap, (size_t)(G + 1) * sizeof(double), false));
    if (!st.counters) {
        TGP_CUDA(h, cudaMalloc((void**)&st.counters, 2 * sizeof(unsigned)));
        TGP_CUDA(h, cudaMemsetAsync(st.counters, 0, 2 * sizeof(unsigned), h->stream));
        TGP_CUDA(h, cudaMalloc((void**)&st.result, 4 * sizeof(double)));
    }
    FirArgs ar{};
    ar.y = dy;
    ar.tab = st.dev;
    ar.plane = st.dev + (size_t)pl.N0 * (D + 1);
    ar.agg = agg;
    ar.epoch = ++st.epoch;
    ar.counters = st.counters;
    ar.partials = st.partials;
    ar.result = st.result;
    ar.lml_user = (lml_out && is_device_ptr(lml_out)) ? lml_out : nullptr;
    if (xc) ar.x = *xc;
    TGP_K(h, "k_fir_logpdf");
    constexpr size_t smem = (size_t)kFirWarps * kFirBufDoubles * sizeof(double);
    static bool attr_set[64] = {false};
    if (!attr_set[h->device & 63]) {   // ask for the shared-memory carve-out that lets kFirCtasPerSm CTAs live on one SM
        TGP_CUDA(h, cudaFuncSetAttribute(k_fir_logpdf<D>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        attr_set[h->device & 63] = true;
    }
    k_fir_logpdf<D><<<G, kFirThreads, smem, h->stream>>>(pl, ar);
    TGP_LAUNCH_CHECK(h);
    *handled = true;
    if (lml_out && !ar.lml_user) {
        TGP_CUDA(h, cudaMemcpyAsync(h->pinned + 8, st.result, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        h->d2h += 8;
        *lml_out = h->pinned[8];
    }
    return TGP_OK;
}

}  // namespace tgp

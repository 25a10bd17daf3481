// tgp_fir.cuh — log-likelihood of a TIME-INVARIANT scalar-observation LGSSM in ONE launch, one pass over y
// (RegularSpacing + homoscedastic noise: BASELINE configs 1, 2, 4; replaces scan_emit + step_logpdf, scan.jl:15-28,
// lgssm.jl:147-159, with predict LGC:46-52 and posterior_and_lml LGC:247-257 as the per-step arithmetic).
//
// Past the covariance transient (host plan, tgp_fir_plan.h) the filter is the constant-coefficient recursion
//      m_t = Abar m_{t-1} + K y_t + c,    v_t = y_t - w'm_{t-1} - hh,    lml_t = -(log 2pi + log S + v_t^2 / S) / 2.
// The steady steps are cut into tiles of 1024 = 32 lanes x 32 steps. Each CTA (two 8-warp CTAs per SM) owns a CONTIGUOUS chunk of
// tiles, its warps take them round-robin. A lane keeps its 32 observations in registers and works on blocks of 8 steps:
//   pass A   u_b = zc + sum_j (Abar^(7-j) K) y_j ;  z <- Abar^8 z + u_b          zero-state response of the lane's run
//   scan     5 shuffle levels with Abar^(32 2^k): lane-exclusive prefix; lane 31 PUBLISHES the tile's zero-state response in a
//            shared-memory ring
//   pass B1  y_j <- y_j - kap_j - sum_{i<j} g_{j-1-i} y_i  (needs no state: runs while the neighbours publish)
//   carry    the state entering the tile = sum_{k=1..nb} Abar^(1024 (k-1)) tot_{t-k}: a stable filter forgets its start, so the nb
//            (1..3, from |Abar^1024|, decided by the plan) PREDECESSOR tiles suffice to double precision. They are being processed
//            by the neighbouring warps of the same CTA at the same moment: no serial chain through the series, no second pass over
//            y, no grid barrier, no global scratch. The first tiles of a chunk take theirs from pass A over the nb tiles BEFORE the
//            chunk (re-read from HBM: 1-3 % extra traffic), so CTAs never talk to each other.
//   pass B2  v_j = (B1) - (w'Abar^j) m from the true block-start state, q += v_j^2 ;  m <- Abar^8 m + u_b.
// 14.4 DFMA per step (two-phase kernel of tgp_steady.cuh: ~25) with every coefficient a constant-bank operand (the plan is a kernel
// parameter); the loop-carried chain is 3 DFMA per 8 steps. y streams through per-warp cp.async buffers (coalesced 16-byte copies
// into padded rows), one tile ahead of the arithmetic.
// The transient (the first N0 steps, where P_t still moves) is run by CTA 0 as a block-wide scan over the affine maps
// x -> A x + a + K_t (y_t - hh - w'x) with the gains K_t, 1/S_t tabulated by the plan. Time shards (multi-GPU): a rank > 0 has no
// transient; its CTA 0 takes the nb tiles before the shard (the "halo", 8 KB per tile) from a buffer its predecessor's kernel
// fills over NVLink at the START of its run, so shards never wait for each other's results.
// Work assignment is static and every reduction has a fixed order: the result is bit-reproducible.
#pragma once
#include <cuda_runtime.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "tgp_ctx.cuh"
#include "tgp_fir_plan.h"

namespace tgp {

// CTA shape: TGP_FIR_WARPS warps, kFirCtasPerSM CTAs per SM (16 warps per SM either way: the register file is full at 128 registers).
// Two 8-warp CTAs per SM instead of one 16-warp CTA let the NEXT call's CTA start (prologue, first loads) in the half of an SM that the
// current call has already vacated while the other half is still streaming.
#ifndef TGP_FIR_WARPS
#define TGP_FIR_WARPS 8
#endif
constexpr int kFirWarps = TGP_FIR_WARPS;
constexpr int kFirThreads = kFirWarps * 32;
constexpr int kFirCtasPerSM = 16 / kFirWarps;
constexpr int kFirHeadLess = 16;                     // tiles CTA 0 is spared for running the transient (twice that when it waits for a halo)
constexpr int kFirPushWarps = 4;                     // warps of the last CTA that copy the halo to the successor rank
constexpr int kFirRing = 64;                          // tile words kept in shared memory (>= 2 rounds of a CTA's warps + look-back)
constexpr int kFirRow = kFirL + 2;                    // doubles between the rows of a staging buffer
constexpr int kFirBufDoubles = 32 * kFirRow;          // staging buffer of one warp

struct FirXchg {                  // time-sharded use (all null / 0 on a single GPU)
    const double* halo;           // nb * 1024 observations preceding this shard's first step, oldest first (rank > 0)
    const unsigned long long* halo_flag;   // wait for *halo_flag >= epoch before reading the halo (null: no wait)
    unsigned long long* ack_out;  // predecessor's ack word (peer memory): set to epoch once the halo has been consumed
    double* push_dst;             // successor's halo buffer for this epoch (peer memory), null on the last rank
    unsigned long long* push_flag;         // successor's halo flag (peer memory)
    const unsigned long long* ack_in;      // own ack word: wait for >= epoch - ring before overwriting the ring slot
    unsigned long long ring;      // halo ring depth
    int local_halo;               // rank > 0 with overlapped shards: the nb tiles before y[0] are in local memory at y - nb * 1024
    char* const* peers;           // mapped exchange buffers of all ranks: the partial log-likelihood goes to every peer
    unsigned long long lml_off;   // byte offset of this rank's {lml, epoch} word (ring slot of this epoch) inside a peer buffer
    unsigned long long epoch;     // exchange epoch of this call (the same on every rank)
    int rank, world;
};

struct FirArgs {
    const double* y;              // observations of this shard, index = time
    const double* tab;            // transient table [N0][D + 1]: K_t, 1/S_t
    const double* plane;          // [D*D][32]: Abar^(32 lane), entry-major
    unsigned long long epoch;     // call counter of the handle (exchange flags)
    unsigned* counters;           // [0] finished CTAs (left at 0 by the last CTA)
    double* partials;             // gridDim.x per-CTA sums of v^2, [gridDim.x]: the transient's sum of v^2 / S_t
    double* result;               // lml of this shard (device)
    double* lml_user;             // caller's device destination, nullable
    unsigned stagger_ns;
    int head_less;                // tiles CTA 0 is spared (0: kFirHeadLess); TGP_FIR_HEADLESS, a tuning knob          // initial delay between the four warp groups of a CTA (0: none)
    unsigned long long* trace;    // optional (TGP_FIR_TRACE): 8 globaltimer stamps per CTA, see fir_trace()
    int early_trigger;            // let the next call's CTAs in as this call's CTAs leave (only when a call fills every SM: then at most
                                  // two calls are ever in flight, which is what the parity-indexed workspace allows)
    FirXchg x;
};

template <int D>
struct FirSmem {
    static constexpr int o_plane = kFirWarps * kFirBufDoubles;
    static constexpr int o_ring = o_plane + D * D * 32 + ((D * D * 32) & 1);   // 16-byte words {value, tag}
    static constexpr int o_red = o_ring + 2 * (kFirRing + 2 * kFirNbMax) * D;
    static constexpr int o_scan = o_red + kFirWarps + 2;                 // transient: per-warp affine maps (D*D + D each)
    static constexpr size_t bytes = (size_t)(o_scan + kFirWarps * (D * D + D)) * sizeof(double);
};

__device__ __forceinline__ unsigned long long fir_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fir_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fir_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// Waits on another warp / another GPU give up (trap: the caller sees a CUDA error instead of a hung device) only after kFirGiveUpNs.
constexpr unsigned long long kFirGiveUpNs = 60ull * 1000000000ull;
__device__ __forceinline__ unsigned long long fir_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fir_spin_check(unsigned& spins, unsigned long long& t0) {
    if ((++spins & 4095u) == 0u) {
        const unsigned long long now = fir_now_ns();
        if (t0 == 0ull) t0 = now;
        else if (now - t0 > kFirGiveUpNs) __trap();
    }
}
__device__ __forceinline__ unsigned long long fir_ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fir_cp16(double* smem_dst, const double* g, unsigned long long pol) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "l"(pol) : "memory");
}
// Staging of a full, aligned tile: coalesced 16-byte cp.async into 32 padded rows (one row = one lane's run; 34 doubles apart, so
// the lanes' 128-bit reads are conflict-free).
__device__ __forceinline__ void fir_issue_tile(double* buf, const double* __restrict__ ys, int lane, unsigned long long pol) {
    // 16-byte chunk c = 32 k + lane (consecutive lanes, consecutive addresses) -> row c >> 4, column chunk c & 15
    double* dst = buf + (lane >> 4) * kFirRow + (lane & 15) * 2;
    const double* src = ys + 2 * lane;
#pragma unroll
    for (int k = 0; k < kFirL / 2; ++k) fir_cp16(dst + k * 2 * kFirRow, src + 64 * k, pol);
}
// The same for the PARTIAL last tile of a series: chunks beyond `nvalid` observations are zero-filled (cp.async src-size), nothing is
// read past the end of y.
static __device__ __noinline__ void fir_issue_tile_zfill(double* buf, const double* __restrict__ ys, int nvalid, int lane, unsigned long long pol) {
    double* dst = buf + (lane >> 4) * kFirRow + (lane & 15) * 2;
#pragma unroll 1
    for (int k = 0; k < kFirL / 2; ++k) {
        const int e = 64 * k + 2 * lane;                                   // first observation of the chunk
        const int nbytes = max(0, min(2, nvalid - e)) * 8;
        const double* src = ys + (nbytes ? e : 0);                          // a valid address even when nothing is read
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + k * 2 * kFirRow);
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(sa), "l"(src), "r"(nbytes), "l"(pol) : "memory");
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

template <int D> __device__ __forceinline__ Vec<D> fir_shfl_up(const Vec<D>& v, int off) {
    Vec<D> r;
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = __shfl_up_sync(0xffffffffu, v[i], off);
    return r;
}

// The shared-memory ring of tile words: entry r (= tile index relative to the chunk + kFirNbMax) holds the zero-state response of
// that tile (or, for r < kFirNbMax, what precedes the chunk); tag[r % kFirRing] == r once it is there.
// Entries r < 2 kFirNbMax (what precedes the chunk, and the chunk's first tiles: on a shard with rank > 0 their pass B runs last)
// have fixed slots; the other tiles of the chunk share the ring. An entry is D 16-byte words
// {value, tag = r}: one 128-bit shared-memory store per word, so a reader that sees the tag sees the value — no fence (a fence here
// would also wait for the warp's cp.async copies in flight).
__device__ __forceinline__ int fir_slot(int r) { return r < 2 * kFirNbMax ? r : 2 * kFirNbMax + ((r - 2 * kFirNbMax) % kFirRing); }
__device__ __forceinline__ void fir_st_word(double2* p, double v, int tag) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"(sa), "d"(v), "d"(__longlong_as_double((long long)tag)) : "memory");
}
__device__ __forceinline__ void fir_ld_word(const double2* p, double& v, int& tag) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
    double t;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(t) : "r"(sa) : "memory");
    tag = (int)__double_as_longlong(t);
}
template <int D>
__device__ __forceinline__ void fir_publish(double2* sring, int r, const Vec<D>& z) {
    double2* p = sring + fir_slot(r) * D;
#pragma unroll
    for (int i = 0; i < D; ++i) fir_st_word(p + i, z[i], r);
}

// One tile from the lane's 32 observations in registers. r: ring entry of the tile. nvalid: steps of the tile that exist (TAIL only).
// pub: publish the zero-state response. full: take the carry and run pass B. next: observations of this warp's next tile (or null):
// its 16 cp.async per lane are spread over pass A instead of being issued back to back. Returns the lane's sum of v^2.
template <int D, bool TAIL, bool ZM = false>
__device__ __forceinline__ double fir_tile_compute(const FirPlan<D>& pl, double (&yv)[kFirL], int r, int nvalid, bool pub, bool full,
                                                   const double* __restrict__ splane, double2* sring, int lane,
                                                   double* buf = nullptr, const double* __restrict__ next = nullptr, unsigned long long pol = 0) {
    // ---- pass A: zero-state response of the lane's run (+ the next tile's copies) -------------------------------------------
    Vec<D> u[kFirNBlk], z = vzero<D>();
#pragma unroll
    for (int b = 0; b < kFirNBlk; ++b) {
        fir_pass_a_block<D, ZM>(pl, b, yv, u[b], z);
        if (!TAIL && next) {      // chunk c = 32 k + lane: row 2 k + (lane >> 4), column chunk lane & 15 -> one base + constants
            double* dst = buf + (lane >> 4) * kFirRow + (lane & 15) * 2;
            const double* src = next + 2 * lane;
#pragma unroll
            for (int k = b * (kFirL / 2 / kFirNBlk); k < (b + 1) * (kFirL / 2 / kFirNBlk); ++k) fir_cp16(dst + k * 2 * kFirRow, src + 64 * k, pol);
        }
    }
    if (!TAIL) fir_cp_commit();
    // ---- warp scan (state at the end of every lane's run, the tile starting from zero), with the data-only half of pass B
    // (y_j <- y_j - kap_j - sum_{i<j} g y_i) filling the shuffle latencies --------------------------------------------------------
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        Vec<D> zu = fir_shfl_up(z, 1 << k);
        if (full && k < kFirNBlk) fir_pass_b1_block<D, ZM>(pl, k, yv);
#pragma unroll
        for (int i = 0; i < D; ++i) zu[i] = lane >= (1 << k) ? zu[i] : 0.0;      // branch-free: lanes without a partner add zero
        z = fir_scan_level<D>(pl, k, z, zu);
    }
    if (pub && lane == 31) fir_publish<D>(sring, r, z);
    if (!full) return 0.0;
    Vec<D> m = fir_shfl_up(z, 1);
    if (lane == 0) m = vzero<D>();
    // ---- carry: state entering the tile from the nb entries before it (lane k D + i fetches component i of entry r - 1 - k) ---------
    double val = 0.0;
    {
        // every lane takes part in the loop (lanes >= nb D are ready at once) and leaves it on a warp-uniform vote: the warp never
        // diverges here, so the shuffles below stay on the fast (converged) path
        const bool poller = lane < pl.nb * D;
        const int want = r - 1 - (poller ? lane / D : 0);
        const double2* wptr = sring + fir_slot(want) * D + (poller ? lane % D : 0);
        unsigned spins = 0;
        unsigned long long t0 = 0ull;
        for (;;) {
            int tag = want;
            if (poller) fir_ld_word(wptr, val, tag);
            if (__all_sync(0xffffffffu, tag == want)) break;
            fir_spin_check(spins, t0);                 // cannot end short of a lost peer rank (halo): fail loudly, do not hang
            __nanosleep(20);
        }
    }
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

// Unaligned series: guarded scalar loads straight from global memory, one out-of-line copy.
template <int D>
__device__ __noinline__ double fir_tile_guarded(const FirPlan<D>& pl, const double* __restrict__ ys, int r, int nvalid, bool pub, bool full,
                                                const double* __restrict__ splane, double2* sring, int lane) {
    double yv[kFirL];
#pragma unroll
    for (int j = 0; j < kFirL; ++j) {
        const int e = lane * kFirL + j;
        yv[j] = e < nvalid ? __ldcg(ys + e) : 0.0;      // L2 (the halo ring is rewritten by a PEER between calls: never through L1)
    }
    return fir_tile_compute<D, true>(pl, yv, r, nvalid, pub, full, splane, sring, lane);
}

// The transient: steps [0, N0) with the tabulated gains, the whole CTA. Thread i owns steps [i c, (i + 1) c), c = ceil(N0 / kFirThreads): it
// composes their affine map (M, b); shuffle scan inside the warps, the warp totals through shared memory; a second sweep from the
// thread's start state forms the innovations. Publishes the filtered mean after step N0 - 1 as ring entry kFirNbMax - 1 (zeros before
// it) and returns the CTA's sum of v_t^2 / S_t (valid in thread 0).
template <int D>
__device__ __noinline__ double fir_head(const FirPlan<D>& pl, const FirArgs& ar, double* sscan, double* sred, double2* sring) {
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const long long N0 = pl.N0;
    const long long c = (N0 + kFirThreads - 1) / kFirThreads;
    const long long t0 = min((long long)tid * c, N0), t1 = min(t0 + c, N0);
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
    if (lane == 31) {       // the warp's map
        double* e = sscan + wp * (D * D + D);
#pragma unroll
        for (int i = 0; i < D * D; ++i) e[i] = M.v[i];
#pragma unroll
        for (int i = 0; i < D; ++i) e[D * D + i] = b[i];
    }
    __syncthreads();
    Vec<D> m;                // state entering the warp: the maps of the warps before it applied to m0
#pragma unroll
    for (int i = 0; i < D; ++i) m[i] = pl.m0[i];
    for (int k = 0; k < wp; ++k) {
        const double* e = sscan + k * (D * D + D);
        Vec<D> mn;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double sacc = e[D * D + i];
#pragma unroll
            for (int j = 0; j < D; ++j) sacc = fma(e[i + D * j], m[j], sacc);
            mn[i] = sacc;
        }
        m = mn;
    }
    {   // exclusive map inside the warp
        Mat<D> Me;
        Vec<D> be;
#pragma unroll
        for (int i = 0; i < D * D; ++i) Me.v[i] = __shfl_up_sync(0xffffffffu, M.v[i], 1);
#pragma unroll
        for (int i = 0; i < D; ++i) be[i] = __shfl_up_sync(0xffffffffu, b[i], 1);
        if (tid == kFirThreads - 1) {   // the last thread's inclusive map closes the transient: the mean entering tile 0
            const Vec<D> t = matvec(M, m);
            Vec<D> mf, zero = vzero<D>();
#pragma unroll
            for (int i = 0; i < D; ++i) mf[i] = t[i] + b[i];
            for (int k = 2; k <= kFirNbMax; ++k) fir_publish<D>(sring, kFirNbMax - k, zero);
            fir_publish<D>(sring, kFirNbMax - 1, mf);
        }
        if (lane > 0) {
            const Vec<D> t = matvec(Me, m);
#pragma unroll
            for (int i = 0; i < D; ++i) m[i] = t[i] + be[i];
        }
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
    if (lane == 0) sred[wp] = q;
    __syncthreads();
    double tot = 0.0;
    if (tid == 0)
        for (int i = 0; i < kFirWarps; ++i) tot += sred[i];
    __syncthreads();
    return tot;
}

// Debug timeline (TGP_FIR_TRACE=1): stamp `slot` of this CTA with the global nanosecond timer.
__device__ __forceinline__ void fir_trace(const FirArgs& ar, int slot) {
    if (ar.trace) ar.trace[(size_t)blockIdx.x * 8 + slot] = fir_now_ns();
}

template <int D>
__global__ void __launch_bounds__(kFirThreads, kFirCtasPerSM)
k_fir_logpdf(const __grid_constant__ FirPlan<D> pl, const __grid_constant__ FirArgs ar) {
    using SM = FirSmem<D>;
    extern __shared__ __align__(16) double smem[];
    double* splane = smem + SM::o_plane;
    double2* sring = reinterpret_cast<double2*>(smem + SM::o_ring);
    double* sred = smem + SM::o_red;
    double* sscan = smem + SM::o_scan;
    __shared__ int s_last, s_push_cnt;
    int* s_push = &s_push_cnt;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    if (tid == 0) { s_push_cnt = 0; fir_trace(ar, 0); }
    const long long G = gridDim.x, b = blockIdx.x;
    // Programmatic dependent launch: the next call's CTAs may take an SM as soon as this call's CTA leaves it (the calls share
    // nothing: counters / partials / result alternate by call parity).
    if (ar.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const double* __restrict__ ys = ar.y + pl.N0;
    const long long Ts = pl.T - pl.N0;
    const long long ntiles = pl.ntiles;
    // This CTA's tiles [c0, c1): equal chunks, except that CTA 0 (which also runs the transient) takes one round less.
    long long c0, c1;
    {
        const long long base = ntiles / G;
        // CTA 0 also runs the transient (rank 0: one round less) or, last, the halo and its first tiles' pass B (rank > 0: two less)
        const long long less = ar.x.local_halo ? 0 : (ar.x.halo ? 2 : 1) * (ar.head_less ? ar.head_less : kFirHeadLess);   // overlapped shard: CTA 0 is like any other
        const long long t0 = (G > 1 && base >= less + 2 * kFirWarps) ? base - less : base;   // tiles of CTA 0
        const long long rest = ntiles - t0;
        c0 = b == 0 ? 0 : t0 + (G > 1 ? rest * (b - 1) / (G - 1) : 0);
        c1 = b == 0 ? (G > 1 ? t0 : ntiles) : t0 + rest * b / (G - 1);
    }
    const bool exch_halo = b == 0 && ar.x.halo != nullptr;              // shard with rank > 0: what precedes tile 0 arrives over NVLink
    // On a shard with rank > 0 the first nb tiles of CTA 0 need the halo: their warps publish pass A now and run pass B last.
    const bool deferred = exch_halo && wp < pl.nb && c0 + wp < c1;
    const long long first = c0 + wp + (deferred ? kFirWarps : 0);       // this warp: tiles first, first + 16, ... < c1
    int n_main = c1 > first ? (int)((c1 - first + kFirWarps - 1) / kFirWarps) : 0;
    // all of them are full tiles except, possibly, the very last tile of the series
    // all of them are full tiles except, possibly, the very last tile of the series (staged too, zero-filled beyond its end)
    const int n_fast = pl.aligned ? n_main : 0;
    double* buf = smem + wp * kFirBufDoubles;
    const unsigned long long pol = fir_policy_evict_first();
    // ---- this warp's staged items, in order: [A] its deferred tile, pass A only; [B] its tiles first, first + 16, ...;
    // [C] (warp 0 of CTA 0 on a shard with rank > 0) the nb halo tiles, pass A only; [D] the deferred tile, pass B. ------------------
    const int nA = (deferred && pl.aligned) ? 1 : 0, nC = (exch_halo && wp == 0 && pl.aligned) ? pl.nb : 0, nD = nA;
    const int n_items = nA + n_fast + nC + nD;
    struct Item { const double* src; int r; int nvalid; bool pub, full; };
    auto item = [&](int it) -> Item {
        if (it < nA) return Item{ys + (c0 + wp) * kFirTile, wp + kFirNbMax, kFirTile, true, false};
        it -= nA;
        if (it < n_fast) {
            const long long t = first + (long long)it * kFirWarps;
            return Item{ys + t * kFirTile, (int)(t - c0) + kFirNbMax, (int)min((long long)kFirTile, Ts - t * kFirTile), true, true};
        }
        it -= n_fast;
        if (it < nC) return Item{ar.x.halo + (size_t)it * kFirTile, kFirNbMax - (pl.nb - it), kFirTile, true, false};      // tile -(nb - it)
        return Item{ys + (c0 + wp) * kFirTile, wp + kFirNbMax, kFirTile, false, true};
    };
    auto issue = [&](const Item& x) {
        if (x.nvalid == kFirTile) fir_issue_tile(buf, x.src, lane, pol);
        else fir_issue_tile_zfill(buf, x.src, x.nvalid, lane, pol);
    };
    auto wait_halo = [&]() {      // the predecessor's push of THIS epoch's ring slot
        if (ar.x.halo_flag) {
            if (lane == 0) {
                unsigned spins = 0;
                unsigned long long t0 = 0ull;
                while (fir_ld_sys(ar.x.halo_flag) != ar.epoch) {
                    fir_spin_check(spins, t0);
                    __nanosleep(100);
                }
            }
            __syncwarp();
            __threadfence_system();
        }
    };
    // CTAs b > 0: the last nb warps (they have a tile less than warps 0, 1 when the chunk is not a multiple of 16) first run pass A
    // over the nb tiles BEFORE the chunk (re-read from HBM) — staged like any other tile, and first in the queue.
    const int halo_k = ((b > 0 || ar.x.local_halo) && wp >= kFirWarps - pl.nb) ? kFirWarps - wp : 0;      // tile c0 - halo_k
    const bool halo_fast = halo_k > 0 && pl.aligned;
    if (ar.stagger_ns && (wp >> 2)) __nanosleep((unsigned)(wp >> 2) * ar.stagger_ns);   // de-phase the 4 warps of each scheduler
    // the lane powers first (one load per thread: ahead of, not behind, the 128 KB of observations this SM is about to request)
    static_assert(D * D * 32 <= 2 * kFirThreads, "at most two lane-power entries per thread");
    const double plane_v = tid < D * D * 32 ? __ldg(ar.plane + tid) : 0.0;
    const double plane_v2 = tid + kFirThreads < D * D * 32 ? __ldg(ar.plane + tid + kFirThreads) : 0.0;
    if (halo_fast) fir_issue_tile(buf, ys + (c0 - halo_k) * kFirTile, lane, pol);
    else if (n_items > 0 && !(nA + n_fast == 0 && nC > 0)) issue(item(0));            // first tile on its way before anything else
    fir_cp_commit();
    if (tid < D * D * 32) splane[tid] = plane_v;
    if (tid + kFirThreads < D * D * 32) splane[tid + kFirThreads] = plane_v2;
    for (int i = tid; i < (kFirRing + 2 * kFirNbMax) * D; i += kFirThreads) fir_st_word(sring + i, 0.0, -1);
    __syncthreads();
    if (tid == 0) fir_trace(ar, 1);
    // ---- what precedes the chunk --------------------------------------------------------------------------------------
    if (b == 0 && !exch_halo && !ar.x.local_halo) {
        const double qh = fir_head<D>(pl, ar, sscan, sred, sring);
        if (tid == 0) ar.partials[G] = qh;
    } else if (halo_k > 0) {
        if (halo_fast) {
            double yv[kFirL];
            fir_cp_wait_all();
            __syncwarp();
            fir_read_tile(buf, lane, yv);
            __syncwarp();
            const double* nx = nullptr;
            if (n_items > 0) {
                const Item x0 = item(0);
                if (x0.nvalid == kFirTile) nx = x0.src;
                else { fir_issue_tile_zfill(buf, x0.src, x0.nvalid, lane, pol); }
            }
            fir_tile_compute<D, false>(pl, yv, kFirNbMax - halo_k, kFirTile, true, false, splane, sring, lane, buf, nx, pol);
        } else {
            fir_tile_guarded<D>(pl, ys + (c0 - halo_k) * kFirTile, kFirNbMax - halo_k, kFirTile, true, false, splane, sring, lane);
        }
    }
    // The last nb tiles of this shard -> the successor's halo ring slot: NVLink stores by the last kFirPushWarps warps of the last CTA at
    // the START of the run; each of them fences after its first tile (the stores have landed by then: the fence is cheap) and the last
    // one to do so raises the successor's flag — early in this kernel's life, long before the successor needs the halo.
    const bool pusher = b == G - 1 && ar.x.push_dst != nullptr && wp >= kFirWarps - kFirPushWarps;
    if (pusher) {
        if (lane == 0 && ar.x.ack_in && ar.epoch > ar.x.ring) {      // the ring slot must have been consumed
            unsigned spins = 0;
            unsigned long long t0 = 0ull;
            while (fir_ld_sys(ar.x.ack_in) + ar.x.ring < ar.epoch) {
                fir_spin_check(spins, t0);
                __nanosleep(100);
            }
        }
        __syncwarp();
        const int n = pl.nb * kFirTile;
        const double* src = ar.y + pl.T - n;
#pragma unroll 8
        for (int e = (wp - (kFirWarps - kFirPushWarps)) * 32 + lane; e < n; e += kFirPushWarps * 32) ar.x.push_dst[e] = __ldg(src + e);
    }
    auto push_done = [&]() {
        __threadfence_system();
        __syncwarp();
        if (lane == 0 && atomicAdd(s_push, 1) == kFirPushWarps - 1) *reinterpret_cast<volatile unsigned long long*>(ar.x.push_flag) = ar.epoch;
    };
    if (tid == 0) fir_trace(ar, 2);
    // ---- the chunk --------------------------------------------------------------------------------------------------------
    double q = 0.0;
    if (deferred && !pl.aligned)
        fir_tile_guarded<D>(pl, ys + (c0 + wp) * kFirTile, wp + kFirNbMax, (int)min(Ts - (c0 + wp) * kFirTile, (long long)kFirTile), true, false,
                            splane, sring, lane);
    if (nA + n_fast == 0 && nC > 0) {      // nothing staged before the halo: its first tile could not be prefetched above
        wait_halo();
        issue(item(0));
        fir_cp_commit();
    }
#pragma unroll 1
    for (int it = 0; it < n_items; ++it) {
        const Item cur = item(it);
        fir_cp_wait_all();
        __syncwarp();
        if (cur.nvalid != kFirTile) {      // the partial last tile of the series (zero-filled beyond its end): masked sums
            double yt[kFirL];
            fir_read_tile(buf, lane, yt);
            q += fir_tile_compute<D, true>(pl, yt, cur.r, cur.nvalid, cur.pub, cur.full, splane, sring, lane);
            __syncwarp();
            if (it + 1 < n_items) {        // (only when one CTA runs the whole series)
                if (nC > 0 && it + 1 == nA + n_fast) wait_halo();
                issue(item(it + 1));
            }
            fir_cp_commit();
            if (pusher && it == 0) push_done();      // (a pusher warp whose first item is the partial tile must still report its push)
            continue;
        }
        double yv[kFirL];
        fir_read_tile(buf, lane, yv);
        __syncwarp();
        const double* __restrict__ next = nullptr;
        if (it + 1 < n_items) {
            if (nC > 0 && it + 1 == nA + n_fast) wait_halo();      // the next item is the first halo tile: its data must have arrived
            const Item nx = item(it + 1);
            if (nx.nvalid == kFirTile) next = nx.src;              // its copies are spread over this tile's pass A
            else fir_issue_tile_zfill(buf, nx.src, nx.nvalid, lane, pol);   // the partial last tile: zero-filled copies, issued now
        }
        q += pl.zero_mean ? fir_tile_compute<D, false, true>(pl, yv, cur.r, kFirTile, cur.pub, cur.full, splane, sring, lane, buf, next, pol)
                          : fir_tile_compute<D, false, false>(pl, yv, cur.r, kFirTile, cur.pub, cur.full, splane, sring, lane, buf, next, pol);
        if (pusher && it == 0) push_done();
        if (tid == 0 && it == 0) fir_trace(ar, 3);
    }
    if (tid == 0) fir_trace(ar, 4);
    if (pusher && n_items == 0) push_done();
#pragma unroll 1
    for (int it = n_fast; it < n_main; ++it) {      // unaligned series (all tiles) or the partial last tile
        const long long s0 = (first + (long long)it * kFirWarps) * kFirTile;
        q += fir_tile_guarded<D>(pl, ys + s0, (int)(first - c0) + it * kFirWarps + kFirNbMax, (int)min(Ts - s0, (long long)kFirTile), true, true,
                                 splane, sring, lane);
    }
    if (!pl.aligned) {
        if (exch_halo && wp == 0) {
            wait_halo();
            for (int k = pl.nb; k >= 1; --k)
                fir_tile_guarded<D>(pl, ar.x.halo + (size_t)(pl.nb - k) * kFirTile, kFirNbMax - k, kFirTile, true, false, splane, sring, lane);
        }
        if (deferred)
            q += fir_tile_guarded<D>(pl, ys + (c0 + wp) * kFirTile, wp + kFirNbMax, (int)min(Ts - (c0 + wp) * kFirTile, (long long)kFirTile), false,
                                     true, splane, sring, lane);
    }
    fir_cp_wait_all();
    // ---- fixed-order reductions; the last CTA to finish forms the log-likelihood ------------------------------------------
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_down_sync(0xffffffffu, q, off);
    __syncthreads();
    if (lane == 0) sred[wp] = q;
    __syncthreads();
    if (tid == 0) {
        fir_trace(ar, 5);
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kFirWarps; ++i) t += sred[i];
        // No CTA of this call leaves before the PREVIOUS call on the stream has completed (a no-op without programmatic dependent
        // launch): with that, the call after this one — which can only start once every CTA of this one has started, i.e. once SMs
        // have been vacated — never overlaps the previous one, so at most two calls are in flight and the parity-indexed
        // counters / partials / result below (and the exchange ring) are never shared by two live calls.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        fir_trace(ar, 6);
        __stcg(ar.partials + b, t);
        if (b == 0 && (exch_halo || ar.x.local_halo)) __stcg(ar.partials + G, 0.0);
        __threadfence();
        s_last = atomicAdd(ar.counters, 1u) == (unsigned)G - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double s = 0.0;
    for (long long i = tid; i < G; i += kFirThreads) s += __ldcg(ar.partials + i);
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
        fir_trace(ar, 7);
        ar.counters[0] = 0u;
        if (ar.x.ack_out)      // the halo of this call has been consumed (and so have those of all earlier calls: see the wait above)
            *reinterpret_cast<volatile unsigned long long*>(ar.x.ack_out) = ar.epoch;
        if (ar.x.peers) {      // this shard's log-likelihood -> every rank's buffer: ONE 16-byte {value, epoch} store per peer, no fence
            for (int p = 0; p < ar.x.world; ++p) {
                double2* w = reinterpret_cast<double2*>(ar.x.peers[p] + ar.x.lml_off);
                asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(w), "d"(lml), "d"(__longlong_as_double((long long)ar.epoch)) : "memory");
            }
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
    const int full_grid = h->sm_count * kFirCtasPerSM;                                                       // 16 warps per SM
    const unsigned G = (unsigned)std::max<long long>(1, std::min<long long>(full_grid, ntiles / (kFirWarps / 2)));
    // counters / partials / result alternate with the parity of the call, so two consecutive calls may overlap (PDL)
    const size_t pstride = ((size_t)full_grid + 8) & ~size_t(7);
    TGP_TRY(fir_grow(h, &st.partials, &st.partials_cap, 2 * pstride * sizeof(double), false));
    if (!st.counters) {
        TGP_CUDA(h, cudaMalloc((void**)&st.counters, 64 * sizeof(unsigned)));
        TGP_CUDA(h, cudaMemsetAsync(st.counters, 0, 64 * sizeof(unsigned), h->stream));
        TGP_CUDA(h, cudaMalloc((void**)&st.result, 8 * sizeof(double)));
    }
    FirArgs ar{};
    ar.y = dy;
    ar.tab = st.dev;
    ar.plane = st.dev + (size_t)pl.N0 * (D + 1);
    ar.epoch = ++st.epoch;
    const int par = (int)(ar.epoch & 1ull);
    ar.counters = st.counters + 32 * par;
    ar.partials = st.partials + pstride * par;
    ar.result = st.result + 4 * par;
    ar.lml_user = (lml_out && is_device_ptr(lml_out)) ? lml_out : nullptr;
    if (xc) { ar.x = *xc; ar.epoch = xc->epoch; }
    static int stagger = -1, pdl = -1;
    if (stagger < 0) { const char* e = getenv("TGP_FIR_STAGGER"); stagger = e ? atoi(e) : 0; }
    if (pdl < 0) { const char* e = getenv("TGP_FIR_PDL"); pdl = e ? atoi(e) : 1; }
    ar.stagger_ns = (unsigned)stagger;
    static int head_less = -1;
    if (head_less < 0) { const char* e = getenv("TGP_FIR_HEADLESS"); head_less = e ? atoi(e) : 0; }
    ar.head_less = head_less;
    static int trace = -1;
    if (trace < 0) trace = getenv("TGP_FIR_TRACE") ? 1 : 0;
    unsigned long long* dtrace = nullptr;
    if (trace) {
        TGP_TRY(dalloc(h, (size_t)G * 8, &dtrace));
        TGP_CUDA(h, cudaMemsetAsync(dtrace, 0, (size_t)G * 64, h->stream));
        ar.trace = dtrace;
    }
    ar.early_trigger = (pdl && (int)G == full_grid) ? 1 : 0;
    TGP_K(h, "k_fir_logpdf");
    constexpr size_t smem = FirSmem<D>::bytes;
    static bool attr_set[64] = {false};
    if (!attr_set[h->device & 63]) {
        TGP_CUDA(h, cudaFuncSetAttribute(k_fir_logpdf<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TGP_CUDA(h, cudaFuncSetAttribute(k_fir_logpdf<D>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set[h->device & 63] = true;
    }
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(G);
        cfg.blockDim = dim3(kFirThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = (pdl && !h->timing) ? 1 : 0;
        TGP_CUDA(h, cudaLaunchKernelEx(&cfg, k_fir_logpdf<D>, pl, ar));
    }
    TGP_LAUNCH_CHECK(h);
    if (trace) {      // debug timeline: per stamp, the earliest / latest CTA relative to the first CTA's entry (microseconds)
        std::vector<unsigned long long> tr((size_t)G * 8);
        TGP_CUDA(h, cudaMemcpyAsync(tr.data(), dtrace, tr.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        unsigned long long t0 = ~0ull;
        for (unsigned c = 0; c < G; ++c) t0 = std::min(t0, tr[(size_t)c * 8]);
        static const char* nm[8] = {"entry", "plane+ring ready", "head/halo done", "first item done", "items done", "CTA done", "prev call complete", "lml written"};
        fprintf(stderr, "[tgp fir trace] T=%lld G=%u N0=%lld nb=%d\n", (long long)pl.T, G, (long long)pl.N0, pl.nb);
        for (int k = 0; k < 8; ++k) {
            unsigned long long lo = ~0ull, hi = 0, c0v = tr[k];
            for (unsigned c = 0; c < G; ++c) { const unsigned long long v = tr[(size_t)c * 8 + k]; if (v) { lo = std::min(lo, v); hi = std::max(hi, v); } }
            if (hi) fprintf(stderr, "   %-20s min %7.2f  max %7.2f  CTA0 %7.2f us\n", nm[k], (lo - t0) * 1e-3, (hi - t0) * 1e-3, c0v ? (c0v - t0) * 1e-3 : -1.0);
        }
        {   // the five slowest CTAs
            std::vector<std::pair<unsigned long long, unsigned>> v;
            for (unsigned c = 0; c < G; ++c) v.push_back({tr[(size_t)c * 8 + 5], c});
            std::sort(v.begin(), v.end());
            for (size_t i = v.size() > 5 ? v.size() - 5 : 0; i < v.size(); ++i) {
                const unsigned c = v[i].second;
                fprintf(stderr, "   slow CTA %3u: ready %6.2f pre %6.2f first %6.2f items %6.2f done %6.2f\n", c, (tr[c * 8 + 1] - t0) * 1e-3,
                        (tr[c * 8 + 2] - t0) * 1e-3, (tr[c * 8 + 3] - t0) * 1e-3, (tr[c * 8 + 4] - t0) * 1e-3, (tr[c * 8 + 5] - t0) * 1e-3);
            }
        }
    }
    *handled = true;
    if (lml_out && !ar.lml_user) {
        TGP_CUDA(h, cudaMemcpyAsync(h->pinned + 8, ar.result, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TGP_CUDA(h, cudaStreamSynchronize(h->stream));
        h->d2h += 8;
        *lml_out = h->pinned[8];
    }
    return TGP_OK;
}

}  // namespace tgp

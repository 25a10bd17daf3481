// tgp_tc_gemm.cuh — the tensor-core contraction of the large-state path (BASELINE config 5; SmallOutputLGC step,
// linear_gaussian_conditionals.jl:46-52, 129-141), hand-written for sm_100a: TMA (cp.async.bulk.tensor, 128-byte swizzle)
// -> shared memory -> tcgen05.mma kind::tf32 with the accumulator in TMEM -> tcgen05.ld epilogue.
//
//   C[m, n] = alpha * sum_k X[k, m] * Y[k, n]  (+ Cin[m, n])        X: K x Mx,  Y: K x N,  both COLUMN-major,
//
// i.e. every operand is "K-major" (the contraction index is the contiguous one), which is how all products of a Kalman
// step can be phrased once A' and H' are stored (see tgp_dense_tc.cu) — no transposed-operand descriptors needed.
//
// Precision: FP32 storage (the reference's ArrayStorage(Float32)); one TF32 MMA keeps 10 mantissa bits, which is not
// enough for the covariance downdate P - B'B, so every operand is kept as a PAIR of planes
//     hi = x with the low 13 mantissa bits cleared (exactly a TF32 number),   lo = x - hi (exact in FP32)
// and a product is three MMAs into the same TMEM accumulator: hi*hi + hi*lo + lo*hi ("3xTF32", ~2^-21 relative).
// A pair buffer is ONE 2-D tensor: hi plane in columns [0, cpad), lo plane in columns [cpad, 2 cpad).
//
// CTA = 192 threads: warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 = TMA producer, warp 5 = MMA issuer
// (one elected lane) and TMEM allocator. Tile 128 x BN x 32, kStages-deep mbarrier ring. One output tile per CTA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tgp {
namespace tc {

constexpr int BM = 128;        // UMMA M (TMEM lanes)
constexpr int BK = 32;         // 32 floats = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 8;      // tf32: 32 bytes per MMA along K
constexpr int kThreads = 192;

struct Pair {                  // hi/lo planes of a column-major matrix (rows x cols), leading dimension ld (floats)
    float* p = nullptr;
    int rows = 0, cols = 0, ld = 0, cpad = 0;
    __host__ __device__ float* hi() const { return p; }
    __host__ __device__ float* lo() const { return p + (size_t)ld * cpad; }
    size_t floats() const { return (size_t)ld * cpad * 2; }
};

struct Epi {
    int Mx = 0, N = 0;          // valid extent of C
    float alpha = 1.f;
    // additive term (all nullable): a pair (hi + lo), a dense double matrix, or a double diagonal / scalar
    const float* cin_hi = nullptr; const float* cin_lo = nullptr; int ld_cin = 0;
    const double* cin_d = nullptr; int ld_cind = 0;     // dense, column-major
    const double* cin_diag = nullptr; int diag_stride = 0;   // diag_stride 0: scalar broadcast on the diagonal
    // outputs (all nullable)
    float* out_hi = nullptr; float* out_lo = nullptr; int ld_out = 0;      // C as a pair
    float* outT_hi = nullptr; float* outT_lo = nullptr; int ld_outT = 0;   // C' as a pair
    double* out_d = nullptr; int ld_outd = 0;                              // C in double
    int symmetric = 0;          // C is symmetric: only tiles touching the upper triangle are computed, (m <= n) is mirrored
    // steady-state detection (nullable): conv[0] = max |C_new - C_old| and conv[1] = max |C_new| as float bits (atomicMax),
    // C_old being what the pair output buffer held before this launch overwrote it (the previous step's covariance)
    unsigned* conv = nullptr;
    // time-blocked mean recursion (nullable): per-row additive constant; rows >= split_row are not stored but their squares
    // are summed per column into colsq[n * colsq_stride] (double atomics)
    const float* row_bias = nullptr;
    int split_row = 1 << 30;
    double* colsq = nullptr;
    long long colsq_stride = 0;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a broken pipeline traps (launch error reported to the caller) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (it > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused, version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                           // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int BN>
struct Cfg {
    static constexpr int kStages = BN == 128 ? 3 : 4;
    static constexpr int kXBytes = BM * 128, kYBytes = BN * 128;
    static constexpr int kStageBytes = 2 * kXBytes + 2 * kYBytes;
    static constexpr int kSmem = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
    // The tensor core adds each K = 8 partial product to the FP32 accumulator with truncation, so the error of one
    // accumulator grows linearly with the number of MMAs issued into it (measured: 3.5e-5 relative at K = 768 with a
    // single accumulator). The main term hi*hi is therefore spread round-robin over kSplit TMEM accumulators (k-block
    // kb -> kb % kSplit), the two small cross terms share one more, and the epilogue adds them up in FP32 (RN).
    static constexpr int kSplit = BN == 64 ? 4 : 3;
    static constexpr int kTmemCols = 512;
    static_assert((kSplit + 1) * BN <= kTmemCols, "accumulators exceed TMEM");
};

// Operand sources of one launch. The Y operand may be the CONCATENATION along K of two tensors (k-blocks < k_switch from tmY,
// the rest from tmY2): the time-blocked mean recursion multiplies [state; observations] without materialising the stack.
struct Src {
    int K = 0;
    int lo_col_x = 0, lo_col_y = 0, lo_col_y2 = 0;   // column offset of the lo plane inside each pair tensor
    int k_switch = 1 << 30;                           // first k-block read from tmY2 (at K coordinate (kb - k_switch) * BK)
    int y_col_shift = 0;                              // added to the column coordinate of tmY (negative columns read as zero)
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_gemm_tn(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmY2,
             const Src src, const Epi e) {
    using C = Cfg<BN>;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (e.symmetric && m0 >= n0 + BN) return;   // tile strictly below the diagonal: produced by its mirror
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + C::kStages * C::kStageBytes);
    uint64_t* empty = full + C::kStages;
    uint64_t* accum = empty + C::kStages;
    uint32_t* tmem_slot = (uint32_t*)(accum + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (src.K + BK - 1) / BK;

    if (warp == 4 && lane == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % C::kStages;
                const uint32_t ph = (kb / C::kStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * C::kStageBytes;
                mbar_expect_tx(&full[s], C::kStageBytes);
                tma_load_2d(st, &tmX, &full[s], kb * BK, m0);
                tma_load_2d(st + C::kXBytes, &tmX, &full[s], kb * BK, src.lo_col_x + m0);
                if (kb < src.k_switch) {
                    tma_load_2d(st + 2 * C::kXBytes, &tmY, &full[s], kb * BK, n0 + src.y_col_shift);
                    tma_load_2d(st + 2 * C::kXBytes + C::kYBytes, &tmY, &full[s], kb * BK, src.lo_col_y + n0 + src.y_col_shift);
                } else {
                    tma_load_2d(st + 2 * C::kXBytes, &tmY2, &full[s], (kb - src.k_switch) * BK, n0);
                    tma_load_2d(st + 2 * C::kXBytes + C::kYBytes, &tmY2, &full[s], (kb - src.k_switch) * BK, src.lo_col_y2 + n0);
                }
            }
        }
    } else if (warp == 5) {
        // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 / 10-12), K-major both, N>>3 at 17, M>>4 at 24
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int kb = 0; kb < nk; ++kb) {
            const int s = kb % C::kStages;
            const uint32_t ph = (kb / C::kStages) & 1;
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + s * C::kStageBytes);
                const uint64_t dxh = umma_desc_sw128(sa), dxl = umma_desc_sw128(sa + C::kXBytes);
                const uint64_t dyh = umma_desc_sw128(sa + 2 * C::kXBytes), dyl = umma_desc_sw128(sa + 2 * C::kXBytes + C::kYBytes);
                const uint32_t acc_main = tmem_base + (uint32_t)((kb % C::kSplit) * BN);
                const uint32_t acc_cross = tmem_base + (uint32_t)(C::kSplit * BN);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);   // +32 bytes along K inside the swizzle atom
                    umma_tf32(acc_cross, dxl + adv, dyh + adv, idesc, (kb | k) != 0);
                    umma_tf32(acc_cross, dxh + adv, dyl + adv, idesc, 1);
                    umma_tf32(acc_main, dxh + adv, dyh + adv, idesc, (kb >= C::kSplit) || k != 0);
                }
                umma_commit(&empty[s]);                 // frees the stage when these MMAs have read it
                if (kb == nk - 1) umma_commit(accum);   // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // epilogue: warp w reads TMEM lanes [32w, 32w+32) = rows m0 + 32w + lane
        mbar_wait(accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = m0 + warp * 32 + lane;
        const bool row_ok = m < e.Mx;
        const bool row_sq = row_ok && m >= e.split_row;       // rows whose squares are summed per column instead of being stored
        const float bias = (row_ok && e.row_bias) ? e.row_bias[m] : 0.f;
        float cmax_d = 0.f, cmax_a = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            float acc[32];
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            tmem_ld32(trow + (uint32_t)(C::kSplit * BN), v);          // cross terms
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = __uint_as_float(v[c]);
            const int nacc = nk < C::kSplit ? nk : C::kSplit;
#pragma unroll 1
            for (int j = nacc - 1; j >= 0; --j) {
                tmem_ld32(trow + (uint32_t)(j * BN), v);
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c] += __uint_as_float(v[c]);
            }
            const int nb0 = n0 + c0;
            if (nb0 >= e.N) break;                               // uniform: the whole chunk is outside C
            // ---- phase A: every global LOAD of the chunk (independent, so they overlap) --------------------------------
            float add[32], old[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) { add[c] = 0.f; old[c] = 0.f; }
            const bool store_row = row_ok && !row_sq;
            if (store_row && e.cin_hi) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = nb0 + c;
                    if (n < e.N && (!e.symmetric || m <= n))
                        add[c] = e.cin_hi[(size_t)m + (size_t)e.ld_cin * n] + e.cin_lo[(size_t)m + (size_t)e.ld_cin * n];
                }
            }
            if (store_row && e.conv && e.out_hi) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = nb0 + c;
                    if (n < e.N && (!e.symmetric || m <= n))
                        old[c] = e.out_hi[(size_t)m + (size_t)e.ld_out * n] + e.out_lo[(size_t)m + (size_t)e.ld_out * n];
                }
            }
            // ---- phase B: values ------------------------------------------------------------------------------------------
            float hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int n = nb0 + c;
                float x = fmaf(e.alpha, acc[c], add[c]) + bias;
                if (row_ok && (e.cin_d || e.cin_diag) && n < e.N) {
                    double xd = (double)x;
                    if (e.cin_d) xd += e.cin_d[(size_t)m + (size_t)e.ld_cind * n];
                    if (e.cin_diag && m == n) xd += e.cin_diag[(size_t)e.diag_stride * m];
                    if (e.out_d && (!e.symmetric || m <= n)) {
                        e.out_d[(size_t)m + (size_t)e.ld_outd * n] = xd;
                        if (e.symmetric) e.out_d[(size_t)n + (size_t)e.ld_outd * m] = xd;
                    }
                    x = (float)xd;
                } else if (row_ok && e.out_d && n < e.N && (!e.symmetric || m <= n)) {
                    e.out_d[(size_t)m + (size_t)e.ld_outd * n] = (double)x;
                    if (e.symmetric) e.out_d[(size_t)n + (size_t)e.ld_outd * m] = (double)x;
                }
                acc[c] = x;
                hi[c] = tf32_hi(x);
                lo[c] = x - hi[c];
            }
            // ---- column sums of squares of the rows >= split_row (whitened residuals of the time-blocked recursion) -------
            if (e.colsq) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    float q = (row_sq && nb0 + c < e.N) ? acc[c] * acc[c] : 0.f;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
                    if (lane == 0 && nb0 + c < e.N && m0 + warp * 32 + 31 >= e.split_row) atomicAdd(e.colsq + (size_t)(nb0 + c) * e.colsq_stride, (double)q);
                }
            }
            if (!store_row) continue;
            // ---- phase C: stores ------------------------------------------------------------------------------------------
            if (e.out_hi) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = nb0 + c;
                    if (n < e.N && (!e.symmetric || m <= n)) {
                        if (e.conv) {
                            cmax_d = fmaxf(cmax_d, fabsf(acc[c] - old[c]));
                            cmax_a = fmaxf(cmax_a, fabsf(acc[c]));
                        }
                        e.out_hi[(size_t)m + (size_t)e.ld_out * n] = hi[c];      // lanes = consecutive m: coalesced
                        e.out_lo[(size_t)m + (size_t)e.ld_out * n] = lo[c];
                    }
                }
            }
            // transposed copy (C'), or the mirror image of a symmetric C: this thread's 32 values are contiguous there
            float* th = e.symmetric ? e.out_hi : e.outT_hi;
            float* tl = e.symmetric ? e.out_lo : e.outT_lo;
            const int ldt = e.symmetric ? e.ld_out : e.ld_outT;
            if (th) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = nb0 + c;
                    if (n < e.N && (!e.symmetric || m < n)) {
                        th[(size_t)n + (size_t)ldt * m] = hi[c];
                        tl[(size_t)n + (size_t)ldt * m] = lo[c];
                    }
                }
            }
        }
        if (e.conv) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                cmax_d = fmaxf(cmax_d, __shfl_xor_sync(0xffffffffu, cmax_d, off));
                cmax_a = fmaxf(cmax_a, __shfl_xor_sync(0xffffffffu, cmax_a, off));
            }
            if (lane == 0) { atomicMax(e.conv, __float_as_uint(cmax_d)); atomicMax(e.conv + 1, __float_as_uint(cmax_a)); }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
    }
}

}  // namespace tc
}  // namespace tgp

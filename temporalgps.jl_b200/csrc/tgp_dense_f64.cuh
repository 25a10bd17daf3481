// tgp_dense_f64.cuh — the FP64 dense products of the large-state step (tgp_dense.cu), hand-written: no library call is left on
// the path. All matrices column-major.
//   k_dgemm<TA, TB, SYMB>   C = alpha * op(A) * op(B) + beta * C      64 x 64 x 16 tiles in shared memory, 4 x 4 outputs per thread,
//                           the next tile's global loads in flight while the current one is multiplied. SYMB: B is read as
//                           Symmetric(B, :U) — the reference's `A * Symmetric(P)` (LGC:49) reads the upper triangle only.
//   k_dgemv_n / k_dgemv_t   y = alpha * op(A) x + beta * y
//   k_trsm_ut               X <- U' \ X for an upper-triangular U (one warp per right-hand side, forward substitution with the
//                           column of U read coalesced and the solved entries kept in shared memory)          (LGC:133-134)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>

namespace tgp {

constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16, kGemmThreads = 256;

template <bool TA, bool TB, bool SYMB>
__global__ void __launch_bounds__(kGemmThreads)
k_dgemm(int M, int N, int K, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, double beta,
        double* __restrict__ C, int ldc, int kchunk, double* __restrict__ part) {
    __shared__ double As[kGemmBK][kGemmBM + 2];
    __shared__ double Bs[kGemmBK][kGemmBN + 2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * kGemmBM, j0 = blockIdx.y * kGemmBN;
    // global -> register staging: 4 elements of each operand tile per thread, coalesced along the operand's contiguous index
    double ra[4], rb[4];
    const int kend = min(K, (int)(blockIdx.z + 1) * kchunk);
    auto load = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i, k;
            if (!TA) { i = tid & 63; k = (tid >> 6) + 4 * r; } else { k = tid & 15; i = (tid >> 4) + 16 * r; }
            const int gi = i0 + i, gk = k0 + k;
            ra[r] = (gi < M && gk < kend) ? (TA ? A[(size_t)gk + (size_t)lda * gi] : A[(size_t)gi + (size_t)lda * gk]) : 0.0;
            int j, kb;
            if (!TB) { kb = tid & 15; j = (tid >> 4) + 16 * r; } else { j = tid & 63; kb = (tid >> 6) + 4 * r; }
            const int gj = j0 + j, gkb = k0 + kb;
            double v = 0.0;
            if (gj < N && gkb < kend) {
                if (SYMB) v = gkb <= gj ? B[(size_t)gkb + (size_t)ldb * gj] : B[(size_t)gj + (size_t)ldb * gkb];
                else v = TB ? B[(size_t)gj + (size_t)ldb * gkb] : B[(size_t)gkb + (size_t)ldb * gj];
            }
            rb[r] = v;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i, k;
            if (!TA) { i = tid & 63; k = (tid >> 6) + 4 * r; } else { k = tid & 15; i = (tid >> 4) + 16 * r; }
            As[k][i] = ra[r];
            int j, kb;
            if (!TB) { kb = tid & 15; j = (tid >> 4) + 16 * r; } else { j = tid & 63; kb = (tid >> 6) + 4 * r; }
            Bs[kb][j] = rb[r];
        }
    };
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    // split K: slice blockIdx.z covers [kb, ke) and stores its raw partial product into part[z] (k_splitk_reduce finishes)
    const int kb = blockIdx.z * kchunk, ke = min(K, kb + kchunk);
    load(kb);
    for (int k0 = kb; k0 < ke; k0 += kGemmBK) {
        stash();
        __syncthreads();
        if (k0 + kGemmBK < ke) load(k0 + kGemmBK);
#pragma unroll
        for (int k = 0; k < kGemmBK; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[k][tx + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[k][ty + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int gj = j0 + ty + 16 * b;
        if (gj >= N) continue;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int gi = i0 + tx + 16 * a;
            if (gi >= M) continue;
            if (part) { part[(size_t)blockIdx.z * M * N + (size_t)gi + (size_t)M * gj] = acc[a][b]; continue; }
            double* c = C + (size_t)gi + (size_t)ldc * gj;
            *c = beta == 0.0 ? alpha * acc[a][b] : fma(alpha, acc[a][b], beta * *c);
        }
    }
}

// C = alpha * (part[0] + part[1] + ... in slice order) + beta * C
__global__ void __launch_bounds__(256) k_splitk_reduce(int M, int N, int S, double alpha, const double* __restrict__ part, double beta,
                                                        double* __restrict__ C, int ldc) {
    const long long n = (long long)M * N;
    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
        double s = 0.0;
        for (int z = 0; z < S; ++z) s += part[(size_t)z * n + e];
        double* c = C + (size_t)(e % M) + (size_t)ldc * (e / M);
        *c = beta == 0.0 ? alpha * s : fma(alpha, s, beta * *c);
    }
}

// W = U^-1 (upper triangular, FP64, column-major M x M, zeros below the diagonal) from the inverses of the 32 x 32 diagonal blocks
// (Dinv, written by k_chol_panel2): blocked back substitution, one warp per COLUMN c of W, lane = row inside the current block row I:
//     x_J = Dinv_J[:, c - 32 J];    x_I = -Dinv_I (sum_{l >= 32 (I + 1)}^{c} U[32 I + lane, l] x_l),  I = J - 1 .. 0.
// With it B = U' \ V and alpha = U' \ r are a GEMM and a GEMV (W' V, W' r): no sequential solve is left on the step.
// Dynamic shared memory: 8 * (Mp + 32) + 32 * Mp doubles, Mp = M rounded up to 32.
__global__ void __launch_bounds__(256) k_tri_inv_f64(const double* __restrict__ U, int M, const double* __restrict__ Dinv, double* __restrict__ W) {
    extern __shared__ double xbuf[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x * 8 + warp;             // the 8 columns of a CTA lie in the same block column J
    const bool live = c < M;
    const int Mp = (M + 31) / 32 * 32;
    double* x = xbuf + (size_t)warp * (Mp + 32);
    double* tb = x + Mp;
    double* tile = xbuf + (size_t)8 * (Mp + 32);
    const int J = (blockIdx.x * 8) / 32;
    const int cmax = min(blockIdx.x * 8 + 7, M - 1);
    if (live) x[32 * J + lane] = Dinv[(size_t)J * 1024 + lane + 32 * (c - 32 * J)];
    __syncwarp();
    for (int I = J - 1; I >= 0; --I) {
        const int l0 = 32 * (I + 1), nl = cmax - l0 + 1;
        __syncthreads();                              // the previous tile has been consumed
        for (int e = tid; e < 32 * nl; e += 256) tile[e] = U[(size_t)(32 * I + (e & 31)) + (size_t)M * (l0 + (e >> 5))];
        __syncthreads();
        if (live) {
            double t0 = 0.0, t1 = 0.0;
            int l = l0;
            for (; l + 1 <= c; l += 2) {
                t0 = fma(tile[(size_t)(l - l0) * 32 + lane], x[l], t0);
                t1 = fma(tile[(size_t)(l - l0 + 1) * 32 + lane], x[l + 1], t1);
            }
            if (l <= c) t0 = fma(tile[(size_t)(l - l0) * 32 + lane], x[l], t0);
            tb[lane] = t0 + t1;
            __syncwarp();
            const double* di = Dinv + (size_t)I * 1024 + lane;     // row `lane` of the upper-triangular block inverse
            double s = 0.0;
#pragma unroll 8
            for (int rp = 0; rp < 32; ++rp) s = fma(di[32 * rp], tb[rp], s);
            x[32 * I + lane] = -s;
            __syncwarp();
        }
    }
    if (!live) return;
    for (int i = lane; i < M; i += 32) W[(size_t)i + (size_t)M * c] = i <= c ? x[i] : 0.0;
}
inline size_t tri_inv_smem(int M) { const size_t Mp = (size_t)(M + 31) / 32 * 32; return sizeof(double) * (8 * (Mp + 32) + 32 * Mp); }

// y (M) = alpha * A (M x N) x + beta * y: 32 rows per CTA (lane = row, A read coalesced across the rows), the columns dealt round-robin
// to the CTA's 8 warps, partial sums folded through shared memory in a fixed order.
__global__ void __launch_bounds__(256) k_dgemv_n(int M, int N, double alpha, const double* __restrict__ A, int lda,
                                                  const double* __restrict__ x, double beta, double* __restrict__ y) {
    __shared__ double part[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    double s0 = 0.0, s1 = 0.0;
    if (i < M) {
        int k = w;
        for (; k + 8 < N; k += 16) {
            s0 = fma(A[(size_t)i + (size_t)lda * k], x[k], s0);
            s1 = fma(A[(size_t)i + (size_t)lda * (k + 8)], x[k + 8], s1);
        }
        if (k < N) s0 = fma(A[(size_t)i + (size_t)lda * k], x[k], s0);
    }
    part[w][lane] = s0 + s1;
    __syncthreads();
    if (w == 0 && i < M) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += part[q][lane];
        y[i] = beta == 0.0 ? alpha * s : fma(alpha, s, beta * y[i]);
    }
}

// y (N) = alpha * A' (A is M x N) x + beta * y: one warp per column.
__global__ void __launch_bounds__(256) k_dgemv_t(int M, int N, double alpha, const double* __restrict__ A, int lda,
                                                  const double* __restrict__ x, double beta, double* __restrict__ y) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= N) return;
    double s = 0.0;
    for (int i = lane; i < M; i += 32) s = fma(A[(size_t)i + (size_t)lda * j], x[i], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[j] = beta == 0.0 ? alpha * s : fma(alpha, s, beta * y[j]);
}

// X (M x N, ldx) <- U' \ X, U upper triangular (M x M, ldu). One warp per column of X; dynamic shared memory: warps * M doubles.
__global__ void k_trsm_ut(int M, int N, const double* __restrict__ U, int ldu, double* __restrict__ X, int ldx) {
    extern __shared__ double xs_all[];
    const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * wpb + w;
    if (j >= N) return;
    double* xs = xs_all + (size_t)w * M;
    double* xc = X + (size_t)ldx * j;
    for (int i = lane; i < M; i += 32) xs[i] = xc[i];
    __syncwarp();
    for (int i = 0; i < M; ++i) {
        const double* u = U + (size_t)ldu * i;     // column i of U: rows k < i multiply the entries solved so far
        double s = 0.0;
        for (int k = lane; k < i; k += 32) s = fma(u[k], xs[k], s);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) xs[i] = (xs[i] - s) / u[i];
        __syncwarp();
    }
    for (int i = lane; i < M; i += 32) xc[i] = xs[i];
}

// ---- host launchers ----------------------------------------------------------------------------------------------------------
enum class Op { N, T };

// ws / ws_doubles: optional workspace for split-K (raw partial products); target_ctas: what fills the machine (about 3 CTAs per SM)
inline cudaError_t dgemm(cudaStream_t st, Op ta, Op tb, bool symb, int M, int N, int K, double alpha, const double* A, int lda,
                         const double* B, int ldb, double beta, double* C, int ldc, double* ws = nullptr, size_t ws_doubles = 0,
                         int target_ctas = 444) {
    const int tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + kGemmBN - 1) / kGemmBN);
    int S = 1;
    if (ws && tiles < target_ctas) {
        S = target_ctas / tiles;
        if (S > 8) S = 8;
        while (S > 1 && ((size_t)S * M * N > ws_doubles || K / S < 4 * kGemmBK)) --S;
    }
    int kchunk = K;
    if (S > 1) {
        kchunk = ((K + S - 1) / S + kGemmBK - 1) / kGemmBK * kGemmBK;
        S = (K + kchunk - 1) / kchunk;
    }
    double* part = S > 1 ? ws : nullptr;
    const dim3 grid((M + kGemmBM - 1) / kGemmBM, (N + kGemmBN - 1) / kGemmBN, S);
    if (symb)                            k_dgemm<false, false, true><<<grid, kGemmThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kchunk, part);
    else if (ta == Op::N && tb == Op::N) k_dgemm<false, false, false><<<grid, kGemmThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kchunk, part);
    else if (ta == Op::N && tb == Op::T) k_dgemm<false, true, false><<<grid, kGemmThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kchunk, part);
    else if (ta == Op::T && tb == Op::N) k_dgemm<true, false, false><<<grid, kGemmThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kchunk, part);
    else                                 k_dgemm<true, true, false><<<grid, kGemmThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kchunk, part);
    if (S > 1) {
        const long long n = (long long)M * N;
        k_splitk_reduce<<<(unsigned)std::min<long long>((n + 255) / 256, 1184), 256, 0, st>>>(M, N, S, alpha, part, beta, C, ldc);
    }
    return cudaGetLastError();
}

inline cudaError_t dgemv(cudaStream_t st, Op ta, int M, int N, double alpha, const double* A, int lda, const double* x, double beta, double* y) {
    if (ta == Op::N) k_dgemv_n<<<(M + 31) / 32, 256, 0, st>>>(M, N, alpha, A, lda, x, beta, y);
    else             k_dgemv_t<<<(N + 7) / 8, 256, 0, st>>>(M, N, alpha, A, lda, x, beta, y);
    return cudaGetLastError();
}

// returns cudaErrorInvalidValue if M is too large for one warp's shared-memory column (M <= 6144 with the default 48 KB)
inline cudaError_t trsm_ut(cudaStream_t st, int M, int N, const double* U, int ldu, double* X, int ldx) {
    int wpb = (int)((48 * 1024) / (sizeof(double) * (size_t)M));
    if (wpb < 1) return cudaErrorInvalidValue;
    if (wpb > 8) wpb = 8;
    k_trsm_ut<<<(N + wpb - 1) / wpb, wpb * 32, (size_t)wpb * M * sizeof(double), st>>>(M, N, U, ldu, X, ldx);
    return cudaGetLastError();
}

}  // namespace tgp

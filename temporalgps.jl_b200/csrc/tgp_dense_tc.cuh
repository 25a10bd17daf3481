// tgp_dense_tc.cuh — FP32-storage tensor-core variant of the large-state path (BASELINE config 5: ArrayStorage(Float32),
// D = 768, M = 256). Included at the end of tgp_dense.cu (shares its Cholesky kernels). Selected with
// tgp_set_option(h, TGP_OPT_DENSE_MATH, TGP_DENSE_TF32X3).
//
// Every product of the step is phrased as  C = X' Y  with X, Y column-major and the contraction index contiguous
// ("TN", both operands K-major), so ONE tcgen05 kernel (tgp_tc_gemm.cuh) serves all of them. With At = A', Ht = H'
// stored once (time-invariant models) and P symmetric:
//     predict (LGC:46-52)      W  = P' At            (= (A P)')            D x D
//                              Pp = W' At + Q        (= A P A' + Q)        D x D   symmetric
//     update (LGC:129-141)     Vt = Pp' Ht           (= (H Pp)')           D x M   (+ its transpose V, M x D)
//                              S  = Vt' Ht + R       (= H Pp H' + R)       M x M   symmetric, emitted in FP64
//                              U  = chol(S) in FP64 (k_chol_panel2 / k_chol_trail),  Winv = U^-1  (k_tri_inv2)
//                              B  = Winv' V          (= U' \ V)            M x D
//                              P  = Pp - B' B                              D x D   symmetric
// Means / residual / likelihood are FP64 GEMVs over the FP32 matrices (k_gemv_pair), lml_t in FP64.
#pragma once
#include "tgp_tc_gemm.cuh"

namespace tgp {

using tc::Pair;

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

struct TcOp {             // a pair buffer and its TMA descriptors (box = 32 floats along K x {128, 64} columns)
    Pair pr;
    CUtensorMap m128, m64;
};

// 2-D FP32 tensor map: `rows` contiguous elements per column, `ncols` columns `col_stride_bytes` apart, box 32 x box_cols, 128-byte swizzle.
static int tc_encode_map(tgp_ctx* h, const float* base, int rows, long long ncols, size_t col_stride_bytes, unsigned box_cols, CUtensorMap* out) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return fail(h, TGP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)ncols};
    const cuuint64_t strides[1] = {(cuuint64_t)col_stride_bytes};
    const cuuint32_t es[2] = {1, 1};
    const cuuint32_t box[2] = {(cuuint32_t)tc::BK, box_cols};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, TGP_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows %d cols %lld)", (int)r, rows, ncols);
    return TGP_OK;
}

// Zero-initialised pair buffer + its maps. extra_cols: additional zero columns behind each plane (a shifted read of the lo
// plane at a negative column then lands in the zero padding of the hi plane).
static int tc_make_op(tgp_ctx* h, int rows, int cols, TcOp* op, int extra_cols = 0) {
    Pair& p = op->pr;
    p.rows = rows; p.cols = cols;
    p.ld = (rows + 31) / 32 * 32;
    p.cpad = (cols + extra_cols + 127) / 128 * 128;
    TGP_TRY(dalloc(h, p.floats(), &p.p));
    TGP_CUDA(h, cudaMemsetAsync(p.p, 0, p.floats() * sizeof(float), h->stream));
    TGP_TRY(tc_encode_map(h, p.p, rows, 2LL * p.cpad, (size_t)p.ld * sizeof(float), 128u, &op->m128));
    TGP_TRY(tc_encode_map(h, p.p, rows, 2LL * p.cpad, (size_t)p.ld * sizeof(float), 64u, &op->m64));
    return TGP_OK;
}

// once per process and device, outside any stream capture
static int tc_prepare(tgp_ctx* h) {
    TGP_CUDA(h, cudaFuncSetAttribute(tc::k_tc_gemm_tn<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<64>::kSmem));
    TGP_CUDA(h, cudaFuncSetAttribute(tc::k_tc_gemm_tn<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<128>::kSmem));
    return TGP_OK;
}

// C = alpha X' Y (+ additive terms). Y may continue into a second tensor Y2 from contraction index k_switch on.
// Tile 128 x 64 for the D x D / D x M products of one step (more CTAs); 128 x 128 when C is wide (the time-blocked phase):
// the operand stream from L2 is what bounds these kernels, and the wider tile does 1.5x the MMAs per byte loaded.
static int tc_gemm(tgp_ctx* h, const char* name, const TcOp& X, const TcOp& Y, int K, const tc::Epi& e, const TcOp* Y2 = nullptr, int k_switch_elems = 0,
                   int y_col_shift = 0) {
    tc::Src src;
    src.K = K; src.lo_col_x = X.pr.cpad; src.lo_col_y = Y.pr.cpad; src.y_col_shift = y_col_shift;
    if (Y2) { src.lo_col_y2 = Y2->pr.cpad; src.k_switch = k_switch_elems / tc::BK; }
    const bool wide = e.N >= 1024 && !e.symmetric;
    const int BN = wide ? 128 : 64;
    dim3 grid((e.Mx + tc::BM - 1) / tc::BM, (e.N + BN - 1) / BN);
    TGP_K(h, name);
    if (wide) tc::k_tc_gemm_tn<128><<<grid, tc::kThreads, tc::Cfg<128>::kSmem, h->stream>>>(X.m128, Y.m128, Y2 ? Y2->m128 : Y.m128, src, e);
    else      tc::k_tc_gemm_tn<64><<<grid, tc::kThreads, tc::Cfg<64>::kSmem, h->stream>>>(X.m128, Y.m64, Y2 ? Y2->m64 : Y.m64, src, e);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

// double column-major (rows x cols, leading dimension rows) -> pair; transpose != 0 writes the transpose.
__global__ void k_to_pair(const double* __restrict__ src, int rows, int cols, int transpose, Pair dst) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)rows * cols; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % rows), j = (int)(e / rows);
        const float x = (float)src[e];
        const float hi = tc::tf32_hi(x);
        const size_t o = transpose ? (size_t)j + (size_t)dst.ld * i : (size_t)i + (size_t)dst.ld * j;
        dst.hi()[o] = hi;
        dst.lo()[o] = x - hi;
    }
}

// out[n] = b1[t s1 + n] - b2[t s2 + n] + sign * sum_k X[k, n] v[k]   (t = *step; b1 / b2 nullable). One warp per n.
__global__ void __launch_bounds__(256) k_gemv_pair(Pair X, int K, int N, const double* __restrict__ v, const double* __restrict__ b1, long long s1,
                                                   const double* __restrict__ b2, long long s2, double sign, double* __restrict__ out,
                                                   const long long* __restrict__ step) {
    const int lane = threadIdx.x & 31;
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const float* hi = X.hi() + (size_t)X.ld * n;
    const float* lo = X.lo() + (size_t)X.ld * n;
    double acc = 0.0;
    for (int k = lane; k < K; k += 32) acc = fma((double)hi[k] + (double)lo[k], v[k], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        const long long t = *step;
        double r = sign * acc;
        if (b1) r += b1[t * s1 + n];
        if (b2) r -= b2[t * s2 + n];
        out[n] = r;
    }
}

// Winv = U^-1 from the inverses of the 32 x 32 diagonal blocks (Dinv, written by k_chol_panel2): blocked back substitution,
// one warp per COLUMN c of Winv, lane = row inside the current block row I:
//     x_J = Dinv_J[:, c - 32 J];    x_I = -Dinv_I (sum_{l > 32 I + 31}^{c} U[32 I + lane, l] x_l),  I = J-1 .. 0.
// Rows of U are read coalesced across the lanes; 8 block steps at M = 256 instead of 256 scalar ones.
__global__ void __launch_bounds__(256) k_tri_inv2(const double* __restrict__ U, int M, const double* __restrict__ Dinv, Pair Winv, Pair WinvT) {
    extern __shared__ double xbuf[];                 // 8 warps x (Mp + 32), then the staged block row of U: 32 x Mp
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
        for (int e = tid; e < 32 * nl; e += 256) tile[e] = __ldg(U + (size_t)(32 * I + (e & 31)) + (size_t)M * (l0 + (e >> 5)));
        __syncthreads();
        if (live) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            int l = l0;
            for (; l + 3 <= c; l += 4) {
                const double* tp = tile + (size_t)(l - l0) * 32 + lane;
                t0 = fma(tp[0], x[l], t0);      t1 = fma(tp[32], x[l + 1], t1);
                t2 = fma(tp[64], x[l + 2], t2); t3 = fma(tp[96], x[l + 3], t3);
            }
            for (; l <= c; ++l) t0 = fma(tile[(size_t)(l - l0) * 32 + lane], x[l], t0);
            tb[lane] = (t0 + t1) + (t2 + t3);
            __syncwarp();
            const double* di = Dinv + (size_t)I * 1024 + lane;     // row `lane` of the upper-triangular block inverse
            double s = 0.0;
#pragma unroll 8
            for (int rp = 0; rp < 32; ++rp) s = fma(__ldg(di + 32 * rp), tb[rp], s);
            x[32 * I + lane] = -s;
            __syncwarp();
        }
    }
    if (!live) return;
    for (int i = lane; i <= c; i += 32) {
        const float xv = (float)x[i];
        const float hi = tc::tf32_hi(xv);
        Winv.hi()[(size_t)i + (size_t)Winv.ld * c] = hi;
        Winv.lo()[(size_t)i + (size_t)Winv.ld * c] = xv - hi;
        WinvT.hi()[(size_t)c + (size_t)WinvT.ld * i] = hi;
        WinvT.lo()[(size_t)c + (size_t)WinvT.ld * i] = xv - hi;
    }
}

__global__ void k_emit_state_pair(const double* __restrict__ m, Pair P, int D, double* __restrict__ m_f, long long s_m, double* __restrict__ P_f,
                                  long long s_P, const long long* __restrict__ step) {
    const long long t = *step;
    const long long n = (long long)D * D;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n + D; e += (long long)gridDim.x * blockDim.x) {
        if (e < n) {
            if (P_f) { const size_t o = (size_t)(e % D) + (size_t)P.ld * (e / D); P_f[t * s_P + e] = (double)P.hi()[o] + (double)P.lo()[o]; }
        } else if (m_f) m_f[t * s_m + (e - n)] = m[e - n];
    }
}

// End of a step: steady-state test on what the covariance epilogue measured (conv[0] = max |P_t - P_{t-1}|, conv[1] = max |P_t|),
// then advance the device-side step counter. ss_at: -1 until the covariance recursion has converged, then the step index.
__global__ void k_advance_conv(long long* step, long long delta, unsigned* conv, float tol, long long* ss_at) {
    if (conv) {
        const float d = __uint_as_float(conv[0]), a = __uint_as_float(conv[1]);
        if (*ss_at < 0 && a > 0.f && d <= tol * a) *ss_at = *step;
        conv[0] = 0u;
        conv[1] = 0u;
    }
    *step += delta;
}

struct TcWs {
    TcOp At, Ht, Pa, Pb, W, Vt, V, B, Winv, WinvT;
    double *S, *m, *mp, *r, *alpha, *lml, *Dinv;
    long long* step;
    unsigned long long* err;
    unsigned* conv = nullptr;      // steady-state detection (time-invariant models only)
    long long* ss_at = nullptr;
};

// One step. cur / nxt: covariance ping-pong (Pa, Pb); on return *cur holds the filtering covariance.
// frozen: the covariance recursion has reached its fixed point (time-invariant model): P, S, U, Winv, B stay as the
// last full step left them and only the mean / likelihood part of the step runs (same arithmetic, frozen gain).
static int tc_step(tgp_ctx* h, const tgp_lgssm& d, const double* dy, TcWs& w, long long t, bool graph_mode, double* lml_steps, double* m_f,
                   int64_t s_m, double* P_f, int64_t s_P, bool frozen = false) {
    const int D = d.D, M = d.M;
    cudaStream_t st = h->stream;
    const long long tt = graph_mode ? 0 : t;
    const bool rev = d.ordering == TGP_REVERSE;
    auto gemv = [&](const char* name, const Pair& X, int K, int N, const double* v, const double* b1, long long s1, const double* b2, long long s2,
                    double sign, double* out) -> int {
        TGP_K(h, name);
        k_gemv_pair<<<(N * 32 + 255) / 256, 256, 0, st>>>(X, K, N, v, b1, s1, b2, s2, sign, out, w.step);
        TGP_LAUNCH_CHECK(h);
        return TGP_OK;
    };
    if (!graph_mode) {   // time-varying parameters: refresh A', H' for this step
        const int nb = (int)std::min<long long>(((long long)D * D + 255) / 256, 1184);
        if (d.sA) { TGP_K(h, "tc:k_to_pair"); k_to_pair<<<nb, 256, 0, st>>>(d.A + tt * d.sA, D, D, 1, w.At.pr); TGP_LAUNCH_CHECK(h); }
        if (d.sH) { TGP_K(h, "tc:k_to_pair"); k_to_pair<<<nb, 256, 0, st>>>(d.H + tt * d.sH, M, D, 1, w.Ht.pr); TGP_LAUNCH_CHECK(h); }
    }
    TcOp* P = &w.Pa;      // current covariance
    TcOp* Pn = &w.Pb;     // scratch for the next one
    auto predict = [&]() -> int {
        TGP_TRY(gemv("tc:k_gemv_pair(predict mean)", w.At.pr, D, D, w.m, d.a, d.sa, nullptr, 0, 1.0, w.mp));
        TGP_CUDA(h, cudaMemcpyAsync(w.m, w.mp, sizeof(double) * D, cudaMemcpyDeviceToDevice, st));
        if (frozen) return TGP_OK;
        tc::Epi e1;
        e1.Mx = D; e1.N = D; e1.out_hi = w.W.pr.hi(); e1.out_lo = w.W.pr.lo(); e1.ld_out = w.W.pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm W=P'At", *P, w.At, D, e1));
        tc::Epi e2;
        e2.Mx = D; e2.N = D; e2.symmetric = 1; e2.cin_d = d.Q + tt * d.sQ; e2.ld_cind = D;
        e2.out_hi = Pn->pr.hi(); e2.out_lo = Pn->pr.lo(); e2.ld_out = Pn->pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm Pp=W'At+Q", w.W, w.At, D, e2));
        std::swap(P, Pn);
        return TGP_OK;
    };
    auto update_mean = [&]() -> int {   // residual, whitened residual, likelihood, mean
        TGP_TRY(gemv("tc:k_gemv_pair(residual)", w.Ht.pr, D, M, w.m, dy, M, d.h, d.sh, -1.0, w.r));
        TGP_TRY(gemv("tc:k_gemv_pair(alpha)", w.Winv.pr, M, M, w.r, nullptr, 0, nullptr, 0, 1.0, w.alpha));
        TGP_K(h, "dense:k_lml");
        k_lml<<<1, 256, 0, st>>>(w.S, w.alpha, M, lml_steps, w.lml, w.step);
        TGP_LAUNCH_CHECK(h);
        TGP_TRY(gemv("tc:k_gemv_pair(mean)", w.B.pr, M, D, w.alpha, w.m, 0, nullptr, 0, 1.0, w.mp));
        TGP_CUDA(h, cudaMemcpyAsync(w.m, w.mp, sizeof(double) * D, cudaMemcpyDeviceToDevice, st));
        if (m_f || P_f) {
            const int nb = (int)std::min<long long>(((long long)D * D + D + 255) / 256, 1184);
            TGP_K(h, "tc:k_emit_state_pair");
            k_emit_state_pair<<<nb, 256, 0, st>>>(w.m, (rev ? w.Pb : w.Pa).pr, D, m_f, s_m, P_f, s_P, w.step);
            TGP_LAUNCH_CHECK(h);
        }
        return TGP_OK;
    };
    auto update = [&]() -> int {
        if (frozen) return update_mean();
        tc::Epi e3;
        e3.Mx = D; e3.N = M; e3.out_hi = w.Vt.pr.hi(); e3.out_lo = w.Vt.pr.lo(); e3.ld_out = w.Vt.pr.ld;
        e3.outT_hi = w.V.pr.hi(); e3.outT_lo = w.V.pr.lo(); e3.ld_outT = w.V.pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm Vt=P'Ht", *P, w.Ht, D, e3));
        tc::Epi e4;
        e4.Mx = M; e4.N = M; e4.symmetric = 1; e4.out_d = w.S; e4.ld_outd = M;
        if (d.R_kind == TGP_R_DENSE) { e4.cin_d = d.R + tt * d.sR; e4.ld_cind = M; }
        else { e4.cin_diag = d.R + tt * d.sR; e4.diag_stride = d.R_kind == TGP_R_DIAG ? 1 : 0; }
        TGP_TRY(tc_gemm(h, "tc:gemm S=Vt'Ht+R", w.Vt, w.Ht, D, e4));
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
        TGP_K(h, "tc:k_tri_inv2");
        k_tri_inv2<<<(M + 7) / 8, 256, sizeof(double) * (8 * ((M + 31) / 32 * 32 + 32) + 32 * ((M + 31) / 32 * 32)), st>>>(w.S, M, w.Dinv, w.Winv.pr, w.WinvT.pr);
        TGP_LAUNCH_CHECK(h);
        tc::Epi e5;
        e5.Mx = M; e5.N = D; e5.out_hi = w.B.pr.hi(); e5.out_lo = w.B.pr.lo(); e5.ld_out = w.B.pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm B=Winv'V", w.Winv, w.V, M, e5));
        tc::Epi e6;
        e6.Mx = D; e6.N = D; e6.symmetric = 1; e6.alpha = -1.f; e6.cin_hi = P->pr.hi(); e6.cin_lo = P->pr.lo(); e6.ld_cin = P->pr.ld;
        e6.out_hi = Pn->pr.hi(); e6.out_lo = Pn->pr.lo(); e6.ld_out = Pn->pr.ld;
        e6.conv = w.conv;
        TGP_TRY(tc_gemm(h, "tc:gemm P=Pp-B'B", w.B, w.B, M, e6));
        std::swap(P, Pn);
        return update_mean();
    };
    if (!rev) { TGP_TRY(predict()); TGP_TRY(update()); }
    else      { TGP_TRY(update()); TGP_TRY(predict()); }
    // two swaps per step: the filtering / predicted covariance is back in w.Pa
    TGP_K(h, "dense:k_advance");
    k_advance_conv<<<1, 1, 0, st>>>(w.step, rev ? -1 : 1, frozen ? nullptr : w.conv, fmaxf((float)h->ss_tol, 2e-6f), w.ss_at);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Time-blocked steady phase (Forward ordering, log-likelihood only). Once P, S, U, B are frozen the step is the constant
// affine map of z_t = [m_t; alpha_t]:
//     [m_t; alpha_t] = G [m_{t-1}; y_t] + g,    G = [[Abar, Kbar], [W1, W2]],
//     W2 = Winv', W1 = -Winv' H A, Kbar = B' Winv', Abar = A + B' W1, g = [a + B' w0; w0], w0 = -Winv'(h + H a)
// (the same arithmetic as the mean part of the step, lml_t = -(M log 2pi + logdet S + alpha_t'alpha_t)/2). The n remaining
// steps are cut into nb blocks of kTbL consecutive steps and ALL blocks advance together, one position j per launch:
//     pass 1   S_{j+1} = [Abar Kbar] [S_j; Y_j] + c      S: D x nb (one column per block), Y_j: M x nb (a strided view of y)
//              from zero states -> the zero-state response of every block;
//     prefix   X_b = Abar^L X_{b-1} + Z_b by recursive doubling over the blocks (log2 nb products with Abar^(L 2^k));
//     pass 2   the same recursion from the true block-initial states, the alpha rows going straight into per-step sums of squares.
// Every product is a (D + M) x (D + M) x nb contraction on the tensor cores (k_tc_gemm_tn): the sequential chain of T GEMVs
// becomes ~2 kTbL + 2 log2(nb) large GEMMs.
constexpr int kTbL = 64;

__global__ void k_y_to_pair(const double* __restrict__ y, long long n, int M, Pair Y) {   // column t of Y = y_t (t < n)
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n * M; e += (long long)gridDim.x * blockDim.x) {
        const long long t = e / M;
        const int i = (int)(e % M);
        const float x = (float)y[e];
        const float hi = tc::tf32_hi(x);
        Y.hi()[(size_t)i + (size_t)Y.ld * t] = hi;
        Y.lo()[(size_t)i + (size_t)Y.ld * t] = x - hi;
    }
}
// dst[r0 + i, c0 + j] = src[i, j]
__global__ void k_copy_pair_block(Pair src, int rows, int cols, Pair dst, int r0, int c0) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)rows * cols; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % rows), j = (int)(e / rows);
        const size_t so = (size_t)i + (size_t)src.ld * j, dn = (size_t)(r0 + i) + (size_t)dst.ld * (c0 + j);
        dst.hi()[dn] = src.hi()[so];
        dst.lo()[dn] = src.lo()[so];
    }
}
// dst[:, 0] = v (double vector), dst[:, b] = src[:, b - 1] for 1 <= b < nb (src nullable: only column 0 is set)
__global__ void k_block_states(const double* __restrict__ v, Pair src, int D, long long nb, Pair dst) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)D * nb; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % D);
        const long long b = e / D;
        float x;
        if (b == 0) x = (float)v[i];
        else if (src.p) x = src.hi()[(size_t)i + (size_t)src.ld * (b - 1)] + src.lo()[(size_t)i + (size_t)src.ld * (b - 1)];
        else continue;
        const float hi = tc::tf32_hi(x);
        dst.hi()[(size_t)i + (size_t)dst.ld * b] = hi;
        dst.lo()[(size_t)i + (size_t)dst.ld * b] = x - hi;
    }
}
__global__ void k_make_bias(const double* __restrict__ c, int D, const double* __restrict__ w0, int M, float* __restrict__ bias) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D + M; i += gridDim.x * blockDim.x) bias[i] = (float)(i < D ? c[i] : w0[i - D]);
}
// lml_t = -(M log 2pi + 2 sum log U_ii + csq_t)/2 for the n blocked steps; total added to *lml_total. One CTA.
__global__ void __launch_bounds__(1024) k_lml_blocked(const double* __restrict__ U, int M, const double* __restrict__ csq, long long n,
                                                      double* __restrict__ lml_steps, double* __restrict__ lml_total) {
    __shared__ double sh[1024];
    __shared__ double s_ld;
    double s = 0.0;
    for (int i = threadIdx.x; i < M; i += 1024) s += 2.0 * log(U[i + (size_t)M * i]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) { if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off]; __syncthreads(); }
    if (threadIdx.x == 0) s_ld = sh[0];
    __syncthreads();
    const double c0 = M * kLog2PiD + s_ld;
    double acc = 0.0;
    for (long long t = threadIdx.x; t < n; t += 1024) {
        const double l = -0.5 * (c0 + csq[t]);
        if (lml_steps) lml_steps[t] = l;
        acc += l;
    }
    __syncthreads();
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) { if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off]; __syncthreads(); }
    if (threadIdx.x == 0) *lml_total += sh[0];
}

static tc::Epi epi_out(int Mx, int N, const TcOp& out) {
    tc::Epi e;
    e.Mx = Mx; e.N = N; e.out_hi = out.pr.hi(); e.out_lo = out.pr.lo(); e.ld_out = out.pr.ld;
    return e;
}

// t0: index of the first blocked step, n: number of steps. w.m holds the filtered mean of step t0 - 1; w.S, w.B, w.Winv(T) are frozen.
static int tc_steady_blocked(tgp_ctx* h, const tgp_lgssm& d, const double* dy, TcWs& w, int64_t t0, int64_t n, double* lml_steps) {
    const int D = d.D, M = d.M, L = kTbL;
    const int Dp = (D + 31) / 32 * 32;                    // the state part of the stacked contraction index, padded to whole k-blocks
    const long long nb = (n + L - 1) / L;
    cudaStream_t st = h->stream;
    auto grid1d = [](long long work) { return (int)std::min<long long>((work + 255) / 256, 2368); };

    // ---- constants ---------------------------------------------------------------------------------------------------------
    TcOp Ap, HA, W1, Ab, Abt, Kb, Gt, P2, Pt2;
    TGP_TRY(tc_make_op(h, D, D, &Ap));
    TGP_TRY(tc_make_op(h, M, D, &HA));
    TGP_TRY(tc_make_op(h, M, D, &W1));
    TGP_TRY(tc_make_op(h, D, D, &Ab));
    TGP_TRY(tc_make_op(h, D, D, &Abt));
    TGP_TRY(tc_make_op(h, D, M, &Kb));
    TGP_TRY(tc_make_op(h, Dp + M, D + M, &Gt));
    TGP_TRY(tc_make_op(h, D, D, &P2));
    TGP_TRY(tc_make_op(h, D, D, &Pt2));
    TGP_K(h, "tc:k_to_pair");
    k_to_pair<<<grid1d((long long)D * D), 256, 0, st>>>(d.A, D, D, 0, Ap.pr);
    TGP_LAUNCH_CHECK(h);
    TGP_TRY(tc_gemm(h, "tc:gemm HA=Ht'A", w.Ht, Ap, D, epi_out(M, D, HA)));
    {   // W1 = -Winv' HA, and W1' into G' (rows = state in, columns D.. = alpha out)
        tc::Epi e = epi_out(M, D, W1);
        e.alpha = -1.f;
        e.outT_hi = Gt.pr.hi() + (size_t)Gt.pr.ld * D; e.outT_lo = Gt.pr.lo() + (size_t)Gt.pr.ld * D; e.ld_outT = Gt.pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm W1=-Winv'HA", w.Winv, HA, M, e));
    }
    {   // Abar = A + B' W1 (and its transpose)
        tc::Epi e = epi_out(D, D, Ab);
        e.cin_hi = Ap.pr.hi(); e.cin_lo = Ap.pr.lo(); e.ld_cin = Ap.pr.ld;
        e.outT_hi = Abt.pr.hi(); e.outT_lo = Abt.pr.lo(); e.ld_outT = Abt.pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm Abar=A+B'W1", w.B, W1, M, e));
    }
    {   // Kbar = B' Winv', and Kbar' into G' (rows Dp.. = observation in, columns 0..D = state out)
        tc::Epi e = epi_out(D, M, Kb);
        e.outT_hi = Gt.pr.hi() + Dp; e.outT_lo = Gt.pr.lo() + Dp; e.ld_outT = Gt.pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm Kbar=B'Winv'", w.B, w.WinvT, M, e));
    }
    TGP_K(h, "tc:k_copy_pair_block");
    k_copy_pair_block<<<grid1d((long long)D * D), 256, 0, st>>>(Abt.pr, D, D, Gt.pr, 0, 0);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "tc:k_copy_pair_block");
    k_copy_pair_block<<<grid1d((long long)M * M), 256, 0, st>>>(w.Winv.pr, M, M, Gt.pr, Dp, D);
    TGP_LAUNCH_CHECK(h);
    double *t1, *w0, *cbar, *csq;
    float* bias;
    TGP_TRY(dalloc(h, M, &t1));
    TGP_TRY(dalloc(h, M, &w0));
    TGP_TRY(dalloc(h, D, &cbar));
    TGP_TRY(dalloc(h, D + M, &bias));
    TGP_TRY(dalloc(h, (size_t)nb * L, &csq));
    TGP_CUDA(h, cudaMemsetAsync(csq, 0, sizeof(double) * nb * L, st));
    auto gemv = [&](const Pair& X, int K, int N, const double* v, const double* b1, double sign, double* out) -> int {
        TGP_K(h, "tc:k_gemv_pair(setup)");
        k_gemv_pair<<<(N * 32 + 255) / 256, 256, 0, st>>>(X, K, N, v, b1, 0, nullptr, 0, sign, out, w.step);
        TGP_LAUNCH_CHECK(h);
        return TGP_OK;
    };
    TGP_TRY(gemv(w.Ht.pr, D, M, d.a, d.h, 1.0, t1));          // t1 = h + H a
    TGP_TRY(gemv(w.Winv.pr, M, M, t1, nullptr, -1.0, w0));    // w0 = -Winv' t1
    TGP_TRY(gemv(w.B.pr, M, D, w0, d.a, 1.0, cbar));          // c = a + B' w0
    TGP_K(h, "tc:k_make_bias");
    k_make_bias<<<(D + M + 255) / 256, 256, 0, st>>>(cbar, D, w0, M, bias);
    TGP_LAUNCH_CHECK(h);

    // ---- observations as a pair tensor and its kTbL strided views (position j of every block) ------------------------------
    const long long nbp = (nb + 127) / 128 * 128;
    Pair Yp;
    Yp.rows = M; Yp.cols = (int)std::min<long long>(nbp * L, 0x7fffffff); Yp.ld = (M + 31) / 32 * 32; Yp.cpad = (int)(nbp * L);
    if (nbp * L > 0x7fffffffLL) return fail(h, TGP_EUNSUPPORTED, "series too long for the time-blocked steady phase");
    TGP_TRY(dalloc(h, Yp.floats(), &Yp.p));
    TGP_CUDA(h, cudaMemsetAsync(Yp.p, 0, Yp.floats() * sizeof(float), st));
    TGP_K(h, "tc:k_y_to_pair");
    k_y_to_pair<<<grid1d(n * M), 256, 0, st>>>(dy + t0 * M, n, M, Yp);
    TGP_LAUNCH_CHECK(h);
    std::vector<TcOp> yv(L);
    for (int j = 0; j < L; ++j) {
        yv[j].pr = Yp;
        yv[j].pr.cpad = (int)nbp;        // lo plane of the view: nbp view-columns behind the hi plane
        TGP_TRY(tc_encode_map(h, Yp.p + (size_t)Yp.ld * j, M, 2 * nbp, (size_t)Yp.ld * L * sizeof(float), 64u, &yv[j].m64));
        TGP_TRY(tc_encode_map(h, Yp.p + (size_t)Yp.ld * j, M, 2 * nbp, (size_t)Yp.ld * L * sizeof(float), 128u, &yv[j].m128));
    }

    // ---- pass 1: zero-state response of every block (block 0 starts from the current mean) ---------------------------------
    TcOp Sa, Sb, Xa, Xb;
    TGP_TRY(tc_make_op(h, D, (int)nb, &Sa));
    TGP_TRY(tc_make_op(h, D, (int)nb, &Sb));
    TGP_TRY(tc_make_op(h, D, (int)nb, &Xa, (int)nb));
    TGP_TRY(tc_make_op(h, D, (int)nb, &Xb, (int)nb));
    Pair none;
    TGP_K(h, "tc:k_block_states");
    k_block_states<<<grid1d(D), 256, 0, st>>>(w.m, none, D, 1, Sa.pr);
    TGP_LAUNCH_CHECK(h);
    TcOp* Sc = &Sa;
    TcOp* Sn = &Sb;
    for (int j = 0; j < L; ++j) {
        tc::Epi e = epi_out(D, (int)nb, j == L - 1 ? Xa : *Sn);
        e.row_bias = bias;
        TGP_TRY(tc_gemm(h, "tc:gemm blocked pass 1", Gt, *Sc, Dp + M, e, &yv[j], Dp));
        std::swap(Sc, Sn);
    }
    // ---- prefix over blocks by recursive doubling: X_b += P_k X_{b - 2^k}, P_0 = Abar^L, P_{k+1} = P_k^2 -------------------------
    TcOp* P = &Ab;      // (P, Pt) -> (P^2, (P^2)')
    TcOp* Pt = &Abt;
    TcOp* Q = &P2;
    TcOp* Qt = &Pt2;
    auto square = [&]() -> int {
        tc::Epi e = epi_out(D, D, *Qt);
        e.outT_hi = Q->pr.hi(); e.outT_lo = Q->pr.lo(); e.ld_outT = Q->pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm power", *P, *Pt, D, e));
        std::swap(P, Q);
        std::swap(Pt, Qt);
        return TGP_OK;
    };
    // (Abar / Abar' may be overwritten by the squarings: G' holds its own copy)
    for (int s2 = 1; s2 < L; s2 <<= 1) TGP_TRY(square());
    TcOp* Xc = &Xa;
    TcOp* Xn = &Xb;
    for (long long sh = 1; sh < nb; sh <<= 1) {
        tc::Epi e = epi_out(D, (int)nb, *Xn);
        e.cin_hi = Xc->pr.hi(); e.cin_lo = Xc->pr.lo(); e.ld_cin = Xc->pr.ld;
        TGP_TRY(tc_gemm(h, "tc:gemm blocked prefix", *Pt, *Xc, D, e, nullptr, 0, (int)-sh));
        std::swap(Xc, Xn);
        if ((sh << 1) < nb) TGP_TRY(square());
    }
    // ---- pass 2: from the true block-initial states; alpha rows -> per-step sums of squares ---------------------------------
    TGP_K(h, "tc:k_block_states");
    k_block_states<<<grid1d((long long)D * nb), 256, 0, st>>>(w.m, Xc->pr, D, nb, Sa.pr);
    TGP_LAUNCH_CHECK(h);
    Sc = &Sa;
    Sn = &Sb;
    for (int j = 0; j < L && j < n; ++j) {
        const long long nvalid = (n - j + L - 1) / L;      // blocks that have a step at position j
        tc::Epi e = epi_out(D + M, (int)nvalid, *Sn);
        e.row_bias = bias;
        e.split_row = D;
        e.colsq = csq + j;
        e.colsq_stride = L;
        TGP_TRY(tc_gemm(h, "tc:gemm blocked pass 2", Gt, *Sc, Dp + M, e, &yv[j], Dp));
        std::swap(Sc, Sn);
    }
    TGP_K(h, "tc:k_lml_blocked");
    k_lml_blocked<<<1, 1024, 0, st>>>(w.S, M, csq, n, lml_steps ? lml_steps + t0 : nullptr, w.lml);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

int dense_filter_tc(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* lml_out, double* lml_steps_user, double* m_f_user, int64_t s_m,
                    double* P_f_user, int64_t s_P) {
    const int D = m->D, M = m->M;
    const int64_t T = m->T;
    cudaStream_t st = h->stream;
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

    TGP_TRY(tc_prepare(h));
    TcWs w;
    TGP_TRY(tc_make_op(h, D, D, &w.At));
    TGP_TRY(tc_make_op(h, D, M, &w.Ht));
    TGP_TRY(tc_make_op(h, D, D, &w.Pa));
    TGP_TRY(tc_make_op(h, D, D, &w.Pb));
    TGP_TRY(tc_make_op(h, D, D, &w.W));
    TGP_TRY(tc_make_op(h, D, M, &w.Vt));
    TGP_TRY(tc_make_op(h, M, D, &w.V));
    TGP_TRY(tc_make_op(h, M, D, &w.B));
    TGP_TRY(tc_make_op(h, M, M, &w.Winv));
    TGP_TRY(tc_make_op(h, M, M, &w.WinvT));
    TGP_TRY(dalloc(h, (size_t)M * M, &w.S));
    TGP_TRY(dalloc(h, (size_t)((M + 31) / 32) * 1024, &w.Dinv));
    TGP_TRY(dalloc(h, D, &w.m));
    TGP_TRY(dalloc(h, D, &w.mp));
    TGP_TRY(dalloc(h, M, &w.r));
    TGP_TRY(dalloc(h, M, &w.alpha));
    TGP_TRY(dalloc(h, 1, &w.lml));
    TGP_TRY(dalloc(h, 1, &w.step));
    TGP_TRY(dalloc(h, 1, &w.err));
    const bool ti = !(m->sA | m->sa | m->sQ | m->sH | m->sh | m->sR);
    if (ti && h->algo == TGP_ALGO_AUTO) {
        TGP_TRY(dalloc(h, 2, &w.conv));
        TGP_TRY(dalloc(h, 1, &w.ss_at));
        TGP_CUDA(h, cudaMemsetAsync(w.conv, 0, 2 * sizeof(unsigned), st));
        TGP_CUDA(h, cudaMemsetAsync(w.ss_at, 0xFF, sizeof(long long), st));
    }
    const int nb = (int)std::min<long long>(((long long)D * D + 255) / 256, 1184);
    TGP_K(h, "tc:k_to_pair");
    k_to_pair<<<nb, 256, 0, st>>>(d.A, D, D, 1, w.At.pr);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "tc:k_to_pair");
    k_to_pair<<<nb, 256, 0, st>>>(d.H, M, D, 1, w.Ht.pr);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "tc:k_to_pair");
    k_to_pair<<<nb, 256, 0, st>>>(d.P0, D, D, 0, w.Pa.pr);
    TGP_LAUNCH_CHECK(h);
    TGP_CUDA(h, cudaMemcpyAsync(w.m, d.m0, sizeof(double) * D, cudaMemcpyDeviceToDevice, st));
    TGP_CUDA(h, cudaMemsetAsync(w.lml, 0, sizeof(double), st));
    TGP_CUDA(h, cudaMemsetAsync(w.err, 0xFF, sizeof(unsigned long long), st));
    const bool rev = m->ordering == TGP_REVERSE;
    long long* pt0 = (long long*)(h->pinned + 8);
    *pt0 = rev ? T - 1 : 0;
    TGP_CUDA(h, cudaMemcpyAsync(w.step, pt0, sizeof(long long), cudaMemcpyHostToDevice, st));

    {
        const size_t tb = sizeof(double) * (8 * ((M + 31) / 32 * 32 + 32) + 32 * ((M + 31) / 32 * 32));
        if (tb > 200 * 1024) return fail(h, TGP_EUNSUPPORTED, "observation dimension M=%d too large for the triangular-inverse kernel", M);
        if (tb > 48 * 1024) TGP_CUDA(h, cudaFuncSetAttribute(k_tri_inv2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb));
    }
    if (ti && T >= 8 && !h->timing) {
        // One step captured into a CUDA graph and replayed. Every kPoll (4) steps the host looks at the steady-state word;
        // once the covariance recursion has converged the remaining steps replay the mean-only graph.
        cudaGraph_t graph[2] = {nullptr, nullptr};
        cudaGraphExec_t exec[2] = {nullptr, nullptr};
        int64_t per_replay[2] = {0, 0};
        TGP_CUDA(h, cudaStreamSynchronize(st));
        for (int fz = 0; fz < (w.conv ? 2 : 1); ++fz) {
            TGP_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int64_t l0 = h->launches;
            int rc = tc_step(h, d, dy, w, -1, true, lml_steps, m_f, dsm, P_f, dsP, fz == 1);
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
                // log-likelihood only, Forward ordering, enough steps left: the rest of the series advances block-parallel
                if (frozen && !rev && !m_f && !P_f && T - (t + 1) >= 8 * kTbL && h->chunk != 1) {
                    TGP_TRY(tc_steady_blocked(h, d, dy, w, t + 1, T - (t + 1), lml_steps));
                    break;
                }
            }
        }
        TGP_CUDA(h, cudaStreamSynchronize(st));
        for (int fz = 0; fz < 2; ++fz) {
            if (exec[fz]) cudaGraphExecDestroy(exec[fz]);
            if (graph[fz]) cudaGraphDestroy(graph[fz]);
        }
    } else {
        for (int64_t n = 0; n < T; ++n) {
            const long long t = rev ? T - 1 - n : n;
            TGP_TRY(tc_step(h, d, dy, w, t, false, lml_steps, m_f, dsm, P_f, dsP));
        }
    }
    unsigned long long* perr = (unsigned long long*)h->pinned;
    TGP_CUDA(h, cudaMemcpyAsync(perr, w.err, 8, cudaMemcpyDeviceToHost, st));
    h->d2h += 8;
    TGP_TRY(deliver_scalar(h, w.lml, lml_out));
    TGP_TRY(flush_outputs(h));
    TGP_CUDA(h, cudaStreamSynchronize(st));
    if (*perr != ~0ull) return fail(h, TGP_ENOTPD, "innovation covariance not positive definite at time index %lld (0-based)", (long long)*perr);
    return TGP_OK;
}

// Test hook: C (Mx x N, column-major float) = X' Y for host X (K x Mx), Y (K x N) column-major float, through the
// tensor-core kernel alone (tests/test_gpu_parity.py compares with NumPy).
int tc_gemm_selftest(tgp_ctx* h, int K, int Mx, int N, const float* X, const float* Y, float* C, int symmetric) {
    cudaStream_t st = h->stream;
    TGP_TRY(tc_prepare(h));
    TcOp ox, oy, oc;
    TGP_TRY(tc_make_op(h, K, Mx, &ox));
    TGP_TRY(tc_make_op(h, K, N, &oy));
    TGP_TRY(tc_make_op(h, Mx, N, &oc));
    double *dx, *dyv;
    TGP_TRY(dalloc(h, (size_t)K * Mx, &dx));
    TGP_TRY(dalloc(h, (size_t)K * N, &dyv));
    std::vector<double> hx((size_t)K * Mx), hy((size_t)K * N);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = X[i];
    for (size_t i = 0; i < hy.size(); ++i) hy[i] = Y[i];
    TGP_CUDA(h, cudaMemcpyAsync(dx, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice, st));
    TGP_CUDA(h, cudaMemcpyAsync(dyv, hy.data(), hy.size() * 8, cudaMemcpyHostToDevice, st));
    TGP_K(h, "tc:k_to_pair");
    k_to_pair<<<256, 256, 0, st>>>(dx, K, Mx, 0, ox.pr);
    TGP_LAUNCH_CHECK(h);
    TGP_K(h, "tc:k_to_pair");
    k_to_pair<<<256, 256, 0, st>>>(dyv, K, N, 0, oy.pr);
    TGP_LAUNCH_CHECK(h);
    tc::Epi e;
    e.Mx = Mx; e.N = N; e.symmetric = symmetric;
    e.out_hi = oc.pr.hi(); e.out_lo = oc.pr.lo(); e.ld_out = oc.pr.ld;
    TGP_TRY(tc_gemm(h, "tc:gemm selftest", ox, oy, K, e));
    std::vector<float> hh((size_t)oc.pr.ld * N), hl((size_t)oc.pr.ld * N);
    TGP_CUDA(h, cudaMemcpyAsync(hh.data(), oc.pr.hi(), hh.size() * 4, cudaMemcpyDeviceToHost, st));
    TGP_CUDA(h, cudaMemcpyAsync(hl.data(), oc.pr.lo(), hl.size() * 4, cudaMemcpyDeviceToHost, st));
    TGP_CUDA(h, cudaStreamSynchronize(st));
    for (int n = 0; n < N; ++n)
        for (int i = 0; i < Mx; ++i) C[(size_t)i + (size_t)Mx * n] = hh[(size_t)i + (size_t)oc.pr.ld * n] + hl[(size_t)i + (size_t)oc.pr.ld * n];
    return TGP_OK;
}

}  // namespace tgp

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
//                              U  = chol(S) in FP64 (k_chol_panel / k_chol_trail),  Winv = U^-1  (k_tri_inv)
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

static int tc_make_op(tgp_ctx* h, int rows, int cols, TcOp* op) {
    Pair& p = op->pr;
    p.rows = rows; p.cols = cols;
    p.ld = (rows + 31) / 32 * 32;
    p.cpad = (cols + 127) / 128 * 128;
    TGP_TRY(dalloc(h, p.floats(), &p.p));
    TGP_CUDA(h, cudaMemsetAsync(p.p, 0, p.floats() * sizeof(float), h->stream));
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return fail(h, TGP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)(2 * p.cpad)};
    const cuuint64_t strides[1] = {(cuuint64_t)p.ld * sizeof(float)};
    const cuuint32_t es[2] = {1, 1};
    for (int which = 0; which < 2; ++which) {
        const cuuint32_t box[2] = {(cuuint32_t)tc::BK, which == 0 ? 128u : 64u};
        CUresult r = enc(which == 0 ? &op->m128 : &op->m64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.p, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(h, TGP_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows %d cols %d)", (int)r, rows, cols);
    }
    return TGP_OK;
}

// once per process and device, outside any stream capture
static int tc_prepare(tgp_ctx* h) {
    TGP_CUDA(h, cudaFuncSetAttribute(tc::k_tc_gemm_tn<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<64>::kSmem));
    return TGP_OK;
}

// C = alpha X' Y (+ additive term), tile 128 x 64.
static int tc_gemm(tgp_ctx* h, const char* name, const TcOp& X, const TcOp& Y, int K, const tc::Epi& e) {
    constexpr int BN = 64;
    dim3 grid((e.Mx + tc::BM - 1) / tc::BM, (e.N + BN - 1) / BN);
    TGP_K(h, name);
    tc::k_tc_gemm_tn<BN><<<grid, tc::kThreads, tc::Cfg<BN>::kSmem, h->stream>>>(X.m128, Y.m64, K, X.pr.cpad, Y.pr.cpad, e);
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

// Winv = U^-1 for the upper Cholesky factor U (M x M, double, column-major, upper triangle read). One warp per ROW j of
// Winv: z' U = e_j'  <=>  z_i = (delta_ij - sum_{j <= k < i} U[k, i] z_k) / U[i, i], i = j..M-1 (column i of U is contiguous).
__global__ void __launch_bounds__(256) k_tri_inv(const double* __restrict__ U, int M, Pair Winv) {
    extern __shared__ double zbuf[];                 // 8 warps x M
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= M) return;
    double* z = zbuf + (size_t)warp * M;
    for (int i = j; i < M; ++i) {
        const double* col = U + (size_t)M * i;
        double acc = 0.0;
        for (int k = j + lane; k < i; k += 32) acc = fma(col[k], z[k], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) z[i] = ((i == j ? 1.0 : 0.0) - acc) / col[i];
        __syncwarp();
    }
    for (int i = j + lane; i < M; i += 32) {
        const float x = (float)z[i];
        const float hi = tc::tf32_hi(x);
        Winv.hi()[(size_t)j + (size_t)Winv.ld * i] = hi;
        Winv.lo()[(size_t)j + (size_t)Winv.ld * i] = x - hi;
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
    TcOp At, Ht, Pa, Pb, W, Vt, V, B, Winv;
    double *S, *m, *mp, *r, *alpha, *lml;
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
            TGP_K(h, "dense:k_chol_panel");
            k_chol_panel<<<1, 1024, sizeof(double) * nbk * (M - k0), st>>>(w.S, M, k0, w.step, w.err);
            TGP_LAUNCH_CHECK(h);
            const int n = M - k0 - nbk;
            if (n > 0) {
                TGP_K(h, "dense:k_chol_trail");
                k_chol_trail<<<(int)std::min<long long>(((long long)n * n + 255) / 256, 1184), 256, 0, st>>>(w.S, M, k0, nbk);
                TGP_LAUNCH_CHECK(h);
            }
        }
        TGP_K(h, "tc:k_tri_inv");
        k_tri_inv<<<(M + 7) / 8, 256, sizeof(double) * 8 * M, st>>>(w.S, M, w.Winv.pr);
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
    TGP_TRY(dalloc(h, (size_t)M * M, &w.S));
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

    const size_t pan_bytes = sizeof(double) * kCholNB * (size_t)M;
    if (pan_bytes > 200 * 1024) return fail(h, TGP_EUNSUPPORTED, "observation dimension M=%d too large for the Cholesky panel kernel", M);
    if (pan_bytes > 48 * 1024) TGP_CUDA(h, cudaFuncSetAttribute(k_chol_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pan_bytes));
    if (sizeof(double) * 8 * M > 48 * 1024) TGP_CUDA(h, cudaFuncSetAttribute(k_tri_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 8 * M)));
    if (ti && T >= 8 && !h->timing) {
        // One step captured into a CUDA graph and replayed. Every kPoll steps the host looks at the steady-state word;
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
        constexpr int64_t kPoll = 16;
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

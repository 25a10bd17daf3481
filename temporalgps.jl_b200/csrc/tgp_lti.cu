// tgp_lti.cu — device-side construction of the per-step transitions of an LTI SDE on an IRREGULAR time grid.
//
// Replaces broadcast_components((F, q, H), x0, t::AbstractVector, storage) — src/gp/lti_sde.jl:136-147:
//     t' = vcat(first(t) - 1, t);  A[i] = exp(F * dt[i]),  Q[i] = P - A[i] P A[i]',  dt = diff(t')
// so a caller with (F, P, t) never ships 2 D^2 doubles per step through the boundary, and never spends T host matrix
// exponentials: one thread per time step, the exponential by scaling and squaring (|F dt| / 2^s <= 1/4, Taylor order 12, remainder
// <= 0.25^13 / 13! = 2.4e-18 of the unit matrix, then s squarings), rows staged through shared memory so the 2 D^2 doubles per step
// leave as coalesced stores. The arrays it writes are exactly what tgp_lgssm.A / .Q take (column-major per step, stride D * D).
//
// The first transition of a kernel stretched in time (TransformedKernel, lti_sde.jl:361-373) uses the UNSTRETCHED drift with
// dt = 1 (the reference subtracts 1 from the already stretched first input), hence the separate F0.
#include "tgp_ctx.cuh"
#include "tgp_dispatch.h"

namespace tgp {

constexpr int kLtiThreads = 64;
constexpr int kLtiTaylor = 12;

template <int D>
struct LtiParam {
    double F[D * D];    // row-major: F[i * D + j]
    double F0[D * D];
    double P[D * D];    // symmetric
};

template <int D>
__device__ __forceinline__ void lti_mm(const double* __restrict__ X, const double* __restrict__ Y, double* __restrict__ Z) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(X[i * D + k], Y[k * D + j], s);
            Z[i * D + j] = s;
        }
}

// E = exp(F * dt), row-major.
template <int D>
__device__ __forceinline__ void lti_expm(const double* __restrict__ F, double dt, double* __restrict__ E) {
    double X[D * D], W[D * D];
    double nrm = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double c = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) c += fabs(F[i * D + j]);
        nrm = fmax(nrm, c);
    }
    nrm *= fabs(dt);
    int s = 0;
    if (nrm > 0.25 && isfinite(nrm)) {
        s = ilogb(nrm) + 3;            // nrm / 2^s in (1/8, 1/4]
        if (s < 0) s = 0;
        if (s > 60) s = 60;
    }
    const double sc = ldexp(dt, -s);
#pragma unroll
    for (int i = 0; i < D * D; ++i) X[i] = F[i] * sc;
    // Horner: E = I + X (I + X/2 (I + X/3 (... (I + X/K))))
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) E[i * D + j] = X[i * D + j] * (1.0 / kLtiTaylor) + (i == j ? 1.0 : 0.0);
#pragma unroll 1
    for (int k = kLtiTaylor - 1; k >= 1; --k) {
        lti_mm<D>(X, E, W);
        const double r = 1.0 / (double)k;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) E[i * D + j] = W[i * D + j] * r + (i == j ? 1.0 : 0.0);
    }
#pragma unroll 1
    for (int q = 0; q < s; ++q) {
        lti_mm<D>(E, E, W);
#pragma unroll
        for (int i = 0; i < D * D; ++i) E[i] = W[i];
    }
}

template <int D>
__global__ void __launch_bounds__(kLtiThreads) k_lti_components(long long T, const __grid_constant__ LtiParam<D> p,
                                                                 const double* __restrict__ t, double* __restrict__ A_out,
                                                                 double* __restrict__ Q_out) {
    constexpr int DD = D * D;
    constexpr bool kStaged = (size_t)DD * kLtiThreads * sizeof(double) <= 40 * 1024;
    __shared__ double stage[kStaged ? DD * kLtiThreads : 1];
    const long long base = (long long)blockIdx.x * kLtiThreads;
    const long long i = base + threadIdx.x;
    const int nvalid = (int)min((long long)kLtiThreads, T - base);
    double E[DD], Q[DD];
    if (i < T) {
        if (i == 0) lti_expm<D>(p.F0, 1.0, E);
        else        lti_expm<D>(p.F, t[i] - t[i - 1], E);
        double W[DD];
        lti_mm<D>(E, p.P, W);                                   // W = A P
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = r; c < D; ++c) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s = fma(W[r * D + k], E[c * D + k], s);   // (A P A')[r][c]
                const double q = p.P[r * D + c] - s;
                Q[r * D + c] = q;
                Q[c * D + r] = q;
            }
    }
    if constexpr (kStaged) {
        // column-major per step: element (r, c) at c * D + r
        if (i < T) {
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int c = 0; c < D; ++c) stage[threadIdx.x * DD + c * D + r] = E[r * D + c];
        }
        __syncthreads();
        for (int k = threadIdx.x; k < nvalid * DD; k += kLtiThreads) A_out[base * DD + k] = stage[k];
        __syncthreads();
        if (i < T) {
#pragma unroll
            for (int k = 0; k < DD; ++k) stage[threadIdx.x * DD + k] = Q[k];       // symmetric: either layout
        }
        __syncthreads();
        for (int k = threadIdx.x; k < nvalid * DD; k += kLtiThreads) Q_out[base * DD + k] = stage[k];
    } else {
        if (i < T) {
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    A_out[i * DD + c * D + r] = E[r * D + c];
                    Q_out[i * DD + c * D + r] = Q[r * D + c];
                }
        }
    }
}

template <int D>
static int lti_run(tgp_ctx* h, int64_t T, const double* F, const double* F0, const double* P, const double* dt, double* dA, double* dQ) {
    LtiParam<D> p;
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {           // inputs are column-major
            p.F[i * D + j] = F[j * D + i];
            p.F0[i * D + j] = (F0 ? F0 : F)[j * D + i];
            p.P[i * D + j] = i <= j ? P[j * D + i] : P[i * D + j];      // Symmetric(P): upper triangle (lti_sde.jl:138)
        }
    const unsigned blocks = (unsigned)((T + kLtiThreads - 1) / kLtiThreads);
    TGP_K(h, "k_lti_components");
    k_lti_components<D><<<blocks, kLtiThreads, 0, h->stream>>>((long long)T, p, dt, dA, dQ);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

int lti_components(tgp_ctx* h, int D, int64_t T, const double* F, const double* F0, const double* P, const double* t, double* A_out,
                   double* Q_out) {
    const double* dt_ = nullptr;
    TGP_TRY(stage_in(h, t, (size_t)T, &dt_));
    double *dA = nullptr, *dQ = nullptr;
    int64_t s1;
    TGP_TRY(stage_out(h, A_out, (size_t)D * D, (int64_t)D * D, T, &dA, &s1));
    TGP_TRY(stage_out(h, Q_out, (size_t)D * D, (int64_t)D * D, T, &dQ, &s1));
    int rc;
    switch (D) {
#define TGP_LTI_CASE(Dv) case Dv: rc = lti_run<Dv>(h, T, F, F0, P, dt_, dA, dQ); break;
        TGP_FOR_EACH_D(TGP_LTI_CASE)
#undef TGP_LTI_CASE
        default: return fail(h, TGP_EUNSUPPORTED, "latent dimension D=%d has no kernel instantiation in this build", D);
    }
    TGP_TRY(rc);
    TGP_TRY(flush_outputs(h));
    if (!is_device_ptr(A_out) || !is_device_ptr(Q_out)) TGP_CUDA(h, cudaStreamSynchronize(h->stream));   // host copies are complete on return
    return TGP_OK;
}

}  // namespace tgp

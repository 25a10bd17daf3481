// kernel_probe.cu — runs the general-path kernel sequence for T = 1 on a fixed small model and dumps the intermediate
// buffers; built twice (rolled / unrolled) to find where the two modes diverge. Diagnostic only.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../temporalgps.jl_b200/csrc/tgp_scan_small.cuh"
using namespace tgp;
constexpr int D = TGP_D;
static void dump(const char* name, const double* d, int n) {
    std::vector<double> h(n); cudaMemcpy(h.data(), d, n * 8, cudaMemcpyDeviceToHost);
    printf("%-8s", name); for (int i = 0; i < n; ++i) printf(" %.6g", h[i]); printf("\n");
}
int main() {
    constexpr int SN = D + Sym<D>::N, EN = Elem<D>::N;
    std::vector<double> A(D * D), a(D), Q(D * D), H(D), m0(D), P0(D * D);
    for (int j = 0; j < D; ++j) for (int i = 0; i < D; ++i) { A[i + D * j] = (i == j ? 0.9 : 0.0) + 0.05 * ((i * 7 + j * 3) % 5 - 2); Q[i + D * j] = (i == j ? 0.5 : 0.05); P0[i + D * j] = (i == j ? 1.0 + 0.1 * i : 0.1); }
    for (int i = 0; i < D; ++i) { a[i] = 0.1 * i; H[i] = 1.0 / (1 + i); m0[i] = 0.3 - 0.1 * i; }
    double h = 0.2, R = 0.3, y = 1.1;
    double *dA, *da, *dQ, *dH, *dh, *dR, *dy, *dm0, *dP0;
    auto up = [](const void* p, size_t n) { double* d; cudaMalloc(&d, n); cudaMemcpy(d, p, n, cudaMemcpyHostToDevice); return d; };
    dA = up(A.data(), D * D * 8); da = up(a.data(), D * 8); dQ = up(Q.data(), D * D * 8); dH = up(H.data(), D * 8); dh = up(&h, 8); dR = up(&R, 8);
    dy = up(&y, 8); dm0 = up(m0.data(), D * 8); dP0 = up(P0.data(), D * D * 8);
    const long long nthreads = kBlock, nwarps = kBlock / 32;
    double *x0buf, *xT, *excl, *wagg, *wstate, *partials, *lml, *steps, *extra; unsigned long long* err;
    cudaMalloc(&x0buf, SN * 8); cudaMalloc(&xT, SN * 8); cudaMalloc(&excl, EN * nthreads * 8); cudaMalloc(&wagg, EN * nwarps * 8);
    cudaMalloc(&wstate, SN * nwarps * 8); cudaMalloc(&partials, 8); cudaMalloc(&lml, 8); cudaMalloc(&steps, 8); cudaMalloc(&extra, 8); cudaMalloc(&err, 8);
    cudaMemset(err, 0xFF, 8);
    k_init_state<D><<<1, 32>>>(dm0, dP0, x0buf, 0, dH, dh, dR, dy, extra, nullptr, nullptr, nullptr, nullptr, 1, 0, err);
    dump("x0buf", x0buf, SN);
    DevModel dm{dA, da, dQ, dH, dh, dR, 1, 1, 1, 1, 1, 1, dy, 1, 1};
    k_filter_reduce<D, true><<<1, kBlock>>>(dm, nullptr, 16, nthreads, excl, wagg, nwarps);
    dump("wagg0", wagg, 4); 
    { std::vector<double> hbuf(EN * nwarps); cudaMemcpy(hbuf.data(), wagg, EN * nwarps * 8, cudaMemcpyDeviceToHost); printf("wagg w0:"); for (int k = 0; k < EN; ++k) printf(" %.6g", hbuf[k * nwarps]); printf("\n"); }
    { std::vector<double> hbuf(EN * nthreads); cudaMemcpy(hbuf.data(), excl, EN * nthreads * 8, cudaMemcpyDeviceToHost); printf("excl t1:"); for (int k = 0; k < EN; ++k) printf(" %.6g", hbuf[k * nthreads + 1]); printf("\nexcl t0:"); for (int k = 0; k < EN; ++k) printf(" %.6g", hbuf[k * nthreads]); printf("\n"); }
    k_filter_mid<D><<<1, kMidThreads>>>(wagg, nwarps, x0buf, wstate, xT);
    { std::vector<double> hbuf(SN * nwarps); cudaMemcpy(hbuf.data(), wstate, SN * nwarps * 8, cudaMemcpyDeviceToHost); printf("wstate w0:"); for (int k = 0; k < SN; ++k) printf(" %.6g", hbuf[k * nwarps]); printf("\n"); }
    dump("xT", xT, SN);
    FilterOut fo{}; fo.lml_steps = steps; fo.s_l = 1; fo.partials = partials; fo.err_step = err;
    k_filter_apply<D, true><<<1, kBlock>>>(dm, nullptr, 16, nthreads, excl, wstate, nwarps, fo);
    dump("lml_step", steps, 1);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

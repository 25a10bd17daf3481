// rolled_probe.cu — compares the device results of the tgp_math.cuh primitives against the host results of the same
// header, in the rolled-loop mode (-DTGP_D=3 -DTGP_REG_D=2). Diagnostic only.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../temporalgps.jl_b200/csrc/tgp_math.cuh"
using namespace tgp;
constexpr int D = TGP_D;
struct Out { double v[8][64]; };

TGP_HD void run(Out& o) {
    Mat<D> A; Vec<D> a, H, m; Sym<D> Q, P;
    for (int j = 0; j < D; ++j) for (int i = 0; i < D; ++i) A(i, j) = (i == j ? 0.9 : 0.0) + 0.05 * ((i * 7 + j * 3) % 5 - 2);
    for (int i = 0; i < D; ++i) { a[i] = 0.1 * i; H[i] = 1.0 / (1 + i); m[i] = 0.3 - 0.1 * i; }
    for (int j = 0; j < D; ++j) for (int i = 0; i <= j; ++i) { Q(i, j) = (i == j ? 0.5 : 0.05); P(i, j) = (i == j ? 1.0 + 0.1 * i : 0.1); }
    // 0: predict
    Vec<D> m1 = m; Sym<D> P1 = P; predict(m1, P1, A, a, Q);
    for (int i = 0; i < D; ++i) o.v[0][i] = m1[i];
    for (int i = 0; i < Sym<D>::N; ++i) o.v[0][D + i] = P1.v[i];
    // 1: update
    double quad; double S = update_scalar(m1, P1, H, 0.2, 0.3, 1.1, &quad);
    o.v[1][0] = S; o.v[1][1] = quad;
    for (int i = 0; i < D; ++i) o.v[1][2 + i] = m1[i];
    for (int i = 0; i < Sym<D>::N; ++i) o.v[1][2 + D + i] = P1.v[i];
    // 2: step const + fold
    StepConst<D> sc = make_step_const<D>(A, a, Q, H, 0.2, 0.3);
    Elem<D> E = elem_identity<D>();
    fold_step(E, sc, 1.1); fold_step(E, sc, -0.4);
    for (int i = 0; i < D * D; ++i) o.v[2][i] = E.A.v[i];
    for (int i = 0; i < D; ++i) { o.v[2][D * D + i] = E.b[i]; o.v[2][D * D + D + i] = E.eta[i]; }
    for (int i = 0; i < Sym<D>::N; ++i) { o.v[3][i] = E.C.v[i]; o.v[3][Sym<D>::N + i] = E.J.v[i]; }
    // 4: combine
    Elem<D> F = elem_identity<D>(); fold_step(F, sc, 0.7);
    Elem<D> G = combine(E, F);
    for (int i = 0; i < D * D; ++i) o.v[4][i] = G.A.v[i];
    for (int i = 0; i < D; ++i) { o.v[4][D * D + i] = G.b[i]; o.v[4][D * D + D + i] = G.eta[i]; }
    for (int i = 0; i < Sym<D>::N; ++i) { o.v[5][i] = G.C.v[i]; o.v[5][Sym<D>::N + i] = G.J.v[i]; }
    // 6: apply
    Vec<D> m2 = m; Sym<D> P2 = P; apply_elem(G, m2, P2);
    for (int i = 0; i < D; ++i) o.v[6][i] = m2[i];
    for (int i = 0; i < Sym<D>::N; ++i) o.v[6][D + i] = P2.v[i];
    // 7: apply identity
    Vec<D> m3 = m; Sym<D> P3 = P; Elem<D> I = elem_identity<D>(); apply_elem(I, m3, P3);
    for (int i = 0; i < D; ++i) o.v[7][i] = m3[i] - m[i];
    for (int i = 0; i < Sym<D>::N; ++i) o.v[7][D + i] = P3.v[i] - P.v[i];
}
__global__ void k(Out* o) { run(*o); }
int main() {
    Out hst = {}, dev = {}; run(hst);
    Out* d; cudaMalloc(&d, sizeof(Out)); cudaMemset(d, 0, sizeof(Out)); k<<<1, 1>>>(d); cudaMemcpy(&dev, d, sizeof(Out), cudaMemcpyDeviceToHost);
    const char* names[] = {"predict", "update", "fold(A,b,eta)", "fold(C,J)", "combine(A,b,eta)", "combine(C,J)", "apply", "apply(identity) - x"};
    for (int r = 0; r < 8; ++r) { double e = 0; for (int i = 0; i < 64; ++i) e = fmax(e, fabs(hst.v[r][i] - dev.v[r][i])); printf("%-18s max|host-device| = %.3e\n", names[r], e); }
    printf("apply(identity)-x on device:"); for (int i = 0; i < D + Sym<D>::N; ++i) printf(" %.3g", dev.v[7][i]); printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
}

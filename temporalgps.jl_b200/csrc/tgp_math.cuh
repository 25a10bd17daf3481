// tgp_math.cuh — register-resident small-matrix algebra for the LGSSM scan kernels.
//
// Everything here is `__host__ __device__` so the same code is exercised by the CUDA kernels and
// by the g++-compiled algebra tests (tests/emul/). Matrices are column-major (Julia layout);
// covariances are kept as packed upper triangles (the reference reads P through
// `Symmetric(P)`, i.e. its upper triangle: linear_gaussian_conditionals.jl:50-51).
//
// The filtering recursion of the reference (scan.jl:22-25 driving lgssm.jl:155-159) is
// re-expressed as an associative scan over 5-tuples (A, b, C, eta, J) (Särkkä &
// García-Fernández 2021):
//   p(x_k | y_k, x_{k-1}) = N(A x_{k-1} + b, C),   p(y_k | x_{k-1}) ∝ N_info(x_{k-1}; eta, J).
// combine(i, j), i earlier:  M = (I + C_i J_j)^-1
//   A = A_j M A_i          b = A_j M (b_i + C_i eta_j) + b_j     C = A_j M C_i A_j' + C_j
//   eta = A_i' M' (eta_j - J_j b_i) + eta_i                      J = A_i' M' J_j A_i + J_i
// For scalar observations J_j = u u' is rank one and M is a Sherman–Morrison update (fold_step).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TGP_HD __host__ __device__ __forceinline__
#else
#define TGP_HD inline __attribute__((always_inline))
#endif

// Latent dimensions up to TGP_REG_D keep every small matrix in registers (loops fully unrolled); larger D fall back
// to rolled loops over thread-local arrays (correct, slower: the warp-cooperative layout for large D is future work).
#ifndef TGP_REG_D
#define TGP_REG_D 6
#endif
// Each latent dimension is its own translation unit (tgp_inst.cu, -DTGP_D=<D>), so the choice is a plain macro.
#if defined(TGP_D) && TGP_D > TGP_REG_D
#if defined(TGP_ROLLED_PRAGMA1)
#define TGP_UNROLL _Pragma("unroll 1")
#else
#define TGP_UNROLL   /* no pragma: the compiler's own heuristics (forcing "unroll 1" miscompiles apply_elem) */
#endif
#else
#define TGP_UNROLL _Pragma("unroll")
#endif

namespace tgp {

constexpr int kRegD = TGP_REG_D;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kInvertJitter = 1e-10;  // lgssm.jl:235

template <int D>
struct Vec {
    double v[D];
    TGP_HD double& operator[](int i) { return v[i]; }
    TGP_HD const double& operator[](int i) const { return v[i]; }
};

template <int D>
struct Mat {  // column-major
    double v[D * D];
    TGP_HD double& operator()(int i, int j) { return v[i + D * j]; }
    TGP_HD const double& operator()(int i, int j) const { return v[i + D * j]; }
};

template <int D>
struct Sym {  // packed upper triangle, (i<=j) at j(j+1)/2 + i
    static constexpr int N = D * (D + 1) / 2;
    double v[N];
    TGP_HD static constexpr int idx(int i, int j) { return i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j; }
    TGP_HD double& operator()(int i, int j) { return v[idx(i, j)]; }
    TGP_HD const double& operator()(int i, int j) const { return v[idx(i, j)]; }
};

template <int D> TGP_HD Vec<D> vzero() { Vec<D> r;
TGP_UNROLL
    for (int i = 0; i < D; ++i) r.v[i] = 0.0; return r; }
template <int D> TGP_HD Sym<D> szero() { Sym<D> r;
TGP_UNROLL
    for (int i = 0; i < Sym<D>::N; ++i) r.v[i] = 0.0; return r; }
template <int D> TGP_HD Mat<D> meye() { Mat<D> r;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) r(i, j) = (i == j) ? 1.0 : 0.0;
    return r; }

template <int D> TGP_HD double dot(const Vec<D>& a, const Vec<D>& b) {
    double s = 0.0;
TGP_UNROLL
    for (int i = 0; i < D; ++i) s = fma(a[i], b[i], s);
    return s;
}
// A x
template <int D> TGP_HD Vec<D> matvec(const Mat<D>& A, const Vec<D>& x) {
    Vec<D> r;
TGP_UNROLL
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
TGP_UNROLL
        for (int j = 0; j < D; ++j) s = fma(A(i, j), x[j], s);
        r[i] = s;
    }
    return r;
}
// A' x
template <int D> TGP_HD Vec<D> matTvec(const Mat<D>& A, const Vec<D>& x) {
    Vec<D> r;
TGP_UNROLL
    for (int j = 0; j < D; ++j) {
        double s = 0.0;
TGP_UNROLL
        for (int i = 0; i < D; ++i) s = fma(A(i, j), x[i], s);
        r[j] = s;
    }
    return r;
}
template <int D> TGP_HD Vec<D> symvec(const Sym<D>& S, const Vec<D>& x) {
    Vec<D> r;
TGP_UNROLL
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
TGP_UNROLL
        for (int j = 0; j < D; ++j) s = fma(S(i, j), x[j], s);
        r[i] = s;
    }
    return r;
}
template <int D> TGP_HD Mat<D> matmul(const Mat<D>& A, const Mat<D>& B) {
    Mat<D> C;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
TGP_UNROLL
            for (int k = 0; k < D; ++k) s = fma(A(i, k), B(k, j), s);
            C(i, j) = s;
        }
    return C;
}
// A * S (S symmetric packed) -> full
template <int D> TGP_HD Mat<D> mat_sym(const Mat<D>& A, const Sym<D>& S) {
    Mat<D> C;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
TGP_UNROLL
            for (int k = 0; k < D; ++k) s = fma(A(i, k), S(k, j), s);
            C(i, j) = s;
        }
    return C;
}
// S * A (S symmetric packed) -> full
template <int D> TGP_HD Mat<D> sym_mat(const Sym<D>& S, const Mat<D>& A) {
    Mat<D> C;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
TGP_UNROLL
            for (int k = 0; k < D; ++k) s = fma(S(i, k), A(k, j), s);
            C(i, j) = s;
        }
    return C;
}
// upper triangle of X * A' + C0 (result assumed symmetric)
template <int D> TGP_HD Sym<D> mat_matT_sym(const Mat<D>& X, const Mat<D>& A, const Sym<D>& C0) {
    Sym<D> R;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i <= j; ++i) {
            double s = C0(i, j);
TGP_UNROLL
            for (int k = 0; k < D; ++k) s = fma(X(i, k), A(j, k), s);
            R(i, j) = s;
        }
    return R;
}
// upper triangle of X' * Y + C0 (result assumed symmetric)
template <int D> TGP_HD Sym<D> matT_mat_sym(const Mat<D>& X, const Mat<D>& Y, const Sym<D>& C0) {
    Sym<D> R;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i <= j; ++i) {
            double s = C0(i, j);
TGP_UNROLL
            for (int k = 0; k < D; ++k) s = fma(X(k, i), Y(k, j), s);
            R(i, j) = s;
        }
    return R;
}
// A S A' + Q, symmetric result
template <int D> TGP_HD Sym<D> congruence(const Mat<D>& A, const Sym<D>& S, const Sym<D>& Q) {
    return mat_matT_sym(mat_sym(A, S), A, Q);
}

// ---------------------------------------------------------------------------------------------
// The ordinary Kalman step (what the reference executes sequentially).
// ---------------------------------------------------------------------------------------------
// predict: linear_gaussian_conditionals.jl:46-52
template <int D>
TGP_HD void predict(Vec<D>& m, Sym<D>& P, const Mat<D>& A, const Vec<D>& a, const Sym<D>& Q) {
    Vec<D> mn = matvec(A, m);
TGP_UNROLL
    for (int i = 0; i < D; ++i) mn[i] += a[i];
    P = congruence(A, P, Q);
    m = mn;
}

// posterior_and_lml for ScalarOutputLGC: linear_gaussian_conditionals.jl:247-257.
// Returns S = H P H' + R (<= 0 or NaN signals "not positive definite"); *quad = alpha^2.
template <int D>
TGP_HD double update_scalar(Vec<D>& m, Sym<D>& P, const Vec<D>& H, double h, double R, double y,
                            double* quad) {
    Vec<D> V = symvec(P, H);
    const double S = dot(V, H) + R;
    const double is = 1.0 / sqrt(S);
    const double alpha = (y - (dot(H, m) + h)) * is;
    Vec<D> B;
TGP_UNROLL
    for (int i = 0; i < D; ++i) B[i] = V[i] * is;
TGP_UNROLL
    for (int i = 0; i < D; ++i) m[i] = fma(B[i], alpha, m[i]);
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i <= j; ++i) P(i, j) = fma(-B[i], B[j], P(i, j));
    *quad = alpha * alpha;
    return S;
}
TGP_HD double lml_from(double S, double quad) { return -0.5 * (kLog2Pi + log(S) + quad); }

// predict(x, emission) in emission space (lgssm.jl:103-115 with LGC:46-52), M = 1.
template <int D>
TGP_HD void emit_scalar(const Vec<D>& m, const Sym<D>& P, const Vec<D>& H, double h, double R,
                        double* mean, double* var) {
    *mean = dot(H, m) + h;
    *var = dot(symvec(P, H), H) + R;
}

// ---------------------------------------------------------------------------------------------
// Scan element and its algebra.
// ---------------------------------------------------------------------------------------------
template <int D>
struct Elem {
    Mat<D> A;
    Vec<D> b;
    Sym<D> C;
    Vec<D> eta;
    Sym<D> J;
    static constexpr int N = D * D + 2 * D + 2 * Sym<D>::N;  // doubles when packed
};

template <int D> TGP_HD Elem<D> elem_identity() {
    Elem<D> e;
    e.A = meye<D>();
    e.b = vzero<D>();
    e.C = szero<D>();
    e.eta = vzero<D>();
    e.J = szero<D>();
    return e;
}

// Per-step quantities of a Forward step (predict with (A,a,Q), then scalar update with (H,h,R)),
// independent of y:  S = H Q H' + R;  kq = Q H'/sqrt(S);  u = A' H'/sqrt(S);
//   A_k = A - kq u',  C_k = Q - kq kq',  b_k = a + kq r,  eta_k = u r,  J_k = u u',
//   r = (y - H a - h)/sqrt(S).
template <int D>
struct StepConst {
    Mat<D> Ak;
    Sym<D> Ck;
    Vec<D> a, kq, u;
    double c0;   // (H a + h)
    double is;   // 1/sqrt(S)
    double S;
};

template <int D>
TGP_HD StepConst<D> make_step_const(const Mat<D>& A, const Vec<D>& a, const Sym<D>& Q, const Vec<D>& H,
                                    double h, double R) {
    StepConst<D> sc;
    Vec<D> qh = symvec(Q, H);
    sc.S = dot(qh, H) + R;
    sc.is = 1.0 / sqrt(sc.S);
    Vec<D> ath = matTvec(A, H);
TGP_UNROLL
    for (int i = 0; i < D; ++i) { sc.kq[i] = qh[i] * sc.is; sc.u[i] = ath[i] * sc.is; sc.a[i] = a[i]; }
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) sc.Ak(i, j) = fma(-sc.kq[i], sc.u[j], A(i, j));
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i <= j; ++i) sc.Ck(i, j) = fma(-sc.kq[i], sc.kq[j], Q(i, j));
    sc.c0 = dot(H, a) + h;
    return sc;
}

// E <- step_k ∘ E  (E earlier). Sherman–Morrison form of combine() for a rank-one J_k.
template <int D>
TGP_HD void fold_step(Elem<D>& E, const StepConst<D>& sc, double y) {
    const double r = (y - sc.c0) * sc.is;
    const Vec<D> w = symvec(E.C, sc.u);
    const double d = 1.0 + dot(sc.u, w);
    const double id = 1.0 / d;
    const Vec<D> v = matTvec(E.A, sc.u);  // A_i' u
    // M A_i = A_i - w v'/d
    Mat<D> MA;
TGP_UNROLL
    for (int j = 0; j < D; ++j) {
        const double vj = v[j] * id;
TGP_UNROLL
        for (int i = 0; i < D; ++i) MA(i, j) = fma(-w[i], vj, E.A(i, j));
    }
    // t = b_i + C_i eta_k = b_i + w r;  M t = t - w (u.t)/d
    Vec<D> t;
TGP_UNROLL
    for (int i = 0; i < D; ++i) t[i] = fma(w[i], r, E.b[i]);
    const double ut = dot(sc.u, t) * id;
TGP_UNROLL
    for (int i = 0; i < D; ++i) t[i] = fma(-w[i], ut, t[i]);
    // eta, J use the *old* b_i, A_i (through v)
    const double s = (r - dot(sc.u, E.b)) * id;
TGP_UNROLL
    for (int i = 0; i < D; ++i) E.eta[i] = fma(v[i], s, E.eta[i]);
TGP_UNROLL
    for (int j = 0; j < D; ++j) {
        const double vj = v[j] * id;
TGP_UNROLL
        for (int i = 0; i <= j; ++i) E.J(i, j) = fma(v[i], vj, E.J(i, j));
    }
    // X = M C_i = C_i - w w'/d
    Sym<D> X;
TGP_UNROLL
    for (int j = 0; j < D; ++j) {
        const double wj = w[j] * id;
TGP_UNROLL
        for (int i = 0; i <= j; ++i) X(i, j) = fma(-w[i], wj, E.C(i, j));
    }
    E.A = matmul(sc.Ak, MA);
    Vec<D> bn = matvec(sc.Ak, t);
TGP_UNROLL
    for (int i = 0; i < D; ++i) E.b[i] = bn[i] + fma(sc.kq[i], r, sc.a[i]);
    E.C = congruence(sc.Ak, X, sc.Ck);
}

// Solve (I + C J) X = RHS for NR right-hand sides by Gauss–Jordan elimination with partial
// pivoting, fully unrolled on registers. I + C J has eigenvalues >= 1 for PSD C, J, so it is never
// singular; pivoting keeps the elimination stable when C J is large (small observation noise).
template <int D, int NR>
TGP_HD void solve_I_plus_CJ(const Sym<D>& C, const Sym<D>& J, double (&rhs)[D][NR]) {
    double a[D][D];
TGP_UNROLL
    for (int i = 0; i < D; ++i)
TGP_UNROLL
        for (int j = 0; j < D; ++j) {
            double s = (i == j) ? 1.0 : 0.0;
TGP_UNROLL
            for (int k = 0; k < D; ++k) s = fma(C(i, k), J(k, j), s);
            a[i][j] = s;
        }
TGP_UNROLL
    for (int k = 0; k < D; ++k) {
        // bring the largest |a[r][k]|, r >= k, to row k by a bubble pass of predicated swaps
TGP_UNROLL
        for (int r = k + 1; r < D; ++r) {
            const bool sw = fabs(a[r][k]) > fabs(a[k][k]);
TGP_UNROLL
            for (int j = k; j < D; ++j) {
                const double x = a[k][j], z = a[r][j];
                a[k][j] = sw ? z : x;
                a[r][j] = sw ? x : z;
            }
TGP_UNROLL
            for (int j = 0; j < NR; ++j) {
                const double x = rhs[k][j], z = rhs[r][j];
                rhs[k][j] = sw ? z : x;
                rhs[r][j] = sw ? x : z;
            }
        }
        const double ip = 1.0 / a[k][k];
TGP_UNROLL
        for (int j = k + 1; j < D; ++j) a[k][j] *= ip;
TGP_UNROLL
        for (int j = 0; j < NR; ++j) rhs[k][j] *= ip;
TGP_UNROLL
        for (int r = 0; r < D; ++r) {
            if (r == k) continue;
            const double f = a[r][k];
TGP_UNROLL
            for (int j = k + 1; j < D; ++j) a[r][j] = fma(-f, a[k][j], a[r][j]);
TGP_UNROLL
            for (int j = 0; j < NR; ++j) rhs[r][j] = fma(-f, rhs[k][j], rhs[r][j]);
        }
    }
}

// combine(Ei, Ej): Ei covers earlier steps.
template <int D>
TGP_HD Elem<D> combine(const Elem<D>& Ei, const Elem<D>& Ej) {
    // RHS columns: [A_i (D) | b_i + C_i eta_j (1) | C_i (D)]
    double rhs[D][2 * D + 1];
    const Vec<D> ce = symvec(Ei.C, Ej.eta);
TGP_UNROLL
    for (int i = 0; i < D; ++i) {
TGP_UNROLL
        for (int j = 0; j < D; ++j) { rhs[i][j] = Ei.A(i, j); rhs[i][D + 1 + j] = Ei.C(i, j); }
        rhs[i][D] = Ei.b[i] + ce[i];
    }
    solve_I_plus_CJ<D, 2 * D + 1>(Ei.C, Ej.J, rhs);
    Mat<D> XA, XC;
    Vec<D> xt;
TGP_UNROLL
    for (int i = 0; i < D; ++i) {
TGP_UNROLL
        for (int j = 0; j < D; ++j) { XA(i, j) = rhs[i][j]; XC(i, j) = rhs[i][D + 1 + j]; }
        xt[i] = rhs[i][D];
    }
    Elem<D> E;
    E.A = matmul(Ej.A, XA);
    Vec<D> bn = matvec(Ej.A, xt);
TGP_UNROLL
    for (int i = 0; i < D; ++i) E.b[i] = bn[i] + Ej.b[i];
    // C = A_j (M C_i) A_j' + C_j ; M C_i is symmetric in exact arithmetic -> symmetrise
    Sym<D> XCs;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i <= j; ++i) XCs(i, j) = 0.5 * (XC(i, j) + XC(j, i));
    E.C = congruence(Ej.A, XCs, Ej.C);
    // eta = (M A_i)' (eta_j - J_j b_i) + eta_i ;  J = (M A_i)' (J_j A_i) + J_i
    Vec<D> z = symvec(Ej.J, Ei.b);
TGP_UNROLL
    for (int i = 0; i < D; ++i) z[i] = Ej.eta[i] - z[i];
    Vec<D> en = matTvec(XA, z);
TGP_UNROLL
    for (int i = 0; i < D; ++i) E.eta[i] = en[i] + Ei.eta[i];
    E.J = matT_mat_sym(XA, sym_mat(Ej.J, Ei.A), Ei.J);
    return E;
}

// Apply an element to the filtering distribution (m, P) that precedes it:
//   M = (I + P J)^-1;  m' = A M (m + P eta) + b;  P' = A M P A' + C.
template <int D>
TGP_HD void apply_elem(const Elem<D>& E, Vec<D>& m, Sym<D>& P) {
    double rhs[D][D + 1];
    const Vec<D> pe = symvec(P, E.eta);
TGP_UNROLL
    for (int i = 0; i < D; ++i) {
TGP_UNROLL
        for (int j = 0; j < D; ++j) rhs[i][1 + j] = P(i, j);
        rhs[i][0] = m[i] + pe[i];
    }
    solve_I_plus_CJ<D, D + 1>(P, E.J, rhs);
    Vec<D> xt;
    Sym<D> XP;
TGP_UNROLL
    for (int i = 0; i < D; ++i) xt[i] = rhs[i][0];
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i <= j; ++i) XP(i, j) = 0.5 * (rhs[i][1 + j] + rhs[j][1 + i]);
    Vec<D> mn = matvec(E.A, xt);
TGP_UNROLL
    for (int i = 0; i < D; ++i) m[i] = mn[i] + E.b[i];
    P = congruence(E.A, XP, E.C);
}

// ---------------------------------------------------------------------------------------------
// Plain affine-Gaussian semigroup (A, b, C): data-free marginals and the backward (RTS) pass.
//   (A2,b2,C2) ∘ (A1,b1,C1) = (A2 A1, A2 b1 + b2, A2 C1 A2' + C2)          
// ---------------------------------------------------------------------------------------------
template <int D>
struct Aff {
    Mat<D> A;
    Vec<D> b;
    Sym<D> C;
    static constexpr int N = D * D + D + Sym<D>::N;
};
template <int D> TGP_HD Aff<D> aff_identity() {
    Aff<D> e; e.A = meye<D>(); e.b = vzero<D>(); e.C = szero<D>(); return e;
}
// later ∘ earlier
template <int D> TGP_HD Aff<D> aff_combine(const Aff<D>& earlier, const Aff<D>& later) {
    Aff<D> e;
    e.A = matmul(later.A, earlier.A);
    Vec<D> t = matvec(later.A, earlier.b);
TGP_UNROLL
    for (int i = 0; i < D; ++i) e.b[i] = t[i] + later.b[i];
    e.C = congruence(later.A, earlier.C, later.C);
    return e;
}
template <int D> TGP_HD void aff_apply(const Aff<D>& e, Vec<D>& m, Sym<D>& P) {
    Vec<D> t = matvec(e.A, m);
TGP_UNROLL
    for (int i = 0; i < D; ++i) m[i] = t[i] + e.b[i];
    P = congruence(e.A, P, e.C);
}

// Upper Cholesky factor of a packed symmetric matrix: S = U'U, U returned packed (upper).
// Returns false if a pivot is not positive.
template <int D> TGP_HD bool chol_upper(const Sym<D>& S, Sym<D>& U) {
    bool ok = true;
TGP_UNROLL
    for (int j = 0; j < D; ++j) {
        double d = S(j, j);
TGP_UNROLL
        for (int k = 0; k < j; ++k) d = fma(-U(k, j), U(k, j), d);
        ok = ok && (d > 0.0);
        const double sd = sqrt(d);
        const double isd = 1.0 / sd;
        U(j, j) = sd;
TGP_UNROLL
        for (int i = j + 1; i < D; ++i) {
            double s = S(j, i);
TGP_UNROLL
            for (int k = 0; k < j; ++k) s = fma(-U(k, j), U(k, i), s);
            U(j, i) = s * isd;
        }
    }
    return ok;
}

// invert_dynamics (lgssm.jl:231-240) as an affine element mapping x_t -> x_{t-1}:
//   U = chol(Pp + 1e-10 I);  G = Pf A' (Pp + eps I)^-1;  g = mf - G mp;  Sigma = Pf - (U G')'(U G').
template <int D>
TGP_HD bool invert_dynamics(const Vec<D>& mf, const Sym<D>& Pf, const Vec<D>& mp, const Sym<D>& Pp,
                            const Mat<D>& A, Aff<D>& out) {
    Sym<D> Pj = Pp, U;
TGP_UNROLL
    for (int i = 0; i < D; ++i) Pj(i, i) += kInvertJitter;
    const bool ok = chol_upper(Pj, U);
    // X = A Pf (D x D); B = U' \ X (forward substitution per column); Gt = U \ B
    Mat<D> X = mat_sym(A, Pf), B, Gt;
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) {
            double s = X(i, j);
TGP_UNROLL
            for (int k = 0; k < i; ++k) s = fma(-U(k, i), B(k, j), s);
            B(i, j) = s / U(i, i);
        }
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = D - 1; i >= 0; --i) {
            double s = B(i, j);
TGP_UNROLL
            for (int k = i + 1; k < D; ++k) s = fma(-U(i, k), Gt(k, j), s);
            Gt(i, j) = s / U(i, i);
        }
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) out.A(i, j) = Gt(j, i);
    Vec<D> gm = matvec(out.A, mp);
TGP_UNROLL
    for (int i = 0; i < D; ++i) out.b[i] = mf[i] - gm[i];
    // Sigma = Pf - B'B  (B == U Gt up to rounding; the reference recomputes U*Gt, lgssm.jl:237)
TGP_UNROLL
    for (int j = 0; j < D; ++j)
TGP_UNROLL
        for (int i = 0; i <= j; ++i) {
            double s = Pf(i, j);
TGP_UNROLL
            for (int k = 0; k < D; ++k) s = fma(-B(k, i), B(k, j), s);
            out.C(i, j) = s;
        }
    return ok;
}

}  // namespace tgp

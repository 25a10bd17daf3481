// tgp_fir_plan.h — plan of the one-launch steady-state logpdf (tgp_fir.cuh): constants, per-lane arithmetic and the host-side
// builder. No CUDA dependency beyond tgp_math.cuh's host/device macros, so tests/emul/ compiles it with g++.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "tgp_math.cuh"

namespace tgp {

constexpr int kFirB = 8;          // block: steps advanced per loop-carried state update
constexpr int kFirNBlk = 4;       // blocks per lane run
constexpr int kFirL = kFirB * kFirNBlk;   // 32 steps per lane
constexpr int kFirTile = 32 * kFirL;      // 1024 steps per warp tile
constexpr int kFirNbMax = 3;      // look-back depth limit (tiles)
constexpr int kFirMaxD = 4;        // instantiated for D <= 4 (larger states: the two-phase kernel / the vector scans)

// Hot-loop constants: passed BY VALUE (kernel parameter space = constant bank, so every coefficient is a DFMA operand).
template <int D>
struct FirPlan {
    double gK[kFirB][D];          // Abar^(7-j) K
    double zc[D];                 // sum_j Abar^(7-j) c
    double A8[D][D];              // Abar^8, [row][col]
    double wA[kFirB][D];          // w' Abar^j
    double g[kFirB];              // g[k] = w' Abar^k K (k < 7 used)
    double kap[kFirB];            // hh + w' sum_{i<j} Abar^(j-1-i) c
    double P2[5][D][D];           // Abar^(32 2^k)
    double PT[kFirNbMax][D][D];   // Abar^(1024 k), PT[0] = I
    double A[D][D], a[D], w[D], m0[D];   // the transient's recursion (x' = A x + a + K_t (y - hh - w'x))
    double hh;                    // H a + h
    double invS;                  // 1 / S_inf
    double c0;                    // data-independent part of the log-likelihood of this shard
    long long N0;                 // steps handled by the transient (0 on a shard with rank > 0)
    long long T;                  // steps of this shard
    long long ntiles;             // ceil((T - N0) / 1024)
    int nb;                       // look-back depth in tiles
    int aligned;                  // y + N0 is 32-byte aligned (16-byte cp.async)
    int zero_mean;                // a = 0 and h = 0: kap, zc vanish (the kernel has a variant without those terms)
};


// ---- per-lane arithmetic of one tile (shared by the kernel and the g++-compiled emulation in tests/emul/) ---------------------
// pass A, one block: u_b = zc + sum_j gK[j] y_j; z <- Abar^8 z + u_b (z = zero-state response of the lane's run so far).
template <int D, bool ZM = false>
TGP_HD void fir_pass_a_block(const FirPlan<D>& pl, int b, const double (&yv)[kFirL], Vec<D>& ub, Vec<D>& z) {
TGP_UNROLL
    for (int i = 0; i < D; ++i) ub[i] = ZM ? pl.gK[0][i] * yv[b * kFirB] : fma(pl.gK[0][i], yv[b * kFirB], pl.zc[i]);
TGP_UNROLL
    for (int j = 1; j < kFirB; ++j)
TGP_UNROLL
        for (int i = 0; i < D; ++i) ub[i] = fma(pl.gK[j][i], yv[b * kFirB + j], ub[i]);
    Vec<D> zn = ub;
    if (b > 0) {
TGP_UNROLL
        for (int i = 0; i < D; ++i)
TGP_UNROLL
            for (int j = 0; j < D; ++j) zn[i] = fma(pl.A8[i][j], z[j], zn[i]);
    }
    z = zn;
}
template <int D>
TGP_HD void fir_pass_a(const FirPlan<D>& pl, const double (&yv)[kFirL], Vec<D> (&u)[kFirNBlk], Vec<D>& z) {
TGP_UNROLL
    for (int b = 0; b < kFirNBlk; ++b) fir_pass_a_block<D>(pl, b, yv, u[b], z);
}
// one level of the lane scan: z <- Abar^(32 2^k) z_up + z
template <int D>
TGP_HD Vec<D> fir_scan_level(const FirPlan<D>& pl, int k, const Vec<D>& z, const Vec<D>& zu) {
    Vec<D> zn;
TGP_UNROLL
    for (int i = 0; i < D; ++i) {
        double s = z[i];
TGP_UNROLL
        for (int j = 0; j < D; ++j) s = fma(pl.P2[k][i][j], zu[j], s);
        zn[i] = s;
    }
    return zn;
}
// carry += Abar^(1024 k) tot_k
template <int D>
TGP_HD void fir_carry_add(const FirPlan<D>& pl, int k, const Vec<D>& tk, Vec<D>& c) {
TGP_UNROLL
    for (int i = 0; i < D; ++i)
TGP_UNROLL
        for (int j = 0; j < D; ++j) c[i] = fma(pl.PT[k][i][j], tk[j], c[i]);
}
// pass B, data-only half (needs no state, so it runs while the carry is in flight): y_j <- y_j - kap_j - sum_{i<j} g_{j-1-i} y_i,
// in place, j descending inside each block.
template <int D, bool ZM = false>
TGP_HD void fir_pass_b1_block(const FirPlan<D>& pl, int b, double (&yv)[kFirL]) {
TGP_UNROLL
    for (int j = kFirB - 1; j >= 0; --j) {
        double v = ZM ? yv[b * kFirB + j] : yv[b * kFirB + j] - pl.kap[j];
TGP_UNROLL
        for (int i = 0; i < j; ++i) v = fma(-pl.g[j - 1 - i], yv[b * kFirB + i], v);
        yv[b * kFirB + j] = v;
    }
}
template <int D>
TGP_HD void fir_pass_b1(const FirPlan<D>& pl, double (&yv)[kFirL]) {
TGP_UNROLL
    for (int b = 0; b < kFirNBlk; ++b) fir_pass_b1_block<D>(pl, b, yv);
}
// pass B, state half: v_j = (the above) - (w'Abar^j) m from the true block-start state m; m <- Abar^8 m + u_b. Returns sum v^2 over the
// first nvalid steps of the run.
template <int D, bool TAIL>
TGP_HD double fir_pass_b2(const FirPlan<D>& pl, const double (&yv)[kFirL], const Vec<D> (&u)[kFirNBlk], Vec<D> m, int nvalid) {
    double q = 0.0, q1 = 0.0;     // two chains: the sum of squares is not the critical path
TGP_UNROLL
    for (int b = 0; b < kFirNBlk; ++b) {
TGP_UNROLL
        for (int j = 0; j < kFirB; ++j) {
            double v = yv[b * kFirB + j];
TGP_UNROLL
            for (int i = 0; i < D; ++i) v = fma(-pl.wA[j][i], m[i], v);
            if (TAIL) {
                if (b * kFirB + j < nvalid) q = fma(v, v, q);
            } else if (j & 1) {
                q1 = fma(v, v, q1);
            } else {
                q = fma(v, v, q);
            }
        }
        Vec<D> mn;
TGP_UNROLL
        for (int i = 0; i < D; ++i) {
            double s = u[b][i];
TGP_UNROLL
            for (int j = 0; j < D; ++j) s = fma(pl.A8[i][j], m[j], s);
            mn[i] = s;
        }
        m = mn;
    }
    return q + q1;
}

// =====================================================================================================================
// Host side: the plan (data-free model set-up, like lgssm_components in the reference: lti_sde.jl:131-174) and the launch.
// =====================================================================================================================
constexpr int64_t kFirMaxN0 = 16384;     // transient budget (steps) before the model is handed to the general scan
constexpr double kFirForget = 1e-18;     // |Abar^(1024 nb)| below this: the carry of older tiles is dropped

template <int D>
struct FirHostPlan {
    FirPlan<D> dev;
    std::vector<double> upload;   // [N0][D + 1] transient table, then [D*D][32] lane powers
    int status = 0;               // 0 ok, 1 not applicable (no convergence / slow forgetting / too short), 2 not positive definite
    long long bad_step = -1;
    long long N0conv = 0;
};

template <int D> inline Mat<D> fir_pow2k(Mat<D> X, int k) { for (int i = 0; i < k; ++i) X = matmul(X, X); return X; }
template <int D> inline double fir_maxabs(const Mat<D>& X) { double m = 0.0; for (int i = 0; i < D * D; ++i) m = fmax(m, fabs(X.v[i])); return m; }

// All pointers are HOST arrays laid out as in tgp_lgssm (column-major, time-invariant). first_shard: run the transient from (m0, P0);
// otherwise every step of the shard is steady (N0 = 0) and only the fixed point is needed. y_addr: device address of y (alignment).
template <int D>
void fir_build_plan(const double* hA, const double* ha, const double* hQ, const double* hH, double hh_, double R, const double* hm0,
                    const double* hP0, int64_t T, double tol, bool first_shard, unsigned long long y_addr, FirHostPlan<D>* out) {
    FirHostPlan<D>& P = *out;
    P.status = 1;
    Mat<D> A;
    Vec<D> a, H, m0;
    Sym<D> Q, Pf;
    for (int i = 0; i < D * D; ++i) A.v[i] = hA[i];
    for (int i = 0; i < D; ++i) { a[i] = ha[i]; H[i] = hH[i]; m0[i] = hm0[i]; }
    for (int j = 0; j < D; ++j)
        for (int i = 0; i <= j; ++i) { Q(i, j) = hQ[i + D * j]; Pf(i, j) = hP0[i + D * j]; }
    // ---- covariance recursion (predict LGC:46-52, update LGC:247-257) until P stops moving --------------------------------
    std::vector<double> tab;
    tab.reserve(4096);
    double sumlogS = 0.0;
    long long N0conv = -1;
    Vec<D> K;
    double S = 0.0;
    for (long long t = 0; t < kFirMaxN0 && t < T; ++t) {
        const Sym<D> Pp = congruence(A, Pf, Q);
        const Vec<D> V = symvec(Pp, H);
        S = dot(V, H) + R;
        if (!(S > 1e-300) || !(S < 1e300)) { P.status = 2; P.bad_step = t; return; }
        const double is = 1.0 / sqrt(S);
        Vec<D> B;
        for (int i = 0; i < D; ++i) { B[i] = V[i] * is; K[i] = B[i] * is; }
        Sym<D> Pn;
        double err = 0.0, nrm = 0.0;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i <= j; ++i) {
                Pn(i, j) = fma(-B[i], B[j], Pp(i, j));
                err = fmax(err, fabs(Pn(i, j) - Pf(i, j)));
                nrm = fmax(nrm, fabs(Pn(i, j)));
            }
        Pf = Pn;
        for (int i = 0; i < D; ++i) tab.push_back(K[i]);
        tab.push_back(is * is);
        sumlogS += log(S);
        if (err <= tol * nrm) { N0conv = t + 1; break; }
    }
    if (N0conv < 0) return;                 // no fixed point within the budget: general scan
    P.N0conv = N0conv;
    // The transient ends where successive P differ by tol; the steady constants come from the fixed point itself: keep iterating
    // (data-free, a few hundred 3x3 steps) until P stops moving in double precision, so that the frozen gain carries no O(tol) bias.
    {
        double prev = 1e300;
        for (long long t = 0; t < 4 * kFirMaxN0; ++t) {
            const Sym<D> Pp = congruence(A, Pf, Q);
            const Vec<D> V = symvec(Pp, H);
            const double Sx = dot(V, H) + R;
            const double is = 1.0 / sqrt(Sx);
            double err = 0.0, nrm = 0.0;
            Sym<D> Pn;
            for (int j = 0; j < D; ++j)
                for (int i = 0; i <= j; ++i) {
                    Pn(i, j) = fma(-(V[i] * is), V[j] * is, Pp(i, j));
                    err = fmax(err, fabs(Pn(i, j) - Pf(i, j)));
                    nrm = fmax(nrm, fabs(Pn(i, j)));
                }
            Pf = Pn;
            if (err <= 4e-16 * nrm || (err >= prev && err <= 1e-14 * nrm)) break;
            prev = err;
        }
    }
    // ---- constants of the steady recursion: one more step from the converged P ---------------------------------------------
    {
        const Sym<D> Pp = congruence(A, Pf, Q);
        const Vec<D> V = symvec(Pp, H);
        S = dot(V, H) + R;
        for (int i = 0; i < D; ++i) K[i] = V[i] / S;
    }
    const Vec<D> w = matTvec(A, H);
    const double hh = dot(H, a) + hh_;
    Mat<D> Ab;
    Vec<D> c;
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < D; ++i) Ab(i, j) = fma(-K[i], w[j], A(i, j));
    for (int i = 0; i < D; ++i) c[i] = fma(-K[i], hh, a[i]);
    FirPlan<D>& d = P.dev;
    Mat<D> pw[kFirB + 1];                   // Abar^j
    pw[0] = meye<D>();
    for (int j = 1; j <= kFirB; ++j) pw[j] = matmul(Ab, pw[j - 1]);
    Vec<D> zc = vzero<D>();
    for (int j = 0; j < kFirB; ++j) {
        const Vec<D> gk = matvec(pw[kFirB - 1 - j], K), ck = matvec(pw[kFirB - 1 - j], c), wa = matTvec(pw[j], w);
        for (int i = 0; i < D; ++i) { d.gK[j][i] = gk[i]; zc[i] += ck[i]; d.wA[j][i] = wa[i]; }
        d.g[j] = dot(wa, K);
        Vec<D> sc = vzero<D>();             // sum_{i<j} Abar^(j-1-i) c
        for (int i = 0; i < j; ++i) { const Vec<D> t = matvec(pw[j - 1 - i], c); for (int k = 0; k < D; ++k) sc[k] += t[k]; }
        d.kap[j] = hh + dot(w, sc);
    }
    for (int i = 0; i < D; ++i) {
        d.zc[i] = zc[i];
        for (int j = 0; j < D; ++j) d.A8[i][j] = pw[kFirB](i, j);
    }
    Mat<D> p32 = fir_pow2k(pw[kFirB], 2);   // Abar^32
    {
        Mat<D> p = p32;
        for (int k = 0; k < 5; ++k) {
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) d.P2[k][i][j] = p(i, j);
            p = matmul(p, p);
        }
        // p = Abar^1024
        Mat<D> q = meye<D>();
        int nb = 0;
        for (int k = 0; k < kFirNbMax; ++k) {
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) d.PT[k][i][j] = q(i, j);
            q = matmul(p, q);
            if (nb == 0 && fir_maxabs(q) <= kFirForget) nb = k + 1;
        }
        if (nb == 0) return;                // forgets too slowly for a 3-tile look-back: the two-phase kernel handles it
        d.nb = nb;
    }
    for (int i = 0; i < D; ++i) {
        d.a[i] = a[i]; d.w[i] = w[i]; d.m0[i] = m0[i];
        for (int j = 0; j < D; ++j) d.A[i][j] = A(i, j);
    }
    d.hh = hh;
    d.invS = 1.0 / S;
    d.zero_mean = 1;
    for (int i = 0; i < D; ++i) if (d.zc[i] != 0.0) d.zero_mean = 0;
    for (int j = 0; j < kFirB; ++j) if (d.kap[j] != 0.0) d.zero_mean = 0;
    // ---- extent of the transient: converged, and the first steady step 32-byte aligned ------------------------------------
    long long N0 = 0;
    if (first_shard) {
        N0 = N0conv;
        while (((y_addr >> 3) + (unsigned long long)N0) & 3ull) ++N0;
    }
    if (T < N0 + 2 * kFirTile) return;      // too short to be worth it
    for (long long t = N0conv; t < N0; ++t) {   // padding steps run with the converged gain
        for (int i = 0; i < D; ++i) tab.push_back(K[i]);
        tab.push_back(1.0 / S);
        sumlogS += log(S);
    }
    if (!first_shard) { tab.clear(); sumlogS = 0.0; }
    d.N0 = N0;
    d.T = T;
    d.ntiles = (T - N0 + kFirTile - 1) / kFirTile;
    d.aligned = ((((y_addr >> 3) + (unsigned long long)N0) & 3ull) == 0) ? 1 : 0;
    d.c0 = -0.5 * ((double)T * kLog2Pi + sumlogS + (double)(T - N0) * log(S));
    P.upload.assign(tab.begin(), tab.end());
    P.upload.resize((size_t)N0 * (D + 1) + (size_t)D * D * 32);
    double* plane = P.upload.data() + (size_t)N0 * (D + 1);
    Mat<D> pl = meye<D>();
    for (int l = 0; l < 32; ++l) {
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) plane[(i * D + j) * 32 + l] = pl(i, j);
        pl = matmul(p32, pl);
    }
    P.status = 0;
}


}  // namespace tgp

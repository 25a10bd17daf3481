// emul_fir.cpp — g++-compiled, single-threaded emulation of the one-launch steady-state logpdf kernel (tgp_fir.cuh), built from the
// SAME plan builder and per-lane arithmetic (tgp_fir_plan.h: fir_build_plan, fir_pass_a, fir_scan_level, fir_carry_add,
// fir_pass_b1, fir_pass_b2); the warp shuffles, the tile words and the look-back are replaced by arrays. TEST HARNESS ONLY.
#include <cstdint>
#include <vector>

#include "../../temporalgps.jl_b200/csrc/tgp_fir_plan.h"

using namespace tgp;

template <int D>
static void tile_pass_a(const FirPlan<D>& pl, const double* ys, int nvalid, double (*yv)[kFirL], Vec<D> (*u)[kFirNBlk], Vec<D>* z) {
    for (int lane = 0; lane < 32; ++lane) {
        for (int j = 0; j < kFirL; ++j) {
            const int e = lane * kFirL + j;
            yv[lane][j] = e < nvalid ? ys[e] : 0.0;
        }
        z[lane] = vzero<D>();
        fir_pass_a<D>(pl, yv[lane], u[lane], z[lane]);
    }
    for (int k = 0; k < 5; ++k) {     // Kogge-Stone, as the shuffles do it
        Vec<D> zn[32];
        for (int lane = 0; lane < 32; ++lane) zn[lane] = lane >= (1 << k) ? fir_scan_level<D>(pl, k, z[lane], z[lane - (1 << k)]) : z[lane];
        for (int lane = 0; lane < 32; ++lane) z[lane] = zn[lane];
    }
}

template <int D>
static int run(const double* A, const double* a, const double* Q, const double* H, double h, double R, const double* m0, const double* P0,
               int64_t T, const double* y, double tol, int align_words, int first_shard, const double* halo, double* lml_out, int64_t* info) {
    FirHostPlan<D> hp;
    fir_build_plan<D>(A, a, Q, H, h, R, m0, P0, T, tol, first_shard != 0, (unsigned long long)align_words << 3, &hp);
    info[0] = hp.status; info[1] = hp.dev.N0; info[2] = hp.dev.nb; info[3] = hp.N0conv; info[4] = hp.bad_step;
    if (hp.status != 0) return hp.status;
    const FirPlan<D>& pl = hp.dev;
    const double* tab = hp.upload.data();
    const double* plane = tab + (size_t)pl.N0 * (D + 1);
    std::vector<Vec<D>> agg(pl.ntiles + kFirNbMax, vzero<D>());
    double qh = 0.0;
    static double yv[32][kFirL];
    static Vec<D> u[32][kFirNBlk];
    Vec<D> z[32];
    if (first_shard) {     // the transient: sequential here (the kernel's warp runs the same recursion as a scan over affine maps)
        Vec<D> m;
        for (int i = 0; i < D; ++i) m[i] = pl.m0[i];
        for (int64_t t = 0; t < pl.N0; ++t) {
            double v = y[t] - pl.hh;
            for (int i = 0; i < D; ++i) v -= pl.w[i] * m[i];
            qh += v * v * tab[t * (D + 1) + D];
            Vec<D> mn;
            for (int i = 0; i < D; ++i) {
                double s = pl.a[i];
                for (int j = 0; j < D; ++j) s += pl.A[i][j] * m[j];
                mn[i] = s + tab[t * (D + 1) + i] * v;
            }
            m = mn;
        }
        agg[kFirNbMax - 1] = m;
    } else {
        for (int k = pl.nb; k >= 1; --k) {
            tile_pass_a<D>(pl, halo + (size_t)(pl.nb - k) * kFirTile, kFirTile, yv, u, z);
            agg[kFirNbMax - k] = z[31];
        }
    }
    const int64_t Ts = pl.T - pl.N0;
    double q = 0.0;
    for (int64_t t = 0; t < pl.ntiles; ++t) {
        const int nvalid = (int)std::min<int64_t>(kFirTile, Ts - t * kFirTile);
        tile_pass_a<D>(pl, y + pl.N0 + t * kFirTile, nvalid, yv, u, z);
        agg[t + kFirNbMax] = z[31];
        Vec<D> c = agg[t + kFirNbMax - 1];
        for (int k = 1; k < pl.nb; ++k) fir_carry_add<D>(pl, k, agg[t + kFirNbMax - 1 - k], c);
        for (int lane = 0; lane < 32; ++lane) {
            Vec<D> m = lane ? z[lane - 1] : vzero<D>();
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) m[i] = fma(plane[(i * D + j) * 32 + lane], c[j], m[i]);
            fir_pass_b1<D>(pl, yv[lane]);
            q += nvalid == kFirTile ? fir_pass_b2<D, false>(pl, yv[lane], u[lane], m, kFirL)
                                    : fir_pass_b2<D, true>(pl, yv[lane], u[lane], m, nvalid - lane * kFirL);
        }
    }
    *lml_out = pl.c0 - 0.5 * (qh + pl.invS * q);
    return 0;
}

extern "C" int emul_fir_logpdf(int D, const double* A, const double* a, const double* Q, const double* H, double h, double R, const double* m0,
                               const double* P0, int64_t T, const double* y, double tol, int align_words, int first_shard, const double* halo,
                               double* lml_out, int64_t* info) {
    switch (D) {
        case 1: return run<1>(A, a, Q, H, h, R, m0, P0, T, y, tol, align_words, first_shard, halo, lml_out, info);
        case 2: return run<2>(A, a, Q, H, h, R, m0, P0, T, y, tol, align_words, first_shard, halo, lml_out, info);
        case 3: return run<3>(A, a, Q, H, h, R, m0, P0, T, y, tol, align_words, first_shard, halo, lml_out, info);
        case 4: return run<4>(A, a, Q, H, h, R, m0, P0, T, y, tol, align_words, first_shard, halo, lml_out, info);
        default: return -1;
    }
}

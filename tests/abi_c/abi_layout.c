/* abi_layout.c — drives libtgpb200.so from plain C with buffers laid out exactly as the Julia glue
 * (temporalgps.jl_b200/julia/TemporalGPsB200.jl) hands them over — the only stand-in for executing that glue here:
 *   time-invariant model   `Fill` parameters: ONE element each, stride 0            (lti_sde.jl:148-160)
 *   time-varying model     Vector{SMatrix{D,D}} / Vector{SVector{D}}: reinterpret(Float64, v), column-major blocks back to back
 *   _filter output         Vector{Gaussian{SVector{D},SMatrix{D,D}}}: records of D + D*D doubles written IN PLACE
 *                          (m_f = base, P_f = base + D, s_m = s_P = D + D*D)
 * Input file (written by tests/test_gpu_abi_c.py, little-endian doubles): D, T, tv, then A (tv ? T : 1 blocks, column-major), a, Q,
 * H, h, R, m0, P0, y (T), expected lml, expected filter records (T x (D + D*D)). Exit code 0 = every check passed.
 * TEST HARNESS ONLY. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tgp_b200.h"

static double* rd(FILE* f, size_t n) {
    double* p = (double*)malloc(n * sizeof(double));
    if (!p || fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    double hdr[3];
    if (fread(hdr, sizeof(double), 3, f) != 3) return 2;
    const int D = (int)hdr[0], tv = (int)hdr[2];
    const int64_t T = (int64_t)hdr[1];
    const size_t n = tv ? (size_t)T : 1;
    double *A = rd(f, n * D * D), *a = rd(f, n * D), *Q = rd(f, n * D * D), *H = rd(f, n * D), *h = rd(f, n), *R = rd(f, n);
    double *m0 = rd(f, D), *P0 = rd(f, D * D), *y = rd(f, T);
    double* lml_ref = rd(f, 1);
    const size_t rec = (size_t)D + (size_t)D * D;
    double* filt_ref = rd(f, (size_t)T * rec);
    fclose(f);

    tgp_handle hd = NULL;
    if (tgp_create(&hd, 0) != TGP_OK) { fprintf(stderr, "tgp_create: %s\n", tgp_last_error(NULL)); return 3; }
    tgp_lgssm m;
    memset(&m, 0, sizeof m);
    m.D = D; m.M = 1; m.T = T; m.ordering = TGP_FORWARD; m.R_kind = TGP_R_SCALAR;
    m.A = A; m.sA = tv ? D * D : 0;      /* Fill -> stride 0; Vector{SMatrix} -> D*D doubles per step */
    m.a = a; m.sa = tv ? D : 0;
    m.Q = Q; m.sQ = tv ? D * D : 0;
    m.H = H; m.sH = tv ? D : 0;
    m.h = h; m.sh = tv ? 1 : 0;
    m.R = R; m.sR = tv ? 1 : 0;
    m.m0 = m0; m.P0 = P0;
    int bad = 0;
    double lml = 0.0;
    int rc = tgp_logpdf(hd, &m, y, &lml, NULL);
    if (rc != TGP_OK) { fprintf(stderr, "tgp_logpdf rc=%d: %s\n", rc, tgp_last_error(hd)); return 4; }
    if (fabs(lml - *lml_ref) > 1e-6 * fabs(*lml_ref)) { fprintf(stderr, "lml %.17g vs %.17g\n", lml, *lml_ref); bad = 1; }
    /* Vector{Gaussian} records, written in place */
    double* recs = (double*)malloc((size_t)T * rec * sizeof(double));
    double lml2 = 0.0;
    rc = tgp_filter(hd, &m, y, recs, (int64_t)rec, recs + D, (int64_t)rec, &lml2);
    if (rc != TGP_OK) { fprintf(stderr, "tgp_filter rc=%d: %s\n", rc, tgp_last_error(hd)); return 5; }
    double worst = 0.0;
    for (size_t i = 0; i < (size_t)T * rec; ++i) {
        const double e = fabs(recs[i] - filt_ref[i]) / (fabs(filt_ref[i]) + 1e-8);
        if (e > worst) worst = e;
    }
    if (worst > 1e-5) { fprintf(stderr, "filter records differ: %.3g\n", worst); bad = 1; }
    if (fabs(lml2 - *lml_ref) > 1e-6 * fabs(*lml_ref)) { fprintf(stderr, "filter lml %.17g vs %.17g\n", lml2, *lml_ref); bad = 1; }
    /* error mapping: a length mismatch is the caller's to check (lgssm.jl:202-208); a negative noise variance must come back as
     * TGP_ENOTPD with the failing index in the message */
    double Rneg = -10.0;
    tgp_lgssm mb = m;
    mb.R = &Rneg; mb.sR = 0;
    rc = tgp_logpdf(hd, &mb, y, &lml, NULL);
    if (rc != TGP_ENOTPD || !strstr(tgp_last_error(hd), "time index")) { fprintf(stderr, "expected TGP_ENOTPD with an index, got %d: %s\n", rc, tgp_last_error(hd)); bad = 1; }
    tgp_destroy(hd);
    printf("abi_layout: D=%d T=%lld tv=%d lml=%.12g worst filter rel err %.3g %s\n", D, (long long)T, tv, lml, worst, bad ? "FAILED" : "ok");
    return bad;
}

/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C restatement of the reference's Partridge-Schwenke water potential:
 *   calc_hoh_pot : sample_potentials/FortPots/Partridge_Schwenke_H2O/calc_h2o_pot.f:1-34
 *   vibpot       : sample_potentials/FortPots/Partridge_Schwenke_H2O/h2opes_v2.f:1-491
 * The Fortran cannot be compiled in this image (no gfortran), so this file follows it
 * statement by statement, keeps its literal constants (4.556335d-6, 0.529177249d0,
 * acos(-1d0)/1.8d2) and its operation order, and is built with -ffp-contract=off so no
 * FMA contraction changes the rounding (the reference Makefile uses plain `gfortran -c`).
 * Pinned by tests/test_oracle_ps.py against the shipped tutorial snapshots.
 */
#include <math.h>
#include <stddef.h>
#include "../pyvibdmc_b200/csrc/ps_h2o_table.h"

typedef struct {
    double c[PS_NTERMS];
    int p[PS_NTERMS][3];
    double reoh, b1, ce, phh1, phh2, deoh, roh, alphaoh;
    int ready;
} PsFolded;

static PsFolded g_ps;

/* h2opes_v2.f:402-459 -- one-time coefficient folding + unit conversion */
static void ps_setup(PsFolded *s)
{
    double reoh = PS_RAW_REOH, thetae = PS_RAW_THETAE, b1 = PS_RAW_B1, roh = PS_RAW_ROH;
    double alphaoh = PS_RAW_ALPHAOH, deoh = PS_RAW_DEOH, phh1 = PS_RAW_PHH1, phh2 = PS_RAW_PHH2;
    const double f5z = PS_RAW_F5Z, fbasis = PS_RAW_FBASIS, fcore = PS_RAW_FCORE, frest = PS_RAW_FREST;
    for (int i = 0; i < PS_NTERMS; ++i) {            /* :421-424 (left-to-right sum) */
        const PsTermRaw *t = &PS_RAW_TERMS[i];
        s->c[i] = f5z * t->c5z + fbasis * t->cbasis + fcore * t->ccore + frest * t->crest;
        s->p[i][0] = t->p1; s->p[i][1] = t->p2; s->p[i][2] = t->p3;
    }
    phh1 = phh1 * f5z;                                /* :431 */
    deoh = deoh * f5z;                                /* :432 */
    reoh = reoh / 0.529177249;                        /* :443 */
    b1 = b1 * 0.529177249 * 0.529177249;              /* :444 */
    for (int i = 0; i < PS_NTERMS; ++i) s->c[i] = s->c[i] * 4.556335e-6;   /* :445-447 */
    const double rad = acos(-1.0) / 1.8e2;            /* :448 */
    s->ce = cos(thetae * rad);                        /* :449 */
    phh1 = phh1 * exp(phh2);                          /* :450 */
    phh1 = phh1 * 4.556335e-6;                        /* :451 */
    phh2 = phh2 * 0.529177249;                        /* :452 */
    deoh = deoh * 4.556335e-6;                        /* :453 */
    roh = roh / 0.529177249;                          /* :454 */
    alphaoh = alphaoh * 0.529177249;                  /* :455 */
    s->c[0] = s->c[0] * 2.0;                          /* :456 */
    s->reoh = reoh; s->b1 = b1; s->phh1 = phh1; s->phh2 = phh2;
    s->deoh = deoh; s->roh = roh; s->alphaoh = alphaoh;
    s->ready = 1;
}

/* h2opes_v2.f:460-489 for one geometry given (r1, r2, theta) */
static double ps_vibpot_one(const PsFolded *s, double r1, double r2, double th)
{
    double fm[15][3];
    const double x1 = (r1 - s->reoh) / s->reoh;
    const double x2 = (r2 - s->reoh) / s->reoh;
    const double x3 = cos(th) - s->ce;
    const double rhh = sqrt(r1 * r1 + r2 * r2 - 2.0 * r1 * r2 * cos(th));
    const double vhh = s->phh1 * exp(-s->phh2 * rhh);
    double ex = exp(-s->alphaoh * (r1 - s->roh));
    const double voh1 = s->deoh * ex * (ex - 2.0);
    ex = exp(-s->alphaoh * (r2 - s->roh));
    const double voh2 = s->deoh * ex * (ex - 2.0);
    fm[0][0] = fm[0][1] = fm[0][2] = 1.0;
    for (int j = 1; j < 15; ++j) {
        fm[j][0] = fm[j - 1][0] * x1;
        fm[j][1] = fm[j - 1][1] * x2;
        fm[j][2] = fm[j - 1][2] * x3;
    }
    double v = 0.0;
    for (int j = 1; j < PS_NTERMS; ++j) {
        const int a = s->p[j][0], b = s->p[j][1], c = s->p[j][2];
        const double term = s->c[j] * (fm[a][0] * fm[b][1] + fm[b][0] * fm[a][1]) * fm[c][2];
        v = v + term;
    }
    const double d1 = r1 - s->reoh, d2 = r2 - s->reoh;
    return v * exp(-s->b1 * (d1 * d1 + d2 * d2)) + s->c[0] + voh1 + voh2 + vhh;
}

/* calc_h2o_pot.f:12-29 -- atoms ordered H, H, O; xyz is the NumPy (n,3,3) C-order array */
void oracle_ps_h2o(const double *xyz, long n, double *v)
{
    if (!g_ps.ready) ps_setup(&g_ps);
    for (long k = 0; k < n; ++k) {
        const double *w = xyz + 9 * k;
        double r1 = 0.0, r2 = 0.0, ct = 0.0;
        for (int j = 0; j < 3; ++j) {
            const double d1 = w[6 + j] - w[j];
            const double d2 = w[6 + j] - w[3 + j];
            r1 = r1 + d1 * d1;
            r2 = r2 + d2 * d2;
            ct = ct + d1 * d2;
        }
        r1 = sqrt(r1);
        r2 = sqrt(r2);
        const double th = acos(ct / r1 / r2);
        v[k] = ps_vibpot_one(&g_ps, r1, r2, th);
    }
}

/* folded parameters, for cross-checking the CUDA host-side folding */
void oracle_ps_folded(double *c245, double *scal8)
{
    if (!g_ps.ready) ps_setup(&g_ps);
    for (int i = 0; i < PS_NTERMS; ++i) c245[i] = g_ps.c[i];
    scal8[0] = g_ps.reoh; scal8[1] = g_ps.b1; scal8[2] = g_ps.ce; scal8[3] = g_ps.phh1;
    scal8[4] = g_ps.phh2; scal8[5] = g_ps.deoh; scal8[6] = g_ps.roh; scal8[7] = g_ps.alphaoh;
}

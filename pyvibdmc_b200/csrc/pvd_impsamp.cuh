// Importance sampling (pyvibdmc.py:549-612, simulation_utilities/imp_samp.py:21-76):
// trial wave functions, finite-difference drift / local kinetic energy, Metropolis step.
#pragma once
#include "pvd_step.cuh"

struct TrialParamsDev {
    // water product wfn (call_trl_h2o.py:7-78): table rows grid / psi, Gaussian bend
    const double *grid;
    const double *wfn;
    const double *slope;        // (wfn[j+1] - wfn[j]) / (grid[j+1] - grid[j]), formed once on the host exactly as np.interp forms it
    const double *dwfn, *d2wfn;       // analytic derivatives (call_trl_h2o.py:17-18): tabulated psi', psi'' and their slopes
    const double *dslope, *d2slope;
    int ntab;
    double g0, inv_step;
    double g_last, w_first, w_last;
    double ang_alpha, theta_eq, ang_pref;
    // 1-D Gaussian (harm_trial_wfn.py:6-40)
    double h_alpha, h_pref;
    double fd_dx, fd_dx2;       // 0.001 and 0.001**2 as Python evaluates them (imp_samp.py:58,75)
};

// np.interp(x, grid, wfn): linear, clamped to the end values, no FMA (NumPy's C loop is plain mul/add)
__device__ __forceinline__ double interp_rows(double x, const TrialParamsDev &p, const double *row, const double *slope)
{
    const int n = p.ntab;
    if (x != x) return x;
    if (x > p.g_last) return row[n - 1];
    if (x < p.g0) return row[0];
    int j = (int)((x - p.g0) * p.inv_step);
    j = j < 0 ? 0 : (j > n - 2 ? n - 2 : j);
    while (j > 0 && x < p.grid[j]) --j;
    while (j < n - 1 && x >= p.grid[j + 1]) ++j;
    const double fj = row[j];
    if (j == n - 1) return fj;
    const double xj = p.grid[j];
    if (x == xj) return fj;
    return __dadd_rn(__dmul_rn(slope[j], x - xj), fj);
}
__device__ __forceinline__ double interp_table(double x, const TrialParamsDev &p) { return interp_rows(x, p, p.wfn, p.slope); }

__device__ __forceinline__ double norm3(double a, double b, double c)
{
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
}

// psi = interp(r_OH1) * interp(r_OH2) * Gaussian(theta), kwargs {'dists':[[0,2],[2,1]],'angs':[[0,2,1]]}
struct TrialH2O {
    static constexpr int NC = 9;
    static constexpr int NDIM = 3;
    static constexpr bool ANALYTIC = false;
    // a finite-difference stencil point that moves ONE hydrogen leaves the other O-H bond (its length and its table value)
    // untouched: psi_moved<0 / 1> takes them from `keep` -- the same operations on the same operands, hence the same bits as a
    // full evaluation -- and a third of the stencil's square roots and table look-ups disappear (trial_drift_ke)
#ifndef PVD_FD_PARTIAL
#define PVD_FD_PARTIAL 1                                  // A/B switch: 0 = every stencil point evaluates both bonds
#endif
    static constexpr bool PARTIAL = PVD_FD_PARTIAL != 0;
    struct Bond { double r, t; };
    template <int WHICH>                                 // 0: H1-O, 1: H2-O
    __device__ static __forceinline__ Bond bond(const double (&x)[9], const TrialParamsDev &p)
    {
        Bond b;
        b.r = norm3(x[3 * WHICH] - x[6], x[3 * WHICH + 1] - x[7], x[3 * WHICH + 2] - x[8]);
        b.t = interp_table(b.r, p);
        return b;
    }
    // MOVED: 0 = H1 moved since `keep` (= bond<1>) was formed, 1 = H2 moved (keep = bond<0>), 2 = evaluate everything.
    // PVD_FD_NOINLINE: the evaluation is ONE out-of-line function (scalar arguments in registers, `moved` a warp-uniform run-time
    // value) instead of a copy per call site: with the copies that inlining the stencil made, k_imp_move was 140-200 KB of code and
    // instruction fetch was its largest stall on an equilibrated ensemble, where the warps of an SM are spread over all of it
    // (no_instruction 1.5-2.5 stalls per issue, profiles/r02_profiles.md).
#ifndef PVD_FD_NOINLINE
#define PVD_FD_NOINLINE 1
#endif
    __device__ static __forceinline__ double psi_core(int moved, double x0, double x1, double x2, double x3, double x4, double x5,
                                                      double x6, double x7, double x8, double keep_r, double keep_t, const TrialParamsDev &p)
    {
        const double ax = x0 - x6, ay = x1 - x7, az = x2 - x8;      // H1 - O
        const double bx = x3 - x6, by = x4 - x7, bz = x5 - x8;      // H2 - O
        double r1, r2, t1, t2;
        if (moved == 1) { r1 = keep_r; t1 = keep_t; } else { r1 = norm3(ax, ay, az); t1 = interp_table(r1, p); }
        if (moved == 0) { r2 = keep_r; t2 = keep_t; } else { r2 = norm3(bx, by, bz); t2 = interp_table(r2, p); }
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
        const double th = acos(dot / __dmul_rn(r1, r2));
        const double dth = th - p.theta_eq;
        const double ang = __dmul_rn(p.ang_pref, exp(__dmul_rn(-p.ang_alpha, __dmul_rn(dth, dth)) / 2.0));
        return __dmul_rn(__dmul_rn(t1, t2), ang);
    }
    __device__ static __noinline__ double psi_call(int moved, double x0, double x1, double x2, double x3, double x4, double x5,
                                                   double x6, double x7, double x8, double keep_r, double keep_t, const TrialParamsDev *p)
    {
        return psi_core(moved, x0, x1, x2, x3, x4, x5, x6, x7, x8, keep_r, keep_t, *p);
    }
    template <int MOVED>
    __device__ static __forceinline__ double psi_moved(const double (&x)[9], const TrialParamsDev &p, const Bond &keep)
    {
#if PVD_FD_NOINLINE
        return psi_call(MOVED, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8], keep.r, keep.t, &p);
#else
        return psi_core(MOVED, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8], keep.r, keep.t, p);
#endif
    }
    __device__ static __forceinline__ double psi(const double (&x)[9], const TrialParamsDev &p)
    {
        const double ax = x[0] - x[6], ay = x[1] - x[7], az = x[2] - x[8];      // H1 - O
        const double bx = x[3] - x[6], by = x[4] - x[7], bz = x[5] - x[8];      // H2 - O
        const double r1 = norm3(ax, ay, az), r2 = norm3(bx, by, bz);
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
        const double th = acos(dot / __dmul_rn(r1, r2));
        const double dth = th - p.theta_eq;
        const double ang = __dmul_rn(p.ang_pref, exp(__dmul_rn(-p.ang_alpha, __dmul_rn(dth, dth)) / 2.0));
        return __dmul_rn(__dmul_rn(interp_table(r1, p), interp_table(r2, p)), ang);
    }
};

// The same product wave function with the reference's ANALYTIC derivatives (call_trl_h2o.py:101-149: dpsi_dx, with the
// chain-rule formulas of ChainRuleHelper, imp_samp_helper.py:10-209, written out for r1 = |H1 - O|, r2 = |O - H2|,
// theta = angle(H1, O, H2)).  One evaluation instead of the 19-point finite-difference stencil.
struct TrialH2OAn {
    static constexpr int NC = 9;
    static constexpr int NDIM = 3;
    static constexpr bool ANALYTIC = true;
    static constexpr bool PARTIAL = false;
    __device__ static __forceinline__ double psi(const double (&x)[9], const TrialParamsDev &p) { return TrialH2O::psi(x, p); }
    __device__ static __forceinline__ void derivs(const double (&x)[9], const TrialParamsDev &p, double &psi0, double (&d1)[9], double (&d2)[9])
    {
        double a[3], b2[3], v2[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { a[j] = x[j] - x[6 + j]; b2[j] = x[6 + j] - x[3 + j]; v2[j] = x[3 + j] - x[6 + j]; }
        const double r1 = norm3(a[0], a[1], a[2]), r2 = norm3(b2[0], b2[1], b2[2]);
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn(a[0], v2[0]), __dmul_rn(a[1], v2[1])), __dmul_rn(a[2], v2[2]));
        const double th = acos(dot / __dmul_rn(r1, r2));
        const double cth = cos(th);
        const double xq = th - p.theta_eq, al = p.ang_alpha;
        const double e = exp(__dmul_rn(-al, __dmul_rn(xq, xq)) / 2.0);
        const double t0 = interp_table(r1, p), t1 = interp_table(r2, p), t2 = __dmul_rn(p.ang_pref, e);
        psi0 = __dmul_rn(__dmul_rn(t0, t1), t2);
        // (dpsi_q/dq) / psi_q and (d2psi_q/dq2) / psi_q for q = r1, r2, theta
        const double w1[3] = {interp_rows(r1, p, p.dwfn, p.dslope) / t0, interp_rows(r2, p, p.dwfn, p.dslope) / t1,
                              __dmul_rn(__dmul_rn(p.ang_pref, -al * xq), e) / t2};
        const double w2[3] = {interp_rows(r1, p, p.d2wfn, p.d2slope) / t0, interp_rows(r2, p, p.d2wfn, p.d2slope) / t1,
                              __dmul_rn(__dmul_rn(p.ang_pref, al * al * xq * xq - al), e) / t2};
        const double ir1 = 1.0 / r1, ir2 = 1.0 / r2, irr = 1.0 / (r1 * r2);
        const double s2 = 1.0 - cth * cth;
        const double dth_dc = -1.0 / sqrt(s2), d2th_dc2 = -1.0 * cth / (s2 * sqrt(s2));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // bond lengths: dq/dx per atom (H1, H2, O) and d2q/dx2 = 1/r - (1/r)(dq/dx)^2 on the two atoms of the bond
            const double dra0 = a[j] * ir1, dra2 = -1.0 * a[j] * ir1;               // r1 wrt H1, O
            const double drc2 = b2[j] * ir2, drc1 = -1.0 * b2[j] * ir2;              // r2 wrt O, H2
            const double d2ra0 = ir1 - ir1 * dra0 * dra0, d2ra2 = ir1 - ir1 * dra2 * dra2;
            const double d2rc2 = ir2 - ir2 * drc2 * drc2, d2rc1 = ir2 - ir2 * drc1 * drc1;
            // cos(theta): vertex O; alpha_1 = H2 - O, alpha_2 = H1 - O, alpha_3 = 2 O - H1 - H2
            const double al1 = v2[j], al2 = a[j], al3 = 2.0 * x[6 + j] - x[j] - x[3 + j];
            const double dc0 = al1 * irr - (cth * ir1) * dra0;
            const double dc1 = al2 * irr - (cth * ir2) * drc1;
            const double dc2 = al3 * irr - (cth * ir1) * dra2 - (cth * ir2) * drc2;
            const double d2c0 = (-2.0 * al1) * (irr * ir1) * dra0 + (2.0 * cth * ir1 * ir1) * dra0 * dra0 + (-1.0 * cth * ir1) * d2ra0;
            const double d2c1 = (-2.0 * al2) * (irr * ir2) * drc1 + (2.0 * cth * ir2 * ir2) * drc1 * drc1 + (-1.0 * cth * ir2) * d2rc1;
            const double d2c2 = 2.0 * irr + (-1.0 * cth * ir1) * d2ra2 + (-1.0 * cth * ir2) * d2rc2 + (2.0 * cth * ir1 * ir1) * dra2 * dra2
                                + (2.0 * cth * ir2 * ir2) * drc2 * drc2 + (-2.0 * al3 * (irr * ir1)) * dra2 + (-2.0 * al3 * (irr * ir2)) * drc2
                                + (2.0 * cth * irr) * dra2 * drc2;
            // internal-coordinate derivatives per atom: q = (r1, r2, theta)
            const double dq[3][3] = {{dra0, 0.0, dth_dc * dc0}, {0.0, drc1, dth_dc * dc1}, {dra2, drc2, dth_dc * dc2}};   // [atom][q]
            const double d2q[3][3] = {{d2ra0, 0.0, dc0 * dc0 * d2th_dc2 + d2c0 * dth_dc},
                                      {0.0, d2rc1, dc1 * dc1 * d2th_dc2 + d2c1 * dth_dc},
                                      {d2ra2, d2rc2, dc2 * dc2 * d2th_dc2 + d2c2 * dth_dc}};
#pragma unroll
            for (int atom = 0; atom < 3; ++atom) {
                const double g0 = dq[atom][0], g1 = dq[atom][1], g2 = dq[atom][2];
                const double h0 = d2q[atom][0], h1 = d2q[atom][1], h2 = d2q[atom][2];
                d1[3 * atom + j] = g0 * w1[0] + g1 * w1[1] + g2 * w1[2];
                const double term1 = g0 * g0 * w2[0] + g1 * g1 * w2[1] + g2 * g2 * w2[2];
                const double term2 = h0 * w1[0] + h1 * w1[1] + h2 * w1[2];
                const double term3 = 2.0 * ((g0 * g1) * (w1[0] * w1[1]) + (g1 * g2) * (w1[1] * w1[2]) + (g2 * g0) * (w1[2] * w1[0]));
                d2[3 * atom + j] = term1 + term2 + term3;
            }
        }
    }
};

struct TrialHarm1D {
    static constexpr int NC = 1;
    static constexpr int NDIM = 1;
    static constexpr bool ANALYTIC = true;
    static constexpr bool PARTIAL = false;
    __device__ static __forceinline__ double psi(const double (&x)[1], const TrialParamsDev &p)
    {
        return __dmul_rn(p.h_pref, exp(__dmul_rn(-p.h_alpha, __dmul_rn(x[0], x[0])) / 2.0));
    }
    // derivative() of harm_trial_wfn.py:36-40: (psi'/psi, psi''/psi) formed exactly as the reference does
    __device__ static __forceinline__ void derivs(const double (&x)[1], const TrialParamsDev &p, double &psi0, double (&d1)[1], double (&d2)[1])
    {
        const double e = exp(__dmul_rn(-p.h_alpha, __dmul_rn(x[0], x[0])) / 2.0);
        psi0 = __dmul_rn(p.h_pref, e);
        d1[0] = __dmul_rn(__dmul_rn(p.h_pref, __dmul_rn(-p.h_alpha, x[0])), e) / psi0;
        const double poly = __dadd_rn(__dmul_rn(__dmul_rn(p.h_alpha, p.h_alpha), __dmul_rn(x[0], x[0])), -p.h_alpha);
        d2[0] = __dmul_rn(__dmul_rn(p.h_pref, poly), e) / psi0;
    }
};

// ImpSamp.drift: psi, grad psi / psi, d2 psi / psi.  Finite differences follow imp_samp.py:56-76
// (dx = 1e-3, the coordinate is walked -dx, +2dx, -dx in place) and the managers' division by psi.
template <class TRIAL>
__device__ __forceinline__ void trial_drift(double (&x)[TRIAL::NC], const TrialParamsDev &p, double &psi0,
                                            double (&d1)[TRIAL::NC], double (&d2)[TRIAL::NC])
{
    constexpr int NC = TRIAL::NC;
    if constexpr (TRIAL::ANALYTIC) {
        TRIAL::derivs(x, p, psi0, d1, d2);
    } else {
        psi0 = TRIAL::psi(x, p);
        double xx[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) xx[c] = x[c];
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
            // dynamic index into a small register array would spill: rebuild the stencil with selects
            double xm[NC], xp[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const double lo = xx[k] - p.fd_dx;
                const double hi = lo + 2.0 * p.fd_dx;
                const double back = hi - p.fd_dx;
                xm[k] = (k == c) ? lo : xx[k];
                xp[k] = (k == c) ? hi : xx[k];
                xx[k] = (k == c) ? back : xx[k];
            }
            const double pm = TRIAL::psi(xm, p), pp = TRIAL::psi(xp, p);
            const double first = (pp - pm) / (2.0 * p.fd_dx);
            const double sec = __dadd_rn(__dadd_rn(pm, -__dmul_rn(2.0, psi0)), pp) / p.fd_dx2;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                if (k == c) { d1[k] = first / psi0; d2[k] = sec / psi0; }
            }
        }
    }
}

// ImpSamp.local_kin (imp_samp.py:50-53): -0.5 * sum_d ( sum_a inv_m[a] * sec[a,d] ), NumPy's reduction order
// (ndim is a template parameter so that every array index is a compile-time constant: no local-memory copies)
template <int NC, int NDIM>
__device__ __forceinline__ double local_kinetic(const double (&d2)[NC], const double *inv_mass)
{
    constexpr int ndim = NDIM, natoms = NC / NDIM;
    double tot = 0.0;
#pragma unroll
    for (int d = 0; d < ndim; ++d) {
        double s = __dmul_rn(inv_mass[0], d2[d]);
#pragma unroll
        for (int a = 1; a < natoms; ++a) s = __dadd_rn(s, __dmul_rn(inv_mass[a], d2[a * ndim + d]));
        tot = (d == 0) ? s : __dadd_rn(tot, s);
    }
    return __dmul_rn(-0.5, tot);
}

// In-step variant of trial_drift + local_kinetic: same arithmetic and the same summation order as
// local_kinetic (s_d accumulates atom by atom, then s_0 + s_1 + s_2), but the second derivatives are folded
// into three running sums instead of an NC-long array and the stencil walks ONE working copy of the
// coordinates (x - dx, + 2 dx, - dx, as imp_samp.py:67-71 does), which keeps the kernel inside 128 registers.
template <class TRIAL>
__device__ __forceinline__ void trial_drift_ke(double (&xx)[TRIAL::NC], const TrialParamsDev &p, const double *inv_mass, double &psi0,
                                               double (&d1)[TRIAL::NC], double &ke)
{
    constexpr int NC = TRIAL::NC, ND = TRIAL::NDIM;
    if constexpr (TRIAL::ANALYTIC) {
        double d2[NC];
        TRIAL::derivs(xx, p, psi0, d1, d2);
        ke = local_kinetic<NC, ND>(d2, inv_mass);
    } else if constexpr (TRIAL::PARTIAL) {
        // atom by atom (compile time), dimension by dimension (run time): the selects that stand in for a dynamic register
        // index span the moved atom's three coordinates only, and the bond that atom is not part of is evaluated once per atom
        static_assert(NC == 9 && ND == 3, "partial stencil: three atoms in three dimensions");
        typename TRIAL::Bond keep = TRIAL::template bond<1>(xx, p);       // the H2-O bond, untouched while H1 is walked
        psi0 = TRIAL::template psi_moved<2>(xx, p, keep);
        double sd[ND] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int atom = 0; atom < 3; ++atom) {
            if (atom == 1) keep = TRIAL::template bond<0>(xx, p);          // H1 as the walk over its coordinates left it
#pragma unroll 1
            for (int d = 0; d < ND; ++d) {
                double orig = 0.0;
#pragma unroll
                for (int k = 0; k < ND; ++k) orig = (k == d) ? xx[3 * atom + k] : orig;
                const double lo = orig - p.fd_dx;
                const double hi = lo + 2.0 * p.fd_dx;
                double pm = 0.0, pp = 0.0;
#pragma unroll 1
                for (int side = 0; side < 2; ++side) {                      // one copy of the evaluation for both stencil points
                    const double at = side ? hi : lo;
#pragma unroll
                    for (int k = 0; k < ND; ++k) xx[3 * atom + k] = (k == d) ? at : xx[3 * atom + k];
                    double pv;
                    if (atom == 0) pv = TRIAL::template psi_moved<0>(xx, p, keep);
                    else if (atom == 1) pv = TRIAL::template psi_moved<1>(xx, p, keep);
                    else pv = TRIAL::template psi_moved<2>(xx, p, keep);
                    pm = side ? pm : pv;
                    pp = side ? pv : pp;
                }
#pragma unroll
                for (int k = 0; k < ND; ++k) xx[3 * atom + k] = (k == d) ? hi - p.fd_dx : xx[3 * atom + k];   // the walked value
                const double first = (pp - pm) / (2.0 * p.fd_dx);
                const double sec = __dadd_rn(__dadd_rn(pm, -__dmul_rn(2.0, psi0)), pp) / p.fd_dx2;
                const double term = __dmul_rn(inv_mass[atom], sec / psi0);
#pragma unroll
                for (int k = 0; k < ND; ++k) {
                    if (k == d) {
                        d1[3 * atom + k] = first / psi0;
                        sd[k] = (atom == 0) ? term : __dadd_rn(sd[k], term);
                    }
                }
            }
        }
        double tot = sd[0];
#pragma unroll
        for (int d = 1; d < ND; ++d) tot = __dadd_rn(tot, sd[d]);
        ke = __dmul_rn(-0.5, tot);
    } else {
        psi0 = TRIAL::psi(xx, p);
        double sd[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) sd[d] = 0.0;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
            double orig = 0.0;
#pragma unroll
            for (int k = 0; k < NC; ++k) orig = (k == c) ? xx[k] : orig;
            const double lo = orig - p.fd_dx;
            const double hi = lo + 2.0 * p.fd_dx;
#pragma unroll
            for (int k = 0; k < NC; ++k) xx[k] = (k == c) ? lo : xx[k];
            const double pm = TRIAL::psi(xx, p);
#pragma unroll
            for (int k = 0; k < NC; ++k) xx[k] = (k == c) ? hi : xx[k];
            const double pp = TRIAL::psi(xx, p);
#pragma unroll
            for (int k = 0; k < NC; ++k) xx[k] = (k == c) ? hi - p.fd_dx : xx[k];   // the walked value, exactly as trial_drift leaves it
            const double first = (pp - pm) / (2.0 * p.fd_dx);
            const double sec = __dadd_rn(__dadd_rn(pm, -__dmul_rn(2.0, psi0)), pp) / p.fd_dx2;
            const double term = __dmul_rn(inv_mass[c / ND], sec / psi0);
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (k == c) d1[k] = first / psi0;
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (d == c % ND) sd[d] = (c < ND) ? term : __dadd_rn(sd[d], term);
        }
        double tot = sd[0];
#pragma unroll
        for (int d = 1; d < ND; ++d) tot = __dadd_rn(tot, sd[d]);
        ke = __dmul_rn(-0.5, tot);
    }
}

// ImpSamp.metropolis (imp_samp.py:29-47) for one walker
// dx_ / dy_ are the drift vectors D = (1/m) f (capped for excited-state importance sampling) at x and y
template <int NC, int NDIM>
__device__ __forceinline__ double metropolis_ratio(const double (&x)[NC], const double (&y)[NC], const double (&dx_)[NC],
                                                   const double (&dy_)[NC], double psi_x, double psi_y, const double *sigma, double dt)
{
    constexpr int ndim = NDIM, natoms = NC / NDIM;
    const double q = psi_y / psi_x;
    const double ratio = __dmul_rn(q, q);
    // prod_{a,d} exp(-u1^2 / 2 s^2) / exp(-u2^2 / 2 s^2) = exp( sum (u2^2 - u1^2) / 2 s^2 ): one exponential instead of
    // 2 * NC (and no intermediate underflow); agrees with the reference's product of ratios to a few ulp
    double expo = 0.0;
#pragma unroll
    for (int d = 0; d < ndim; ++d) {
#pragma unroll
        for (int a = 0; a < natoms; ++a) {
            const int c = a * ndim + d;
            const double inv_two_s2 = 0.5 / __dmul_rn(sigma[a], sigma[a]);
            const double dxm = __dmul_rn(dx_[c], dt), dym = __dmul_rn(dy_[c], dt);
            const double u1 = __dadd_rn(__dadd_rn(x[c], -y[c]), -dym);
            const double u2 = __dadd_rn(__dadd_rn(y[c], -x[c]), -dxm);
            expo = fma((u2 - u1) * (u2 + u1), inv_two_s2, expo);
        }
    }
    double acc = __dmul_rn(exp(expo), ratio);
    if (__dmul_rn(psi_x, psi_y) <= 0.0) acc = 0.0;
    return acc;
}

// excited_state_imp_samp (pyvibdmc.py:562-591): per atom the drift vector is shortened by
// factor = (-1 + sqrt(1 + 2 m v^2)) / (m v^2), v^2 = |D + 1e-50|^2; the vector score is sqrt(sum_a m_a |D'_a|^2 / sum_a m_a |D_a|^2).
// shifted: the reference scales the shifted copy for D_x (:568) but the unshifted vector for D_y (:586).
template <int NC>
__device__ __forceinline__ double cap_drift(const double (&d)[NC], const double *inv_mass, const double *mass, bool shifted,
                                            double (&dc)[NC])
{
    static_assert(NC % 3 == 0, "excited-state drift capping is written for 3-D atoms");
    double numer = 0.0, denom = 0.0;
#pragma unroll
    for (int a = 0; a < NC / 3; ++a) {
        const double ms = 1.0 / inv_mass[a];
        const double sx = d[3 * a] + 1e-50, sy = d[3 * a + 1] + 1e-50, sz = d[3 * a + 2] + 1e-50;
        const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy)), __dmul_rn(sz, sz)));
        const double v2 = __dmul_rn(nrm, nrm);
        const double msv2 = __dmul_rn(ms, v2);
        const double factor = (-1.0 + sqrt(__dadd_rn(1.0, __dmul_rn(__dmul_rn(2.0, ms), v2)))) / msv2;
        dc[3 * a] = __dmul_rn(factor, shifted ? sx : d[3 * a]);
        dc[3 * a + 1] = __dmul_rn(factor, shifted ? sy : d[3 * a + 1]);
        dc[3 * a + 2] = __dmul_rn(factor, shifted ? sz : d[3 * a + 2]);
        const double n2 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dc[3 * a], dc[3 * a]), __dmul_rn(dc[3 * a + 1], dc[3 * a + 1])),
                                         __dmul_rn(dc[3 * a + 2], dc[3 * a + 2])));
        const double n1 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d[3 * a], d[3 * a]), __dmul_rn(d[3 * a + 1], d[3 * a + 1])),
                                         __dmul_rn(d[3 * a + 2], d[3 * a + 2])));
        const double tn = __dmul_rn(__dmul_rn(n2, n2), mass[a]), td = __dmul_rn(__dmul_rn(n1, n1), mass[a]);
        numer = (a == 0) ? tn : __dadd_rn(numer, tn);
        denom = (a == 0) ? td : __dadd_rn(denom, td);
    }
    return sqrt(numer / denom);
}

struct ImpArgs {
    TrialParamsDev trial;
    double inv_mass[PVD_MAX_ATOMS];
    double mass[PVD_MAX_ATOMS];
    double *vscore;            // excited-state importance sampling: per-walker vector score (in place), else nullptr
    const double *inj_um;      // injected Metropolis uniforms or nullptr
    unsigned long long *acc_count;   // accepted moves of this shard in the current step
};

// first-step exception with importance sampling (pyvibdmc.py:553-554, 760-769): drift on the start
// ensemble, E_L = V + local kinetic energy.
template <class TRIAL, class POT>
__global__ void __launch_bounds__(PVD_CTA, 2) k_imp_init(const StepArgs a, const ImpArgs im, double *x, double *f, double *psi, double *lk, double *v,
                                                         long long first, long long count)
{
    constexpr int NC = TRIAL::NC;
    // count < 0: the whole shard; else the walkers [first, first + count) (walkers received from another shard)
    const long long n = count < 0 ? a.st[a.parity].n : first + count;
    for (long long i = (count < 0 ? 0 : first) + blockIdx.x * (long long)PVD_CTA + threadIdx.x; i < n; i += (long long)gridDim.x * PVD_CTA) {
        double xx[NC], d1[NC], d2[NC], p0;
#pragma unroll
        for (int c = 0; c < NC; ++c) xx[c] = x[c * a.cap + i];
        trial_drift<TRIAL>(xx, im.trial, p0, d1, d2);
        const double ke = local_kinetic<NC, TRIAL::NDIM>(d2, im.inv_mass);
        if constexpr (NC % 3 == 0) {
            if (im.vscore) {                       // vector score of the start ensemble (pyvibdmc.py:570-573)
                double dr[NC], dcap[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) dr[c] = __dmul_rn(im.inv_mass[c / 3], d1[c]);
                im.vscore[i] = cap_drift<NC>(dr, im.inv_mass, im.mass, true, dcap);
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) f[c * a.cap + i] = d1[c];
        psi[i] = p0;
        lk[i] = ke;
        v[i] = __dadd_rn(POT::eval(xx, a.pot), ke);
    }
}

// imp_move_randomly (pyvibdmc.py:549-612) + potential + local energy (:786-812), in place.
// Writes the accepted/kept walker, its drift, psi, local kinetic energy and E_L; counts acceptances.
// The last CTA publishes dt_eff = dt * n_accept / N (pyvibdmc.py:603, 372-378) for the branching kernel.
template <class TRIAL, class POT, int RNG, int VAR>
__global__ void __launch_bounds__(PVD_CTA, 2) k_imp_move(const StepArgs a, const ImpArgs im, double *x, double *f, double *psi, double *lk, double *v)
{
    constexpr int NC = TRIAL::NC;
    constexpr bool SECOND = VAR == PVD_IMP_SECOND_DISPLACEMENT, EXCITED = VAR == PVD_IMP_EXCITED_STATE && NC % 3 == 0;
    __shared__ unsigned s_cnt[PVD_WARPS];
    __shared__ unsigned s_last;
    DevState *sip = &a.st[a.parity];
#if PVD_FD_NOINLINE
    // the out-of-line trial function takes the table description by pointer: a copy in shared memory (a pointer into the kernel's
    // parameter block would make the compiler copy the block onto every thread's stack)
    __shared__ TrialParamsDev s_trial;
    if (threadIdx.x == 0) s_trial = im.trial;
    __syncthreads();
#define PVD_IMP_TRIAL s_trial
#else
#define PVD_IMP_TRIAL im.trial
#endif
    if (sip->err || sip->n <= 0) return;           // the branching kernel forwards the dead state
    const double vref_now = sip->vref;
    const long long n = sip->n, step = sip->step;
    unsigned my_acc = 0;
    // the walker's current position and drift, and the proposed position, wait in shared memory (thread-private
    // columns) while the finite-difference stencil is evaluated on a working copy: they would otherwise pin 54
    // registers through the heaviest code
    extern __shared__ __align__(16) unsigned char s_imp_raw[];
    double (*s_xf)[PVD_CTA] = reinterpret_cast<double (*)[PVD_CTA]>(s_imp_raw);          // [3 * NC][PVD_CTA]
    const int tid = threadIdx.x;
    for (long long i = blockIdx.x * (long long)PVD_CTA + threadIdx.x; i < n; i += (long long)gridDim.x * PVD_CTA) {
        double xx[NC];
        double psi_x, ke_x = 0.0;
        // the Gaussian displacement (injected or Philox + Box-Muller), scaled by sigma
        if (a.inj_disp) {
#pragma unroll
            for (int c = 0; c < NC; ++c) xx[c] = a.inj_disp[c * a.cap + i];
        } else {
            walker_normals<NC, RNG>(a.seed, i, step, xx);
#pragma unroll
            for (int c = 0; c < NC; ++c) xx[c] = __dmul_rn(a.sigc[c], xx[c]);
        }
        if constexpr (!SECOND) {
            // imp_move_randomly (:549-612): displaced = coords + disps + (inv_m * f_x) * dt, f_x and psi carried from the last step
            double xo[NC], fo[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) { xo[c] = x[c * a.cap + i]; fo[c] = f[c * a.cap + i]; }
            // slot 1 of the stash holds the drift vector D_x = (1/m) f_x (capped in the excited-state variant)
#pragma unroll
            for (int c = 0; c < NC; ++c) fo[c] = __dmul_rn(im.inv_mass[c / TRIAL::NDIM], fo[c]);
            if constexpr (EXCITED) {
                double dcap[NC];
                cap_drift<NC>(fo, im.inv_mass, im.mass, true, dcap);
#pragma unroll
                for (int c = 0; c < NC; ++c) fo[c] = dcap[c];
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                xx[c] = __dadd_rn(__dadd_rn(xo[c], xx[c]), __dmul_rn(fo[c], a.dt));
                s_xf[c][tid] = xo[c];
                s_xf[NC + c][tid] = fo[c];
                s_xf[2 * NC + c][tid] = xx[c];
            }
            psi_x = psi[i];
        } else {
            // imp_move_randomly_second_type (:614-649): the diffusion step is always taken, the drift is evaluated at the
            // diffused position and only the drift step displaced = x' + (inv_m * f_x') * dt goes through Metropolis
#pragma unroll
            for (int c = 0; c < NC; ++c) { xx[c] = __dadd_rn(x[c * a.cap + i], xx[c]); s_xf[c][tid] = xx[c]; }
            double f1[NC];
            trial_drift_ke<TRIAL>(xx, PVD_IMP_TRIAL, im.inv_mass, psi_x, f1, ke_x);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                xx[c] = __dadd_rn(s_xf[c][tid], __dmul_rn(__dmul_rn(im.inv_mass[c / TRIAL::NDIM], f1[c]), a.dt));
                s_xf[NC + c][tid] = f1[c];                                  // the drift itself: rejected walkers keep it
                s_xf[2 * NC + c][tid] = xx[c];
            }
        }
        double fy[NC], psi_y, ke_new;
        trial_drift_ke<TRIAL>(xx, PVD_IMP_TRIAL, im.inv_mass, psi_y, fy, ke_new);
        double xo[NC], y[NC];
        double acc, vs_new = 1.0;
        {
            double fo[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) { xo[c] = s_xf[c][tid]; fo[c] = s_xf[NC + c][tid]; y[c] = s_xf[2 * NC + c][tid]; }
            double dy[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) dy[c] = __dmul_rn(im.inv_mass[c / TRIAL::NDIM], fy[c]);
            if constexpr (SECOND) {
                double dxv[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) dxv[c] = __dmul_rn(im.inv_mass[c / TRIAL::NDIM], fo[c]);
                acc = metropolis_ratio<NC, TRIAL::NDIM>(xo, y, dxv, dy, psi_x, psi_y, a.sigma, a.dt);
            } else if constexpr (EXCITED) {
                double dcap[NC];
                vs_new = cap_drift<NC>(dy, im.inv_mass, im.mass, false, dcap);
                acc = metropolis_ratio<NC, TRIAL::NDIM>(xo, y, fo, dcap, psi_x, psi_y, a.sigma, a.dt);
            } else acc = metropolis_ratio<NC, TRIAL::NDIM>(xo, y, fo, dy, psi_x, psi_y, a.sigma, a.dt);
            if constexpr (SECOND) {
                // rejected walkers keep x' with ITS drift / psi / kinetic energy: store them now, overwritten below on acceptance
#pragma unroll
                for (int c = 0; c < NC; ++c) { x[c * a.cap + i] = xo[c]; f[c * a.cap + i] = fo[c]; }
                psi[i] = psi_x;
                lk[i] = ke_x;
            }
        }
        double u;
        if (im.inj_um) u = im.inj_um[i];
        else { const uint4 r = pvd_draw(a.seed, i, step, PVD_STREAM_METRO, 0u); u = u53(r.x, r.y); }
        const bool ok = acc > u;
        double ke = SECOND ? ke_x : lk[i];
        if (ok) {
            ke = ke_new;
#pragma unroll
            for (int c = 0; c < NC; ++c) { xo[c] = y[c]; x[c * a.cap + i] = y[c]; f[c * a.cap + i] = fy[c]; }
            psi[i] = psi_y;
            lk[i] = ke;
            ++my_acc;
        }
        double el = __dadd_rn(POT::eval(xo, a.pot), ke);
        if constexpr (EXCITED) {
            const double vs = ok ? vs_new : im.vscore[i];
            if (ok) im.vscore[i] = vs;
            el = vref_now - (vref_now - el) * vs;                           // pyvibdmc.py:810-811
        }
        v[i] = el;
    }
    // acceptance count -> dt_eff
    for (int off = 16; off > 0; off >>= 1) my_acc += __shfl_xor_sync(0xffffffffu, my_acc, off);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = my_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < PVD_WARPS; ++w) tot += s_cnt[w];
        atomicAdd(im.acc_count, (unsigned long long)tot);
        __threadfence();
        const unsigned d = atomicAdd(&sip->done, 1u);
        s_last = (d == gridDim.x - 1) ? 1u : 0u;
        if (s_last) {
            __threadfence();
            const unsigned long long nacc = atomicAdd(im.acc_count, 0ull);
            sip->n_accept = (long long)nacc;
            if (a.world == 1) sip->dt_eff = __dmul_rn(a.dt, (double)nacc / (double)n);
            else { a.sums[0] = (double)nacc; a.sums[1] = (double)n; }          // reduced over the shards, then k_imp_set_dt
            sip->done = 0u;
            *im.acc_count = 0ull;
        }
    }
}

// multi-GPU: dt_eff from the globally reduced acceptance count (sums[0] = n_accept, sums[1] = n)
__global__ void k_imp_set_dt(DevState *st, int parity, const double *sums, double dt)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st[parity].dt_eff = __dmul_rn(dt, sums[0] / sums[1]);
        st[parity].n_accept = (long long)sums[0];
    }
}

// ---------------------------------------------------------------- stand-alone entry points (AoS host layout)
template <class TRIAL>
__global__ void k_trial_drift_aos(const double *__restrict__ xyz, long long n, const TrialParamsDev p, double *psi, double *dlog, double *d2)
{
    constexpr int NC = TRIAL::NC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x[NC], a1[NC], a2[NC], p0;
#pragma unroll
        for (int c = 0; c < NC; ++c) x[c] = xyz[i * NC + c];
        trial_drift<TRIAL>(x, p, p0, a1, a2);
        psi[i] = TRIAL::psi(x, p);
#pragma unroll
        for (int c = 0; c < NC; ++c) { dlog[i * NC + c] = a1[c]; d2[i * NC + c] = a2[c]; }
    }
}

template <int NC>
__global__ void k_metropolis_aos(const double *x, const double *y, const double *fx, const double *fy, const double *psx,
                                 const double *psy, long long n, int ndim, const double *sigma, const double *inv_mass, double dt, double *acc)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double a[NC], b[NC], c[NC], d[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) { a[k] = x[i * NC + k]; b[k] = y[i * NC + k]; c[k] = fx[i * NC + k]; d[k] = fy[i * NC + k]; }
        (void)ndim;
#pragma unroll
        for (int k = 0; k < NC; ++k) { c[k] = __dmul_rn(inv_mass[k / (NC == 9 ? 3 : 1)], c[k]); d[k] = __dmul_rn(inv_mass[k / (NC == 9 ? 3 : 1)], d[k]); }
        acc[i] = metropolis_ratio<NC, (NC == 9 ? 3 : 1)>(a, b, c, d, psx[i], psy[i], sigma, dt);
    }
}

template <int NC>
__global__ void k_local_kin_aos(const double *d2, long long n, int ndim, const double *inv_mass, double *ke)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double a[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) a[k] = d2[i * NC + k];
        (void)ndim;
        ke[i] = local_kinetic<NC, (NC == 9 ? 3 : 1)>(a, inv_mass);
    }
}

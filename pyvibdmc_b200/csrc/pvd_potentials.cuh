// fp64 per-walker potential energy surfaces (device functions) + their host-side set-up.
#pragma once
#include "pvd_common.cuh"
#include "ps_h2o_table.h"

// ---------------------------------------------------------------- Partridge-Schwenke H2O
// Replaces calc_hoh_pot (calc_h2o_pot.f:1-34) + vibpot (h2opes_v2.f:460-489).
// Folded coefficients live in constant memory: every lane of a warp reads the same coefficient at
// the same time, so each one is a constant-bank operand of its DFMA (no load instruction).
struct PsScalars {
    double reoh, inv_reoh, b1, ce, phh1, phh2, deoh, roh, alphaoh, c0;
};
__constant__ double c_ps_horner[PS_NTERMS];
__constant__ PsScalars c_ps;

// Host: the one-time folding of h2opes_v2.f:421-459 in the Fortran's operation order
// (this translation unit is compiled with -fmad=false on the host side by construction:
// host code never contracts; see build flags).  Returns folded c[245] and the 8 scalars.
static inline void ps_fold_host(double *c, double *scal8)
{
    double reoh = PS_RAW_REOH, thetae = PS_RAW_THETAE, b1 = PS_RAW_B1, roh = PS_RAW_ROH;
    double alphaoh = PS_RAW_ALPHAOH, deoh = PS_RAW_DEOH, phh1 = PS_RAW_PHH1, phh2 = PS_RAW_PHH2;
    for (int i = 0; i < PS_NTERMS; ++i) {
        const PsTermRaw &t = PS_RAW_TERMS[i];
        volatile double a = PS_RAW_F5Z * t.c5z;
        volatile double b = PS_RAW_FBASIS * t.cbasis;
        volatile double d = PS_RAW_FCORE * t.ccore;
        volatile double e = PS_RAW_FREST * t.crest;
        volatile double s = a + b;
        s = s + d;
        s = s + e;
        c[i] = s;
    }
    phh1 = phh1 * PS_RAW_F5Z;
    deoh = deoh * PS_RAW_F5Z;
    reoh = reoh / 0.529177249;
    b1 = b1 * 0.529177249 * 0.529177249;
    for (int i = 0; i < PS_NTERMS; ++i) c[i] = c[i] * 4.556335e-6;
    const double rad = acos(-1.0) / 1.8e2;
    const double ce = cos(thetae * rad);
    phh1 = phh1 * exp(phh2);
    phh1 = phh1 * 4.556335e-6;
    phh2 = phh2 * 0.529177249;
    deoh = deoh * 4.556335e-6;
    roh = roh / 0.529177249;
    alphaoh = alphaoh * 0.529177249;
    c[0] = c[0] * 2.0;
    scal8[0] = reoh; scal8[1] = b1; scal8[2] = ce; scal8[3] = phh1;
    scal8[4] = phh2; scal8[5] = deoh; scal8[6] = roh; scal8[7] = alphaoh;
}

static inline cudaError_t ps_upload_constants()
{
    double c[PS_NTERMS], s[8], h[PS_NTERMS];
    ps_fold_host(c, s);
    for (int k = 0; k < PS_NTERMS; ++k) h[k] = c[PS_HORNER_TERM[k]];
    h[0] = 0.0;   // term 0 is the constant added outside the damped sum (h2opes_v2.f:480,487)
    PsScalars p{s[0], 1.0 / s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], c[0]};
    cudaError_t e = cudaMemcpyToSymbol(c_ps_horner, h, sizeof(h));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_ps, &p, sizeof(p));
}

// One (a,b) group: Horner in x3 times the symmetrised x1/x2 monomial.  The group structure is a
// template parameter so every exponent and coefficient slot is a compile-time constant.
// Groups are evaluated PS_BATCH at a time with their Horner recurrences advanced in lock-step, so that
// PS_BATCH independent DFMA chains are in flight (a single chain is latency-bound: one DFMA per ~8 cycles).
constexpr int PS_BATCH = 5;
static_assert(PS_NGROUPS % PS_BATCH == 0, "25 groups = 5 batches of 5");

// one Horner step of group Q at power L (no-op when the group has no x3^L coefficient)
template <int Q, int L>
__device__ __forceinline__ void ps_step(double &h, double x3)
{
    constexpr int len = PS_GLEN[Q];
    constexpr int idx = PS_GSTART[Q] + (L < len ? L : 0);
    if constexpr (len - 1 == L) h = c_ps_horner[idx];
    else if constexpr (len - 1 > L) h = fma(h, x3, c_ps_horner[idx]);
}

template <int G0, int L>
__device__ __forceinline__ void ps_batch_level(double x3, double (&h)[PS_BATCH])
{
    if constexpr (L >= 0) {
        ps_step<G0 + 0, L>(h[0], x3);
        ps_step<G0 + 1, L>(h[1], x3);
        ps_step<G0 + 2, L>(h[2], x3);
        ps_step<G0 + 3, L>(h[3], x3);
        ps_step<G0 + 4, L>(h[4], x3);
        ps_batch_level<G0, L - 1>(x3, h);
    }
}

template <int Q>
__device__ __forceinline__ double ps_sym(const double (&p1)[9], const double (&p2)[9])
{
    constexpr int a = PS_GA[Q], b = PS_GB[Q];
    if constexpr (a == b) return 2.0 * (p1[a] * p2[a]);
    else return fma(p1[a], p2[b], p1[b] * p2[a]);
}

template <int G>
__device__ __forceinline__ void ps_group(double x3, const double (&p1)[9], const double (&p2)[9], double &acc)
{
    double h[PS_BATCH];
    ps_batch_level<G, PS_GLEN[G] - 1>(x3, h);          // groups are ordered by non-increasing length
    const double t0 = ps_sym<G + 0>(p1, p2) * h[0], t1 = ps_sym<G + 1>(p1, p2) * h[1];
    const double t2 = fma(ps_sym<G + 2>(p1, p2), h[2], t0), t3 = fma(ps_sym<G + 3>(p1, p2), h[3], t1);
    acc += fma(ps_sym<G + 4>(p1, p2), h[4], t2) + t3;
    if constexpr (G + PS_BATCH < PS_NGROUPS) ps_group<G + PS_BATCH>(x3, p1, p2, acc);
}

// 1/sqrt(t) for the squared distances of the PES (t of order 1..100 bohr^2; callers clamp t >= 1e-30 so that two
// coinciding atoms give a huge finite 1/r instead of NaN)
__device__ __forceinline__ double ps_rsqrt(double t)
{
    double y = (double)rsqrtf((float)t);
    const double h = 0.5 * t;
    y = y * fma(-h, y * y, 1.5);
    y = y * fma(-h, y * y, 1.5);
    return y;
}
__device__ __forceinline__ double ps_sqrt_from(double t, double y)
{
    const double r = t * y;
    return fma(0.5 * y, fma(-r, r, t), r);
}

// x: 9 Cartesians, atoms ordered H, H, O (calc_h2o_pot.f:18-19).
// The 244 polynomial terms c_j (x1^a x2^b + x1^b x2^a) x3^l are evaluated as 25 (a,b)-groups,
// each a Horner polynomial in x3 (SURVEY hard part 6): ~330 DFMA instead of ~1460 flops.
// cos(theta) is used directly (the reference goes through acos and back: <= 1 ulp apart).
__device__ __forceinline__ double ps_h2o_energy(const double (&x)[9])
{
    const double d1x = x[6] - x[0], d1y = x[7] - x[1], d1z = x[8] - x[2];
    const double d2x = x[6] - x[3], d2y = x[7] - x[4], d2z = x[8] - x[5];
    const double r1s = fmax(d1x * d1x + d1y * d1y + d1z * d1z, 1e-30);
    const double r2s = fmax(d2x * d2x + d2y * d2y + d2z * d2z, 1e-30);
    const double ct = d1x * d2x + d1y * d2y + d1z * d2z;
    // 1/r and r from one float-seeded Newton iteration chain each (two steps: 22 -> 44 -> 88 bits, then one residual
    // correction of r): about a third of the instructions of sqrt() + a division, same result to 1-2 ulp
    const double y1 = ps_rsqrt(r1s), y2 = ps_rsqrt(r2s);
    const double r1 = ps_sqrt_from(r1s, y1), r2 = ps_sqrt_from(r2s, y2);
    const double costh = ct * (y1 * y2);
    const double a1 = r1 - c_ps.reoh, a2 = r2 - c_ps.reoh;
    const double x1 = a1 * c_ps.inv_reoh, x2 = a2 * c_ps.inv_reoh;
    const double x3 = costh - c_ps.ce;
    const double rhh2 = fmax(r1s + r2s - 2.0 * ct, 1e-30);
    const double rhh = ps_sqrt_from(rhh2, ps_rsqrt(rhh2));
    const double vhh = c_ps.phh1 * exp(-c_ps.phh2 * rhh);
    const double e1 = exp(-c_ps.alphaoh * (r1 - c_ps.roh));
    const double e2 = exp(-c_ps.alphaoh * (r2 - c_ps.roh));
    const double voh = c_ps.deoh * (e1 * (e1 - 2.0) + e2 * (e2 - 2.0));
    const double damp = exp(-c_ps.b1 * (a1 * a1 + a2 * a2));

    double p1[9], p2[9];
    p1[0] = 1.0; p2[0] = 1.0;
#pragma unroll
    for (int k = 1; k < 9; ++k) { p1[k] = p1[k - 1] * x1; p2[k] = p2[k - 1] * x2; }

    double acc = 0.0;
    ps_group<0>(x3, p1, p2, acc);
    return fma(acc, damp, c_ps.c0) + voh + vhh;
}

// ---------------------------------------------------------------- potential policies for the step kernels
struct PotParamsDev {
    double k[PVD_MAX_COMP];   // HARMONIC: k[c] ; MORSE1D: k[0]=De, k[1]=alpha
};

struct PotH2O {
    static constexpr int NC = 9;
    static constexpr int MIN_CTAS = 2;       // resident CTAs per SM the fused step kernel is compiled for (register budget)
    __device__ static __forceinline__ double eval(const double (&x)[9], const PotParamsDev &) { return ps_h2o_energy(x); }
};
// harmonicOscillator1D.py:13-17 : ((0.5*m)*w^2) * (x*x), summed over components
template <int NCOMP>
struct PotHarm {
    static constexpr int NC = NCOMP;
    static constexpr int MIN_CTAS = 4;       // cheap potential: the step is latency-bound, so run more warps per SM
    __device__ static __forceinline__ double eval(const double (&x)[NCOMP], const PotParamsDev &p)
    {
        double v = __dmul_rn(p.k[0], __dmul_rn(x[0], x[0]));      // no FMA contraction: bit-exact vs NumPy
#pragma unroll
        for (int c = 1; c < NCOMP; ++c) v = __dadd_rn(v, __dmul_rn(p.k[c], __dmul_rn(x[c], x[c])));
        return v;
    }
};
// morse_osc_1d.py:4-12
struct PotMorse {
    static constexpr int NC = 1;
    static constexpr int MIN_CTAS = 4;
    __device__ static __forceinline__ double eval(const double (&x)[1], const PotParamsDev &p)
    {
        const double t = 1.0 - exp(-p.k[1] * x[0]);
        return p.k[0] * (t * t);
    }
};

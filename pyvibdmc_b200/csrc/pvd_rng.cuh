// Counter-based random numbers: Philox4x32-10 (Salmon et al., SC'11) + normals by ziggurat (default) or Box-Muller.
// Replaces the NumPy global MT19937 stream used by DMC_Sim.move_randomly / birth_or_death
// (pyvibdmc.py:392,544-546,601).  Streams are addressed by (seed; walker slot, step, purpose),
// so results do not depend on the launch configuration.
#pragma once
#include "pvd_common.cuh"

enum : unsigned { PVD_STREAM_DISP = 0u, PVD_STREAM_BRANCH = 1u, PVD_STREAM_METRO = 2u };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// counter layout: x = walker slot (low 32), y = walker slot (high) | call index << 24,
//                 z = step (low 32),        w = step (high 24 bits) | purpose << 24
__device__ __forceinline__ uint4 pvd_draw(uint64_t seed, long long slot, long long step, unsigned purpose, unsigned call)
{
    const uint4 ctr = make_uint4((unsigned)slot, (unsigned)((unsigned long long)slot >> 32) | (call << 24),
                                 (unsigned)step, ((unsigned)((unsigned long long)step >> 32) & 0xFFFFFFu) | (purpose << 24));
    return philox4x32_10(ctr, make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
}

// uniform in [0,1) with 53 random bits -- same range as np.random.random
__device__ __forceinline__ double u53(unsigned lo, unsigned hi)
{
    const unsigned long long r = ((unsigned long long)hi << 32) | lo;
    return (double)(r >> 11) * 0x1.0p-53;
}

// One Philox call -> two standard normals.
//   PVD_RNG_FP64: Box-Muller evaluated entirely in double (log, sqrt, sincospi).
//   PVD_RNG_FAST: radius and angle from SFU (MUFU lg2/sqrt/sin/cos) float32 intrinsics with a 64-bit
//                 exponent-extended uniform, result widened to double; relative error ~1e-6 per
//                 normal, tails to 9.4 sigma.
template <int MODE>
__device__ __forceinline__ void normal_pair(uint4 r, double &z0, double &z1)
{
    if (MODE == PVD_RNG_FP64) {
        // Box-Muller in double, with the transcendental functions specialised for random-bit inputs
        // (no generic range reduction, no special cases), ~2x fewer instructions than log/sqrt/sincospi:
        //   u = (1+f) 2^-(e+1): e = leading zeros of y (geometric), f = 52 bits from x and w[19:0]
        //   angle = 2 pi N / 2^44, N from z and w[31:20]
        const int e = r.y ? __clz((int)r.y) : 32;
        const unsigned long long mant = ((unsigned long long)r.x << 20) | (unsigned long long)(r.w & 0xFFFFFu);
        double m = __longlong_as_double((long long)(0x3FF0000000000000ull | mant));        // [1,2)
        int k = e + 1;
        if (m > 1.4142135623730951) { m *= 0.5; --k; }                                   // m in [0.7071, 1.4142]
        // ln m = 2 atanh(s), s = (m-1)/(m+1); reciprocal by float seed + two Newton steps (rel. error < 2^-80)
        const double den = m + 1.0;
        double ri = (double)__frcp_rn((float)den);
        ri = fma(ri, fma(-den, ri, 1.0), ri);
        ri = fma(ri, fma(-den, ri, 1.0), ri);
        const double s = (m - 1.0) * ri, s2 = s * s;
        double p = 1.0 / 21.0;
        p = fma(p, s2, 1.0 / 19.0); p = fma(p, s2, 1.0 / 17.0); p = fma(p, s2, 1.0 / 15.0); p = fma(p, s2, 1.0 / 13.0); p = fma(p, s2, 1.0 / 11.0);
        p = fma(p, s2, 1.0 / 9.0); p = fma(p, s2, 1.0 / 7.0); p = fma(p, s2, 1.0 / 5.0); p = fma(p, s2, 1.0 / 3.0);
        p = fma(p * s2, s, s);                                                           // atanh(s)
        const double t = fma((double)k, 1.3862943611198906, -4.0 * p);                   // -2 ln u = 2k ln2 - 4 atanh(s)  (> 0)
        // sqrt(t) = t * rsqrt(t): float seed + two Newton steps
        double y = (double)rsqrtf((float)t);
        y = y * fma(-0.5 * t, y * y, 1.5);
        y = y * fma(-0.5 * t, y * y, 1.5);
        const double rad = t * y;
        // angle: octant o (3 bits) + 41-bit fraction; a in [-pi/4, pi/4], then a quadrant rotation
        const unsigned long long ang = ((unsigned long long)r.z << 12) | (unsigned long long)(r.w >> 20);   // 44 bits
        const int o = (int)(ang >> 41);
        long long frac = (long long)(ang & 0x1FFFFFFFFFFull);
        if (o & 1) frac -= (1ll << 41);
        const double a = (double)frac * (0.78539816339744831 * 0x1.0p-41);
        const double a2 = a * a;
        double sn = -1.0 / 355687428096000.0;                                             // -1/17!
        sn = fma(sn, a2, 1.0 / 1307674368000.0); sn = fma(sn, a2, -1.0 / 6227020800.0); sn = fma(sn, a2, 1.0 / 39916800.0);
        sn = fma(sn, a2, -1.0 / 362880.0); sn = fma(sn, a2, 1.0 / 5040.0); sn = fma(sn, a2, -1.0 / 120.0);
        sn = fma(sn, a2, 1.0 / 6.0);
        sn = fma(-sn * a2, a, a);                                                         // sin a
        double cs = 1.0 / 20922789888000.0;                                               // 1/16!
        cs = fma(cs, a2, -1.0 / 87178291200.0); cs = fma(cs, a2, 1.0 / 479001600.0); cs = fma(cs, a2, -1.0 / 3628800.0);
        cs = fma(cs, a2, 1.0 / 40320.0); cs = fma(cs, a2, -1.0 / 720.0); cs = fma(cs, a2, 1.0 / 24.0);
        cs = fma(cs, a2, -0.5);
        cs = fma(cs, a2, 1.0);                                                            // cos a
        const int j = ((o + 1) >> 1) & 3;                                                 // nearest multiple of pi/2
        const double c0 = (j & 1) ? sn : cs, s0 = (j & 1) ? cs : sn;
        const double cosv = (j == 1 || j == 2) ? -c0 : c0;
        const double sinv = (j == 2 || j == 3) ? -s0 : s0;
        z0 = rad * cosv;
        z1 = rad * sinv;
    } else {
        // -ln(u), u = m * 2^-(e+1), m in [1,2): e = leading zeros of a 64-bit word, m from the next 24 bits
        unsigned long long w = ((unsigned long long)r.y << 32) | r.x;
        const int e = w ? __clzll((long long)w) : 63;
        w = (w << e) << 1;                                   // drop the leading one
        const float m = __uint_as_float(0x3F800000u | (unsigned)(w >> 41));
        const float nlog2 = (float)(e + 1) - __log2f(m);     // -log2(u) > 0
        const float rad = sqrtf(1.3862943611198906f * nlog2);
        const float ang = (float)(int)r.z * 1.4629180792671596e-9f;   // 2*pi*2^-32 * signed word: [-pi, pi)
        float s, c;
        __sincosf(ang, &s, &c);
        z0 = (double)(rad * c);
        z1 = (double)(rad * s);
        (void)r.w;
    }
}

// NP Box-Muller pairs at once, fp64, stage by stage across the pairs: every polynomial coefficient is
// materialised once and used by NP back-to-back independent DFMAs (fewer constant moves, NP-way ILP).
// Same arithmetic per pair as normal_pair<PVD_RNG_FP64>.
template <int NP>
__device__ __forceinline__ void normal_pairs_fp64(const uint4 (&r)[NP], double (&z0)[NP], double (&z1)[NP])
{
    double m[NP], kk[NP], s[NP], s2[NP], p[NP], a[NP], a2[NP], sn[NP], cs[NP], rad[NP];
    int oct[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const int e = r[q].y ? __clz((int)r[q].y) : 32;
        const unsigned long long mant = ((unsigned long long)r[q].x << 20) | (unsigned long long)(r[q].w & 0xFFFFFu);
        double mm = __longlong_as_double((long long)(0x3FF0000000000000ull | mant));
        int k = e + 1;
        if (mm > 1.4142135623730951) { mm *= 0.5; --k; }
        m[q] = mm;
        kk[q] = (double)k;
        const unsigned long long ang = ((unsigned long long)r[q].z << 12) | (unsigned long long)(r[q].w >> 20);
        const int o = (int)(ang >> 41);
        long long frac = (long long)(ang & 0x1FFFFFFFFFFull);
        if (o & 1) frac -= (1ll << 41);
        oct[q] = o;
        a[q] = (double)frac * (0.78539816339744831 * 0x1.0p-41);
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const double den = m[q] + 1.0;
        double ri = (double)__frcp_rn((float)den);
        ri = fma(ri, fma(-den, ri, 1.0), ri);
        ri = fma(ri, fma(-den, ri, 1.0), ri);
        s[q] = (m[q] - 1.0) * ri;
        s2[q] = s[q] * s[q];
        a2[q] = a[q] * a[q];
        p[q] = 1.0 / 21.0;
        sn[q] = -1.0 / 355687428096000.0;
        cs[q] = 1.0 / 20922789888000.0;
    }
#define PVD_STAGE(ARR, X, C)                                         \
    _Pragma("unroll") for (int q = 0; q < NP; ++q) ARR[q] = fma(ARR[q], X[q], C);
    PVD_STAGE(p, s2, 1.0 / 19.0) PVD_STAGE(p, s2, 1.0 / 17.0) PVD_STAGE(p, s2, 1.0 / 15.0) PVD_STAGE(p, s2, 1.0 / 13.0)
    PVD_STAGE(p, s2, 1.0 / 11.0) PVD_STAGE(p, s2, 1.0 / 9.0) PVD_STAGE(p, s2, 1.0 / 7.0) PVD_STAGE(p, s2, 1.0 / 5.0)
    PVD_STAGE(p, s2, 1.0 / 3.0)
    PVD_STAGE(sn, a2, 1.0 / 1307674368000.0) PVD_STAGE(sn, a2, -1.0 / 6227020800.0) PVD_STAGE(sn, a2, 1.0 / 39916800.0)
    PVD_STAGE(sn, a2, -1.0 / 362880.0) PVD_STAGE(sn, a2, 1.0 / 5040.0) PVD_STAGE(sn, a2, -1.0 / 120.0) PVD_STAGE(sn, a2, 1.0 / 6.0)
    PVD_STAGE(cs, a2, -1.0 / 87178291200.0) PVD_STAGE(cs, a2, 1.0 / 479001600.0) PVD_STAGE(cs, a2, -1.0 / 3628800.0)
    PVD_STAGE(cs, a2, 1.0 / 40320.0) PVD_STAGE(cs, a2, -1.0 / 720.0) PVD_STAGE(cs, a2, 1.0 / 24.0) PVD_STAGE(cs, a2, -0.5)
    PVD_STAGE(cs, a2, 1.0)
#undef PVD_STAGE
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const double at = fma(p[q] * s2[q], s[q], s[q]);                                   // atanh(s)
        const double t = fma(kk[q], 1.3862943611198906, -4.0 * at);                        // -2 ln u
        double y = (double)rsqrtf((float)t);
        y = y * fma(-0.5 * t, y * y, 1.5);
        y = y * fma(-0.5 * t, y * y, 1.5);
        rad[q] = t * y;
        const double sine = fma(-sn[q] * a2[q], a[q], a[q]);
        const int j = ((oct[q] + 1) >> 1) & 3;
        const double c0 = (j & 1) ? sine : cs[q], s0 = (j & 1) ? cs[q] : sine;
        z0[q] = rad[q] * ((j == 1 || j == 2) ? -c0 : c0);
        z1[q] = rad[q] * ((j == 2 || j == 3) ? -s0 : s0);
    }
}

// ---------------------------------------------------------------- ziggurat normals (PVD_RNG_ZIGGURAT)
// Marsaglia & Tsang's ziggurat (J. Stat. Softw. 5(8), 2000) in Doornik's corrected form (layer index, sign and
// abscissa from DISJOINT random bits), 1024 layers, fp64: the method NumPy's Generator.normal uses (256 layers
// there).  The distribution is exact (no approximation of any transcendental on the common path): with
// probability 0.9957 a normal costs one table look-up, one subtraction, one multiplication and one comparison;
// the rest (wedges: one exp; tail beyond R = 4.03: two logs) is handled per walker after all its components took
// the common path, so a warp diverges about once per tile of 32 x 9 normals.
//   bits of one normal (64 of the 128 a Philox call returns): lo[9:0] layer, lo[10] sign, hi:lo[31:12] 52-bit abscissa
//   retries: stream purpose 16 + component, call = attempt -> still a pure function of (seed; walker, step, component)
constexpr int PVD_ZIG_N = 1024;
__device__ double2 g_zig_xr[PVD_ZIG_N];         // {x_i, x_{i+1} / x_i}; x_0 = V / f(R) (base strip), x_1 = R, x_N = 0
__device__ double g_zig_f[PVD_ZIG_N + 1];       // f(x_i) = exp(-x_i^2 / 2), f(x_N) = 1
__shared__ double2 s_zig_xr[PVD_ZIG_N];         // per-CTA copy (allocated only in kernels that call zig_stage)

// copy the table into shared memory; every thread of the CTA must call it before the first normal
__device__ __forceinline__ void zig_stage()
{
    for (int i = threadIdx.x; i < PVD_ZIG_N; i += blockDim.x) s_zig_xr[i] = g_zig_xr[i];
    __syncthreads();
}
__device__ __forceinline__ void zig_decode(unsigned lo, unsigned hi, int &layer, double &mag, unsigned &sgn)
{
    layer = (int)(lo & (unsigned)(PVD_ZIG_N - 1));
    sgn = (lo << 21) & 0x80000000u;
    mag = __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12))) - 1.0;      // [0, 1), 52 bits
}
__device__ __forceinline__ double zig_signed(double x, unsigned sgn)
{
    return __hiloint2double(__double2hiint(x) ^ (int)sgn, __double2loint(x));
}
// uniform in (0, 1] with 53 random bits (argument of a logarithm)
__device__ __forceinline__ double u53_open(unsigned lo, unsigned hi)
{
    const unsigned long long r = ((unsigned long long)hi << 32) | lo;
    return (double)((r >> 11) + 1ull) * 0x1.0p-53;
}
// Everything but the common path, for component `comp` whose first candidate (lo, hi) was not accepted outright.
__device__ __forceinline__ double zig_slow(uint64_t seed, long long slot, long long step, int comp, unsigned lo, unsigned hi)
{
    const double R = s_zig_xr[1].x;
    for (unsigned attempt = 0; attempt < 250u; ++attempt) {
        int layer;
        double mag;
        unsigned sgn;
        zig_decode(lo, hi, layer, mag, sgn);
        const double2 t = s_zig_xr[layer];
        if (mag < t.y) return zig_signed(mag * t.x, sgn);                    // a fresh candidate on the common path
        uint4 q = pvd_draw(seed, slot, step, 16u + (unsigned)comp, attempt);
        if (layer == 0) {
            // tail beyond R (Marsaglia 1964): x = -ln(U1)/R, accept when -2 ln(U2) > x^2
            for (unsigned more = attempt + 1u;; ++more) {
                const double xx = -log(u53_open(q.x, q.y)) / R, yy = -log(u53_open(q.z, q.w));
                if (yy + yy > xx * xx || more >= 250u) return zig_signed(R + xx, sgn);
                q = pvd_draw(seed, slot, step, 16u + (unsigned)comp, more);
            }
        }
        // wedge of layer i: uniform height between f(x_i) and f(x_{i+1}) against the density
        const double x = mag * t.x, f0 = g_zig_f[layer], f1 = g_zig_f[layer + 1];
        if (fma(u53(q.x, q.y), f1 - f0, f0) < exp(-0.5 * x * x)) return zig_signed(x, sgn);
        lo = q.z;
        hi = q.w;
    }
    return 0.0;    // unreachable in practice (each attempt succeeds with probability > 1/2)
}
// one normal, common path or not, decided on the spot (kernels with a run-time number of components)
__device__ __forceinline__ double zig_normal(uint64_t seed, long long slot, long long step, int comp, unsigned lo, unsigned hi)
{
    int layer;
    double mag;
    unsigned sgn;
    zig_decode(lo, hi, layer, mag, sgn);
    const double2 t = s_zig_xr[layer];
    if (mag < t.y) return zig_signed(mag * t.x, sgn);
    return zig_slow(seed, slot, step, comp, lo, hi);
}
// NC normals of one walker: all components take the common path first, the exceptions are settled afterwards
template <int NC>
__device__ __forceinline__ void walker_normals_zig(uint64_t seed, long long slot, long long step, double (&z)[NC])
{
    static_assert(NC <= 32, "one pending bit per component");
    constexpr int NP = (NC + 1) / 2;
    unsigned pend = 0u;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const uint4 r = pvd_draw(seed, slot, step, PVD_STREAM_DISP, (unsigned)k);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 2 * k + h;
            if (c < NC) {
                int layer;
                double mag;
                unsigned sgn;
                zig_decode(h ? r.z : r.x, h ? r.w : r.y, layer, mag, sgn);
                const double2 t = s_zig_xr[layer];
                z[c] = zig_signed(mag * t.x, sgn);
                if (!(mag < t.y)) pend |= 1u << c;
            }
        }
    }
    while (pend) {
        const int c = __ffs((int)pend) - 1;
        pend &= pend - 1u;
        const uint4 r = pvd_draw(seed, slot, step, PVD_STREAM_DISP, (unsigned)(c >> 1));
        const double v = zig_slow(seed, slot, step, c, (c & 1) ? r.z : r.x, (c & 1) ? r.w : r.y);
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (j == c) z[j] = v;
    }
}

// NC standard normals for walker `slot` at `step` into z[0..NC)
template <int NC, int MODE>
__device__ __forceinline__ void walker_normals(uint64_t seed, long long slot, long long step, double (&z)[NC])
{
    constexpr int NP = (NC + 1) / 2;
    if constexpr (MODE == PVD_RNG_ZIGGURAT) {
        walker_normals_zig<NC>(seed, slot, step, z);
    } else if constexpr (MODE == PVD_RNG_FP64) {
        uint4 r[NP];
        double z0[NP], z1[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) r[k] = pvd_draw(seed, slot, step, PVD_STREAM_DISP, (unsigned)k);
        normal_pairs_fp64<NP>(r, z0, z1);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            z[2 * k] = z0[k];
            if (2 * k + 1 < NC) z[2 * k + 1] = z1[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            double a, b;
            normal_pair<MODE>(pvd_draw(seed, slot, step, PVD_STREAM_DISP, (unsigned)k), a, b);
            z[2 * k] = a;
            if (2 * k + 1 < NC) z[2 * k + 1] = b;
        }
    }
}

// Counter-based random numbers: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.
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
        const double u1 = ((double)((((unsigned long long)r.y << 32) | r.x) >> 11) + 1.0) * 0x1.0p-53;   // (0,1]
        const double u2 = u53(r.z, r.w);                                                                 // [0,1)
        const double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        z0 = rad * c;
        z1 = rad * s;
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

// NC standard normals for walker `slot` at `step` into z[0..NC)
template <int NC, int MODE>
__device__ __forceinline__ void walker_normals(uint64_t seed, long long slot, long long step, double (&z)[NC])
{
#pragma unroll
    for (int k = 0; k < (NC + 1) / 2; ++k) {
        double a, b;
        normal_pair<MODE>(pvd_draw(seed, slot, step, PVD_STREAM_DISP, (unsigned)k), a, b);
        z[2 * k] = a;
        if (2 * k + 1 < NC) z[2 * k + 1] = b;
    }
}

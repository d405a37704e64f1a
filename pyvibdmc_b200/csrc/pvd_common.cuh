// Shared device/host helpers for libpvd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <string>
#include <vector>
#include "../../include/pvd_b200.h"

// ---------------------------------------------------------------- host-side error plumbing
extern thread_local std::string g_pvd_err;
extern std::atomic<long long> g_pvd_launches;

static inline int pvd_fail(int code, const std::string &msg)
{
    g_pvd_err = msg;
    return code;
}

#define PVD_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return pvd_fail(PVD_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

#define PVD_CHECK_LAUNCH()                                                                          \
    do {                                                                                            \
        g_pvd_launches.fetch_add(1, std::memory_order_relaxed);                                     \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess)                                                                     \
            return pvd_fail(PVD_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__));  \
    } while (0)

#define PVD_REQUIRE(cond, msg)                                                                      \
    do {                                                                                            \
        if (!(cond)) return pvd_fail(PVD_E_ARG, std::string(msg));                                  \
    } while (0)

static const char *const PVD_MASSIVE_MSG = "Massive walker birth or death event!!!!!!! Dying...";

// ---------------------------------------------------------------- tiling
// A tile is one warp's worth of walkers (light kernels: a few such sub-tiles).  Every warp of the (persistent) grid
// takes tiles from one ticket counter per step parity; the tile loops of the step kernels have no __syncthreads.
constexpr int PVD_TILE = 32;           // walkers per tile == one warp, one walker per lane
constexpr int PVD_CTA = 256;           // threads per CTA
constexpr int PVD_WARPS = PVD_CTA / 32;
constexpr int PVD_TICKET_STRIDE = 32;  // uints between the two parities' ticket counters (128 B: one per L2 line)

// error bits kept in DevState::err
enum : unsigned {
    PVD_ERR_WEIGHT = 1u,    // non-finite weight or weight > 1.5 N0 + 1   (pyvibdmc.py:397-400)
    PVD_ERR_POP = 2u,       // population outside [0.5, 1.5] N0            (pyvibdmc.py:409-413, 714-717)
    PVD_ERR_CAPACITY = 4u,  // shard buffer too small for the branched population
    PVD_ERR_EMPTY = 8u,     // no walkers left on this shard
    PVD_ERR_COMM = 16u      // a peer's per-step message did not arrive (mailbox collective)
};

// Device-resident simulation state.  Two copies are kept (index = step parity): the kernel of
// step s reads st[s&1] and its finalisation writes st[(s+1)&1], so warps that start late never
// observe a half-updated state, and the host never has to synchronise between steps.
struct DevState {
    long long n;            // walkers on this shard
    long long step;         // number of completed propagation steps (also the RNG counter)
    double vref;
    double pop_global;      // len(walkers) or sum(w) over all shards
    double dt_eff;          // effective time step for weighting (imp-samp: dt * accept fraction)
    double eff_time;        // accumulated effective time (pyvibdmc.py:372-378)
    unsigned err;
    int dw_active;          // descendant-weighting window open (who_from is carried)
    unsigned done;          // finished-warp counter of the step that READS this copy
    int buf;                // ping-pong buffer that holds the valid walkers (meaningful to the host when err != 0)
    long long n_accept;     // imp-samp: accepted moves in the current step (global)
    long long n_kill;       // continuous: walkers below the lower threshold in the current step
};

// Exact accumulation: floating sums that feed Vref are kept as 128-bit fixed-point integers
// (quantum 2^-80, range +-2^46), so the result does not depend on which warp processed which tile
// or in which order partial sums are combined: Vref is bit-reproducible from run to run.
struct Fx128 {
    long long hi;
    unsigned long long lo;
};
__device__ __forceinline__ Fx128 fx_zero() { return Fx128{0ll, 0ull}; }
__device__ __forceinline__ Fx128 fx_add(Fx128 a, Fx128 b)
{
    Fx128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (long long)(r.lo < a.lo ? 1 : 0);
    return r;
}
// truncating conversion of a finite double (|x| clamped below 2^46) to fixed point
__device__ __forceinline__ Fx128 fx_from_double(double x)
{
    if (!(fabs(x) < 7.0e13)) x = (x != x) ? 0.0 : copysign(7.0e13, x);
    const long long bits = __double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7ff);
    unsigned long long m = (unsigned long long)bits & 0xFFFFFFFFFFFFFull;
    if (e) m |= 1ull << 52; else e = 1;
    const int sh = e - 995;                       // value * 2^80 = m * 2^(e - 1075 + 80)
    unsigned long long h = 0ull, l = 0ull;
    if (sh >= 64) h = m << (sh - 64);
    else if (sh > 0) { l = m << sh; h = m >> (64 - sh); }
    else if (sh == 0) l = m;
    else if (sh > -64) l = m >> (-sh);
    if (bits < 0) { l = ~l + 1ull; h = ~h + (l == 0ull ? 1ull : 0ull); }
    return Fx128{(long long)h, l};
}
__device__ __forceinline__ Fx128 fx_mul_small(Fx128 a, int k)       // k in [0, 2^20)
{
    const unsigned long long lo_lo = (a.lo & 0xffffffffull) * (unsigned long long)k;
    const unsigned long long lo_hi = (a.lo >> 32) * (unsigned long long)k;
    Fx128 r;
    r.lo = lo_lo + (lo_hi << 32);
    const unsigned long long carry = (lo_hi >> 32) + ((r.lo < lo_lo) ? 1ull : 0ull);
    r.hi = a.hi * (long long)k + (long long)carry;
    return r;
}
__device__ __forceinline__ double fx_to_double(Fx128 a)
{
    const bool neg = a.hi < 0;
    unsigned long long h = (unsigned long long)a.hi, l = a.lo;
    if (neg) { l = ~l + 1ull; h = ~h + (l == 0ull ? 1ull : 0ull); }
    const double v = (double)h * 0x1.0p-16 + (double)l * 0x1.0p-80;      // exact scalings (== ldexp), one rounding in the sum
    return neg ? -v : v;
}
__device__ __forceinline__ Fx128 fx_warp_sum(Fx128 a)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Fx128 b;
        b.hi = __shfl_xor_sync(0xffffffffu, a.hi, off);
        b.lo = __shfl_xor_sync(0xffffffffu, a.lo, off);
        a = fx_add(a, b);
    }
    return a;
}

// per-warp partial sums, combined by the last warp to finish.  cv / v are exact (fixed point);
// the remaining sums are integer-valued doubles, so every field is order independent.
struct WarpPartial {
    Fx128 cv;       // sum count*V   (continuous: sum w*V over kept walkers)
    Fx128 v;        // sum V before branching
    Fx128 cw;       // continuous: sum w over kept walkers
    double c;       // discrete: sum count
    double vmin, vmax;
    double wmin, wmax;
    double births, deaths, n_in, n_acc;
};

// Programmatic dependent launch (host side: launch_pdl).  A kernel launched with programmatic stream serialisation may
// become resident while its predecessor in the stream is still draining; pdl_wait() returns when the predecessor has
// completed and its writes are visible, and lets this kernel's own successor be launched in turn.  Nothing that reads or
// writes simulation state may precede it.  Without the launch attribute it returns at once.
__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// decoupled look-back status word: [63:34] step stamp, [33:32] state, [31:0] value
constexpr unsigned long long PVD_ST_AGG = 1ull, PVD_ST_PREFIX = 2ull;
__device__ __forceinline__ unsigned long long pack_status(long long step, unsigned long long state, unsigned value)
{
    return (((unsigned long long)((step + 1) & 0x3FFFFFFFll)) << 34) | (state << 32) | (unsigned long long)value;
}
__device__ __forceinline__ bool status_valid(unsigned long long w, long long step)
{
    return (w >> 34) == (unsigned long long)((step + 1) & 0x3FFFFFFFll) && ((w >> 32) & 3ull) != 0ull;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// Tickets in batches: a ticket t covers tiles [t*batch, (t+1)*batch).  One counter serves about one atomic per
// 2-2.5 ns whatever the grid, so kernels with light tiles take several tiles per ticket; tile ids are still
// handed out in increasing order, which is all the look-back needs.
struct TileFeed {
    long long next = 0, end = 0;
};
__device__ __forceinline__ long long feed_next(TileFeed &f, unsigned *tickets, long long ntiles, int batch)
{
    if (f.next >= f.end) {
        const int lane = threadIdx.x & 31;
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(&tickets[0], 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        f.next = (long long)t * batch;
        f.end = f.next + batch < ntiles ? f.next + batch : ntiles;
        if (f.next >= ntiles) { f.end = f.next; return -1; }
    }
    return f.next++;
}

// The same feed in two halves, so that the ticket's round trip to L2 (~0.45 us) overlaps other work: feed_issue()
// starts the atomic (lane 0, only when the current batch is used up), feed_take() consumes its result later.
__device__ __forceinline__ unsigned feed_issue(const TileFeed &f, unsigned *tickets)
{
    unsigned t = 0;
    if (f.next >= f.end && (threadIdx.x & 31) == 0) t = atomicAdd(&tickets[0], 1u);
    return t;
}
__device__ __forceinline__ long long feed_take(TileFeed &f, unsigned issued, long long ntiles, int batch)
{
    if (f.next >= f.end) {
        const unsigned t = __shfl_sync(0xffffffffu, issued, 0);
        f.next = (long long)t * batch;
        f.end = f.next + batch < ntiles ? f.next + batch : ntiles;
        if (f.next >= ntiles) { f.end = f.next; return -1; }
    }
    return f.next++;
}

// warp-wide inclusive scan of one int per lane
__device__ __forceinline__ int warp_incl_scan(int c)
{
    const int lane = threadIdx.x & 31;
    int x = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    return x;
}

// Chained scan with decoupled look-back, one status word per tile, driven by the tile's own warp.
// publish_aggregate() is called as soon as a tile's total is known; resolve_prefix() -- possibly
// much later, after the warp has already computed its next tile -- walks back over the
// predecessors' status words (32 at a time) and returns the number of output slots used by
// earlier tiles, then upgrades the tile's status to an inclusive prefix.
__device__ __forceinline__ void publish_aggregate(unsigned long long *status, long long tile, long long step, int tile_total)
{
    if ((threadIdx.x & 31) == 0)
        st_relaxed_u64(&status[tile], pack_status(step, tile == 0 ? PVD_ST_PREFIX : PVD_ST_AGG, (unsigned)tile_total));
}
__device__ __forceinline__ long long resolve_prefix(unsigned long long *status, long long tile, long long step, int tile_total, bool *timed_out = nullptr)
{
    const int lane = threadIdx.x & 31;
    if (tile == 0) return 0;
    long long running = 0;
    long long look = tile - 1;
    while (true) {
        // two windows of 32 predecessors per round trip to L2: the nearer one usually holds only aggregates
        // (its tiles resolve at about the same time as this one), the farther one an inclusive prefix
        const long long idx0 = look - lane, idx1 = look - 32 - lane;
        unsigned long long w0 = pack_status(step, PVD_ST_PREFIX, 0u), w1 = w0;   // virtual tiles before tile 0
        bool ok = true;
        unsigned spins = 0;
        long long t0 = 0;
        while (true) {
            if (idx0 >= 0) w0 = ld_relaxed_u64(&status[idx0]);
            if (idx1 >= 0) w1 = ld_relaxed_u64(&status[idx1]);
            ok = status_valid(w0, step) && status_valid(w1, step);
            if (__all_sync(0xffffffffu, ok)) break;
            __nanosleep(20);              // short back-off (64 ns cost 1.5 us per step at 20 000 walkers: most waiting is in the tail)
            // a predecessor that never publishes (a bug, or a warp that left a dying run) must not hang the device
            if (++spins >= 4096u) {
                if (t0 == 0) t0 = clock64();
                else if (__any_sync(0xffffffffu, clock64() - t0 > 120000000000ll)) { if (timed_out) *timed_out = true; return 0; }
            }
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, ((w0 >> 32) & 3ull) == PVD_ST_PREFIX);
        const unsigned m1 = __ballot_sync(0xffffffffu, ((w1 >> 32) & 3ull) == PVD_ST_PREFIX);
        long long contrib = 0;
        if (m0) {
            const int first = __ffs(m0) - 1;                            // nearest tile holding an inclusive prefix
            if (lane <= first) contrib = (long long)(unsigned)(w0 & 0xffffffffull);
        } else {
            contrib = (long long)(unsigned)(w0 & 0xffffffffull);
            if (m1) {
                const int first = __ffs(m1) - 1;
                if (lane <= first) contrib += (long long)(unsigned)(w1 & 0xffffffffull);
            } else contrib += (long long)(unsigned)(w1 & 0xffffffffull);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, off);
        running += contrib;
        if (m0 | m1) break;
        look -= 64;
    }
    if (lane == 0) st_relaxed_u64(&status[tile], pack_status(step, PVD_ST_PREFIX, (unsigned)(running + tile_total)));
    return running;
}

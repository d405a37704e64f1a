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
constexpr int PVD_TILE = 256;          // walkers per tile == threads per CTA (one walker per thread)
constexpr int PVD_WARPS = PVD_TILE / 32;

// error bits kept in DevState::err
enum : unsigned {
    PVD_ERR_WEIGHT = 1u,    // non-finite weight or weight > 1.5 N0 + 1   (pyvibdmc.py:397-400)
    PVD_ERR_POP = 2u,       // population outside [0.5, 1.5] N0            (pyvibdmc.py:409-413, 714-717)
    PVD_ERR_CAPACITY = 4u,  // shard buffer too small for the branched population
    PVD_ERR_EMPTY = 8u      // no walkers left on this shard
};

// Device-resident simulation state.  Two copies are kept (index = step parity): the kernel of
// step s reads st[s&1] and its finalisation writes st[(s+1)&1], so CTAs that start late never
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
    unsigned ticket;        // dynamic tile counter of the step that READS this copy
    unsigned done;          // finished-tile counter of the step that READS this copy
    long long n_accept;     // imp-samp: accepted moves in the current step (this shard)
    long long pad;
};

// per-tile partial sums, reduced in a fixed order by the last CTA (deterministic Vref)
struct TilePartial {
    double cv;      // sum count*V   (continuous: sum w*V over kept walkers)
    double c;       // sum count     (continuous: sum w over kept walkers)
    double v;       // sum V before branching
    double vmin, vmax;
    double wmin, wmax;
    int births, deaths;
    int n_in;       // walkers this tile consumed
    int n_acc;      // imp-samp accepted
};

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

// block-wide exclusive scan of one int per thread (PVD_TILE threads); returns exclusive prefix,
// total in *total.  smem: PVD_WARPS+1 ints.
__device__ __forceinline__ int block_excl_scan(int c, int *smem, int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) smem[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int t = lane < PVD_WARPS ? smem[lane] : 0;
        int s = t;
#pragma unroll
        for (int off = 1; off < PVD_WARPS; off <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += y;
        }
        if (lane < PVD_WARPS) smem[lane] = s - t;     // exclusive warp offsets
        if (lane == PVD_WARPS - 1) smem[PVD_WARPS] = s;
    }
    __syncthreads();
    const int excl = x - c + smem[wid];
    *total = smem[PVD_WARPS];
    __syncthreads();
    return excl;
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
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Tile-level exclusive prefix over the whole grid (single-pass chained scan with decoupled
// look-back).  Called by all threads; returns the number of output slots used by earlier tiles.
__device__ __forceinline__ long long tile_lookback(unsigned long long *status, int tile, long long step,
                                                   int tile_total, long long *smem_prefix)
{
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        long long running = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(&status[0], pack_status(step, PVD_ST_PREFIX, (unsigned)tile_total));
        } else {
            if (lane == 0) st_relaxed_u64(&status[tile], pack_status(step, PVD_ST_AGG, (unsigned)tile_total));
            int look = tile - 1;
            while (true) {
                const int idx = look - lane;
                unsigned long long w = 0;
                bool ok;
                do {
                    if (idx >= 0) {
                        w = ld_relaxed_u64(&status[idx]);
                        ok = status_valid(w, step);
                    } else {
                        w = pack_status(step, PVD_ST_PREFIX, 0u);   // virtual tile before tile 0
                        ok = true;
                    }
                } while (!__all_sync(0xffffffffu, ok));
                const bool is_prefix = ((w >> 32) & 3ull) == PVD_ST_PREFIX;
                const unsigned mask = __ballot_sync(0xffffffffu, is_prefix);
                const unsigned val = (unsigned)(w & 0xffffffffull);
                if (mask) {
                    const int first = __ffs(mask) - 1;              // nearest tile holding an inclusive prefix
                    long long contrib = lane <= first ? (long long)val : 0ll;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, off);
                    running += contrib;
                    break;
                }
                long long contrib = (long long)val;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, off);
                running += contrib;
                look -= 32;
            }
            if (lane == 0)
                st_relaxed_u64(&status[tile], pack_status(step, PVD_ST_PREFIX, (unsigned)(running + tile_total)));
        }
        if (lane == 0) *smem_prefix = running;
    }
    __syncthreads();
    const long long p = *smem_prefix;
    __syncthreads();
    return p;
}

// Continuous weighting (pyvibdmc.py:432-454 + _branch :340-356).
//
// Reference semantics: w *= exp(-(V-Vref) dt); every walker below the lower threshold is, in
// ascending index order, overwritten by a copy of the CURRENT maximum-weight walker (first index on
// ties), whose weight is halved and shared.  That loop is sequential by construction (argmax after
// every halving).  Parallel-equivalent used here:
//   1. k_cont_update   : weight update, exact sums over the surviving walkers, kill flags as one ballot word per
//                        32 walkers (no warp waits for another: no ticket counter, no look-back scan)
//                        and the log-spaced histogram (64 bins / octave) of the updated weights
//   3. k_cont_prefix   : suffix sums of the histogram -> the bin that holds the K-th largest weight
//                        (K = number of kills) and, per bin, where its members start in the
//                        candidate array
//      k_cont_collect  : candidates = all walkers at or above that bin, bucketed by bin
//      k_cont_rank     : one CTA per bin orders its bucket by (w desc, index asc) in shared memory
//                        (1024 sub-buckets + all-pairs rank inside a sub-bucket) and writes it to
//                        its place in the sorted candidate array.  If some bin is too full for that (> 8192: many
//                        identical weights), the candidates are appended unordered instead and
//                        the donor assignment sorts them with a single-CTA bitonic network in global memory.
//                        CTA 0 of the same kernel makes the ascending kill list from the ballot words meanwhile.
//   4. cont_assign     : (the last CTA of k_cont_rank to finish) if the K-th largest weight is > half the largest, the K argmax steps are
//                        exactly "j-th largest donates to j-th kill" and are applied in parallel;
//                        otherwise (halved pieces re-enter the top: start-up transients) the
//                        reference loop is replayed exactly by one thread over the sorted
//                        candidates.  The optional upper threshold (argpartition of the smallest
//                        weights) is replayed with block-wide argmin/argmax reductions.
//   5. k_cont_copy     : walkers (all per-walker arrays) are copied donor-root -> killed slot
//   6. k_cont_finish   : min/max weight, Vref, log record (last CTA)
#pragma once
#include "pvd_step.cuh"

constexpr int PVD_HIST_BINS = 4096;            // 64 octaves x 64 bins
constexpr int PVD_HIST_BASE = (1023 - 32) * 64;
constexpr unsigned PVD_RANK_MAX_BIN = 8192;    // fuller bins fall back to the single-CTA sort

struct ContCand { double w; int idx; int root; };
struct ContWork {
    unsigned n_kill;        // walkers below the lower threshold (set by the update kernel's scan)
    unsigned n_cand;        // candidates appended by k_cont_collect
    unsigned n_copy;        // (dst, src) pairs to copy
    unsigned n_upper;       // kills caused by the upper threshold
    int edge_bin;
    int fast;               // 1/2: k_cont_copy pairs j-th largest (sorted / cand array) with the j-th kill itself
    int ranked;             // 1: candidates are bucketed by bin and ranked in parallel; 0: unordered + single-CTA sort
    unsigned done;
    unsigned done_rank;     // CTAs of k_cont_rank that have finished (the last one assigns the donors)
    double sub_w, sub_wv;   // weight removed by upper-threshold kills (corrects the exact sums)
    unsigned long long wmax_bits, wmin_bits;
};

struct ContArgs {
    double *w;
    const double *v;
    int *kill_idx;
    unsigned *kill_mask;        // one ballot word per 32 walkers: who fell below the lower threshold in this step (k_cont_update)
    int *copy_dst, *copy_src;
    ContCand *cand;
    ContCand *sorted;           // candidates in (w desc, idx asc) order (ranked path)
    unsigned *hist;
    unsigned *bin_start;        // [PVD_HIST_BINS] candidates in higher bins
    unsigned *bin_fill;         // [PVD_HIST_BINS] running fill of each bin's bucket
    ContWork *work;
    long long cand_cap;
    double lower, upper;
    int has_upper;
};

__device__ __forceinline__ int weight_bin(double w)
{
    const long long bits = __double_as_longlong(w);
    long long b = (bits >> 46) - PVD_HIST_BASE;          // exponent and top 6 mantissa bits
    if (!(w > 0.0)) b = 0;
    return (int)(b < 0 ? 0 : (b >= PVD_HIST_BINS ? PVD_HIST_BINS - 1 : b));
}
__device__ __forceinline__ double bin_lower_edge(int bin)
{
    return __longlong_as_double(((long long)bin + PVD_HIST_BASE) << 46);
}

// ---- 1+2. weight update + kill list + exact sums + histogram of the updated weights (energies from memory: a.vin)
// A tile is SUB sub-tiles of 32 walkers (lane l of sub-tile s owns walker tile * 32 SUB + s * 32 + l); the kill flags of a
// sub-tile are one ballot.
// Where the energy of walker i comes from: memory (external / NN / importance-sampled energies)...
struct ContFromMemory {
    static constexpr int RNG_MODE = PVD_RNG_FP64;      // draws no normals
    static constexpr int MIN_CTAS = 4;
    static constexpr int SUB = 8;
    __device__ static __forceinline__ double produce(const StepArgs &a, long long i, long long) { return a.vin[i]; }
};
// ...or the fused move + built-in potential (in place: continuous weighting never compacts)
template <class POT, int RNG>
struct ContFused {
    static constexpr int RNG_MODE = RNG;
    static constexpr int MIN_CTAS = POT::MIN_CTAS;
    static constexpr int SUB = POT::MIN_CTAS >= 4 ? 4 : 1;      // heavy potential: small tiles keep the warps balanced
    __device__ static __forceinline__ double produce(const StepArgs &a, long long i, long long step)
    {
        double x[POT::NC], v;
        ProduceFused<POT, RNG>::run(a, i, step, true, x, v);
        double *po = a.xout + i;
#pragma unroll
        for (int c = 0; c < POT::NC; ++c) { *po = x[c]; po += a.cap; }
        a.vout[i] = v;
        return v;
    }
};

#ifndef PVD_CONT_DYNAMIC
#define PVD_CONT_DYNAMIC 0          // A/B: 1 = tiles from the global ticket counter (dynamic over the whole grid), 0 = contiguous chunk per CTA
#endif
template <class PROD>
__global__ void __launch_bounds__(PVD_CTA, PROD::MIN_CTAS) k_cont_update(const StepArgs a, const ContArgs ca)
{
    constexpr int SUB = PROD::SUB, TILE = PVD_TILE * SUB;
    __shared__ unsigned s_hist[PVD_HIST_BINS];
    __shared__ unsigned s_ticket;
    for (int b = threadIdx.x; b < PVD_HIST_BINS; b += PVD_CTA) s_hist[b] = 0;
    if (threadIdx.x == 0) s_ticket = 0u;
    if constexpr (PROD::RNG_MODE == PVD_RNG_ZIGGURAT) zig_stage();
    __syncthreads();
    pdl_wait();                                        // everything above is independent of the previous kernel
    if (!step_prologue(a)) return;
    const DevState *sip = &a.st[a.parity];
    const long long n = sip->n, step = sip->step;
    const double vref = sip->vref, dt = sip->dt_eff;
    const long long ntiles = (n + TILE - 1) / TILE;
    const int lane = threadIdx.x & 31;
    // No warp waits for another: a tile leaves its kill flags as one ballot word per 32 walkers (kill_mask, bit l of word
    // i / 32 = walker i), and the ascending kill list is made from the words afterwards (k_cont_prefix).  Tiles are dealt to
    // CTAs in contiguous chunks, the warps of a CTA share their chunk through a shared-memory ticket: no global ticket
    // counter (one atomic per 2.5 ns at best: 31 250 tiles of a 1e6-walker H2O step kept it busy for most of the step) and no
    // chained look-back over tile totals, which is what the deferred-compaction step did for discrete weighting (pvd_gather.cuh).
    LaneAcc acc;
#if PVD_CONT_DYNAMIC
    unsigned *tickets = step_tickets(a, a.parity);
    TileFeed feed;
    long long tile = feed_next(feed, tickets, ntiles, 1);
    while (tile >= 0) {
        const unsigned issued = feed_issue(feed, tickets);     // the next ticket travels while this tile is computed
#else
    const long long t_lo = (ntiles * (long long)blockIdx.x) / gridDim.x, t_hi = (ntiles * ((long long)blockIdx.x + 1)) / gridDim.x;
    while (true) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(&s_ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        const long long tile = t_lo + t;
        if (tile >= t_hi) break;
#endif
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            const long long i = tile * TILE + s * PVD_TILE + lane;
            bool kill = false;
            if (i < n) {
                const double v = PROD::produce(a, i, step);
                const double wn = __dmul_rn(ca.w[i], exp(__dmul_rn(-1.0 * (v - vref), dt)));      // :433
                ca.w[i] = wn;
                kill = wn < ca.lower;                                                             // :436
                atomicAdd(&s_hist[weight_bin(wn)], 1u);
                const Fx128 fv = fx_from_double(v);
                acc.v = fx_add(acc.v, fv);
                if (!kill) {
                    acc.cw = fx_add(acc.cw, fx_from_double(wn));
                    acc.cv = fx_add(acc.cv, fx_from_double(__dmul_rn(wn, v)));
                }
                acc.vmin = fmin(acc.vmin, v); acc.vmax = fmax(acc.vmax, v);
                acc.n_in += 1.0; acc.n_acc += 1.0;
                acc.births += kill ? 1.0 : 0.0;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, kill);
            if (lane == 0) ca.kill_mask[tile * SUB + s] = mask;
        }
#if PVD_CONT_DYNAMIC
        tile = feed_take(feed, issued, ntiles, 1);
#endif
    }
    __syncthreads();
    for (int b = threadIdx.x; b < PVD_HIST_BINS; b += PVD_CTA)
        if (s_hist[b]) atomicAdd(&ca.hist[b], s_hist[b]);
    // population does not change; Vref is produced by k_cont_finish, after branching
    cta_finish_step(a, acc, ntiles, true, n, true);
}

// (w desc, idx asc) ordering
__device__ __forceinline__ bool cand_before(const ContCand &p, const ContCand &q)
{
    return p.w > q.w || (p.w == q.w && p.idx < q.idx);
}

// ---- 3a. suffix sums of the histogram (one CTA, one bin per 4 threads' worth of work)
__global__ void __launch_bounds__(1024) k_cont_prefix(const StepArgs a, const ContArgs ca)
{
    pdl_wait();
    __shared__ unsigned s_warp[32];
    __shared__ int s_edge;
    __shared__ unsigned s_maxbin;
    const DevState *so = &a.st[a.parity];          // continuous weighting: population and flags do not change within a step
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    // the number of walkers below the threshold is one of the step's sums (k_cont_update's last CTA left it there; the sums stay
    // local to the shard until k_cont_finish); the kill LIST is made from the ballot words by CTA 0 of k_cont_rank, off the
    // critical path (nothing before the donor assignment reads it)
    const unsigned nk = (so->err || so->n <= 0) ? 0u : (unsigned)a.sums[PVD_SUM_BIRTHS];
    if (threadIdx.x == 0) {
        ca.work->n_kill = nk; ca.work->n_cand = 0; ca.work->n_copy = 0; ca.work->n_upper = 0; ca.work->done = 0; ca.work->done_rank = 0;
        ca.work->sub_w = 0.0; ca.work->sub_wv = 0.0; ca.work->ranked = 0; ca.work->fast = 0;
        ca.work->wmax_bits = 0ull; ca.work->wmin_bits = 0x7FF0000000000000ull;
    }
    if (so->err || nk == 0) return;
    __syncthreads();
    constexpr int PER = PVD_HIST_BINS / 1024;                 // 4 consecutive bins per thread, highest bins first
    if (t == 0) { s_edge = -1; s_maxbin = 0u; }
    unsigned h[PER], mine = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { h[k] = ca.hist[PVD_HIST_BINS - 1 - (t * PER + k)]; mine += h[k]; }
    unsigned incl = mine;
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned v = s_warp[lane], x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        s_warp[lane] = x - v;                                 // exclusive prefix over warps
    }
    __syncthreads();
    unsigned above = s_warp[wid] + incl - mine;               // members of strictly higher bins than this thread's first bin
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int bin = PVD_HIST_BINS - 1 - (t * PER + k);
        ca.bin_start[bin] = above;
        if (above < nk && above + h[k] >= nk) s_edge = bin;   // exactly one bin satisfies this when enough weights are positive
        above += h[k];
    }
    __syncthreads();
    const int edge = s_edge < 0 ? 0 : s_edge;                 // fewer than K positive weights: take everything
    unsigned mb = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int bin = PVD_HIST_BINS - 1 - (t * PER + k);
        if (bin >= edge) mb = max(mb, h[k]);
    }
    for (int off = 16; off > 0; off >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, off));
    if (lane == 0 && mb) atomicMax(&s_maxbin, mb);
    __syncthreads();
    if (t == 0) {
        ca.work->edge_bin = edge;
        const bool ranked = s_maxbin <= PVD_RANK_MAX_BIN;
        ca.work->ranked = ranked ? 1 : 0;
        if (ranked) ca.work->n_cand = ca.bin_start[edge] + ca.hist[edge];
    }
}

// ---- 3b. candidates: everything at or above the bin that contains the K-th largest weight
__global__ void __launch_bounds__(PVD_CTA) k_cont_collect(const StepArgs a, const ContArgs ca)
{
    pdl_wait();
    const DevState *so = &a.st[a.parity];
    const unsigned nk = ca.work->n_kill;
    if (so->err || nk == 0) return;
    const long long n = so->n;
    const int edge_bin = ca.work->edge_bin;
    const bool ranked = ca.work->ranked != 0;
    for (long long i = blockIdx.x * (long long)PVD_CTA + threadIdx.x; i < n; i += (long long)gridDim.x * PVD_CTA) {
        const double w = ca.w[i];
        const int bin = weight_bin(w);
        if (bin >= edge_bin) {
            const unsigned slot = ranked ? ca.bin_start[bin] + atomicAdd(&ca.bin_fill[bin], 1u) : atomicAdd(&ca.work->n_cand, 1u);
            if ((long long)slot < ca.cand_cap) ca.cand[slot] = ContCand{w, (int)i, (int)i};
        }
    }
}

// ---- 3c. order inside each bin (one CTA per bin) -> sorted candidates
// The bin's bucket is loaded into shared memory, split into 1024 sub-buckets by the next 10 mantissa bits
// (weights inside one 1/64-octave bin are close to uniformly spread, so sub-buckets hold a handful of members),
// and every member's final position is the start of its sub-bucket plus an all-pairs count inside it.
constexpr int PVD_RANK_SUB = 1024;
constexpr size_t PVD_RANK_SMEM = (size_t)PVD_RANK_MAX_BIN * (sizeof(ContCand) + sizeof(unsigned short)) + 3 * PVD_RANK_SUB * sizeof(unsigned);
__device__ __forceinline__ int weight_subbin(double w)
{
    return (int)((__double_as_longlong(w) >> 36) & (PVD_RANK_SUB - 1));
}
// block-wide argmax (first index on ties) / argmin over w[0..n) with a skip mask
__device__ inline int block_arg_extreme(const double *w, long long n, bool want_max, const unsigned char *skip, double *s_val, int *s_idx)
{
    double best = want_max ? -INFINITY : INFINITY;
    int bi = -1;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        if (skip && skip[i]) continue;
        const double x = w[i];
        if (bi < 0 || (want_max ? x > best : x < best)) { best = x; bi = (int)i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        const bool take = oi >= 0 && (bi < 0 || (want_max ? ob > best : ob < best) || (ob == best && oi < bi));
        if (take) { best = ob; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            const double ob = s_val[k];
            const int oi = s_idx[k];
            const bool take = oi >= 0 && (s_idx[0] < 0 || (want_max ? ob > s_val[0] : ob < s_val[0]) || (ob == s_val[0] && oi < s_idx[0]));
            if (take) { s_val[0] = ob; s_idx[0] = oi; }
        }
    }
    __syncthreads();
    const int r = s_idx[0];
    __syncthreads();
    return r;
}

// ---- 4. donor assignment (one CTA of 1024 threads: the last CTA of k_cont_rank to finish)
__device__ __forceinline__ void cont_assign(const StepArgs &a, const ContArgs &ca, ContCand *queue, int *root, unsigned char *skip)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_flag;
    const DevState *so = &a.st[a.parity];
    if (so->err) return;
    const long long n = so->n;
    const int K = (int)ca.work->n_kill;
    int ncopy = 0;
    if (K > 0) {
        long long C = ca.work->n_cand;
        if (C > ca.cand_cap) C = ca.cand_cap;
        const bool ranked = ca.work->ranked != 0;
        const ContCand *srt = ranked ? ca.sorted : ca.cand;
        if (!ranked) {
            // over-full bin (many identical weights): bitonic sort of the unordered candidates, padded to a power of two with -inf
            long long P = 1;
            while (P < C) P <<= 1;
            for (long long t = C + threadIdx.x; t < P; t += blockDim.x) ca.cand[t] = ContCand{-INFINITY, 0x7fffffff, -1};
            __syncthreads();
            for (long long k = 2; k <= P; k <<= 1) {
                for (long long j = k >> 1; j > 0; j >>= 1) {
                    for (long long t = threadIdx.x; t < P; t += blockDim.x) {
                        const long long l = t ^ j;
                        if (l > t) {
                            const ContCand x = ca.cand[t], y = ca.cand[l];
                            const bool up = (t & k) == 0;
                            if (up ? cand_before(y, x) : cand_before(x, y)) { ca.cand[t] = y; ca.cand[l] = x; }
                        }
                    }
                    __syncthreads();
                }
            }
        }
        // fast path: the K largest weights are all above half of the largest -> no halved piece is ever a donor
        if (threadIdx.x == 0) s_flag = (C >= K && srt[K - 1].w > 0.5 * srt[0].w) ? 1 : 0;
        __syncthreads();
        if (s_flag && !ca.has_upper) {
            // applied by all CTAs of k_cont_copy straight from the sorted candidates
            if (threadIdx.x == 0) ca.work->fast = ranked ? 1 : 2;
            ncopy = K;
        } else if (s_flag) {
            for (int j = threadIdx.x; j < K; j += blockDim.x) {
                const int d = srt[j].idx, k = ca.kill_idx[j];
                const double h = srt[j].w / 2.0;
                ca.w[d] = h;
                ca.w[k] = h;
                ca.copy_dst[j] = k;
                ca.copy_src[j] = d;
            }
            ncopy = K;
        } else {
            // exact sequential replay of _branch over the sorted candidates (A) and the FIFO of halves (B)
            if (threadIdx.x == 0) {
                long long ia = 0, ib = 0, nb = 0;
                for (int j = 0; j < K; ++j) {
                    const int k = ca.kill_idx[j];
                    // current maximum: front of A vs the run of equal values at the front of B (first index wins)
                    double mb = -INFINITY;
                    long long best_b = -1;
                    if (ib < nb) {
                        mb = queue[ib].w;
                        best_b = ib;
                        for (long long t = ib + 1; t < nb && queue[t].w == mb; ++t)
                            if (queue[t].idx < queue[best_b].idx) best_b = t;
                    }
                    const bool have_a = ia < C;
                    bool use_a;
                    if (!have_a && best_b < 0) {
                        // candidates exhausted (every weight below the threshold): degenerate ensemble,
                        // fall back to a plain scan for the maximum
                        double bw = -INFINITY; int bi = 0;
                        for (long long t = 0; t < n; ++t) if (ca.w[t] > bw) { bw = ca.w[t]; bi = (int)t; }
                        const double h = bw / 2.0;
                        ca.w[bi] = h; ca.w[k] = h;
                        ca.copy_dst[j] = k; ca.copy_src[j] = root ? root[bi] : bi;
                        continue;
                    }
                    if (!have_a) use_a = false;
                    else if (best_b < 0) use_a = true;
                    else use_a = srt[ia].w > mb || (srt[ia].w == mb && srt[ia].idx < queue[best_b].idx);
                    ContCand top;
                    if (use_a) top = srt[ia++];
                    else { top = queue[best_b]; queue[best_b] = queue[ib]; ++ib; }
                    const double h = top.w / 2.0;
                    ca.w[top.idx] = h;
                    ca.w[k] = h;
                    ca.copy_dst[j] = k;
                    ca.copy_src[j] = top.root;
                    queue[nb++] = ContCand{h, top.idx, top.root};
                    queue[nb++] = ContCand{h, k, top.root};
                }
            }
            ncopy = K;
        }
        __syncthreads();
        // chains: a killed slot refilled earlier may itself have donated later; src already holds roots.
    }
    // optional upper threshold (pyvibdmc.py:440-445): kill the n_above smallest weights, each refilled from the
    // current maximum.  Replayed with block-wide reductions; roots are tracked through `root`.
    if (ca.has_upper) {
        for (long long i = threadIdx.x; i < n; i += blockDim.x) { root[i] = (int)i; skip[i] = 0; }
        __syncthreads();
        for (int j = threadIdx.x; j < ncopy; j += blockDim.x) root[ca.copy_dst[j]] = ca.copy_src[j];
        __syncthreads();
        // count weights above the threshold
        int cnt = 0;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) cnt += (ca.w[i] > ca.upper) ? 1 : 0;
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if ((threadIdx.x & 31) == 0) s_idx[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_idx[k]; s_flag = t; }
        __syncthreads();
        const int n_above = s_flag;
        __syncthreads();
        // the n_above smallest weights, fixed before any of them is refilled (argpartition on the pre-branch array)
        for (int j = 0; j < n_above; ++j) {
            const int k = block_arg_extreme(ca.w, n, false, skip, s_val, s_idx);
            if (threadIdx.x == 0) { skip[k] = 1; ca.copy_dst[ncopy + j] = k; }
            __syncthreads();
        }
        for (int j = 0; j < n_above; ++j) {
            const int d = block_arg_extreme(ca.w, n, true, nullptr, s_val, s_idx);
            if (threadIdx.x == 0) {
                const int k = ca.copy_dst[ncopy + j];
                ca.work->sub_w += ca.w[k];
                ca.work->sub_wv += ca.w[k] * ca.v[root[k]];
                const double h = ca.w[d] / 2.0;
                ca.w[d] = h; ca.w[k] = h;
                ca.copy_src[ncopy + j] = root[d];
                root[k] = root[d];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) ca.work->n_upper = (unsigned)n_above;
        ncopy += n_above;
    }
    if (threadIdx.x == 0) ca.work->n_copy = (unsigned)ncopy;
    // reset the histogram and the bucket fills for the next step
    for (int b = threadIdx.x; b < PVD_HIST_BINS; b += blockDim.x) { ca.hist[b] = 0; ca.bin_fill[b] = 0; }
}

// ascending kill list from the ballot words of k_cont_update (one CTA of 1024 threads), in rounds of 16 words per thread: four
// 16-byte loads in flight per thread (the buffer is padded), the words stay in registers, one block-wide scan per round places
// their set bits
__device__ __forceinline__ void cont_kill_list(const ContArgs &ca, long long n, unsigned *s_warp, unsigned *s_tot)
{
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const long long nwords = (n + 31) / 32;
    unsigned base_k = 0;
    for (long long r0 = 0; r0 < nwords; r0 += 16 * 1024) {
        const long long w0 = r0 + 16ll * t;
        unsigned m[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (w0 + 4 * q < nwords) v = __ldcg(reinterpret_cast<const uint4 *>(ca.kill_mask + w0) + q);
            m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
        }
        unsigned mine_k = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (w0 + j >= nwords) m[j] = 0u;                  // stale words beyond the ensemble
            mine_k += __popc(m[j]);
        }
        unsigned incl_k = mine_k;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl_k, off);
            if (lane >= off) incl_k += y;
        }
        __syncthreads();                                      // the previous round's s_warp / s_tot have been read
        if (lane == 31) s_warp[wid] = incl_k;
        __syncthreads();
        if (wid == 0) {
            const unsigned v = s_warp[lane];
            unsigned x = v;
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, x, off);
                if (lane >= off) x += y;
            }
            s_warp[lane] = x - v;
            if (lane == 31) *s_tot = x;
        }
        __syncthreads();
        unsigned o = base_k + s_warp[wid] + incl_k - mine_k;
        base_k += *s_tot;
        if (mine_k) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                unsigned mm = m[j];
                while (mm) {
                    const int bit = __ffs((int)mm) - 1;
                    ca.kill_idx[o++] = (int)((w0 + j) * 32 + bit);
                    mm &= mm - 1u;
                }
            }
        }
    }
}

// ---- 3c + 4. CTA 0 makes the kill list, the others order the bins; the last CTA to finish assigns the donors
__global__ void __launch_bounds__(1024) k_cont_rank(const StepArgs a, const ContArgs ca, ContCand *queue, int *root, unsigned char *skip)
{
    pdl_wait();
    extern __shared__ __align__(16) unsigned char s_rank_raw[];
    ContCand *s_c = reinterpret_cast<ContCand *>(s_rank_raw);
    unsigned *sub_cnt = reinterpret_cast<unsigned *>(s_c + PVD_RANK_MAX_BIN);
    unsigned *sub_start = sub_cnt + PVD_RANK_SUB;
    unsigned *sub_fill = sub_start + PVD_RANK_SUB;
    unsigned short *perm = reinterpret_cast<unsigned short *>(sub_fill + PVD_RANK_SUB);
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_tot, s_last;
    const DevState *so = &a.st[a.parity];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const bool active = !so->err && ca.work->n_kill != 0;
    if (blockIdx.x == 0) {
        if (active) cont_kill_list(ca, so->n, s_warp, &s_tot);
    } else if (active && ca.work->ranked) {
        const int edge = ca.work->edge_bin;
        const int nb = (int)gridDim.x - 1;
        // CTAs stride over the bins from the top; each CTA needs 150 KB of shared memory, so the grid is kept small
        for (int bin = PVD_HIST_BINS - 1 - ((int)blockIdx.x - 1); bin >= edge; bin -= nb) {
            unsigned cnt = ca.hist[bin];
            const unsigned start = ca.bin_start[bin];
            if (cnt == 0 || (long long)start >= ca.cand_cap) continue;
            if ((long long)start + cnt > ca.cand_cap) cnt = (unsigned)(ca.cand_cap - start);
            __syncthreads();                                      // previous bin's shared arrays are no longer read
            sub_cnt[t] = 0; sub_fill[t] = 0;                      // blockDim.x == PVD_RANK_SUB
            for (unsigned e = t; e < cnt; e += blockDim.x) s_c[e] = ca.cand[start + e];
            __syncthreads();
            for (unsigned e = t; e < cnt; e += blockDim.x) atomicAdd(&sub_cnt[weight_subbin(s_c[e].w)], 1u);
            __syncthreads();
            // members of higher sub-buckets come first: thread t owns sub-bucket 1023 - t
            const unsigned mine = sub_cnt[PVD_RANK_SUB - 1 - t];
            unsigned incl = mine;
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += y;
            }
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            if (wid == 0) {
                const unsigned v = s_warp[lane];
                unsigned x = v;
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned y = __shfl_up_sync(0xffffffffu, x, off);
                    if (lane >= off) x += y;
                }
                s_warp[lane] = x - v;
            }
            __syncthreads();
            sub_start[PVD_RANK_SUB - 1 - t] = s_warp[wid] + incl - mine;
            __syncthreads();
            for (unsigned e = t; e < cnt; e += blockDim.x) {
                const int sk = weight_subbin(s_c[e].w);
                perm[sub_start[sk] + atomicAdd(&sub_fill[sk], 1u)] = (unsigned short)e;
            }
            __syncthreads();
            for (unsigned p = t; p < cnt; p += blockDim.x) {
                const ContCand me = s_c[perm[p]];
                const int sk = weight_subbin(me.w);
                const unsigned s0 = sub_start[sk], c = sub_cnt[sk];
                unsigned r = s0;
                for (unsigned q = s0; q < s0 + c; ++q) r += cand_before(s_c[perm[q]], me) ? 1u : 0u;
                ca.sorted[start + r] = me;
            }
        }
    }
    // the last CTA to get here (every CTA does, whatever it had to do) assigns the donors: it sees the kill list and the sorted
    // candidates of all the others
    __threadfence();
    __syncthreads();
    if (t == 0) {
        const unsigned d = atomicAdd(&ca.work->done_rank, 1u);
        s_last = (d == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    cont_assign(a, ca, queue, root, skip);
}

// ---- 5. copy donor -> killed for every per-walker array (in place: donors are never kill targets)
__global__ void __launch_bounds__(PVD_CTA) k_cont_copy(const StepArgs a, const ContArgs ca, double *x, double *v, int *who, double *f,
                                                       double *psi, double *lk, long long *src_out)
{
    pdl_wait();
    const DevState *so = &a.st[a.parity];
    if (so->err) return;
    const unsigned m = ca.work->n_copy;
    const bool dw = so->dw_active != 0;
    const int fast = ca.work->fast;
    const ContCand *srt = fast == 1 ? ca.sorted : ca.cand;
    for (long long t = blockIdx.x * (long long)PVD_CTA + threadIdx.x; t < (long long)m; t += (long long)gridDim.x * PVD_CTA) {
        int dst, src;
        if (fast) {
            // j-th largest weight donates to the j-th killed walker (ascending index): halve and share (:340-356)
            const ContCand d = srt[t];
            src = d.idx; dst = ca.kill_idx[t];
            const double h = d.w / 2.0;
            ca.w[src] = h;
            ca.w[dst] = h;
        } else { dst = ca.copy_dst[t]; src = ca.copy_src[t]; }
        for (int c = 0; c < a.nc; ++c) x[c * a.cap + dst] = x[c * a.cap + src];
        v[dst] = v[src];
        if (dw && who) who[dst] = who[src];
        if (f) {
            for (int c = 0; c < a.nc; ++c) f[c * a.cap + dst] = f[c * a.cap + src];
            psi[dst] = psi[src];
            lk[dst] = lk[src];
            if (a.vsout) a.vsout[dst] = a.vsout[src];
        }
        if (src_out) src_out[dst] = src;
    }
}

// ---- 6. min / max weight after branching, upper-threshold correction of the sums, Vref + log record
__global__ void __launch_bounds__(1024) k_cont_finish(const StepArgs a, const ContArgs ca)
{
    // one CTA of 1024 threads per SM: the same loads in flight as four times as many CTAs of 256, a quarter of the atomics on
    // the two extrema and the done counter (they serialise in L2 at ~2.5 ns each)
    pdl_wait();
    __shared__ double s_mx[32], s_mn[32];
    __shared__ unsigned s_last;
    const DevState *so = &a.st[a.parity];
    if (a.st[a.parity].err) return;
    const long long n = so->n;
    double mx = 0.0, mn = INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double w = ca.w[i];
        mx = fmax(mx, w); mn = fmin(mn, w);
    }
    mx = warp_max(mx); mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) { s_mx[threadIdx.x >> 5] = mx; s_mn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { mx = fmax(mx, s_mx[k]); mn = fmin(mn, s_mn[k]); }
        // positive doubles order like their bit patterns
        atomicMax(&ca.work->wmax_bits, (unsigned long long)__double_as_longlong(fmax(mx, 0.0)));
        atomicMin(&ca.work->wmin_bits, (unsigned long long)__double_as_longlong(fmax(mn, 0.0)));
        __threadfence();
        const unsigned d = atomicAdd(&ca.work->done, 1u);
        s_last = (d == gridDim.x - 1) ? 1u : 0u;
        if (s_last) {
            __threadfence();
            double *s = a.sums;
            sum_put_double(s, PVD_SUM_CV, sum_get(s, PVD_SUM_CV) - ca.work->sub_wv);
            sum_put_double(s, PVD_SUM_C, sum_get(s, PVD_SUM_C) - ca.work->sub_w);
            s[PVD_SUM_BIRTHS] = (double)(ca.work->n_kill + ca.work->n_upper);     // "Walkers Branched"
            double *e = s + PVD_SUM_EXT + 4 * a.rank;
            e[2] = __longlong_as_double((long long)atomicAdd(&ca.work->wmin_bits, 0ull));
            e[3] = __longlong_as_double((long long)atomicAdd(&ca.work->wmax_bits, 0ull));
            if (a.world == 1) finalize_from_sums(a, true);
        }
    }
    if (a.world > 1 && a.mbox[0]) {
        __syncthreads();                       // s_last was written by thread 0
        if (s_last) mailbox_send(a, a.st[a.parity].step);
    }
}

// plain weight update for the stand-alone entry point's bookkeeping of `src`
__global__ void k_iota_i64(long long *p, long long n)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) p[t] = t;
}

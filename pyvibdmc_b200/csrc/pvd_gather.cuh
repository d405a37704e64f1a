// Discrete-weighting time steps with DEFERRED compaction ("gather" steps): the body of DMC_Sim.propagate
//   move_randomly (pyvibdmc.py:540-547) -> potential (:786-793) -> birth_or_death (:380-431) -> calc_vref (:651-661)
// with the np.repeat gather of birth_or_death (:415-431) moved to the START of the next step.
//
// Why.  k_step_discrete (pvd_step.cuh) compacts inside the step: every 32-walker tile needs the number of copies all
// earlier tiles produced (a chained scan with look-back), so every warp waits for the slowest of the ~60 warps that hold
// the tiles just before its own.  Warps of an SM sub-partition are not scheduled fairly, so there is always a slow one:
// ncu (profiles/r02_step_kernel.md) shows 3.7 polls of the look-back per tile and ~30 % of all warp time spent waiting
// on status words or the ticket counter -- and MORE resident warps make the step slower, not faster.
//
// How.  A step writes what it computes where it computed it -- moved coordinates, V and the copy count of walker slot o
// at slot o -- plus one number per tile (the tile's total).  No warp ever waits for another warp.  The ordered prefix
// the compaction needs has two levels: tiles are dealt to CTAs in contiguous chunks (the warps of a CTA share the chunk
// through a shared-memory ticket), each CTA leaves the inclusive prefix of its tile totals (tincl) when it is done, and
// the last CTA of the step, which already combines the per-CTA records into Vref, scans the per-CTA totals (cbase).
// The NEXT step pulls: a CTA's output tiles are contiguous and so are their sources (the map is monotone), so the CTA
// finds the first source tile of its chunk once (chunk in cbase by bisection in shared memory, tile in a 32-wide window
// of tincl), loads the copy counts from there on, scans them block-wide and writes the source slot of every output
// slot of the chunk into shared memory (gather_build_map): two round trips to L2 per CTA and step.  A tile then reads
// its 32 source slots from shared memory and gathers its walkers -- contiguous runs, because the map is monotone.
// (Doing the pull per tile -- three dependent round trips in front of every tile's loads -- cost 25 us of a 109 us step.)
// The arithmetic per walker, the random-number addressing (compacted slot, step) and the np.repeat order are those of
// k_step_discrete: trajectories are bit-identical (tests/test_gpu_gather.py).  A segment of steps ends with
// k_gather_materialise, which leaves the ensemble compacted in the other buffer exactly as k_step_discrete would have.
// Tried on this kernel and rejected (B200, 1e6 walkers, us per step against 90.2): cp.async double-buffering of the next
// tile's coordinates (100.8), three CTAs per SM at 80 registers with 16 bytes of spills (96.7), odd warps started half a
// tile late (90.8 .. 94.5), and an "early start" protocol in which step k+1 starts on a software flag as soon as step k's
// prefix is published and only its copy-count stage waits for Vref(k) (96.0: the two GPU-scope fences and the polling
// cost more than the hardware's dependent-launch hand-over saves; it won 2 us at 200 000 walkers).  For several GPUs a variant in which a
// step only POSTS its sums and warp 0 of the next kernel's CTA 0 collects and finalises while every other warp already works on its first
// tile (verdict awaited at the copy-count stage, first tile's coordinates parked in shared memory) did overlap the exchange (2 GPUs:
// 128.6 -> 119.1 us with the same binary) but the extra state made the MULTI instantiation spill at its 128-register limit, which cost
// more than the overlap won (105-107 us for the posted-and-collected-in-place exchange that is in the tree).
#pragma once
#include "pvd_step.cuh"

constexpr int PVD_GATHER_MAX_TPC = 4096;       // tiles per CTA chunk the shared-memory arrays hold (131 072 walkers per CTA)
constexpr int PVD_GATHER_MAX_GRID = 1024;

// what a buffer in deferred form carries besides x / V / who: written by the step that wrote the buffer
struct GatherMeta {
    long long n_slots;      // slots holding (moved walker, copy count); 0: the buffer is compacted
    int tpc;                // tiles per CTA chunk of the step that wrote it
    int nchunks;            // CTAs of that step
};

struct GatherArgs {
    const int *cnt_in;      // copy count of every slot of the input buffer
    int *cnt_out;
    const int *tincl_in;    // per tile: inclusive prefix of the tile totals inside the tile's chunk
    int *tincl_out;
    const int *cbase_in;    // [nchunks + 1]: exclusive prefix of the chunk totals
    int *cbase_out;
    const GatherMeta *meta_in;
    GatherMeta *meta_out;
    int deferred_in;        // 0: the input buffer is compacted (first step of a segment)
    long long *seg_step0;   // the first step of a segment leaves [0] the step counter it starts from, [1] the error bits it found
    int stagger_ns;         // odd warps start their first tile this much later (warps of a scheduler out of phase: see k_step_gather)
};

// ---------------------------------------------------------------- the pull, once per CTA and pass
// The CTA's output tiles are contiguous, and so are their sources (the map is monotone): instead of three dependent
// round trips to L2 per TILE (chunk -> tincl window -> copy counts; measured 25 us of a 109 us step at 1e6 walkers), the
// CTA builds the source map of up to PVD_GATHER_SUBT tiles in shared memory with two round trips per PASS: warp 0 finds
// the first source tile, every thread loads 16 consecutive copy counts, one block-wide scan, and every source writes its
// slot number into the output slots it feeds.  A tile's pull is then one shared-memory read per lane.
constexpr int PVD_GATHER_SUBT = 256;            // tiles per pass: the map holds 8192 output slots (32 KB)
constexpr int PVD_GATHER_PER_THREAD = 16;       // copy counts per thread and round (4096 source slots per round of a 256-thread CTA)

// all threads of the CTA; o_lo: first output slot of the pass (a multiple of 32), n_map: its output slots
__device__ inline void gather_build_map(const GatherArgs &g, const int *s_cbase, int nchunks, int tpc, int n_slots,
                                        int o_lo, int n_map, int *s_map, int *s_tmp)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid == 0) {
        int lo = 0, hi = nchunks - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_cbase[mid] <= o_lo) lo = mid; else hi = mid - 1;
        }
        const int cb = s_cbase[lo];
        const int rel = o_lo - cb;
        const int ntiles_in = (n_slots + PVD_TILE - 1) / PVD_TILE;
        const int tb = lo * tpc;
        const int nt = min(tpc, ntiles_in - tb);
        // a chunk's copies are spread almost evenly over its tiles: the tile sought is near rel * nt / (chunk total)
        const int ctot = s_cbase[lo + 1] - cb;
        int ws = (int)(((long long)rel * nt) / (ctot > 0 ? ctot : 1)) - 15;
        ws = max(min(ws, nt - 32), 0);
        int t0, prev;
        while (true) {
            const int idx = ws + lane;
            const int val = idx < nt ? __ldcg(&g.tincl_in[tb + idx]) : 0x7fffffff;
            const unsigned m = __ballot_sync(0xffffffffu, val > rel);
            if (m == 0u) { ws += 32; continue; }
            const int f = __ffs((int)m) - 1;
            if (f == 0 && ws > 0) { ws = max(ws - 31, 0); continue; }
            t0 = ws + f;
            prev = __shfl_sync(0xffffffffu, val, f > 0 ? f - 1 : 0);
            if (f == 0) prev = 0;
            break;
        }
        if (lane == 0) { s_tmp[0] = (tb + t0) * PVD_TILE; s_tmp[1] = cb + prev - o_lo; }
    }
    __syncthreads();
    int s_cur = s_tmp[0], pos = s_tmp[1];                      // first source slot; output slot (relative) of its first copy: <= 0
    while (true) {
        const int first = s_cur + PVD_GATHER_PER_THREAD * (int)threadIdx.x;
        int c[PVD_GATHER_PER_THREAD];
#pragma unroll
        for (int q = 0; q < PVD_GATHER_PER_THREAD / 4; ++q) {
            int4 v = make_int4(0, 0, 0, 0);
            if (first + 4 * q < n_slots) v = __ldcg(reinterpret_cast<const int4 *>(g.cnt_in + first) + q);   // (the buffer is a multiple of 32 long)
            c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
        }
        int tsum = 0;
#pragma unroll
        for (int j = 0; j < PVD_GATHER_PER_THREAD; ++j) {
            if (first + j >= n_slots) c[j] = 0;               // stale counts beyond the ensemble
            tsum += c[j];
        }
        const int wincl = warp_incl_scan(tsum);
        if (lane == 31) s_tmp[2 + wid] = wincl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < PVD_WARPS; ++w) {
            const int t = s_tmp[2 + w];
            if (w < wid) wbase += t;
            total += t;
        }
        int e = pos + wbase + wincl - tsum;
#pragma unroll
        for (int j = 0; j < PVD_GATHER_PER_THREAD; ++j)
            for (int m = 0; m < c[j]; ++m, ++e)
                if ((unsigned)e < (unsigned)n_map) s_map[e] = first + j;
        pos += total;
        s_cur += PVD_GATHER_PER_THREAD * PVD_CTA;
        __syncthreads();                                      // s_tmp[2..] is re-used; after the last round: the map is complete
        if (pos >= n_map || s_cur >= n_slots) break;
    }
}

// ---------------------------------------------------------------- the pull for ONE output tile (used by the materialisation, a light kernel at full occupancy)
// All lanes call it with the same arguments.  Returns the source slot of output slot o0 + lane (-1 beyond the
// ensemble).  s_src: 32 ints private to the warp.
__device__ __forceinline__ int gather_sources(const GatherArgs &g, const int *s_cbase, int nchunks, int tpc, int n_slots,
                                              int o0, int n_out, int *s_src)
{
    const int lane = threadIdx.x & 31;
    // chunk: the last b with cbase[b] <= o0 (empty chunks repeat their successor's base and are skipped by "last")
    int lo = 0, hi = nchunks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_cbase[mid] <= o0) lo = mid; else hi = mid - 1;
    }
    const int cb = s_cbase[lo];
    const int rel = o0 - cb;
    const int ntiles_in = (n_slots + PVD_TILE - 1) / PVD_TILE;
    const int tb = lo * tpc;
    const int nt = min(tpc, ntiles_in - tb);
    // first tile of the chunk whose inclusive prefix exceeds rel: a tile holds about 32 copies, so it is near rel / 32
    int ws = min(rel >> 5, nt - 1) - 15;
    ws = max(min(ws, nt - 32), 0);
    int t0, prev;
    while (true) {
        const int idx = ws + lane;
        const int val = idx < nt ? __ldcg(&g.tincl_in[tb + idx]) : 0x7fffffff;
        const unsigned m = __ballot_sync(0xffffffffu, val > rel);
        if (m == 0u) { ws += 32; continue; }                 // (the chunk's last prefix exceeds rel: the window stays inside)
        const int f = __ffs((int)m) - 1;
        if (f == 0 && ws > 0) { ws = max(ws - 31, 0); continue; }
        t0 = ws + f;
        prev = __shfl_sync(0xffffffffu, val, f > 0 ? f - 1 : 0);
        if (f == 0) prev = 0;
        break;
    }
    // expand the copy counts from source tile t0 on, 64 slots per round, into the tile's 32 output slots
    int pos = cb + prev - o0;                                // output slot (relative) of the first copy of the first source: <= 0
    int k = (tb + t0) * PVD_TILE + lane;
    const int need = min(PVD_TILE, n_out - o0);
    s_src[lane] = -1;
    __syncwarp();
    while (true) {
        const int c0 = k < n_slots ? __ldcg(&g.cnt_in[k]) : 0;
        const int c1 = k + PVD_TILE < n_slots ? __ldcg(&g.cnt_in[k + PVD_TILE]) : 0;
        const int i0 = warp_incl_scan(c0);
        const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
        const int i1 = warp_incl_scan(c1);
        const int tot1 = __shfl_sync(0xffffffffu, i1, 31);
        int e = pos + i0 - c0;
        for (int m = 0; m < c0; ++m, ++e)
            if ((unsigned)e < (unsigned)PVD_TILE) s_src[e] = k;
        e = pos + tot0 + i1 - c1;
        for (int m = 0; m < c1; ++m, ++e)
            if ((unsigned)e < (unsigned)PVD_TILE) s_src[e] = k + PVD_TILE;
        pos += tot0 + tot1;
        k += 2 * PVD_TILE;
        if (pos >= need || k - lane >= n_slots) break;
    }
    __syncwarp();
    const int src = s_src[lane];
    __syncwarp();
    return src;
}

// ---------------------------------------------------------------- end of a gather step
// Chunk bookkeeping on top of cta_finish_step: this CTA's inclusive tile prefixes, and (last CTA) the scan of the per-CTA
// totals.  Called by every thread of every CTA after its tile loop; s_ttot holds the totals of the CTA's tiles.
template <bool MULTI>
__device__ inline void gather_finish_step(const StepArgs &a, const GatherArgs &g, const LaneAcc &acc, long long n, int ntiles, int tpc,
                                          const int *s_ttot, int *s_scan)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tb = blockIdx.x * tpc;
    const int nt = max(min(tpc, ntiles - tb), 0);
    __syncthreads();                                         // every warp of the CTA has written its tile totals
    if (wid == 0) {
        int running = 0;
        for (int base = 0; base < nt; base += 32) {
            const int v = base + lane < nt ? s_ttot[base + lane] : 0;
            const int incl = warp_incl_scan(v) + running;
            if (base + lane < nt) g.tincl_out[tb + base + lane] = incl;
            running = __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    // sums, CTA record, last CTA -> Vref / population / log record (pvd_step.cuh); the last CTA also leaves cbase
    __shared__ unsigned s_is_last;
    {
        __shared__ WarpPartial s_part[PVD_WARPS];
        const WarpPartial p = acc_warp_reduce(acc);
        if (lane == 0) s_part[wid] = p;
        __syncthreads();
        if (threadIdx.x == 0) {
            LaneAcc r;
            for (int w = 0; w < PVD_WARPS; ++w) acc_merge(r, s_part[w]);
            WarpPartial q;
            q.cv = r.cv; q.v = r.v; q.cw = r.cw; q.c = r.c; q.births = r.births; q.deaths = r.deaths; q.n_in = r.n_in; q.n_acc = r.n_acc;
            q.vmin = r.vmin; q.vmax = r.vmax; q.wmin = r.wmin; q.wmax = r.wmax;
            a.part[blockIdx.x] = q;
            __threadfence();                                 // the record and (warp 0 above, same CTA) tincl before the arrival
            const unsigned d = atomicAdd(&a.st[a.parity].done, 1u);
            s_is_last = (d == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (!s_is_last) return;
        __threadfence();
        LaneAcc r;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += PVD_CTA) {
            const WarpPartial *q = &a.part[b];
            WarpPartial t;
            t.cv.hi = __ldcg(&q->cv.hi); t.cv.lo = __ldcg(&q->cv.lo); t.v.hi = __ldcg(&q->v.hi); t.v.lo = __ldcg(&q->v.lo);
            t.cw.hi = 0ll; t.cw.lo = 0ull;
            t.c = __ldcg(&q->c); t.births = __ldcg(&q->births); t.deaths = __ldcg(&q->deaths);
            t.n_in = __ldcg(&q->n_in); t.n_acc = __ldcg(&q->n_acc);
            t.vmin = __ldcg(&q->vmin); t.vmax = __ldcg(&q->vmax); t.wmin = INFINITY; t.wmax = -INFINITY;
            s_scan[b] = (int)t.c;                            // the chunk's total (exact in double)
            acc_merge(r, t);
        }
        const WarpPartial pw = acc_warp_reduce(r);
        if (lane == 0) s_part[wid] = pw;
        __syncthreads();
        if (wid == 1) {
            // exclusive prefix of the chunk totals
            int running = 0;
            for (int base = 0; base < (int)gridDim.x; base += 32) {
                const int v = base + lane < (int)gridDim.x ? s_scan[base + lane] : 0;
                const int incl = warp_incl_scan(v) + running;
                if (base + lane < (int)gridDim.x) g.cbase_out[base + lane] = incl - v;
                running = __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) {
                g.cbase_out[gridDim.x] = running;
                GatherMeta mo;
                mo.n_slots = n;
                mo.tpc = tpc;
                mo.nchunks = (int)gridDim.x;
                *g.meta_out = mo;
            }
        }
        if (threadIdx.x == 0) {
            LaneAcc f;
            for (int w = 0; w < PVD_WARPS; ++w) acc_merge(f, s_part[w]);
            double *s = a.sums;
            for (int k = 0; k < PVD_SUM_EXT + 4 * a.world; ++k) s[k] = 0.0;
            sum_put(s, PVD_SUM_CV, f.cv);
            sum_put_double(s, PVD_SUM_C, f.c);
            sum_put(s, PVD_SUM_V, f.v);
            s[PVD_SUM_BIRTHS] = f.c - (f.n_in - f.deaths);       // sum max(c - 1, 0) = sum c - #(c >= 1)
            s[PVD_SUM_DEATHS] = f.deaths; s[PVD_SUM_NIN] = f.n_in; s[PVD_SUM_NACC] = f.n_acc;
            unsigned e = *a.err_accum;
            if (f.c > (double)a.cap) e |= PVD_ERR_CAPACITY;      // the next step has one slot per copy
            s[PVD_SUM_ERR] = err_encode(e);
            double *ex = s + PVD_SUM_EXT + 4 * a.rank;
            ex[0] = f.vmin; ex[1] = f.vmax; ex[2] = f.wmin; ex[3] = f.wmax;
            a.st[a.parity ^ 1].n = (long long)f.c;
            if (a.world == 1) finalize_from_sums(a, false);
        }
        if constexpr (MULTI) if (a.world > 1 && a.mbox[0]) {
            mailbox_send(a, a.st[a.parity].step);
            __syncthreads();
            if (threadIdx.x < 32) mailbox_collect_and_finalize(a, false);
        }
    }
}

// ---------------------------------------------------------------- the step
// MULTI: the instantiation that contains the NVLink exchange (several GPUs)
template <class POT, int RNG, int MINB, bool MULTI>
__global__ void __launch_bounds__(PVD_CTA, MINB) k_step_gather(const StepArgs a, const GatherArgs g)
{
    constexpr int NC = POT::NC;
    extern __shared__ __align__(16) unsigned char s_dyn[];   // source map of a pass, [tpc] tile totals of this CTA, [grid + 1] chunk bases / scan scratch
    __shared__ int s_tmp[2 + PVD_WARPS];
    __shared__ ulonglong2 s_acc_cv[PVD_CTA], s_acc_v[PVD_CTA];    // per-thread exact sums {lo, hi}: sum count*V, sum V
    __shared__ unsigned s_next;
    if (threadIdx.x == 0) s_next = 0u;
    if constexpr (RNG == PVD_RNG_ZIGGURAT) zig_stage();           // constant table: independent of the previous step
    else __syncthreads();
    pdl_wait();
    if (!g.deferred_in && blockIdx.x == 0 && threadIdx.x == 0) { g.seg_step0[0] = a.st[a.parity].step; g.seg_step0[1] = (long long)a.st[a.parity].err; }
    if (!step_prologue(a)) return;
    const DevState *sip = &a.st[a.parity];
    const int n = (int)sip->n;
    const long long step = sip->step;
    const double vref = sip->vref;
    const int ntiles = (n + PVD_TILE - 1) / PVD_TILE;
    const int tpc = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    int *s_map = reinterpret_cast<int *>(s_dyn);                 // [min(tpc, SUBT) * 32] source slot of every output slot of the pass
    int *s_ttot = s_map + min(tpc, PVD_GATHER_SUBT) * PVD_TILE;
    int *s_cbase = s_ttot + tpc;
    const bool dw = sip->dw_active != 0;
    const double n0 = (double)a.n0;
    const double w_limit = (n0 + n0 * 0.5) + 1.0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool branch_now = branch_this_step(a.do_branch, step);
    int in_slots = n, in_tpc = 0, in_chunks = 0;
    if (g.deferred_in) {
        in_slots = (int)__ldcg(&g.meta_in->n_slots);
        in_tpc = __ldcg(&g.meta_in->tpc);
        in_chunks = __ldcg(&g.meta_in->nchunks);
        for (int b = threadIdx.x; b <= in_chunks; b += PVD_CTA) s_cbase[b] = __ldcg(&g.cbase_in[b]);
        __syncthreads();
    }
    const int tb = blockIdx.x * tpc;
    const int nt = max(min(tpc, ntiles - tb), 0);
    double vmin = INFINITY, vmax = -INFINITY;
    int csum = 0, deaths = 0;                                   // warp-uniform
    if (g.stagger_ns > 0 && (wid & 1)) __nanosleep((unsigned)g.stagger_ns);
    s_acc_cv[threadIdx.x] = make_ulonglong2(0ull, 0ull);
    s_acc_v[threadIdx.x] = make_ulonglong2(0ull, 0ull);

    for (int p0 = 0; p0 < nt; p0 += PVD_GATHER_SUBT) {
        const int ntp = min(PVD_GATHER_SUBT, nt - p0);          // tiles of this pass
        const int o_lo = (tb + p0) * PVD_TILE;
        if (p0 > 0) {
            __syncthreads();                                    // every warp is done with the previous pass (map, ticket)
            if (threadIdx.x == 0) s_next = 0u;
        }
        if (g.deferred_in) gather_build_map(g, s_cbase, in_chunks, in_tpc, in_slots, o_lo, min(ntp * PVD_TILE, n - o_lo), s_map, s_tmp);
        else if (p0 > 0) __syncthreads();
    while (true) {
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(&s_next, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntp) break;
        const int tile = tb + p0 + t;
        const int o = tile * PVD_TILE + lane;
        const bool active = o < n;
        int src = o;
#ifndef PVD_EXP_NO_PULL
        if (g.deferred_in && active) src = s_map[t * PVD_TILE + lane];
#endif
        double x[NC], v = 0.0;
        int cnt = 0;
        {
            // (prefetching the next tile's coordinates with cp.async while this one is evaluated was measured: +6 us per step)
            const double *px = a.xin + src;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                x[c] = 1.0 + c;                                                   // idle lanes: harmless geometry
                if (active) x[c] = __ldcs(px);
                px += a.cap;
            }
        }
        if (active) {
            if (a.inj_disp) {
#pragma unroll
                for (int c = 0; c < NC; ++c) x[c] = x[c] + a.inj_disp[c * a.cap + o];
            } else {
                double z[NC];
#ifdef PVD_EXP_NO_RNG
#pragma unroll
                for (int c = 0; c < NC; ++c) z[c] = 1e-3 * (double)((o + c + (int)step) & 7) - 3.5e-3;
#else
                walker_normals<NC, RNG>(a.seed, (long long)o, step, z);
#endif
#pragma unroll
                for (int c = 0; c < NC; ++c) x[c] = __dadd_rn(x[c], __dmul_rn(a.sigc[c], z[c]));
            }
            double *po = a.xout + o;
#pragma unroll
            for (int c = 0; c < NC; ++c) { *po = x[c]; po += a.cap; }
            if (dw) a.who_out[o] = a.who_in[src];
        }
#ifdef PVD_EXP_NO_PES
        v = 0.0211 + 1e-3 * (x[0] - 1.828) + 1e-3 * (x[4] - 1.77) + 1e-4 * (x[1] + x[2] + x[3] + x[5] + x[6] + x[7] + x[8]);
#else
        v = active ? POT::eval(x, a.pot) : 0.0;
#endif
        bool bad = false;
        if (active) {
            if (branch_now) {
                double u;
                if (a.inj_u) u = a.inj_u[o];
                else { const uint4 r = pvd_draw(a.seed, (long long)o, step, PVD_STREAM_BRANCH, 0u); u = u53(r.x, r.y); }
                cnt = discrete_count(v, vref, a.dt, u, w_limit, bad);
            } else cnt = 1;
            a.vout[o] = v;
            g.cnt_out[o] = cnt;
            // running sums of this lane: the exact (fixed-point) ones live in shared memory, not in registers -- the
            // kernel is bounded by how many warps an SM holds
            const Fx128 fv = fx_from_double_fast(v);       // (same value as fx_from_double for every |V| >= 2^-27 Hartree: pvd_run.cuh)
            ulonglong2 sv = s_acc_v[threadIdx.x];
            const Fx128 nv = fx_add(Fx128{(long long)sv.y, sv.x}, fv);
            s_acc_v[threadIdx.x] = make_ulonglong2(nv.lo, (unsigned long long)nv.hi);
            if (cnt > 0) {
                ulonglong2 sc = s_acc_cv[threadIdx.x];
                const Fx128 nc = fx_add(Fx128{(long long)sc.y, sc.x}, cnt == 1 ? fv : fx_mul_small(fv, cnt));
                s_acc_cv[threadIdx.x] = make_ulonglong2(nc.lo, (unsigned long long)nc.hi);
            }
            vmin = fmin(vmin, v); vmax = fmax(vmax, v);
        }
        if (bad) atomicOr(a.err_accum, PVD_ERR_WEIGHT);
        const int tot = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) s_ttot[p0 + t] = tot;
        csum += tot;
        deaths += __popc(__ballot_sync(0xffffffffu, active && cnt == 0));
    }
    }
    LaneAcc acc;
    {
        const ulonglong2 sc = s_acc_cv[threadIdx.x], sv = s_acc_v[threadIdx.x];
        acc.cv = Fx128{(long long)sc.y, sc.x};
        acc.v = Fx128{(long long)sv.y, sv.x};
        acc.vmin = vmin; acc.vmax = vmax;
        if (threadIdx.x == 0) {
            // whole-CTA figures (exact integers): walkers in = the chunk's slots, births = sum max(c - 1, 0) = sum c - #(c >= 1)
            const long long n_in = max(min((long long)n, (long long)(tb + nt) * PVD_TILE) - (long long)tb * PVD_TILE, 0ll);
            acc.n_in = (double)n_in; acc.n_acc = (double)n_in;
        }
        if (lane == 0) { acc.c = (double)csum; acc.deaths = (double)deaths; }
    }
    gather_finish_step<MULTI>(a, g, acc, (long long)n, ntiles, tpc, s_ttot, s_cbase);
}

// ---------------------------------------------------------------- end of a segment: deferred form -> compacted ensemble
// x / V / who of buffer `in` (deferred form) -> buffer `out`, compacted in np.repeat order: what k_step_discrete leaves.
// One launch; the run may have died inside the segment (the failing step's input is then the valid ensemble: reference
// semantics, pyvibdmc.py:397-413).  k_first/k_count: the host's view of the segment; the device state says how far it got.
struct MaterialiseArgs {
    double *x[2];
    double *v[2];
    int *who[2];
    const int *cnt[2];
    const int *tincl[2];
    const int *cbase[2];
    const GatherMeta *meta[2];
    DevState *st;
    const long long *seg_step0;
    long long cap;
    int nc, parity_end, buf0;
};

// (a version that builds shared-memory maps of 64 tiles per CTA like the step kernel was measured: 68 us instead of 54 us at 1e6 walkers)
__global__ void __launch_bounds__(PVD_CTA) k_gather_materialise(const MaterialiseArgs m)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ int s_src[PVD_WARPS][32];
    int *s_cbase = reinterpret_cast<int *>(s_dyn);
    pdl_wait();
    DevState *st = &m.st[m.parity_end];
    const long long done = __ldcg(&st->step) - __ldcg(m.seg_step0);     // successful steps of the segment
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (__ldcg(m.seg_step0 + 1) != 0) return;               // dead before the segment began: whoever saw it die left the ensemble in st.buf
    if (done <= 0) {                                        // died in its first step: the compacted input is still the ensemble
        if (blockIdx.x == 0 && threadIdx.x == 0) { m.st[0].buf = m.buf0; m.st[1].buf = m.buf0; }
        return;
    }
    const int in = m.buf0 ^ (int)(done & 1), out = in ^ 1;
    const int n = (int)__ldcg(&st->n);
    GatherArgs g{};
    g.cnt_in = m.cnt[in];
    g.tincl_in = m.tincl[in];
    const int in_slots = (int)__ldcg(&m.meta[in]->n_slots), in_tpc = __ldcg(&m.meta[in]->tpc), in_chunks = __ldcg(&m.meta[in]->nchunks);
    for (int b = threadIdx.x; b <= in_chunks; b += PVD_CTA) s_cbase[b] = __ldcg(&m.cbase[in][b]);
    __syncthreads();
    const bool dw = __ldcg(&st->dw_active) != 0;
    const int ntiles = (n + PVD_TILE - 1) / PVD_TILE;
    const int nwarps = gridDim.x * PVD_WARPS;
    for (int tile = blockIdx.x * PVD_WARPS + wid; tile < ntiles; tile += nwarps) {
        const int o = tile * PVD_TILE + lane;
        const int src = gather_sources(g, s_cbase, in_chunks, in_tpc, in_slots, tile * PVD_TILE, n, s_src[wid]);
        if (o < n) {
            const double *px = m.x[in] + src;
            double *po = m.x[out] + o;
            for (int c = 0; c < m.nc; ++c) { *po = __ldcs(px); px += m.cap; po += m.cap; }
            m.v[out][o] = __ldcs(&m.v[in][src]);
            if (dw) m.who[out][o] = m.who[in][src];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { m.st[0].buf = out; m.st[1].buf = out; }
}

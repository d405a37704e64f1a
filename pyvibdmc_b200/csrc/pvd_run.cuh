// The resident discrete-weighting loop: MANY time steps of DMC_Sim.propagate (pyvibdmc.py:701-876) in ONE kernel launch.
//   move_randomly (:540-547) -> potential (:786-793) -> birth_or_death (:380-431) -> calc_vref (:651-661), step after step
// Same per-tile arithmetic as k_step_discrete (pvd_step.cuh), hence bit-identical trajectories; what changes is how the
// steps are chained.  A time step has exactly one global dependency, Vref (the copy counts of step s+1 need Vref(s),
// which needs every energy of step s).  Everything else only needs the walkers it reads, so instead of a kernel boundary
// per step:
//   * every output tile (32 consecutive slots of the compacted ensemble) has a counter of the slots written so far
//     (`ready`); tile j of step s+1 starts -- load, displacement, potential -- as soon as output tile j of step s is
//     complete, whatever the rest of step s is doing;
//   * only the copy-count stage waits for Vref(s): the warp that completes the last CTA record of step s finalises it
//     (exact sums -> Vref, population, guards, log record; with several GPUs the NVLink mailbox exchange) and publishes a
//     sequence number; by then most warps are in the middle of their first tile of step s+1;
//   * per-step sums go warp -> CTA (shared-memory integer atomics) -> global integer atomics: order independent, so Vref
//     is bit-reproducible, and nobody re-reads hundreds of per-CTA records on the critical path.
// No kernel launch, no grid-wide barrier, no pipeline drain between steps: the fixed cost per step of the one-launch-
// per-step design (launch hand-over, fill/drain of a ~7 us tile, last-CTA reduction: ~14 us) disappears, and with several
// GPUs the exchange of step s overlaps the move + potential of step s+1.
// Hazards (why two parities of everything suffice): a scatter of step s+1 needs Vref(s), i.e. every tile of step s has
// been loaded and evaluated, so the ping-pong buffers are never overwritten early; ready/status words of parity p are
// consumed (and the ready word cleared) before the copy-count stage of step s+1 completes, which precedes Vref(s+1) and
// therefore every write of step s+2 to the same parity.
// Launched cooperatively (all CTAs co-resident: warps poll flags written by other CTAs).
#pragma once
#include "pvd_step.cuh"

// Cache policy of the walker traffic.  Loads: streaming (evict-first) like the one-launch-per-step kernel -- safe in a
// resident kernel because every tile's loads follow an acquire of its counter, which invalidates the SM's L1 (CCTL.IVALL),
// so a line cached before the data was rewritten cannot be hit.  Stores: write-back (they reach L2 in any case).
#ifndef PVD_RUN_LD
#define PVD_RUN_LD(p) __ldcs(p)
#endif
#ifndef PVD_RUN_ST
#define PVD_RUN_ST(p, v) (*(p) = (v))
#endif
constexpr int PVD_RUN_NACC = 16;
// accumulator slots (64-bit integers, per parity): 0-3 sum count*V limbs, 4-7 sum V limbs, 8 sum count, 9 deaths,
// 10 ordered key of min V (atomicMin), 11 ordered key of max V (atomicMax)
struct RunCtl {
    // what the copy-count stage of a step needs, published by the finalisation of the previous step as ONE 16-byte word
    // (a single vector store / load: no fence on the critical path):
    //   pub[0] = (steps finalised so far, low 32 bits) << 32 | population of this shard (bit 31: the run is dead),  pub[1] = Vref
    unsigned long long pub[2];
    unsigned long long done_seq;        // steps whose bookkeeping (state copy, log record, re-armed counters) is complete and fenced
    unsigned long long pad0[13];
    unsigned tk[2][32];                 // ticket counter per step parity, one per 128-byte line
    unsigned arrive[2][32];             // CTAs that have delivered their record of the step of this parity
    unsigned long long acc[2][PVD_RUN_NACC];
};

// identity of every accumulator; seq = 0
__global__ void k_run_ctl_init(RunCtl *c)
{
    const int t = threadIdx.x;
    if (t == 0) { c->pub[0] = 0ull; c->pub[1] = 0ull; c->done_seq = 0ull; }
    if (t < 64) { (&c->tk[0][0])[t] = 0u; (&c->arrive[0][0])[t] = 0u; }
    if (t < 2 * PVD_RUN_NACC) (&c->acc[0][0])[t] = (t % PVD_RUN_NACC) == 10 ? ~0ull : 0ull;
}

struct RunArgs {
    RunCtl *ctl;
    unsigned *ready;                    // [2][ntiles_cap]: slots of output tile j written so far, by parity of the WRITING step
    long long ntiles_cap;
    long long nsteps;
    unsigned long long seq0;            // steps finalised by earlier resident launches (the sequence number this launch starts from)
    int dynamic;                        // 1: tiles from a ticket counter (A/B); 0: static interleaved assignment
    int single;                         // 1: one time step per launch (programmatic dependent launch): the kernel boundary orders the
                                        //    scattered walkers before the next step's loads, so nothing is fenced or announced
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Every poll loop of the resident kernel gives up eventually (a peer GPU died, or a bug): the run is marked dead with
// PVD_ERR_COMM instead of hanging the device.  The limit is twice the mailbox time-out plus a minute, so that a rank waiting
// for a slow peer is reported by the exchange itself, not by its bystanders.
struct SpinGuard {
    long long limit;
    long long t0 = 0;
    unsigned n = 0;
    __device__ __forceinline__ explicit SpinGuard(long long lim) : limit(lim) {}
    __device__ __forceinline__ bool expired()
    {
        if (++n < 1024u) return false;
        if (t0 == 0) { t0 = clock64(); return false; }
        return clock64() - t0 > limit;
    }
};

// monotone map double -> uint64 (so that integer atomicMin/Max order doubles) and back
__device__ __forceinline__ unsigned long long dkey(double d)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k)
{
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
}

// truncating conversion to 128-bit fixed point (quantum 2^-80) in a handful of instructions: the integer part in units
// of 2^-16 by floor + cvt, the remainder (exact by one FMA) in units of 2^-80.  Same value as fx_from_double for every
// |x| >= 2^-27 (no bits below the quantum); below that it rounds down instead of towards zero.
__device__ __forceinline__ Fx128 fx_from_double_fast(double x)
{
    x = fmin(fmax(x, -7.0e13), 7.0e13);                 // NaN -> -7e13 (such a walker raises PVD_ERR_WEIGHT anyway)
    const double hd = floor(x * 65536.0);
    const double rem = fma(hd, -0x1.0p-16, x);          // exact, in [0, 2^-16)
    Fx128 r;
    r.hi = (long long)hd;
    r.lo = (unsigned long long)(rem * 0x1.0p80);
    return r;
}

// 128-bit sum split into four 32-bit limbs, each added to its own 64-bit accumulator (no carries between atomics)
__device__ __forceinline__ void limb_add_shared(unsigned long long *acc4, Fx128 x)
{
    atomicAdd(&acc4[0], x.lo & 0xffffffffull);
    atomicAdd(&acc4[1], x.lo >> 32);
    atomicAdd(&acc4[2], (unsigned long long)x.hi & 0xffffffffull);
    atomicAdd(&acc4[3], (unsigned long long)(x.hi >> 32));          // sign-extended: wraps correctly modulo 2^64
}
__device__ __forceinline__ Fx128 limb_combine(unsigned long long l0, unsigned long long l1, unsigned long long l2, unsigned long long l3)
{
    // value = l0 + l1 2^32 + l2 2^64 + l3 2^96 (l3 signed), modulo 2^128
    Fx128 r;
    r.lo = l0 + (l1 << 32);
    const unsigned long long c0 = r.lo < l0 ? 1ull : 0ull;
    r.hi = (long long)((l1 >> 32) + l2 + (l3 << 32) + c0);
    return r;
}

// A fence at GPU scope costs ~2 us whatever the warp has in flight, so written slots are announced in batches: one fence
// per PVD_RUN_RDY tiles (and at the end of the step).  A tile is consumed one whole time step after it was produced, so
// the delay never makes a consumer wait, except where it matters -- small ensembles, where every tile is a step's last.
constexpr int PVD_RUN_RDY = 8;

// One warp's shared memory: two stashes (the tile being computed, whose displaced coordinates are parked BEFORE the
// potential so that they do not occupy registers during it, and the previous tile, scattered afterwards) and the
// per-lane sums of the current step (kept out of the register file as well: the kernel is occupancy bound).
template <int NC>
struct RunStash {
    double x[NC + 1][32];     // components, then V
    int cnt[32], who[32];     // (the exclusive scan of cnt is recomputed at scatter time: shared memory is what limits occupancy)
};
template <int NC>
struct RunWarpMem {
    RunStash<NC> stash[2];
    unsigned long long cv_hi[32], cv_lo[32], v_hi[32], v_lo[32];
    double vmin[32], vmax[32];
    int rdy_base[PVD_RUN_RDY], rdy_total[PVD_RUN_RDY];          // slots written but not yet announced
};

__device__ __forceinline__ void red_relaxed_add_u32(unsigned *p, unsigned v)
{
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Scatter of a finished tile: look-back for the output offset, then the survivors go from the warp's stash to the other
// ping-pong buffer.  Returns the offset (the number of slots written is the tile's total); -1: the look-back timed out.
// Element offsets are 32-bit (the host checks NC * capacity < 2^31).
template <int NC>
__device__ __forceinline__ int run_scatter_stores(const StepArgs &a, const RunArgs &ra, const RunStash<NC> &stash, int ps, int q,
                                                  int tile, long long gstep, int total, bool dw)
{
    const int lane = threadIdx.x & 31;
    unsigned long long *status = a.status + (long long)ps * ra.ntiles_cap;
    bool lost = false;
    const int base = (int)resolve_prefix(status, (long long)tile, gstep, total, &lost);
    if (lost) { atomicOr(a.err_accum, PVD_ERR_COMM); return -1; }
    const int pc = stash.cnt[lane];
    const int o = base + warp_incl_scan(pc) - pc;
    if (pc > 0) {
        const unsigned cap = (unsigned)a.cap;
        if ((unsigned)(o + pc) > cap) atomicOr(a.err_accum, PVD_ERR_CAPACITY);
        else {
            double *xo = q ? const_cast<double *>(a.xin) : a.xout;
            double *vo = q ? const_cast<double *>(a.vin) : a.vout;
            const double pv = stash.x[NC][lane];
#pragma unroll 1
            for (int k = 0; k < pc; ++k) {
                unsigned e = (unsigned)(o + k);
                PVD_RUN_ST(&vo[e], pv);
                if (dw) {
                    int *wo = q ? const_cast<int *>(a.who_in) : a.who_out;
                    __stcg(&wo[e], stash.who[lane]);
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) { PVD_RUN_ST(&xo[e], stash.x[c][lane]); e += cap; }
            }
        }
    }
    return base;
}

// The slots [base, base + total) are written AND fenced: add them to the `ready` counter of every output tile they touch
// (the consumer, tile j of the next step, acquires the counter before it loads).
__device__ __forceinline__ void run_publish_ready(const RunArgs &ra, int ps, int base, int total)
{
    if (total <= 0 || base < 0) return;
    const int lane = threadIdx.x & 31;
    unsigned *ready_out = ra.ready + (long long)ps * ra.ntiles_cap;
    const int first = base >> 5, last = (base + total - 1) >> 5;
    for (int jt = first + lane; jt <= last; jt += 32) {
        const int lo = (jt << 5) > base ? (jt << 5) : base;
        const int hi = ((jt + 1) << 5) < base + total ? ((jt + 1) << 5) : base + total;
        if ((long long)jt < ra.ntiles_cap) red_relaxed_add_u32(&ready_out[jt], (unsigned)(hi - lo));
    }
}

#define PVD_RUN_DIE()                                                                                   \
    do {                                                                                                \
        atomicOr(a.err_accum, PVD_ERR_COMM);                                                            \
        atomicOr(&a.st[0].err, PVD_ERR_COMM);                                                           \
        atomicOr(&a.st[1].err, PVD_ERR_COMM);                                                           \
        return;                                                                                         \
    } while (0)

// MULTI: the instantiation that contains the NVLink exchange (several GPUs); the single-GPU one carries none of its code or registers
template <class POT, int RNG, int TPB, int MINB, bool MULTI>
__global__ void __launch_bounds__(TPB, MINB) k_run_discrete(const StepArgs a, const RunArgs ra)
{
    constexpr int NC = POT::NC;
    constexpr int NWARP = TPB / 32;
    extern __shared__ __align__(16) unsigned char s_dyn[];       // NWARP x RunWarpMem<NC> (more than the 48 KB a static array may take)
    RunWarpMem<NC> *s_warp = reinterpret_cast<RunWarpMem<NC> *>(s_dyn);
    __shared__ unsigned long long s_acc[2][PVD_RUN_NACC];
    __shared__ unsigned s_arrive[2];
    if (threadIdx.x < 2 * PVD_RUN_NACC) {
        const int k = threadIdx.x % PVD_RUN_NACC;
        s_acc[threadIdx.x / PVD_RUN_NACC][k] = k == 10 ? ~0ull : 0ull;
    }
    if (threadIdx.x < 2) s_arrive[threadIdx.x] = 0u;
    if constexpr (RNG == PVD_RNG_ZIGGURAT) zig_stage();
    else __syncthreads();
    if (ra.single) pdl_wait();                                  // everything above overlapped the previous step's tail

    const int lane = threadIdx.x & 31;
    const long long spin_limit = 2 * a.mbox_timeout_ticks + 120000000000ll;
    RunWarpMem<NC> &wm = s_warp[threadIdx.x >> 5];
    RunCtl *ctl = ra.ctl;
    // the state copy this launch starts from is valid by stream order; step number and DW flag are launch constants
    // (a step that fails kills the run, and the host toggles dw_active only between launches)
    long long gstep0;
    bool dw;
    {
        const DevState *s0 = &a.st[a.parity];
        const unsigned e0 = __ldcg(&s0->err);
        const long long n_start = __ldcg(&s0->n);
        if (e0 || n_start <= 0) {
            // dead run / empty shard: nothing is touched, both state copies say so (the host may read either)
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                DevState d = load_state_cg(s0);
                if (!e0) d.err |= PVD_ERR_EMPTY;
                a.st[0] = d;
                a.st[1] = d;
            }
            return;
        }
        gstep0 = __ldcg(&s0->step);
        dw = __ldcg(&s0->dw_active) != 0;
    }

    for (int k = 0; k < (int)ra.nsteps; ++k) {
        const int ps = (a.parity + k) & 1;                      // state / ticket / status / accumulator parity of this step
        const int q = k & 1;                                    // ping-pong direction of this step
        unsigned *tk = &ctl->tk[ps][0];
        const DevState *sip = &a.st[ps];
        const long long gstep = gstep0 + k;
        const bool branch_now = branch_this_step(a.do_branch, gstep);
        bool st_ok = false;
        int n = 0;
        double vref = 0.0;
        // n and Vref of this step, published by the finalisation of the previous one: 0 not yet, 1 loaded, -1 the run died there
        const unsigned want_seq = (unsigned)(ra.seq0 + (unsigned long long)k);
        auto poll_state = [&]() -> int {
            unsigned long long w0, w1;
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(&ctl->pub[0]) : "memory");
            if ((int)((unsigned)(w0 >> 32) - want_seq) < 0) return 0;
            if ((unsigned)w0 & 0x80000000u) return -1;
            n = (int)(unsigned)w0;
            vref = __longlong_as_double((long long)w1);
            st_ok = true;
            return 1;
        };
        if (k == 0) {                                           // valid by stream order (checked above: alive, n > 0)
            n = (int)__ldcg(&sip->n);
            vref = __ldcg(&sip->vref);
            st_ok = true;
        }
        wm.cv_hi[lane] = 0ull; wm.cv_lo[lane] = 0ull; wm.v_hi[lane] = 0ull; wm.v_lo[lane] = 0ull;
        wm.vmin[lane] = INFINITY; wm.vmax[lane] = -INFINITY;
        int csum = 0, deaths = 0;                               // warp-uniform
        int pending = -1, pending_total = 0, cur = 0;
        int nrdy = 0;                                           // ranges of slots written but not yet announced (wm.rdy_*)
        // Which tiles this warp takes.  Static (default): tile = slot + j * (warps of the grid), slot = the warp's index with
        // the CTA index running fastest (neighbouring tiles on different SMs), rotated from step to step so that the warps
        // with one tile more than the others change.  Tiles are still taken in increasing order by co-resident warps, which
        // is all the look-back needs -- and no counter is hammered: one atomic per tile on a single address (31 250 + one
        // per warp, per step) runs at ~2.5 ns each, a floor of ~85 us per step at 1e6 walkers.  Dynamic (ra.dynamic): tickets.
        const unsigned nwarps_grid = gridDim.x * (unsigned)NWARP;
        unsigned t = 0;
        if (ra.dynamic) {
            if (lane == 0) t = atomicAdd(tk, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
        } else {
            t = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x + (unsigned)k * 977u) % nwarps_grid;
        }

        unsigned r_pre = 0u;                                    // the tile's counter as fetched ahead of time (see below)
        bool pre = false;
        while (true) {
            // ---- may tile t of this step start?  (k == 0: the whole input ensemble is there by stream order)
            bool have = false;
            if (k == 0) have = (long long)t * PVD_TILE < n;
            else {
                unsigned *ready_in = ra.ready + (long long)(ps ^ 1) * ra.ntiles_cap;      // written by the previous step
                SpinGuard guard(spin_limit);
                while (true) {
                    unsigned r;
                    if (pre) { r = r_pre; pre = false; }          // fetched while the previous tile was being scattered
                    else r = (long long)t < ra.ntiles_cap ? ld_acquire_u32(&ready_in[t]) : 0u;
                    if (r == (unsigned)PVD_TILE) { have = true; break; }
                    if (!st_ok && poll_state() < 0) return;      // the previous step failed: the run is dead, every warp leaves here
                    if (st_ok) {
                        const long long left = (long long)n - (long long)t * PVD_TILE;
                        if (left <= 0) break;                    // no such tile in this step
                        if ((long long)r == (left < PVD_TILE ? left : PVD_TILE)) { have = true; break; }
                    }
                    if (guard.expired()) PVD_RUN_DIE();
                    __nanosleep(guard.n < 8u ? 20 : 100);
                }
                if (have && lane == 0) ready_in[t] = 0u;         // consumed: the word is free for step s+2's writers
            }
            if (!have) break;

            // ---- move + potential (needs nothing but the walkers)
            const int i = (int)t * PVD_TILE + lane;
            RunStash<NC> &stash = wm.stash[cur];
            double v;
            unsigned tn = 0;
            {
                const bool active = st_ok ? i < n : true;       // a full tile that became ready before Vref is all walkers
                double x[NC];
                const double *xi = q ? a.xout : a.xin;
                unsigned e = (unsigned)i;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    x[c] = 1.0 + c;
                    if (active) x[c] = PVD_RUN_LD(&xi[e]);
                    e += (unsigned)a.cap;
                }
                if (active) {
                    double z[NC];
                    walker_normals<NC, RNG>(a.seed, (long long)i, gstep, z);
#pragma unroll
                    for (int c = 0; c < NC; ++c) x[c] = __dadd_rn(x[c], __dmul_rn(a.sigc[c], z[c]));
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) stash.x[c][lane] = x[c];
                v = active ? POT::eval(x, a.pot) : 0.0;
            }
            double u;
            { const uint4 r = pvd_draw(a.seed, (long long)i, gstep, PVD_STREAM_BRANCH, 0u); u = u53(r.x, r.y); }

            // ---- copy counts need Vref of the previous step
            if (!st_ok) {
                SpinGuard guard(spin_limit);
                int got;
                while ((got = poll_state()) == 0) {
                    if (guard.expired()) PVD_RUN_DIE();
                    if (guard.n > 16u) __nanosleep(20);
                }
                if (got < 0) return;
            }
            int cnt = 0, who = 0;
            bool bad = false;
            const bool act2 = i < n;
            if (act2) {
                const double n0 = (double)a.n0;
                cnt = branch_now ? discrete_count(v, vref, a.dt, u, (n0 + n0 * 0.5) + 1.0, bad) : 1;
                if (dw) who = __ldcg(&(q ? a.who_out : a.who_in)[i]);
                const Fx128 fv = fx_from_double_fast(v);
                Fx128 sv{(long long)wm.v_hi[lane], wm.v_lo[lane]};
                sv = fx_add(sv, fv);
                wm.v_hi[lane] = (unsigned long long)sv.hi; wm.v_lo[lane] = sv.lo;
                if (cnt > 0) {
                    Fx128 cv{(long long)wm.cv_hi[lane], wm.cv_lo[lane]};
                    cv = fx_add(cv, cnt == 1 ? fv : fx_mul_small(fv, cnt));
                    wm.cv_hi[lane] = (unsigned long long)cv.hi; wm.cv_lo[lane] = cv.lo;
                }
                wm.vmin[lane] = fmin(wm.vmin[lane], v); wm.vmax[lane] = fmax(wm.vmax[lane], v);
            }
            if (bad) atomicOr(a.err_accum, PVD_ERR_WEIGHT);
            const int incl = warp_incl_scan(cnt);
            const int tile_total = __shfl_sync(0xffffffffu, incl, 31);
            publish_aggregate(a.status + (long long)ps * ra.ntiles_cap, (long long)t, gstep, tile_total);
            csum += tile_total;
            deaths += __popc(__ballot_sync(0xffffffffu, act2 && cnt == 0));
            // Dynamic: the next ticket travels while the previous tile is scattered (requesting it any earlier makes the
            // successors' look-back wait for this warp's publish: measured slower).  Static: the next tile is known, so its
            // counter is fetched now and looked at after the scatter, whose look-back round trips hide this one.
            if (ra.dynamic) { if (lane == 0) tn = atomicAdd(tk, 1u); }
            else if (k > 0 && (long long)(t + nwarps_grid) < ra.ntiles_cap) {
                r_pre = ld_acquire_u32(&(ra.ready + (long long)(ps ^ 1) * ra.ntiles_cap)[t + nwarps_grid]);
                pre = true;
            }
            stash.x[NC][lane] = v;
            stash.cnt[lane] = cnt;
            stash.who[lane] = who;
            __syncwarp();
            if (pending >= 0) {
                const int base = run_scatter_stores<NC>(a, ra, wm.stash[cur ^ 1], ps, q, pending, gstep, pending_total, dw);
                if (lane == 0) { wm.rdy_base[nrdy] = base; wm.rdy_total[nrdy] = pending_total; }
                if (ra.single) nrdy = 0;
                else if (++nrdy == PVD_RUN_RDY) {
                    __threadfence();                            // orders the stores of the last PVD_RUN_RDY scatters before the counters
                    __syncwarp();
                    for (int r = 0; r < PVD_RUN_RDY; ++r) run_publish_ready(ra, ps, wm.rdy_base[r], wm.rdy_total[r]);
                    __syncwarp();
                    nrdy = 0;
                }
            }
            pending = (int)t;
            pending_total = tile_total;
            cur ^= 1;
            t = ra.dynamic ? __shfl_sync(0xffffffffu, tn, 0) : t + nwarps_grid;
        }
        // ---- this warp has no more tiles in step k (st_ok holds: an out-of-range ticket is only recognised with the state).
        // First the sums (Vref is what every other warp will soon wait for), then the scatter of the last tile.
        // warp -> CTA record (shared-memory integer atomics)
        bool finalise = false;
        {
            const Fx128 cv = fx_warp_sum(Fx128{(long long)wm.cv_hi[lane], wm.cv_lo[lane]});
            const Fx128 sv = fx_warp_sum(Fx128{(long long)wm.v_hi[lane], wm.v_lo[lane]});
            const double mn = warp_min(wm.vmin[lane]), mx = warp_max(wm.vmax[lane]);
            if (lane == 0) {
                limb_add_shared(&s_acc[ps][0], cv);
                limb_add_shared(&s_acc[ps][4], sv);
                atomicAdd(&s_acc[ps][8], (unsigned long long)csum);
                atomicAdd(&s_acc[ps][9], (unsigned long long)deaths);
                atomicMin(&s_acc[ps][10], dkey(mn));
                atomicMax(&s_acc[ps][11], dkey(mx));
            }
        }
        unsigned arrived = 0;
        if (lane == 0) { __threadfence_block(); arrived = atomicAdd(&s_arrive[ps], 1u); }
        arrived = __shfl_sync(0xffffffffu, arrived, 0);
        if (arrived == NWARP - 1) {
            // ---- last warp of this CTA for step k: CTA record -> global accumulators.  The atomics return their old values
            // and the arrival is issued only after all of them have come back: they are performed in L2 (the point of
            // coherence of everything this kernel exchanges) before the arrival can be seen -- no fence (a fence at GPU scope
            // costs ~2 us, and three of them in a row were most of the step-to-step latency).
            __threadfence_block();
            unsigned long long back = 0ull;
            if (lane < 12) {
                const unsigned long long val = *(volatile unsigned long long *)&s_acc[ps][lane];
                s_acc[ps][lane] = lane == 10 ? ~0ull : 0ull;    // re-armed for step k + 2
                if (lane == 10) back = atomicMin(&ctl->acc[ps][lane], val);
                else if (lane == 11) back = atomicMax(&ctl->acc[ps][lane], val);
                else back = atomicAdd(&ctl->acc[ps][lane], val);
            }
            if (lane == 0) s_arrive[ps] = 0u;
            unsigned dep = (unsigned)(back ^ (back >> 32));
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) dep ^= __shfl_xor_sync(0xffffffffu, dep, off);
            asm volatile("and.b32 %0, %0, 0;" : "+r"(dep));     // zero, but only once every atomic above has returned
            unsigned g_arrived = 0;
            if (lane == 0) g_arrived = atomicAdd(&ctl->arrive[ps][0], 1u + dep);
            g_arrived = __shfl_sync(0xffffffffu, g_arrived, 0);
            finalise = g_arrived == gridDim.x - 1;
        }
        if (finalise) {
            // ---- last CTA record of step k: finalise the step (one warp of the whole grid)
            // (lane 12: the bookkeeping of the previous step -- which re-armed the counters the NEXT step will use and wrote
            // the state copy read below -- must be complete before this step is published; it has been for a long time)
            unsigned long long l = 0ull;
            if (lane < 12) l = ld_relaxed_u64(&ctl->acc[ps][lane]);
            else if (lane == 12) {
                SpinGuard guard(spin_limit);
                while ((l = ld_relaxed_u64(&ctl->done_seq)) < ra.seq0 + (unsigned long long)k) {
                    if (guard.expired()) break;
                }
            }
            unsigned long long L[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) L[j] = __shfl_sync(0xffffffffu, l, j);
            const long long csum_g = (long long)L[8];
            double *s = a.sums;
            if (lane == 0) {
                const Fx128 cv = limb_combine(L[0], L[1], L[2], L[3]), sv = limb_combine(L[4], L[5], L[6], L[7]);
                const long long deaths_g = (long long)L[9];
                for (int j = 0; j < PVD_SUM_EXT + 4 * a.world; ++j) s[j] = 0.0;
                sum_put(s, PVD_SUM_CV, cv);
                sum_put_double(s, PVD_SUM_C, (double)csum_g);
                sum_put(s, PVD_SUM_V, sv);
                s[PVD_SUM_DEATHS] = (double)deaths_g;
                s[PVD_SUM_BIRTHS] = (double)(csum_g - (long long)n + deaths_g);   // sum max(c-1,0) = sum c - #(c >= 1)
                s[PVD_SUM_NIN] = (double)n;
                s[PVD_SUM_NACC] = (double)n;
                unsigned e = __ldcg(a.err_accum);
                if (csum_g > a.cap) e |= PVD_ERR_CAPACITY;      // the scatter skipped what did not fit
                s[PVD_SUM_ERR] = err_encode(e);
                double *ex = s + PVD_SUM_EXT + 4 * a.rank;
                ex[0] = dkey_inv(L[10]); ex[1] = dkey_inv(L[11]); ex[2] = INFINITY; ex[3] = -INFINITY;
            }
            __syncwarp();
            bool comm_ok = true;
            if constexpr (MULTI) if (a.world > 1 && a.mbox[0]) {
                // several GPUs: the shard's sums go to every peer's mailbox, the world's sums come back (pvd_step.cuh); the
                // other warps of the grid are busy with the move + potential of step k + 1 meanwhile
                mailbox_post(a, ps, gstep, lane, 32);
                mailbox_collect_and_finalize(a, false, ps, &comm_ok);
            }
            if (lane == 0) {
                // the critical part first: Vref and the guards in calc_vref's / birth_or_death's own arithmetic
                // (pyvibdmc.py:651-661, 409-413; the same expressions as finalize_from_sums), published as one 16-byte word
                const double tot_c = sum_get(s, PVD_SUM_C), tot_cv = sum_get(s, PVD_SUM_CV), n0 = (double)a.n0;
                const double v_bar = tot_cv / tot_c;
                const double correction = (tot_c - n0) / n0;
                const double vref_new = v_bar - (a.alpha * correction);
                unsigned err = err_decode(s[PVD_SUM_ERR]);
                if (tot_c < n0 - n0 * 0.5 || tot_c > n0 + n0 * 0.5) err |= PVD_ERR_POP;
                if (!(tot_c > 0.0) || csum_g <= 0) err |= PVD_ERR_EMPTY;
                if (!comm_ok) err |= PVD_ERR_COMM;
                const unsigned long long w0 = ((unsigned long long)(want_seq + 1u) << 32) | (err ? 0x80000000ull : (unsigned long long)(unsigned)csum_g);
                asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(&ctl->pub[0]), "l"(w0), "l"((unsigned long long)__double_as_longlong(vref_new)) : "memory");
                // then the bookkeeping nobody waits for: the state copy of the next step (for the host and for the next
                // finalisation), the log record, the counters of this parity re-armed for step k + 2
                a.st[ps ^ 1].n = csum_g;
                if (!comm_ok) { forward_dead_state(a, ps); a.st[ps ^ 1].err |= PVD_ERR_COMM; }
                else finalize_from_sums(a, false, ps);
                if (a.st[ps ^ 1].err) a.st[ps] = a.st[ps ^ 1];   // the run died in this step: the host may read either copy
            }
            if (lane < 12) st_relaxed_u64(&ctl->acc[ps][lane], lane == 10 ? ~0ull : 0ull);
            if (lane == 0) {
                ctl->arrive[ps][0] = 0u;
                ctl->tk[ps][0] = 0u;                             // every warp has drawn its last ticket of this step
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) st_relaxed_u64(&ctl->done_seq, ra.seq0 + (unsigned long long)k + 1ull);
            __syncwarp();
        }
        // ---- scatter of this warp's last tile of the step; everything not yet announced is announced behind one fence
        if (pending >= 0) {
            const int base = run_scatter_stores<NC>(a, ra, wm.stash[cur ^ 1], ps, q, pending, gstep, pending_total, dw);
            if (lane == 0) { wm.rdy_base[nrdy] = base; wm.rdy_total[nrdy] = pending_total; }
            ++nrdy;
        }
        if (nrdy > 0 && !ra.single) {
            __threadfence();
            __syncwarp();
            for (int r = 0; r < nrdy; ++r) run_publish_ready(ra, ps, wm.rdy_base[r], wm.rdy_total[r]);
            __syncwarp();
        }
    }
}

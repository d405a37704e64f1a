// The per-time-step kernels: replaces the body of DMC_Sim.propagate (pyvibdmc.py:701-876):
//   move_randomly (:540-547) -> potential (:786-793) -> birth_or_death (:380-454) -> calc_vref (:651-661)
// Discrete weighting is ONE kernel per step.  Every warp takes tiles of 32 walkers from ticket
// counters, moves them (Philox + Box-Muller), evaluates V, draws the integer copy count, and a
// single-pass chained scan (decoupled look-back, one status word per tile, driven by the tile's own
// warp) gives each tile its output offset, so each surviving walker is written exactly once,
// already compacted and in np.repeat order, into the other half of a ping-pong buffer.  No
// __syncthreads, no shared memory.  The last warp to finish reduces the per-warp partial sums in
// a fixed order and produces Vref, the population and the per-step log record on the device.
#pragma once
#include "pvd_common.cuh"
#include "pvd_rng.cuh"
#include "pvd_potentials.cuh"

struct StepArgs {
    // walker arrays, SoA: component c of walker i at [c*cap + i]
    const double *xin;
    double *xout;
    const double *vin;          // branch-only kernels: energies used for weighting
    double *vout;
    const int *who_in;
    int *who_out;
    double *w;                  // continuous weights (in place)
    // importance-sampling companions (carried through branching): f_x (NC comps), psi, local kinetic
    const double *fin;
    double *fout;
    const double *psin;
    double *psout;
    const double *lkin;
    double *lkout;
    // control
    DevState *st;               // st[2], indexed by step parity
    unsigned *tickets;          // [2][PVD_WARPS * PVD_TICKET_STRIDE], indexed by step parity
    unsigned *err_accum;        // error bits raised while the step is in flight
    unsigned long long *status; // look-back status words, one per tile
    WarpPartial *part;          // one record per warp of the grid
    pvd_step_stats *ring;
    long long ring_len;
    double *sums;               // PVD_NSUMS doubles: this shard's contribution to the global reduction
    const double *inj_disp;     // injected displacements (SoA, same stride) or nullptr
    const double *inj_u;        // injected uniforms or nullptr
    int *counts_out;            // optional: per-walker copy counts (stand-alone entry point)
    long long *idx_out;         // optional: np.repeat(arange(n), counts)
    int *kill_idx;              // continuous: ascending indices of walkers below the lower threshold
    unsigned *hist;             // continuous: log-spaced histogram of the updated weights
    long long cap;
    long long n0;               // global target population N0
    double dt, alpha, lower, upper;
    unsigned long long seed;
    int parity, do_branch, world, rank, ndim, nc;
    int flip;                   // 1 when this step writes the other ping-pong buffer
    double sigma[PVD_MAX_ATOMS];
    PotParamsDev pot;
};

constexpr int PVD_SUM_CV = 0, PVD_SUM_C = 1, PVD_SUM_BIRTHS = 2, PVD_SUM_DEATHS = 3, PVD_SUM_V = 4,
              PVD_SUM_NIN = 5, PVD_SUM_ERR = 6, PVD_SUM_NACC = 7, PVD_SUM_EXT = 8;   // + 4*rank: vmin,vmax,wmin,wmax

__device__ __forceinline__ unsigned *step_tickets(const StepArgs &a, int parity)
{
    return a.tickets + parity * (PVD_WARPS * PVD_TICKET_STRIDE);
}

// ---------------------------------------------------------------- finalisation
// Turns the (already globally reduced) sums into Vref / population / log record and publishes the
// next step's state copy.  Runs in one thread: by the last warp (single GPU) or by k_finalize after
// the NCCL all-reduce (multi-GPU).
__device__ inline void finalize_from_sums(const StepArgs &a, bool continuous)
{
    const DevState &si = a.st[a.parity];
    DevState &so = a.st[a.parity ^ 1];
    const double *s = a.sums;
    const double tot_c = s[PVD_SUM_C], tot_cv = s[PVD_SUM_CV];
    double vmin = INFINITY, vmax = -INFINITY, wmin = INFINITY, wmax = -INFINITY;
    for (int r = 0; r < a.world; ++r) {
        const double *e = s + PVD_SUM_EXT + 4 * r;
        if (e[0] <= e[1]) { vmin = fmin(vmin, e[0]); vmax = fmax(vmax, e[1]); }
        if (e[2] <= e[3]) { wmin = fmin(wmin, e[2]); wmax = fmax(wmax, e[3]); }
    }
    const double n0 = (double)a.n0;
    // calc_vref (pyvibdmc.py:651-661): v_bar - (alpha * correction)
    const double v_bar = tot_cv / tot_c;
    const double correction = (tot_c - n0) / n0;
    const double vref = v_bar - (a.alpha * correction);
    unsigned err = si.err | (unsigned)s[PVD_SUM_ERR];
    if (!continuous && (tot_c < n0 - n0 * 0.5 || tot_c > n0 + n0 * 0.5)) err |= PVD_ERR_POP;   // :409-413
    if (!(tot_c > 0.0)) err |= PVD_ERR_EMPTY;
    // a failed step does not count and leaves the pre-step ensemble (input buffer) as the valid one,
    // like the reference, which raises before touching the walker arrays (pyvibdmc.py:397-413)
    so.step = err ? si.step : si.step + 1;
    so.buf = err ? si.buf : (si.buf ^ a.flip);
    if (err) so.n = si.n;
    so.vref = err ? si.vref : vref;
    so.pop_global = tot_c;
    so.err = err;
    so.dw_active = si.dw_active;
    so.dt_eff = a.dt;           // imp-samp kernels overwrite this each step before weighting
    so.eff_time = si.eff_time;
    so.done = 0u;
    so.n_accept = 0;
    so.n_kill = 0;
    pvd_step_stats &r = a.ring[si.step % a.ring_len];
    r.vref = vref;
    r.pop = tot_c;
    r.v_avg = s[PVD_SUM_V] / s[PVD_SUM_NIN];
    r.v_max = vmax;
    r.v_min = vmin;
    r.w_max = wmax;
    r.w_min = wmin;
    r.dt_eff = si.dt_eff;
    r.births = (long long)s[PVD_SUM_BIRTHS];
    r.deaths = (long long)s[PVD_SUM_DEATHS];
    r.rejected = (long long)s[PVD_SUM_NIN] - (long long)s[PVD_SUM_NACC];
    r.step = si.step;
}

__global__ void k_finalize(const StepArgs a, int continuous)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (a.st[a.parity].err) return;   // the local step already forwarded the dead state
        finalize_from_sums(a, continuous != 0);
    }
}

// state forwarding when the run is already dead (error raised in an earlier step)
__device__ inline void forward_dead_state(const StepArgs &a)
{
    const DevState &si = a.st[a.parity];
    DevState &so = a.st[a.parity ^ 1];
    so = si;
    so.done = 0u;
}

// per-lane running sums of one warp over all the tiles it processed in this step
struct LaneAcc {
    Fx128 cv = Fx128{0ll, 0ull}, v = Fx128{0ll, 0ull}, cw = Fx128{0ll, 0ull};
    double c = 0.0, vmin = INFINITY, vmax = -INFINITY, wmin = INFINITY, wmax = -INFINITY;
    double births = 0.0, deaths = 0.0, n_in = 0.0, n_acc = 0.0;
};

// End of a step kernel, called by every warp of the grid (converged): publish this warp's partial
// record; the last warp to arrive combines all records (every field is exact / order independent,
// so Vref is bit-reproducible), publishes the shard's sums and, on a single GPU, finalises the step.
// n_local_fixed < 0: the new local population is the inclusive prefix of the last tile.
__device__ inline void warp_finish_step(const StepArgs &a, const LaneAcc &acc, long long ntiles, bool continuous, long long n_local_fixed)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)((gridDim.x * (long long)blockDim.x) >> 5);
    WarpPartial p;
    p.cv = fx_warp_sum(acc.cv); p.v = fx_warp_sum(acc.v); p.cw = fx_warp_sum(acc.cw);
    p.c = warp_sum(acc.c); p.births = warp_sum(acc.births); p.deaths = warp_sum(acc.deaths);
    p.n_in = warp_sum(acc.n_in); p.n_acc = warp_sum(acc.n_acc);
    p.vmin = warp_min(acc.vmin); p.vmax = warp_max(acc.vmax); p.wmin = warp_min(acc.wmin); p.wmax = warp_max(acc.wmax);
    unsigned last = 0;
    if (lane == 0) {
        a.part[gwarp] = p;
        __threadfence();
        const unsigned d = atomicAdd(&a.st[a.parity].done, 1u);
        last = (d == (unsigned)(nwarps - 1)) ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    LaneAcc r;
    for (int wv = lane; wv < nwarps; wv += 32) {
        const WarpPartial *q = &a.part[wv];
        Fx128 t;
        t.hi = __ldcg(&q->cv.hi); t.lo = __ldcg(&q->cv.lo); r.cv = fx_add(r.cv, t);
        t.hi = __ldcg(&q->v.hi); t.lo = __ldcg(&q->v.lo); r.v = fx_add(r.v, t);
        t.hi = __ldcg(&q->cw.hi); t.lo = __ldcg(&q->cw.lo); r.cw = fx_add(r.cw, t);
        r.c += __ldcg(&q->c); r.births += __ldcg(&q->births); r.deaths += __ldcg(&q->deaths);
        r.n_in += __ldcg(&q->n_in); r.n_acc += __ldcg(&q->n_acc);
        r.vmin = fmin(r.vmin, __ldcg(&q->vmin)); r.vmax = fmax(r.vmax, __ldcg(&q->vmax));
        r.wmin = fmin(r.wmin, __ldcg(&q->wmin)); r.wmax = fmax(r.wmax, __ldcg(&q->wmax));
    }
    r.cv = fx_warp_sum(r.cv); r.v = fx_warp_sum(r.v); r.cw = fx_warp_sum(r.cw);
    r.c = warp_sum(r.c); r.births = warp_sum(r.births); r.deaths = warp_sum(r.deaths);
    r.n_in = warp_sum(r.n_in); r.n_acc = warp_sum(r.n_acc);
    r.vmin = warp_min(r.vmin); r.vmax = warp_max(r.vmax); r.wmin = warp_min(r.wmin); r.wmax = warp_max(r.wmax);
    if (lane == 0) {
        double *s = a.sums;
        for (int k = 0; k < PVD_SUM_EXT + 4 * a.world; ++k) s[k] = 0.0;
        s[PVD_SUM_CV] = fx_to_double(r.cv);
        s[PVD_SUM_C] = continuous ? fx_to_double(r.cw) : r.c;
        s[PVD_SUM_V] = fx_to_double(r.v);
        s[PVD_SUM_BIRTHS] = r.births; s[PVD_SUM_DEATHS] = r.deaths; s[PVD_SUM_NIN] = r.n_in; s[PVD_SUM_NACC] = r.n_acc;
        s[PVD_SUM_ERR] = (double)(*a.err_accum);
        double *e = s + PVD_SUM_EXT + 4 * a.rank;
        e[0] = r.vmin; e[1] = r.vmax; e[2] = r.wmin; e[3] = r.wmax;
        long long n_new = n_local_fixed;
        if (n_local_fixed < 0) n_new = (long long)(ld_relaxed_u64(&a.status[ntiles - 1]) & 0xffffffffull);
        a.st[a.parity ^ 1].n = n_new;
        // every warp of this step has drawn its last ticket: re-arm this parity's counters for step s+2
        unsigned *tk = step_tickets(a, a.parity);
        for (int w = 0; w < PVD_WARPS; ++w) tk[w * PVD_TICKET_STRIDE] = 0u;
        if (a.world == 1) finalize_from_sums(a, continuous);
    }
}

// ---------------------------------------------------------------- producers: how a tile obtains (x, V)
// Fused producer: load, displace (Philox or injected), evaluate the built-in potential.
template <class POT, int RNG>
struct ProduceFused {
    static constexpr int NC = POT::NC;
    __device__ static __forceinline__ void run(const StepArgs &a, long long i, long long step, bool active,
                                               double (&x)[POT::NC], double &v)
    {
        if (!active) {
#pragma unroll
            for (int c = 0; c < NC; ++c) x[c] = 1.0 + c;   // harmless geometry for idle lanes
            v = 0.0;
            return;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) x[c] = __ldcs(&a.xin[c * a.cap + i]);
        if (a.inj_disp) {
#pragma unroll
            for (int c = 0; c < NC; ++c) x[c] = x[c] + a.inj_disp[c * a.cap + i];
        } else {
            double z[NC];
            walker_normals<NC, RNG>(a.seed, i, step, z);
#pragma unroll
            for (int c = 0; c < NC; ++c) x[c] = __dadd_rn(x[c], __dmul_rn(a.sigma[c / a.ndim], z[c]));
        }
        v = POT::eval(x, a.pot);
    }
};

// ---------------------------------------------------------------- the fused discrete step
// copy count of birth_or_death (pyvibdmc.py:393-403); bad -> the reference would raise (:397-400)
__device__ __forceinline__ int discrete_count(double v, double vref, double dt, double u, double w_limit, bool &bad)
{
    const double w = exp(__dmul_rn(-1.0 * (v - vref), dt));
    bad = !(w <= w_limit);                       // catches NaN, +inf and w > 1.5 N0 + 1
    if (bad) return 0;
    const double fl = floor(w);
    int c = (int)fl;
    c += (u < (w - fl)) ? 1 : 0;
    return c;
}

// branch_every (pyvibdmc.py:139,828): do_branch > 0 always, < 0 every |do_branch| steps, 0 never
__device__ __forceinline__ bool branch_this_step(int do_branch, long long step)
{
    return do_branch > 0 || (do_branch < 0 && (step % (long long)(-do_branch)) == 0);
}

// common prologue: returns false when the kernel has nothing to do (dead run / empty shard)
__device__ __forceinline__ bool step_prologue(const StepArgs &a)
{
    const DevState *sip = &a.st[a.parity];
    if (sip->err) {
        if (blockIdx.x == 0 && threadIdx.x == 0) forward_dead_state(a);
        return false;
    }
    if (sip->n <= 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { forward_dead_state(a); a.st[a.parity ^ 1].err |= PVD_ERR_EMPTY; }
        return false;
    }
    return true;
}

template <class POT, int RNG>
__global__ void __launch_bounds__(PVD_CTA) k_step_discrete(const StepArgs a)
{
    constexpr int NC = POT::NC;
    if (!step_prologue(a)) return;
    const DevState *sip = &a.st[a.parity];
    const long long n = sip->n, step = sip->step;
    const double vref = sip->vref;
    const long long ntiles = (n + PVD_TILE - 1) / PVD_TILE;
    const bool dw = sip->dw_active != 0;
    const double n0 = (double)a.n0;
    const double w_limit = (n0 + n0 * 0.5) + 1.0;
    const int lane = threadIdx.x & 31;
    const bool branch_now = branch_this_step(a.do_branch, step);
    unsigned *tickets = step_tickets(a, a.parity);
    LaneAcc acc;

    while (true) {
        const long long tile = warp_take_tile(tickets, ntiles);
        if (tile < 0) break;
        const long long i = tile * PVD_TILE + lane;
        const bool active = i < n;

        double x[NC], v;
        ProduceFused<POT, RNG>::run(a, i, step, active, x, v);

        int cnt = 0;
        bool bad = false;
        if (active) {
            if (branch_now) {
                double u;
                if (a.inj_u) u = a.inj_u[i];
                else { const uint4 r = pvd_draw(a.seed, i, step, PVD_STREAM_BRANCH, 0u); u = u53(r.x, r.y); }
                cnt = discrete_count(v, vref, a.dt, u, w_limit, bad);
            } else cnt = 1;
        }
        if (bad) atomicOr(a.err_accum, PVD_ERR_WEIGHT);

        const int incl = warp_incl_scan(cnt);
        const int tile_total = __shfl_sync(0xffffffffu, incl, 31);
        const long long o = warp_lookback(a.status, tile, step, tile_total) + (incl - cnt);
        if (cnt > 0) {
            if (o + cnt > a.cap) atomicOr(a.err_accum, PVD_ERR_CAPACITY);
            else {
                const int who = dw ? a.who_in[i] : 0;
                for (int k = 0; k < cnt; ++k) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) a.xout[c * a.cap + o + k] = x[c];
                    a.vout[o + k] = v;
                    if (dw) a.who_out[o + k] = who;
                }
            }
        }
        if (active) {
            const Fx128 fv = fx_from_double(v);
            acc.v = fx_add(acc.v, fv);
            if (cnt > 0) acc.cv = fx_add(acc.cv, cnt == 1 ? fv : fx_mul_small(fv, cnt));
            acc.c += (double)cnt;
            acc.vmin = fmin(acc.vmin, v); acc.vmax = fmax(acc.vmax, v);
            acc.births += (double)(cnt > 1 ? cnt - 1 : 0); acc.deaths += (cnt == 0) ? 1.0 : 0.0;
            acc.n_in += 1.0; acc.n_acc += 1.0;
        }
    }
    warp_finish_step(a, acc, ntiles, false, -1);
}

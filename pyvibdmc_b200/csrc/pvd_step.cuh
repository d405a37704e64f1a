// The per-time-step kernels: replaces the body of DMC_Sim.propagate (pyvibdmc.py:701-876):
//   move_randomly (:540-547) -> potential (:786-793) -> birth_or_death (:380-454) -> calc_vref (:651-661)
// Discrete weighting is ONE kernel per step: every CTA takes tiles of 256 walkers from a ticket
// counter, moves them, evaluates V, draws the integer copy count, and a single-pass chained scan
// (decoupled look-back) gives each tile its output offset, so each surviving walker is written
// exactly once, already compacted and in np.repeat order, into the other half of a ping-pong
// buffer.  The last CTA to finish reduces the per-tile partial sums in a fixed order and
// produces Vref, the population and the per-step log record; nothing returns to the host.
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
    unsigned *err_accum;        // error bits raised while the step is in flight
    unsigned long long *status; // look-back status words, one per tile
    TilePartial *part;
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
    double sigma[PVD_MAX_ATOMS];
    PotParamsDev pot;
};

constexpr int PVD_SUM_CV = 0, PVD_SUM_C = 1, PVD_SUM_BIRTHS = 2, PVD_SUM_DEATHS = 3, PVD_SUM_V = 4,
              PVD_SUM_NIN = 5, PVD_SUM_ERR = 6, PVD_SUM_NACC = 7, PVD_SUM_EXT = 8;   // + 4*rank: vmin,vmax,wmin,wmax

// ---------------------------------------------------------------- finalisation
// Turns the (already globally reduced) sums into Vref / population / log record and publishes the
// next step's state copy.  Runs in one thread: by the last CTA (single GPU) or by k_finalize after
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
        if (s[PVD_SUM_NIN] > 0.0 && e[0] <= e[1]) { vmin = fmin(vmin, e[0]); vmax = fmax(vmax, e[1]); }
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
    so.step = si.step + 1;
    so.vref = vref;
    so.pop_global = tot_c;
    so.err = err;
    so.dw_active = si.dw_active;
    so.dt_eff = si.dt_eff;
    so.eff_time = si.eff_time;
    so.ticket = 0u;
    so.done = 0u;
    so.n_accept = 0;
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
    so.ticket = 0u;
    so.done = 0u;
}

// Last-CTA reduction of the tile partials (fixed order => run-to-run deterministic Vref).
// Must be called by all PVD_TILE threads of the CTA that finished last.
__device__ inline void reduce_partials_and_publish(const StepArgs &a, int ntiles, long long n_out_local, bool continuous,
                                                   double *sred /* >= 11*PVD_WARPS doubles */)
{
    double cv = 0, c = 0, v = 0, vmin = INFINITY, vmax = -INFINITY, wmin = INFINITY, wmax = -INFINITY;
    double births = 0, deaths = 0, nin = 0, nacc = 0;
    for (int t = threadIdx.x; t < ntiles; t += PVD_TILE) {
        const TilePartial *p = &a.part[t];
        cv += __ldcg(&p->cv); c += __ldcg(&p->c); v += __ldcg(&p->v);
        vmin = fmin(vmin, __ldcg(&p->vmin)); vmax = fmax(vmax, __ldcg(&p->vmax));
        wmin = fmin(wmin, __ldcg(&p->wmin)); wmax = fmax(wmax, __ldcg(&p->wmax));
        births += (double)__ldcg(&p->births); deaths += (double)__ldcg(&p->deaths);
        nin += (double)__ldcg(&p->n_in); nacc += (double)__ldcg(&p->n_acc);
    }
    double vals[11] = {cv, c, v, births, deaths, nin, nacc, vmin, vmax, wmin, wmax};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 7; ++k) vals[k] = warp_sum(vals[k]);
    vals[7] = warp_min(vals[7]); vals[8] = warp_max(vals[8]);
    vals[9] = warp_min(vals[9]); vals[10] = warp_max(vals[10]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 11; ++k) sred[k * PVD_WARPS + wid] = vals[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[11];
        for (int k = 0; k < 11; ++k) {
            double acc = sred[k * PVD_WARPS];
            for (int w = 1; w < PVD_WARPS; ++w) {
                const double x = sred[k * PVD_WARPS + w];
                acc = (k < 7) ? acc + x : ((k == 7 || k == 9) ? fmin(acc, x) : fmax(acc, x));
            }
            t[k] = acc;
        }
        double *s = a.sums;
        for (int k = 0; k < PVD_SUM_EXT + 4 * a.world; ++k) s[k] = 0.0;
        s[PVD_SUM_CV] = t[0]; s[PVD_SUM_C] = t[1]; s[PVD_SUM_V] = t[2];
        s[PVD_SUM_BIRTHS] = t[3]; s[PVD_SUM_DEATHS] = t[4]; s[PVD_SUM_NIN] = t[5]; s[PVD_SUM_NACC] = t[6];
        s[PVD_SUM_ERR] = (double)(*a.err_accum);
        double *e = s + PVD_SUM_EXT + 4 * a.rank;
        e[0] = t[7]; e[1] = t[8]; e[2] = t[9]; e[3] = t[10];
        a.st[a.parity ^ 1].n = n_out_local;
        if (a.world == 1) finalize_from_sums(a, continuous);
    }
    __syncthreads();
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

template <class POT, int RNG>
__global__ void __launch_bounds__(PVD_TILE) k_step_discrete(const StepArgs a)
{
    constexpr int NC = POT::NC;
    __shared__ int s_scan[PVD_WARPS + 1];
    __shared__ long long s_prefix;
    __shared__ int s_tile;
    __shared__ int s_last;
    __shared__ double s_red[11 * PVD_WARPS];

    DevState *sip = &a.st[a.parity];
    const long long n = sip->n, step = sip->step;
    const double vref = sip->vref;
    if (sip->err) {
        if (blockIdx.x == 0 && threadIdx.x == 0) forward_dead_state(a);
        return;
    }
    if (n <= 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { forward_dead_state(a); a.st[a.parity ^ 1].err |= PVD_ERR_EMPTY; }
        return;
    }
    const int ntiles = (int)((n + PVD_TILE - 1) / PVD_TILE);
    const bool dw = sip->dw_active != 0;
    const double n0 = (double)a.n0;
    const double w_limit = (n0 + n0 * 0.5) + 1.0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // branch_every (pyvibdmc.py:139,828): do_branch > 0 always, < 0 every |do_branch| steps, 0 never
    const bool branch_now = a.do_branch > 0 || (a.do_branch < 0 && (step % (long long)(-a.do_branch)) == 0);

    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&sip->ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const long long i = (long long)tile * PVD_TILE + threadIdx.x;
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

        int tile_total;
        const int excl = block_excl_scan(cnt, s_scan, &tile_total);
        const long long prefix = tile_lookback(a.status, tile, step, tile_total, &s_prefix);
        const long long o = prefix + excl;
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
        // tile partial sums
        double pcv = warp_sum((double)cnt * v), pc = warp_sum((double)cnt), pv = warp_sum(active ? v : 0.0);
        double pmin = warp_min(active ? v : INFINITY), pmax = warp_max(active ? v : -INFINITY);
        int pb = warp_sum_i(cnt > 1 ? cnt - 1 : 0), pd = warp_sum_i((active && cnt == 0) ? 1 : 0);
        if (lane == 0) {
            s_red[0 * PVD_WARPS + wid] = pcv; s_red[1 * PVD_WARPS + wid] = pc; s_red[2 * PVD_WARPS + wid] = pv;
            s_red[3 * PVD_WARPS + wid] = pmin; s_red[4 * PVD_WARPS + wid] = pmax;
            s_red[5 * PVD_WARPS + wid] = (double)pb; s_red[6 * PVD_WARPS + wid] = (double)pd;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t[7];
            for (int k = 0; k < 7; ++k) {
                double acc = s_red[k * PVD_WARPS];
                for (int w = 1; w < PVD_WARPS; ++w) {
                    const double y = s_red[k * PVD_WARPS + w];
                    acc = (k == 3) ? fmin(acc, y) : (k == 4 ? fmax(acc, y) : acc + y);
                }
                t[k] = acc;
            }
            TilePartial p;
            p.cv = t[0]; p.c = t[1]; p.v = t[2]; p.vmin = t[3]; p.vmax = t[4];
            p.wmin = INFINITY; p.wmax = -INFINITY;
            p.births = (int)t[5]; p.deaths = (int)t[6];
            const long long rem = n - (long long)tile * PVD_TILE;
            p.n_in = (int)(rem < PVD_TILE ? rem : PVD_TILE);
            p.n_acc = p.n_in;
            a.part[tile] = p;
            __threadfence();
            const unsigned d = atomicAdd(&sip->done, 1u);
            s_last = (d == (unsigned)(ntiles - 1)) ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            // inclusive prefix of the last tile == new local population
            long long n_new = 0;
            if (threadIdx.x == 0) n_new = (long long)(ld_relaxed_u64(&a.status[ntiles - 1]) & 0xffffffffull);
            reduce_partials_and_publish(a, ntiles, n_new, false, s_red);
        }
    }
}

// The per-time-step kernels: replaces the body of DMC_Sim.propagate (pyvibdmc.py:701-876):
//   move_randomly (:540-547) -> potential (:786-793) -> birth_or_death (:380-454) -> calc_vref (:651-661)
// Discrete weighting is ONE kernel per step.  Every warp takes tiles of 32 walkers from a ticket
// counter, moves them (Philox bits -> ziggurat or Box-Muller normals), evaluates V, draws the integer copy count, and a
// single-pass chained scan (decoupled look-back, one status word per tile, driven by the tile's own
// warp) gives each tile its output offset, so each surviving walker is written exactly once,
// already compacted and in np.repeat order, into the other half of a ping-pong buffer.  The
// tile loop has no __syncthreads (shared memory only holds each warp's private stash and the ziggurat table).  The last CTA to finish combines the exact
// (fixed-point) per-CTA partial sums and produces Vref, the population and the per-step log record on the device.
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
    const double *vsin;         // excited-state importance sampling: vector score (pyvibdmc.py:570-573, 608-609), else nullptr
    double *vsout;
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
    int ticket_batch;           // consecutive tiles a warp takes per ticket (1 for heavy tiles, more for light ones)
    double sigma[PVD_MAX_ATOMS];
    double sigc[PVD_MAX_COMP];  // sigma expanded per component (sigc[c] = sigma[c / ndim])
    unsigned long long mbox_epoch; // run epoch (bumped by every upload) folded into the mailbox stamps
    long long mbox_timeout_ticks; // SM clock ticks the exchange waits for a peer's message (PVD_MBOX_TIMEOUT_S, default 120 s)
    int mbox_fold;              // 1: the step kernel's last CTA also waits for the peers and finalises (no k_finalize_mailbox launch)
    double *mbox[PVD_MAX_WORLD]; // NVLink mailbox collective: every rank's mailbox mapped into this process (mbox[0] == nullptr: off)
    PotParamsDev pot;
};

// Layout of the per-step reduction vector (PVD_NSUMS doubles; every rank adds its own, by NCCL or through the mailbox).
// The three floating sums (sum c*V, sum c or sum w, sum V) travel as THREE doubles each: 44-bit chunks of the exact 128-bit
// fixed-point value (quantum 2^-80), so that adding the contributions of up to PVD_MAX_WORLD ranks is exact in ANY order --
// Vref does not depend on how the collective associates (a ring, a tree, the mailbox's rank order: bit-identical), and it is
// formed from the same fixed-point total on one GPU and on eight.  The rest are integer-valued doubles / per-rank slots.
constexpr int PVD_SUM_CV = 0, PVD_SUM_C = 3, PVD_SUM_V = 6, PVD_SUM_BIRTHS = 9, PVD_SUM_DEATHS = 10,
              PVD_SUM_NIN = 11, PVD_SUM_ERR = 12, PVD_SUM_NACC = 13, PVD_SUM_EXT = 16;   // + 4*rank: vmin,vmax,wmin,wmax
static_assert(PVD_NSUMS == PVD_SUM_EXT + 4 * PVD_MAX_WORLD, "include/pvd_b200.h: PVD_NSUMS");

__device__ __forceinline__ void sum_put(double *s, int slot, Fx128 v)
{
    // I = hi * 2^64 + lo (two's complement, units of 2^-80) -> chunks of 44, 44 and the remaining (signed) bits
    const unsigned long long m44 = (1ull << 44) - 1ull;
    const unsigned long long c0 = v.lo & m44;
    const unsigned long long c1 = ((v.lo >> 44) | ((unsigned long long)v.hi << 20)) & m44;
    const long long c2 = v.hi >> 24;                            // arithmetic shift keeps the sign
    s[slot] = (double)c0 * 0x1.0p-80;
    s[slot + 1] = (double)c1 * 0x1.0p-36;
    s[slot + 2] = (double)c2 * 0x1.0p8;
}
__device__ __forceinline__ Fx128 sum_get_fx(const double *s, int slot)
{
    // every sum is an exact integer multiple of its quantum (at most 44 + 3 bits)
    const long long c0 = (long long)(s[slot] * 0x1.0p80), c1 = (long long)(s[slot + 1] * 0x1.0p36), c2 = (long long)(s[slot + 2] * 0x1.0p-8);
    Fx128 r{0ll, (unsigned long long)c0};
    r = fx_add(r, Fx128{c1 >> 20, (unsigned long long)c1 << 44});          // c1 * 2^44
    r = fx_add(r, Fx128{c2 << 24, 0ull});                                   // c2 * 2^88
    return r;
}
__device__ __forceinline__ double sum_get(const double *s, int slot) { return fx_to_double(sum_get_fx(s, slot)); }
__device__ __forceinline__ void sum_put_double(double *s, int slot, double x) { sum_put(s, slot, fx_from_double(x)); }

// Error bits cross the per-step reduction as a SUM of doubles (NCCL all-reduce or the mailbox): bit b travels as 16^b, so
// that up to PVD_MAX_WORLD ranks raising the same bit cannot carry into another one (2 x CAPACITY must not read as COMM).
__device__ __forceinline__ double err_encode(unsigned e)
{
    double v = 0.0, p = 1.0;
    for (int b = 0; b < 8; ++b, p *= 16.0)
        if (e & (1u << b)) v += p;
    return v;
}
__device__ __forceinline__ unsigned err_decode(double v)
{
    unsigned e = 0u;
    unsigned long long u = (unsigned long long)v;
    for (int b = 0; b < 8; ++b, u >>= 4)
        if (u & 15ull) e |= 1u << b;
    return e;
}

__device__ __forceinline__ unsigned *step_tickets(const StepArgs &a, int parity)
{
    return a.tickets + parity * (PVD_WARPS * PVD_TICKET_STRIDE);
}

// DevState as it is in L2 (persistent kernels must not read a copy an earlier step left in L1)
__device__ __forceinline__ DevState load_state_cg(const DevState *p)
{
    DevState s;
    s.n = __ldcg(&p->n); s.step = __ldcg(&p->step); s.vref = __ldcg(&p->vref); s.pop_global = __ldcg(&p->pop_global);
    s.dt_eff = __ldcg(&p->dt_eff); s.eff_time = __ldcg(&p->eff_time); s.err = __ldcg(&p->err); s.dw_active = __ldcg(&p->dw_active);
    s.done = __ldcg(&p->done); s.buf = __ldcg(&p->buf); s.n_accept = __ldcg(&p->n_accept); s.n_kill = __ldcg(&p->n_kill);
    return s;
}

// ---------------------------------------------------------------- finalisation
// Turns the (already globally reduced) sums into Vref / population / log record and publishes the
// next step's state copy.  Runs in one thread: by the last warp (single GPU) or by k_finalize after
// the NCCL all-reduce (multi-GPU).
__device__ inline void finalize_from_sums(const StepArgs &a, bool continuous, int parity)
{
    const DevState si = load_state_cg(&a.st[parity]);
    DevState &so = a.st[parity ^ 1];
    const double *s = a.sums;
    const double tot_c = sum_get(s, PVD_SUM_C), tot_cv = sum_get(s, PVD_SUM_CV);
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
    unsigned err = si.err | err_decode(s[PVD_SUM_ERR]);
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
    so.eff_time = si.eff_time + (a.fin ? si.dt_eff : 0.0);      // accumulated effective time (pyvibdmc.py:372-378)
    so.done = 0u;
    so.n_accept = 0;
    so.n_kill = 0;
    pvd_step_stats &r = a.ring[si.step % a.ring_len];
    r.vref = vref;
    r.pop = tot_c;
    r.v_avg = sum_get(s, PVD_SUM_V) / s[PVD_SUM_NIN];
    r.v_max = vmax;
    r.v_min = vmin;
    r.w_max = wmax;
    r.w_min = wmin;
    r.dt_eff = si.dt_eff;
    r.births = (long long)s[PVD_SUM_BIRTHS];
    r.deaths = (long long)s[PVD_SUM_DEATHS];
    r.rejected = a.fin ? (long long)s[PVD_SUM_NIN] - si.n_accept : 0;
    r.step = si.step;
}

__device__ inline void finalize_from_sums(const StepArgs &a, bool continuous) { finalize_from_sums(a, continuous, a.parity); }
__device__ inline void forward_dead_state(const StepArgs &a, int parity);
__device__ inline void forward_dead_state(const StepArgs &a);

// ---------------------------------------------------------------- per-step collective over NVLink peer memory
// The only exchange of a time step is PVD_NSUMS doubles per shard.  Instead of a NCCL all-reduce kernel
// between the step kernel and the finalisation, the last CTA of the step kernel stores its shard's sums straight into
// every peer's mailbox (slot [step parity][source rank]) over NVLink; one warp on each GPU waits for the world's messages,
// adds the slots in rank order (so every GPU gets bit-identical Vref) and finalises.  Two parities suffice: a rank cannot
// start step s + 2 before it has finalised s + 1, which needs every peer's s + 1 message, which a peer sends only after
// finalising s.
// Every value is its own message: one 16-byte store {bits(value), stamp ^ bits(value)} with stamp = (run epoch, step + 1).
// The reader accepts an entry when word1 ^ word0 equals the stamp it expects -- so there is no data-then-fence-then-flag
// sequence on either side (the two system-scope fences were most of the exchange's latency), and an entry whose two halves
// arrived from different steps (if a 16-byte store were ever torn) cannot pass: it would need old value == new value.
constexpr int PVD_MBOX_STRIDE = PVD_NSUMS + 8;          // 16-byte entries per slot
__device__ __forceinline__ long long mbox_slot(int parity, int rank) { return ((long long)parity * PVD_MAX_WORLD + rank) * PVD_MBOX_STRIDE; }
__device__ __forceinline__ unsigned long long mbox_stamp(const StepArgs &a, long long step) { return (a.mbox_epoch << 40) | (unsigned long long)(step + 1); }

// posts a.sums to every rank's mailbox; called by `nthreads` threads (tid = 0 .. nthreads-1) after a.sums is complete and visible to them
__device__ inline void mailbox_post(const StepArgs &a, int parity, long long step, int tid, int nthreads)
{
    const int nmsg = PVD_SUM_EXT + 4 * a.world;
    const long long slot = mbox_slot(parity, a.rank);
    const unsigned long long stamp = mbox_stamp(a, step);
    for (int t = tid; t < a.world * nmsg; t += nthreads) {
        const int peer = t / nmsg, k = t - peer * nmsg;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(a.sums[k]);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.mbox[peer]) + 2 * (slot + k);
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(bits), "l"(stamp ^ bits) : "memory");
    }
}

// called by ALL threads of the CTA that holds the shard's final sums in a.sums
__device__ inline void mailbox_send(const StepArgs &a, long long step)
{
    __syncthreads();
    mailbox_post(a, a.parity, step, (int)threadIdx.x, (int)blockDim.x);
}

// One warp: waits for the world's messages of this step in its own mailbox, adds them in rank order into a.sums; false: a peer's
// message did not arrive in time (a peer died or fell far behind: the device is never hung).
// (out of line, scalar arguments: its polling arrays must not cost the step kernels that contain it any registers, and a
// by-reference StepArgs would make every caller copy the kernel's parameter block onto its stack)
__device__ __noinline__ bool mailbox_collect_impl(const unsigned long long *mine, unsigned long long want, int world, int parity,
                                                  long long timeout_ticks, double *sums)
{
    const int lane = threadIdx.x & 31;
    const int nmsg = PVD_SUM_EXT + 4 * world;
    bool ok = true;
    const long long t0 = clock64();
    for (int k = lane; k < nmsg; k += 32) {
        // all ranks' entries of this value are requested together; only the late ones are asked for again
        unsigned long long bits[PVD_MAX_WORLD];
        unsigned pending = (1u << world) - 1u;
        while (pending) {
            unsigned long long w0[PVD_MAX_WORLD], w1[PVD_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < PVD_MAX_WORLD; ++r)
                if (pending & (1u << r))
                    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[r]), "=l"(w1[r]) : "l"(mine + 2 * (mbox_slot(parity, r) + k)) : "memory");
#pragma unroll
            for (int r = 0; r < PVD_MAX_WORLD; ++r)
                if ((pending & (1u << r)) && (w1[r] ^ w0[r]) == want) { bits[r] = w0[r]; pending &= ~(1u << r); }
            if (pending && clock64() - t0 > timeout_ticks) { ok = false; break; }
        }
        double v = 0.0;
#pragma unroll
        for (int r = 0; r < PVD_MAX_WORLD; ++r)
            if (r < world && !(pending & (1u << r))) v += __longlong_as_double((long long)bits[r]);
        sums[k] = v;
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    return ok;
}
// One warp: waits for the world's messages of this step in its own mailbox, adds them in rank order into a.sums; false: a peer's
// message did not arrive in time (a peer died or fell far behind: the device is never hung).
__device__ __forceinline__ bool mailbox_collect(const StepArgs &a, int parity, long long cur_step)
{
    return mailbox_collect_impl(reinterpret_cast<const unsigned long long *>(a.mbox[a.rank]), mbox_stamp(a, cur_step), a.world, parity,
                                a.mbox_timeout_ticks, a.sums);
}

// One warp: collect, then finalise (lane 0).
__device__ inline void mailbox_collect_and_finalize(const StepArgs &a, bool continuous, int parity, bool *collect_only = nullptr)
{
    const long long cur_step = __ldcg(&a.st[parity].step);
    const bool ok = mailbox_collect(a, parity, cur_step);
    if (collect_only) { *collect_only = ok; return; }
    if ((threadIdx.x & 31) == 0) {
        if (!ok) {
            forward_dead_state(a, parity);
            a.st[parity ^ 1].err |= PVD_ERR_COMM;
        } else finalize_from_sums(a, continuous, parity);
    }
}
__device__ inline void mailbox_collect_and_finalize(const StepArgs &a, bool continuous) { mailbox_collect_and_finalize(a, continuous, a.parity); }

__global__ void __launch_bounds__(32) k_finalize_mailbox(const StepArgs a, int continuous)
{
    // programmatic dependent launch on both sides: this warp is resident before the step kernel has drained, and the
    // next step's CTAs are launched (and stage their tables) while it waits for the peers' messages
    pdl_wait();
    if (a.st[a.parity].err) return;            // the local step already forwarded the dead state
    mailbox_collect_and_finalize(a, continuous != 0);
}

__global__ void k_finalize(const StepArgs a, int continuous)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (a.st[a.parity].err) return;   // the local step already forwarded the dead state
        finalize_from_sums(a, continuous != 0);
    }
}

// state forwarding when the run is already dead (error raised in an earlier step)
__device__ inline void forward_dead_state(const StepArgs &a, int parity)
{
    const DevState si = load_state_cg(&a.st[parity]);
    DevState &so = a.st[parity ^ 1];
    so = si;
    so.done = 0u;
}
__device__ inline void forward_dead_state(const StepArgs &a) { forward_dead_state(a, a.parity); }

// per-lane running sums of one warp over all the tiles it processed in this step
struct LaneAcc {
    Fx128 cv = Fx128{0ll, 0ull}, v = Fx128{0ll, 0ull}, cw = Fx128{0ll, 0ull};
    double c = 0.0, vmin = INFINITY, vmax = -INFINITY, wmin = INFINITY, wmax = -INFINITY;
    double births = 0.0, deaths = 0.0, n_in = 0.0, n_acc = 0.0;
};

__device__ __forceinline__ void acc_merge(LaneAcc &r, const WarpPartial &q)
{
    r.cv = fx_add(r.cv, q.cv); r.v = fx_add(r.v, q.v); r.cw = fx_add(r.cw, q.cw);
    r.c += q.c; r.births += q.births; r.deaths += q.deaths; r.n_in += q.n_in; r.n_acc += q.n_acc;
    r.vmin = fmin(r.vmin, q.vmin); r.vmax = fmax(r.vmax, q.vmax);
    r.wmin = fmin(r.wmin, q.wmin); r.wmax = fmax(r.wmax, q.wmax);
}
__device__ __forceinline__ WarpPartial acc_warp_reduce(const LaneAcc &acc)
{
    WarpPartial p;
    p.cv = fx_warp_sum(acc.cv); p.v = fx_warp_sum(acc.v); p.cw = fx_warp_sum(acc.cw);
    p.c = warp_sum(acc.c); p.births = warp_sum(acc.births); p.deaths = warp_sum(acc.deaths);
    p.n_in = warp_sum(acc.n_in); p.n_acc = warp_sum(acc.n_acc);
    p.vmin = warp_min(acc.vmin); p.vmax = warp_max(acc.vmax); p.wmin = warp_min(acc.wmin); p.wmax = warp_max(acc.wmax);
    return p;
}

// End of a step kernel, called by every thread of every CTA after its ticket loop (the only CTA
// barriers of the kernel are here, when the CTA has nothing left to do).  Warps -> CTA record ->
// the last CTA to arrive combines all CTA records (every field is exact / order independent, so
// Vref is bit-reproducible), publishes the shard's sums and, on a single GPU, finalises the step.
// n_local_fixed < 0: the new local population is the sum of the copy counts.
__device__ inline void cta_finish_step(const StepArgs &a, const LaneAcc &acc, long long ntiles, bool continuous, long long n_local_fixed,
                                       bool defer_finalize = false)
{
    __shared__ WarpPartial s_part[PVD_WARPS];
    __shared__ unsigned s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
        __threadfence();
        const unsigned d = atomicAdd(&a.st[a.parity].done, 1u);
        s_last = (d == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    LaneAcc r;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += PVD_CTA) {
        const WarpPartial *q = &a.part[b];
        WarpPartial t;
        t.cv.hi = __ldcg(&q->cv.hi); t.cv.lo = __ldcg(&q->cv.lo); t.v.hi = __ldcg(&q->v.hi); t.v.lo = __ldcg(&q->v.lo);
        t.cw.hi = __ldcg(&q->cw.hi); t.cw.lo = __ldcg(&q->cw.lo);
        t.c = __ldcg(&q->c); t.births = __ldcg(&q->births); t.deaths = __ldcg(&q->deaths);
        t.n_in = __ldcg(&q->n_in); t.n_acc = __ldcg(&q->n_acc);
        t.vmin = __ldcg(&q->vmin); t.vmax = __ldcg(&q->vmax); t.wmin = __ldcg(&q->wmin); t.wmax = __ldcg(&q->wmax);
        acc_merge(r, t);
    }
    const WarpPartial pw = acc_warp_reduce(r);
    if (lane == 0) s_part[wid] = pw;
    __syncthreads();
    if (threadIdx.x == 0) {
        LaneAcc f;
        for (int w = 0; w < PVD_WARPS; ++w) acc_merge(f, s_part[w]);
        double *s = a.sums;
        for (int k = 0; k < PVD_SUM_EXT + 4 * a.world; ++k) s[k] = 0.0;
        sum_put(s, PVD_SUM_CV, f.cv);
        if (continuous) sum_put(s, PVD_SUM_C, f.cw);
        else sum_put_double(s, PVD_SUM_C, f.c);
        sum_put(s, PVD_SUM_V, f.v);
        s[PVD_SUM_BIRTHS] = f.births; s[PVD_SUM_DEATHS] = f.deaths; s[PVD_SUM_NIN] = f.n_in; s[PVD_SUM_NACC] = f.n_acc;
        s[PVD_SUM_ERR] = err_encode(*a.err_accum);
        double *e = s + PVD_SUM_EXT + 4 * a.rank;
        e[0] = f.vmin; e[1] = f.vmax; e[2] = f.wmin; e[3] = f.wmax;
        long long n_new = n_local_fixed;
        if (n_local_fixed < 0) n_new = (long long)f.c;      // discrete: the new local population is the sum of the copy counts (exact in double)
        a.st[a.parity ^ 1].n = n_new;
        // every warp of this step has drawn its last ticket: re-arm this parity's counters for step s+2
        unsigned *tk = step_tickets(a, a.parity);
        for (int w = 0; w < PVD_WARPS; ++w) tk[w * PVD_TICKET_STRIDE] = 0u;
        if (a.world == 1 && !defer_finalize) finalize_from_sums(a, continuous);
    }
    if (a.world > 1 && a.mbox[0] && !defer_finalize) {
        mailbox_send(a, a.st[a.parity].step);
        // discrete steps: the same CTA collects the peers' messages and finalises, which takes a kernel launch and its
        // dependency wait out of the per-step chain (a.sums is re-used: the local sums have been sent)
        if (a.mbox_fold) {
            __syncthreads();
            if (threadIdx.x < 32) mailbox_collect_and_finalize(a, continuous);
        }
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
            for (int c = 0; c < NC; ++c) x[c] = __dadd_rn(x[c], __dmul_rn(a.sigc[c], z[c]));
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
    if (sip->n <= 0 && a.world == 1) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { forward_dead_state(a); a.st[a.parity ^ 1].err |= PVD_ERR_EMPTY; }
        return false;
    }
    // (several GPUs: a shard that has run empty takes part in the step with zero tiles -- its peers wait for its message, and the
    // next rebalancing refills it; only an empty WORLD is an error, raised by finalize_from_sums on every rank at once)
    return true;
}

// One warp's stash for the software pipeline: the tile computed in iteration k is scattered in
// iteration k+1, after the warp has computed its next tile, so that the look-back (which must see
// the aggregate of every earlier tile, including tiles that started just before this one) almost
// never has to wait.
template <int NC>
struct TileStash {
    double x[NC + 1][32];     // components, then V
    int cnt[32], excl[32], who[32];
};

// Experiments that did NOT help on this kernel (B200, 1e6 walkers, 2 CTAs/SM; kept out of the code): drawing
// tickets two tiles ahead (+5 us), cp.async prefetch of the next tile's coordinates into shared memory (+8 us),
// requesting the look-back status words before the potential (+3 us), one 256-walker tile per CTA with
// block-level scans (+33 us), 3 or 4 CTAs/SM with spills (+2..4 us), two passes over a 3-pair Box-Muller body
// to shrink the loop below the 32 KB instruction cache (+3.5 us), 64-walker tiles with a double-buffered stash (+9 us),
// status words requested right after the potential (+2.7 us), a three-tile-deep pipeline that resolves a tile behind the next
// tile's random numbers (neutral) or two full tiles late (+6 us), a static first tile per warp (+2.4 us), 192-thread CTAs at
// three per SM (96 registers, 18 warps: +2 us), a rolled Philox round loop (neutral).  What did help: fewer executed
// instructions, the ticket overlapped with the scatter, two look-back windows per round trip, dependent launch.
template <class POT, int RNG>
__global__ void __launch_bounds__(PVD_CTA, POT::MIN_CTAS) k_step_discrete(const StepArgs a)
{
    constexpr int NC = POT::NC;
    __shared__ TileStash<NC> s_stash[PVD_WARPS];
    if constexpr (RNG == PVD_RNG_ZIGGURAT) zig_stage();           // constant table: independent of the previous step
    // launched with programmatic stream serialisation: everything above may overlap the previous step's tail;
    // nothing below may run before that step has completed and its writes are visible
    pdl_wait();
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
    TileStash<NC> &stash = s_stash[threadIdx.x >> 5];
    LaneAcc acc;
    TileFeed feed;
    long long pending = -1;
    int pending_total = 0;

    long long tile = feed_next(feed, tickets, ntiles, a.ticket_batch);
    while (true) {
        double x[NC], v = 0.0;
        int cnt = 0, incl = 0, tile_total = 0, who = 0;
        unsigned issued = 0u;
        if (tile >= 0) {
            const long long i = tile * PVD_TILE + lane;
            const bool active = i < n;
            const double *px = a.xin + i;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                x[c] = 1.0 + c;                                                   // idle lanes: harmless geometry
                if (active) x[c] = __ldcs(px);
                px += a.cap;
            }
            if (active) {
                if (a.inj_disp) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) x[c] = x[c] + a.inj_disp[c * a.cap + i];
                } else {
                    double z[NC];
                    walker_normals<NC, RNG>(a.seed, i, step, z);
#pragma unroll
                    for (int c = 0; c < NC; ++c) x[c] = __dadd_rn(x[c], __dmul_rn(a.sigc[c], z[c]));
                }
            }
            v = active ? POT::eval(x, a.pot) : 0.0;
            bool bad = false;
            if (active) {
                if (branch_now) {
                    double u;
                    if (a.inj_u) u = a.inj_u[i];
                    else { const uint4 r = pvd_draw(a.seed, i, step, PVD_STREAM_BRANCH, 0u); u = u53(r.x, r.y); }
                    cnt = discrete_count(v, vref, a.dt, u, w_limit, bad);
                } else cnt = 1;
                if (dw) who = a.who_in[i];
                const Fx128 fv = fx_from_double(v);
                acc.v = fx_add(acc.v, fv);
                if (cnt > 0) acc.cv = fx_add(acc.cv, cnt == 1 ? fv : fx_mul_small(fv, cnt));
                acc.c += (double)cnt;
                acc.vmin = fmin(acc.vmin, v); acc.vmax = fmax(acc.vmax, v);
                acc.births += (double)(cnt > 1 ? cnt - 1 : 0); acc.deaths += (cnt == 0) ? 1.0 : 0.0;
                acc.n_in += 1.0; acc.n_acc += 1.0;
            }
            if (bad) atomicOr(a.err_accum, PVD_ERR_WEIGHT);
            incl = warp_incl_scan(cnt);
            tile_total = __shfl_sync(0xffffffffu, incl, 31);
            publish_aggregate(a.status, tile, step, tile_total);       // successors can use it right away
            issued = feed_issue(feed, tickets);                        // the next ticket travels while the previous tile is scattered
        }
        if (pending >= 0) {
            // scatter the previous tile: its predecessors have had a whole tile's worth of time to publish
            const long long o = resolve_prefix(a.status, pending, step, pending_total) + stash.excl[lane];
            const int pc = stash.cnt[lane];
            if (pc > 0) {
                if (o + pc > a.cap) atomicOr(a.err_accum, PVD_ERR_CAPACITY);
                else {
                    const double pv = stash.x[NC][lane];
#pragma unroll 1
                    for (int k = 0; k < pc; ++k) {
                        double *po = a.xout + (o + k);
#pragma unroll
                        for (int c = 0; c < NC; ++c) { *po = stash.x[c][lane]; po += a.cap; }
                        a.vout[o + k] = pv;
                        if (dw) a.who_out[o + k] = stash.who[lane];
                    }
                }
            }
            __syncwarp();
        }
        if (tile < 0) break;
#pragma unroll
        for (int c = 0; c < NC; ++c) stash.x[c][lane] = x[c];
        stash.x[NC][lane] = v;
        stash.cnt[lane] = cnt;
        stash.excl[lane] = incl - cnt;
        stash.who[lane] = who;
        __syncwarp();
        pending = tile;
        pending_total = tile_total;
        tile = feed_take(feed, issued, ntiles, a.ticket_batch);
    }
    cta_finish_step(a, acc, ntiles, false, -1);
}


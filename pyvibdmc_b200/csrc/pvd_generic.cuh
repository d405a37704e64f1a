// Layout-generic kernels: AoS<->SoA transposes, stand-alone potential / displacement kernels,
// and the branch-only step (used when V does not come from the fused producer: external Python
// potentials, importance sampling, the stand-alone pvd_branch_discrete entry point).
#pragma once
#include "pvd_step.cuh"

// ---------------------------------------------------------------- layout
// host arrays are (n, nc) row-major ("AoS"); device walkers are SoA with stride cap.
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, long long n, int nc, long long cap)
{
    const long long total = n * nc;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / nc;
        const int c = (int)(t - i * nc);
        soa[c * cap + i] = aos[t];          // coalesced read; writes are nc interleaved streams
    }
}
__global__ void k_soa_to_aos(const double *__restrict__ soa, double *__restrict__ aos, long long n, int nc, long long cap)
{
    const long long total = n * nc;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / nc;
        const int c = (int)(t - i * nc);
        aos[t] = soa[c * cap + i];
    }
}
__global__ void k_iota_int(int *p, long long n, int offset)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        p[t] = offset + (int)t;
}
__global__ void k_int_to_i64(const int *in, long long *out, long long n)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        out[t] = in[t];
}
__global__ void k_fill_double(double *p, long long n, double v)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        p[t] = v;
}

// ---------------------------------------------------------------- stand-alone potential on an AoS batch
// One thread per walker.  The (n, NC) tile of a CTA is staged through shared memory with
// coalesced loads; each thread then reads its own NC values (stride NC doubles: conflict-free
// for odd NC).
template <class POT>
__global__ void __launch_bounds__(PVD_CTA) k_pot_aos(const double *__restrict__ aos, long long n, double *__restrict__ v,
                                                      const PotParamsDev pot)
{
    constexpr int NC = POT::NC;
    __shared__ double tile[PVD_CTA * NC];
    for (long long base = (long long)blockIdx.x * PVD_CTA; base < n; base += (long long)gridDim.x * PVD_CTA) {
        const long long rem = n - base;
        const int cnt = (int)(rem < PVD_CTA ? rem : PVD_CTA);
        const double *src = aos + base * NC;
        for (int t = threadIdx.x; t < cnt * NC; t += PVD_CTA) tile[t] = __ldcs(&src[t]);
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            double x[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) x[c] = tile[threadIdx.x * NC + c];
            v[base + threadIdx.x] = POT::eval(x, pot);
        }
        __syncthreads();
    }
}

// potential on resident SoA walkers (start ensemble, first-step exception pyvibdmc.py:760-762)
template <class POT>
__global__ void __launch_bounds__(PVD_CTA) k_pot_soa(const double *__restrict__ x_soa, const DevState *st, int parity,
                                                      long long cap, double *__restrict__ v, const PotParamsDev pot)
{
    constexpr int NC = POT::NC;
    const long long n = st[parity].n;
    for (long long i = blockIdx.x * (long long)PVD_CTA + threadIdx.x; i < n; i += (long long)gridDim.x * PVD_CTA) {
        double x[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) x[c] = x_soa[c * cap + i];
        v[i] = POT::eval(x, pot);
    }
}

// harmonic potential with a run-time number of components (plug-in entry point pvd_pes_harmonic)
__global__ void k_pot_harm_rt(const double *__restrict__ aos, long long n, int nc, const PotParamsDev pot, double *__restrict__ v)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double acc = __dmul_rn(pot.k[0], __dmul_rn(aos[i * nc], aos[i * nc]));
        for (int c = 1; c < nc; ++c) {
            const double x = aos[i * nc + c];
            acc = __dadd_rn(acc, __dmul_rn(pot.k[c], __dmul_rn(x, x)));
        }
        v[i] = acc;
    }
}

// ---------------------------------------------------------------- displacement (move_randomly, pyvibdmc.py:540-547)
// Generic in the number of components: normals are generated pairwise per walker exactly as in
// walker_normals<>, so the fused kernels and this kernel produce identical streams.
template <int RNG>
__global__ void __launch_bounds__(256, 4) k_displace_soa(double *__restrict__ x, const DevState *st, int parity, long long n_fixed, long long step_fixed,
                               long long cap, int nc, int ndim, unsigned long long seed, const double *__restrict__ inj_disp,
                               const StepArgs *sig_src, const double *__restrict__ sigma_dev, double *__restrict__ z_out)
{
    (void)sig_src;
    __shared__ double s_sig[PVD_MAX_COMP];                 // sigma per component (no integer division in the loop)
    const long long n = st ? st[parity].n : n_fixed;
    const long long step = st ? st[parity].step : step_fixed;
    if (st && st[parity].err) return;
    for (int c = threadIdx.x; c < nc; c += blockDim.x) s_sig[c] = sigma_dev ? sigma_dev[c / ndim] : 1.0;
    if constexpr (RNG == PVD_RNG_ZIGGURAT) zig_stage();
    else __syncthreads();
    const int npairs = (nc + 1) / 2;
    constexpr int G = 3;                                   // pairs generated together: three independent chains in flight
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll 1
        for (int k0 = 0; k0 < npairs; k0 += G) {
            double z0[G], z1[G];
            if (inj_disp) {
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int k = k0 + j;
                    z0[j] = (2 * k < nc) ? inj_disp[(2 * k) * cap + i] : 0.0;
                    z1[j] = (2 * k + 1 < nc) ? inj_disp[(2 * k + 1) * cap + i] : 0.0;
                }
            } else {
                uint4 rr[G];
#pragma unroll
                for (int j = 0; j < G; ++j) rr[j] = pvd_draw(seed, i, step, PVD_STREAM_DISP, (unsigned)(k0 + j));
                if (RNG == PVD_RNG_FP64) {
                    normal_pairs_fp64<G>(rr, z0, z1);
                } else if (RNG == PVD_RNG_ZIGGURAT) {
                    // common path for the 2G normals of the group, then the rare exceptions (as walker_normals_zig)
                    unsigned pend = 0u;
#pragma unroll
                    for (int j = 0; j < G; ++j) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            int layer;
                            double mag;
                            unsigned sgn;
                            zig_decode(h ? rr[j].z : rr[j].x, h ? rr[j].w : rr[j].y, layer, mag, sgn);
                            const double2 t = s_zig_xr[layer];
                            (h ? z1[j] : z0[j]) = zig_signed(mag * t.x, sgn);
                            if (!(mag < t.y) && 2 * (k0 + j) + h < nc) pend |= 1u << (2 * j + h);
                        }
                    }
                    while (pend) {
                        const int b = __ffs((int)pend) - 1;
                        pend &= pend - 1u;
                        const uint4 r = pvd_draw(seed, i, step, PVD_STREAM_DISP, (unsigned)(k0 + (b >> 1)));     // rare: redraw instead of indexing rr[]
                        const double v = zig_slow(seed, i, step, 2 * k0 + b, (b & 1) ? r.z : r.x, (b & 1) ? r.w : r.y);
#pragma unroll
                        for (int j = 0; j < G; ++j) {
                            if (b == 2 * j) z0[j] = v;
                            if (b == 2 * j + 1) z1[j] = v;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < G; ++j) normal_pair<RNG>(rr[j], z0[j], z1[j]);
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int k = k0 + j;
                    if (2 * k < nc) z0[j] = __dmul_rn(s_sig[2 * k], z0[j]);
                    if (2 * k + 1 < nc) z1[j] = __dmul_rn(s_sig[2 * k + 1], z1[j]);
                }
            }
            double *pz = (z_out ? z_out : x) + (long long)(2 * k0) * cap + i;
            if (z_out) {
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int k = k0 + j;
                    if (2 * k < nc) pz[0] = z0[j];
                    if (2 * k + 1 < nc) pz[cap] = z1[j];
                    pz += 2 * cap;
                }
            } else {
                // all loads of the group first, then the stores
                double a0[G], a1[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int k = k0 + j;
                    a0[j] = (2 * k < nc) ? pz[(long long)(2 * j) * cap] : 0.0;
                    a1[j] = (2 * k + 1 < nc) ? pz[(long long)(2 * j + 1) * cap] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int k = k0 + j;
                    if (2 * k < nc) pz[(long long)(2 * j) * cap] = __dadd_rn(a0[j], z0[j]);
                    if (2 * k + 1 < nc) pz[(long long)(2 * j + 1) * cap] = __dadd_rn(a1[j], z1[j]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- branch-only discrete step
// Same counting / chained scan / compaction / finalisation as k_step_discrete, but the energies
// come from memory (a.vin) and every per-walker array is copied memory->memory with a run-time
// number of components.  dt comes from the device state (importance sampling scales it).
// A tile is PVD_BR_SUB sub-tiles of 32 walkers (lane l of sub-tile s owns walker tile*128 + s*32 + l); see k_cont_update.
constexpr int PVD_BR_SUB = 4;
constexpr int PVD_BR_TILE = PVD_TILE * PVD_BR_SUB;
#ifndef PVD_BR_MINB
#define PVD_BR_MINB 2               // A/B: resident CTAs per SM the register budget is cut for (2: 128 registers, 3: 80, 4: 64)
#endif
__global__ void __launch_bounds__(PVD_CTA, PVD_BR_MINB) k_branch_discrete(const StepArgs a)
{
    if (!step_prologue(a)) return;
    const DevState *sip = &a.st[a.parity];
    const long long n = sip->n, step = sip->step;
    const double vref = sip->vref, dt = sip->dt_eff;
    const long long ntiles = (n + PVD_BR_TILE - 1) / PVD_BR_TILE;
    const bool dw = sip->dw_active != 0;
    const double n0 = (double)a.n0;
    const double w_limit = (n0 + n0 * 0.5) + 1.0;
    const int lane = threadIdx.x & 31;
    const bool branch_now = branch_this_step(a.do_branch, step);
    unsigned *tickets = step_tickets(a, a.parity);
    LaneAcc acc;
    TileFeed feed;
    // the previous tile of this warp, scattered one iteration late (see k_step_discrete)
    long long pend = -1;
    int pend_total = 0;
    int pend_cnt[PVD_BR_SUB], pend_excl[PVD_BR_SUB];

    long long tile = feed_next(feed, tickets, ntiles, 1);
    while (true) {
        int cnt[PVD_BR_SUB], excl[PVD_BR_SUB];
        int tile_total = 0;
        unsigned issued = 0u;
        if (tile >= 0) {
            bool bad = false;
#pragma unroll
            for (int s = 0; s < PVD_BR_SUB; ++s) {
                const long long i = tile * PVD_BR_TILE + s * PVD_TILE + lane;
                int c = 0;
                if (i < n) {
                    const double v = a.vin[i];
                    if (branch_now) {
                        double u;
                        bool b1 = false;
                        if (a.inj_u) u = a.inj_u[i];
                        else { const uint4 r = pvd_draw(a.seed, i, step, PVD_STREAM_BRANCH, 0u); u = u53(r.x, r.y); }
                        c = discrete_count(v, vref, dt, u, w_limit, b1);
                        bad |= b1;
                    } else c = 1;
                    if (a.counts_out) a.counts_out[i] = c;
                    const Fx128 fv = fx_from_double(v);
                    acc.v = fx_add(acc.v, fv);
                    if (c > 0) acc.cv = fx_add(acc.cv, c == 1 ? fv : fx_mul_small(fv, c));
                    acc.c += (double)c;
                    acc.vmin = fmin(acc.vmin, v); acc.vmax = fmax(acc.vmax, v);
                    acc.births += (double)(c > 1 ? c - 1 : 0); acc.deaths += (c == 0) ? 1.0 : 0.0;
                    acc.n_in += 1.0; acc.n_acc += 1.0;
                }
                const int incl = warp_incl_scan(c);
                cnt[s] = c;
                excl[s] = tile_total + incl - c;
                tile_total += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (bad) atomicOr(a.err_accum, PVD_ERR_WEIGHT);
            publish_aggregate(a.status, tile, step, tile_total);
            issued = feed_issue(feed, tickets);                        // the next ticket travels while the previous tile is copied
        }
        if (pend >= 0) {
            const long long base = resolve_prefix(a.status, pend, step, pend_total);
#pragma unroll
            for (int s = 0; s < PVD_BR_SUB; ++s) {
                const int pc = pend_cnt[s];
                if (pc <= 0) continue;
                const long long pi_ = pend * PVD_BR_TILE + s * PVD_TILE + lane;
                const long long o = base + pend_excl[s];
                if (o + pc > a.cap) { atomicOr(a.err_accum, PVD_ERR_CAPACITY); continue; }
                // every array is read ONCE (nine components at a time, all loads in flight together) and written pc times
#pragma unroll 1
                for (int c0 = 0; c0 < a.nc; c0 += 9) {
                    double r[9];
                    const double *pi = a.xin + pi_ + (long long)c0 * a.cap;
#pragma unroll
                    for (int j = 0; j < 9; ++j) { if (c0 + j < a.nc) r[j] = __ldcs(pi); pi += a.cap; }
#pragma unroll 1
                    for (int k = 0; k < pc; ++k) {
                        double *po = a.xout + (o + k) + (long long)c0 * a.cap;
#pragma unroll
                        for (int j = 0; j < 9; ++j) { if (c0 + j < a.nc) *po = r[j]; po += a.cap; }
                    }
                }
                if (a.fin) {
#pragma unroll 1
                    for (int c0 = 0; c0 < a.nc; c0 += 9) {
                        double r[9];
                        const double *fi = a.fin + pi_ + (long long)c0 * a.cap;
#pragma unroll
                        for (int j = 0; j < 9; ++j) { if (c0 + j < a.nc) r[j] = __ldcs(fi); fi += a.cap; }
#pragma unroll 1
                        for (int k = 0; k < pc; ++k) {
                            double *fo = a.fout + (o + k) + (long long)c0 * a.cap;
#pragma unroll
                            for (int j = 0; j < 9; ++j) { if (c0 + j < a.nc) *fo = r[j]; fo += a.cap; }
                        }
                    }
                }
                {
                    const double v = a.vout ? a.vin[pi_] : 0.0;
                    const int who = dw ? a.who_in[pi_] : 0;
                    const double ps = a.fin ? a.psin[pi_] : 0.0, lk = a.fin ? a.lkin[pi_] : 0.0, vs = a.vsin ? a.vsin[pi_] : 0.0;
#pragma unroll 1
                    for (int k = 0; k < pc; ++k) {
                        if (a.vout) a.vout[o + k] = v;
                        if (dw) a.who_out[o + k] = who;
                        if (a.idx_out) a.idx_out[o + k] = pi_;
                        if (a.fin) { a.psout[o + k] = ps; a.lkout[o + k] = lk; }
                        if (a.vsin) a.vsout[o + k] = vs;
                    }
                }
            }
        }
        if (tile < 0) break;
        pend = tile; pend_total = tile_total;
#pragma unroll
        for (int s = 0; s < PVD_BR_SUB; ++s) { pend_cnt[s] = cnt[s]; pend_excl[s] = excl[s]; }
        tile = feed_take(feed, issued, ntiles, 1);
    }
    cta_finish_step(a, acc, ntiles, false, -1);
}

// ---------------------------------------------------------------- first Vref (pyvibdmc.py:760-769) and stand-alone calc_vref
// single CTA, fixed-order reduction; writes local sums then (world==1) finalises into st[parity]
// itself (no step advance): used once after upload.
__global__ void __launch_bounds__(1024) k_init_sums(const double *__restrict__ v, const double *__restrict__ w,
                                                    const DevState *st, int parity, long long n_fixed, double *sums, int world, int rank)
{
    __shared__ double s1[32], s2[32];
    const long long n = st ? st[parity].n : n_fixed;
    double a = 0.0, b = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double wi = w ? w[i] : 1.0;
        a += wi * v[i];
        b += wi;
    }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0, tb = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { ta += s1[k]; tb += s2[k]; }
        for (int k = 0; k < PVD_SUM_EXT + 4 * world; ++k) sums[k] = 0.0;
        sum_put_double(sums, PVD_SUM_CV, ta);
        sum_put_double(sums, PVD_SUM_C, tb);
        (void)rank;
    }
}
__global__ void k_init_finalize(DevState *st, int parity, const double *sums, double alpha, long long n0_, double dt)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double n0 = (double)n0_;
        const double tot_c = sum_get(sums, PVD_SUM_C);
        const double v_bar = sum_get(sums, PVD_SUM_CV) / tot_c;
        const double correction = (tot_c - n0) / n0;
        DevState &s = st[parity];
        s.vref = v_bar - (alpha * correction);
        s.pop_global = tot_c;
        s.dt_eff = dt;
    }
}

// descendant weights (calc_desc_wts, pyvibdmc.py:663-672): histogram of who_from (discrete) or
// weight-sum per parent (continuous)
__global__ void k_desc_wts(const int *__restrict__ who, const double *__restrict__ w, const DevState *st, int parity,
                           long long n_fixed, int base, long long n_parent, double *__restrict__ out)
{
    const long long n = st ? st[parity].n : n_fixed;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long q = (long long)who[i] - base;
        if (q >= 0 && q < n_parent) atomicAdd(&out[q], w ? w[i] : 1.0);
    }
}

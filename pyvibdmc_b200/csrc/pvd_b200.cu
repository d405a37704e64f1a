// libpvd_b200: C-ABI implementation (see include/pvd_b200.h).  sm_100a only.
#include <stdlib.h>
#include "pvd_common.cuh"
#include "pvd_rng.cuh"
#include "pvd_potentials.cuh"
#include "pvd_step.cuh"
#include "pvd_run.cuh"
#include "pvd_gather.cuh"
#include "pvd_generic.cuh"
#include "pvd_continuous.cuh"
#include "pvd_impsamp.cuh"
#include "pvd_impext.cuh"
#include "pvd_nn.cuh"
#include "pvd_descriptor.cuh"

#include <mutex>

#ifndef PVD_RUN_VARIANT_H2O
#define PVD_RUN_VARIANT_H2O 2563
#endif
thread_local std::string g_pvd_err;
std::atomic<long long> g_pvd_launches{0};
static thread_local double g_last_kernel_ms = 0.0;

// Launch with programmatic stream serialisation (see pdl_wait in pvd_common.cuh).  Only for kernels that call pdl_wait()
// before they touch simulation state.  PVD_NO_PDL=1 turns the attribute off (A/B runs).
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
{
    static const bool use_pdl = getenv("PVD_NO_PDL") == nullptr;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = grid;
    lc.blockDim = block;
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    lc.attrs = at;
    lc.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------- small RAII helpers (host)
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    // grows only: staging buffers are re-used across calls (cudaFree / cudaMalloc of a 70 MB buffer cost ~2 ms each)
    cudaError_t alloc(size_t b)
    {
        if (p && b <= bytes) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; }
        bytes = b;
        return cudaMalloc(&p, b ? b : 1);
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    cudaError_t init() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
};

// ---------------------------------------------------------------- ziggurat table (host, long double)
// Layers of equal area V under f(x) = exp(-x^2/2), x >= 0 (Marsaglia & Tsang 2000): x_1 = R, x_{i+1} = f^-1(V/x_i + f(x_i)),
// closed by x_N = 0; V = R f(R) + int_R^inf f.  R is found by bisection on the closing condition.
#include <math.h>
struct ZigHost {
    double x[PVD_ZIG_N + 1], f[PVD_ZIG_N + 1];
    double2 xr[PVD_ZIG_N];
    ZigHost()
    {
        static long double xl[PVD_ZIG_N + 1];
        auto closing = [&](long double R, long double &V) -> long double {
            const long double fR = expl(-0.5L * R * R);
            V = R * fR + sqrtl(2.0L * atanl(1.0L)) * erfcl(R / sqrtl(2.0L));     // sqrt(pi/2) erfc(R/sqrt 2)
            xl[0] = V / fR;
            xl[1] = R;
            for (int i = 1; i < PVD_ZIG_N - 1; ++i) {
                const long double y = V / xl[i] + expl(-0.5L * xl[i] * xl[i]);
                if (y >= 1.0L) return 1.0L;                                      // reached the top too early: R too small
                xl[i + 1] = sqrtl(-2.0L * logl(y));
            }
            return V / xl[PVD_ZIG_N - 1] + expl(-0.5L * xl[PVD_ZIG_N - 1] * xl[PVD_ZIG_N - 1]) - 1.0L;
        };
        long double lo = 2.0L, hi = 7.0L, V = 0.0L;
        for (int it = 0; it < 200; ++it) {
            const long double mid = 0.5L * (lo + hi);
            if (closing(mid, V) > 0.0L) lo = mid; else hi = mid;
        }
        closing(hi, V);
        xl[PVD_ZIG_N] = 0.0L;
        for (int i = 0; i <= PVD_ZIG_N; ++i) {
            x[i] = (double)xl[i];
            f[i] = (double)expl(-0.5L * xl[i] * xl[i]);
        }
        f[PVD_ZIG_N] = 1.0;
        for (int i = 0; i < PVD_ZIG_N; ++i) xr[i] = make_double2(x[i], (double)(xl[i + 1] / xl[i]));
    }
};
static const ZigHost &zig_host()
{
    static const ZigHost z;
    return z;
}
static cudaError_t zig_upload()
{
    const ZigHost &z = zig_host();
    cudaError_t e = cudaMemcpyToSymbol(g_zig_xr, z.xr, sizeof(z.xr));
    return e != cudaSuccess ? e : cudaMemcpyToSymbol(g_zig_f, z.f, sizeof(z.f));
}
extern "C" int pvd_ziggurat_table(double *x, double *f)
{
    PVD_REQUIRE(x && f, "NULL argument");
    const ZigHost &z = zig_host();
    for (int i = 0; i <= PVD_ZIG_N; ++i) { x[i] = z.x[i]; f[i] = z.f[i]; }
    return PVD_OK;
}

static int g_num_sms = 0;
static std::once_flag g_init_once;
static cudaError_t g_init_err = cudaSuccess;
static int g_const_device = -1;

static int ensure_device_ready()
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return pvd_fail(PVD_E_NODEVICE, std::string("no CUDA device available: ") + cudaGetErrorString(e));
    int dev = 0;
    PVD_CUDA(cudaGetDevice(&dev));
    if (g_const_device != dev) {          // constants are per-device (one process per GPU in practice)
        PVD_CUDA(ps_upload_constants());
        PVD_CUDA(zig_upload());
        cudaDeviceProp prop;
        PVD_CUDA(cudaGetDeviceProperties(&prop, dev));
        g_num_sms = prop.multiProcessorCount;
        g_const_device = dev;
    }
    return PVD_OK;
}
static inline int grid_for(long long n, int per_block, int max_waves = 8)
{
    long long g = (n + per_block - 1) / per_block;
    const long long cap = (long long)(g_num_sms > 0 ? g_num_sms : 148) * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" {

int pvd_abi_version(void) { return PVD_ABI_VERSION; }
int pvd_sizeof_config(void) { return (int)sizeof(pvd_config); }
int pvd_sizeof_step_stats(void) { return (int)sizeof(pvd_step_stats); }
const char *pvd_last_error(void) { return g_pvd_err.c_str(); }
int64_t pvd_launch_count(void) { return (int64_t)g_pvd_launches.load(); }
int pvd_last_kernel_ms(double *ms) { PVD_REQUIRE(ms, "ms is NULL"); *ms = g_last_kernel_ms; return PVD_OK; }

int pvd_device_count(int *count)
{
    PVD_REQUIRE(count, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return pvd_fail(PVD_E_NODEVICE, cudaGetErrorString(e)); }
    *count = n;
    return PVD_OK;
}
int pvd_set_device(int device)
{
    PVD_CUDA(cudaSetDevice(device));
    return ensure_device_ready();
}

}  // extern "C"

// ---------------------------------------------------------------- FP64 peak micro-benchmark
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.9999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;     // never true; keeps the chain alive
}

extern "C" int pvd_measure_fp64_peak(double *flops_per_s)
{
    PVD_REQUIRE(flops_per_s, "flops_per_s is NULL");
    if (int rc = ensure_device_ready()) return rc;
    DevBuf out;
    PVD_CUDA(out.alloc(8));
    EventPair ev;
    PVD_CUDA(ev.init());
    const int iters = 1 << 14, blocks = g_num_sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        PVD_CUDA(cudaEventRecord(ev.a));
        k_fp64_peak<<<blocks, 256>>>(out.as<double>(), iters, 1.0);
        PVD_CHECK_LAUNCH();
        PVD_CUDA(cudaEventRecord(ev.b));
        PVD_CUDA(cudaEventSynchronize(ev.b));
        float ms = 0;
        PVD_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
        const double fl = 2.0 * 8.0 * (double)iters * 256.0 * blocks / (ms * 1e-3);
        if (rep > 0 && fl > best) best = fl;
    }
    *flops_per_s = best;
    return PVD_OK;
}

// ---------------------------------------------------------------- stand-alone host entry points
// Shared pattern: H2D, one kernel timed with events, D2H.
template <class LaunchFn>
static int run_host_kernel(const void *hin, size_t in_bytes, void *hout, size_t out_bytes, LaunchFn launch)
{
    if (int rc = ensure_device_ready()) return rc;
    DevBuf din, dout;
    PVD_CUDA(din.alloc(in_bytes));
    PVD_CUDA(dout.alloc(out_bytes));
    EventPair ev;
    PVD_CUDA(ev.init());
    if (in_bytes) PVD_CUDA(cudaMemcpy(din.p, hin, in_bytes, cudaMemcpyHostToDevice));
    PVD_CUDA(cudaEventRecord(ev.a));
    launch(din.p, dout.p);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaEventRecord(ev.b));
    if (out_bytes) PVD_CUDA(cudaMemcpy(hout, dout.p, out_bytes, cudaMemcpyDeviceToHost));
    PVD_CUDA(cudaEventSynchronize(ev.b));
    float ms = 0;
    PVD_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    g_last_kernel_ms = ms;
    return PVD_OK;
}

extern "C" {

int pvd_pes_h2o(const double *xyz, int64_t n, double *v)
{
    PVD_REQUIRE(n >= 0 && (n == 0 || (xyz && v)), "pvd_pes_h2o: bad arguments");
    if (n == 0) return ensure_device_ready();
    PotParamsDev pot{};
    return run_host_kernel(xyz, (size_t)n * 72, v, (size_t)n * 8, [&](void *in, void *out) {
        k_pot_aos<PotH2O><<<grid_for(n, PVD_CTA, 16), PVD_CTA>>>((const double *)in, n, (double *)out, pot);
    });
}

int pvd_pes_h2o_params(double *c245, double *scal8)
{
    PVD_REQUIRE(c245 && scal8, "NULL output");
    ps_fold_host(c245, scal8);
    return PVD_OK;
}

int pvd_pes_harmonic(const double *x, int64_t n, int32_t ncomp, const double *k, double *v)
{
    PVD_REQUIRE(n >= 0 && ncomp >= 1 && ncomp <= PVD_MAX_COMP && k && (n == 0 || (x && v)), "pvd_pes_harmonic: bad arguments");
    if (n == 0) return ensure_device_ready();
    PotParamsDev pot{};
    for (int c = 0; c < ncomp; ++c) pot.k[c] = k[c];
    return run_host_kernel(x, (size_t)n * ncomp * 8, v, (size_t)n * 8, [&](void *in, void *out) {
        k_pot_harm_rt<<<grid_for(n, 256, 16), 256>>>((const double *)in, n, ncomp, pot, (double *)out);
    });
}

int pvd_pes_morse1d(const double *x, int64_t n, double de, double alpha, double *v)
{
    PVD_REQUIRE(n >= 0 && (n == 0 || (x && v)), "pvd_pes_morse1d: bad arguments");
    if (n == 0) return ensure_device_ready();
    PotParamsDev pot{};
    pot.k[0] = de;
    pot.k[1] = alpha;
    return run_host_kernel(x, (size_t)n * 8, v, (size_t)n * 8, [&](void *in, void *out) {
        k_pot_aos<PotMorse><<<grid_for(n, PVD_CTA, 16), PVD_CTA>>>((const double *)in, n, (double *)out, pot);
    });
}

static int displace_impl(double *xyz, int64_t n, int32_t nc, int32_t ndim, const double *sigma, uint64_t seed, uint64_t step,
                         int32_t rng_mode, bool normals_only)
{
    PVD_REQUIRE(n >= 0 && nc >= 1 && nc <= PVD_MAX_COMP && ndim >= 1 && (n == 0 || xyz), "pvd_displace: bad arguments");
    PVD_REQUIRE(rng_mode == PVD_RNG_FP64 || rng_mode == PVD_RNG_FAST || rng_mode == PVD_RNG_ZIGGURAT, "pvd_displace: unknown rng_mode");
    if (int rc = ensure_device_ready()) return rc;
    if (n == 0) return PVD_OK;
    // stage AoS -> SoA (stride n), displace, SoA -> AoS
    DevBuf aos, soa, zbuf, sig;
    PVD_CUDA(aos.alloc((size_t)n * nc * 8));
    PVD_CUDA(soa.alloc((size_t)n * nc * 8));
    PVD_CUDA(sig.alloc(PVD_MAX_ATOMS * 8));
    double sg[PVD_MAX_ATOMS];
    for (int a = 0; a < PVD_MAX_ATOMS; ++a) sg[a] = (sigma && a < (nc + ndim - 1) / ndim) ? sigma[a] : 1.0;
    PVD_CUDA(cudaMemcpy(sig.p, sg, sizeof(sg), cudaMemcpyHostToDevice));
    EventPair ev;
    PVD_CUDA(ev.init());
    const int g = grid_for(n, 256, 16);
    if (!normals_only) {
        PVD_CUDA(cudaMemcpy(aos.p, xyz, (size_t)n * nc * 8, cudaMemcpyHostToDevice));
        k_aos_to_soa<<<g, 256>>>(aos.as<double>(), soa.as<double>(), n, nc, n);
        PVD_CHECK_LAUNCH();
    }
    PVD_CUDA(cudaEventRecord(ev.a));
    double *zout = normals_only ? soa.as<double>() : nullptr;
    if (rng_mode == PVD_RNG_FP64)
        k_displace_soa<PVD_RNG_FP64><<<g, 256>>>(soa.as<double>(), nullptr, 0, n, (long long)step, n, nc, ndim, seed, nullptr, nullptr,
                                                 sig.as<double>(), zout);
    else if (rng_mode == PVD_RNG_ZIGGURAT)
        k_displace_soa<PVD_RNG_ZIGGURAT><<<g, 256>>>(soa.as<double>(), nullptr, 0, n, (long long)step, n, nc, ndim, seed, nullptr, nullptr,
                                                     sig.as<double>(), zout);
    else
        k_displace_soa<PVD_RNG_FAST><<<g, 256>>>(soa.as<double>(), nullptr, 0, n, (long long)step, n, nc, ndim, seed, nullptr, nullptr,
                                                 sig.as<double>(), zout);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaEventRecord(ev.b));
    k_soa_to_aos<<<g, 256>>>(soa.as<double>(), aos.as<double>(), n, nc, n);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpy(xyz, aos.p, (size_t)n * nc * 8, cudaMemcpyDeviceToHost));
    float ms = 0;
    PVD_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    g_last_kernel_ms = ms;
    return PVD_OK;
}

int pvd_displace(double *xyz, int64_t n, int32_t natoms, int32_t ndim, const double *sigma, uint64_t seed, uint64_t step,
                 int32_t rng_mode)
{
    PVD_REQUIRE(natoms >= 1 && natoms <= PVD_MAX_ATOMS && sigma, "pvd_displace: bad natoms / sigma");
    return displace_impl(xyz, n, natoms * ndim, ndim, sigma, seed, step, rng_mode, false);
}
int pvd_normals(double *z, int64_t n, int32_t ncomp, uint64_t seed, uint64_t step, int32_t rng_mode)
{
    return displace_impl(z, n, ncomp, 1 << 20, nullptr, seed, step, rng_mode, true);
}

}  // extern "C"

__global__ void k_philox_kat(const unsigned *ctr, const unsigned *key, unsigned *out)
{
    const uint4 r = philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

extern "C" int pvd_philox_kat(const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4)
{
    PVD_REQUIRE(ctr4 && key2 && out4, "NULL argument");
    if (int rc = ensure_device_ready()) return rc;
    DevBuf b;
    PVD_CUDA(b.alloc(40));
    unsigned h[6] = {ctr4[0], ctr4[1], ctr4[2], ctr4[3], key2[0], key2[1]};
    PVD_CUDA(cudaMemcpy(b.p, h, 24, cudaMemcpyHostToDevice));
    k_philox_kat<<<1, 1>>>(b.as<unsigned>(), b.as<unsigned>() + 4, b.as<unsigned>() + 6);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpy(out4, b.as<unsigned>() + 6, 16, cudaMemcpyDeviceToHost));
    return PVD_OK;
}

// ---------------------------------------------------------------- the simulation handle
struct pvd_sim {
    pvd_config cfg;
    int nc = 0;
    long long cap = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // walker arrays (ping-pong for discrete compaction)
    DevBuf x[2], v[2], who[2], w, f[2], psi[2], lk[2], vs[2];
    DevBuf st, err_accum, status, part, ring, sums, sigma_dev, tickets;
    DevBuf run_ctl, run_ready;                     // resident multi-step kernel (pvd_run.cuh)
    unsigned long long run_seq = 0;                // steps finalised by resident launches since the last upload
    long long ntiles_cap = 0;
    bool resident = true;                          // pvd_sim_set_resident
    int resident_mode = 1;                         // 0 off, 1 by ensemble size, 2 always
    int run_grid = 0, run_minb = 0, run_occ = 0;                // cooperative grid (all CTAs co-resident) and the occupancy variant in use
    DevBuf g_cnt[2], g_tincl[2], g_cbase[2], g_meta[2], g_seg;  // deferred-compaction steps (pvd_gather.cuh): per ping-pong buffer
    int gather_grid = 0, gather_id = 0;
    bool deferred = false;                         // x[cur] is in deferred form (a chain of gather steps is open: gather_flush() closes it)
    int defer_buf0 = 0;                            // buffer the open chain started from (compacted input of its first step)
    long long defer_steps = 0;                     // steps enqueued in the open chain
    size_t gather_smem = 0;
    DevBuf inj_disp, inj_u, inj_um, stage, stage2;   // staging for host<->device transposes / injections
    DevBuf parent_x, parent_w;
    DevBuf dump_desc, dump_parent, dump_pw;        // wave-function dump in flight (descendant weights, parent ensemble): device copies
    void *dump_host = nullptr;
    size_t dump_host_bytes = 0;
    cudaEvent_t dump_ready = nullptr, dump_done = nullptr;
    bool dump_pending = false;
    long long dump_n = 0;
    DevBuf xfer;                                   // packed walkers on their way to / from another shard (device-to-device rebalancing)
    DevBuf kill_idx, kill_mask, hist, cand, cand_sorted, bin_start, bin_fill, cont_work, copy_dst, copy_src, cont_queue, cont_root, cont_skip;
    DevBuf trial_table, acc_count;
    DevBuf impx_y, impx_fy, impx_sec, impx_psiy, impx_invm;      // importance sampling with a user trial wave function (pvd_impext.cuh)
    NNDeviceWeights nn_w;
    TrialParamsDev trial_params{};
    int nn_grid = 1;
    long long parent_n = 0;
    void *sums_ext = nullptr;   // caller-owned reduction buffer (multi-GPU)
    long long ntrial = 0;
    int parity = 0;      // state copy the next enqueued step reads
    int cur = 0;         // buffer holding the current walkers
    long long n_uploaded = 0;
    bool uploaded = false, ext_moved = false;
    int grid = 1, grid_light = 1;
    int ticket_batch = 1;
    // asynchronous snapshot (checkpoints / dumps): device copies on the compute stream, D2H on a side stream into pinned memory
    DevBuf snap_x, snap_v, snap_w, snap_who, snap_state;
    void *snap_host = nullptr;
    size_t snap_host_bytes = 0;
    cudaStream_t snap_stream = nullptr;
    cudaEvent_t snap_ready = nullptr, snap_done = nullptr;
    bool snap_pending = false;
    int snap_parity = 0;
    DevBuf mbox;                                   // this rank's mailbox (NVLink collective)
    double *peer_mbox[PVD_MAX_WORLD]{};            // every rank's mailbox mapped here (own: mbox.p)
    bool mbox_connected = false, mbox_step = false;
    unsigned long long mbox_epoch = 0;             // bumped by every upload: stamps of an earlier run can never match
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    PotParamsDev pot{};
    double sigma[PVD_MAX_ATOMS]{};
    double inv_mass[PVD_MAX_ATOMS]{};
};

static StepArgs make_args(pvd_sim *s, int do_branch)
{
    StepArgs a{};
    const int in = s->cur, out = s->cur ^ 1;
    a.xin = s->x[in].as<double>();
    a.xout = s->x[out].as<double>();
    a.vin = s->v[in].as<double>();
    a.vout = s->v[out].as<double>();
    a.who_in = s->who[in].as<int>();
    a.who_out = s->who[out].as<int>();
    a.w = s->w.as<double>();
    if (s->cfg.trial != PVD_TRIAL_NONE) {
        a.fin = s->f[in].as<double>(); a.fout = s->f[out].as<double>();
        a.psin = s->psi[in].as<double>(); a.psout = s->psi[out].as<double>();
        a.lkin = s->lk[in].as<double>(); a.lkout = s->lk[out].as<double>();
        if (s->cfg.imp_variant == PVD_IMP_EXCITED_STATE) { a.vsin = s->vs[in].as<double>(); a.vsout = s->vs[out].as<double>(); }
    }
    a.st = s->st.as<DevState>();
    a.err_accum = s->err_accum.as<unsigned>();
    a.status = s->status.as<unsigned long long>();
    a.part = s->part.as<WarpPartial>();
    a.tickets = s->tickets.as<unsigned>();
    a.ring = s->ring.as<pvd_step_stats>();
    a.ring_len = s->cfg.stats_ring;
    a.sums = s->sums_ext ? (double *)s->sums_ext : s->sums.as<double>();
    a.kill_idx = s->kill_idx.as<int>();
    a.hist = s->hist.as<unsigned>();
    a.cap = s->cap;
    a.n0 = s->cfg.num_walkers;
    a.dt = s->cfg.delta_t;
    a.alpha = s->cfg.alpha;
    a.lower = s->cfg.thresh_lower;
    a.upper = s->cfg.thresh_upper;
    a.seed = s->cfg.seed;
    a.parity = s->parity;
    a.do_branch = do_branch;
    a.world = s->cfg.world_size;
    a.rank = s->cfg.rank;
    a.ndim = s->cfg.ndim;
    a.nc = s->nc;
    a.flip = s->cfg.weighting == PVD_WEIGHT_DISCRETE ? 1 : 0;
    a.ticket_batch = s->ticket_batch;
    for (int r = 0; r < PVD_MAX_WORLD; ++r) a.mbox[r] = s->mbox_step ? s->peer_mbox[r] : nullptr;
    a.mbox_epoch = s->mbox_epoch;
    // ranks are only loosely synchronised by the host (checkpoint pickling, file systems): a slow peer is not a dead peer
    static const long long mbox_ticks = [] { const char *e = getenv("PVD_MBOX_TIMEOUT_S"); const double sec = e ? atof(e) : 120.0; return (long long)((sec > 0 ? sec : 120.0) * 2.0e9); }();
    a.mbox_timeout_ticks = mbox_ticks;
    static const bool fold_ok = getenv("PVD_NO_MBOX_FOLD") == nullptr;       // A/B switch
    a.mbox_fold = (fold_ok && s->mbox_step && s->cfg.weighting == PVD_WEIGHT_DISCRETE && s->cfg.trial == PVD_TRIAL_NONE) ? 1 : 0;
    for (int i = 0; i < PVD_MAX_ATOMS; ++i) a.sigma[i] = s->sigma[i];
    for (int c = 0; c < PVD_MAX_COMP; ++c) a.sigc[c] = s->sigma[(c / (s->cfg.ndim > 0 ? s->cfg.ndim : 1)) % PVD_MAX_ATOMS];
    a.pot = s->pot;
    return a;
}

static int run_ctl_reset(pvd_sim *s);
static int cont_enqueue_step(pvd_sim *, StepArgs &);
static int cont_enqueue_branch_only(pvd_sim *, StepArgs &, long long *src_out = nullptr);
static int imp_enqueue_step(pvd_sim *, StepArgs &, const double *);
static int imp_enqueue_move(pvd_sim *, StepArgs &, const double *);
static int imp_enqueue_branch(pvd_sim *, StepArgs &);
static int imp_initial_drift(pvd_sim *, long long first = 0, long long count = -1);
static int nn_enqueue_discrete_step(pvd_sim *, StepArgs &);

static int gather_flush(pvd_sim *s);
#define SIM_CHECK(s) PVD_REQUIRE((s) != nullptr, "NULL simulation handle")
// Every entry point that may look at the walker arrays goes through SIM_DEVICE: an open chain of deferred-compaction steps
// (pvd_gather.cuh) is materialised first, so nothing but pvd_sim_run / _run_mailbox / _step_injected ever sees the deferred form.
// SIM_DEVICE_RAW: entry points that read only the state copies / log ring, or continue the chain.
#define SIM_DEVICE_RAW(s) PVD_CUDA(cudaSetDevice((s)->cfg.device))
#define SIM_DEVICE(s)                                                \
    do {                                                             \
        PVD_CUDA(cudaSetDevice((s)->cfg.device));                    \
        if (int rc__ = gather_flush(s)) return rc__;                 \
    } while (0)

// launch the potential of the configured kind on resident SoA walkers -> v[cur]
static int launch_pot_soa(pvd_sim *s)
{
    const int g = s->grid;
    double *x = s->x[s->cur].as<double>(), *v = s->v[s->cur].as<double>();
    const DevState *st = s->st.as<DevState>();
    switch (s->cfg.potential) {
    case PVD_POT_H2O_PS: k_pot_soa<PotH2O><<<g, PVD_CTA, 0, s->stream>>>(x, st, s->parity, s->cap, v, s->pot); break;
    case PVD_POT_HARMONIC:
        if (s->nc == 1) k_pot_soa<PotHarm<1>><<<g, PVD_CTA, 0, s->stream>>>(x, st, s->parity, s->cap, v, s->pot);
        else if (s->nc == 3) k_pot_soa<PotHarm<3>><<<g, PVD_CTA, 0, s->stream>>>(x, st, s->parity, s->cap, v, s->pot);
        else return pvd_fail(PVD_E_ARG, "built-in harmonic potential supports 1 or 3 components");
        break;
    case PVD_POT_MORSE1D: k_pot_soa<PotMorse><<<g, PVD_CTA, 0, s->stream>>>(x, st, s->parity, s->cap, v, s->pot); break;
    case PVD_POT_NN_H4O2: return nn_launch(s->stream, x, 1, s->cap, st, s->parity, 0, s->cap, v, s->nn_w);
    default: return pvd_fail(PVD_E_STATE, "no built-in potential configured");
    }
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}

extern "C" {

int pvd_sim_create(const pvd_config *cfg, pvd_sim **out)
{
    PVD_REQUIRE(cfg && out, "NULL argument");
    PVD_REQUIRE(cfg->imp_variant >= PVD_IMP_STANDARD && cfg->imp_variant <= PVD_IMP_EXCITED_STATE, "unknown imp_variant");
    PVD_REQUIRE(cfg->imp_variant != PVD_IMP_EXCITED_STATE || (cfg->trial != PVD_TRIAL_NONE && cfg->ndim == 3), "excited-state importance sampling needs a trial wave function over 3-D atoms");
    PVD_REQUIRE(cfg->natoms >= 1 && cfg->natoms <= PVD_MAX_ATOMS && cfg->ndim >= 1 && cfg->ndim <= 3, "bad natoms/ndim");
    PVD_REQUIRE(cfg->num_walkers >= 1 && cfg->capacity >= 1 && cfg->delta_t > 0, "bad num_walkers/capacity/delta_t");
    PVD_REQUIRE(cfg->world_size >= 1 && cfg->world_size <= PVD_MAX_WORLD && cfg->rank >= 0 && cfg->rank < cfg->world_size, "bad rank/world_size");
    PVD_REQUIRE(cfg->weighting == PVD_WEIGHT_DISCRETE || cfg->weighting == PVD_WEIGHT_CONTINUOUS, "bad weighting");
    PVD_REQUIRE(cfg->stats_ring >= 1, "stats_ring must be >= 1");
    PVD_REQUIRE(cfg->rng_mode >= PVD_RNG_FP64 && cfg->rng_mode <= PVD_RNG_ZIGGURAT, "unknown rng_mode");
    PVD_CUDA(cudaSetDevice(cfg->device));
    if (int rc = ensure_device_ready()) return rc;
    pvd_sim *s = new pvd_sim();
    s->cfg = *cfg;
    s->nc = cfg->natoms * cfg->ndim;
    const int nc = s->nc;
    if (cfg->potential == PVD_POT_H2O_PS && nc != 9) { delete s; return pvd_fail(PVD_E_ARG, "PS water needs 3 atoms x 3 dims"); }
    if (cfg->potential == PVD_POT_NN_H4O2 && nc != 18) { delete s; return pvd_fail(PVD_E_ARG, "h4o2 NN needs 6 atoms x 3 dims"); }
    if (cfg->potential == PVD_POT_MORSE1D && nc != 1) { delete s; return pvd_fail(PVD_E_ARG, "Morse needs 1 component"); }
    s->cap = (cfg->capacity + 31) / 32 * 32;
    const long long cap = s->cap;
    const long long ntiles = (cap + PVD_TILE - 1) / PVD_TILE;
    for (int a = 0; a < cfg->natoms; ++a) {
        s->sigma[a] = sqrt(cfg->delta_t / cfg->masses[a]);      // pyvibdmc.py:199
        s->inv_mass[a] = 1.0 / cfg->masses[a];
    }
    for (int c = 0; c < PVD_MAX_COMP; ++c) s->pot.k[c] = cfg->pot_params[c];
    auto fail = [&](cudaError_t e) { std::string m = cudaGetErrorString(e); delete s; return pvd_fail(PVD_E_CUDA, "pvd_sim_create: " + m); };
    cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) return fail(e)
    TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->own_stream = true;
    const bool imp = cfg->trial != PVD_TRIAL_NONE;
    for (int b = 0; b < 2; ++b) {
        TRY(s->x[b].alloc((size_t)cap * nc * 8));
        TRY(s->v[b].alloc((size_t)cap * 8));
        TRY(s->who[b].alloc((size_t)cap * 4));
        if (imp) {
            TRY(s->f[b].alloc((size_t)cap * nc * 8));
            TRY(s->psi[b].alloc((size_t)cap * 8));
            TRY(s->lk[b].alloc((size_t)cap * 8));
            if (cfg->imp_variant == PVD_IMP_EXCITED_STATE) TRY(s->vs[b].alloc((size_t)cap * 8));
        }
    }
    if (cfg->weighting == PVD_WEIGHT_CONTINUOUS) {
        TRY(s->w.alloc((size_t)cap * 8));
        TRY(s->kill_idx.alloc((size_t)cap * 4));
        TRY(s->kill_mask.alloc(((size_t)cap / 32 + 16) * 4));
        TRY(s->hist.alloc(PVD_HIST_BINS * 4));
        TRY(cudaMemset(s->hist.p, 0, PVD_HIST_BINS * 4));
        TRY(s->cand.alloc((size_t)2 * cap * sizeof(ContCand)));       // candidates (fallback sort pads to a power of two)
        TRY(s->cand_sorted.alloc((size_t)cap * sizeof(ContCand)));
        TRY(s->bin_start.alloc(PVD_HIST_BINS * 4));
        TRY(cudaFuncSetAttribute(k_cont_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PVD_RANK_SMEM));
        TRY(s->bin_fill.alloc(PVD_HIST_BINS * 4));
        TRY(cudaMemset(s->bin_fill.p, 0, PVD_HIST_BINS * 4));
        TRY(s->cont_queue.alloc((size_t)2 * cap * sizeof(ContCand)));
        TRY(s->copy_dst.alloc((size_t)2 * cap * 4));
        TRY(s->copy_src.alloc((size_t)2 * cap * 4));
        TRY(s->cont_root.alloc((size_t)cap * 4));
        TRY(s->cont_skip.alloc((size_t)cap));
        TRY(s->cont_work.alloc(sizeof(ContWork)));
        TRY(cudaMemset(s->cont_work.p, 0, sizeof(ContWork)));
    }
    TRY(s->acc_count.alloc(8));
    TRY(cudaMemset(s->acc_count.p, 0, 8));
    TRY(s->st.alloc(2 * sizeof(DevState)));
    TRY(cudaMemset(s->st.p, 0, 2 * sizeof(DevState)));
    TRY(s->err_accum.alloc(4));
    TRY(cudaMemset(s->err_accum.p, 0, 4));
    s->ntiles_cap = ntiles;
    TRY(s->status.alloc((size_t)2 * ntiles * 8));         // two parities (the resident kernel overlaps consecutive steps)
    TRY(cudaMemset(s->status.p, 0, (size_t)2 * ntiles * 8));
    TRY(s->run_ctl.alloc(sizeof(RunCtl)));
    TRY(s->run_ready.alloc((size_t)2 * ntiles * 4));
    k_run_ctl_init<<<1, 64>>>(s->run_ctl.as<RunCtl>());
    TRY(cudaGetLastError());
    // persistent-style grid: enough CTAs to cover the capacity, at most 8 per SM
    s->grid = grid_for(cap, PVD_CTA, 2);
    {
        // heavy tiles (fused H2O step: ~9 us of fp64 work per tile) take one tile per ticket; light ones share a ticket
        const bool heavy = cfg->potential == PVD_POT_H2O_PS && cfg->trial == PVD_TRIAL_NONE && cfg->weighting == PVD_WEIGHT_DISCRETE;
        const char *e = getenv(heavy ? "PVD_TICKET_BATCH_HEAVY" : "PVD_TICKET_BATCH_LIGHT");
        s->ticket_batch = e ? atoi(e) : 1;
        if (s->ticket_batch < 1 || s->ticket_batch > 2) s->ticket_batch = 1;   // > 2 serialises the look-back chain (measured: 250x slower)
    }
    s->grid_light = grid_for(cap, PVD_CTA, 4);       // kernels without a heavy potential: more warps per SM hide the memory latency
    TRY(s->part.alloc((size_t)s->grid_light * sizeof(WarpPartial)));
    TRY(s->tickets.alloc(2 * PVD_WARPS * PVD_TICKET_STRIDE * 4));
    TRY(cudaMemset(s->tickets.p, 0, 2 * PVD_WARPS * PVD_TICKET_STRIDE * 4));
    TRY(s->ring.alloc((size_t)cfg->stats_ring * sizeof(pvd_step_stats)));
    TRY(cudaMemset(s->ring.p, 0, (size_t)cfg->stats_ring * sizeof(pvd_step_stats)));
    TRY(s->sums.alloc(PVD_NSUMS * 8));
    TRY(cudaMemset(s->sums.p, 0, PVD_NSUMS * 8));
    TRY(s->sigma_dev.alloc(PVD_MAX_ATOMS * 8));
    TRY(cudaMemcpy(s->sigma_dev.p, s->sigma, PVD_MAX_ATOMS * 8, cudaMemcpyHostToDevice));
    s->nn_grid = grid_for(cap, NN_TILE, 3);
    // staging for host <-> device transposes (uploads, downloads): allocated once, not inside a timed upload / download
    TRY(s->stage.alloc((size_t)cap * nc * 8));
    TRY(cudaEventCreate(&s->ev0));
    TRY(cudaEventCreate(&s->ev1));
    if (cfg->weighting == PVD_WEIGHT_DISCRETE && cfg->trial == PVD_TRIAL_NONE &&
        (cfg->potential == PVD_POT_H2O_PS || cfg->potential == PVD_POT_HARMONIC || cfg->potential == PVD_POT_MORSE1D)) {
        // buffers of the deferred-compaction step (pvd_gather.cuh), so that the first segment does not allocate
        for (int b = 0; b < 2; ++b) {
            TRY(s->g_cnt[b].alloc((size_t)cap * 4));
            TRY(s->g_tincl[b].alloc((size_t)ntiles * 4));
            TRY(s->g_cbase[b].alloc((size_t)(PVD_GATHER_MAX_GRID + 2) * 4));
            TRY(s->g_meta[b].alloc(sizeof(GatherMeta)));
        }
        TRY(s->g_seg.alloc(16));
    }
#undef TRY
    *out = s;
    return PVD_OK;
}

int pvd_sim_destroy(pvd_sim *s)
{
    if (!s) return PVD_OK;
    cudaSetDevice(s->cfg.device);
    cudaDeviceSynchronize();
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    for (int r = 0; r < PVD_MAX_WORLD; ++r)
        if (s->peer_mbox[r] && s->peer_mbox[r] != s->mbox.p) cudaIpcCloseMemHandle(s->peer_mbox[r]);
    if (s->snap_host) cudaFreeHost(s->snap_host);
    if (s->dump_host) cudaFreeHost(s->dump_host);
    if (s->dump_ready) cudaEventDestroy(s->dump_ready);
    if (s->dump_done) cudaEventDestroy(s->dump_done);
    if (s->snap_stream) cudaStreamDestroy(s->snap_stream);
    if (s->snap_ready) cudaEventDestroy(s->snap_ready);
    if (s->snap_done) cudaEventDestroy(s->snap_done);
    if (s->nn_w.packed) cudaFree(s->nn_w.packed);
    if (s->nn_w.images) cudaFree(s->nn_w.images);
    if (s->nn_w.vecs) cudaFree(s->nn_w.vecs);
    delete s;
    return PVD_OK;
}

int pvd_sim_set_stream(pvd_sim *s, void *cuda_stream)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    if (s->own_stream && s->stream) { cudaStreamDestroy(s->stream); s->own_stream = false; }
    s->stream = (cudaStream_t)cuda_stream;      // 0 selects the legacy default stream
    return PVD_OK;
}

int pvd_sim_sync(pvd_sim *s)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

static int sim_init_sums(pvd_sim *s)
{
    const double *w = s->cfg.weighting == PVD_WEIGHT_CONTINUOUS ? s->w.as<double>() : nullptr;
    k_init_sums<<<1, 1024, 0, s->stream>>>(s->v[s->cur].as<double>(), w, s->st.as<DevState>(), s->parity, 0,
                                           s->sums_ext ? (double *)s->sums_ext : s->sums.as<double>(),
                                           s->cfg.world_size, s->cfg.rank);
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}

int pvd_sim_init_finalize(pvd_sim *s)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    k_init_finalize<<<1, 32, 0, s->stream>>>(s->st.as<DevState>(), s->parity,
                                             s->sums_ext ? (double *)s->sums_ext : s->sums.as<double>(), s->cfg.alpha, s->cfg.num_walkers,
                                             s->cfg.delta_t);
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}

int pvd_sim_upload(pvd_sim *s, const double *xyz, int64_t n, const double *w)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    PVD_REQUIRE(xyz && n >= 1 && n <= s->cap, "pvd_sim_upload: n must be in [1, capacity]");
    s->deferred = false;                                      // whatever chain was open belonged to the ensemble that is replaced
    const int nc = s->nc;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    PVD_CUDA(s->stage.alloc((size_t)n * nc * 8));
    PVD_CUDA(cudaMemcpyAsync(s->stage.p, xyz, (size_t)n * nc * 8, cudaMemcpyHostToDevice, s->stream));
    s->cur = 0;
    s->parity = 0;
    ++s->mbox_epoch;
    if (int rc = run_ctl_reset(s)) return rc;
    k_aos_to_soa<<<grid_for(n * nc, 256, 16), 256, 0, s->stream>>>(s->stage.as<double>(), s->x[0].as<double>(), n, nc, s->cap);
    PVD_CHECK_LAUNCH();
    DevState h[2];
    memset(h, 0, sizeof(h));
    h[0].n = n;
    h[0].dt_eff = s->cfg.delta_t;
    h[1] = h[0];
    PVD_CUDA(cudaMemcpyAsync(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaMemsetAsync(s->err_accum.p, 0, 4, s->stream));
    if (s->cfg.weighting == PVD_WEIGHT_CONTINUOUS) {
        if (w) PVD_CUDA(cudaMemcpyAsync(s->w.p, w, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
        else {
            k_fill_double<<<grid_for(n, 256, 16), 256, 0, s->stream>>>(s->w.as<double>(), n, 1.0);
            PVD_CHECK_LAUNCH();
        }
    }
    PVD_CUDA(cudaStreamSynchronize(s->stream));   // h[] and xyz staging are host stack / caller memory
    s->n_uploaded = n;
    s->uploaded = true;
    s->ext_moved = false;
    if (s->cfg.trial == PVD_TRIAL_EXTERNAL) return PVD_OK;       // caller continues with pvd_sim_imp_ext_init (drift terms from the host)
    if (s->cfg.potential == PVD_POT_EXTERNAL) return PVD_OK;     // caller continues with pvd_sim_set_pots
    if (s->cfg.trial != PVD_TRIAL_NONE) {
        // first-step exception with importance sampling: E_L = V + local kinetic (pyvibdmc.py:763-767)
        if (int rc = imp_initial_drift(s)) return rc;
    } else {
        if (int rc = launch_pot_soa(s)) return rc;
    }
    if (int rc = sim_init_sums(s)) return rc;
    if (s->cfg.world_size == 1) return pvd_sim_init_finalize(s);
    return PVD_OK;
}

int pvd_sim_set_pots(pvd_sim *s, const double *v, int64_t n)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && v && n == s->n_uploaded, "pvd_sim_set_pots: call after pvd_sim_upload with the same n");
    PVD_CUDA(cudaMemcpyAsync(s->v[s->cur].p, v, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    if (int rc = sim_init_sums(s)) return rc;
    if (s->cfg.world_size == 1) return pvd_sim_init_finalize(s);
    return PVD_OK;
}

int pvd_sim_sums_ptr(pvd_sim *s, void **device_ptr)
{
    SIM_CHECK(s);
    PVD_REQUIRE(device_ptr, "NULL argument");
    *device_ptr = s->sums_ext ? s->sums_ext : s->sums.p;
    return PVD_OK;
}

int pvd_sim_set_sums_ptr(pvd_sim *s, void *device_ptr)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    if (device_ptr) PVD_CUDA(cudaMemcpy(device_ptr, s->sums_ext ? s->sums_ext : s->sums.p, PVD_NSUMS * 8, cudaMemcpyDeviceToDevice));
    s->sums_ext = device_ptr;
    return PVD_OK;
}

}  // extern "C"

// one fused step on the stream; `inj` pointers may be null
static int enqueue_step(pvd_sim *s, int do_branch, const double *inj_disp, const double *inj_u, const double *inj_um)
{
    StepArgs a = make_args(s, do_branch);
    a.inj_disp = inj_disp;
    a.inj_u = inj_u;
    const int g = s->grid;
    const bool cont = s->cfg.weighting == PVD_WEIGHT_CONTINUOUS;
    const bool fast = s->cfg.rng_mode == PVD_RNG_FAST;
    if (s->cfg.trial != PVD_TRIAL_NONE) {
        if (int rc = imp_enqueue_step(s, a, inj_um)) return rc;
    } else if (cont) {
        if (int rc = cont_enqueue_step(s, a)) return rc;
    } else {
#ifndef PVD_STEP_KERNEL
#define PVD_STEP_KERNEL k_step_discrete
#endif
        // Programmatic dependent launch: the next step's CTAs may become resident (and stage their tables) while this
        // step drains; they read nothing of the walker state before pdl_wait() (see k_step_discrete).
#define LAUNCH_DISC_R(POT, R) PVD_CUDA(launch_pdl(PVD_STEP_KERNEL<POT, R>, dim3((unsigned)gp), dim3(PVD_CTA), 0, s->stream, a))
#define LAUNCH_DISC(POT)                                                                            \
    do {                                                                                            \
        const int gp = POT::MIN_CTAS >= 4 ? s->grid_light : g;                                      \
        if (fast) LAUNCH_DISC_R(POT, PVD_RNG_FAST);                                                 \
        else if (s->cfg.rng_mode == PVD_RNG_ZIGGURAT) LAUNCH_DISC_R(POT, PVD_RNG_ZIGGURAT);         \
        else LAUNCH_DISC_R(POT, PVD_RNG_FP64);                                                      \
    } while (0)
        switch (s->cfg.potential) {
        case PVD_POT_H2O_PS: LAUNCH_DISC(PotH2O); break;
        case PVD_POT_HARMONIC:
            if (s->nc == 1) LAUNCH_DISC(PotHarm<1>);
            else if (s->nc == 3) LAUNCH_DISC(PotHarm<3>);
            else return pvd_fail(PVD_E_ARG, "built-in harmonic potential supports 1 or 3 components");
            break;
        case PVD_POT_MORSE1D: LAUNCH_DISC(PotMorse); break;
        case PVD_POT_NN_H4O2:
            if (int rc = nn_enqueue_discrete_step(s, a)) return rc;
            break;
        default: return pvd_fail(PVD_E_STATE, "pvd_sim_run needs a built-in potential (use the ext_* calls for PVD_POT_EXTERNAL)");
        }
#undef LAUNCH_DISC
#undef LAUNCH_DISC_R
        PVD_CHECK_LAUNCH();
        s->cur ^= 1;
    }
    s->parity ^= 1;
    return PVD_OK;
}

// ---------------------------------------------------------------- resident multi-step launch (pvd_run.cuh)
typedef void (*run_kernel_t)(const StepArgs, const RunArgs);
struct RunVariant {
    run_kernel_t kern = nullptr;
    int tpb = 0, id = 0;
    size_t smem = 0;
};

template <class POT, int TPB, int MINB>
static RunVariant run_variant_rng(int rng_mode, bool multi = false)
{
    RunVariant v;
    v.tpb = TPB;
    v.id = (TPB * 8 + MINB) * 2 + (multi ? 1 : 0);
    v.smem = (size_t)(TPB / 32) * sizeof(RunWarpMem<POT::NC>);
    if (multi) {
        if (rng_mode == PVD_RNG_FAST) v.kern = k_run_discrete<POT, PVD_RNG_FAST, TPB, MINB, true>;
        else if (rng_mode == PVD_RNG_ZIGGURAT) v.kern = k_run_discrete<POT, PVD_RNG_ZIGGURAT, TPB, MINB, true>;
        else v.kern = k_run_discrete<POT, PVD_RNG_FP64, TPB, MINB, true>;
    } else {
        if (rng_mode == PVD_RNG_FAST) v.kern = k_run_discrete<POT, PVD_RNG_FAST, TPB, MINB, false>;
        else if (rng_mode == PVD_RNG_ZIGGURAT) v.kern = k_run_discrete<POT, PVD_RNG_ZIGGURAT, TPB, MINB, false>;
        else v.kern = k_run_discrete<POT, PVD_RNG_FP64, TPB, MINB, false>;
    }
    return v;
}

// the kernel variant for this simulation (kern == nullptr when the resident loop does not cover it)
static RunVariant run_variant_for(const pvd_sim *s)
{
    static const bool off = getenv("PVD_NO_RESIDENT") != nullptr;           // A/B switch: one launch per time step
    if (off || !s->resident || s->cfg.weighting != PVD_WEIGHT_DISCRETE || s->cfg.trial != PVD_TRIAL_NONE) return RunVariant{};
    // Measured on a B200 (H2O, us per step, one launch per step vs resident): 1 000 walkers 12.4 / 10.8, 20 000 15.8 / 12.4,
    // 100 000 25.1 / 20.4, 400 000 51.7 / 48.4, 1 000 000 103.5 / 110: the resident kernel removes the fixed cost of a step
    // but spends more per tile (counters, fences), so above ~600 000 walkers per GPU the step-per-launch kernel is used
    // unless resident stepping is forced (pvd_sim_set_resident(s, 2) / PVD_RUN_MAX_WALKERS).
    // With the deferred-compaction step (pvd_gather.cuh: 34.1 / 63.8 / 90.2 us at 200 000 / 600 000 / 1 000 000 walkers) the
    // cross-over moves down to ~300 000 walkers per GPU.
    static const long long max_walkers = [] { const char *e = getenv("PVD_RUN_MAX_WALKERS"); return e ? atoll(e) : (getenv("PVD_NO_GATHER") ? 600000ll : 300000ll); }();
    if (s->resident_mode != 2 && (s->cfg.num_walkers + s->cfg.world_size - 1) / s->cfg.world_size > max_walkers) return RunVariant{};
    const int rng = s->cfg.rng_mode;
    switch (s->cfg.potential) {
    case PVD_POT_H2O_PS: {
        // occupancy variants of the same kernel (A/B: PVD_RUN_VARIANT = 2562 | 2563 | 3842 | 2564 = threads per CTA, CTAs per SM)
        static const int want = [] { const char *e = getenv("PVD_RUN_VARIANT"); return e ? atoi(e) : PVD_RUN_VARIANT_H2O; }();
        const bool multi = s->cfg.world_size > 1;
        if (!multi) {                                   // (the A/B occupancy variants exist for one GPU only)
            if (want == 2562) return run_variant_rng<PotH2O, 256, 2>(rng);
            if (want == 3842) return run_variant_rng<PotH2O, 384, 2>(rng);
            if (want == 2564) return run_variant_rng<PotH2O, 256, 4>(rng);
        }
        return run_variant_rng<PotH2O, 256, 3>(rng, multi);
    }
    case PVD_POT_HARMONIC:
        if (s->nc == 1) return run_variant_rng<PotHarm<1>, 256, 4>(rng, s->cfg.world_size > 1);
        if (s->nc == 3) return run_variant_rng<PotHarm<3>, 256, 4>(rng, s->cfg.world_size > 1);
        return RunVariant{};
    case PVD_POT_MORSE1D: return run_variant_rng<PotMorse, 256, 4>(rng, s->cfg.world_size > 1);
    default: return RunVariant{};
    }
}

// nsteps time steps in one cooperative launch; the caller has checked run_variant_for()
static int enqueue_run(pvd_sim *s, long long nsteps, int do_branch)
{
    if (nsteps <= 0) return PVD_OK;
    PVD_REQUIRE(nsteps < (1ll << 30), "enqueue_run: at most 2^30 time steps per launch");
    PVD_REQUIRE(s->cap * s->nc < (1ll << 31), "enqueue_run: 32-bit element offsets need components x capacity < 2^31");
    const RunVariant rv = run_variant_for(s);
    PVD_REQUIRE(rv.kern != nullptr, "enqueue_run: configuration not covered by the resident kernel");
    if (s->run_grid == 0 || s->run_minb != rv.id) {
        PVD_CUDA(cudaFuncSetAttribute(rv.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rv.smem));
        PVD_CUDA(cudaFuncSetAttribute(rv.kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        PVD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rv.kern, rv.tpb, rv.smem));
        PVD_REQUIRE(occ >= 1, "resident kernel does not fit an SM");
        s->run_grid = grid_for(s->cap, rv.tpb, occ);       // never more CTAs than can be co-resident
        s->run_occ = occ;
        s->run_minb = rv.id;
        if (getenv("PVD_RUN_VERBOSE")) fprintf(stderr, "[pvd] resident kernel variant %d: %d CTAs/SM, grid %d x %d threads, %zu B dynamic smem\n", rv.id, occ, s->run_grid, rv.tpb, rv.smem);
    }
    StepArgs a = make_args(s, do_branch);
    RunArgs ra{};
    ra.ctl = s->run_ctl.as<RunCtl>();
    ra.ready = s->run_ready.as<unsigned>();
    ra.ntiles_cap = s->ntiles_cap;
    ra.nsteps = nsteps;
    ra.seq0 = s->run_seq;
    static const int dyn = getenv("PVD_RUN_STATIC") ? 0 : 1;      // A/B switch: tickets (default) or static interleaved tiles
    ra.dynamic = dyn;
    static const int single = getenv("PVD_RUN_SINGLE") ? 1 : 0;   // A/B switch: one launch per time step of the same kernel
    if (single) {
        // the kernel boundary orders a step's scattered walkers before the next step's loads
        cudaLaunchAttribute ats[1];
        ats[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        ats[0].val.programmaticStreamSerializationAllowed = 1;
        for (long long k = 0; k < nsteps; ++k) {
            StepArgs ak = make_args(s, do_branch);
            ra.nsteps = 1;
            ra.seq0 = s->run_seq;
            ra.single = 1;
            cudaLaunchConfig_t lc{};
            lc.gridDim = dim3((unsigned)s->run_grid);
            lc.blockDim = dim3((unsigned)rv.tpb);
            lc.dynamicSmemBytes = rv.smem;
            lc.stream = s->stream;
            lc.attrs = ats;
            lc.numAttrs = 1;
            PVD_CUDA(cudaLaunchKernelEx(&lc, rv.kern, ak, ra));
            PVD_CHECK_LAUNCH();
            s->run_seq += 1;
            s->cur ^= 1;
            s->parity ^= 1;
        }
        return PVD_OK;
    }
    // output-tile counters start from zero in every launch (whatever ran in between: injected steps, rebalancing, uploads)
    PVD_CUDA(cudaMemsetAsync(s->run_ready.p, 0, (size_t)2 * s->ntiles_cap * 4, s->stream));
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)s->run_grid);
    lc.blockDim = dim3((unsigned)rv.tpb);
    lc.dynamicSmemBytes = rv.smem;
    lc.stream = s->stream;
    lc.attrs = at;
    lc.numAttrs = 1;
    PVD_CUDA(cudaLaunchKernelEx(&lc, rv.kern, a, ra));
    PVD_CHECK_LAUNCH();
    s->run_seq += (unsigned long long)nsteps;
    if (nsteps & 1) { s->cur ^= 1; s->parity ^= 1; }
    return PVD_OK;
}

static int run_ctl_reset(pvd_sim *s)
{
    k_run_ctl_init<<<1, 64, 0, s->stream>>>(s->run_ctl.as<RunCtl>());
    PVD_CHECK_LAUNCH();
    s->run_seq = 0;
    return PVD_OK;
}

// ---------------------------------------------------------------- steps with deferred compaction (pvd_gather.cuh)
typedef void (*gather_kernel_t)(const StepArgs, const GatherArgs);
struct GatherVariant {
    gather_kernel_t kern = nullptr;
    int minb = 0, id = 0;
};
template <class POT, int MINB>
static GatherVariant gather_variant_rng(int rng_mode, int pot_id, bool multi)
{
    GatherVariant v;
    v.minb = MINB;
    v.id = (pot_id * 64 + rng_mode * 8 + MINB) * 2 + (multi ? 1 : 0);
    if (multi) {
        if (rng_mode == PVD_RNG_FAST) v.kern = k_step_gather<POT, PVD_RNG_FAST, MINB, true>;
        else if (rng_mode == PVD_RNG_ZIGGURAT) v.kern = k_step_gather<POT, PVD_RNG_ZIGGURAT, MINB, true>;
        else v.kern = k_step_gather<POT, PVD_RNG_FP64, MINB, true>;
    } else {
        if (rng_mode == PVD_RNG_FAST) v.kern = k_step_gather<POT, PVD_RNG_FAST, MINB, false>;
        else if (rng_mode == PVD_RNG_ZIGGURAT) v.kern = k_step_gather<POT, PVD_RNG_ZIGGURAT, MINB, false>;
        else v.kern = k_step_gather<POT, PVD_RNG_FP64, MINB, false>;
    }
    return v;
}

// the kernel for this simulation (kern == nullptr: not covered, or not selected)
static GatherVariant gather_variant_for(const pvd_sim *s, bool forced = false)
{
    static const bool off = getenv("PVD_NO_GATHER") != nullptr;             // A/B switch: compaction inside the step (k_step_discrete)
    if (s->cfg.weighting != PVD_WEIGHT_DISCRETE || s->cfg.trial != PVD_TRIAL_NONE) return GatherVariant{};
    if (!forced && (off || s->resident_mode == 0)) return GatherVariant{};
    const int rng = s->cfg.rng_mode;
    switch (s->cfg.potential) {
    case PVD_POT_H2O_PS: {
        static const int want = [] { const char *e = getenv("PVD_GATHER_MINB"); return e ? atoi(e) : 2; }();
        const bool multi = s->cfg.world_size > 1;
        if (want == 3 && !multi) return gather_variant_rng<PotH2O, 3>(rng, 1, false);      // (A/B occupancy variant: one GPU only)
        return gather_variant_rng<PotH2O, 2>(rng, 1, multi);
    }
    case PVD_POT_HARMONIC:
        if (s->nc == 1) return gather_variant_rng<PotHarm<1>, 4>(rng, 2, s->cfg.world_size > 1);
        if (s->nc == 3) return gather_variant_rng<PotHarm<3>, 4>(rng, 3, s->cfg.world_size > 1);
        return GatherVariant{};
    case PVD_POT_MORSE1D: return gather_variant_rng<PotMorse, 4>(rng, 4, s->cfg.world_size > 1);
    default: return GatherVariant{};
    }
}

// nsteps time steps, each ONE launch of k_step_gather, then k_gather_materialise: on return (in stream order) the
// simulation is in exactly the state nsteps launches of k_step_discrete would have left (compacted ensemble in x[cur]).
static int enqueue_gather_segment(pvd_sim *s, long long nsteps, int do_branch, const GatherVariant &gv,
                                  const double *inj_disp = nullptr, const double *inj_u = nullptr)
{
    if (nsteps <= 0) return PVD_OK;
    PVD_REQUIRE(s->cap < (1ll << 31), "gather steps: capacity must be below 2^31 walkers per shard");
    if (s->gather_grid == 0 || s->gather_id != gv.id) {
        const int grid = grid_for(s->cap, PVD_CTA, gv.minb);
        PVD_REQUIRE(grid <= PVD_GATHER_MAX_GRID, "gather steps: grid too large");
        const long long tpc_max = (s->ntiles_cap + grid - 1) / grid;
        PVD_REQUIRE(tpc_max <= PVD_GATHER_MAX_TPC, "gather steps: shard too large for one chunk per CTA");
        s->gather_smem = (size_t)((tpc_max < PVD_GATHER_SUBT ? tpc_max : PVD_GATHER_SUBT) * PVD_TILE + tpc_max + grid + 2) * 4;
        PVD_CUDA(cudaFuncSetAttribute(gv.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->gather_smem));
        if (const char *e = getenv("PVD_GATHER_CARVEOUT")) PVD_CUDA(cudaFuncSetAttribute(gv.kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
        PVD_CUDA(cudaFuncSetAttribute(k_gather_materialise, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((PVD_GATHER_MAX_GRID + 2) * 4)));
        for (int b = 0; b < 2; ++b) {
            PVD_CUDA(s->g_cnt[b].alloc((size_t)s->cap * 4));
            PVD_CUDA(s->g_tincl[b].alloc((size_t)s->ntiles_cap * 4));
            PVD_CUDA(s->g_cbase[b].alloc((size_t)(PVD_GATHER_MAX_GRID + 2) * 4));
            PVD_CUDA(s->g_meta[b].alloc(sizeof(GatherMeta)));
        }
        PVD_CUDA(s->g_seg.alloc(16));
        s->gather_grid = grid;
        s->gather_id = gv.id;
    }
    if (!s->deferred) { s->defer_buf0 = s->cur; s->defer_steps = 0; }
    for (long long k = 0; k < nsteps; ++k) {
        StepArgs a = make_args(s, do_branch);
        a.inj_disp = inj_disp;
        a.inj_u = inj_u;
        GatherArgs g{};
        const int in = s->cur, out = s->cur ^ 1;
        g.cnt_in = s->g_cnt[in].as<int>(); g.cnt_out = s->g_cnt[out].as<int>();
        g.tincl_in = s->g_tincl[in].as<int>(); g.tincl_out = s->g_tincl[out].as<int>();
        g.cbase_in = s->g_cbase[in].as<int>(); g.cbase_out = s->g_cbase[out].as<int>();
        g.meta_in = s->g_meta[in].as<GatherMeta>(); g.meta_out = s->g_meta[out].as<GatherMeta>();
        g.deferred_in = s->defer_steps > 0 ? 1 : 0;
        static const int stagger = [] { const char *e = getenv("PVD_GATHER_STAGGER_NS"); return e ? atoi(e) : 0; }();
        g.stagger_ns = stagger;
        g.seg_step0 = s->g_seg.as<long long>();     // the step counter the segment starts from stays on the device (no host synchronisation)
        PVD_CUDA(launch_pdl(gv.kern, dim3((unsigned)s->gather_grid), dim3(PVD_CTA), s->gather_smem, s->stream, a, g));
        PVD_CHECK_LAUNCH();
        s->cur ^= 1;
        s->parity ^= 1;
        s->defer_steps += 1;
        s->deferred = true;
    }
    return PVD_OK;
}

// Closes an open chain of deferred-compaction steps: one k_gather_materialise leaves the ensemble compacted in x[cur], exactly
// what the same number of k_step_discrete launches would have left (also when the run died inside the chain).
static int gather_flush(pvd_sim *s)
{
    if (!s->deferred) return PVD_OK;
    MaterialiseArgs m{};
    for (int b = 0; b < 2; ++b) {
        m.x[b] = s->x[b].as<double>(); m.v[b] = s->v[b].as<double>(); m.who[b] = s->who[b].as<int>();
        m.cnt[b] = s->g_cnt[b].as<int>(); m.tincl[b] = s->g_tincl[b].as<int>(); m.cbase[b] = s->g_cbase[b].as<int>();
        m.meta[b] = s->g_meta[b].as<GatherMeta>();
    }
    m.st = s->st.as<DevState>();
    m.cap = s->cap;
    m.nc = s->nc;
    m.parity_end = s->parity;
    m.buf0 = s->defer_buf0;
    m.seg_step0 = s->g_seg.as<long long>();
    const int mg = grid_for(s->cap, PVD_CTA, 8);
    PVD_CUDA(launch_pdl(k_gather_materialise, dim3((unsigned)mg), dim3(PVD_CTA), (size_t)(s->gather_grid + 2) * 4, s->stream, m));
    PVD_CHECK_LAUNCH();
    s->cur ^= 1;
    s->deferred = false;
    s->defer_steps = 0;
    return PVD_OK;
}

extern "C" {

int pvd_sim_run(pvd_sim *s, int64_t nsteps, int32_t branch_every)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    PVD_REQUIRE(s->uploaded, "pvd_sim_run: upload walkers first");
    PVD_REQUIRE(s->cfg.world_size == 1, "pvd_sim_run is single-shard; use step_local/step_finalize for multi-GPU");
    PVD_REQUIRE(branch_every >= 1, "branch_every must be >= 1");
    PVD_CUDA(cudaEventRecord(s->ev0, s->stream));
    const int do_branch = (branch_every == 1) ? 1 : -branch_every;       // negative: the kernel decides from its step counter
    const GatherVariant gv = gather_variant_for(s, s->resident_mode == 3);
    const bool use_resident = s->resident_mode != 3 && run_variant_for(s).kern;
    if (use_resident || !gv.kern)
        if (int rc = gather_flush(s)) return rc;              // (an open chain of deferred-compaction steps ends here)
    if (use_resident) {
        // discrete weighting with a built-in potential, small ensembles: the whole segment is ONE resident launch
        if (int rc = enqueue_run(s, nsteps, do_branch)) return rc;
    } else if (gv.kern) {
        // large ensembles: one launch per step, compaction deferred to the next step's gather
        if (int rc = enqueue_gather_segment(s, nsteps, do_branch, gv)) return rc;
    } else {
        for (int64_t k = 0; k < nsteps; ++k)
            if (int rc = enqueue_step(s, do_branch, nullptr, nullptr, nullptr)) return rc;
    }
    PVD_CUDA(cudaEventRecord(s->ev1, s->stream));
    return PVD_OK;
}

int pvd_sim_set_resident(pvd_sim *s, int32_t enable)
{
    SIM_CHECK(s);
    PVD_REQUIRE(enable >= 0 && enable <= 3, "pvd_sim_set_resident: 0 one self-compacting launch per step, 1 automatic (by ensemble size), 2 resident kernel always, 3 deferred-compaction steps always");
    s->resident = enable != 0 && enable != 3;
    s->resident_mode = enable;
    return PVD_OK;
}

int pvd_sim_last_run_ms(pvd_sim *s, double *ms)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    PVD_REQUIRE(ms, "NULL argument");
    PVD_CUDA(cudaEventSynchronize(s->ev1));
    float f = 0;
    PVD_CUDA(cudaEventElapsedTime(&f, s->ev0, s->ev1));
    *ms = f;
    return PVD_OK;
}

int pvd_sim_step_injected(pvd_sim *s, const double *disp, const double *u_branch, const double *u_metro)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    if (!(s->resident_mode == 3 && gather_variant_for(s, true).kern))
        if (int rc = gather_flush(s)) return rc;              // (mode 3 continues an open chain: the injected steps exercise the pull)
    PVD_REQUIRE(s->uploaded && disp, "pvd_sim_step_injected: upload first / disp required");
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    const int nc = s->nc;
    if (!s->inj_disp.p) {
        PVD_CUDA(s->inj_disp.alloc((size_t)s->cap * nc * 8));
        PVD_CUDA(s->inj_u.alloc((size_t)s->cap * 8));
        PVD_CUDA(s->inj_um.alloc((size_t)s->cap * 8));
    }
    PVD_CUDA(s->stage.alloc((size_t)n * nc * 8));
    PVD_CUDA(cudaMemcpyAsync(s->stage.p, disp, (size_t)n * nc * 8, cudaMemcpyHostToDevice, s->stream));
    k_aos_to_soa<<<grid_for(n * nc, 256, 16), 256, 0, s->stream>>>(s->stage.as<double>(), s->inj_disp.as<double>(), n, nc, s->cap);
    PVD_CHECK_LAUNCH();
    if (u_branch) PVD_CUDA(cudaMemcpyAsync(s->inj_u.p, u_branch, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    if (u_metro) PVD_CUDA(cudaMemcpyAsync(s->inj_um.p, u_metro, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    const GatherVariant gv = s->resident_mode == 3 ? gather_variant_for(s, true) : GatherVariant{};
    if (gv.kern) {
        // the deferred-compaction step on injected numbers (a one-step segment): parity tests of the kernel the large runs use
        if (int rc = enqueue_gather_segment(s, 1, 1, gv, s->inj_disp.as<double>(), u_branch ? s->inj_u.as<double>() : nullptr)) return rc;
    } else if (int rc = enqueue_step(s, 1, s->inj_disp.as<double>(), u_branch ? s->inj_u.as<double>() : nullptr,
                                     u_metro ? s->inj_um.as<double>() : nullptr))
        return rc;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

// ---- external potential: move on the device, V from the caller, weight/branch on the device
// moves the walkers (Philox normals) and leaves them as AoS (n, atoms, dims) in the staging buffer on the device
static int ext_move_impl(pvd_sim *s, long long *n_out)
{
    PVD_REQUIRE(s->cfg.trial == PVD_TRIAL_NONE, "external potentials with built-in importance sampling are not supported");
    const int nc = s->nc;
    const int g = s->grid;
    double *x = s->x[s->cur].as<double>();
    if (s->cfg.rng_mode == PVD_RNG_FAST)
        k_displace_soa<PVD_RNG_FAST><<<g, 256, 0, s->stream>>>(x, s->st.as<DevState>(), s->parity, 0, 0, s->cap, nc, s->cfg.ndim,
                                                              s->cfg.seed, nullptr, nullptr, s->sigma_dev.as<double>(), nullptr);
    else if (s->cfg.rng_mode == PVD_RNG_ZIGGURAT)
        k_displace_soa<PVD_RNG_ZIGGURAT><<<g, 256, 0, s->stream>>>(x, s->st.as<DevState>(), s->parity, 0, 0, s->cap, nc, s->cfg.ndim,
                                                                  s->cfg.seed, nullptr, nullptr, s->sigma_dev.as<double>(), nullptr);
    else
        k_displace_soa<PVD_RNG_FP64><<<g, 256, 0, s->stream>>>(x, s->st.as<DevState>(), s->parity, 0, 0, s->cap, nc, s->cfg.ndim,
                                                              s->cfg.seed, nullptr, nullptr, s->sigma_dev.as<double>(), nullptr);
    PVD_CHECK_LAUNCH();
    DevState h[2];
    PVD_CUDA(cudaMemcpyAsync(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    const long long n = h[s->parity].n;
    PVD_CUDA(s->stage.alloc((size_t)n * nc * 8));
    k_soa_to_aos<<<grid_for(n * nc, 256, 16), 256, 0, s->stream>>>(x, s->stage.as<double>(), n, nc, s->cap);
    PVD_CHECK_LAUNCH();
    *n_out = n;
    s->ext_moved = true;
    return PVD_OK;
}

int pvd_sim_ext_move(pvd_sim *s, double *xyz_out, int64_t *n_out)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && xyz_out && n_out, "pvd_sim_ext_move: bad arguments");
    long long n = 0;
    if (int rc = ext_move_impl(s, &n)) return rc;
    PVD_CUDA(cudaMemcpyAsync(xyz_out, s->stage.p, (size_t)n * s->nc * 8, cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    *n_out = n;
    return PVD_OK;
}

/* device-tensor plug-in: the moved walkers stay in HBM.  *xyz_dev: float64 (n, atoms, dims), C order, on the simulation's device, valid
 * until the next call on this handle; every kernel that produced it has completed when the call returns. */
int pvd_sim_ext_move_device(pvd_sim *s, void **xyz_dev, int64_t *n_out)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && xyz_dev && n_out, "pvd_sim_ext_move_device: bad arguments");
    long long n = 0;
    if (int rc = ext_move_impl(s, &n)) return rc;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    *xyz_dev = s->stage.p;
    *n_out = n;
    return PVD_OK;
}

/* the walkers as they are (no move), AoS on the device: first-step energies of a device-tensor potential */
int pvd_sim_coords_device(pvd_sim *s, void **xyz_dev, int64_t *n_out)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && xyz_dev && n_out, "pvd_sim_coords_device: bad arguments");
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    PVD_CUDA(s->stage.alloc((size_t)n * s->nc * 8));
    k_soa_to_aos<<<grid_for(n * s->nc, 256, 16), 256, 0, s->stream>>>(s->x[s->cur].as<double>(), s->stage.as<double>(), n, s->nc, s->cap);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    *xyz_dev = s->stage.p;
    *n_out = n;
    return PVD_OK;
}

int pvd_sim_ext_finish(pvd_sim *s, const double *v, int64_t n, int32_t do_branch)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->ext_moved && v, "pvd_sim_ext_finish: call pvd_sim_ext_move first");
    PVD_CUDA(cudaMemcpyAsync(s->v[s->cur].p, v, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    StepArgs a = make_args(s, do_branch);
    if (s->cfg.weighting == PVD_WEIGHT_CONTINUOUS) {
        if (int rc = cont_enqueue_branch_only(s, a)) return rc;
    } else {
        k_branch_discrete<<<s->grid_light, PVD_CTA, 0, s->stream>>>(a);
        PVD_CHECK_LAUNCH();
        s->cur ^= 1;
    }
    s->parity ^= 1;
    s->ext_moved = false;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

/* v_dev: float64 (n) on the simulation's device (the plug-in's result, e.g. a torch / CuPy array: whatever stream wrote it is
 * synchronised here, device-wide, before the energies are used) */
int pvd_sim_ext_finish_device(pvd_sim *s, const void *v_dev, int64_t n, int32_t do_branch)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->ext_moved && v_dev, "pvd_sim_ext_finish_device: call pvd_sim_ext_move_device first");
    PVD_CUDA(cudaDeviceSynchronize());
    PVD_CUDA(cudaMemcpyAsync(s->v[s->cur].p, v_dev, (size_t)n * 8, cudaMemcpyDeviceToDevice, s->stream));
    StepArgs a = make_args(s, do_branch);
    if (s->cfg.weighting == PVD_WEIGHT_CONTINUOUS) {
        if (int rc = cont_enqueue_branch_only(s, a)) return rc;
    } else {
        k_branch_discrete<<<s->grid_light, PVD_CTA, 0, s->stream>>>(a);
        PVD_CHECK_LAUNCH();
        s->cur ^= 1;
    }
    s->parity ^= 1;
    s->ext_moved = false;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

int pvd_sim_set_pots_device(pvd_sim *s, const void *v_dev, int64_t n)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && v_dev && n == s->n_uploaded, "pvd_sim_set_pots_device: call after pvd_sim_upload with the same n");
    PVD_CUDA(cudaDeviceSynchronize());
    PVD_CUDA(cudaMemcpyAsync(s->v[s->cur].p, v_dev, (size_t)n * 8, cudaMemcpyDeviceToDevice, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    if (int rc = sim_init_sums(s)) return rc;
    if (s->cfg.world_size == 1) return pvd_sim_init_finalize(s);
    return PVD_OK;
}

// ---- multi-GPU split step
int pvd_sim_step_local(pvd_sim *s, int32_t do_branch)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded, "upload walkers first");
    // identical to a fused step; with world_size > 1 the last CTA only publishes the shard's sums
    return enqueue_step(s, do_branch, nullptr, nullptr, nullptr);
}

int pvd_sim_step_finalize(pvd_sim *s)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    // enqueue_step already flipped the parity: finalise the step that read parity^1
    StepArgs a = make_args(s, 1);
    a.parity = s->parity ^ 1;
    k_finalize<<<1, 32, 0, s->stream>>>(a, s->cfg.weighting == PVD_WEIGHT_CONTINUOUS ? 1 : 0);
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}

// ---- multi-GPU step with the NVLink mailbox collective (see k_finalize_mailbox)
int pvd_sim_mailbox_handle(pvd_sim *s, void *handle64)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(handle64, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!s->mbox.p) {
        const size_t bytes = (size_t)2 * PVD_MAX_WORLD * PVD_MBOX_STRIDE * 16;      // 16-byte entries {value, stamp ^ value}
        PVD_CUDA(s->mbox.alloc(bytes));
        PVD_CUDA(cudaMemset(s->mbox.p, 0, bytes));
    }
    cudaIpcMemHandle_t h;
    PVD_CUDA(cudaIpcGetMemHandle(&h, s->mbox.p));
    memcpy(handle64, &h, 64);
    return PVD_OK;
}

int pvd_sim_mailbox_connect(pvd_sim *s, const void *handles, int32_t n)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(handles && n == s->cfg.world_size && n <= PVD_MAX_WORLD && s->mbox.p, "pvd_sim_mailbox_connect: one handle per rank, after pvd_sim_mailbox_handle");
    for (int r = 0; r < n; ++r) {
        if (r == s->cfg.rank) { s->peer_mbox[r] = s->mbox.as<double>(); continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * r, 64);
        void *p = nullptr;
        PVD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_mbox[r] = (double *)p;
    }
    s->mbox_connected = true;
    return PVD_OK;
}

int pvd_sim_run_mailbox(pvd_sim *s, int64_t nsteps, int32_t branch_every)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    PVD_REQUIRE(s->uploaded && s->mbox_connected, "pvd_sim_run_mailbox: upload walkers and connect the mailboxes first");
    PVD_REQUIRE(s->cfg.trial == PVD_TRIAL_NONE, "importance sampling needs two exchanges per step: use the split step with an all-reduce");
    PVD_REQUIRE(branch_every >= 1, "branch_every must be >= 1");
    const int cont = s->cfg.weighting == PVD_WEIGHT_CONTINUOUS ? 1 : 0;
    PVD_CUDA(cudaEventRecord(s->ev0, s->stream));
    const GatherVariant gv = gather_variant_for(s, s->resident_mode == 3);
    if (!(!(s->resident_mode != 3 && run_variant_for(s).kern) && gv.kern))
        if (int rc = gather_flush(s)) return rc;
    if (!(s->resident_mode != 3 && run_variant_for(s).kern) && gv.kern) {
        // large shards: deferred-compaction steps, the last CTA of each exchanges the sums through the mailboxes
        s->mbox_step = true;
        const int rc = enqueue_gather_segment(s, nsteps, branch_every == 1 ? 1 : -branch_every, gv);
        s->mbox_step = false;
        if (rc) return rc;
        PVD_CUDA(cudaEventRecord(s->ev1, s->stream));
        return PVD_OK;
    }
    if (run_variant_for(s).kern) {
        // resident launch: the warp that finalises a step exchanges the sums while the others already move the next step
        s->mbox_step = true;
        const int rc = enqueue_run(s, nsteps, branch_every == 1 ? 1 : -branch_every);
        s->mbox_step = false;
        if (rc) return rc;
        PVD_CUDA(cudaEventRecord(s->ev1, s->stream));
        return PVD_OK;
    }
    for (int64_t k = 0; k < nsteps; ++k) {
        s->mbox_step = true;
        const int rc = enqueue_step(s, branch_every == 1 ? 1 : -branch_every, nullptr, nullptr, nullptr);
        if (rc) { s->mbox_step = false; return rc; }
        StepArgs a = make_args(s, 1);
        s->mbox_step = false;
        if (a.mbox_fold) continue;                 // discrete steps: the step kernel's last CTA has collected and finalised
        a.parity = s->parity ^ 1;                  // enqueue_step already flipped the parity
        PVD_CUDA(launch_pdl(k_finalize_mailbox, dim3(1), dim3(32), 0, s->stream, a, cont));
        PVD_CHECK_LAUNCH();
    }
    PVD_CUDA(cudaEventRecord(s->ev1, s->stream));
    return PVD_OK;
}

// ---- descendant weighting
int pvd_sim_dw_begin(pvd_sim *s, int64_t global_offset)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded, "upload walkers first");
    // _parent = copy(coords), _parent_wts = copy(w), who_from = arange(N), _desc_wt = True (pyvibdmc.py:739-745)
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    s->parent_n = n;
    PVD_CUDA(s->parent_x.alloc((size_t)s->cap * s->nc * 8));
    PVD_CUDA(cudaMemcpyAsync(s->parent_x.p, s->x[s->cur].p, (size_t)s->cap * s->nc * 8, cudaMemcpyDeviceToDevice, s->stream));
    if (s->cfg.weighting == PVD_WEIGHT_CONTINUOUS) {
        PVD_CUDA(s->parent_w.alloc((size_t)s->cap * 8));
        PVD_CUDA(cudaMemcpyAsync(s->parent_w.p, s->w.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, s->stream));
    }
    k_iota_int<<<grid_for(n, 256, 16), 256, 0, s->stream>>>(s->who[s->cur].as<int>(), n, (int)global_offset);
    PVD_CHECK_LAUNCH();
    h[0].dw_active = 1;
    h[1].dw_active = 1;
    PVD_CUDA(cudaMemcpyAsync(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

/* Re-opens a descendant-weighting window from a checkpoint written inside one (pyvibdmc.py:299-338 keeps _who_from, _parent,
 * _parent_wts in the pickle): who_from of the current walkers and the parent ensemble go back to the device. */
int pvd_sim_dw_resume(pvd_sim *s, const int64_t *who_from, int64_t n, const double *parent_xyz, const double *parent_w, int64_t n_parent)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && who_from && parent_xyz && n == s->n_uploaded && n_parent >= 1 && n_parent <= s->cap, "pvd_sim_dw_resume: bad arguments");
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<int> w32((size_t)n);
    for (int64_t i = 0; i < n; ++i) w32[(size_t)i] = (int)who_from[i];
    PVD_CUDA(cudaMemcpy(s->who[s->cur].p, w32.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    PVD_CUDA(s->parent_x.alloc((size_t)s->cap * s->nc * 8));
    PVD_CUDA(s->stage.alloc((size_t)n_parent * s->nc * 8));
    PVD_CUDA(cudaMemcpy(s->stage.p, parent_xyz, (size_t)n_parent * s->nc * 8, cudaMemcpyHostToDevice));
    k_aos_to_soa<<<grid_for(n_parent * s->nc, 256, 16), 256, 0, s->stream>>>(s->stage.as<double>(), s->parent_x.as<double>(), n_parent, s->nc, s->cap);
    PVD_CHECK_LAUNCH();
    if (s->cfg.weighting == PVD_WEIGHT_CONTINUOUS && parent_w) {
        PVD_CUDA(s->parent_w.alloc((size_t)s->cap * 8));
        PVD_CUDA(cudaMemcpy(s->parent_w.p, parent_w, (size_t)n_parent * 8, cudaMemcpyHostToDevice));
    }
    s->parent_n = n_parent;
    DevState h[2];
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    h[0].dw_active = 1;
    h[1].dw_active = 1;
    PVD_CUDA(cudaMemcpy(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice));
    return PVD_OK;
}

static int dw_collect(pvd_sim *s, double *desc_wts, int64_t n_parent, bool close_window);
int pvd_sim_dw_end(pvd_sim *s, double *desc_wts, int64_t n_parent) { return dw_collect(s, desc_wts, n_parent, true); }
/* calc_desc_wts without closing the window (DEBUG_save_desc_wt_tracker, pyvibdmc.py:849-852) */
int pvd_sim_dw_peek(pvd_sim *s, double *desc_wts, int64_t n_parent) { return dw_collect(s, desc_wts, n_parent, false); }

/* DEBUG_mass_change (pyvibdmc.py:749-753): new masses -> new displacement widths sigma = sqrt(dt / m); like the reference,
 * nothing else (the importance-sampling 1/m factors keep their initial values) */
int pvd_sim_set_masses(pvd_sim *s, const double *masses, int32_t natoms)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(masses && natoms == s->cfg.natoms, "pvd_sim_set_masses: one mass per atom");
    for (int a = 0; a < natoms; ++a) {
        PVD_REQUIRE(masses[a] > 0.0, "masses must be positive");
        s->cfg.masses[a] = masses[a];
        s->sigma[a] = sqrt(s->cfg.delta_t / masses[a]);
    }
    PVD_CUDA(cudaMemcpyAsync(s->sigma_dev.p, s->sigma, PVD_MAX_ATOMS * 8, cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

static int dw_collect(pvd_sim *s, double *desc_wts, int64_t n_parent, bool close_window)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(desc_wts && n_parent >= 1, "bad arguments");
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    PVD_REQUIRE(h[s->parity].dw_active, "no descendant-weighting window is open");
    PVD_CUDA(s->stage2.alloc((size_t)n_parent * 8));
    PVD_CUDA(cudaMemsetAsync(s->stage2.p, 0, (size_t)n_parent * 8, s->stream));
    const double *w = s->cfg.weighting == PVD_WEIGHT_CONTINUOUS ? s->w.as<double>() : nullptr;
    k_desc_wts<<<s->grid, 256, 0, s->stream>>>(s->who[s->cur].as<int>(), w, s->st.as<DevState>(), s->parity, 0, 0, n_parent,
                                               s->stage2.as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpyAsync(desc_wts, s->stage2.p, (size_t)n_parent * 8, cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    if (close_window) {
        h[0].dw_active = 0;
        h[1].dw_active = 0;
        PVD_CUDA(cudaMemcpy(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice));
    }
    return PVD_OK;
}

__global__ void k_set_dw_active(DevState *st, int v)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { st[0].dw_active = v; st[1].dw_active = v; }
}

/* Closes the descendant-weighting window WITHOUT stopping the loop (wave-function dumps of the reference, pyvibdmc.py:856-872):
 * the descendant weights are counted and the parent ensemble is copied in stream order, then travel to pinned host memory on the
 * side stream while the compute stream goes on with the next time steps; pvd_sim_dw_end_wait hands them over. */
int pvd_sim_dw_end_begin(pvd_sim *s, int64_t n_parent)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(n_parent >= 1 && n_parent == s->parent_n && s->parent_x.p, "pvd_sim_dw_end_begin: no window open for this parent count");
    PVD_REQUIRE(!s->dump_pending, "a wave-function dump is already in flight: call pvd_sim_dw_end_wait first");
    const int nc = s->nc;
    const bool cont = s->cfg.weighting == PVD_WEIGHT_CONTINUOUS;
    if (!s->snap_stream) PVD_CUDA(cudaStreamCreateWithFlags(&s->snap_stream, cudaStreamNonBlocking));
    if (!s->dump_ready) {
        PVD_CUDA(cudaEventCreateWithFlags(&s->dump_ready, cudaEventDisableTiming));
        PVD_CUDA(cudaEventCreateWithFlags(&s->dump_done, cudaEventDisableTiming));
    }
    PVD_CUDA(s->dump_desc.alloc((size_t)n_parent * 8));
    PVD_CUDA(s->dump_parent.alloc((size_t)n_parent * nc * 8));
    if (cont) PVD_CUDA(s->dump_pw.alloc((size_t)n_parent * 8));
    const size_t need = (size_t)n_parent * (nc + 2) * 8;
    if (need > s->dump_host_bytes) {
        if (s->dump_host) PVD_CUDA(cudaFreeHost(s->dump_host));
        PVD_CUDA(cudaMallocHost(&s->dump_host, need));
        s->dump_host_bytes = need;
    }
    PVD_CUDA(cudaMemsetAsync(s->dump_desc.p, 0, (size_t)n_parent * 8, s->stream));
    k_desc_wts<<<s->grid, 256, 0, s->stream>>>(s->who[s->cur].as<int>(), cont ? s->w.as<double>() : nullptr, s->st.as<DevState>(), s->parity, 0, 0,
                                               n_parent, s->dump_desc.as<double>());
    PVD_CHECK_LAUNCH();
    k_soa_to_aos<<<grid_for(n_parent * nc, 256, 16), 256, 0, s->stream>>>(s->parent_x.as<double>(), s->dump_parent.as<double>(), n_parent, nc, s->cap);
    PVD_CHECK_LAUNCH();
    if (cont) PVD_CUDA(cudaMemcpyAsync(s->dump_pw.p, s->parent_w.p, (size_t)n_parent * 8, cudaMemcpyDeviceToDevice, s->stream));
    k_set_dw_active<<<1, 32, 0, s->stream>>>(s->st.as<DevState>(), 0);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaEventRecord(s->dump_ready, s->stream));
    PVD_CUDA(cudaStreamWaitEvent(s->snap_stream, s->dump_ready, 0));
    char *h = (char *)s->dump_host;
    PVD_CUDA(cudaMemcpyAsync(h, s->dump_desc.p, (size_t)n_parent * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    PVD_CUDA(cudaMemcpyAsync(h + (size_t)n_parent * 8, s->dump_parent.p, (size_t)n_parent * nc * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    if (cont) PVD_CUDA(cudaMemcpyAsync(h + (size_t)n_parent * (nc + 1) * 8, s->dump_pw.p, (size_t)n_parent * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    PVD_CUDA(cudaEventRecord(s->dump_done, s->snap_stream));
    s->dump_pending = true;
    s->dump_n = n_parent;
    return PVD_OK;
}

/* blocks only until the side-stream copies have landed (the compute stream is not synchronised) */
int pvd_sim_dw_end_wait(pvd_sim *s, double *desc_wts, double *parent_xyz, double *parent_w, int64_t n_parent)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->dump_pending && n_parent == s->dump_n && desc_wts && parent_xyz, "pvd_sim_dw_end_wait: no dump in flight for this parent count");
    PVD_CUDA(cudaEventSynchronize(s->dump_done));
    const char *h = (const char *)s->dump_host;
    const size_t n = (size_t)n_parent;
    memcpy(desc_wts, h, n * 8);
    memcpy(parent_xyz, h + n * 8, n * s->nc * 8);
    if (parent_w && s->cfg.weighting == PVD_WEIGHT_CONTINUOUS) memcpy(parent_w, h + n * (s->nc + 1) * 8, n * 8);
    s->dump_pending = false;
    return PVD_OK;
}

int pvd_sim_dw_parent(pvd_sim *s, double *xyz, double *w, int64_t *n_parent)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(n_parent, "NULL argument");
    *n_parent = s->parent_n;
    if (!xyz) return PVD_OK;
    PVD_REQUIRE(s->parent_x.p, "no parent ensemble stored");
    const long long n = s->parent_n;
    PVD_CUDA(s->stage.alloc((size_t)n * s->nc * 8));
    k_soa_to_aos<<<grid_for(n * s->nc, 256, 16), 256, 0, s->stream>>>(s->parent_x.as<double>(), s->stage.as<double>(), n, s->nc, s->cap);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpyAsync(xyz, s->stage.p, (size_t)n * s->nc * 8, cudaMemcpyDeviceToHost, s->stream));
    if (w && s->parent_w.p) PVD_CUDA(cudaMemcpyAsync(w, s->parent_w.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

// ---- queries
int pvd_sim_state(pvd_sim *s, int64_t *n, double *vref, int64_t *step, int32_t *err)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const DevState &c = h[s->parity];
    if (n) *n = c.n;
    if (vref) *vref = c.vref;
    if (step) *step = c.step;
    if (err) *err = (int32_t)c.err;
    if (c.err & (PVD_ERR_WEIGHT | PVD_ERR_POP | PVD_ERR_EMPTY)) return pvd_fail(PVD_E_MASSIVE, PVD_MASSIVE_MSG);
    if (c.err & PVD_ERR_CAPACITY) return pvd_fail(PVD_E_MASSIVE, std::string(PVD_MASSIVE_MSG) + " (shard capacity exceeded)");
    if (c.err & PVD_ERR_COMM) return pvd_fail(PVD_E_STATE, "a peer's per-step message did not arrive in time (NVLink mailbox exchange; PVD_MBOX_TIMEOUT_S, default 120 s): a rank died or stalled");
    return PVD_OK;
}

int pvd_sim_download(pvd_sim *s, double *xyz, double *pots, double *w, int64_t *who_from, int64_t capacity, int64_t *n_out)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    if (n_out) *n_out = n;
    PVD_REQUIRE(capacity >= n || (!xyz && !pots && !w && !who_from), "pvd_sim_download: host buffers too small");
    const int nc = s->nc;
    if (h[s->parity].err) s->cur = h[s->parity].buf;      // dead run: the failing step's input buffer is the valid one
    if (xyz) {
        PVD_CUDA(s->stage.alloc((size_t)n * nc * 8));
        k_soa_to_aos<<<grid_for(n * nc, 256, 16), 256, 0, s->stream>>>(s->x[s->cur].as<double>(), s->stage.as<double>(), n, nc, s->cap);
        PVD_CHECK_LAUNCH();
        PVD_CUDA(cudaMemcpyAsync(xyz, s->stage.p, (size_t)n * nc * 8, cudaMemcpyDeviceToHost, s->stream));
    }
    if (pots) PVD_CUDA(cudaMemcpyAsync(pots, s->v[s->cur].p, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    if (w && s->w.p) PVD_CUDA(cudaMemcpyAsync(w, s->w.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    if (who_from) {
        PVD_CUDA(s->stage2.alloc((size_t)n * 8));
        k_int_to_i64<<<grid_for(n, 256, 16), 256, 0, s->stream>>>(s->who[s->cur].as<int>(), s->stage2.as<long long>(), n);
        PVD_CHECK_LAUNCH();
        PVD_CUDA(cudaMemcpyAsync(who_from, s->stage2.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    }
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

// ---- asynchronous snapshot: the walkers as they are NOW (in stream order) travel to the host on a side stream while the
// compute stream goes on with the next time steps (checkpoints and dumps of the reference, pyvibdmc.py:729-736, 861-872)
int pvd_sim_snapshot_begin(pvd_sim *s)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded, "pvd_sim_snapshot_begin: upload walkers first");
    PVD_REQUIRE(!s->snap_pending, "a snapshot is already in flight: call pvd_sim_snapshot_wait first");
    const int nc = s->nc;
    const size_t cap = (size_t)s->cap;
    const bool cont = s->w.p != nullptr;
    if (!s->snap_stream) PVD_CUDA(cudaStreamCreateWithFlags(&s->snap_stream, cudaStreamNonBlocking));     // (shared with the wave-function dumps)
    if (!s->snap_host) {
        PVD_CUDA(cudaEventCreateWithFlags(&s->snap_ready, cudaEventDisableTiming));
        PVD_CUDA(cudaEventCreateWithFlags(&s->snap_done, cudaEventDisableTiming));
        PVD_CUDA(s->snap_x.alloc(cap * nc * 8));
        PVD_CUDA(s->snap_v.alloc(cap * 8));
        PVD_CUDA(s->snap_who.alloc(cap * 4));
        if (cont) PVD_CUDA(s->snap_w.alloc(cap * 8));
        PVD_CUDA(s->snap_state.alloc(2 * sizeof(DevState)));
        s->snap_host_bytes = cap * nc * 8 + cap * 8 * 2 + cap * 4 + 2 * sizeof(DevState);
        PVD_CUDA(cudaMallocHost(&s->snap_host, s->snap_host_bytes));
    }
    // device-side copies in stream order (the whole capacity: the population is only known on the device)
    k_soa_to_aos<<<grid_for((long long)cap * nc, 256, 16), 256, 0, s->stream>>>(s->x[s->cur].as<double>(), s->snap_x.as<double>(), (long long)cap, nc, s->cap);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpyAsync(s->snap_v.p, s->v[s->cur].p, cap * 8, cudaMemcpyDeviceToDevice, s->stream));
    PVD_CUDA(cudaMemcpyAsync(s->snap_who.p, s->who[s->cur].p, cap * 4, cudaMemcpyDeviceToDevice, s->stream));
    if (cont) PVD_CUDA(cudaMemcpyAsync(s->snap_w.p, s->w.p, cap * 8, cudaMemcpyDeviceToDevice, s->stream));
    PVD_CUDA(cudaMemcpyAsync(s->snap_state.p, s->st.p, 2 * sizeof(DevState), cudaMemcpyDeviceToDevice, s->stream));
    PVD_CUDA(cudaEventRecord(s->snap_ready, s->stream));
    // device -> pinned host on the side stream
    PVD_CUDA(cudaStreamWaitEvent(s->snap_stream, s->snap_ready, 0));
    char *h = (char *)s->snap_host;
    PVD_CUDA(cudaMemcpyAsync(h, s->snap_x.p, cap * nc * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    h += cap * nc * 8;
    PVD_CUDA(cudaMemcpyAsync(h, s->snap_v.p, cap * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    h += cap * 8;
    if (cont) PVD_CUDA(cudaMemcpyAsync(h, s->snap_w.p, cap * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    h += cap * 8;
    PVD_CUDA(cudaMemcpyAsync(h, s->snap_who.p, cap * 4, cudaMemcpyDeviceToHost, s->snap_stream));
    h += cap * 4;
    PVD_CUDA(cudaMemcpyAsync(h, s->snap_state.p, 2 * sizeof(DevState), cudaMemcpyDeviceToHost, s->snap_stream));
    PVD_CUDA(cudaEventRecord(s->snap_done, s->snap_stream));
    s->snap_pending = true;
    s->snap_parity = s->parity;
    return PVD_OK;
}

// blocks only until the side-stream copy has landed (the compute stream is not synchronised); any output may be NULL
int pvd_sim_snapshot_wait(pvd_sim *s, double *xyz, double *pots, double *w, int64_t *who_from, int64_t capacity, int64_t *n_out,
                          double *vref_out)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->snap_pending, "no snapshot in flight");
    PVD_CUDA(cudaEventSynchronize(s->snap_done));
    const int nc = s->nc;
    const size_t cap = (size_t)s->cap;
    const char *h = (const char *)s->snap_host;
    const DevState *st = reinterpret_cast<const DevState *>(h + cap * nc * 8 + cap * 8 * 2 + cap * 4);
    const long long n = st[s->snap_parity].n;
    if (n_out) *n_out = n;
    if (vref_out) *vref_out = st[s->snap_parity].vref;
    if (!xyz && !pots && !w && !who_from) return PVD_OK;          // size query: the snapshot stays pending
    PVD_REQUIRE(capacity >= n, "pvd_sim_snapshot_wait: host buffers too small");
    if (xyz) memcpy(xyz, h, (size_t)n * nc * 8);
    if (pots) memcpy(pots, h + cap * nc * 8, (size_t)n * 8);
    if (w && s->w.p) memcpy(w, h + cap * nc * 8 + cap * 8, (size_t)n * 8);
    if (who_from) {
        const int *src = reinterpret_cast<const int *>(h + cap * nc * 8 + cap * 8 * 2);
        for (long long i = 0; i < n; ++i) who_from[i] = src[i];
    }
    s->snap_pending = false;
    return PVD_OK;
}

int pvd_sim_stats(pvd_sim *s, int64_t first_step, int64_t count, pvd_step_stats *out)
{
    SIM_CHECK(s);
    SIM_DEVICE_RAW(s);
    PVD_REQUIRE(out && count >= 0 && count <= s->cfg.stats_ring && first_step >= 0, "pvd_sim_stats: bad range");
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    const long long L = s->cfg.stats_ring;
    long long done = 0;
    while (done < count) {
        const long long pos = (first_step + done) % L;
        const long long chunk = (count - done < L - pos) ? count - done : L - pos;
        PVD_CUDA(cudaMemcpy(out + done, s->ring.as<pvd_step_stats>() + pos, (size_t)chunk * sizeof(pvd_step_stats), cudaMemcpyDeviceToHost));
        done += chunk;
    }
    return PVD_OK;
}

// ---- stand-alone discrete branching (injection mode) built on the branch-only kernel
int pvd_branch_discrete(const double *v, int64_t n, double vref, double dt, const double *u, int64_t n0, int32_t *counts,
                        int64_t *idx, int64_t idx_capacity, int64_t *stats3)
{
    PVD_REQUIRE(v && u && n >= 1 && n0 >= 1 && stats3, "pvd_branch_discrete: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    const long long cap = ((idx_capacity > n ? idx_capacity : n) + 31) / 32 * 32;
    const long long ntiles = (n + PVD_TILE - 1) / PVD_TILE;
    DevBuf dv, du, dst, derr, dstatus, dpart, dring, dsums, dcounts, didx, dtick;
    const int g = grid_for(n, PVD_CTA, 8);
    PVD_CUDA(dv.alloc((size_t)n * 8)); PVD_CUDA(du.alloc((size_t)n * 8));
    PVD_CUDA(dst.alloc(2 * sizeof(DevState))); PVD_CUDA(derr.alloc(4));
    PVD_CUDA(dstatus.alloc((size_t)ntiles * 8)); PVD_CUDA(dpart.alloc((size_t)g * sizeof(WarpPartial)));
    PVD_CUDA(dtick.alloc(2 * PVD_WARPS * PVD_TICKET_STRIDE * 4));
    PVD_CUDA(cudaMemset(dtick.p, 0, 2 * PVD_WARPS * PVD_TICKET_STRIDE * 4));
    PVD_CUDA(dring.alloc(sizeof(pvd_step_stats))); PVD_CUDA(dsums.alloc(PVD_NSUMS * 8));
    PVD_CUDA(dcounts.alloc((size_t)n * 4)); PVD_CUDA(didx.alloc((size_t)cap * 8));
    PVD_CUDA(cudaMemcpy(dv.p, v, (size_t)n * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(cudaMemcpy(du.p, u, (size_t)n * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(cudaMemset(dstatus.p, 0, (size_t)ntiles * 8));
    PVD_CUDA(cudaMemset(derr.p, 0, 4));
    DevState h[2];
    memset(h, 0, sizeof(h));
    h[0].n = n; h[0].vref = vref; h[0].dt_eff = dt;
    PVD_CUDA(cudaMemcpy(dst.p, h, sizeof(h), cudaMemcpyHostToDevice));
    StepArgs a{};
    a.vin = dv.as<double>();
    a.st = dst.as<DevState>(); a.err_accum = derr.as<unsigned>(); a.status = dstatus.as<unsigned long long>();
    a.part = dpart.as<WarpPartial>(); a.tickets = dtick.as<unsigned>(); a.ring = dring.as<pvd_step_stats>(); a.ring_len = 1; a.sums = dsums.as<double>();
    a.inj_u = du.as<double>(); a.counts_out = dcounts.as<int>(); a.idx_out = didx.as<long long>();
    a.cap = cap; a.n0 = n0; a.dt = dt; a.alpha = 1.0 / (2.0 * dt); a.parity = 0; a.do_branch = 1; a.world = 1; a.rank = 0; a.nc = 0; a.ndim = 1; a.ticket_batch = 1;
    EventPair ev;
    PVD_CUDA(ev.init());
    PVD_CUDA(cudaEventRecord(ev.a));
    k_branch_discrete<<<g, PVD_CTA>>>(a);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaEventRecord(ev.b));
    PVD_CUDA(cudaMemcpy(h, dst.p, sizeof(h), cudaMemcpyDeviceToHost));
    pvd_step_stats r;
    PVD_CUDA(cudaMemcpy(&r, dring.p, sizeof(r), cudaMemcpyDeviceToHost));
    float ms = 0;
    PVD_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    g_last_kernel_ms = ms;
    stats3[0] = r.births; stats3[1] = r.deaths; stats3[2] = (int64_t)r.pop;
    if (counts) PVD_CUDA(cudaMemcpy(counts, dcounts.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (h[1].err) return pvd_fail(PVD_E_MASSIVE, PVD_MASSIVE_MSG);
    if (idx) {
        PVD_REQUIRE(idx_capacity >= h[1].n, "pvd_branch_discrete: idx buffer too small");
        PVD_CUDA(cudaMemcpy(idx, didx.p, (size_t)h[1].n * 8, cudaMemcpyDeviceToHost));
    }
    return PVD_OK;
}

int pvd_calc_vref(const double *v, const double *w, int64_t n, int64_t n0, double alpha, double *vref)
{
    PVD_REQUIRE(v && vref && n >= 1 && n0 >= 1, "pvd_calc_vref: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    DevBuf dv, dw, dst, dsums;
    PVD_CUDA(dv.alloc((size_t)n * 8));
    PVD_CUDA(cudaMemcpy(dv.p, v, (size_t)n * 8, cudaMemcpyHostToDevice));
    if (w) { PVD_CUDA(dw.alloc((size_t)n * 8)); PVD_CUDA(cudaMemcpy(dw.p, w, (size_t)n * 8, cudaMemcpyHostToDevice)); }
    PVD_CUDA(dst.alloc(2 * sizeof(DevState)));
    PVD_CUDA(dsums.alloc(PVD_NSUMS * 8));
    k_init_sums<<<1, 1024>>>(dv.as<double>(), w ? dw.as<double>() : nullptr, nullptr, 0, n, dsums.as<double>(), 1, 0);
    PVD_CHECK_LAUNCH();
    k_init_finalize<<<1, 32>>>(dst.as<DevState>(), 0, dsums.as<double>(), alpha, n0, 1.0);
    PVD_CHECK_LAUNCH();
    DevState h;
    PVD_CUDA(cudaMemcpy(&h, dst.p, sizeof(h), cudaMemcpyDeviceToHost));
    *vref = h.vref;
    return PVD_OK;
}

int pvd_desc_wts(const int64_t *who_from, const double *w, int64_t n, int64_t n_parent, double *out)
{
    PVD_REQUIRE(who_from && out && n >= 0 && n_parent >= 1, "pvd_desc_wts: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    std::vector<int> who32((size_t)n);
    for (int64_t i = 0; i < n; ++i) who32[(size_t)i] = (int)who_from[i];
    DevBuf dwho, dw, dout;
    PVD_CUDA(dwho.alloc((size_t)n * 4));
    PVD_CUDA(cudaMemcpy(dwho.p, who32.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    if (w) { PVD_CUDA(dw.alloc((size_t)n * 8)); PVD_CUDA(cudaMemcpy(dw.p, w, (size_t)n * 8, cudaMemcpyHostToDevice)); }
    PVD_CUDA(dout.alloc((size_t)n_parent * 8));
    PVD_CUDA(cudaMemset(dout.p, 0, (size_t)n_parent * 8));
    k_desc_wts<<<grid_for(n, 256, 16), 256>>>(dwho.as<int>(), w ? dw.as<double>() : nullptr, nullptr, 0, n, 0, n_parent, dout.as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpy(out, dout.p, (size_t)n_parent * 8, cudaMemcpyDeviceToHost));
    return PVD_OK;
}

}  // extern "C"

#include "pvd_api_tail.inl"

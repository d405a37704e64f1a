// Importance sampling with a USER trial wave function (PVD_TRIAL_EXTERNAL): the plug-in contract of the reference --
// ImpSampManager.call_trial / call_derivs (imp_samp_manager.py:92-139, 197-224) through ImpSamp.drift (imp_samp.py:21-27) --
// is honoured once per time step on the host; everything else of imp_move_randomly (pyvibdmc.py:549-612) stays on the GPU:
//   k_impx_propose : displaced = coords + disps + (1/m) f_x dt        (:556-559, 593; Philox + fp64 Box-Muller like k_imp_move)
//   host           : f_y, psi_2, psi_sec_der_disp = impsamp.drift(displaced)                                   (:595)
//   k_impx_accept  : ImpSamp.metropolis (imp_samp.py:29-47) for any (atoms x dims), accept where met_nums > u (:606-612),
//                    local kinetic energy -1/2 sum (1/m) psi''/psi (imp_samp.py:49-53) of the accepted walkers, dt_eff (:603)
//   then V (built-in kernel or the user's getpot), E_L = V + T_L (:807-809) and k_branch_discrete carrying f_x, psi, T_L.
// Run-time number of components (<= PVD_MAX_COMP): the shapes (N, A, 3) and (N, 1, 1) of the reference and anything else.
#pragma once
#include "pvd_impsamp.cuh"

// standard normals of walker i, component c (same numbers as walker_normals<NC, MODE>: pair k = c / 2 of Philox call k)
template <int MODE>
__device__ __forceinline__ double impx_normal(uint64_t seed, long long i, long long step, int c)
{
    double z0, z1;
    normal_pair<MODE>(pvd_draw(seed, i, step, PVD_STREAM_DISP, (unsigned)(c >> 1)), z0, z1);
    return (c & 1) ? z1 : z0;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_impx_propose(const StepArgs a, const double *__restrict__ inv_mass, const double *__restrict__ x,
                                                      const double *__restrict__ f, double *__restrict__ y)
{
    const DevState *sip = &a.st[a.parity];
    if (sip->err || sip->n <= 0) return;
    const long long n = sip->n, step = sip->step;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        for (int c = 0; c < a.nc; c += 2) {
            double z0, z1;
            if (a.inj_disp) { z0 = a.inj_disp[c * a.cap + i]; z1 = c + 1 < a.nc ? a.inj_disp[(c + 1) * a.cap + i] : 0.0; }
            else {
                normal_pair<MODE>(pvd_draw(a.seed, i, step, PVD_STREAM_DISP, (unsigned)(c >> 1)), z0, z1);
                z0 = __dmul_rn(a.sigc[c], z0);
                z1 = __dmul_rn(a.sigc[c + 1 < a.nc ? c + 1 : c], z1);
            }
            for (int h = 0; h < 2 && c + h < a.nc; ++h) {
                const long long e = (long long)(c + h) * a.cap + i;
                const double d = __dmul_rn(inv_mass[(c + h) / a.ndim], f[e]);                   // D_x = (1/m) f_x
                y[e] = __dadd_rn(__dadd_rn(x[e], h ? z1 : z0), __dmul_rn(d, a.dt));
            }
        }
    }
}

// fy, sec: AoS (n, nc) as the host's call_derivs returned them; psiy: (n)
__global__ void __launch_bounds__(256) k_impx_accept(const StepArgs a, const double *__restrict__ inv_mass, double *x, double *f, double *psi,
                                                     double *lk, const double *__restrict__ y, const double *__restrict__ fy,
                                                     const double *__restrict__ psiy, const double *__restrict__ sec, const double *inj_um,
                                                     unsigned long long *acc_count)
{
    __shared__ unsigned s_cnt[8];
    __shared__ unsigned s_last;
    DevState *sip = &a.st[a.parity];
    if (sip->err || sip->n <= 0) return;
    const long long n = sip->n, step = sip->step;
    const int nc = a.nc, ndim = a.ndim, natoms = nc / ndim;
    unsigned my_acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // ImpSamp.metropolis: prod exp(-(x - y - D_y dt)^2 / 2 s^2) / exp(-(y - x - D_x dt)^2 / 2 s^2) * (psi_y / psi_x)^2, in the
        // summed-exponent form of metropolis_ratio (pvd_impsamp.cuh), same loop order (dimension outside, atoms inside)
        const double psx = psi[i], psy = psiy[i];
        double expo = 0.0;
        for (int d = 0; d < ndim; ++d)
            for (int at = 0; at < natoms; ++at) {
                const int c = at * ndim + d;
                const long long e = (long long)c * a.cap + i;
                const double inv_two_s2 = 0.5 / __dmul_rn(a.sigma[at], a.sigma[at]);
                const double xc = x[e], yc = y[e];
                const double dxm = __dmul_rn(__dmul_rn(inv_mass[at], f[e]), a.dt), dym = __dmul_rn(__dmul_rn(inv_mass[at], fy[i * nc + c]), a.dt);
                const double u1 = __dadd_rn(__dadd_rn(xc, -yc), -dym);
                const double u2 = __dadd_rn(__dadd_rn(yc, -xc), -dxm);
                expo = fma((u2 - u1) * (u2 + u1), inv_two_s2, expo);
            }
        const double q = psy / psx;
        double acc = __dmul_rn(exp(expo), __dmul_rn(q, q));
        if (__dmul_rn(psx, psy) <= 0.0) acc = 0.0;
        double u;
        if (inj_um) u = inj_um[i];
        else { const uint4 r = pvd_draw(a.seed, i, step, PVD_STREAM_METRO, 0u); u = u53(r.x, r.y); }
        if (acc > u) {
            // local_kin (imp_samp.py:49-53): -0.5 * sum over dims of (sum over atoms of (1/m) psi''/psi)
            double tot = 0.0;
            for (int d = 0; d < ndim; ++d) {
                double s = __dmul_rn(inv_mass[0], sec[i * nc + d]);
                for (int at = 1; at < natoms; ++at) s = __dadd_rn(s, __dmul_rn(inv_mass[at], sec[i * nc + at * ndim + d]));
                tot = d == 0 ? s : __dadd_rn(tot, s);
            }
            for (int c = 0; c < nc; ++c) {
                const long long e = (long long)c * a.cap + i;
                x[e] = y[e];
                f[e] = fy[i * nc + c];
            }
            psi[i] = psy;
            lk[i] = __dmul_rn(-0.5, tot);
            ++my_acc;
        }
    }
    for (int off = 16; off > 0; off >>= 1) my_acc += __shfl_xor_sync(0xffffffffu, my_acc, off);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = my_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_cnt[w];
        atomicAdd(acc_count, (unsigned long long)tot);
        __threadfence();
        const unsigned d = atomicAdd(&sip->done, 1u);
        s_last = (d == gridDim.x - 1) ? 1u : 0u;
        if (s_last) {
            __threadfence();
            const unsigned long long nacc = atomicAdd(acc_count, 0ull);
            sip->n_accept = (long long)nacc;
            if (a.world == 1) sip->dt_eff = __dmul_rn(a.dt, (double)nacc / (double)n);       // pyvibdmc.py:603, 372-378
            else { a.sums[0] = (double)nacc; a.sums[1] = (double)n; }
            sip->done = 0u;
            *acc_count = 0ull;
        }
    }
}

// drift terms of the start ensemble (first-step exception, pyvibdmc.py:553-554, 760-769): f, psi as given, T_L from psi''/psi
__global__ void k_impx_init(long long n, long long cap, int nc, int ndim, const double *__restrict__ inv_mass, const double *__restrict__ fx,
                            const double *__restrict__ psi_in, const double *__restrict__ sec, double *f, double *psi, double *lk)
{
    const int natoms = nc / ndim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double tot = 0.0;
        for (int d = 0; d < ndim; ++d) {
            double s = __dmul_rn(inv_mass[0], sec[i * nc + d]);
            for (int at = 1; at < natoms; ++at) s = __dadd_rn(s, __dmul_rn(inv_mass[at], sec[i * nc + at * ndim + d]));
            tot = d == 0 ? s : __dadd_rn(tot, s);
        }
        for (int c = 0; c < nc; ++c) f[(long long)c * cap + i] = fx[i * nc + c];
        psi[i] = psi_in[i];
        lk[i] = __dmul_rn(-0.5, tot);
    }
}

// E_L = V + T_L (pyvibdmc.py:807-809) on the stored energies
__global__ void k_impx_add_lk(const DevState *st, int parity, double *v, const double *__restrict__ lk)
{
    const long long n = st[parity].n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] = __dadd_rn(v[i], lk[i]);
}

// stand-alone ImpSamp.metropolis / ImpSamp.local_kin for any (atoms x dims): AoS host layout
__global__ void k_metropolis_rt(const double *x, const double *y, const double *fx, const double *fy, const double *psx, const double *psy,
                                long long n, int nc, int ndim, const double *sigma, const double *inv_mass, double dt, double *acc)
{
    const int natoms = nc / ndim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double expo = 0.0;
        for (int d = 0; d < ndim; ++d)
            for (int at = 0; at < natoms; ++at) {
                const long long e = i * nc + at * ndim + d;
                const double inv_two_s2 = 0.5 / __dmul_rn(sigma[at], sigma[at]);
                const double dxm = __dmul_rn(__dmul_rn(inv_mass[at], fx[e]), dt), dym = __dmul_rn(__dmul_rn(inv_mass[at], fy[e]), dt);
                const double u1 = __dadd_rn(__dadd_rn(x[e], -y[e]), -dym);
                const double u2 = __dadd_rn(__dadd_rn(y[e], -x[e]), -dxm);
                expo = fma((u2 - u1) * (u2 + u1), inv_two_s2, expo);
            }
        const double q = psy[i] / psx[i];
        double r = __dmul_rn(exp(expo), __dmul_rn(q, q));
        if (__dmul_rn(psx[i], psy[i]) <= 0.0) r = 0.0;
        acc[i] = r;
    }
}

__global__ void k_local_kin_rt(const double *d2, long long n, int nc, int ndim, const double *inv_mass, double *ke)
{
    const int natoms = nc / ndim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double tot = 0.0;
        for (int d = 0; d < ndim; ++d) {
            double s = __dmul_rn(inv_mass[0], d2[i * nc + d]);
            for (int at = 1; at < natoms; ++at) s = __dadd_rn(s, __dmul_rn(inv_mass[at], d2[i * nc + at * ndim + d]));
            tot = d == 0 ? s : __dadd_rn(tot, s);
        }
        ke[i] = __dmul_rn(-0.5, tot);
    }
}

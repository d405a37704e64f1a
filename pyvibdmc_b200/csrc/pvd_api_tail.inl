// Host-side pieces that need the full pvd_sim definition: continuous weighting, importance
// sampling, NN potential glue, and the remaining stand-alone entry points.

// ---------------------------------------------------------------- continuous weighting
static ContArgs make_cont_args(pvd_sim *s)
{
    ContArgs ca{};
    ca.w = s->w.as<double>();
    ca.v = s->v[s->cur].as<double>();
    ca.kill_idx = s->kill_idx.as<int>();
    ca.kill_mask = s->kill_mask.as<unsigned>();
    ca.copy_dst = s->copy_dst.as<int>();
    ca.copy_src = s->copy_src.as<int>();
    ca.cand = s->cand.as<ContCand>();
    ca.sorted = s->cand_sorted.as<ContCand>();
    ca.bin_start = s->bin_start.as<unsigned>();
    ca.bin_fill = s->bin_fill.as<unsigned>();
    ca.hist = s->hist.as<unsigned>();
    ca.work = s->cont_work.as<ContWork>();
    ca.cand_cap = s->cap;
    ca.lower = s->cfg.thresh_lower;
    ca.upper = s->cfg.thresh_upper;
    ca.has_upper = (s->cfg.thresh_upper == s->cfg.thresh_upper) ? 1 : 0;
    return ca;
}

// kill list + histogram are produced by the k_cont_update instance the caller launched; then:
// ranked candidates -> donor assignment -> copies -> Vref
static int cont_launch_tail(pvd_sim *s, const StepArgs &a, const ContArgs &ca, long long *src_out)
{
    const int g = s->grid_light;
    const bool imp = s->cfg.trial != PVD_TRIAL_NONE;
    // six kernels per step: launched with programmatic stream serialisation so that each one's launch latency
    // overlaps its predecessor's tail (every one of them starts with pdl_wait())
    PVD_CUDA(launch_pdl(k_cont_prefix, dim3(1), dim3(1024), 0, s->stream, a, ca));
    PVD_CHECK_LAUNCH();
    PVD_CUDA(launch_pdl(k_cont_collect, dim3((unsigned)g), dim3(PVD_CTA), 0, s->stream, a, ca));
    PVD_CHECK_LAUNCH();
    // grid: CTA 0 makes the kill list, one CTA per SM orders the bins; the last one to finish assigns the donors
    PVD_CUDA(launch_pdl(k_cont_rank, dim3((unsigned)(g_num_sms > 0 ? g_num_sms : 148) + 1u), dim3(PVD_RANK_SUB), (size_t)PVD_RANK_SMEM, s->stream, a, ca,
                        s->cont_queue.as<ContCand>(), s->cont_root.as<int>(), s->cont_skip.as<unsigned char>()));
    PVD_CHECK_LAUNCH();
    PVD_CUDA(launch_pdl(k_cont_copy, dim3((unsigned)g), dim3(PVD_CTA), 0, s->stream, a, ca, s->x[s->cur].as<double>(), s->v[s->cur].as<double>(),
                        s->who[s->cur].as<int>(), imp ? s->f[s->cur].as<double>() : (double *)nullptr,
                        imp ? s->psi[s->cur].as<double>() : (double *)nullptr, imp ? s->lk[s->cur].as<double>() : (double *)nullptr, src_out));
    PVD_CHECK_LAUNCH();
    PVD_CUDA(launch_pdl(k_cont_finish, dim3((unsigned)(g_num_sms > 0 ? g_num_sms : 148)), dim3(1024), 0, s->stream, a, ca));
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}

static void cont_in_place(pvd_sim *s, StepArgs &a)
{
    a.xin = a.xout = s->x[s->cur].as<double>();
    a.vin = a.vout = s->v[s->cur].as<double>();
    a.who_in = a.who_out = s->who[s->cur].as<int>();
    if (s->cfg.imp_variant == PVD_IMP_EXCITED_STATE) a.vsin = a.vsout = s->vs[s->cur].as<double>();
    a.flip = 0;
}

// weight update + branching + Vref on energies already stored in v[cur] (in place)
static int cont_enqueue_branch_only(pvd_sim *s, StepArgs &a, long long *src_out)
{
    cont_in_place(s, a);
    ContArgs ca = make_cont_args(s);
    PVD_CUDA(launch_pdl(k_cont_update<ContFromMemory>, dim3((unsigned)s->grid_light), dim3(PVD_CTA), 0, s->stream, a, ca));
    PVD_CHECK_LAUNCH();
    return cont_launch_tail(s, a, ca, src_out);
}

// move + built-in potential + weight update in one kernel, then the branching tail
static int cont_enqueue_step(pvd_sim *s, StepArgs &a)
{
    const bool fast = s->cfg.rng_mode == PVD_RNG_FAST;
    cont_in_place(s, a);
    ContArgs ca = make_cont_args(s);
#define LAUNCH_CONT(POT)                                                                                             \
    do {                                                                                                             \
        const int gp = POT::MIN_CTAS >= 4 ? s->grid_light : s->grid;                                                 \
        if (fast) { PVD_CUDA(launch_pdl(k_cont_update<ContFused<POT, PVD_RNG_FAST>>, dim3((unsigned)gp), dim3(PVD_CTA), 0, s->stream, a, ca)); } \
        else if (s->cfg.rng_mode == PVD_RNG_ZIGGURAT) { PVD_CUDA(launch_pdl(k_cont_update<ContFused<POT, PVD_RNG_ZIGGURAT>>, dim3((unsigned)gp), dim3(PVD_CTA), 0, s->stream, a, ca)); } \
        else { PVD_CUDA(launch_pdl(k_cont_update<ContFused<POT, PVD_RNG_FP64>>, dim3((unsigned)gp), dim3(PVD_CTA), 0, s->stream, a, ca)); }      \
    } while (0)
    switch (s->cfg.potential) {
    case PVD_POT_H2O_PS: LAUNCH_CONT(PotH2O); break;
    case PVD_POT_HARMONIC:
        if (s->nc == 1) LAUNCH_CONT(PotHarm<1>);
        else if (s->nc == 3) LAUNCH_CONT(PotHarm<3>);
        else return pvd_fail(PVD_E_ARG, "built-in harmonic potential supports 1 or 3 components");
        break;
    case PVD_POT_MORSE1D: LAUNCH_CONT(PotMorse); break;
    default: return pvd_fail(PVD_E_STATE, "continuous weighting on the device needs a built-in fp64 potential");
    }
#undef LAUNCH_CONT
    PVD_CHECK_LAUNCH();
    return cont_launch_tail(s, a, ca, nullptr);
}

extern "C" int pvd_branch_continuous(double *w, const double *v, int64_t n, double vref, double dt, double lower, double upper,
                                     int64_t *src, double *stats3)
{
    PVD_REQUIRE(w && v && src && stats3 && n >= 1, "pvd_branch_continuous: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    pvd_config cfg{};
    cfg.natoms = 1; cfg.ndim = 1; cfg.weighting = PVD_WEIGHT_CONTINUOUS; cfg.potential = PVD_POT_EXTERNAL;
    cfg.world_size = 1; cfg.num_walkers = n; cfg.capacity = n; cfg.delta_t = dt; cfg.alpha = 1.0 / (2.0 * dt);
    cfg.thresh_lower = lower; cfg.thresh_upper = upper; cfg.masses[0] = 1.0; cfg.stats_ring = 1;
    PVD_CUDA(cudaGetDevice(&cfg.device));
    pvd_sim *s = nullptr;
    if (int rc = pvd_sim_create(&cfg, &s)) return rc;
    struct Guard { pvd_sim *s; ~Guard() { pvd_sim_destroy(s); } } guard{s};
    std::vector<double> x0((size_t)n, 0.0);
    if (int rc = pvd_sim_upload(s, x0.data(), n, w)) return rc;
    // state: energies and the injected Vref
    PVD_CUDA(cudaMemcpy(s->v[s->cur].p, v, (size_t)n * 8, cudaMemcpyHostToDevice));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    h[0].vref = vref; h[0].dt_eff = dt; h[1] = h[0];
    PVD_CUDA(cudaMemcpy(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice));
    DevBuf dsrc;
    PVD_CUDA(dsrc.alloc((size_t)n * 8));
    k_iota_i64<<<grid_for(n, 256, 16), 256, 0, s->stream>>>(dsrc.as<long long>(), n);
    PVD_CHECK_LAUNCH();
    StepArgs a = make_args(s, 1);
    if (int rc = cont_enqueue_branch_only(s, a, dsrc.as<long long>())) return rc;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    PVD_CUDA(cudaMemcpy(w, s->w.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    PVD_CUDA(cudaMemcpy(src, dsrc.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    pvd_step_stats r;
    PVD_CUDA(cudaMemcpy(&r, s->ring.p, sizeof(r), cudaMemcpyDeviceToHost));
    stats3[0] = (double)r.births; stats3[1] = r.w_max; stats3[2] = r.w_min;
    return PVD_OK;
}

// ---------------------------------------------------------------- importance sampling
static int fill_trial_params(pvd_sim *s, ImpArgs &im)
{
    memset(&im, 0, sizeof(im));
    for (int a = 0; a < PVD_MAX_ATOMS; ++a) { im.inv_mass[a] = s->inv_mass[a]; im.mass[a] = s->cfg.masses[a]; }
    im.vscore = s->cfg.imp_variant == PVD_IMP_EXCITED_STATE ? s->vs[s->cur].as<double>() : nullptr;
    im.trial = s->trial_params;
    im.acc_count = s->acc_count.as<unsigned long long>();
    if ((s->cfg.trial == PVD_TRIAL_H2O_FD || s->cfg.trial == PVD_TRIAL_H2O_AN) && !s->trial_table.p)
        return pvd_fail(PVD_E_STATE, "water trial wfn: call pvd_sim_set_trial_table first");
    return PVD_OK;
}

// caller's table size (doubles) for a trial id
static size_t trial_table_doubles(int32_t trial, int64_t ntab)
{
    if (trial == PVD_TRIAL_H2O_FD) return (size_t)(2 * ntab + 2);
    if (trial == PVD_TRIAL_H2O_AN) return (size_t)(4 * ntab + 2);
    return (size_t)ntab;
}

// device image of a trial table: the caller's values followed (water product wfn) by the np.interp slopes of every row
static std::vector<double> trial_table_with_slopes(int32_t trial, const double *table, int64_t ntab, size_t total)
{
    std::vector<double> host(table, table + total);
    const int rows = trial == PVD_TRIAL_H2O_FD ? 1 : (trial == PVD_TRIAL_H2O_AN ? 3 : 0);
    host.resize(total + (size_t)rows * ntab, 0.0);
    for (int r = 0; r < rows; ++r)
        for (int64_t j = 0; j + 1 < ntab; ++j) {
            volatile double df = table[(r + 1) * ntab + j + 1] - table[(r + 1) * ntab + j], dx = table[j + 1] - table[j];
            host[total + (size_t)r * ntab + (size_t)j] = df / dx;
        }
    return host;
}

static int host_trial_params(int32_t trial, const double *table, int64_t ntab, const double *dev_table, TrialParamsDev &p)
{
    memset(&p, 0, sizeof(p));
    p.fd_dx = 0.001;
    p.fd_dx2 = 0.001 * 0.001;               // Python: dx ** 2
    if (trial == PVD_TRIAL_HARM1D) {
        PVD_REQUIRE(table && ntab >= 1, "HARM1D trial needs {alpha}");
        p.h_alpha = table[0];
        p.h_pref = pow(table[0] / 3.141592653589793, 0.25);
        return PVD_OK;
    }
    if (trial == PVD_TRIAL_H2O_FD || trial == PVD_TRIAL_H2O_AN) {
        // table layout: grid[ntab], psi[ntab], (psi'[ntab], psi''[ntab],) then {ang_alpha, theta_eq}; slopes appended on the device
        PVD_REQUIRE(table && ntab >= 4, "H2O trial needs the (rows, ntab) table followed by {alpha_theta, theta_eq}");
        const int64_t rows = trial == PVD_TRIAL_H2O_AN ? 4 : 2;
        const size_t total = trial_table_doubles(trial, ntab);
        p.grid = dev_table;
        p.wfn = dev_table + ntab;
        p.slope = dev_table + total;
        if (trial == PVD_TRIAL_H2O_AN) {
            p.dwfn = dev_table + 2 * ntab;
            p.d2wfn = dev_table + 3 * ntab;
            p.dslope = dev_table + total + ntab;
            p.d2slope = dev_table + total + 2 * ntab;
        }
        p.ntab = (int)ntab;
        p.g0 = table[0];
        p.g_last = table[ntab - 1];
        p.w_first = table[ntab];
        p.w_last = table[2 * ntab - 1];
        p.inv_step = (double)(ntab - 1) / (table[ntab - 1] - table[0]);
        p.ang_alpha = table[rows * ntab];
        p.theta_eq = table[rows * ntab + 1];
        p.ang_pref = pow(p.ang_alpha / 3.141592653589793, 0.25);
        return PVD_OK;
    }
    return pvd_fail(PVD_E_ARG, "unknown trial wave function id");
}

extern "C" int pvd_sim_set_trial_table(pvd_sim *s, const double *table, int64_t ntab)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->cfg.trial != PVD_TRIAL_NONE, "this handle was created without a trial wave function");
    const size_t total = trial_table_doubles(s->cfg.trial, ntab);
    std::vector<double> host = trial_table_with_slopes(s->cfg.trial, table, ntab, total);
    PVD_CUDA(s->trial_table.alloc(host.size() * 8));
    PVD_CUDA(cudaMemcpy(s->trial_table.p, host.data(), host.size() * 8, cudaMemcpyHostToDevice));
    s->ntrial = ntab;
    return host_trial_params(s->cfg.trial, table, ntab, s->trial_table.as<double>(), s->trial_params);
}

#define IMP_DISPATCH(KERNEL_CALL)                                                                                           \
    do {                                                                                                                    \
        const int t = s->cfg.trial, p = s->cfg.potential;                                                                   \
        if (t == PVD_TRIAL_H2O_FD && p == PVD_POT_H2O_PS) { KERNEL_CALL(TrialH2O, PotH2O); }                                 \
        else if (t == PVD_TRIAL_H2O_AN && p == PVD_POT_H2O_PS) { KERNEL_CALL(TrialH2OAn, PotH2O); }                          \
        else if (t == PVD_TRIAL_HARM1D && p == PVD_POT_HARMONIC && s->nc == 1) { KERNEL_CALL(TrialHarm1D, PotHarm<1>); }     \
        else if (t == PVD_TRIAL_HARM1D && p == PVD_POT_MORSE1D) { KERNEL_CALL(TrialHarm1D, PotMorse); }                      \
        else return pvd_fail(PVD_E_ARG, "unsupported built-in (trial, potential) combination");                             \
    } while (0)

static int imp_initial_drift(pvd_sim *s, long long first, long long count)
{
    ImpArgs im;
    if (int rc = fill_trial_params(s, im)) return rc;
    StepArgs a = make_args(s, 1);
    const int g = s->grid;
    double *x = s->x[s->cur].as<double>(), *f = s->f[s->cur].as<double>(), *psi = s->psi[s->cur].as<double>();
    double *lk = s->lk[s->cur].as<double>(), *v = s->v[s->cur].as<double>();
#define CALL_INIT(T, P) k_imp_init<T, P><<<g, PVD_CTA, 0, s->stream>>>(a, im, x, f, psi, lk, v, first, count)
    IMP_DISPATCH(CALL_INIT);
#undef CALL_INIT
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}

static int imp_enqueue_move(pvd_sim *s, StepArgs &a, const double *inj_um)
{
    ImpArgs im;
    if (int rc = fill_trial_params(s, im)) return rc;
    im.inj_um = inj_um;
    const int g = s->grid;
    const bool fast = s->cfg.rng_mode == PVD_RNG_FAST;
    const int variant = s->cfg.imp_variant;
    double *x = s->x[s->cur].as<double>(), *f = s->f[s->cur].as<double>(), *psi = s->psi[s->cur].as<double>();
    double *lk = s->lk[s->cur].as<double>(), *v = s->v[s->cur].as<double>();
#define LAUNCH_MOVE(T, P, R, S)                                                                                            \
    do {                                                                                                                  \
        PVD_CUDA(cudaFuncSetAttribute(k_imp_move<T, P, R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));      \
        k_imp_move<T, P, R, S><<<g, PVD_CTA, sm, s->stream>>>(a, im, x, f, psi, lk, v);                                    \
    } while (0)
#define CALL_MOVE(T, P)                                                                                 \
    do {                                                                                                \
        const size_t sm = (size_t)3 * T::NC * PVD_CTA * sizeof(double);                                  \
        if (variant == PVD_IMP_SECOND_DISPLACEMENT) {                                                   \
            if (fast) LAUNCH_MOVE(T, P, PVD_RNG_FAST, PVD_IMP_SECOND_DISPLACEMENT); else LAUNCH_MOVE(T, P, PVD_RNG_FP64, PVD_IMP_SECOND_DISPLACEMENT); \
        } else if (variant == PVD_IMP_EXCITED_STATE) {                                                  \
            if (fast) LAUNCH_MOVE(T, P, PVD_RNG_FAST, PVD_IMP_EXCITED_STATE); else LAUNCH_MOVE(T, P, PVD_RNG_FP64, PVD_IMP_EXCITED_STATE); \
        } else {                                                                                        \
            if (fast) LAUNCH_MOVE(T, P, PVD_RNG_FAST, PVD_IMP_STANDARD); else LAUNCH_MOVE(T, P, PVD_RNG_FP64, PVD_IMP_STANDARD); \
        }                                                                                               \
    } while (0)
    IMP_DISPATCH(CALL_MOVE);
#undef CALL_MOVE
#undef LAUNCH_MOVE
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}

static int imp_enqueue_branch(pvd_sim *s, StepArgs &a)
{
    if (s->cfg.weighting == PVD_WEIGHT_CONTINUOUS) return cont_enqueue_branch_only(s, a);
    // discrete: branch on E_L with the effective time step, carrying f_x, psi and the local kinetic energy
    k_branch_discrete<<<s->grid_light, PVD_CTA, 0, s->stream>>>(a);
    PVD_CHECK_LAUNCH();
    s->cur ^= 1;
    return PVD_OK;
}

static int imp_enqueue_step(pvd_sim *s, StepArgs &a, const double *inj_um)
{
    if (int rc = imp_enqueue_move(s, a, inj_um)) return rc;
    return imp_enqueue_branch(s, a);
}

extern "C" {

// multi-GPU importance sampling: the acceptance fraction that scales the time step (pyvibdmc.py:372-378, 603) is global,
// so the step is cut after the Metropolis move: move -> all-reduce of `sums` -> branch -> all-reduce -> finalize
int pvd_sim_imp_move_local(pvd_sim *s)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && s->cfg.trial != PVD_TRIAL_NONE, "pvd_sim_imp_move_local: needs an importance-sampled simulation with walkers");
    StepArgs a = make_args(s, 1);
    return imp_enqueue_move(s, a, nullptr);
}

int pvd_sim_imp_branch_local(pvd_sim *s, int32_t do_branch)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->uploaded && s->cfg.trial != PVD_TRIAL_NONE, "pvd_sim_imp_branch_local: needs an importance-sampled simulation with walkers");
    StepArgs a = make_args(s, do_branch);
    if (s->cfg.world_size > 1) {
        k_imp_set_dt<<<1, 32, 0, s->stream>>>(s->st.as<DevState>(), s->parity, a.sums, s->cfg.delta_t);
        PVD_CHECK_LAUNCH();
    }
    if (int rc = imp_enqueue_branch(s, a)) return rc;
    s->parity ^= 1;
    return PVD_OK;
}

int pvd_sim_download_imp(pvd_sim *s, double *fx, double *psi, double *sec, int64_t capacity)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->cfg.trial != PVD_TRIAL_NONE, "no trial wave function configured");
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    PVD_REQUIRE(capacity >= n, "pvd_sim_download_imp: host buffers too small");
    const int nc = s->nc;
    // the second derivatives are a pure function of the coordinates: recompute them (only the local
    // kinetic energy is carried on the device)
    ImpArgs im;
    if (int rc = fill_trial_params(s, im)) return rc;
    DevBuf aos, dpsi, dlog, d2;
    PVD_CUDA(aos.alloc((size_t)n * nc * 8)); PVD_CUDA(dpsi.alloc((size_t)n * 8));
    PVD_CUDA(dlog.alloc((size_t)n * nc * 8)); PVD_CUDA(d2.alloc((size_t)n * nc * 8));
    k_soa_to_aos<<<grid_for(n * nc, 256, 16), 256, 0, s->stream>>>(s->x[s->cur].as<double>(), aos.as<double>(), n, nc, s->cap);
    PVD_CHECK_LAUNCH();
    if (s->cfg.trial == PVD_TRIAL_H2O_FD)
        k_trial_drift_aos<TrialH2O><<<grid_for(n, 128, 16), 128, 0, s->stream>>>(aos.as<double>(), n, im.trial, dpsi.as<double>(), dlog.as<double>(), d2.as<double>());
    else if (s->cfg.trial == PVD_TRIAL_H2O_AN)
        k_trial_drift_aos<TrialH2OAn><<<grid_for(n, 128, 16), 128, 0, s->stream>>>(aos.as<double>(), n, im.trial, dpsi.as<double>(), dlog.as<double>(), d2.as<double>());
    else
        k_trial_drift_aos<TrialHarm1D><<<grid_for(n, 128, 16), 128, 0, s->stream>>>(aos.as<double>(), n, im.trial, dpsi.as<double>(), dlog.as<double>(), d2.as<double>());
    PVD_CHECK_LAUNCH();
    if (fx) {
        // carried drift (the one the next Metropolis step will use)
        k_soa_to_aos<<<grid_for(n * nc, 256, 16), 256, 0, s->stream>>>(s->f[s->cur].as<double>(), aos.as<double>(), n, nc, s->cap);
        PVD_CHECK_LAUNCH();
        PVD_CUDA(cudaMemcpyAsync(fx, aos.p, (size_t)n * nc * 8, cudaMemcpyDeviceToHost, s->stream));
    }
    if (psi) PVD_CUDA(cudaMemcpyAsync(psi, s->psi[s->cur].p, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    if (sec) PVD_CUDA(cudaMemcpyAsync(sec, d2.p, (size_t)n * nc * 8, cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

int pvd_trial_drift(int32_t trial, const double *xyz, int64_t n, int32_t natoms, int32_t ndim, const double *table, int64_t ntab,
                    double *psi, double *dlog, double *d2)
{
    PVD_REQUIRE(xyz && psi && dlog && d2 && n >= 0 && natoms >= 1 && ndim >= 1, "pvd_trial_drift: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    if (n == 0) return PVD_OK;
    const int nc = natoms * ndim;
    PVD_REQUIRE(((trial == PVD_TRIAL_H2O_FD || trial == PVD_TRIAL_H2O_AN) && nc == 9) || (trial == PVD_TRIAL_HARM1D && nc == 1), "trial / shape mismatch");
    DevBuf dt, dx, dpsi, dlogb, d2b;
    const size_t total = trial_table_doubles(trial, ntab);
    std::vector<double> host = trial_table_with_slopes(trial, table, ntab, total);
    PVD_CUDA(dt.alloc(host.size() * 8));
    PVD_CUDA(cudaMemcpy(dt.p, host.data(), host.size() * 8, cudaMemcpyHostToDevice));
    TrialParamsDev p;
    if (int rc = host_trial_params(trial, table, ntab, dt.as<double>(), p)) return rc;
    PVD_CUDA(dx.alloc((size_t)n * nc * 8)); PVD_CUDA(dpsi.alloc((size_t)n * 8));
    PVD_CUDA(dlogb.alloc((size_t)n * nc * 8)); PVD_CUDA(d2b.alloc((size_t)n * nc * 8));
    PVD_CUDA(cudaMemcpy(dx.p, xyz, (size_t)n * nc * 8, cudaMemcpyHostToDevice));
    EventPair ev;
    PVD_CUDA(ev.init());
    PVD_CUDA(cudaEventRecord(ev.a));
    if (trial == PVD_TRIAL_H2O_FD)
        k_trial_drift_aos<TrialH2O><<<grid_for(n, 128, 16), 128>>>(dx.as<double>(), n, p, dpsi.as<double>(), dlogb.as<double>(), d2b.as<double>());
    else if (trial == PVD_TRIAL_H2O_AN)
        k_trial_drift_aos<TrialH2OAn><<<grid_for(n, 128, 16), 128>>>(dx.as<double>(), n, p, dpsi.as<double>(), dlogb.as<double>(), d2b.as<double>());
    else
        k_trial_drift_aos<TrialHarm1D><<<grid_for(n, 128, 16), 128>>>(dx.as<double>(), n, p, dpsi.as<double>(), dlogb.as<double>(), d2b.as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaEventRecord(ev.b));
    PVD_CUDA(cudaMemcpy(psi, dpsi.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    PVD_CUDA(cudaMemcpy(dlog, dlogb.p, (size_t)n * nc * 8, cudaMemcpyDeviceToHost));
    PVD_CUDA(cudaMemcpy(d2, d2b.p, (size_t)n * nc * 8, cudaMemcpyDeviceToHost));
    float ms = 0;
    PVD_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    g_last_kernel_ms = ms;
    return PVD_OK;
}

int pvd_metropolis(const double *x, const double *y, const double *fx, const double *fy, const double *psi_x, const double *psi_y,
                   int64_t n, int32_t natoms, int32_t ndim, const double *sigma, const double *inv_mass, double dt, double *acc)
{
    PVD_REQUIRE(x && y && fx && fy && psi_x && psi_y && sigma && inv_mass && acc && n >= 0, "pvd_metropolis: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    if (n == 0) return PVD_OK;
    const int nc = natoms * ndim;
    PVD_REQUIRE(natoms >= 1 && natoms <= PVD_MAX_ATOMS && ndim >= 1 && nc <= PVD_MAX_COMP, "pvd_metropolis: bad natoms / ndim");
    DevBuf b[6], ds, dm, dacc;
    const double *src[4] = {x, y, fx, fy};
    for (int k = 0; k < 4; ++k) { PVD_CUDA(b[k].alloc((size_t)n * nc * 8)); PVD_CUDA(cudaMemcpy(b[k].p, src[k], (size_t)n * nc * 8, cudaMemcpyHostToDevice)); }
    PVD_CUDA(b[4].alloc((size_t)n * 8)); PVD_CUDA(cudaMemcpy(b[4].p, psi_x, (size_t)n * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(b[5].alloc((size_t)n * 8)); PVD_CUDA(cudaMemcpy(b[5].p, psi_y, (size_t)n * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(ds.alloc(natoms * 8)); PVD_CUDA(cudaMemcpy(ds.p, sigma, natoms * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(dm.alloc(natoms * 8)); PVD_CUDA(cudaMemcpy(dm.p, inv_mass, natoms * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(dacc.alloc((size_t)n * 8));
    const int g = grid_for(n, 128, 16);
    if (!((nc == 9 && ndim == 3) || (nc == 1 && ndim == 1)))       // any other (atoms x dims): the run-time kernel (imp_samp.py:29-47 is shape generic)
        k_metropolis_rt<<<g, 128>>>(b[0].as<double>(), b[1].as<double>(), b[2].as<double>(), b[3].as<double>(), b[4].as<double>(),
                                    b[5].as<double>(), n, nc, ndim, ds.as<double>(), dm.as<double>(), dt, dacc.as<double>());
    else if (nc == 9)
        k_metropolis_aos<9><<<g, 128>>>(b[0].as<double>(), b[1].as<double>(), b[2].as<double>(), b[3].as<double>(), b[4].as<double>(),
                                       b[5].as<double>(), n, ndim, ds.as<double>(), dm.as<double>(), dt, dacc.as<double>());
    else
        k_metropolis_aos<1><<<g, 128>>>(b[0].as<double>(), b[1].as<double>(), b[2].as<double>(), b[3].as<double>(), b[4].as<double>(),
                                       b[5].as<double>(), n, ndim, ds.as<double>(), dm.as<double>(), dt, dacc.as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpy(acc, dacc.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return PVD_OK;
}

int pvd_local_kin(const double *d2, int64_t n, int32_t natoms, int32_t ndim, const double *inv_mass, double *ke)
{
    PVD_REQUIRE(d2 && inv_mass && ke && n >= 0, "pvd_local_kin: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    if (n == 0) return PVD_OK;
    const int nc = natoms * ndim;
    PVD_REQUIRE(natoms >= 1 && natoms <= PVD_MAX_ATOMS && ndim >= 1 && nc <= PVD_MAX_COMP, "pvd_local_kin: bad natoms / ndim");
    DevBuf dd, dm, dk;
    PVD_CUDA(dd.alloc((size_t)n * nc * 8)); PVD_CUDA(cudaMemcpy(dd.p, d2, (size_t)n * nc * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(dm.alloc(natoms * 8)); PVD_CUDA(cudaMemcpy(dm.p, inv_mass, natoms * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(dk.alloc((size_t)n * 8));
    if (!((nc == 9 && ndim == 3) || (nc == 1 && ndim == 1)))
        k_local_kin_rt<<<grid_for(n, 128, 16), 128>>>(dd.as<double>(), n, nc, ndim, dm.as<double>(), dk.as<double>());
    else if (nc == 9) k_local_kin_aos<9><<<grid_for(n, 128, 16), 128>>>(dd.as<double>(), n, ndim, dm.as<double>(), dk.as<double>());
    else k_local_kin_aos<1><<<grid_for(n, 128, 16), 128>>>(dd.as<double>(), n, ndim, dm.as<double>(), dk.as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpy(ke, dk.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return PVD_OK;
}

// ---------------------------------------------------------------- importance sampling with a user trial wave function
// (PVD_TRIAL_EXTERNAL, csrc/pvd_impext.cuh): the host evaluates psi and its derivatives once per step, the GPU does the rest
static int impx_buffers(pvd_sim *s)
{
    if (s->impx_y.p) return PVD_OK;
    PVD_CUDA(s->impx_y.alloc((size_t)s->cap * s->nc * 8));
    PVD_CUDA(s->impx_fy.alloc((size_t)s->cap * s->nc * 8));
    PVD_CUDA(s->impx_sec.alloc((size_t)s->cap * s->nc * 8));
    PVD_CUDA(s->impx_psiy.alloc((size_t)s->cap * 8));
    PVD_CUDA(s->impx_invm.alloc(PVD_MAX_ATOMS * 8));
    PVD_CUDA(cudaMemcpy(s->impx_invm.p, s->inv_mass, PVD_MAX_ATOMS * 8, cudaMemcpyHostToDevice));
    return PVD_OK;
}

extern "C" {

int pvd_sim_imp_ext_init(pvd_sim *s, const double *fx, const double *psi, const double *sec, const double *v)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->cfg.trial == PVD_TRIAL_EXTERNAL && s->uploaded, "pvd_sim_imp_ext_init: needs trial = PVD_TRIAL_EXTERNAL and uploaded walkers");
    PVD_REQUIRE(fx && psi && sec && (v || s->cfg.potential != PVD_POT_EXTERNAL), "pvd_sim_imp_ext_init: NULL argument");
    if (int rc = impx_buffers(s)) return rc;
    const long long n = s->n_uploaded;
    const int nc = s->nc;
    PVD_CUDA(cudaMemcpyAsync(s->impx_fy.p, fx, (size_t)n * nc * 8, cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaMemcpyAsync(s->impx_sec.p, sec, (size_t)n * nc * 8, cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaMemcpyAsync(s->impx_psiy.p, psi, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    k_impx_init<<<grid_for(n, 256, 8), 256, 0, s->stream>>>(n, s->cap, nc, s->cfg.ndim, s->impx_invm.as<double>(), s->impx_fy.as<double>(),
                                                           s->impx_psiy.as<double>(), s->impx_sec.as<double>(), s->f[s->cur].as<double>(),
                                                           s->psi[s->cur].as<double>(), s->lk[s->cur].as<double>());
    PVD_CHECK_LAUNCH();
    // first-step exception: V on the start ensemble, E_L = V + T_L, Vref from it (pyvibdmc.py:760-769)
    if (v) PVD_CUDA(cudaMemcpyAsync(s->v[s->cur].p, v, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    else if (int rc = launch_pot_soa(s)) return rc;
    k_impx_add_lk<<<grid_for(n, 256, 8), 256, 0, s->stream>>>(s->st.as<DevState>(), s->parity, s->v[s->cur].as<double>(), s->lk[s->cur].as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    if (int rc = sim_init_sums(s)) return rc;
    if (s->cfg.world_size == 1) return pvd_sim_init_finalize(s);
    return PVD_OK;                     // sharded: the caller all-reduces the sums, then pvd_sim_init_finalize
}

int pvd_sim_imp_ext_propose(pvd_sim *s, const double *disp, double *xyz_out, int64_t *n_out)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->cfg.trial == PVD_TRIAL_EXTERNAL && s->uploaded && s->impx_y.p && xyz_out && n_out, "pvd_sim_imp_ext_propose: call pvd_sim_imp_ext_init first");
    StepArgs a = make_args(s, 1);
    if (disp) {
        // injected displacements (n, atoms, dims), already scaled by sigma: parity replays of reference trajectories
        PVD_CUDA(cudaStreamSynchronize(s->stream));
        DevState h0[2];
        PVD_CUDA(cudaMemcpy(h0, s->st.p, sizeof(h0), cudaMemcpyDeviceToHost));
        const long long n0 = h0[s->parity].n;
        if (!s->inj_disp.p) {
            PVD_CUDA(s->inj_disp.alloc((size_t)s->cap * s->nc * 8));
            PVD_CUDA(s->inj_u.alloc((size_t)s->cap * 8));
            PVD_CUDA(s->inj_um.alloc((size_t)s->cap * 8));
        }
        PVD_CUDA(s->stage.alloc((size_t)n0 * s->nc * 8));
        PVD_CUDA(cudaMemcpyAsync(s->stage.p, disp, (size_t)n0 * s->nc * 8, cudaMemcpyHostToDevice, s->stream));
        k_aos_to_soa<<<grid_for(n0 * s->nc, 256, 16), 256, 0, s->stream>>>(s->stage.as<double>(), s->inj_disp.as<double>(), n0, s->nc, s->cap);
        PVD_CHECK_LAUNCH();
        a.inj_disp = s->inj_disp.as<double>();
    }
    const int g = grid_for(s->cap, 256, 8);
    if (s->cfg.rng_mode == PVD_RNG_FAST)
        k_impx_propose<PVD_RNG_FAST><<<g, 256, 0, s->stream>>>(a, s->impx_invm.as<double>(), s->x[s->cur].as<double>(), s->f[s->cur].as<double>(), s->impx_y.as<double>());
    else
        k_impx_propose<PVD_RNG_FP64><<<g, 256, 0, s->stream>>>(a, s->impx_invm.as<double>(), s->x[s->cur].as<double>(), s->f[s->cur].as<double>(), s->impx_y.as<double>());
    PVD_CHECK_LAUNCH();
    DevState h[2];
    PVD_CUDA(cudaMemcpyAsync(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    const long long n = h[s->parity].n;
    const int nc = s->nc;
    PVD_CUDA(s->stage.alloc((size_t)n * nc * 8));
    k_soa_to_aos<<<grid_for(n * nc, 256, 16), 256, 0, s->stream>>>(s->impx_y.as<double>(), s->stage.as<double>(), n, nc, s->cap);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpyAsync(xyz_out, s->stage.p, (size_t)n * nc * 8, cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    *n_out = n;
    s->ext_moved = true;
    return PVD_OK;
}

int pvd_sim_imp_ext_accept(pvd_sim *s, const double *fy, const double *psiy, const double *secy, int64_t n, const double *u_metro)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->cfg.trial == PVD_TRIAL_EXTERNAL && s->ext_moved && fy && psiy && secy, "pvd_sim_imp_ext_accept: call pvd_sim_imp_ext_propose first");
    const int nc = s->nc;
    PVD_CUDA(cudaMemcpyAsync(s->impx_fy.p, fy, (size_t)n * nc * 8, cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaMemcpyAsync(s->impx_sec.p, secy, (size_t)n * nc * 8, cudaMemcpyHostToDevice, s->stream));
    PVD_CUDA(cudaMemcpyAsync(s->impx_psiy.p, psiy, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    const double *um = nullptr;
    if (u_metro) {
        if (!s->inj_um.p) PVD_CUDA(s->inj_um.alloc((size_t)s->cap * 8));
        PVD_CUDA(cudaMemcpyAsync(s->inj_um.p, u_metro, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
        um = s->inj_um.as<double>();
    }
    StepArgs a = make_args(s, 1);
    k_impx_accept<<<grid_for(s->cap, 256, 8), 256, 0, s->stream>>>(a, s->impx_invm.as<double>(), s->x[s->cur].as<double>(), s->f[s->cur].as<double>(),
                                                                  s->psi[s->cur].as<double>(), s->lk[s->cur].as<double>(), s->impx_y.as<double>(),
                                                                  s->impx_fy.as<double>(), s->impx_psiy.as<double>(), s->impx_sec.as<double>(), um,
                                                                  s->acc_count.as<unsigned long long>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

int pvd_sim_imp_ext_finish(pvd_sim *s, const double *v, int64_t n, int32_t do_branch, const double *u_branch)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(s->cfg.trial == PVD_TRIAL_EXTERNAL && s->ext_moved, "pvd_sim_imp_ext_finish: call pvd_sim_imp_ext_propose / _accept first");
    PVD_REQUIRE(v || s->cfg.potential != PVD_POT_EXTERNAL, "pvd_sim_imp_ext_finish: an external potential needs its energies");
    if (v) PVD_CUDA(cudaMemcpyAsync(s->v[s->cur].p, v, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    else if (int rc = launch_pot_soa(s)) return rc;
    k_impx_add_lk<<<grid_for(s->cap, 256, 8), 256, 0, s->stream>>>(s->st.as<DevState>(), s->parity, s->v[s->cur].as<double>(), s->lk[s->cur].as<double>());
    PVD_CHECK_LAUNCH();
    StepArgs a = make_args(s, do_branch);
    if (u_branch) {
        PVD_CUDA(cudaStreamSynchronize(s->stream));
        DevState h0[2];
        PVD_CUDA(cudaMemcpy(h0, s->st.p, sizeof(h0), cudaMemcpyDeviceToHost));
        if (!s->inj_u.p) PVD_CUDA(s->inj_u.alloc((size_t)s->cap * 8));
        PVD_CUDA(cudaMemcpyAsync(s->inj_u.p, u_branch, (size_t)h0[s->parity].n * 8, cudaMemcpyHostToDevice, s->stream));
        a.inj_u = s->inj_u.as<double>();
    }
    if (s->cfg.world_size > 1) {
        // sharded: k_impx_accept published (accepted, walkers) of this shard, the caller all-reduced them: global acceptance -> dt_eff
        k_imp_set_dt<<<1, 32, 0, s->stream>>>(s->st.as<DevState>(), s->parity, a.sums, s->cfg.delta_t);
        PVD_CHECK_LAUNCH();
    }
    if (int rc = imp_enqueue_branch(s, a)) return rc;
    s->parity ^= 1;
    s->ext_moved = false;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    return PVD_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- walker rebalancing between shards
int pvd_sim_export_tail(pvd_sim *s, int64_t count, double *xyz, double *pots, double *w, int64_t *who)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    PVD_REQUIRE(count >= 0 && count < n && xyz && pots, "pvd_sim_export_tail: bad count");
    const int nc = s->nc;
    const long long first = n - count;
    PVD_CUDA(s->stage.alloc((size_t)count * nc * 8));
    // SoA tail -> AoS: treat the tail as its own SoA array with the same stride
    k_soa_to_aos<<<grid_for(count * nc, 256, 16), 256, 0, s->stream>>>(s->x[s->cur].as<double>() + first, s->stage.as<double>(), count, nc, s->cap);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpyAsync(xyz, s->stage.p, (size_t)count * nc * 8, cudaMemcpyDeviceToHost, s->stream));
    PVD_CUDA(cudaMemcpyAsync(pots, s->v[s->cur].as<double>() + first, (size_t)count * 8, cudaMemcpyDeviceToHost, s->stream));
    if (w && s->w.p) PVD_CUDA(cudaMemcpyAsync(w, s->w.as<double>() + first, (size_t)count * 8, cudaMemcpyDeviceToHost, s->stream));
    if (who) {
        PVD_CUDA(s->stage2.alloc((size_t)count * 8));
        k_int_to_i64<<<grid_for(count, 256, 16), 256, 0, s->stream>>>(s->who[s->cur].as<int>() + first, s->stage2.as<long long>(), count);
        PVD_CHECK_LAUNCH();
        PVD_CUDA(cudaMemcpyAsync(who, s->stage2.p, (size_t)count * 8, cudaMemcpyDeviceToHost, s->stream));
    }
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    h[0].n = h[1].n = first;
    h[1 - s->parity] = h[s->parity];
    h[0].n = first; h[1].n = first;
    PVD_CUDA(cudaMemcpy(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice));
    return PVD_OK;
}

int pvd_sim_import(pvd_sim *s, int64_t count, const double *xyz, const double *pots, const double *w, const int64_t *who)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    PVD_REQUIRE(count >= 0 && n + count <= s->cap && xyz && pots, "pvd_sim_import: not enough capacity");
    const int nc = s->nc;
    PVD_CUDA(s->stage.alloc((size_t)count * nc * 8));
    PVD_CUDA(cudaMemcpyAsync(s->stage.p, xyz, (size_t)count * nc * 8, cudaMemcpyHostToDevice, s->stream));
    k_aos_to_soa<<<grid_for(count * nc, 256, 16), 256, 0, s->stream>>>(s->stage.as<double>(), s->x[s->cur].as<double>() + n, count, nc, s->cap);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpyAsync(s->v[s->cur].as<double>() + n, pots, (size_t)count * 8, cudaMemcpyHostToDevice, s->stream));
    if (w && s->w.p) PVD_CUDA(cudaMemcpyAsync(s->w.as<double>() + n, w, (size_t)count * 8, cudaMemcpyHostToDevice, s->stream));
    if (who) {
        std::vector<int> w32((size_t)count);
        for (int64_t i = 0; i < count; ++i) w32[(size_t)i] = (int)who[i];
        PVD_CUDA(cudaMemcpy(s->who[s->cur].as<int>() + n, w32.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    }
    // importance sampling: drift, psi and local kinetic energy are functions of the coordinates alone: rebuild them
    if (s->cfg.trial != PVD_TRIAL_NONE && count > 0)
        if (int rc = imp_initial_drift(s, n, count)) return rc;
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    h[1 - s->parity] = h[s->parity];
    h[0].n = n + count; h[1].n = n + count;
    PVD_CUDA(cudaMemcpy(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice));
    return PVD_OK;
}

// ---- the same transfers without the host: walkers are packed into / unpacked from one device buffer, which travels GPU to GPU
// (NCCL send / recv over NVLink on the caller's side).  Row layout (float64): x[nc] | V | w | who_from | then, with importance
// sampling, f_x[nc] | psi | T_L (| vector score) -- the companions travel with their walker instead of being recomputed.
static int payload_cols(const pvd_sim *s)
{
    int c = s->nc + 3;
    if (s->cfg.trial != PVD_TRIAL_NONE) c += s->nc + 2 + (s->cfg.imp_variant == PVD_IMP_EXCITED_STATE ? 1 : 0);
    return c;
}

struct PackArgs {
    double *x, *v, *w, *f, *psi, *lk, *vs;
    int *who;
    long long cap;
    int nc, ncols, imp;
};

__global__ void k_pack_walkers(PackArgs a, long long first, long long count, double *payload, int unpack)
{
    const long long total = count * a.ncols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / a.ncols;
        const int j = (int)(e - i * a.ncols);
        const long long wk = first + i;
        double *p = payload + e;
        if (j < a.nc) { double *q = a.x + (long long)j * a.cap + wk; if (unpack) *q = *p; else *p = *q; continue; }
        int k = j - a.nc;
        if (k == 0) { if (unpack) a.v[wk] = *p; else *p = a.v[wk]; continue; }
        if (k == 1) { if (a.w) { if (unpack) a.w[wk] = *p; else *p = a.w[wk]; } else if (!unpack) *p = 1.0; continue; }
        if (k == 2) { if (unpack) a.who[wk] = (int)*p; else *p = (double)a.who[wk]; continue; }
        k -= 3;
        if (k < a.nc) { double *q = a.f + (long long)k * a.cap + wk; if (unpack) *q = *p; else *p = *q; continue; }
        k -= a.nc;
        double *q = k == 0 ? a.psi + wk : k == 1 ? a.lk + wk : a.vs + wk;
        if (unpack) *q = *p; else *p = *q;
    }
}

static PackArgs pack_args(pvd_sim *s)
{
    PackArgs a{};
    a.x = s->x[s->cur].as<double>(); a.v = s->v[s->cur].as<double>(); a.w = s->w.as<double>(); a.who = s->who[s->cur].as<int>();
    a.cap = s->cap; a.nc = s->nc; a.ncols = payload_cols(s);
    a.imp = s->cfg.trial != PVD_TRIAL_NONE;
    if (a.imp) { a.f = s->f[s->cur].as<double>(); a.psi = s->psi[s->cur].as<double>(); a.lk = s->lk[s->cur].as<double>(); a.vs = s->vs[s->cur].as<double>(); }
    return a;
}

int pvd_sim_export_tail_device(pvd_sim *s, int64_t count, void **payload_dev, int32_t *ncols)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(payload_dev && ncols, "NULL argument");
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    PVD_REQUIRE(count >= 0 && count < n, "pvd_sim_export_tail_device: bad count");
    const PackArgs a = pack_args(s);
    PVD_CUDA(s->xfer.alloc((size_t)count * a.ncols * 8));
    k_pack_walkers<<<grid_for(count * a.ncols, 256, 16), 256, 0, s->stream>>>(a, n - count, count, s->xfer.as<double>(), 0);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    h[1 - s->parity] = h[s->parity];
    h[0].n = n - count; h[1].n = n - count;
    PVD_CUDA(cudaMemcpy(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice));
    *payload_dev = s->xfer.p;
    *ncols = a.ncols;
    return PVD_OK;
}

int pvd_sim_import_device(pvd_sim *s, int64_t count, const void *payload_dev, int32_t ncols)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_CUDA(cudaDeviceSynchronize());            // the payload was written by the caller's stream (NCCL recv)
    DevState h[2];
    PVD_CUDA(cudaMemcpy(h, s->st.p, sizeof(h), cudaMemcpyDeviceToHost));
    const long long n = h[s->parity].n;
    const PackArgs a = pack_args(s);
    PVD_REQUIRE(payload_dev && count >= 0 && n + count <= s->cap, "pvd_sim_import_device: not enough capacity");
    PVD_REQUIRE(ncols == a.ncols, "pvd_sim_import_device: payload layout does not match this simulation");
    k_pack_walkers<<<grid_for(count * a.ncols, 256, 16), 256, 0, s->stream>>>(a, n, count, (double *)payload_dev, 1);
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaStreamSynchronize(s->stream));
    h[1 - s->parity] = h[s->parity];
    h[0].n = n + count; h[1].n = n + count;
    PVD_CUDA(cudaMemcpy(s->st.p, h, sizeof(h), cudaMemcpyHostToDevice));
    return PVD_OK;
}

}  // extern "C"

#include "pvd_nn_host.inl"

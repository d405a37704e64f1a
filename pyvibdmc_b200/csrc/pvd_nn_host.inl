// NN potential: host glue.
static int nn_enqueue_discrete_step(pvd_sim *s, StepArgs &a)
{
    // move (in place) -> descriptor + MLP -> branch-only step
    const int g = s->grid_light;
    double *x = s->x[s->cur].as<double>();
    if (a.inj_disp) {
        k_displace_soa<PVD_RNG_FP64><<<g, 256, 0, s->stream>>>(x, s->st.as<DevState>(), s->parity, 0, 0, s->cap, s->nc, s->cfg.ndim, s->cfg.seed,
                                                               a.inj_disp, nullptr, s->sigma_dev.as<double>(), nullptr);
    } else if (s->cfg.rng_mode == PVD_RNG_FAST)
        k_displace_soa<PVD_RNG_FAST><<<g, 256, 0, s->stream>>>(x, s->st.as<DevState>(), s->parity, 0, 0, s->cap, s->nc, s->cfg.ndim, s->cfg.seed,
                                                               nullptr, nullptr, s->sigma_dev.as<double>(), nullptr);
    else if (s->cfg.rng_mode == PVD_RNG_ZIGGURAT)
        k_displace_soa<PVD_RNG_ZIGGURAT><<<g, 256, 0, s->stream>>>(x, s->st.as<DevState>(), s->parity, 0, 0, s->cap, s->nc, s->cfg.ndim, s->cfg.seed,
                                                                   nullptr, nullptr, s->sigma_dev.as<double>(), nullptr);
    else
        k_displace_soa<PVD_RNG_FP64><<<g, 256, 0, s->stream>>>(x, s->st.as<DevState>(), s->parity, 0, 0, s->cap, s->nc, s->cfg.ndim, s->cfg.seed,
                                                               nullptr, nullptr, s->sigma_dev.as<double>(), nullptr);
    PVD_CHECK_LAUNCH();
    if (int rc = nn_launch(s->stream, x, 1, s->cap, s->st.as<DevState>(), s->parity, 0, s->cap, s->v[s->cur].as<double>(), s->nn_w))
        return rc;
    k_branch_discrete<<<s->grid_light, PVD_CTA, 0, s->stream>>>(a);
    return PVD_OK;
}

extern "C" {

int pvd_sim_set_nn_weights(pvd_sim *s, const float *packed, int64_t nfloats)
{
    SIM_CHECK(s);
    SIM_DEVICE(s);
    PVD_REQUIRE(packed && nfloats == NN_NPARAM, "expected 31081 packed float32 weights");
    return nn_upload_weights(s->nn_w, packed);
}

/* which kernel evaluates the network (process-wide; negative = leave as it is): path 0 tcgen05 with two tiles in flight
 * (default), 1 tcgen05 with one tile, 2 float32 FMA on CUDA cores (float32-accurate cross-check); terms 3 | 4 cross terms of
 * the fp16 split; threads 512 | 1024 per CTA of the default path.  The environment (PVD_NN_FP32 / _TC1 / _TERMS / _THREADS) only
 * sets the initial values, once. */
int pvd_nn_config(int32_t path, int32_t terms, int32_t threads)
{
    PVD_REQUIRE(path <= 2 && (terms < 0 || terms == 3 || terms == 4) && (threads < 0 || threads == 512 || threads == 1024), "pvd_nn_config: bad arguments");
    (void)nn_cfg();
    if (path >= 0) g_nn_cfg.path = path;
    if (terms >= 0) g_nn_cfg.terms = terms;
    if (threads >= 0) g_nn_cfg.threads = threads;
    return PVD_OK;
}

int pvd_nn_h4o2_set_weights(const float *packed, int64_t nfloats)
{
    PVD_REQUIRE(packed && nfloats == NN_NPARAM, "expected 31081 packed float32 weights");
    if (int rc = ensure_device_ready()) return rc;
    return nn_upload_weights(g_nn, packed);
}

int pvd_nn_h4o2(const double *xyz, int64_t n, double *v)
{
    PVD_REQUIRE(n >= 0 && (n == 0 || (xyz && v)), "pvd_nn_h4o2: bad arguments");
    if (int rc = ensure_device_ready()) return rc;
    if (!g_nn.packed) return pvd_fail(PVD_E_STATE, "pvd_nn_h4o2: call pvd_nn_h4o2_set_weights first");
    if (n == 0) return PVD_OK;
    int rc_launch = PVD_OK;
    const int rc = run_host_kernel(xyz, (size_t)n * 18 * 8, v, (size_t)n * 8, [&](void *in, void *out) {
        rc_launch = nn_launch(nullptr, (const double *)in, 0, 0, nullptr, 0, n, n, (double *)out, g_nn);
        g_pvd_launches.fetch_sub(1);      // run_host_kernel counts the launch itself
    });
    return rc_launch ? rc_launch : rc;
}

int pvd_coulomb_descriptor(const double *xyz, int64_t n, int32_t natoms, const double *z, double *desc)
{
    PVD_REQUIRE(xyz && desc && z && n >= 0 && natoms == 6, "pvd_coulomb_descriptor: built for the 6-atom water dimer");
    PVD_REQUIRE(z[0] == 8 && z[1] == 1 && z[2] == 1 && z[3] == 8 && z[4] == 1 && z[5] == 1, "charges must be [8,1,1,8,1,1]");
    if (int rc = ensure_device_ready()) return rc;
    if (n == 0) return PVD_OK;
    DevBuf dx, dv, dd, dw;
    PVD_CUDA(dx.alloc((size_t)n * 18 * 8)); PVD_CUDA(dv.alloc((size_t)n * 8)); PVD_CUDA(dd.alloc((size_t)n * NN_IN * 8));
    PVD_CUDA(dw.alloc((size_t)NN_NPARAM * 4));
    PVD_CUDA(cudaMemset(dw.p, 0, (size_t)NN_NPARAM * 4));
    PVD_CUDA(cudaMemcpy(dx.p, xyz, (size_t)n * 18 * 8, cudaMemcpyHostToDevice));
    if (int rc = nn_prepare_launch()) return rc;
    k_nn_h4o2<<<grid_for(n, NN_TILE, 3), NN_THREADS, NN_SMEM_BYTES>>>(dx.as<double>(), 0, 0, nullptr, 0, n, dw.as<float>(), dv.as<double>(), dd.as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpy(desc, dd.p, (size_t)n * NN_IN * 8, cudaMemcpyDeviceToHost));
    return PVD_OK;
}

int pvd_distit(const double *xyz, int64_t n, int32_t natoms, int32_t method, const double *pair_scale, const double *diag,
               const double *r_eq, const int32_t *atom_lists, const int32_t *atom_list_ofs, int32_t n_atom_lists,
               const int32_t *groups, int32_t ngroups, int32_t gsize, int32_t full_mat, double *out)
{
    PVD_REQUIRE(n >= 0 && natoms >= 2 && natoms <= PVD_MAX_ATOMS && (n == 0 || (xyz && out)), "pvd_distit: bad arguments");
    PVD_REQUIRE(method >= PVD_DESC_DISTANCE && method <= PVD_DESC_SPF, "pvd_distit: method is 0 (distance), 1 (coulomb) or 2 (spf)");
    PVD_REQUIRE(method != PVD_DESC_COULOMB || (pair_scale && diag), "pvd_distit: coulomb needs pair_scale and diag");
    PVD_REQUIRE(n_atom_lists >= 0 && ngroups >= 0 && gsize >= 0 && ngroups * gsize <= natoms, "pvd_distit: bad sorting lists");
    PVD_REQUIRE(n_atom_lists == 0 || (atom_lists && atom_list_ofs && atom_list_ofs[0] == 0 && atom_list_ofs[n_atom_lists] == natoms),
                "pvd_distit: sorted_atoms must hold every atom once");
    PVD_REQUIRE(ngroups == 0 || groups, "pvd_distit: NULL groups");
    const bool sort = n_atom_lists > 0 || ngroups > 0;
    PVD_REQUIRE(method != PVD_DESC_SPF || r_eq || (full_mat && !sort), "pvd_distit: spf needs r_eq");
    if (int rc = ensure_device_ready()) return rc;
    if (n == 0) return PVD_OK;
    const int npairs = natoms * (natoms - 1) / 2;
    static thread_local DistitParams P;          // 6 KB: not on the stack
    memset(&P, 0, sizeof(P));
    P.natoms = natoms; P.method = method; P.full_mat = full_mat ? 1 : 0; P.sort = sort ? 1 : 0;
    P.n_atom_lists = n_atom_lists; P.ngroups = ngroups; P.gsize = gsize;
    for (int k = 0; k < (n_atom_lists ? natoms : 0); ++k) {
        PVD_REQUIRE(atom_lists[k] >= 0 && atom_lists[k] < natoms, "pvd_distit: atom index out of range");
        P.atom_lists[k] = atom_lists[k];
    }
    for (int k = 0; k <= n_atom_lists && n_atom_lists; ++k) P.atom_list_ofs[k] = atom_list_ofs[k];
    for (int k = 0; k < ngroups * gsize; ++k) {
        PVD_REQUIRE(groups[k] >= 0 && groups[k] < natoms, "pvd_distit: atom index out of range");
        P.groups[k] = groups[k];
    }
    if (method == PVD_DESC_COULOMB) {
        for (int k = 0; k < npairs; ++k) P.pair_scale[k] = pair_scale[k];
        for (int k = 0; k < natoms; ++k) P.diag[k] = diag[k];
    }
    if (method == PVD_DESC_SPF && r_eq)
        for (int k = 0; k < (sort ? natoms * natoms : npairs); ++k) P.r_eq[k] = r_eq[k];
    const size_t per = full_mat ? (size_t)natoms * natoms : (size_t)npairs;
    DevBuf dx, dp, dout;
    PVD_CUDA(dx.alloc((size_t)n * natoms * 3 * 8)); PVD_CUDA(dp.alloc(sizeof(P))); PVD_CUDA(dout.alloc((size_t)n * per * 8));
    PVD_CUDA(cudaMemcpy(dx.p, xyz, (size_t)n * natoms * 3 * 8, cudaMemcpyHostToDevice));
    PVD_CUDA(cudaMemcpy(dp.p, &P, sizeof(P), cudaMemcpyHostToDevice));
    k_distit<<<grid_for(n, 128, 16), 128>>>(dx.as<double>(), n, dp.as<DistitParams>(), dout.as<double>());
    PVD_CHECK_LAUNCH();
    PVD_CUDA(cudaMemcpy(out, dout.p, (size_t)n * per * 8, cudaMemcpyDeviceToHost));
    return PVD_OK;
}

}  // extern "C"

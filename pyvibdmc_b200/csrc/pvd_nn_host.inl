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

}  // extern "C"

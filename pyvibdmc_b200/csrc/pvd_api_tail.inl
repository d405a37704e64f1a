// Host-side pieces that need the full pvd_sim definition.
static int cont_enqueue_step(pvd_sim *, StepArgs &) { return pvd_fail(PVD_E_STATE, "continuous weighting: not built yet"); }
static int cont_enqueue_branch_only(pvd_sim *, StepArgs &) { return pvd_fail(PVD_E_STATE, "continuous weighting: not built yet"); }
static int imp_enqueue_step(pvd_sim *, StepArgs &, const double *) { return pvd_fail(PVD_E_STATE, "importance sampling: not built yet"); }
static int imp_initial_drift(pvd_sim *) { return pvd_fail(PVD_E_STATE, "importance sampling: not built yet"); }
static int nn_enqueue_discrete_step(pvd_sim *, StepArgs &) { return pvd_fail(PVD_E_STATE, "NN potential: not built yet"); }

extern "C" {
#define PVD_TODO(name) return pvd_fail(PVD_E_STATE, name ": not built yet")
int pvd_branch_continuous(double *, const double *, int64_t, double, double, double, double, int64_t *, double *) { PVD_TODO("pvd_branch_continuous"); }
int pvd_trial_drift(int32_t, const double *, int64_t, int32_t, int32_t, const double *, int64_t, double *, double *, double *) { PVD_TODO("pvd_trial_drift"); }
int pvd_metropolis(const double *, const double *, const double *, const double *, const double *, const double *, int64_t, int32_t, int32_t, const double *, const double *, double, double *) { PVD_TODO("pvd_metropolis"); }
int pvd_local_kin(const double *, int64_t, int32_t, int32_t, const double *, double *) { PVD_TODO("pvd_local_kin"); }
int pvd_nn_h4o2_set_weights(const float *, int64_t) { PVD_TODO("pvd_nn_h4o2_set_weights"); }
int pvd_nn_h4o2(const double *, int64_t, double *) { PVD_TODO("pvd_nn_h4o2"); }
int pvd_coulomb_descriptor(const double *, int64_t, int32_t, const double *, double *) { PVD_TODO("pvd_coulomb_descriptor"); }
int pvd_sim_set_trial_table(pvd_sim *, const double *, int64_t) { PVD_TODO("pvd_sim_set_trial_table"); }
int pvd_sim_set_nn_weights(pvd_sim *, const float *, int64_t) { PVD_TODO("pvd_sim_set_nn_weights"); }
int pvd_sim_download_imp(pvd_sim *, double *, double *, double *, int64_t) { PVD_TODO("pvd_sim_download_imp"); }
int pvd_sim_export_tail(pvd_sim *, int64_t, double *, double *, double *, int64_t *) { PVD_TODO("pvd_sim_export_tail"); }
int pvd_sim_import(pvd_sim *, int64_t, const double *, const double *, const double *, const int64_t *) { PVD_TODO("pvd_sim_import"); }
}

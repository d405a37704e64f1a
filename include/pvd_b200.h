/* pvd_b200.h -- C ABI of the B200-native PyVibDMC walker-propagation path.
 *
 * One shared library (pyvibdmc_b200/_lib/libpvd_b200.so, sm_100a) exports exactly the entry
 * points declared here.  Signatures use plain pointers and sizes only (no torch types); the
 * Python host (pyvibdmc_b200/_capi.py) binds them with ctypes.  Each entry point cites the
 * reference interface it replaces (paths relative to /root/reference/pyvibdmc).
 *
 * Conventions
 *   - all arrays are float64 C-contiguous in the reference's layouts: coordinates are
 *     "AoS" (n, natoms, ndim); per-walker scalars are (n,).
 *   - every function returns 0 on success, a PVD_E_* code otherwise; pvd_last_error()
 *     returns a human-readable message for the calling thread's last failure.
 *   - "host" entry points take host pointers and do H2D / kernel / D2H internally (they are the
 *     drop-in plug-in calls); pvd_sim_* entry points keep the walker ensemble resident in HBM.
 *   - a simulation handle is not thread-safe; use one host thread per handle.
 */
#ifndef PVD_B200_H
#define PVD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVD_ABI_VERSION 1

/* error codes */
enum {
    PVD_OK = 0,
    PVD_E_CUDA = 1,        /* a CUDA runtime call failed (message has the CUDA error string) */
    PVD_E_ARG = 2,         /* invalid argument */
    PVD_E_MASSIVE = 3,     /* "Massive walker birth or death event" (pyvibdmc.py:400,413,717) */
    PVD_E_STATE = 4,       /* call not valid in the handle's current state */
    PVD_E_NODEVICE = 5     /* no CUDA device available */
};

/* built-in potentials (replace the user functions called through Potential.getpot,
 * simulation_utilities/potential_manager.py:71-99) */
enum {
    PVD_POT_EXTERNAL = 0,  /* V supplied by the host each step (arbitrary user Python potential) */
    PVD_POT_HARMONIC = 1,  /* sum_c k_c x_c^2 : sample_potentials/PythonPots/harmonicOscillator1D.py:5-49 */
    PVD_POT_H2O_PS = 2,    /* Partridge-Schwenke water: FortPots/Partridge_Schwenke_H2O/calc_h2o_pot.f + h2opes_v2.f */
    PVD_POT_MORSE1D = 3,   /* De (1-exp(-a x))^2 : sample_potentials/PythonPots/morse_osc_1d.py:4-12 */
    PVD_POT_NN_H4O2 = 4    /* Coulomb descriptor + 15-120-120-120-1 MLP: TensorflowPots/call_sample_model.py:4-9 */
};

/* built-in trial wave functions (replace ImpSampManager.call_trial / call_derivs,
 * simulation_utilities/imp_samp_manager.py:92-139,197-224) */
enum {
    PVD_TRIAL_NONE = 0,
    PVD_TRIAL_HARM1D = 1,  /* analytic Gaussian + derivatives: PythonPots/harm_trial_wfn.py:6-40 */
    PVD_TRIAL_H2O_FD = 2,  /* water product wfn, finite-difference derivatives:
                              FortPots/.../call_trl_h2o.py:63-78 + imp_samp.py:56-76; table = [grid | psi | alpha, theta_eq] */
    PVD_TRIAL_H2O_AN = 3,  /* same wfn, analytic derivatives: call_trl_h2o.py:101-149 (dpsi_dx) + imp_samp_helper.py:10-209;
                              table = [grid | psi | psi' | psi'' | alpha, theta_eq] */
    PVD_TRIAL_EXTERNAL = 4 /* ANY user trial wave function / derivative function: the host answers ImpSampManager.call_trial /
                              call_derivs (imp_samp_manager.py:92-139, 197-224) once per step, see pvd_sim_imp_ext_* */
};

enum { PVD_WEIGHT_DISCRETE = 0, PVD_WEIGHT_CONTINUOUS = 1 };
/* how the Gaussian displacements of move_randomly (pyvibdmc.py:540-547, np.random.normal) are generated from Philox bits:
 * FP64 = Box-Muller evaluated in double; FAST = Box-Muller with float32 SFU intrinsics (~1e-6 relative);
 * ZIGGURAT = 1024-layer ziggurat in double (exact distribution, the method of NumPy's Generator.normal) */
enum { PVD_RNG_FP64 = 0, PVD_RNG_FAST = 1, PVD_RNG_ZIGGURAT = 2 };
#define PVD_ZIGGURAT_LAYERS 1024

#define PVD_MAX_ATOMS 16
#define PVD_MAX_COMP (3 * PVD_MAX_ATOMS)

/* ------------------------------------------------------------------ library / device */
int pvd_abi_version(void);
const char *pvd_last_error(void);
int pvd_device_count(int *count);
int pvd_set_device(int device);
/* FP64 FMA throughput of the current device in FLOP/s, measured with a register-resident
 * DFMA chain (roofline denominator for the fp64 potential kernels). */
int pvd_measure_fp64_peak(double *flops_per_s);
/* Device time (ms) of the most recent host entry point's kernel, measured with CUDA events. */
int pvd_last_kernel_ms(double *ms);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t pvd_launch_count(void);

/* ------------------------------------------------------------------ stand-alone host entry points
 * (the plug-in surface: what a maintainer binds behind Potential.getpot / ImpSamp.*) */

/* replaces water_pot(cds) = calc_hoh_pot(cds, len(cds))  (h2o_potential.py:6-7, calc_h2o_pot.f:1-34).
 * xyz: (n,3,3) bohr, atoms ordered H,H,O ; v: (n,) Hartree */
int pvd_pes_h2o(const double *xyz, int64_t n, double *v);
/* folded PS parameters as uploaded to the device: c[245], scal[8] = reoh,b1,ce,phh1,phh2,deoh,roh,alphaoh */
int pvd_pes_h2o_params(double *c245, double *scal8);

/* replaces oh_stretch_harm-style functions (harmonicOscillator1D.py:13-17):
 * v[i] = sum_c k[c] * x[i,c]^2 with k[c] = (0.5*mass)*omega^2 computed by the caller. x: (n,ncomp) */
int pvd_pes_harmonic(const double *x, int64_t n, int32_t ncomp, const double *k, double *v);
/* replaces oh_stretch_morse (morse_osc_1d.py:4-12): v = de * (1 - exp(-alpha x))^2 ; x: (n,) */
int pvd_pes_morse1d(const double *x, int64_t n, double de, double alpha, double *v);

/* replaces DMC_Sim.move_randomly (pyvibdmc.py:540-547): xyz[i,a,:] += N(0, sigma[a]) in place.
 * Normals come from Philox4x32-10 keyed by `seed`, counter (walker index, step). */
int pvd_displace(double *xyz, int64_t n, int32_t natoms, int32_t ndim, const double *sigma,
                 uint64_t seed, uint64_t step, int32_t rng_mode);
/* raw generator output for statistical tests: z[i,c] exactly as pvd_displace would add with sigma=1 */
int pvd_normals(double *z, int64_t n, int32_t ncomp, uint64_t seed, uint64_t step, int32_t rng_mode);
/* the ziggurat table the device uses (host-computed at load): x[0..LAYERS] layer abscissae (x[0] = V/f(R), x[1] = R,
 * x[LAYERS] = 0), f[i] = exp(-x[i]^2/2).  Needs no device: lets a test re-derive the generator on the CPU. */
int pvd_ziggurat_table(double *x, double *f);
/* Philox4x32-10 known-answer hook: out[4] = philox(ctr[4], key[2]) computed on the device */
int pvd_philox_kat(const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4);

/* replaces the discrete branch of DMC_Sim.birth_or_death (pyvibdmc.py:391-431) given the
 * caller's uniforms u (injection mode).  counts: (n,) int32 ; idx: (capacity,) int64 receives
 * np.repeat(arange(n), counts) ; stats[3] = births, deaths, final_pop.  Returns PVD_E_MASSIVE when
 * the reference would raise (non-finite / too large weight, population outside [0.5,1.5] n0). */
int pvd_branch_discrete(const double *v, int64_t n, double vref, double dt, const double *u, int64_t n0,
                        int32_t *counts, int64_t *idx, int64_t idx_capacity, int64_t *stats3);

/* replaces the continuous branch of birth_or_death + _branch (pyvibdmc.py:432-454, 340-356).
 * w: (n,) in/out ; src: (n,) int64 out, src[i] = walker whose coords/V/who_from walker i holds
 * afterwards ; upper = NaN for "no upper threshold" ; stats[3] = n_branched, max_w, min_w (after). */
int pvd_branch_continuous(double *w, const double *v, int64_t n, double vref, double dt,
                          double lower, double upper, int64_t *src, double *stats3);

/* replaces calc_vref (pyvibdmc.py:651-661); w may be NULL (discrete). */
int pvd_calc_vref(const double *v, const double *w, int64_t n, int64_t n0, double alpha, double *vref);
/* replaces calc_desc_wts (pyvibdmc.py:663-672); w NULL -> discrete counts. out: (n_parent,) */
int pvd_desc_wts(const int64_t *who_from, const double *w, int64_t n, int64_t n_parent, double *out);

/* replaces ImpSamp.drift with finite differences (imp_samp.py:21-26,56-76 + manager division
 * by psi, imp_samp_manager.py:117-120) for a built-in trial wfn.
 * psi: (n,) ; dlog: (n,natoms,ndim) = grad psi / psi ; d2: same shape = d2psi/dx2 / psi.
 * table: trial-specific parameters (H2O: 2 x ntab doubles = grid row then psi row; HARM1D: {alpha}). */
int pvd_trial_drift(int32_t trial, const double *xyz, int64_t n, int32_t natoms, int32_t ndim,
                    const double *table, int64_t ntab, double *psi, double *dlog, double *d2);
/* replaces ImpSamp.metropolis (imp_samp.py:29-47) + local_kin (:50-53), for any (natoms x ndim) like the reference
 * (natoms <= PVD_MAX_ATOMS; 3 x 3 and 1 x 1 run unrolled kernels, every other shape a run-time one, same arithmetic).
 * acc: (n,) acceptance ratios ; sigma, inv_mass: (natoms,) */
int pvd_metropolis(const double *x, const double *y, const double *fx, const double *fy,
                   const double *psi_x, const double *psi_y, int64_t n, int32_t natoms, int32_t ndim,
                   const double *sigma, const double *inv_mass, double dt, double *acc);
int pvd_local_kin(const double *d2, int64_t n, int32_t natoms, int32_t ndim, const double *inv_mass, double *ke);

/* replaces sample_h4o2_pot (call_sample_model.py:4-9): Coulomb descriptor (distance_descriptors.py
 * :102-113,154-168) + Dense(120,swish)x3 + Dense(1,relu), cm^-1 -> Hartree.
 * weights: packed float32 [W0(15x120) b0(120) W1(120x120) b1 W2(120x120) b2 W3(120) b3(1)] row-major (in,out).
 * xyz: (n,6,3) bohr, atoms O,H,H,O,H,H. */
int pvd_nn_h4o2_set_weights(const float *packed, int64_t nfloats);
/* kernel selection for the network (negative = keep): path 0 tcgen05 / two tiles in flight (default), 1 tcgen05 / one tile,
 * 2 float32 FMA on CUDA cores (float32-accurate cross-check); terms 3 | 4; threads 512 | 1024.  Looked up at launch from a
 * process-wide setting, never from the environment (which only provides the initial values). */
int pvd_nn_config(int32_t path, int32_t terms, int32_t threads);
int pvd_nn_h4o2(const double *xyz, int64_t n, double *v);
int pvd_coulomb_descriptor(const double *xyz, int64_t n, int32_t natoms, const double *z, double *desc);
/* replaces DistIt.run (simulation_utilities/tensorflow_descriptors/distance_descriptors.py:177-213; helpers :88-168) for any
 * molecule of up to PVD_MAX_ATOMS atoms.  xyz: (n, natoms, 3).  method: 0 distance, 1 coulomb, 2 spf.
 * pair_scale[npairs] = Z_i Z_j and diag[natoms] = 0.5 Z^2.4 (coulomb only; pairs in itertools.combinations order);
 * r_eq: distances of the equilibrium structure, per pair when nothing is sorted, else the natoms x natoms matrix of the
 * equilibrium structure sorted the same way (spf only).  sorted_atoms as a flattened list + offsets (n_atom_lists + 1),
 * sorted_groups as ngroups x gsize.  full_mat = 0: out is (n, npairs), upper triangle; 1: (n, natoms, natoms).
 * Bit-identical to the reference's NumPy arithmetic, summation order included: the group sort (:146) adds the squares of a
 * column in np.sum's unrolled pairwise order from 8 atoms on and the group totals left to right -- pairwise when n == 1,
 * as NumPy does for a single walker, so pass the batch the reference would see (r_eq comes from TWO copies of eq_xyz, :75).
 * Ties between exactly equal norms keep index order (the reference's order then depends on the host's SIMD argsort). */
int pvd_distit(const double *xyz, int64_t n, int32_t natoms, int32_t method, const double *pair_scale, const double *diag,
               const double *r_eq, const int32_t *atom_lists, const int32_t *atom_list_ofs, int32_t n_atom_lists,
               const int32_t *groups, int32_t ngroups, int32_t gsize, int32_t full_mat, double *out);

/* ------------------------------------------------------------------ device-resident simulation
 * replaces the state + per-step body of DMC_Sim.propagate (pyvibdmc.py:701-876). */
typedef struct pvd_sim pvd_sim;   /* opaque */

enum { PVD_IMP_STANDARD = 0,            /* imp_move_randomly (pyvibdmc.py:549-612) */
       PVD_IMP_SECOND_DISPLACEMENT = 1, /* imp_move_randomly_second_type (pyvibdmc.py:614-649) */
       PVD_IMP_EXCITED_STATE = 2        /* excited_state_imp_samp: capped drift + vector score (pyvibdmc.py:562-591, 608-611, 810-811) */ };

typedef struct {
    int32_t natoms, ndim;
    int32_t weighting;            /* PVD_WEIGHT_* */
    int32_t potential;            /* PVD_POT_* */
    int32_t trial;                /* PVD_TRIAL_* */
    int32_t rng_mode;             /* PVD_RNG_*; the importance-sampling move draws ZIGGURAT requests with FP64 Box-Muller */
    int32_t device;               /* CUDA device ordinal */
    int32_t rank, world_size;     /* shard id / number of shards (multi-GPU); 0,1 for a single GPU */
    int64_t num_walkers;          /* N0: target population of THIS shard's share is num_walkers/world_size */
    int64_t capacity;             /* walker slots per buffer on this device (>= 1.5*N0_local + slack) */
    double delta_t;
    double alpha;                 /* 1/(2 dt) unless DEBUG_alpha (pyvibdmc.py:200-203) */
    double thresh_lower, thresh_upper;   /* continuous weighting; upper = NaN when absent */
    uint64_t seed;
    double masses[PVD_MAX_ATOMS];
    double pot_params[PVD_MAX_COMP]; /* HARMONIC: k[c]; MORSE1D: de, alpha */
    int64_t stats_ring;           /* length of the per-step statistics ring (>= steps between drains) */
    int32_t imp_variant;          /* PVD_IMP_*: which importance-sampling move (pyvibdmc.py:549-612 or :614-649) */
    int32_t reserved_;
} pvd_config;

/* per-step record written by the step kernel's finalisation (one per executed step) */
typedef struct {
    double vref;                  /* after calc_vref */
    double pop;                   /* len(walkers) (discrete) or sum(w) (continuous), global */
    double v_avg, v_max, v_min;   /* of the energies used for weighting, before branching (log lines) */
    double w_max, w_min;          /* continuous: after branching */
    double dt_eff;                /* effective time step used (imp-samp) */
    int64_t births, deaths;       /* discrete ; continuous: births = n_branched */
    int64_t rejected;             /* imp-samp Metropolis rejections */
    int64_t step;                 /* propagation step index this record belongs to */
} pvd_step_stats;

/* sizeof(pvd_config) / sizeof(pvd_step_stats) as this library was compiled: lets a binding check its struct layouts */
int pvd_sizeof_config(void);
int pvd_sizeof_step_stats(void);

int pvd_sim_create(const pvd_config *cfg, pvd_sim **out);
int pvd_sim_destroy(pvd_sim *s);
/* use an externally owned stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = legacy default stream.
 * Until this is called the handle runs on its own non-blocking stream. */
int pvd_sim_set_stream(pvd_sim *s, void *cuda_stream);

/* upload the start ensemble (DMC_Sim._initialize, pyvibdmc.py:155-169,214-219); w may be NULL.
 * Also evaluates V on the start ensemble and the first Vref (first-step exception, :760-769). */
int pvd_sim_upload(pvd_sim *s, const double *xyz, int64_t n, const double *w);
/* PVD_POT_EXTERNAL only: energies of the uploaded start ensemble, computed by the caller's getpot */
int pvd_sim_set_pots(pvd_sim *s, const double *v, int64_t n);
/* multi-GPU: after pvd_sim_upload (+ all-reduce of the sums buffer) compute the first Vref */
int pvd_sim_init_finalize(pvd_sim *s);
/* trial-wfn parameters for importance sampling (same meaning as pvd_trial_drift's table) */
int pvd_sim_set_trial_table(pvd_sim *s, const double *table, int64_t ntab);
/* trainable-potential weights for PVD_POT_NN_H4O2 */
int pvd_sim_set_nn_weights(pvd_sim *s, const float *packed, int64_t nfloats);

/* enqueue `nsteps` whole time steps (move -> V -> weight/branch -> Vref -> record) without
 * host synchronisation; branch_mask_every = branch_every (pyvibdmc.py:828-837). */
int pvd_sim_run(pvd_sim *s, int64_t nsteps, int32_t branch_every);
/* How a pvd_sim_run / pvd_sim_run_mailbox segment of discrete-weighting steps with a built-in potential is executed.
 * enable = 1 (default), by ensemble size: up to ~300 000 walkers per GPU ONE resident kernel that overlaps consecutive
 * time steps (csrc/pvd_run.cuh); above, one launch per step of the deferred-compaction step (csrc/pvd_gather.cuh: the
 * np.repeat gather of birth_or_death, pyvibdmc.py:415-431, is done by the next step's loads, so no warp waits for
 * another) plus one materialisation at the end of the segment.  2: the resident kernel always; 3: the deferred-compaction
 * step always (pvd_sim_step_injected uses it too: parity tests); 0: one self-compacting launch per time step.  All four
 * give the same numbers bit for bit.  Replaces the loop `for prop_step in range(...)` of DMC_Sim.propagate
 * (pyvibdmc.py:703) for the steps between two host-side events. */
int pvd_sim_set_resident(pvd_sim *s, int32_t enable);
/* one step with injected random numbers (parity tests): disp (n,natoms,ndim) already scaled by
 * sigma, u_branch (n_after_move,) for birth/death, u_metro (n,) for Metropolis (may be NULL). */
int pvd_sim_step_injected(pvd_sim *s, const double *disp, const double *u_branch, const double *u_metro);
/* external-potential stepping: move, hand coordinates to the host, take V back and weight/branch */
int pvd_sim_ext_move(pvd_sim *s, double *xyz_out, int64_t *n_out);
int pvd_sim_ext_finish(pvd_sim *s, const double *v, int64_t n, int32_t do_branch);
/* Device-tensor potential plug-in (the reference's NN_Potential exists so that user models run on GPUs,
 * potential_manager.py:177-214): the same per-step contract as pvd_sim_ext_move / _finish, but the coordinates handed to the
 * callable and the energies it returns stay in HBM.  *xyz_dev: float64 (n, atoms, dims), C order, on the simulation's device,
 * owned by the handle and valid until the next call on it (wrap it with __cuda_array_interface__ / DLPack, do not free it);
 * v_dev: float64 (n) on the same device.  pvd_sim_coords_device: the current walkers without a move (first-step energies);
 * pvd_sim_set_pots_device: device counterpart of pvd_sim_set_pots. */
int pvd_sim_ext_move_device(pvd_sim *s, void **xyz_dev, int64_t *n_out);
int pvd_sim_ext_finish_device(pvd_sim *s, const void *v_dev, int64_t n, int32_t do_branch);
int pvd_sim_coords_device(pvd_sim *s, void **xyz_dev, int64_t *n_out);
int pvd_sim_set_pots_device(pvd_sim *s, const void *v_dev, int64_t n);

/* multi-GPU split step: local part, then the caller all-reduces `sums` (device pointer to
 * PVD_NSUMS doubles, obtained from pvd_sim_sums_ptr) over NCCL, then finalisation. */
#define PVD_MAX_WORLD 8
#define PVD_NSUMS (16 + 4 * PVD_MAX_WORLD)
int pvd_sim_sums_ptr(pvd_sim *s, void **device_ptr);
/* use a caller-owned device buffer (e.g. a torch tensor NCCL can reduce) of PVD_NSUMS doubles instead */
int pvd_sim_set_sums_ptr(pvd_sim *s, void *device_ptr);
int pvd_sim_step_local(pvd_sim *s, int32_t do_branch);
int pvd_sim_step_finalize(pvd_sim *s);
/* importance sampling across shards: the acceptance fraction that scales the time step (pyvibdmc.py:372-378, 603) is
 * global, so the local step is cut after the Metropolis move:
 *   pvd_sim_imp_move_local -> all-reduce(sums) -> pvd_sim_imp_branch_local -> all-reduce(sums) -> pvd_sim_step_finalize */
int pvd_sim_imp_move_local(pvd_sim *s);
int pvd_sim_imp_branch_local(pvd_sim *s, int32_t do_branch);

/* the same exchange without a collective kernel: the step kernel's last CTA stores the shard's sums into every peer's
 * mailbox over NVLink (CUDA IPC mappings), a one-warp kernel waits for the world's stamps and finalises.
 *   every rank: pvd_sim_mailbox_handle(out 64 bytes) -> all-gather the handles -> pvd_sim_mailbox_connect(all, world)
 *   then pvd_sim_run_mailbox(nsteps) enqueues whole time steps with no host or NCCL involvement */
int pvd_sim_mailbox_handle(pvd_sim *s, void *handle64);
int pvd_sim_mailbox_connect(pvd_sim *s, const void *handles, int32_t n);
int pvd_sim_run_mailbox(pvd_sim *s, int64_t nsteps, int32_t branch_every);

/* descendant weighting (pyvibdmc.py:739-747, 663-672, 856-869) */
int pvd_sim_dw_begin(pvd_sim *s, int64_t global_offset);
/* dmc_restart inside an open window (pyvibdmc.py:299-338: _who_from / _parent / _parent_wts travel in the checkpoint): who_from
 * (n) of the uploaded walkers and the parent ensemble (n_parent, atoms, dims) [+ parent weights] back onto the device. */
int pvd_sim_dw_resume(pvd_sim *s, const int64_t *who_from, int64_t n, const double *parent_xyz, const double *parent_w, int64_t n_parent);
int pvd_sim_dw_end(pvd_sim *s, double *desc_wts, int64_t n_parent);
/* The same without stopping the loop (north_star: "wavefunction dumps are async D2H on a side stream"): _begin counts the
 * descendant weights, copies the parent ensemble and closes the window in stream order, and starts their transfer to pinned host
 * memory on the side stream; the caller enqueues the next time steps and collects the arrays with _wait, which waits for the
 * transfer only.  Replaces the synchronous save at pyvibdmc.py:856-872. */
int pvd_sim_dw_end_begin(pvd_sim *s, int64_t n_parent);
int pvd_sim_dw_end_wait(pvd_sim *s, double *desc_wts, double *parent_xyz, double *parent_w, int64_t n_parent);
/* calc_desc_wts without closing the window (DEBUG_save_desc_wt_tracker, pyvibdmc.py:849-852) */
int pvd_sim_dw_peek(pvd_sim *s, double *desc_wts, int64_t n_parent);
/* DEBUG_mass_change (pyvibdmc.py:749-753): sigma = sqrt(dt / m) from new masses */
int pvd_sim_set_masses(pvd_sim *s, const double *masses, int32_t natoms);
int pvd_sim_dw_parent(pvd_sim *s, double *xyz, double *w, int64_t *n_parent);

/* blocking queries */
int pvd_sim_sync(pvd_sim *s);
int pvd_sim_state(pvd_sim *s, int64_t *n, double *vref, int64_t *step, int32_t *err);
/* asynchronous snapshot (checkpoints / dumps, pyvibdmc.py:729-736, 861-872): _begin copies the walkers as they are now (in
 * stream order) and sends them to pinned host memory on a side stream; the compute stream can be given the next time steps
 * right away; _wait blocks only on the side-stream copy.  With all outputs NULL, _wait returns the walker count and leaves
 * the snapshot pending. */
int pvd_sim_snapshot_begin(pvd_sim *s);
int pvd_sim_snapshot_wait(pvd_sim *s, double *xyz, double *pots, double *w, int64_t *who_from, int64_t capacity, int64_t *n_out,
                          double *vref_out);
int pvd_sim_download(pvd_sim *s, double *xyz, double *pots, double *w, int64_t *who_from, int64_t capacity, int64_t *n);
int pvd_sim_stats(pvd_sim *s, int64_t first_step, int64_t count, pvd_step_stats *out);
/* device time in ms between the first and last kernel of the most recent pvd_sim_run */
int pvd_sim_last_run_ms(pvd_sim *s, double *ms);
/* importance-sampling per-walker arrays (f_x, psi, sec) for checkpoints */
int pvd_sim_download_imp(pvd_sim *s, double *fx, double *psi, double *sec, int64_t capacity);

/* ---- importance sampling with a USER trial wave function (config.trial = PVD_TRIAL_EXTERNAL; any atoms x dims).
 * Replaces imp_move_randomly (pyvibdmc.py:549-612) around the plug-in call impsamp.drift(cds) (imp_samp.py:21-27), which the
 * host makes once per step:
 *   pvd_sim_upload -> pvd_sim_imp_ext_init(f_x, psi, psi''/psi of the start ensemble [, V])       first-step exception (:553-554, 760-769)
 *   per step: pvd_sim_imp_ext_propose -> displaced coordinates (n, atoms, dims)                   (:556-559, 593)
 *             host: f_y, psi_2, sec_y = impsamp.drift(displaced)                                  (:595)
 *             pvd_sim_imp_ext_accept(f_y, psi_2, sec_y)   Metropolis (imp_samp.py:29-47), acceptance, dt_eff (:597-612)
 *             pvd_sim_imp_ext_finish(V or NULL, n, do_branch)   E_L = V + T_L (:807-809), weighting / branching, Vref
 * V = NULL evaluates the configured built-in potential on the GPU; with PVD_POT_EXTERNAL the caller downloads the accepted
 * coordinates (pvd_sim_download) and passes getpot's result.  disp / u_metro / u_branch: injected displacements (n, atoms, dims,
 * already scaled by sigma), Metropolis and birth/death uniforms -- replays of reference trajectories in the parity tests -- or NULL.
 * Sharded (world_size > 1): every call works on this rank's walkers; the caller all-reduces the shared sums (pvd_sim_set_sums_ptr)
 * after _init (then pvd_sim_init_finalize), after _accept (global acceptance fraction -> dt_eff, pyvibdmc.py:372-378) and after
 * _finish (then pvd_sim_step_finalize). */
int pvd_sim_imp_ext_init(pvd_sim *s, const double *fx, const double *psi, const double *sec, const double *v_or_null);
int pvd_sim_imp_ext_propose(pvd_sim *s, const double *disp_or_null, double *xyz_out, int64_t *n_out);
int pvd_sim_imp_ext_accept(pvd_sim *s, const double *fy, const double *psi_y, const double *sec_y, int64_t n, const double *u_metro_or_null);
int pvd_sim_imp_ext_finish(pvd_sim *s, const double *v_or_null, int64_t n, int32_t do_branch, const double *u_branch_or_null);
/* walker rebalancing between shards: remove the last `count` walkers into host/peer buffers, or append */
int pvd_sim_export_tail(pvd_sim *s, int64_t count, double *xyz, double *pots, double *w, int64_t *who);
int pvd_sim_import(pvd_sim *s, int64_t count, const double *xyz, const double *pots, const double *w, const int64_t *who);
/* The same transfers GPU to GPU (SURVEY 8e "periodic P2P walker rebalancing"): the last `count` walkers are packed into one device
 * buffer of `ncols` float64 per walker -- x[atoms*dims] | V | w | who_from | and, with importance sampling, f_x[atoms*dims] | psi |
 * T_L (| vector score): the companions travel with their walker -- which the caller sends over NVLink (NCCL send / recv on device
 * pointers, no host staging) and the receiver appends with pvd_sim_import_device.  *payload_dev is owned by the handle and valid
 * until the next call on it. */
int pvd_sim_export_tail_device(pvd_sim *s, int64_t count, void **payload_dev, int32_t *ncols);
int pvd_sim_import_device(pvd_sim *s, int64_t count, const void *payload_dev, int32_t ncols);

#ifdef __cplusplus
}
#endif
#endif /* PVD_B200_H */

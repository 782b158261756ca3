/* ssb.h — C-ABI of the B200-native ssa_sdpd engine (libssb_core.so + one model unit per model).
 *
 * The reference engine has NO foreign-function interface: its ABI is a process contract
 * (`ssa_sdpd.exe -s SEED -t THREADS`, cwd = result directory, exit code; E/propensity_file_template.cpp:101-142,
 * spatialpy/solvers/solver.py:553-597, E = spatialpy/solvers/c_base/ssa_sdpd-c-simulation-engine) with every
 * model input baked in as C++ literals (solver.py:100-158).  This header is the in-process replacement of
 * that contract: plain pointers and sizes, no C++ or torch types, all functions return 0 on success and a
 * non-zero code on failure (the Python side raises SimulationError("Solver execution failed, return code = N"),
 * matching solver.py:595-597).  No function calls exit() or throws across the boundary.
 *
 * Ownership: every pointer inside `ssb_model` is BORROWED for the duration of ssb_create() only (the engine
 * copies to device memory); the library never frees caller memory.  One handle is single-threaded; distinct
 * handles may run concurrently on distinct GPUs.
 */
#ifndef SSB_H
#define SSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_ABI_VERSION 4

/* error codes */
#define SSB_OK 0
#define SSB_ERR_NAN 1          /* NaN/Inf in x, v or rho — reference: check_particle_nan() exit(1), E/src/particle.cpp:88-126 */
#define SSB_ERR_RDME 2         /* negative population / propensity overflow — reference exit(1) sites, E/src/simulate_rdme.cpp:279,293,345,388,399,467 */
#define SSB_ERR_CUDA 3
#define SSB_ERR_ARG 4
#define SSB_ERR_IO 5
#define SSB_ERR_CANCELLED 6    /* ssb_cancel() — reference: SIGINT to the process group on timeout, solver.py:579-586 */
#define SSB_ERR_MODEL_UNIT 7   /* model unit missing / compiled for different sizes */
#define SSB_ERR_HALO 8         /* slab decomposition: a neighbouring rank's halo message did not arrive within the device-side timeout */

/* flags (ssb_model.flags) */
#define SSB_FLAG_CORRECTED_NSM_SELECT 1u   /* draw the reaction/diffusion channel with the textbook NSM rule instead of the reference's
                                              rand1*srrate rule (simulate_rdme.cpp:253-261,317-321) */
#define SSB_FLAG_CORRECTED_STOICH 2u       /* index the dense stoichiometry as N[s][rxn] instead of the reference's transposed/out-of-bounds
                                              read (E/src/model.cpp:186-187) */
#define SSB_FLAG_NO_VTK 4u                 /* stage outputs (ssb_get_output) but do not write outputN.vtk files */
#define SSB_FLAG_LITERAL_KERNELS 16u       /* evaluate every pair expression in the reference's literal order (k_force<true>, separate diffusion-matrix
                                              sweep) instead of the restructured sweeps; slower, used by the parity tests as a cross-check */
#define SSB_FLAG_LEAP_DIFFUSION 32u        /* sSSA: advance the diffusion channel per window (binomial jump counts, multinomial destinations)
                                              instead of one event per jump; same law in the windowed scheme, O(1) per species and window */
#define SSB_FLAG_SKIP_STATIC_FORCES 8u     /* static domains: skip F/Fbp/Frho (never consumed when static, simulate.cpp:68,137); default on via Python */
#define SSB_FLAG_NO_STEP_OVERSHOOT 128u    /* moving domains: do not execute the reference's one event past each step's end
                                             (`while(tt <= end_time)` tests the previous event's time, simulate_rdme.cpp:233-238) */
#define SSB_FLAG_BINARY_STORE 64u         /* also write outputN.ssb next to (or, with SSB_FLAG_NO_VTK, instead of) outputN.vtk: the same snapshot as raw
                                             little-endian arrays at full fp64 precision (layout in spatialpy_b200/vtk.py, read by Result.read_step;
                                             SURVEY.md 8f item 1 - the reference's pure-Python ASCII parser, vtkreader.py:29-56, bounds large N*T) */
#define SSB_FLAG_CORRECTED_OUTPUT_STEPS 256u /* ssb_run: write file k at step output_steps[k] (one file per time point) instead of the reference's
                                             gate (simulate_threads.cpp:231-247: files at steps 0, 1, f, 2f, ... plus the final state, i.e. one file
                                             more than time points and file k >= 2 holding step (k-1)*f) */
#define SSB_FLAG_CORRECTED_PDE_INDEX 512u  /* PDE flux: read the species-major diffusion table as [s*num_types + type-1] (the entry
                                             simulate_rdme.cpp:146 uses) instead of the reference's [S_c*(type-1)+s] (model.cpp:163); identical when
                                             the coefficients do not depend on the type */

/* Flat model description.  Replaces the generated-literal inputs of solver.py:100-419. */
typedef struct ssb_model {
    int32_t abi_version;            /* SSB_ABI_VERSION */
    uint32_t flags;
    int64_t n_particles;            /* __NUMBER_OF_VOXELS__ (solver.py:138) */
    int32_t dimension;              /* system->dimension (solver.py:407-413) */
    int32_t static_domain;          /* system->static_domain (solver.py:381) */
    int32_t num_types;              /* ParticleSystem::num_types = len(listOfTypeIDs)-1 (solver.py:379) */
    int32_t num_chem_species;       /* S_c (solver.py:106-111) */
    int32_t num_chem_rxns;
    int32_t num_stoch_species;      /* S_d (solver.py:112-117) */
    int32_t num_stoch_rxns;
    int32_t num_data_fn;
    double dt;                      /* system->dt (solver.py:389) */
    uint32_t nt;                    /* system->nt (solver.py:390) */
    uint32_t n_output_steps;
    const uint32_t *output_steps;   /* get_next_output() table (solver.py:290-299) */
    double h, rho0, c0, P0;         /* solver.py:391-398 */
    double xlo, xhi, ylo, yhi, zlo, zhi;   /* solver.py:400-405 */
    double gravity[3];              /* solver.py:415-417 */
    /* particles, index = id (init_create_particle, solver.py:312-331) */
    const double *x;                /* [N*3] xyz interleaved */
    const int32_t *type;            /* [N], 1-based */
    const double *nu, *mass, *c, *rho;   /* [N] each */
    const int32_t *solid;           /* [N] solidTag = domain.fixed */
    const uint32_t *u0;             /* [N*S] voxel-major input_u0 (solver.py:211-220); also seeds C[] (template:77-81) */
    const double *data_fn;          /* [ndf*N] input_data_fn (solver.py:239-249) */
    const int32_t *N_dense;         /* [S*R] input_N_dense row-major species x rxn (solver.py:223-233) */
    const int64_t *irN, *jcN;       /* CSC of N (solver.py:235-236) */
    const int32_t *prN;             /* (solver.py:237) */
    const int64_t *irG, *jcG;       /* dependency graph CSC, columns [species..., reactions...] (solver.py:256-257) */
    const double *diffusion_matrix; /* [S*num_types] input_subdomain_diffusion_matrix (solver.py:269-286) */
    const char *const *species_names; /* [S] input_species_names (solver.py:259-263) */
    /* sSSA window controller: tau = rdme_epsilon / max_i max_s Ddiag_i[s]  (<=0 selects the default 0.0125;
     * the splitting error of the windowed scheme is first order in it while the spatial distribution relaxes) */
    double rdme_epsilon;
    int32_t device;                 /* CUDA device ordinal */
    int32_t reserved;
    /* spatial slab decomposition (NULL on a single GPU): owned[i] = 1 if this rank integrates particle i, 0 if it is a
     * ghost copy of a particle owned by a neighbouring slab; rng_id[i] = global particle id used as the Philox counter
     * and as the serial order of the reference's particle vector (BVF sweep), so results do not depend on the partition */
    const int32_t *owned;
    const int32_t *rng_id;
} ssb_model;

typedef struct ssb_handle ssb_handle;

/* progress callback: called from the stepping thread after each engine step; return non-zero to cancel */
typedef int (*ssb_progress_cb)(void *user, uint32_t step, uint32_t nt);

/* Library / device */
int ssb_abi_version(void);
int ssb_device_count(int *count);

/* Lifecycle.  ssb_create copies the model to the device (replaces init_all_particles + initialize_rdme,
 * template:134-138).  ssb_load_kernels dlopen()s the per-model unit compiled by codegen (replaces linking
 * the generated model.cpp, E/build/SConstruct). */
int ssb_create(const ssb_model *model, ssb_handle **out);
int ssb_load_kernels(ssb_handle *h, const char *model_unit_path);
int ssb_destroy(ssb_handle *h);

/* Run trajectories first_traj .. first_traj+ntraj-1; trajectory k uses seed+k (solver.py:558-559) and writes
 * output%u.vtk + output0_boundingBox.vtk into out_dirs[k-first_traj] (E/src/output.cpp:104-229).
 * Replaces `exe -s seed -t T` (solver.py:553-569) and run_simulation() (E/src/simulate_threads.cpp:171-312). */
int ssb_run(ssb_handle *h, uint64_t seed, int32_t ntraj, int32_t first_traj,
            const char *const *out_dirs, ssb_progress_cb cb, void *cb_user);

/* Step-wise control (parity taps and benchmarking): reset state to the model's initial condition for one
 * trajectory, then advance n engine steps (each = 3 SDPD substeps + RDME, simulate_threads.cpp:232-281).
 * No files are written by ssb_step. */
int ssb_reset(ssb_handle *h, uint64_t seed);
int ssb_step(ssb_handle *h, uint32_t nsteps);

/* Counters: ParticleSystem::total_reactions / total_diffusion (E/include/particle_system.hpp:79-80) of the most
 * recent trajectory, the wall seconds of its stepping loop, and the number of sSSA windows launched. */
int ssb_counters(ssb_handle *h, int64_t *reactions, int64_t *diffusions, double *seconds, int64_t *windows);

/* Parity taps: copy a named per-particle field, in particle-id order, to caller memory.
 * Names: x v vt F Fbp (f64, N*3) | rho old_rho Frho bvf_phi mass nu srrate sdrate (f64, N) | type solid (i32, N)
 *        C Q (f64, N*S_c, voxel-major) | xx (u32, N*S_d) | Ddiag (f64, N*S_d) | rrate (f64, N*R)
 *        nbr_count (i32, N).  `bytes` is the capacity of `dst`; returns SSB_ERR_ARG on mismatch. */
int ssb_get_field(ssb_handle *h, const char *name, void *dst, int64_t bytes);
/* Neighbour lists in id space (CSR): ptr[N+1] int64, idx[nnz] int32 (neighbour ids), dist/dWdr/Dij[nnz] f64.
 * Call with idx == NULL to obtain nnz in *nnz_out. */
int ssb_get_neighbors(ssb_handle *h, int64_t *ptr, int32_t *idx, double *dist, double *dWdr, double *Dij,
                      int64_t *nnz_out);

/* State hand-over (slab re-partition: the particles a rank owns change, so the state of a trajectory moves between handles).
 * ssb_set_field is the inverse of ssb_get_field for the fields that make up a particle's state between two engine steps —
 * x v vt F Fbp (f64, N*3) | rho old_rho Frho bvf_phi nu (f64, N) | C Q (f64, N*S_c voxel-major) | xx (u32, N*S_d) — given in
 * particle-id order; replacing x invalidates the neighbour lists (rebuilt at the next step).  ssb_get_step / ssb_set_step
 * read and set the engine-step counter (system->current_step, E/include/particle_system.hpp:63) and the Philox window epoch. */
int ssb_set_field(ssb_handle *h, const char *name, const void *src, int64_t bytes);
int ssb_get_step(ssb_handle *h, uint32_t *step, uint64_t *epoch);
int ssb_set_step(ssb_handle *h, uint32_t step, uint64_t epoch);

/* The engine's own file writers (E/src/output.cpp:104-229 byte format; the outputN.ssb side-store) applied to a snapshot in CALLER
 * memory, id order: x v [np*3] | scal = rho,mass,bvf_phi,nu [4*np] | C [Sc*np] species-major | type [np] | xx [Sd*np] species-major;
 * lims6 = xlo xhi ylo yhi zlo zhi (output0_boundingBox.vtk is written with file_index 0 / step 0); what = 1 VTK, 2 binary, 3 both.
 * Used for snapshots assembled on the host (slab-decomposed runs, batched ensembles).  No handle, no device work. */
int ssb_write_snapshot(const char *dir, uint32_t file_index, uint32_t step, int32_t rdme_initialized, int64_t np, int32_t Sc, int32_t Sd,
                       const char *const *species_names, const double *lims6, const double *x, const double *v, const double *scal,
                       const double *C, const int32_t *type, const uint32_t *xx, uint32_t what);

int ssb_cancel(ssb_handle *h);                 /* async-signal-safe flag; the running ssb_run returns SSB_ERR_CANCELLED */
const char *ssb_last_error(ssb_handle *h);     /* message for the last non-zero return (valid until next call) */

/* Kernel accounting for bench.py: number of engine kernels launched since the last ssb_reset/ssb_run start. */
int ssb_launch_count(ssb_handle *h, int64_t *launches);

/* Measurement hooks (bench.py).  ssb_step_timed = ssb_step bracketed by CUDA events on the engine's own stream.
 * ssb_profile(1) brackets every launch group with event pairs; categories:
 * 0 cell list  1 predictor  2 neighbour search  3 force sweep  4 corrector  5 finish/BVF  6 diffusion matrix
 * 7 RDME init  8 sSSA windows  9 output staging.  ssb_io_bytes = bytes copied H2D at reset / D2H by output staging. */
int ssb_step_timed(ssb_handle *h, uint32_t nsteps, double *device_ms);
int ssb_profile(ssb_handle *h, int enable);
int ssb_profile_read(ssb_handle *h, int category, double *ms_total, int64_t *launches);
int ssb_io_bytes(ssb_handle *h, int64_t *h2d, int64_t *d2h);
int ssb_nbr_stats(ssb_handle *h, int32_t *capacity, int64_t *total);

/* Spatial slab decomposition (one process per GPU; spatialpy_b200/slab.py does the NCCL send/recv between the phases).
 * ssb_step_phase runs one piece of an engine step (phases: 0 PRE = cell list + predictor + search + force sweep,
 * 1 CORRECTOR, 2 FINISH, 3 RDME_PREP -> *out = local max Ddiag, 4 RDME_INIT (arg = global max) -> *out = windows per step,
 * 5 RDME_WINDOW (arg = window index), 6 RDME_CLOSE, 7 END, 8 RDME_MIN -> *out = earliest pending event of this rank, 9 RDME_EXTRA
 * (arg = global earliest pending event: the reference's one event past the end of the step, simulate_rdme.cpp:233-238; runs the
 * event window only — follow it with the inbox exchange and RDME_CLOSE with arg < 0, which delivers the molecule if it jumped across
 * a face under the Philox epoch the overshoot reserved for it)).  ssb_halo_pack / ssb_halo_unpack move the field group that a
 * phase produced between storage and a caller-owned DEVICE buffer for the particle ids listed in dev_ids (group 0: F[3]
 * Fbp[3] Frho Q[S_c]; 1: rho_new; 2: v[3] bvf_phi; 3: rho); ssb_halo_inbox_pack reads-and-clears the molecules that jumped
 * into ghost voxels in the last sSSA window, ssb_halo_inbox_add delivers them to the owner.  ssb_mark/ssb_mark_elapsed_ms
 * record CUDA events on the engine's stream. */
int ssb_step_phase(ssb_handle *h, int phase, double arg, double *out);
int ssb_halo_pack(ssb_handle *h, int group, const int32_t *dev_ids, int32_t n, double *dev_out);
int ssb_halo_unpack(ssb_handle *h, int group, const int32_t *dev_ids, int32_t n, const double *dev_in);
int ssb_halo_inbox_pack(ssb_handle *h, const int32_t *dev_ids, int32_t n, uint32_t *dev_out);
int ssb_halo_inbox_add(ssb_handle *h, const int32_t *dev_ids, int32_t n, const uint32_t *dev_in);
int ssb_halo_width(ssb_handle *h, int group, int32_t *width);
/* Native slab transport — the product path of spatial decomposition (one process, or one thread, per GPU; all GPUs of a B200 box are
 * NVLink peers).  The reference has no decomposition (one shared-memory process, E/src/simulate_threads.cpp:171-312); the phase
 * boundaries respected here are its substeps (:232-281).  ssb_slab_setup allocates this rank's RECEIVE WINDOWS (one per slab face;
 * host arrays of local particle ids: send_* = my owned particles that are ghosts on that neighbour, recv_* = my ghosts owned by it,
 * both in the order the neighbour lists them, i.e. by global id) and its scalar board.  ssb_slab_export fills a blob of
 * ssb_slab_blob_bytes() bytes (CUDA IPC handles + raw pointers + process id); the caller gathers the blobs of all ranks in rank
 * order (torch.distributed, MPI, a thread hub — control plane only) and hands them to ssb_slab_connect, which maps the neighbours'
 * windows and every board (cudaIpcOpenMemHandle, or the pointer itself when the rank lives in this process).  ssb_slab_step then runs
 * engine steps whose halo traffic never touches the host: pack kernels store straight into the neighbour's window over NVLink and
 * raise a sequence flag, a one-warp kernel waits for the incoming flag, unpack reads the local window; the three scalar reductions
 * of a step (max Ddiag -> windows per step, earliest pending event -> the step-end overshoot event, step displacement -> travel
 * bound) go through the boards the same way.  It stops early, on all ranks after the same step, once particles may have travelled
 * `travel_limit` (> 0) since ssb_slab_setup / ssb_reset; *done = steps executed, *travel = the bound so far.
 * ssb_slab_disconnect unmaps the neighbours' memory: every rank calls it, all ranks meet, then handles may be destroyed. */
int ssb_slab_setup(ssb_handle *h, int32_t rank, int32_t world, const int32_t *send_lo, int32_t n_send_lo, const int32_t *recv_lo, int32_t n_recv_lo,
                   const int32_t *send_hi, int32_t n_send_hi, const int32_t *recv_hi, int32_t n_recv_hi);
int ssb_slab_blob_bytes(void);
int ssb_slab_export(ssb_handle *h, void *blob, int64_t bytes);
int ssb_slab_connect(ssb_handle *h, const void *blobs, int64_t bytes);
int ssb_slab_disconnect(ssb_handle *h);
int ssb_slab_step(ssb_handle *h, uint32_t nsteps, double travel_limit, uint32_t *done, double *travel);
/* Verlet-skin bookkeeping of moving domains: chosen skin (fraction of h), largest single-step displacement seen, list rebuilds */
int ssb_skin_stats(ssb_handle *h, double *skin, double *step_disp_max, int64_t *rebuilds);
int ssb_mark(ssb_handle *h, int which);
int ssb_mark_elapsed_ms(ssb_handle *h, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* SSB_H */

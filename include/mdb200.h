/*
 * mdb200.h -- C ABI of the B200-native engine for MolecularDynamics.jl's per-step hot path.
 *
 * The reference (Julia, /root/reference) has no FFI today: its boundary is the Julia-level
 * run_simulation!/initialize_state/Parameters API and the Potential + evaluate plugin contract.
 * Each entry point below names the reference code it replaces (paths relative to /root/reference).
 * The Julia-side binding a maintainer adds (`ccall` stubs + GPUSystem) is julia/MolecularDynamicsB200.jl
 * and is walked through in INTEGRATION.md; the Python mirror used by tests/bench is
 * moleculardynamics.jl_b200/ (ctypes).
 *
 * Conventions
 *  - plain C, opaque handle, int status returns (0 = MDB_OK); no exceptions cross the boundary;
 *    mdb_last_error() gives the message of the last failure on that handle (or of mdb_create).
 *  - host arrays are AoS [n][dim] Float64 / Int32 in C order, i.e. exactly
 *    reinterpret(Float64, ::Vector{MVector{dim,Float64}}) (src/initialization.jl:44,93,97,137);
 *    particle order on the host side is always the caller's original order.
 *  - the library owns all device memory; one handle is driven by one host thread.
 *  - there is NO CPU fallback: without a CUDA device mdb_create fails with MDB_ERR_NO_DEVICE.
 */
#ifndef MDB200_H
#define MDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_VERSION 100

typedef struct mdb_engine_s *mdb_handle;

enum mdb_status {
    MDB_OK = 0,
    MDB_ERR_INVALID_ARG = 1,
    MDB_ERR_CUDA = 2,
    MDB_ERR_UNSUPPORTED_POTENTIAL = 3, /* Potential subtype without a device functor (src/types.jl:4-6 errors likewise) */
    MDB_ERR_UNSUPPORTED_CELL = 4,      /* a feature that needs a diagonal unit cell (slab decomposition, user potentials) was
                                          asked for with a general one (to_unitcell matrix branch, src/initialization.jl:13-15) */
    MDB_ERR_BOX_TOO_SMALL = 5,         /* cutoff >= L/2 in a periodic direction */
    MDB_ERR_NO_DEVICE = 6,
    MDB_ERR_NCCL = 7,
    MDB_ERR_STATE = 8,                 /* e.g. run before velocities were set */
    MDB_ERR_NONFINITE = 9,             /* overlap / blow-up detected: non-finite energy */
    MDB_ERR_NVRTC = 10,
    MDB_ERR_IO = 11                    /* frame / checkpoint file could not be written or read */
};

/* Potential subtype -> device functor tag (src/potentials.jl, README.md:82-145) */
enum mdb_potential {
    MDB_POT_PSEUDOHS = 0, /* PseudoHS: no parameters (lambda = 50 fixed, src/potentials.jl:11-29) */
    MDB_POT_LJ = 1,       /* LennardJones: params = {epsilon, r_cut} (src/potentials.jl:160-164, 66-77) */
    MDB_POT_LJ_XPLOR = 2, /* LennardJonesXPLOR: params = {epsilon, r_on, r_cut} (src/potentials.jl:217-249) */
    MDB_POT_POLY = 3,     /* non-additive Polydisperse plugin: params = {rcut, non_additivity} (README.md:89-145) */
    MDB_POT_SOFT = 4,     /* overlap-removal penalty u = k/2 (1 - r/tol)^2, r < tol: params = {k, tol}; stands in for Packmol's
                             pack_monoatomic!(coordinates, maxs, tol) together with mdb_fire_minimize (src/initialization.jl:20-30) */
    MDB_POT_USER = 100    /* user CUDA-C `evaluate` body compiled with NVRTC (mdb_set_user_potential) */
};

/* Ensemble subtypes (src/types.jl:34-51) */
enum mdb_ensemble { MDB_NVE = 0, MDB_NVT = 1, MDB_BROWNIAN = 2 };

/* Neighbour strategy.  Results are identical (same pair set, same per-pair arithmetic); only cost differs. */
enum mdb_mode {
    MDB_MODE_AUTO = 0,
    MDB_MODE_CELLS = 1, /* cell list rebuilt and particles re-sorted EVERY step (the reference's shape, src/simulation.jl:100) */
    MDB_MODE_LIST = 2,  /* Verlet list with skin, rebuilt from the cell list when the displacement bound is reached */
    MDB_MODE_SMALL = 3  /* n <= 4096: the whole step loop runs inside one persistent CTA, thousands of steps per launch
                           (MDB_MODE_AUTO picks it for such systems; other entry points keep using cells / list) */
};

typedef struct mdb_config {
    int32_t dim;            /* 2 or 3                       (initialize_state(dimension=...), src/initialization.jl:116) */
    int32_t potential;      /* enum mdb_potential           (Parameters.potential, src/types.jl:12) */
    int64_t n_particles;    /* global particle count        (Parameters.n_particles, src/types.jl:10) */
    double unitcell[9];     /* row-major 3x3, upper-left dim x dim used: the matrix of SimulationState.unitcell, lattice vectors
                               in its columns (x = U*frac, src/boundary.jl:7-17); diagonal (per-axis arithmetic) or general /
                               triclinic (fractional-coordinate wrap, nearest image and cell grid) */
    double cutoff;          /* neighbour cutoff             (initialize_state(cutoff=1.5), src/initialization.jl:118) */
    double pot_params[8];   /* see enum mdb_potential */
    uint64_t seed;          /* key of the counter-based RNG (replaces SimulationState.rng, src/types.jl:21) */
    int32_t device;         /* CUDA device ordinal */
    int32_t mode;           /* enum mdb_mode */
    double skin;            /* Verlet skin for MDB_MODE_LIST; <= 0 picks a default */
    int32_t use_graph;      /* 1: replay each step as a CUDA graph (conditional rebuild node); 0: eager launches */
    int32_t rank;           /* slab decomposition along x: this handle owns x in [rank, rank+1) * Lx / nranks */
    int32_t nranks;         /* 1 = single domain */
    int32_t no_fuse;        /* 0 (default): NVE runs in list mode fuse the next step's kick-drift into the force kernel
                               (bit-identical results, one sweep less per step); 1: keep the reference's kernel order */
    double skin_inner;      /* skin of the inner (tight) list derived from the Verlet list; <= 0 picks a default */
    int32_t slab_transport; /* nranks > 1: how ghost columns / migrants / reductions travel.  0 default (mdb_comm_init: peer memory
                               over NVLink with cudaIpc-mapped mailboxes, NCCL send/recv if they cannot be mapped;
                               mdb_comm_init_local: device copies), 1 classic (NCCL send/recv + all-reduce, or device copies),
                               2 peer memory (in the in-process ring too: the same kernels, on one device) */
    int32_t reserved;
} mdb_config;

typedef struct mdb_stats {
    int64_t steps;            /* MD steps executed since creation */
    int64_t rebuilds;         /* cell/list rebuilds */
    int64_t kernel_launches;  /* kernels launched by this library (graph kernel nodes counted per replay) */
    int64_t n_owned;          /* particles owned by this handle */
    int64_t n_ghost;          /* ghost particles currently held (nranks > 1) */
    int64_t list_capacity;    /* neighbour slots per particle (MDB_MODE_LIST) */
    int64_t max_neighbors;    /* largest neighbour count seen at the last rebuild */
    double r_search;          /* min(cutoff, potential range): radius the force kernels search */
    double cell_len[3];
    int32_t ncell[3];
    int32_t mode;             /* resolved mode */
    double last_run_ms;       /* device time of the last mdb_run_* call (CUDA events) */
    double last_force_ms;     /* device time of the pair-force kernel of the last stand-alone force evaluation */
    /* eager mode (use_graph = 0) doubles as the profiling mode: per-kernel CUDA-event totals over the last run call */
    double prof_kick_ms;      /* K5 kick-drift-wrap (or K8 Brownian move) */
    double prof_force_ms;     /* K4 pair-force kernel (with the fused second kick) */
    double prof_rebuild_ms;   /* K1-K3 (+ list build) when they ran */
    int64_t prof_steps;
    int64_t slab_transport;   /* last run: 0 single domain, 1 in-process ring (device copies), 2 NCCL send/recv, 3 peer memory */
    int64_t slab_graph;       /* last run: 1 = the slab step replayed as one CUDA graph (peer memory), 2 = NCCL graph experiment */
} mdb_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int mdb_version(void);
/* message of the most recent failure (per handle; h == NULL: last failure of mdb_create on this thread) */
const char *mdb_last_error(mdb_handle h);
/* replaces CellListMap.ParticleSystem(...) construction at src/initialization.jl:100-107 + SimulationState (:140-142) */
int mdb_create(const mdb_config *cfg, mdb_handle *out);
int mdb_destroy(mdb_handle h);

/* ---- state transfer (host arrays in original particle order) --------------------------------------------- */
/* positions/diameters required; velocities, forces, images may be NULL (zeros: src/initialization.jl:97-99,137).
 * Positions are wrapped into the unit cell on upload with wrap_to_box arithmetic (src/boundary.jl:7-17), images
 * accumulate the crossings, so x + U*img is preserved.  With nranks > 1 every rank passes the GLOBAL arrays and
 * keeps the particles of its slab. */
int mdb_upload(mdb_handle h, const double *positions, const double *velocities, const double *forces,
               const double *diameters, const int32_t *images);
/* nranks > 1, after one mdb_upload has planned the slab: hand back the rows of the particles this rank owns (what
 * mdb_download_owned returned: original indices + rows), without touching the global arrays.  The next force evaluation
 * rebuilds the neighbour structures and migrates whatever the host moved across a slab face. */
int mdb_upload_owned(mdb_handle h, int64_t count, const int32_t *ids, const double *positions, const double *velocities,
                     const double *forces, const double *diameters, const int32_t *images);
/* state.velocities = initialize_velocities(...) assignment (README.md:38-41) */
int mdb_set_velocities(mdb_handle h, const double *velocities);
/* any pointer may be NULL.  nranks == 1: arrays are [n_particles][dim] in original order.
 * nranks > 1: use mdb_download_owned. */
int mdb_download(mdb_handle h, double *positions, double *velocities, double *forces, int32_t *images);
/* owned particles of this rank, in device slot order, with their original indices; arrays sized for capacity rows;
 * *count receives the number written. */
int mdb_download_owned(mdb_handle h, int64_t capacity, int32_t *ids, double *positions, double *velocities,
                       double *forces, int32_t *images, int64_t *count);

/* ---- the force path alone ----------------------------------------------------------------------------- */
/* reset_output! + CellListMap.map_pairwise!(energy_and_forces!) (src/simulation.jl:99-104, src/pairwise.jl:26-39,
 * src/minimize.jl:71-72): fills the resident forces; returns energy, virial and the number of pairs that passed
 * the potential's own range test (each unordered pair once; rank-local share when nranks > 1). */
int mdb_compute_forces(mdb_handle h, double *energy, double *virial, int64_t *n_pairs);
/* debug: number of unordered pairs with minimum-image d2 <= cutoff^2 (what map_pairwise! would visit), and optionally
 * each particle's neighbour count in original order (nranks == 1 only). */
int mdb_count_pairs(mdb_handle h, double cutoff, int64_t *n_pairs, int32_t *per_particle);

/* ---- the step loop (src/simulation.jl:88-108 and :231-250) ---------------------------------------------- */
/* thermo, when not NULL, receives one row per step: {U, W, KE, n_pairs} = potential energy, virial, kinetic energy after
 * the ensemble step (0 for Brownian), interacting pairs.  Forces are NOT primed before step 0. */
/* integrate_half! -> forces -> integrate_second_half! -> ensemble_step!(::NVE)  (src/integrate.jl:8-44) */
int mdb_run_nve(mdb_handle h, int64_t nsteps, double dt, double *thermo);
/* ... -> ensemble_step!(::NVT): bussi! with T0 = ktemp_per_step[s] = ensemble.ktemp(step+1) (src/integrate.jl:46-53,
 * src/thermostat.jl:20-48); nf = dim*(N-1) as src/initialization.jl:124 */
int mdb_run_nvt(mdb_handle h, int64_t nsteps, double dt, const double *ktemp_per_step, double tau, double *thermo);
/* forces -> integrate_brownian! (src/simulation.jl:231-250, src/integrate.jl:66-82; intended semantics) */
int mdb_run_brownian(mdb_handle h, int64_t nsteps, double dt, double ktemp, double *thermo);
/* {U, W, KE, n_pairs} of the most recent force evaluation / step */
int mdb_thermo(mdb_handle h, double out[4]);

/* ---- FIRE minimiser (src/minimize.jl:31-135), second caller of the force path ----------------------------- */
typedef struct mdb_fire_params {
    int64_t max_steps;
    double tol, dt_initial, dt_max, alpha0, f_inc, f_dec;
    int32_t n_min;
    int32_t reserved;
} mdb_fire_params;
/* out: {energy, F_rms, steps_done}; *converged = 1 when F_norm/sqrt(ndof) < tol */
int mdb_fire_minimize(mdb_handle h, const mdb_fire_params *p, double out[3], int32_t *converged);

/* ---- thermostat test hooks --------------------------------------------------------------------------- */
/* scale factor of bussi! (src/thermostat.jl:38-43) evaluated ON THE DEVICE from injected noises */
int mdb_bussi_scale_from(mdb_handle h, double kinetic_energy, double ktemp, double nf, double dt, double tau, double r1,
                         double r2, double *scale);
/* the (r1, r2) = (randn, sum_noises(nf-1)) the device draws for RNG step `step` (src/thermostat.jl:1-18,35-36) */
int mdb_bussi_noises(mdb_handle h, uint64_t step, double nf, double *r1, double *r2);
/* engine-wide step counter that keys the RNG; run calls advance it */
int mdb_get_rng_step(mdb_handle h, uint64_t *step);
int mdb_set_rng_step(mdb_handle h, uint64_t step);

/* ---- user-defined Potential (src/types.jl:1-6 plugin contract) on the device, compiled with NVRTC --------- */
/* `body` is the CUDA-C body of
 *   __device__ bool evaluate(double r, double sigma1, double sigma2, const double* p, double& u, double& f)
 * returning true when the pair interacts (false with u = f = 0 otherwise).  params fills p[0..6] (n_params <= 7);
 * `range` = largest r at which it can return true.  The body is compiled with NVRTC (sm_100a, -fmad=false) into the same
 * kernel templates the built-in potentials use (kernels.cuh next to the library).  A comment containing MDB_DENSE_HITS
 * selects in-line evaluation instead of the deferred-hit queue (for potentials where most pairs in range interact).
 * Must be called before mdb_upload; sets cfg.potential = MDB_POT_USER.  Compile errors: MDB_ERR_NVRTC + compiler log. */
int mdb_set_user_potential(mdb_handle h, const char *body, const double *params, int32_t n_params, double range);

/* ---- multi-GPU slabs (new; the reference is single-process) --------------------------------------------- */
/* 128-byte NCCL unique id: rank 0 calls mdb_comm_unique_id and ships it to the others (torch.distributed broadcast) */
int mdb_comm_unique_id(char id[128]);
/* join the slab ring: ncclCommInitRank.  The communicator bootstraps the peer-memory transport (the ranks' mailboxes are
 * exchanged as cudaIpc handles with one ncclAllGather at the first run after every mdb_upload) and carries the chunk-end
 * ncclAllReduce of the thermo rows; with mdb_config.slab_transport = 1 (or when the mailboxes cannot be mapped) ghost +
 * migration exchange use ncclSend/ncclRecv and the per-step consensus ncclAllReduce */
int mdb_comm_init(mdb_handle h, const char id[128]);
/* in-process ring of handles on the same device (tests / single-GPU emulation of the slab protocol; no NCCL) */
int mdb_comm_init_local(mdb_handle *handles, int32_t count);

/* ---- trajectory frames: the step after the path --------------------------------------- */
/* Replaces the reads of positions and images by write_to_file_lammps (src/io.jl:78-170) at its call sites in the step loop
 * (src/simulation.jl:139-171).  mdb_frame_capture packs, on the device and in ORIGINAL particle order, one record per
 * particle {radius = diameter/2, x[dim] wrapped, xu[dim] = x + U*img (unwrapped, src/io.jl:62-70)} into frame slot
 * `slot` (0 .. MDB_FRAME_SLOTS-1) and starts its copy to pinned host memory on a second stream; it returns without waiting,
 * so the step loop continues while the frame travels.
 * nranks > 1 (rank-local calls, no collective): the frame holds the particles the rank OWNS, in device slot order, with
 * their original ids beside them; mdb_frame_write_lammps writes "<path>.<rank>", a complete LAMMPS dump of those atoms
 * (ids = original index + 1) -- the one-file-per-processor layout of LAMMPS' `dump ... file.%`; mdb_frame_wait returns the
 * rows of the owned particles (mdb_stats.n_owned of them at capture time). */
#define MDB_FRAME_SLOTS 2
int mdb_frame_capture(mdb_handle h, int32_t slot);
/* waits for the copy; *frame = library-owned pinned array [n_particles][*width], *width = 2*dim + 1; valid until the next
 * capture into the slot */
int mdb_frame_wait(mdb_handle h, int32_t slot, const double **frame, int32_t *width);
/* queues the frame for a background thread of the library that formats it exactly like write_to_file_lammps (same header
 * lines and "%lf" columns, src/io.jl:97-167; append != 0 is mode="a") and returns at once; a later capture into the slot
 * waits until the file is written */
int mdb_frame_write_lammps(mdb_handle h, int32_t slot, const char *path, int64_t step, int32_t append);
/* blocks until every queued frame is on disk; reports the first I/O error (MDB_ERR_IO) */
int mdb_frame_flush(mdb_handle h);

/* ---- device-side set-up and exact restart: the step before the path -------------------- */
/* state.velocities = initialize_velocities(ktemp, rng, N, dim) (src/initialization.jl:32-47) on the device: standard normals
 * from the counter-based RNG keyed by (seed, original particle id, stream), centre-of-mass motion removed, rescaled so
 * that sum(v^2) / ((N-1) dim) == ktemp.  nranks > 1: collective over the slab ring; every rank draws for the particles it
 * owns (same normals as a single domain, keyed by particle id), the two global sums are all-reduced. */
int mdb_init_velocities(mdb_handle h, double ktemp, uint64_t stream);
/* initialize_random (src/initialization.jl:20-30) on the device, first half: replaces the resident positions by uniform
 * random points of the cell (counter-based RNG keyed by seed, particle id, stream), resets the images.  Second half: a
 * handle created with MDB_POT_SOFT {k, tol} removes the overlaps with mdb_fire_minimize (zero energy <=> no pair closer
 * than tol; check with mdb_count_pairs(h, tol)), which is what Packmol's pack_monoatomic! does on the host (nranks == 1). */
int mdb_random_positions(mdb_handle h, uint64_t stream);
/* Exact binary checkpoint: positions, velocities, forces, images and ids in device slot order plus the RNG step counter.
 * Saving invalidates the resident Verlet list, so the saved run and a run restored with mdb_checkpoint_load on a handle
 * created with the same mdb_config continue bit-identically.
 * nranks > 1: COLLECTIVE by convention (every rank calls it at the same point of the run: the invalidated list must be
 * invalid on all ranks at once); each rank writes / reads "<path>.<rank>" holding the particles it owns plus what a fresh
 * handle needs to plan its slab, so a ring of new handles (mdb_create + mdb_comm_init*) restores without the global arrays. */
int mdb_checkpoint_save(mdb_handle h, const char *path);
int mdb_checkpoint_load(mdb_handle h, const char *path);

/* ---- introspection ------------------------------------------------------------------------------------ */
int mdb_get_stats(mdb_handle h, mdb_stats *out);
/* device pointer to resident arrays, for zero-copy interop (CUDA.jl CuArray / torch tensor views).
 * which: 0 pos4 {x,y,z,sigma} [n], 1 velocities SoA, 2 forces SoA, 3 images SoA, 4 ids; *stride = SoA component stride */
int mdb_device_ptr(mdb_handle h, int32_t which, void **ptr, int64_t *stride);
/* the CUDA stream the engine launches on (cudaStream_t), so callers can order their own work / events after it */
int mdb_stream(mdb_handle h, void **stream);
int mdb_synchronize(mdb_handle h);
/* measured FP64 (DFMA) throughput of the device in TFLOP/s (8 independent FMA chains per thread, best of 5): the
 * denominator bench.py uses next to ncu's FP64-pipe utilisation; not part of the path */
int mdb_measure_fp64_peak(mdb_handle h, double *tflops);
/* identity of the dominant kernel as built into THIS library (the fused NVE pair-force kernel of the handle's dimension and
 * potential): info[0] = registers per thread, [1] = static shared memory (bytes), [2] = resident CTAs per SM the engine
 * launches, [3] = kernel variant (0 direct gathers, 1 cp.async-staged gathers, 2 TMA operand ring), [4] = threads per CTA, [5] = local (stack) bytes per
 * thread.  bench.py refuses ncu-derived numbers (profiles/rNN_traffic.json) captured on a different build. */
int mdb_force_kernel_info(mdb_handle h, int32_t info[6]);

#ifdef __cplusplus
}
#endif
#endif /* MDB200_H */

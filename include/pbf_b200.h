/*
 * pbf_b200.h — C ABI of the B200-native Position-Based-Fluids substep backend.
 *
 * This is the drop-in boundary for the one hot path of CTKnight/FluidSimulator:
 * the PBF substep.  The reference exposes that path as three C++ free functions
 * (reference cuda/include/fluid/cuda.h:7-9):
 *
 *     int  fluid::cuda_version();
 *     bool fluid::cuda_device_available(int* count, const char** error);
 *     void fluid::cuda_step(const Params& params, State& state);
 *
 * called from exactly one site (reference app/src/main.cpp:171-187 probe,
 * :250-254 step).  Everything below is plain C: pointers, sizes and PODs.
 * The C++ shim that re-creates the three reference symbols on top of this ABI
 * lives in fluidsimulator_b200/csrc/host/cuda_shim.cpp; INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returning int returns PBF_OK (0) or a negative PBF_E_* code;
 *     the human-readable reason is available from pbf_last_error().
 *   - there is NO CPU fallback: without a usable CUDA device pbf_create() fails.
 *   - particle arrays are SoA float32 in ORIGINAL particle order, exactly like
 *     fluid::State (reference core/include/fluid/core.h:121-132).
 */
#ifndef PBF_B200_H
#define PBF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBF_ABI_VERSION 1

/* status codes */
#define PBF_OK 0
#define PBF_E_INVALID (-1)   /* bad argument / bad state                       */
#define PBF_E_CUDA (-2)      /* a CUDA runtime call failed                     */
#define PBF_E_NODEVICE (-3)  /* no CUDA device                                 */
#define PBF_E_CAPACITY (-4)  /* a device table could not be grown              */
#define PBF_E_COMM (-5)      /* NCCL / slab exchange failure                   */

/* arithmetic modes (pbf_set_mode) */
#define PBF_MODE_STRICT 0 /* IEEE binary32, oracle operation order, no FMA contraction:
                             bit-identical to the reference CPU path (core.cpp:119-615) */
#define PBF_MODE_FAST 1   /* FMA contraction + approximate sqrt; tolerance-gated        */

/* POD mirror of fluid::Params (reference core/include/fluid/core.h:11-79),
 * minus `backend` (meaningless here) and the plane arrays (pbf_set_planes). */
typedef struct pbf_params {
  float dt;                      /* core.h:12 */
  float density;                 /* core.h:18 */
  float particle_mass;           /* core.h:20 */
  float h;                       /* core.h:22 */
  float particle_radius;         /* core.h:23 (unused by the substep) */
  float epsilon;                 /* core.h:24 */
  int32_t solver_iterations;     /* core.h:25 */
  float neighbor_reserve_factor; /* core.h:26 (host-vector hint; ignored) */
  int32_t use_uniform_grid;      /* core.h:27 (must be non-zero; the O(N^2) path is out of scope) */
  int32_t enable_scorr;          /* core.h:28 */
  int32_t enable_xsph;           /* core.h:29 */
  int32_t enable_vorticity;      /* core.h:30 */
  float scorr_k;                 /* core.h:31 */
  int32_t scorr_n;               /* core.h:32 */
  float scorr_dq_coeff;          /* core.h:33 */
  float visc_c;                  /* core.h:34 */
  float plane_restitution;       /* core.h:35 */
  float plane_friction;          /* core.h:36 */
  float vort_epsilon;            /* core.h:37 */
  float vort_norm_eps;           /* core.h:38 */
  float external_force[3];       /* core.h:39-43 */
} pbf_params;

/* Fills *p with the defaults of fluid::Params (core.h:12-43). */
void pbf_default_params(pbf_params* p);

typedef struct pbf_ctx pbf_ctx;

int pbf_abi_version(void);

/* Replaces fluid::cuda_device_available (reference cuda_stub.cu:740-762).
 * Returns the device count (>= 0) or a negative code; *err (if non-NULL) gets a
 * borrowed static string or NULL. */
int pbf_device_count(const char** err);

/* One context = one GPU = one slab of particles.  `capacity` is a particle-count
 * hint (tables grow on demand).  Returns NULL on failure (see pbf_last_error(NULL)). */
pbf_ctx* pbf_create(int device, size_t capacity);
void pbf_destroy(pbf_ctx* ctx);

/* Borrowed string, valid until the next failing call on the same ctx.
 * ctx == NULL reports the last context-less failure (pbf_create, pbf_device_count). */
const char* pbf_last_error(const pbf_ctx* ctx);

int pbf_set_params(pbf_ctx* ctx, const pbf_params* params);
/* Plane SoA, reference core.h:44-78; normals are used as given (already normalised). */
int pbf_set_planes(pbf_ctx* ctx, int count, const float* nx, const float* ny,
                   const float* nz, const float* d);
int pbf_set_mode(pbf_ctx* ctx, int mode);
/* Is the current parameter set inside the domain where PBF_MODE_STRICT is bit-identical to the
 * reference CPU path?  1 = yes; 0 = no, *why (may be NULL) names the reason: an s_corr exponent
 * outside {2,3,4} (reference: std::pow, core.cpp:59-71) or solver_iterations == 0 with XSPH /
 * vorticity enabled (reference: those passes run on stale scratch, core.cpp:423-571).  Every
 * shipped scene and every CLI flag combination of the reference is inside the domain. */
int pbf_strict_exact(const pbf_ctx* ctx, const char** why);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream) so that the
 * caller's CUDA events bracket the work.  NULL = the context's own stream. */
int pbf_set_stream(pbf_ctx* ctx, void* cuda_stream);
/* 1 = replay each substep as a CUDA graph (default), 0 = plain stream launches. */
int pbf_set_graph(pbf_ctx* ctx, int enabled);
/* Kernel family of the neighbour build and the passes that walk the neighbour list (what replaces
 * reference cuda_stub.cu:155-416 either way; results are bit-identical):
 *   PBF_BRICK_OFF         one thread per particle, neighbours gathered from global memory through L1,
 *                         32-bit list entries.  Default: measured faster on B200.
 *   PBF_BRICK_PERSISTENT  bricks of 4x4x4 grid cells; the brick plus one halo cell layer is staged into
 *                         shared memory by TMA bulk copies, gathers are LDS.128 through 16-bit
 *                         tile-relative list entries; persistent CTAs with a ring of tile slots.
 *   PBF_BRICK_PER_CTA     the same with one CTA per brick.
 * The environment variable PBF_BRICK=0|1|2 sets the initial choice of a context.  A batch the brick
 * path cannot hold (tile capacity, sparse cell table of a diverged scene) is transparently replayed on
 * the global-gather family. */
#define PBF_BRICK_OFF 0
#define PBF_BRICK_PERSISTENT 1
#define PBF_BRICK_PER_CTA 2
int pbf_set_brick(pbf_ctx* ctx, int mode);
/* Returns 1 if the last pbf_step batch ran on the brick path, 0 if not; optional outputs: batches
 * replayed on the global-gather family so far, largest tile (records) of the last batch. */
int pbf_brick_status(const pbf_ctx* ctx, uint64_t* fallbacks, uint32_t* max_tile);

/* Batches that were replayed so far because a device table (cell table, neighbour table, slab
 * message or particle capacity) had to grow or a peer timed out.  Replays are transparent — results
 * never depend on them — but they cost time: a benchmark reports this count for its timed region. */
uint64_t pbf_batches_retried(const pbf_ctx* ctx);

/* Asynchronous frame output (device-resident stepping; replaces the blocking read of positions for
 * the VTK writer at reference app/src/main.cpp:259-273).  pbf_snapshot_begin enqueues, after the
 * substeps issued so far, a copy of the positions into one of two library-owned pinned host
 * buffers and returns at once; the copy runs on its own stream under the next pbf_step batch.
 * pbf_snapshot_wait blocks until that copy has landed and returns the three arrays (original
 * particle order; valid until the next pbf_snapshot_begin on the same slot), the particle count and
 * the simulation time of the snapshot.  Not available on slab contexts. */
int pbf_snapshot_begin(pbf_ctx* ctx, int slot);
int pbf_snapshot_wait(pbf_ctx* ctx, int slot, const float** px, const float** py, const float** pz,
                      size_t* n, float* time);

/* Host SoA -> device (the six H2D copies of reference cuda_stub.cu:791-796).
 * Resets nothing else; time is kept (use pbf_set_time). */
int pbf_upload(pbf_ctx* ctx, size_t n, const float* px, const float* py,
               const float* pz, const float* vx, const float* vy, const float* vz);
/* Device -> host SoA in original particle order (reference cuda_stub.cu:1092-1097).
 * Any pointer may be NULL to skip that array. */
int pbf_download(pbf_ctx* ctx, float* px, float* py, float* pz, float* vx,
                 float* vy, float* vz);

/* Runs `nsteps` substeps (reference core.cpp:119-615 each) device-resident:
 * no host<->device particle traffic, no host synchronisation between substeps.
 * Synchronises the stream before returning and validates the device tables
 * (a substep that overflowed a table is re-run after growing it; the result is
 * independent of table sizes). */
int pbf_step(pbf_ctx* ctx, int nsteps);

/* Page-locks (or releases) a caller-owned host array so that pbf_upload / pbf_download /
 * pbf_step_host move it at full PCIe speed instead of through the driver's staging buffers.
 * For callers whose arrays persist across steps, like fluid::State's std::vectors (the shim
 * registers them once and re-registers when a vector is re-allocated).  Not required. */
int pbf_host_register(pbf_ctx* ctx, void* ptr, size_t bytes);
int pbf_host_unregister(pbf_ctx* ctx, void* ptr);

/* The reference's cuda_step contract in one call (cuda_stub.cu:764-1099): host
 * arrays in, `nsteps` substeps, host arrays out, all inside the call. */
int pbf_step_host(pbf_ctx* ctx, size_t n, float* px, float* py, float* pz,
                  float* vx, float* vy, float* vz, int nsteps);

size_t pbf_count(const pbf_ctx* ctx);
float pbf_time(const pbf_ctx* ctx);       /* State::time, float-accumulated (core.cpp:614) */
int pbf_set_time(pbf_ctx* ctx, float t);

/* ---- parity / debug surface (mirrors what fluid::step leaves in State::cpu,
 * reference core.h:92-115).  All refer to the LAST substep executed. ---------- */

/* Occupied-cell count and total neighbour-list length of the last substep. */
int pbf_debug_sizes(pbf_ctx* ctx, size_t* ncells, size_t* nneighbors);
/* grid_entries (core.h:112): per sorted slot the cell coordinates and particle id;
 * grid_keys/grid_starts/grid_ends (core.h:113-115): per occupied cell.
 * Arrays: entry_* [n]; cell_xyz [3*ncells] (x,y,z interleaved); cell_start/end [ncells].
 * Always in the reference's order (lexicographic x, y, z, then particle id), also when the
 * substep used the sparse cell table.  On a slab context: the slab's own particles, local ids. */
int pbf_debug_grid(pbf_ctx* ctx, int32_t* entry_cx, int32_t* entry_cy,
                   int32_t* entry_cz, int32_t* entry_particle, int32_t* cell_xyz,
                   int32_t* cell_start, int32_t* cell_end);
/* neighbor_prefix_sum (inclusive, per ORIGINAL particle id) and neighbor_indices
 * (original ids, oracle traversal order) — core.h:110-111, core.cpp:205-247. */
int pbf_debug_neighbors(pbf_ctx* ctx, int32_t* prefix_sum, int32_t* indices);

/* Per-particle float scratch in original order (core.h:92-109). */
enum pbf_scratch_id {
  PBF_SCRATCH_PRED_X = 0, PBF_SCRATCH_PRED_Y, PBF_SCRATCH_PRED_Z,
  PBF_SCRATCH_DELTA_X, PBF_SCRATCH_DELTA_Y, PBF_SCRATCH_DELTA_Z,
  PBF_SCRATCH_LAMBDA, PBF_SCRATCH_RHO,
  PBF_SCRATCH_DV_X, PBF_SCRATCH_DV_Y, PBF_SCRATCH_DV_Z,
  PBF_SCRATCH_OMEGA_X, PBF_SCRATCH_OMEGA_Y, PBF_SCRATCH_OMEGA_Z, PBF_SCRATCH_OMEGA_MAG,
  PBF_SCRATCH_ETA_X, PBF_SCRATCH_ETA_Y, PBF_SCRATCH_ETA_Z,
  PBF_SCRATCH_COUNT
};
/* Enables retention of the scratch arrays (extra stores in the kernels). */
int pbf_debug_enable(pbf_ctx* ctx, int enabled);
int pbf_debug_scratch(pbf_ctx* ctx, int scratch_id, float* out);

/* ---- measurement surface ------------------------------------------------- */

enum pbf_stage_id {
  PBF_STAGE_PREDICT = 0, /* a3+a4: integrate, cell coordinates, bounds          */
  PBF_STAGE_SORT,        /* a5+a6: counting sort by cell key, cell start/end     */
  PBF_STAGE_CELLS,       /* order cells by particle id, reorder into sorted order */
  PBF_STAGE_NEIGHBORS,   /* a7: neighbour list                                   */
  PBF_STAGE_LAMBDA,      /* a8                                                   */
  PBF_STAGE_DELTA,       /* a9+a10 (+a11 on the last iteration)                  */
  PBF_STAGE_XSPH,        /* a12                                                  */
  PBF_STAGE_VORT_OMEGA,  /* a13 first pass                                       */
  PBF_STAGE_VORT_APPLY,  /* a13 second pass + apply                              */
  PBF_STAGE_FINALIZE,    /* a14 + scatter back to original order                 */
  PBF_STAGE_EXCHANGE,    /* slab halo / migration traffic (multi-GPU only)       */
  PBF_STAGE_COUNT
};
const char* pbf_stage_name(int stage);
/* With profiling on, substeps are launched un-graphed with CUDA events around
 * every stage; totals accumulate until pbf_profile_reset. */
int pbf_profile_enable(pbf_ctx* ctx, int enabled);
int pbf_profile_reset(pbf_ctx* ctx);
int pbf_profile_get(pbf_ctx* ctx, int stage, double* total_ms, uint64_t* launches);
/* Number of this library's kernels launched since creation (graph replays counted
 * by the kernels they contain). */
uint64_t pbf_launch_count(const pbf_ctx* ctx);

/* ---- slab decomposition (one ctx per GPU; x-slabs) ----
 * The reference has no multi-GPU path.  Large scenes are split into slabs of x-cells; every
 * substep exchanges migrating particles and two ghost layers with the x-neighbours and refreshes
 * the ghosts once per solver iteration.  STRICT results are bit-identical to a single GPU. */

/* Host-only planning: cuts[0..nranks] on x-cell boundaries (cuts[0] = INT32_MIN, cuts[nranks] =
 * INT32_MAX) with >= 2 cell layers per slab; slab r owns the x-cells [cuts[r], cuts[r+1]).  Cell
 * of a position: floor(x * (1.0f / h)) (core.cpp:28-34).  The cuts minimise the modelled work of
 * the busiest slab: owned particles + ghost_weight * first ghost layers + ghost_weight / 4 *
 * second ghost layers (slabs between two neighbours carry two ghost sides and therefore own
 * fewer particles).  pbf_slab_plan, pbf_slab_upload and the automatic re-balancing use the
 * process default (0.5, environment PBF_SLAB_GHOST_WEIGHT in [0, 1]; 0 = equal owned counts). */
int pbf_slab_plan(size_t n, const float* px, float h, int nranks, int32_t* cuts);
/* The same planner on a histogram: hist[l] = particles in the x-layer first_layer + l. */
int pbf_slab_plan_hist(const uint64_t* hist, int32_t nlayers, int32_t first_layer, int nranks,
                       float ghost_weight, int32_t* cuts);

/* Multi-process transport: one NCCL communicator over the slabs.  It carries the halo messages
 * themselves (ncclSend/ncclRecv between x-neighbours) or, by default, only the cudaIpc handles of
 * the peer windows and the per-batch status reduction (see pbf_slab_set_p2p). */
#define PBF_COMM_ID_BYTES 128
/* Rank 0 calls this and ships the bytes to every rank by any host channel
 * (torch.distributed broadcast, a file, MPI...). */
int pbf_comm_unique_id(void* id_bytes);
/* Joins the NCCL communicator of `nranks` slabs and turns ctx into slab `rank`.  Collective. */
int pbf_comm_init(pbf_ctx* ctx, int rank, int nranks, const void* id_bytes);

/* How halo data moves on the substep path.  1 = direct peer stores: the pack kernels write into
 * the neighbour's memory (cudaIpc window between processes, plain pointers inside one process)
 * and an exchange is one flag kernel — no message copy, no collective call; default after
 * pbf_comm_init (environment PBF_SLAB_P2P=0 disables it).  0 = messages: ncclSend/ncclRecv
 * between processes, peer copies inside a process (default for pbf_group_create).  Collective:
 * every rank must choose the same mode before the next pbf_step. */
int pbf_slab_set_p2p(pbf_ctx* ctx, int enabled);

/* Distributes a global particle set: every rank passes the SAME global arrays (after
 * pbf_set_params) and keeps the particles of its slab; global ids = indices into these arrays. */
int pbf_slab_upload(pbf_ctx* ctx, size_t n_global, const float* px, const float* py,
                    const float* pz, const float* vx, const float* vy, const float* vz);
/* This rank's own particles only (unique global ids, in any order): the per-rank analogue of
 * pbf_upload.  The cuts of the last pbf_slab_upload stay; particles outside them migrate during
 * the next substep.  global_id == NULL: the same particles, in the same order, as the last
 * pbf_slab_download returned (n must match) — only positions and velocities are replaced. */
int pbf_slab_upload_owned(pbf_ctx* ctx, size_t n, const int64_t* global_id, const float* px,
                          const float* py, const float* pz, const float* vx, const float* vy,
                          const float* vz);
/* Particles currently owned by this slab (changes as particles migrate), its cuts, and the
 * owned particles with their global ids (in the slab's storage order, which is arbitrary). */
size_t pbf_slab_owned(const pbf_ctx* ctx);
int pbf_slab_cuts(const pbf_ctx* ctx, int32_t* lo, int32_t* hi);
/* Re-balancing: new cuts for this slab (e.g. from pbf_slab_plan on the gathered positions).  By
 * convention collective — the cuts of all ranks must tile the x axis.  Particles that now lie in
 * another slab migrate during the next substep (one hop per slab of distance; the hop count
 * grows on demand), so results do not depend on when or how the cuts move. */
int pbf_slab_set_cuts(pbf_ctx* ctx, int32_t lo, int32_t hi);
/* Automatic re-balancing: at the end of a pbf_step batch, when the largest slab owns more than
 * `threshold` times the mean (default 1.1; 0 = never), all ranks re-plan equal-count cuts on the
 * x-layer histogram of the whole scene.  Collective setting: the same value on every rank. */
int pbf_slab_set_rebalance(pbf_ctx* ctx, float threshold);
uint64_t pbf_slab_rebalance_count(const pbf_ctx* ctx);
int pbf_slab_download(pbf_ctx* ctx, int64_t* global_id, float* px, float* py,
                      float* pz, float* vx, float* vy, float* vz);
/* Payload bytes this slab sent during the last substep of the last batch (pbf_slab_stats counts
 * message capacities, which is what NCCL messages move; peer stores move the payload only). */
int pbf_slab_payload(const pbf_ctx* ctx, uint64_t* bytes_last_substep);
/* Which data plane the last pbf_step batch of this slab used for its halo exchanges. */
#define PBF_TRANSPORT_LOCAL_COPIES 1  /* one process: cudaMemcpyPeerAsync + events */
#define PBF_TRANSPORT_NCCL_MESSAGES 2 /* ncclSend / ncclRecv between x-neighbours */
#define PBF_TRANSPORT_PEER_STORES 3   /* pack kernels store into the neighbour's window, flag kernels */
int pbf_slab_transport(const pbf_ctx* ctx);
/* Exchanges and bytes sent since creation, ghosts held after the last substep, migration hops. */
int pbf_slab_stats(const pbf_ctx* ctx, uint64_t* exchanges, uint64_t* bytes_sent,
                   int32_t* ghosts, int32_t* hops);
/* pbf_step on a slab context is collective: every rank calls it with the same nsteps. */

/* Single-process transport: `n` contexts (on the same or on different devices) become slabs
 * 0..n-1 linked by peer copies; pbf_group_step runs one host thread per slab.  Destroy the
 * group before its contexts. */
typedef struct pbf_group pbf_group;
pbf_group* pbf_group_create(pbf_ctx** ctxs, int n);
void pbf_group_destroy(pbf_group* group);
int pbf_group_upload(pbf_group* group, size_t n_global, const float* px, const float* py,
                     const float* pz, const float* vx, const float* vy, const float* vz);
int pbf_group_step(pbf_group* group, int nsteps);
/* All slabs gathered into global arrays in original particle order (any pointer may be NULL). */
int pbf_group_download(pbf_group* group, float* px, float* py, float* pz, float* vx,
                       float* vy, float* vz);
size_t pbf_group_count(const pbf_group* group);

#ifdef __cplusplus
}
#endif
#endif /* PBF_B200_H */

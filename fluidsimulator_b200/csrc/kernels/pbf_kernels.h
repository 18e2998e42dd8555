// pbf_kernels.h — host-callable launchers of the sm_100a kernels (one per stage of
// the PBF substep).  Every launcher enqueues on `stream`,
// returns the number of kernels it launched, and never synchronises.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pbf_device.cuh"

namespace pbf {

// <<<grid, block, 0, stream>>> with programmatic stream serialization (see pdl_wait in pbf_device.cuh)
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PBF_PDL ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define PBF_LAUNCH(kernel, grid, block, stream, ...) \
  ::pbf::launch_pdl(kernel, dim3(grid), dim3(block), stream, __VA_ARGS__)

struct GridBuffers {
  GridDesc* desc;
  StatusBlock* status;
  uint32_t* keys[2];
  uint32_t* vals[2];
  uint32_t* chunk_total;  // one per 2048 table cells
  int2* cell_range;     // cell_cap entries
  uint32_t* cell_count; // cell_cap entries, zero between substeps (counting sort)
  uint32_t* cell_excl;  // cell_cap entries: exclusive prefix inside a scan chunk
  uint32_t* slot_id;    // n entries: particle id per slot before the cells are ordered by id
  unsigned long long* cell_key;  // cell_cap entries: packed coordinates per hash slot (sparse table), else empty
  uint32_t cell_cap;
  BrickRec* bricks;     // brick path: brick_cap records, rewritten every substep (nullptr = brick path off)
  int brick_cap;
};

struct NeighborList {
  uint32_t* idx;     // [(ceil(n/32)) * K * 32], pair-interleaved (see pbf_device.cuh); K even
  uint32_t* count;   // [n]
  int K;
  // Brick path (brick.cu): `idx` then holds 16-bit tile-relative entries, entry k of slot i at
  // ((uint16_t*)idx)[(i/32)*K*32 + (k/4)*128 + (i%32)*4 + k%4], and every pass that walks the list
  // runs as one CTA per brick.  nullptr = global-gather family (32-bit entries).
  const BrickRec* bricks = nullptr;
  const GridDesc* desc = nullptr;
  int brick_cap = 0;
  unsigned int* brick_ctl = nullptr;  // persistent brick kernels: {brick ticket, finished CTAs}, zero between launches
  bool brick_persist = true;          // persistent CTAs with a ring of tile slots (false: one CTA per brick)
};

// SoA host staging <-> float4 persistent state
int launch_pack_state(const float* const soa[6], float4* pos_o, float4* vel_o, int n, cudaStream_t s);
int launch_unpack_state(const float4* pos_o, const float4* vel_o, float* const soa[6], int n, cudaStream_t s);

// a3+a4: integrate, predicted positions, cell bounds (always STRICT arithmetic: the
// grid tables are bit-exact in both modes).  slab: bounds are taken after migration instead.
int launch_predict(float4* pos_o, float4* vel_o, float4* pred_o, const StepConsts& c,
                   const GridBuffers& g, NRef n, bool slab, cudaStream_t s);
int launch_grid_finalize(const GridBuffers& g, int pad, NRef n, cudaStream_t s);
// a4+a5+a6: dense keys, counting sort by cell, cell start/end table.  launch_cells_reorder then
// orders every cell by particle id (`gid` = global ids in slab mode, nullptr = the slot index),
// writes the sorted keys / ids to g.keys[*out] / g.vals[*out] and gathers pred/pos into sorted order.
int launch_sort(const float4* pred_o, const StepConsts& c, const GridBuffers& g, NRef n,
                int* out, cudaStream_t s);
int launch_cells_reorder(const float4* pred_o, const float4* pos_o, float4* pred_s, float4* pos_s,
                         const uint32_t* gid, const GridBuffers& g, NRef n, cudaStream_t s);
// a7: neighbour list in the oracle's traversal order.
int launch_neighbors(const float4* pred_s, const StepConsts& c, const GridBuffers& g,
                     const NeighborList& nl, NRef n, cudaStream_t s);

struct SolveBuffers {
  float4* pred[2];   // ping-pong (pred xyz, lambda)
  float4* pos_s;     // (pos xyz, bits(orig id))
  float4* vel[2];    // (vel xyz, m/rho)
  PosVel* pv;        // XSPH input record, written by the last delta pass instead of vel[0] when XSPH runs
  float4* omega;     // (omega xyz, |omega|)
  float* rho;
  const float4* planes;  // (nx, ny, nz, d) x nplanes, device
  float4* pos_o;     // persistent, original order (written by finalize)
  float4* vel_o;
  StatusBlock* status;
  DebugPtrs dbg;
  HaloOut halo;      // slab mode: where the next delta / xsph launch also stores its boundary layers
};

// a8..a14 for one substep: I x (lambda, delta), velocity update, XSPH, vorticity,
// restitution, scatter to original order.  `strict` picks the arithmetic policy.
// stage_cb (may be null) is invoked between stages for profiling.
PosVel* xsph_record(const SolveBuffers& b, const StepConsts& c);  // b.pv when XSPH runs, else nullptr
typedef void (*StageCallback)(void* user, int stage_id, int begin);
int launch_solve(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c,
                 int iterations, NRef n, bool strict, cudaStream_t s,
                 StageCallback cb, void* cb_user, int phase = 0);
// positions (sorted order, original id in pos_s.w) -> three SoA arrays in original order
int launch_scatter_positions(const float4* pos_sorted, const float4* pos_s, float* x, float* y, float* z, int n,
                             cudaStream_t s);
// ---- brick path (kernels/brick.cu): the same stages as one CTA per brick of grid cells, halo staged
// into shared memory with cp.async.bulk, 16-bit tile-relative neighbour entries ------------------
int brick_setup();  // opt the kernels in to their dynamic shared memory (once per process / device)
int launch_brick_table(const GridBuffers& g, cudaStream_t s);
int launch_neighbors_brick(const float4* pred_s, const StepConsts& c, const GridBuffers& g, const NeighborList& nl,
                           cudaStream_t s);
int launch_lambda_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool strict,
                        cudaStream_t s);
int launch_delta_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool last,
                       bool is_final, bool strict, cudaStream_t s);
int launch_xsph_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, bool is_final,
                      bool strict, cudaStream_t s);
int launch_vort_omega_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                            bool strict, cudaStream_t s);
int launch_vort_apply_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                            bool strict, cudaStream_t s);

// The individual passes (the slab driver puts halo exchanges between them).  `cur` selects the
// pred ping-pong buffer a pass reads; delta writes pred[cur ^ 1].
int launch_lambda(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, NRef n,
                  bool strict, cudaStream_t s, Span span = Span{nullptr, 0});
int launch_delta(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool last,
                 bool is_final, NRef n, bool strict, cudaStream_t s);
int launch_xsph(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, bool is_final,
                NRef n, bool strict, cudaStream_t s);
int launch_vort_omega(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                      NRef n, bool strict, cudaStream_t s);
int launch_vort_apply(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                      NRef n, bool strict, cudaStream_t s);
int launch_commit_only(const SolveBuffers& b, const StepConsts& c, NRef n, bool strict, cudaStream_t s);

// ---- x-slab decomposition (kernels/slab.cu) --------------------------------------
// Messages between x-neighbours are fixed-capacity float4 arrays; element 0 is a header whose .x
// holds the element count as bits.
struct SlabBuffers {
  SlabCounts* counts;       // device
  StatusBlock* status;
  uint32_t* gid_o;          // global particle id per owned slot (any order)
  uint32_t* holes;          // 6 * mcap: vacated slots, then two scratch lists of 2 * mcap each
  float4* send[2];          // message to the left / right neighbour
  float4* recv[2];          // message from the left / right neighbour
  int cut_lo, cut_hi;       // owned x-cells [cut_lo, cut_hi); INT_MIN / INT_MAX at the ends
  int cap;                  // owned-particle capacity
  int tot_cap;              // capacity of the sorted arrays (owned + ghosts)
  int mcap;                 // migration message capacity (particles)
  int gcap;                 // ghost message capacity (particles)
  PeerSync sync;            // exchange fused into the next consumer kernel (mine == nullptr: not fused)
};
// migration, one hop: particles whose predicted x-cell left the slab go into the two messages and
// the slots they vacate are refilled from the tail of the owned range (O(migrants) data movement)
int launch_slab_split(float4* pos_o, float4* pred_o, const SlabBuffers& sb, const StepConsts& c,
                      cudaStream_t s);
// append the received particles; on the last hop also extends the cell bounds by them and flags
// arrivals that are still outside the slab
int launch_slab_merge(float4* pos_o, float4* pred_o, const SlabBuffers& sb, const StepConsts& c, bool last_hop,
                      cudaStream_t s);
// boundary counts + ghost messages (pred, pos of the two boundary layers on each side)
int launch_slab_ghost_pack(const uint32_t* keys_sorted, const float4* pred_s, const float4* pos_s,
                           const GridBuffers& g, const SlabBuffers& sb, cudaStream_t s);
// append received ghosts after the owned slots and insert their cells into the table
int launch_slab_ghost_unpack(float4* pred_s, float4* pos_s, const GridBuffers& g, const SlabBuffers& sb,
                             const StepConsts& c, cudaStream_t s);
// per-iteration refresh of one float4 array: messages -> ghost slots (the boundary slots are stored
// into the messages by the producing delta / xsph launch, SolveBuffers::halo)
int launch_slab_halo_unpack(float4* arr, const SlabBuffers& sb, cudaStream_t s);
// re-balancing support: this slab's owned count into the status block; x-cell range and per-layer
// histogram of the owned particles (positions after the batch)
int launch_slab_report(const SlabBuffers& sb, int rank, cudaStream_t s);
// (x-cells of pos + vel * lookahead: where the particles will be when the new cuts are in use)
int launch_slab_xrange(const float4* pos_o, const float4* vel_o, const SlabBuffers& sb, const StepConsts& c, float lookahead,
                       int* out_min_max, cudaStream_t s);
int launch_slab_xhist(const float4* pos_o, const float4* vel_o, const SlabBuffers& sb, const StepConsts& c, float lookahead,
                      int x_min, int layers, unsigned long long* hist, cudaStream_t s);
// ghost velocities (pred - pos)/dt and m/rho, recomputed locally after the last pred refresh
int launch_slab_ghost_vel(const float4* pred_final, const float4* pos_s, const float* rho, float4* vel, PosVel* pv,
                          const SlabBuffers& sb, const StepConsts& c, bool strict, cudaStream_t s);

}  // namespace pbf

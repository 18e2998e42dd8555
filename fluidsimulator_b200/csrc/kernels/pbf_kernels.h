// pbf_kernels.h — host-callable launchers of the sm_100a kernels (one per stage of
// the PBF substep).  Every launcher enqueues on `stream`,
// returns the number of kernels it launched, and never synchronises.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pbf_device.cuh"

namespace pbf {

constexpr int kRadixBits = 8;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kSortTile = 2048;  // keys per block in the radix passes

inline int sort_blocks(int n) { return (n + kSortTile - 1) / kSortTile; }

struct GridBuffers {
  GridDesc* desc;
  StatusBlock* status;
  uint32_t* keys[2];
  uint32_t* vals[2];
  uint32_t* hist;       // kRadixBins * sort_blocks(n)
  uint32_t* chunk_total;  // one per 2048 hist entries
  int2* cell_range;     // cell_cap entries
  uint32_t cell_cap;
  int sort_passes;      // ceil(log2(cell_cap) / 8)
};

struct NeighborList {
  uint32_t* idx;     // [(ceil(n/32)) * K * 32], pair-interleaved (see pbf_device.cuh); K even
  uint32_t* count;   // [n]
  int K;
};

// SoA host staging <-> float4 persistent state
int launch_pack_state(const float* const soa[6], float4* pos_o, float4* vel_o, int n, cudaStream_t s);
int launch_unpack_state(const float4* pos_o, const float4* vel_o, float* const soa[6], int n, cudaStream_t s);

// a3+a4: integrate, predicted positions, cell bounds (always STRICT arithmetic: the
// grid tables are bit-exact in both modes).
int launch_predict(float4* pos_o, float4* vel_o, float4* pred_o, const StepConsts& c,
                   const GridBuffers& g, int n, cudaStream_t s);
// a4+a5: dense keys + LSD radix sort (stable => ties keep ascending particle id).
// On return the sorted keys/vals are in g.keys[out]/g.vals[out]; returns launches, sets *out.
int launch_sort(const float4* pred_o, const StepConsts& c, const GridBuffers& g, int n,
                int* out, cudaStream_t s);
// a6 + reorder: dense cell start/end table, gather pred/pos into sorted order.
int launch_cells_reorder(const uint32_t* keys, const uint32_t* vals, const float4* pred_o,
                         const float4* pos_o, float4* pred_s, float4* pos_s,
                         const GridBuffers& g, int n, cudaStream_t s);
// a7: neighbour list in the oracle's traversal order.
int launch_neighbors(const float4* pred_s, const StepConsts& c, const GridBuffers& g,
                     const NeighborList& nl, int n, cudaStream_t s);

struct SolveBuffers {
  float4* pred[2];   // ping-pong (pred xyz, lambda)
  float4* pos_s;     // (pos xyz, bits(orig id))
  float4* vel[2];    // (vel xyz, m/rho)
  float4* omega;     // (omega xyz, |omega|)
  float* rho;
  const float4* planes;  // (nx, ny, nz, d) x nplanes, device
  float4* pos_o;     // persistent, original order (written by finalize)
  float4* vel_o;
  StatusBlock* status;
  DebugPtrs dbg;
};

// a8..a14 for one substep: I x (lambda, delta), velocity update, XSPH, vorticity,
// restitution, scatter to original order.  `strict` picks the arithmetic policy.
// stage_cb (may be null) is invoked between stages for profiling.
typedef void (*StageCallback)(void* user, int stage_id, int begin);
int launch_solve(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c,
                 int iterations, int n, bool strict, cudaStream_t s,
                 StageCallback cb, void* cb_user);

}  // namespace pbf

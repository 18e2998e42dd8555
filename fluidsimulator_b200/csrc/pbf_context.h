// pbf_context.h — host-side driver state behind one pbf_ctx (one GPU, one slab).
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "kernels/pbf_kernels.h"
#include "pbf_b200.h"

namespace pbf {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;  // elements
  cudaError_t reserve(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct StageTimer {
  std::vector<cudaEvent_t> begin, end;  // recorded pairs awaiting resolution
  std::vector<int> stage;
  double total_ms[PBF_STAGE_COUNT] = {0};
  uint64_t launches[PBF_STAGE_COUNT] = {0};
  std::vector<cudaEvent_t> pool;
};

}  // namespace pbf

struct pbf_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string error;

  pbf_params params{};
  std::vector<float4> planes_host;
  pbf::StepConsts consts{};
  int mode = PBF_MODE_STRICT;
  bool use_graph = true;
  bool debug = false;
  bool profile = false;

  size_t n = 0;         // particles
  size_t cap = 0;       // particle capacity of the per-particle buffers
  float time = 0.0f;

  // persistent state (original order) + batch backup
  pbf::DevBuf<float4> pos_o, vel_o, pos_bak, vel_bak, pred_o;
  // sorted-order work arrays
  pbf::DevBuf<float4> pred_a, pred_b, pos_s, vel_a, vel_b, omega;
  pbf::DevBuf<float> rho;
  pbf::DevBuf<float4> planes_dev;
  // grid
  pbf::DevBuf<pbf::GridDesc> desc;
  pbf::DevBuf<pbf::StatusBlock> status;
  pbf::DevBuf<uint32_t> keys0, keys1, vals0, vals1, hist, chunk_total;
  pbf::DevBuf<int2> cell_range;
  uint32_t cell_cap = 1u << 22;
  int sorted_buf = 0;  // which keys/vals buffer holds the last substep's sorted order
  // neighbour list
  pbf::DevBuf<uint32_t> nbr_idx, nbr_count;
  int K = 96;
  // SoA staging for upload / download
  pbf::DevBuf<float> soa[6];
  // debug scratch (sorted order)
  pbf::DevBuf<float> dbg_lambda, dbg_rho;
  pbf::DevBuf<float4> dbg_delta, dbg_dv, dbg_eta;

  pbf::StatusBlock* status_host = nullptr;  // pinned
  pbf::StatusBlock last_status{};           // of the last completed batch
  pbf::GridDesc last_desc{};

  // CUDA graph of one substep (re-captured when the configuration changes)
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_key = 0;
  int graph_kernels = 0;

  pbf::StageTimer timer;
  uint64_t launch_count = 0;
  uint64_t batches_retried = 0;
};

// cuda_backend.hpp — the three entry points the reference application binds for its CUDA
// backend (reference cuda/include/fluid/cuda.h:7-9), re-created on top of the C ABI.
#pragma once

#include <cstddef>
#include <vector>

#ifdef PBF_USE_REFERENCE_HEADERS
#include "fluid/core.h"   // the reference's own header (drop-in build)
#else
#include "fluid_types.hpp"
#endif

namespace fluid {

int cuda_version();                                           // cuda.h:7
bool cuda_device_available(int* count, const char** error);   // cuda.h:8
void cuda_step(const Params& params, State& state);           // cuda.h:9

// Extensions used by this repo's own application (device-resident stepping, row f1 of
// SURVEY §8): the same backend object cuda_step uses, without the per-call host round trip.
namespace b200 {
struct Options {
  int device = 0;
  std::vector<int> devices;  // more than one entry: x-slabs over these devices (pbf_group_*)
  bool fast_mode = false;    // PBF_MODE_FAST instead of the bit-exact default
};
void configure(const Options& options);            // before the first step
void upload(const Params& params, const State& state);
void step_resident(const Params& params, int nsteps);   // no host traffic
void download_positions(State& state);             // pos_* only (what the VTK writer reads); blocking
// Overlapped frame output (one device): snapshot_begin enqueues a copy of the positions into one
// of two pinned buffers of the backend and returns at once — the copy runs under the next
// step_resident() batch; snapshot_wait blocks until it has landed.  The arrays stay valid until
// the next snapshot_begin on the same slot.  snapshots_available() is false for --devices runs.
struct Snapshot {
  const float* pos_x = nullptr;
  const float* pos_y = nullptr;
  const float* pos_z = nullptr;
  std::size_t count = 0;
  float time = 0.0f;
};
bool snapshots_available();
void snapshot_begin(int slot);
Snapshot snapshot_wait(int slot);
void download(State& state);                       // pos_* and vel_*
float device_time();
void set_device_time(float t);
void shutdown();
}  // namespace b200

}  // namespace fluid

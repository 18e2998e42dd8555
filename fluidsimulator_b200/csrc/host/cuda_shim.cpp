// cuda_shim.cpp — fluid::cuda_version / cuda_device_available / cuda_step over the C ABI.
//
// Replaces reference cuda/src/cuda_stub.cu:736-1099 behind the call site
// app/src/main.cpp:171-187 (probe) and :250-254 (step).  Contract kept:
//   - caller owns Params/State; on return pos_*/vel_* hold the new values in original particle
//     order and state.time is advanced by dt (float accumulate, core.cpp:614);
//   - state.cpu is left untouched (the reference CUDA backend does the same);
//   - empty state: only time advances (cuda_stub.cu:765-769).
// Difference: CUDA failures are not silent.  The reference checks no error codes; here any
// failure prints the reason to stderr and aborts — main() has no recovery path either way.
#include "cuda_backend.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pbf_b200.h"

namespace fluid {
namespace {

struct Backend {
  pbf_ctx* ctx = nullptr;            // the only context, or slab 0 of the group
  std::vector<pbf_ctx*> slabs;       // --devices: one context per device ...
  pbf_group* group = nullptr;        // ... linked as x-slabs
  b200::Options options;
  pbf_params last_params{};
  bool params_valid = false;
  std::vector<float> plane_cache;  // nx.., ny.., nz.., d.. of the planes last sent
  // State's six vectors, page-locked in place (cuda_step moves 48 B per particle per call)
  void* pinned_ptr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::size_t pinned_bytes[6] = {0, 0, 0, 0, 0, 0};
};

Backend& backend() {
  static Backend b;
  return b;
}

[[noreturn]] void die(const char* what, pbf_ctx* ctx) {
  std::fprintf(stderr, "pbf_b200 backend: %s failed: %s\n", what, pbf_last_error(ctx));
  std::abort();
}

void check(int rc, const char* what) {
  if (rc != PBF_OK) die(what, backend().ctx);
}

pbf_ctx* context(std::size_t capacity) {
  Backend& b = backend();
  if (!b.ctx) {
    const int mode = b.options.fast_mode ? PBF_MODE_FAST : PBF_MODE_STRICT;
    if (b.options.devices.size() > 1) {
      for (int dev : b.options.devices) {
        pbf_ctx* c = pbf_create(dev, 0);
        if (!c) die("pbf_create", nullptr);
        if (pbf_set_mode(c, mode) != PBF_OK) die("pbf_set_mode", c);
        b.slabs.push_back(c);
      }
      b.group = pbf_group_create(b.slabs.data(), static_cast<int>(b.slabs.size()));
      if (!b.group) die("pbf_group_create", nullptr);
      b.ctx = b.slabs[0];
    } else {
      b.ctx = pbf_create(b.options.device, capacity);
      if (!b.ctx) die("pbf_create", nullptr);
      check(pbf_set_mode(b.ctx, mode), "pbf_set_mode");
    }
  }
  return b.ctx;
}

// every context of the backend (one, or all slabs)
std::vector<pbf_ctx*> all_contexts() {
  Backend& b = backend();
  return b.slabs.empty() ? std::vector<pbf_ctx*>{b.ctx} : b.slabs;
}

void check_slabs(int rc, const char* what) {
  if (rc == PBF_OK) return;
  for (pbf_ctx* c : all_contexts()) {
    const char* msg = pbf_last_error(c);
    if (msg && msg[0]) die(what, c);
  }
  die(what, backend().ctx);
}

pbf_params to_pod(const Params& p) {
  pbf_params q;
  std::memset(&q, 0, sizeof(q));
  q.dt = p.dt;
  q.density = p.density;
  q.particle_mass = p.particle_mass;
  q.h = p.h;
  q.particle_radius = p.particle_radius;
  q.epsilon = p.epsilon;
  q.solver_iterations = p.solver_iterations;
  q.neighbor_reserve_factor = p.neighbor_reserve_factor;
  q.use_uniform_grid = p.use_uniform_grid ? 1 : 0;
  q.enable_scorr = p.enable_scorr ? 1 : 0;
  q.enable_xsph = p.enable_xsph ? 1 : 0;
  q.enable_vorticity = p.enable_vorticity ? 1 : 0;
  q.scorr_k = p.scorr_k;
  q.scorr_n = p.scorr_n;
  q.scorr_dq_coeff = p.scorr_dq_coeff;
  q.visc_c = p.visc_c;
  q.plane_restitution = p.plane_restitution;
  q.plane_friction = p.plane_friction;
  q.vort_epsilon = p.vort_epsilon;
  q.vort_norm_eps = p.vort_norm_eps;
  q.external_force[0] = p.external_forces.x;
  q.external_force[1] = p.external_forces.y;
  q.external_force[2] = p.external_forces.z;
  return q;
}

// Params are re-sent only when they changed (cuda_step receives them on every call).
void sync_params(const Params& params, std::size_t capacity) {
  Backend& b = backend();
  pbf_ctx* ctx = context(capacity);
  const std::size_t np = params.planes.size();
  std::vector<float> planes;
  planes.reserve(4 * np);
  planes.insert(planes.end(), params.planes.nx.begin(), params.planes.nx.end());
  planes.insert(planes.end(), params.planes.ny.begin(), params.planes.ny.end());
  planes.insert(planes.end(), params.planes.nz.begin(), params.planes.nz.end());
  planes.insert(planes.end(), params.planes.d.begin(), params.planes.d.end());
  (void)ctx;
  if (!b.params_valid || planes != b.plane_cache) {
    for (pbf_ctx* c : all_contexts())
      if (pbf_set_planes(c, static_cast<int>(np), planes.data(), planes.data() + np, planes.data() + 2 * np,
                         planes.data() + 3 * np) != PBF_OK)
        die("pbf_set_planes", c);
    b.plane_cache = planes;
    b.params_valid = false;
  }
  const pbf_params pod = to_pod(params);
  if (!b.params_valid || std::memcmp(&pod, &b.last_params, sizeof(pod)) != 0) {
    for (pbf_ctx* c : all_contexts())
      if (pbf_set_params(c, &pod) != PBF_OK) die("pbf_set_params", c);
    b.last_params = pod;
    b.params_valid = true;
  }
}

}  // namespace

int cuda_version() { return 1; }  // cuda_stub.cu:736-738

bool cuda_device_available(int* count, const char** error) {  // cuda_stub.cu:740-762
  if (count) *count = 0;
  if (error) *error = nullptr;
  const char* err = nullptr;
  const int n = pbf_device_count(&err);
  if (n < 0) {
    if (error) *error = err;
    return false;
  }
  if (count) *count = n;
  return n > 0;
}

namespace {

// The caller's vectors live across calls (the reference's main.cpp allocates State once and never
// resizes it, app/src/main.cpp:189-230): page-lock them where they are, so the six copies of the
// contract run at PCIe speed and pbf_step_host can replay the whole call as one graph.  CUDA
// requires page-locked memory to be unregistered BEFORE it is freed, which a library cannot
// guarantee for memory it does not own: a caller that resizes or destroys State between calls must
// set PBF_PIN_STATE=0 (copies then go through the driver's staging buffers, as for any pageable
// memory).  A changed data() pointer is detected and never unregistered through the stale address
// — the stale registration is dropped from the books, not touched.  Failure to register is not an
// error either.
bool pinning_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("PBF_PIN_STATE");
    return !(e && e[0] == '0');
  }();
  return on;
}

void pin_state(State& state) {
  if (!pinning_enabled()) return;
  Backend& b = backend();
  std::vector<float>* arrays[6] = {&state.pos_x, &state.pos_y, &state.pos_z, &state.vel_x, &state.vel_y, &state.vel_z};
  for (int a = 0; a < 6; ++a) {
    void* ptr = arrays[a]->data();
    const std::size_t bytes = arrays[a]->size() * sizeof(float);
    if (ptr == b.pinned_ptr[a] && bytes == b.pinned_bytes[a]) continue;
    // same storage, other size: still ours to unregister.  Other storage: the old block has been
    // freed by the vector — forget it (unregistering a freed range is undefined).
    if (b.pinned_ptr[a] && b.pinned_ptr[a] == ptr) pbf_host_unregister(b.ctx, b.pinned_ptr[a]);
    b.pinned_ptr[a] = nullptr;
    b.pinned_bytes[a] = 0;
    if (bytes && pbf_host_register(b.ctx, ptr, bytes) == PBF_OK) {
      b.pinned_ptr[a] = ptr;
      b.pinned_bytes[a] = bytes;
    }
  }
}

void unpin_state() {
  Backend& b = backend();
  for (int a = 0; a < 6; ++a) {
    if (b.pinned_ptr[a] && b.ctx) pbf_host_unregister(b.ctx, b.pinned_ptr[a]);
    b.pinned_ptr[a] = nullptr;
    b.pinned_bytes[a] = 0;
  }
}

}  // namespace

void cuda_step(const Params& params, State& state) {  // cuda_stub.cu:764-1099
  const std::size_t n = state.size();
  if (n == 0) {
    state.time += params.dt;
    return;
  }
  sync_params(params, n);
  pin_state(state);
  pbf_ctx* ctx = backend().ctx;
  check(pbf_set_time(ctx, state.time), "pbf_set_time");
  check(pbf_step_host(ctx, n, state.pos_x.data(), state.pos_y.data(), state.pos_z.data(), state.vel_x.data(),
                      state.vel_y.data(), state.vel_z.data(), 1),
        "pbf_step_host");
  state.time = pbf_time(ctx);
}

namespace b200 {

void configure(const Options& options) { backend().options = options; }

void upload(const Params& params, const State& state) {
  sync_params(params, state.size());
  Backend& b = backend();
  if (b.group) {
    check_slabs(pbf_group_upload(b.group, state.size(), state.pos_x.data(), state.pos_y.data(), state.pos_z.data(),
                                 state.vel_x.data(), state.vel_y.data(), state.vel_z.data()),
                "pbf_group_upload");
  } else {
    check(pbf_upload(b.ctx, state.size(), state.pos_x.data(), state.pos_y.data(), state.pos_z.data(),
                     state.vel_x.data(), state.vel_y.data(), state.vel_z.data()),
          "pbf_upload");
  }
  for (pbf_ctx* c : all_contexts())
    if (pbf_set_time(c, state.time) != PBF_OK) die("pbf_set_time", c);
}

void step_resident(const Params& params, int nsteps) {
  Backend& b = backend();
  sync_params(params, pbf_count(b.ctx));
  if (b.group)
    check_slabs(pbf_group_step(b.group, nsteps), "pbf_group_step");
  else
    check(pbf_step(b.ctx, nsteps), "pbf_step");
}

void download_positions(State& state) {
  Backend& b = backend();
  if (b.group)
    check_slabs(pbf_group_download(b.group, state.pos_x.data(), state.pos_y.data(), state.pos_z.data(), nullptr,
                                   nullptr, nullptr),
                "pbf_group_download");
  else
    check(pbf_download(b.ctx, state.pos_x.data(), state.pos_y.data(), state.pos_z.data(), nullptr, nullptr, nullptr),
          "pbf_download");
  state.time = pbf_time(b.ctx);
}

bool snapshots_available() { return backend().group == nullptr; }

void snapshot_begin(int slot) { check(pbf_snapshot_begin(backend().ctx, slot), "pbf_snapshot_begin"); }

Snapshot snapshot_wait(int slot) {
  Snapshot s;
  check(pbf_snapshot_wait(backend().ctx, slot, &s.pos_x, &s.pos_y, &s.pos_z, &s.count, &s.time), "pbf_snapshot_wait");
  return s;
}

void download(State& state) {
  Backend& b = backend();
  if (b.group)
    check_slabs(pbf_group_download(b.group, state.pos_x.data(), state.pos_y.data(), state.pos_z.data(),
                                   state.vel_x.data(), state.vel_y.data(), state.vel_z.data()),
                "pbf_group_download");
  else
    check(pbf_download(b.ctx, state.pos_x.data(), state.pos_y.data(), state.pos_z.data(), state.vel_x.data(),
                       state.vel_y.data(), state.vel_z.data()),
          "pbf_download");
  state.time = pbf_time(b.ctx);
}

float device_time() { return pbf_time(backend().ctx); }
void set_device_time(float t) { check(pbf_set_time(backend().ctx, t), "pbf_set_time"); }

void shutdown() {
  Backend& b = backend();
  unpin_state();
  if (b.group) {
    pbf_group_destroy(b.group);  // before its contexts
    b.group = nullptr;
    for (pbf_ctx* c : b.slabs) pbf_destroy(c);
    b.slabs.clear();
  } else if (b.ctx) {
    pbf_destroy(b.ctx);
  }
  b.ctx = nullptr;
  b.params_valid = false;
}

}  // namespace b200
}  // namespace fluid

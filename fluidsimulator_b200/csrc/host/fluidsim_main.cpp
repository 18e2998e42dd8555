// fluidsim_main.cpp — the application around the B200 backend: same 16 options, validation,
// stdout lines and output cadence as the reference application (reference app/src/main.cpp),
// so a user of `fluidsim --backend=cuda` can switch binaries.  Differences, all additive:
//   --solver-iterations N   Params::solver_iterations has no flag in the reference (core.h:25);
//                           BASELINE.json's iteration sweep needs one.
//   --mode strict|fast      arithmetic mode of the backend (strict = bit-identical to the CPU path).
//   --device N              CUDA device index.
//   --per-step-host         use the reference's cuda_step contract (host round trip per step)
//                           instead of device-resident stepping.
//   --devices 0,1,2,3       split the scene into x-slabs over several GPUs of this process
//                           (halo exchange by peer copies; results identical to one GPU).
// Device-resident runs execute all substeps up to the next output step as ONE batch (no host
// synchronisation in between), and frames are rendered and written by a background thread while
// the GPU already computes the next batch (row f1 of SURVEY §8: the ASCII writer costs ~78 bytes
// per particle and dominated the wall time of a run with output).
// The state stays on the GPU between steps; positions come back only on output steps
// (main.cpp:259-273 reads state.pos_* only there).  --backend=cpu is refused: this product
// has no CPU fallback (the reference's own binary provides it).
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cuda_backend.hpp"
#include "frame_writer.hpp"
#include "options.hpp"
#include "scene_loader.hpp"

namespace {

using fluid::b200::Option;

bool parse_double(const fluid::b200::ParsedOptions& parsed, const char* name, double& out) {
  try {
    out = std::stod(parsed.value(name, ""));
  } catch (const std::exception&) {
    std::cerr << "Invalid " << name << " value." << std::endl;
    return false;
  }
  return true;
}

// Frames are handed to one writer thread; at most two are in flight, so a slow disk throttles
// the simulation instead of exhausting memory.  A job either owns its arrays (copies of State) or
// points at one of the backend's two pinned snapshot buffers (`slot` >= 0): that buffer is not
// reused before wait_slot() has seen the job written.
struct FrameJob {
  std::vector<float> x, y, z;
  const float* px = nullptr;
  const float* py = nullptr;
  const float* pz = nullptr;
  std::size_t count = 0;
  int slot = -1;
  double time = 0.0;
  std::size_t index = 0;
};

class AsyncFrames {
 public:
  explicit AsyncFrames(const fluid::b200::FrameWriter& writer) : writer_(writer), thread_([this] { run(); }) {}
  ~AsyncFrames() { finish(); }
  void submit(std::unique_ptr<FrameJob> job) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return queue_.size() < 2; });
    if (job->slot >= 0) slot_busy_[job->slot] = true;
    queue_.push_back(std::move(job));
    cv_.notify_all();
  }
  // blocks until the frame that reads snapshot buffer `slot` is on disk
  void wait_slot(int slot) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !slot_busy_[slot]; });
  }
  bool finish() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (done_) return ok_;
      done_ = true;
      cv_.notify_all();
    }
    thread_.join();
    return ok_;
  }

 private:
  void run() {
    for (;;) {
      std::unique_ptr<FrameJob> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !queue_.empty() || done_; });
        if (queue_.empty()) return;
        job = std::move(queue_.front());
        queue_.pop_front();
        cv_.notify_all();
      }
      fluid::b200::FrameView view;
      view.pos_x = job->px ? job->px : job->x.data();
      view.pos_y = job->py ? job->py : job->y.data();
      view.pos_z = job->pz ? job->pz : job->z.data();
      view.count = job->px ? job->count : job->x.size();
      view.time = job->time;
      if (!writer_.write(view, job->index)) ok_ = false;
      if (job->slot >= 0) {
        std::lock_guard<std::mutex> lk(mu_);
        slot_busy_[job->slot] = false;
        cv_.notify_all();
      }
    }
  }
  const fluid::b200::FrameWriter& writer_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::unique_ptr<FrameJob>> queue_;
  bool slot_busy_[2] = {false, false};
  bool done_ = false;
  bool ok_ = true;
  std::thread thread_;
};

bool parse_devices(const std::string& text, std::vector<int>& out) {
  out.clear();
  std::size_t pos = 0;
  while (pos <= text.size()) {
    const std::size_t comma = text.find(',', pos);
    const std::string item = text.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
    try {
      std::size_t used = 0;
      const int d = std::stoi(item, &used);
      if (used != item.size() || d < 0) return false;
      out.push_back(d);
    } catch (const std::exception&) {
      return false;
    }
    if (comma == std::string::npos) break;
    pos = comma + 1;
  }
  return !out.empty();
}

}  // namespace

int main(int argc, char** argv) {
  const std::vector<Option> options = {
      {"backend", true, "Backend to use: cuda (this build has no CPU backend)."},
      {"no-output", false, "Disable output for benchmarking."},
      {"debug-print", false, "Print per-step timing."},
      {"steps", true, "Number of simulation steps to run."},
      {"steps-per-sec", true, "Simulation steps per second (sets dt = 1 / value)."},
      {"enable-scorr", false, "Enable s_corr constraint."},
      {"enable-xsph", false, "Enable XSPH viscosity."},
      {"enable-vorticity", false, "Enable vorticity confinement."},
      {"plane-restitution", true, "Restitution for plane collisions (0 = no bounce)."},
      {"plane-friction", true, "Tangential damping for plane collisions (0..1)."},
      {"threads", true, "Accepted for compatibility; the GPU backend does not use host threads."},
      {"no-omp", false, "Accepted for compatibility; the GPU backend does not use OpenMP."},
      {"fps", true, "Output frames per second (controls output stride)."},
      {"duration", true, "Simulation duration in seconds (overrides --steps)."},
      {"scene", true, "Scene JSON to load (legacy format)."},
      {"output-dir", true, "Directory for VTK output (for ParaView)."},
      {"solver-iterations", true, "Constraint solver iterations per step (default 4)."},
      {"mode", true, "Arithmetic mode: strict (bit-identical to the CPU path, default) or fast."},
      {"device", true, "CUDA device index (default 0)."},
      {"per-step-host", false, "Round-trip the state through host memory every step (cuda_step contract)."},
      {"devices", true, "Comma-separated CUDA devices: split the scene into x-slabs over them (e.g. 0,1,2,3)."},
  };
  const fluid::b200::ParsedOptions parsed = fluid::b200::parse_options(argc, argv, options);
  if (!parsed.ok) {
    std::cerr << parsed.error << std::endl;
    std::cerr << fluid::b200::usage_text(argv[0], options);
    return 1;
  }
  if (parsed.has("help")) {
    std::cout << fluid::b200::usage_text(argv[0], options);
    return 0;
  }

  const std::string backend = parsed.value("backend", "cuda");
  const bool output_enabled = !parsed.has("no-output");
  const bool debug_print = parsed.has("debug-print");
  const std::string output_dir = parsed.value("output-dir", "output");
  const std::string scene_path = parsed.value("scene", "");
  const double nan = std::numeric_limits<double>::quiet_NaN();
  double steps_per_sec = nan, fps = -1.0, duration = -1.0, restitution = nan, friction = nan;
  int steps = 1, iterations = -1, device = 0;
  try {
    steps = std::stoi(parsed.value("steps", "1"));
  } catch (const std::exception&) {
    std::cerr << "Invalid steps value." << std::endl;
    std::cerr << fluid::b200::usage_text(argv[0], options);
    return 1;
  }
  if (steps < 1) {
    std::cerr << "Steps must be >= 1." << std::endl;
    return 1;
  }
  if (parsed.has("steps-per-sec")) {
    if (!parse_double(parsed, "steps-per-sec", steps_per_sec)) return 1;
    if (!(steps_per_sec > 0.0)) {
      std::cerr << "steps-per-sec must be > 0." << std::endl;
      return 1;
    }
  }
  if (parsed.has("fps")) {
    if (!parse_double(parsed, "fps", fps)) return 1;
    if (!(fps > 0.0)) {
      std::cerr << "fps must be > 0." << std::endl;
      return 1;
    }
  }
  if (parsed.has("duration")) {
    if (!parse_double(parsed, "duration", duration)) return 1;
    if (!(duration > 0.0)) {
      std::cerr << "duration must be > 0." << std::endl;
      return 1;
    }
  }
  if (parsed.has("plane-restitution")) {
    if (!parse_double(parsed, "plane-restitution", restitution)) return 1;
    if (restitution < 0.0) {
      std::cerr << "plane-restitution must be >= 0." << std::endl;
      return 1;
    }
  }
  if (parsed.has("plane-friction")) {
    if (!parse_double(parsed, "plane-friction", friction)) return 1;
    if (friction < 0.0 || friction > 1.0) {
      std::cerr << "plane-friction must be in [0, 1]." << std::endl;
      return 1;
    }
  }
  if (parsed.has("threads")) {
    int threads = 0;
    try {
      threads = std::stoi(parsed.value("threads", ""));
    } catch (const std::exception&) {
      std::cerr << "Invalid threads value." << std::endl;
      return 1;
    }
    if (threads < 1) {
      std::cerr << "threads must be >= 1." << std::endl;
      return 1;
    }
  }
  if (parsed.has("solver-iterations")) {
    try {
      iterations = std::stoi(parsed.value("solver-iterations", ""));
    } catch (const std::exception&) {
      std::cerr << "Invalid solver-iterations value." << std::endl;
      return 1;
    }
    if (iterations < 0) {
      std::cerr << "solver-iterations must be >= 0." << std::endl;
      return 1;
    }
  }
  if (parsed.has("device")) {
    try {
      device = std::stoi(parsed.value("device", ""));
    } catch (const std::exception&) {
      std::cerr << "Invalid device value." << std::endl;
      return 1;
    }
  }
  std::vector<int> devices;
  if (parsed.has("devices")) {
    if (!parse_devices(parsed.value("devices", ""), devices)) {
      std::cerr << "Invalid devices value (expected e.g. 0,1,2,3)." << std::endl;
      return 1;
    }
    if (parsed.has("per-step-host")) {
      std::cerr << "--devices and --per-step-host exclude each other." << std::endl;
      return 1;
    }
    device = devices[0];
  }
  const std::string mode = parsed.value("mode", "strict");
  if (mode != "strict" && mode != "fast") {
    std::cerr << "mode must be strict or fast." << std::endl;
    return 1;
  }
  if (backend == "cpu") {
    std::cerr << "Unsupported backend: cpu (this build is the B200 CUDA backend only; there is no CPU fallback)"
              << std::endl;
    return 1;
  }
  if (backend != "cuda") {
    std::cerr << "Unsupported backend: " << backend << std::endl;
    std::cerr << fluid::b200::usage_text(argv[0], options);
    return 1;
  }

  std::cout << "FluidSimulator rewrite scaffold" << std::endl;
  std::cout << "core_version=" << fluid::core_version_b200() << std::endl;
  int device_count = 0;
  const char* cuda_error = nullptr;
  if (!fluid::cuda_device_available(&device_count, &cuda_error)) {
    if (cuda_error)
      std::cerr << "CUDA backend unavailable: " << cuda_error << std::endl;
    else
      std::cerr << "CUDA backend unavailable: no CUDA devices detected." << std::endl;
    return 1;
  }
  std::cout << "backend=cuda" << std::endl;
  std::cout << "cuda_devices=" << device_count << std::endl;

  fluid::Params params;
  params.backend = fluid::Params::Backend::Cuda;
  fluid::State state;
  if (!scene_path.empty()) {
    std::string error;
    if (!fluid::b200::load_scene_json(scene_path, params, state, &error)) {
      std::cerr << "Failed to load scene: " << error << std::endl;
      return 1;
    }
  } else {
    fluid::b200::default_test_scene(params, state);
  }
  if (steps_per_sec == steps_per_sec) params.dt = static_cast<float>(1.0 / steps_per_sec);  // main.cpp:203-205
  if (duration > 0.0) {                                                                       // main.cpp:206-211
    steps = static_cast<int>(std::ceil(duration / params.dt));
    if (steps < 1) steps = 1;
  }
  if (parsed.has("enable-scorr")) params.enable_scorr = true;
  if (parsed.has("enable-xsph")) params.enable_xsph = true;
  if (parsed.has("enable-vorticity")) params.enable_vorticity = true;
  if (restitution == restitution) params.plane_restitution = static_cast<float>(restitution);
  if (friction == friction) params.plane_friction = static_cast<float>(friction);
  if (iterations >= 0) params.solver_iterations = iterations;

  int output_interval = 1;  // main.cpp:239-244
  if (output_enabled && fps > 0.0) {
    const double steps_per_frame = 1.0 / (params.dt * fps);
    output_interval = std::max(1, static_cast<int>(std::lround(steps_per_frame)));
  }
  fluid::b200::FrameWriter frames(output_dir, "frame");
  fluid::b200::SeriesWriter series(output_dir, "series");
  std::size_t frame_index = 0;

  fluid::b200::Options backend_options;
  backend_options.device = device;
  backend_options.devices = devices;
  backend_options.fast_mode = (mode == "fast");
  fluid::b200::configure(backend_options);
  const bool per_step_host = parsed.has("per-step-host");
  const bool resident = !per_step_host && state.size() > 0;
  if (resident) fluid::b200::upload(params, state);
  if (devices.size() > 1) std::cout << "slabs=" << devices.size() << std::endl;

  AsyncFrames async_frames(frames);
  // Overlapped output (one device, resident stepping): the positions of a frame are copied into
  // one of two pinned buffers on a copy stream WHILE the next batch of substeps runs; the frame
  // goes to the writer thread one batch later.  PBF_BLOCKING_FRAMES=1 keeps the blocking download.
  const char* blocking_env = std::getenv("PBF_BLOCKING_FRAMES");
  const bool overlapped = resident && fluid::b200::snapshots_available() && !(blocking_env && blocking_env[0] == '1');
  int pending_slot = -1;           // snapshot begun but not yet handed to the writer
  std::size_t pending_index = 0;
  auto flush_pending = [&]() {
    if (pending_slot < 0) return;
    const fluid::b200::Snapshot snap = fluid::b200::snapshot_wait(pending_slot);
    std::unique_ptr<FrameJob> job(new FrameJob());
    job->px = snap.pos_x;
    job->py = snap.pos_y;
    job->pz = snap.pos_z;
    job->count = snap.count;
    job->slot = pending_slot;
    job->time = snap.time;
    job->index = pending_index;
    async_frames.submit(std::move(job));
    pending_slot = -1;
  };
  int step = 0;
  while (step < steps) {
    const auto t0 = std::chrono::steady_clock::now();
    // steps [step, last] run as one batch; `last` is the next step after which a frame is due
    int last = step;
    if (resident && !debug_print) {
      last = steps - 1;
      if (output_enabled) {
        const int next_frame = ((step + output_interval - 1) / output_interval) * output_interval;
        last = std::min(last, next_frame);
      }
    }
    const int batch = last - step + 1;
    const bool wants_frame = output_enabled && (last % output_interval == 0);
    if (!resident) {
      fluid::cuda_step(params, state);
    } else {
      fluid::b200::step_resident(params, batch);   // the previous frame's copy runs under this batch
      if (overlapped) flush_pending();
      if (wants_frame && !overlapped) fluid::b200::download_positions(state);
    }
    step = last + 1;
    const auto t1 = std::chrono::steady_clock::now();
    const double step_ms = std::chrono::duration<double, std::milli>(t1 - t0).count() / batch;
    if (wants_frame && overlapped) {
      const int slot = static_cast<int>(frame_index & 1);
      async_frames.wait_slot(slot);                // the frame that used this buffer is on disk
      fluid::b200::snapshot_begin(slot);
      pending_slot = slot;
      pending_index = frame_index;
      series.add(fluid::b200::device_time(), fluid::b200::frame_filename("frame", frame_index));
      frame_index++;
    } else if (wants_frame) {
      std::unique_ptr<FrameJob> job(new FrameJob());
      job->x = state.pos_x;
      job->y = state.pos_y;
      job->z = state.pos_z;
      job->time = state.time;
      job->index = frame_index;
      async_frames.submit(std::move(job));
      series.add(state.time, fluid::b200::frame_filename("frame", frame_index));
      frame_index++;
    }
    if (debug_print) std::cout << "step_done=" << step << " step_ms=" << step_ms << std::endl;
  }
  flush_pending();
  if (!async_frames.finish()) {
    std::cerr << "Failed to write VTK frame." << std::endl;
    return 1;
  }
  if (!per_step_host && state.size() > 0) fluid::b200::download(state);
  std::cout << "particle_count=" << state.size() << std::endl;
  std::cout << "end_time=" << state.time << std::endl;
  std::cout << "output_enabled=" << (output_enabled ? "true" : "false") << std::endl;
  if (output_enabled && !series.write()) {
    std::cerr << "Failed to write PVD index." << std::endl;
    return 1;
  }
  fluid::b200::shutdown();
  return 0;
}

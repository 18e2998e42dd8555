// host_tools.cpp — small command-line probes of the host layer, used by the tests:
//   pbf_host_tools scene <file.json>            prints particle count, params and FNV-1a
//                                               hashes of the emitted arrays and planes
//   pbf_host_tools test-scene                   same for the built-in test scene
//   pbf_host_tools frame <file.json> <out_dir>  writes frame_000000.vtp + series.pvd of t0
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "frame_writer.hpp"
#include "scene_loader.hpp"

namespace {

std::uint64_t fnv1a(const void* data, std::size_t bytes, std::uint64_t h = 1469598103934665603ull) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (std::size_t i = 0; i < bytes; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

std::uint64_t hash_vec(const std::vector<float>& v) { return fnv1a(v.data(), v.size() * sizeof(float)); }

unsigned bits(float f) {
  unsigned u;
  std::memcpy(&u, &f, 4);
  return u;
}

void dump(const fluid::Params& p, const fluid::State& s) {
  std::printf("count=%zu\n", s.size());
  std::printf("pos_x=%016llx\npos_y=%016llx\npos_z=%016llx\n", (unsigned long long)hash_vec(s.pos_x),
              (unsigned long long)hash_vec(s.pos_y), (unsigned long long)hash_vec(s.pos_z));
  std::printf("vel_x=%016llx\nvel_y=%016llx\nvel_z=%016llx\n", (unsigned long long)hash_vec(s.vel_x),
              (unsigned long long)hash_vec(s.vel_y), (unsigned long long)hash_vec(s.vel_z));
  std::printf("planes=%zu\n", p.planes.size());
  std::printf("plane_nx=%016llx\nplane_ny=%016llx\nplane_nz=%016llx\nplane_d=%016llx\n",
              (unsigned long long)hash_vec(p.planes.nx), (unsigned long long)hash_vec(p.planes.ny),
              (unsigned long long)hash_vec(p.planes.nz), (unsigned long long)hash_vec(p.planes.d));
  std::printf("particle_mass=%08x\ndensity=%08x\nh=%08x\nparticle_radius=%08x\nepsilon=%08x\n", bits(p.particle_mass),
              bits(p.density), bits(p.h), bits(p.particle_radius), bits(p.epsilon));
  std::printf("scorr_n=%d\nscorr_k=%08x\nvisc_c=%08x\n", p.scorr_n, bits(p.scorr_k), bits(p.visc_c));
  std::printf("force=%08x,%08x,%08x\n", bits(p.external_forces.x), bits(p.external_forces.y), bits(p.external_forces.z));
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s scene <json> | test-scene | frame <json> <dir>\n", argv[0]);
    return 2;
  }
  const std::string cmd = argv[1];
  fluid::Params params;
  fluid::State state;
  std::string error;
  if (cmd == "test-scene") {
    fluid::b200::default_test_scene(params, state);
    dump(params, state);
    return 0;
  }
  if ((cmd == "scene" && argc >= 3) || (cmd == "frame" && argc >= 4)) {
    if (!fluid::b200::load_scene_json(argv[2], params, state, &error)) {
      std::fprintf(stderr, "Failed to load scene: %s\n", error.c_str());
      return 1;
    }
    if (cmd == "scene") {
      dump(params, state);
      return 0;
    }
    fluid::b200::FrameWriter frames(argv[3]);
    fluid::b200::SeriesWriter series(argv[3]);
    fluid::b200::FrameView view;
    view.pos_x = state.pos_x.data();
    view.pos_y = state.pos_y.data();
    view.pos_z = state.pos_z.data();
    view.count = state.size();
    view.time = static_cast<float>(1.0 / 120.0);  // State::time is a float (core.h:129)
    if (!frames.write(view, 0)) return 1;
    series.add(view.time, fluid::b200::frame_filename("frame", 0));
    return series.write() ? 0 : 1;
  }
  std::fprintf(stderr, "unknown command\n");
  return 2;
}

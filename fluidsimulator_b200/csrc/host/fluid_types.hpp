// fluid_types.hpp — the boundary types of the PBF substep, laid out field for field like the
// reference's (reference core/include/fluid/core.h:11-132) so that code compiled against this
// header and code compiled against the reference header agree on the object layout.
//
// Only what the substep boundary needs is here: Params (physics knobs + plane SoA), State (six
// SoA float vectors + time) and the CpuScratch member State carries (the CUDA path leaves it
// untouched, as the reference CUDA backend does).  The CPU solver itself is NOT part of this
// product (there is no CPU fallback).
#pragma once

#include <cmath>
#include <cstddef>
#include <vector>

namespace fluid {

struct Params {
  float dt = 1.0f / 60.0f;                   // core.h:12
  enum class Backend { Cpu, Cuda };          // core.h:13-16
  Backend backend = Backend::Cpu;            // core.h:17
  float density = 6000.0f;                   // core.h:18
  float particle_mass = 0.0f;                // core.h:20
  float h = 0.0f;                            // core.h:22
  float particle_radius = 0.01f;             // core.h:23
  float epsilon = 600.0f;                    // core.h:24
  int solver_iterations = 4;                 // core.h:25
  float neighbor_reserve_factor = 1.5f;      // core.h:26
  bool use_uniform_grid = true;              // core.h:27
  bool enable_scorr = false;                 // core.h:28
  bool enable_xsph = false;                  // core.h:29
  bool enable_vorticity = false;             // core.h:30
  float scorr_k = 0.00005f;                  // core.h:31
  int scorr_n = 4;                           // core.h:32
  float scorr_dq_coeff = 0.3f;               // core.h:33
  float visc_c = 0.0002f;                    // core.h:34
  float plane_restitution = 0.0f;            // core.h:35
  float plane_friction = 0.0f;               // core.h:36
  float vort_epsilon = 0.5f;                 // core.h:37
  float vort_norm_eps = 1e-6f;               // core.h:38
  struct Vec3 {                              // core.h:39-43
    float x = 0.0f;
    float y = -9.8f;
    float z = 0.0f;
  } external_forces;
  struct PlaneSoA {                          // core.h:44-78
    std::vector<float> nx, ny, nz, d;
    std::size_t size() const { return nx.size(); }
    void clear() { nx.clear(); ny.clear(); nz.clear(); d.clear(); }
    void add(float a, float b, float c, float dist) {
      nx.push_back(a); ny.push_back(b); nz.push_back(c); d.push_back(dist);
    }
    void add_normalized(float a, float b, float c, float dist) {  // core.h:69-77
      const float len_sq = a * a + b * b + c * c;
      if (len_sq > 0.0f) {
        const float inv_len = 1.0f / std::sqrt(len_sq);
        add(a * inv_len, b * inv_len, c * inv_len, dist);
      } else {
        add(a, b, c, dist);
      }
    }
  } planes;
};

struct CpuScratch {                          // core.h:81-119 (layout only)
  struct CellKey { int x = 0, y = 0, z = 0; };
  struct CellEntry { CellKey key; int particle = 0; };
  std::vector<float> pred_x, pred_y, pred_z, delta_x, delta_y, delta_z, lambda, rho;
  std::vector<float> dv_x, dv_y, dv_z, omega_x, omega_y, omega_z, omega_mag, eta_x, eta_y, eta_z;
  std::vector<int> neighbor_indices, neighbor_prefix_sum;
  std::vector<CellEntry> grid_entries;
  std::vector<CellKey> grid_keys;
  std::vector<int> grid_starts, grid_ends;
};

struct State {                               // core.h:121-132
  std::vector<float> pos_x, pos_y, pos_z, vel_x, vel_y, vel_z;
  CpuScratch cpu;
  float time = 0.0f;
  std::size_t size() const { return pos_x.size(); }
};

// core.h:9 — printed by the application as core_version=
inline int core_version_b200() { return 1; }

}  // namespace fluid

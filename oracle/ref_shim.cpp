// ref_shim.cpp — extern "C" wrapper around the UNMODIFIED reference CPU solver.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).  This file contains no solver
// logic: it owns a fluid::Params + fluid::State pair and forwards to the reference's
// own fluid::step (reference core/src/core.cpp:119), fluid::init_scene_from_json
// (core/src/init.cpp:158) and fluid::init_test_scene (init.cpp:119).  The reference
// sources are compiled from /root/reference by oracle/Makefile into oracle/_ref/.
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "fluid/core.h"
#include "fluid/init.h"
#include "oracle_api.h"

struct oracle_sim {
  fluid::Params params;
  fluid::State state;
};

extern "C" {

const char* oracle_kind(void) { return "reference"; }

int oracle_has_openmp(void) {
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}

void oracle_set_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

oracle_sim* oracle_create(void) {
  oracle_sim* sim = new oracle_sim();
  sim->state = fluid::make_state(0);
  return sim;
}

void oracle_destroy(oracle_sim* sim) { delete sim; }

int oracle_load_scene(oracle_sim* sim, const char* json_path, char* err, size_t errlen) {
  std::string error;
  const bool ok = fluid::init_scene_from_json(json_path, sim->params, sim->state, &error);
  if (!ok && err && errlen > 0) {
    std::strncpy(err, error.c_str(), errlen - 1);
    err[errlen - 1] = '\0';
  }
  return ok ? 0 : -1;
}

int oracle_init_test_scene(oracle_sim* sim) {
  fluid::init_test_scene(sim->params, sim->state);
  return 0;
}

void oracle_set_params(oracle_sim* sim, const pbf_params* p) {
  fluid::Params& q = sim->params;
  q.dt = p->dt;
  q.density = p->density;
  q.particle_mass = p->particle_mass;
  q.h = p->h;
  q.particle_radius = p->particle_radius;
  q.epsilon = p->epsilon;
  q.solver_iterations = p->solver_iterations;
  q.neighbor_reserve_factor = p->neighbor_reserve_factor;
  q.use_uniform_grid = p->use_uniform_grid != 0;
  q.enable_scorr = p->enable_scorr != 0;
  q.enable_xsph = p->enable_xsph != 0;
  q.enable_vorticity = p->enable_vorticity != 0;
  q.scorr_k = p->scorr_k;
  q.scorr_n = p->scorr_n;
  q.scorr_dq_coeff = p->scorr_dq_coeff;
  q.visc_c = p->visc_c;
  q.plane_restitution = p->plane_restitution;
  q.plane_friction = p->plane_friction;
  q.vort_epsilon = p->vort_epsilon;
  q.vort_norm_eps = p->vort_norm_eps;
  q.external_forces.x = p->external_force[0];
  q.external_forces.y = p->external_force[1];
  q.external_forces.z = p->external_force[2];
}

void oracle_get_params(const oracle_sim* sim, pbf_params* p) {
  const fluid::Params& q = sim->params;
  p->dt = q.dt;
  p->density = q.density;
  p->particle_mass = q.particle_mass;
  p->h = q.h;
  p->particle_radius = q.particle_radius;
  p->epsilon = q.epsilon;
  p->solver_iterations = q.solver_iterations;
  p->neighbor_reserve_factor = q.neighbor_reserve_factor;
  p->use_uniform_grid = q.use_uniform_grid ? 1 : 0;
  p->enable_scorr = q.enable_scorr ? 1 : 0;
  p->enable_xsph = q.enable_xsph ? 1 : 0;
  p->enable_vorticity = q.enable_vorticity ? 1 : 0;
  p->scorr_k = q.scorr_k;
  p->scorr_n = q.scorr_n;
  p->scorr_dq_coeff = q.scorr_dq_coeff;
  p->visc_c = q.visc_c;
  p->plane_restitution = q.plane_restitution;
  p->plane_friction = q.plane_friction;
  p->vort_epsilon = q.vort_epsilon;
  p->vort_norm_eps = q.vort_norm_eps;
  p->external_force[0] = q.external_forces.x;
  p->external_force[1] = q.external_forces.y;
  p->external_force[2] = q.external_forces.z;
}

void oracle_set_planes(oracle_sim* sim, int count, const float* nx, const float* ny,
                       const float* nz, const float* d) {
  sim->params.planes.clear();
  for (int i = 0; i < count; ++i) sim->params.planes.add(nx[i], ny[i], nz[i], d[i]);
}

int oracle_plane_count(const oracle_sim* sim) {
  return static_cast<int>(sim->params.planes.size());
}

void oracle_get_planes(const oracle_sim* sim, float* nx, float* ny, float* nz, float* d) {
  const auto& pl = sim->params.planes;
  for (std::size_t i = 0; i < pl.size(); ++i) {
    nx[i] = pl.nx[i];
    ny[i] = pl.ny[i];
    nz[i] = pl.nz[i];
    d[i] = pl.d[i];
  }
}

void oracle_set_state(oracle_sim* sim, size_t n, const float* px, const float* py,
                      const float* pz, const float* vx, const float* vy, const float* vz) {
  fluid::State& s = sim->state;
  s.pos_x.assign(px, px + n);
  s.pos_y.assign(py, py + n);
  s.pos_z.assign(pz, pz + n);
  s.vel_x.assign(vx, vx + n);
  s.vel_y.assign(vy, vy + n);
  s.vel_z.assign(vz, vz + n);
  s.cpu.resize(n);
}

size_t oracle_count(const oracle_sim* sim) { return sim->state.size(); }

static void copy_out(const std::vector<float>& v, float* out) {
  if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(float));
}

void oracle_get_state(const oracle_sim* sim, float* px, float* py, float* pz, float* vx,
                      float* vy, float* vz) {
  const fluid::State& s = sim->state;
  copy_out(s.pos_x, px);
  copy_out(s.pos_y, py);
  copy_out(s.pos_z, pz);
  copy_out(s.vel_x, vx);
  copy_out(s.vel_y, vy);
  copy_out(s.vel_z, vz);
}

float oracle_time(const oracle_sim* sim) { return sim->state.time; }
void oracle_set_time(oracle_sim* sim, float t) { sim->state.time = t; }

void oracle_step(oracle_sim* sim, int nsteps) {
  for (int s = 0; s < nsteps; ++s) fluid::step(sim->params, sim->state);
}

size_t oracle_ncells(const oracle_sim* sim) { return sim->state.cpu.grid_keys.size(); }
size_t oracle_nneighbors(const oracle_sim* sim) {
  return sim->state.cpu.neighbor_indices.size();
}

void oracle_get_grid(const oracle_sim* sim, int32_t* entry_cx, int32_t* entry_cy,
                     int32_t* entry_cz, int32_t* entry_particle, int32_t* cell_xyz,
                     int32_t* cell_start, int32_t* cell_end) {
  const fluid::CpuScratch& c = sim->state.cpu;
  const std::size_t n = sim->state.size();
  for (std::size_t i = 0; i < n && i < c.grid_entries.size(); ++i) {
    if (entry_cx) entry_cx[i] = c.grid_entries[i].key.x;
    if (entry_cy) entry_cy[i] = c.grid_entries[i].key.y;
    if (entry_cz) entry_cz[i] = c.grid_entries[i].key.z;
    if (entry_particle) entry_particle[i] = c.grid_entries[i].particle;
  }
  for (std::size_t k = 0; k < c.grid_keys.size(); ++k) {
    if (cell_xyz) {
      cell_xyz[3 * k + 0] = c.grid_keys[k].x;
      cell_xyz[3 * k + 1] = c.grid_keys[k].y;
      cell_xyz[3 * k + 2] = c.grid_keys[k].z;
    }
    if (cell_start) cell_start[k] = c.grid_starts[k];
    if (cell_end) cell_end[k] = c.grid_ends[k];
  }
}

void oracle_get_neighbors(const oracle_sim* sim, int32_t* prefix_sum, int32_t* indices) {
  const fluid::CpuScratch& c = sim->state.cpu;
  if (prefix_sum && !c.neighbor_prefix_sum.empty())
    std::memcpy(prefix_sum, c.neighbor_prefix_sum.data(),
                c.neighbor_prefix_sum.size() * sizeof(int));
  if (indices && !c.neighbor_indices.empty())
    std::memcpy(indices, c.neighbor_indices.data(), c.neighbor_indices.size() * sizeof(int));
}

void oracle_get_scratch(const oracle_sim* sim, int id, float* out) {
  const fluid::CpuScratch& c = sim->state.cpu;
  const std::vector<float>* v = nullptr;
  switch (id) {
    case PBF_SCRATCH_PRED_X: v = &c.pred_x; break;
    case PBF_SCRATCH_PRED_Y: v = &c.pred_y; break;
    case PBF_SCRATCH_PRED_Z: v = &c.pred_z; break;
    case PBF_SCRATCH_DELTA_X: v = &c.delta_x; break;
    case PBF_SCRATCH_DELTA_Y: v = &c.delta_y; break;
    case PBF_SCRATCH_DELTA_Z: v = &c.delta_z; break;
    case PBF_SCRATCH_LAMBDA: v = &c.lambda; break;
    case PBF_SCRATCH_RHO: v = &c.rho; break;
    case PBF_SCRATCH_DV_X: v = &c.dv_x; break;
    case PBF_SCRATCH_DV_Y: v = &c.dv_y; break;
    case PBF_SCRATCH_DV_Z: v = &c.dv_z; break;
    case PBF_SCRATCH_OMEGA_X: v = &c.omega_x; break;
    case PBF_SCRATCH_OMEGA_Y: v = &c.omega_y; break;
    case PBF_SCRATCH_OMEGA_Z: v = &c.omega_z; break;
    case PBF_SCRATCH_OMEGA_MAG: v = &c.omega_mag; break;
    case PBF_SCRATCH_ETA_X: v = &c.eta_x; break;
    case PBF_SCRATCH_ETA_Y: v = &c.eta_y; break;
    case PBF_SCRATCH_ETA_Z: v = &c.eta_z; break;
    default: break;
  }
  if (v) copy_out(*v, out);
}

}  // extern "C"

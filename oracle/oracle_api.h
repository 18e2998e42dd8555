/*
 * oracle_api.h — C interface shared by the two CPU oracles of the PBF substep.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load these libraries, and only as the checker or the timed CPU baseline.
 *
 *   oracle/_ref/libpbf_oracle_ref.so   kind "reference": the UNMODIFIED reference
 *       sources (/root/reference/core/src/core.cpp + init.cpp) compiled where they
 *       lie and wrapped by oracle/ref_shim.cpp.
 *   oracle/_build/libpbf_oracle_port.so kind "port": oracle/pbf_oracle.c, a plain-C
 *       restatement of fluid::step (reference core/src/core.cpp:119-615).
 *
 * Both export exactly the symbols below, so a test can run either.
 */
#ifndef PBF_ORACLE_API_H
#define PBF_ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#include "pbf_b200.h" /* pbf_params, pbf_scratch_id */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_sim oracle_sim;

const char* oracle_kind(void); /* "reference" or "port" */
int oracle_has_openmp(void);
void oracle_set_threads(int nthreads);
int oracle_max_threads(void);

oracle_sim* oracle_create(void);
void oracle_destroy(oracle_sim* sim);

/* reference-only (the port returns -1): fluid::init_scene_from_json (init.cpp:158-418)
 * and fluid::init_test_scene (init.cpp:119-156). */
int oracle_load_scene(oracle_sim* sim, const char* json_path, char* err, size_t errlen);
int oracle_init_test_scene(oracle_sim* sim);

void oracle_set_params(oracle_sim* sim, const pbf_params* p);
void oracle_get_params(const oracle_sim* sim, pbf_params* p);
void oracle_set_planes(oracle_sim* sim, int count, const float* nx, const float* ny,
                       const float* nz, const float* d);
int oracle_plane_count(const oracle_sim* sim);
void oracle_get_planes(const oracle_sim* sim, float* nx, float* ny, float* nz, float* d);

void oracle_set_state(oracle_sim* sim, size_t n, const float* px, const float* py,
                      const float* pz, const float* vx, const float* vy, const float* vz);
size_t oracle_count(const oracle_sim* sim);
void oracle_get_state(const oracle_sim* sim, float* px, float* py, float* pz, float* vx,
                      float* vy, float* vz);
float oracle_time(const oracle_sim* sim);
void oracle_set_time(oracle_sim* sim, float t);

/* nsteps calls of fluid::step (core.cpp:119). */
void oracle_step(oracle_sim* sim, int nsteps);

/* State::cpu after the last step (core.h:92-115). */
size_t oracle_ncells(const oracle_sim* sim);
size_t oracle_nneighbors(const oracle_sim* sim);
void oracle_get_grid(const oracle_sim* sim, int32_t* entry_cx, int32_t* entry_cy,
                     int32_t* entry_cz, int32_t* entry_particle, int32_t* cell_xyz,
                     int32_t* cell_start, int32_t* cell_end);
void oracle_get_neighbors(const oracle_sim* sim, int32_t* prefix_sum, int32_t* indices);
void oracle_get_scratch(const oracle_sim* sim, int scratch_id, float* out);

#ifdef __cplusplus
}
#endif
#endif

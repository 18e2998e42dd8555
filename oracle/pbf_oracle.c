/*
 * pbf_oracle.c — plain-C restatement of the reference CPU PBF substep.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h): the checker for the CUDA path
 * and the "port" CPU baseline.  It restates fluid::step, reference
 * core/src/core.cpp:119-615, operation for operation in IEEE binary32.  Build it
 * WITHOUT FMA contraction and WITHOUT -march/-ffast-math (oracle/Makefile uses
 * -O2 -ffp-contract=off): the stock reference build is x86-64 baseline (no FMA),
 * and every float result below depends on that.
 *
 * Parity pin: tests/test_oracle.py requires this file to be bit-identical to the
 * unmodified reference (oracle/_ref, kind "reference") on state, grid tables,
 * neighbour lists and every scratch array, and both to match the committed
 * fixtures under tests/golden/ (generated from the reference by
 * tests/golden/make_golden.py).  The reference ships no golden vectors of its own
 * (SURVEY.md §4), so executing its source is the only pin there is.
 *
 * Differences from the reference that cannot change any result:
 *   - the neighbour list is built count-then-fill so the fill parallelises; the
 *     per-particle order (27 cells, dz/dy/dx with dx innermost, ascending id inside
 *     a cell) is the reference's (core.cpp:211-241);
 *   - the brute-force O(N^2) branch (core.cpp:248-268) is not restated: no scene or
 *     CLI flag reaches it (use_uniform_grid defaults to true, core.h:27).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_api.h"

#define ORACLE_PI 3.14159265358979323846f /* core.cpp:10 */

typedef struct {
  int x, y, z;
} cell_t;

typedef struct {
  cell_t key;
  int particle;
} entry_t; /* CpuScratch::CellEntry, core.h:87-90 */

struct oracle_sim {
  pbf_params p;
  int nplanes;
  float *pl_nx, *pl_ny, *pl_nz, *pl_d;
  size_t n;
  float* st[6]; /* pos x,y,z, vel x,y,z */
  float time;
  /* scratch (core.h:92-115) */
  float* scratch[PBF_SCRATCH_COUNT];
  entry_t* entries;
  cell_t* cell_key;
  int *cell_start, *cell_end;
  size_t ncells;
  int* nbr_prefix; /* inclusive */
  int* nbr_index;
  size_t nbr_cap, nbr_count;
};

/* ---- small helpers ------------------------------------------------------- */

static int cell_less(const cell_t* a, const cell_t* b) { /* core.cpp:12-21 */
  if (a->x != b->x) return a->x < b->x;
  if (a->y != b->y) return a->y < b->y;
  return a->z < b->z;
}

static int cell_eq(const cell_t* a, const cell_t* b) { /* core.cpp:23-26 */
  return a->x == b->x && a->y == b->y && a->z == b->z;
}

static int entry_cmp(const void* va, const void* vb) { /* core.cpp:174-183 */
  const entry_t* a = (const entry_t*)va;
  const entry_t* b = (const entry_t*)vb;
  if (cell_less(&a->key, &b->key)) return -1;
  if (cell_less(&b->key, &a->key)) return 1;
  return (a->particle > b->particle) - (a->particle < b->particle);
}

static cell_t cell_of(float x, float y, float z, float cell) { /* core.cpp:28-34 */
  const float inv = 1.0f / cell;
  cell_t c;
  c.x = (int)floorf(x * inv);
  c.y = (int)floorf(y * inv);
  c.z = (int)floorf(z * inv);
  return c;
}

static float poly6(float r2, float h) { /* core.cpp:35-46 */
  const float h2 = h * h;
  if (r2 > h2) return 0.0f;
  const float diff = h2 - r2;
  const float diff3 = diff * diff * diff;
  const float h4 = h2 * h2;
  const float h9 = h4 * h4 * h;
  const float coeff = 315.0f / (64.0f * ORACLE_PI * h9);
  return coeff * diff3;
}

static float spiky_factor(float r, float h) { /* core.cpp:48-57 */
  if (r > h) return 0.0f;
  const float h2 = h * h;
  const float h6 = h2 * h2 * h2;
  const float diff = h - r;
  const float coeff = -45.0f / (ORACLE_PI * h6);
  return coeff * diff * diff;
}

static float pow_ratio(float ratio, int n) { /* core.cpp:59-71 */
  if (n == 2) return ratio * ratio;
  if (n == 3) return ratio * ratio * ratio;
  if (n == 4) {
    const float r2 = ratio * ratio;
    return r2 * r2;
  }
  return powf(ratio, (float)n);
}

/* index of the occupied cell `key`, or -1 (core.cpp:216-226) */
static long find_cell(const oracle_sim* s, const cell_t* key) {
  size_t lo = 0, hi = s->ncells;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (cell_less(&s->cell_key[mid], key))
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo == s->ncells || !cell_eq(&s->cell_key[lo], key)) return -1;
  return (long)lo;
}

static void resize_particles(oracle_sim* s, size_t n) {
  for (int a = 0; a < 6; ++a) s->st[a] = (float*)realloc(s->st[a], (n ? n : 1) * sizeof(float));
  for (int a = 0; a < PBF_SCRATCH_COUNT; ++a) {
    s->scratch[a] = (float*)realloc(s->scratch[a], (n ? n : 1) * sizeof(float));
    memset(s->scratch[a], 0, n * sizeof(float));
  }
  s->entries = (entry_t*)realloc(s->entries, (n ? n : 1) * sizeof(entry_t));
  s->cell_key = (cell_t*)realloc(s->cell_key, (n ? n : 1) * sizeof(cell_t));
  s->cell_start = (int*)realloc(s->cell_start, (n ? n : 1) * sizeof(int));
  s->cell_end = (int*)realloc(s->cell_end, (n ? n : 1) * sizeof(int));
  s->nbr_prefix = (int*)realloc(s->nbr_prefix, (n ? n : 1) * sizeof(int));
  memset(s->nbr_prefix, 0, n * sizeof(int));
  s->n = n;
  s->ncells = 0;
  s->nbr_count = 0;
}

/* ---- one substep: core.cpp:119-615 ---------------------------------------- */

/* visits the candidates of particle i in the reference order and either counts or
 * stores the ones with r2 < h2 (core.cpp:205-247) */
static int neighbours_of(const oracle_sim* s, size_t i, float h, float h2, int* out) {
  const float* px = s->scratch[PBF_SCRATCH_PRED_X];
  const float* py = s->scratch[PBF_SCRATCH_PRED_Y];
  const float* pz = s->scratch[PBF_SCRATCH_PRED_Z];
  const float xi = px[i], yi = py[i], zi = pz[i];
  const cell_t base = cell_of(xi, yi, zi, h);
  int count = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        cell_t key;
        key.x = base.x + dx;
        key.y = base.y + dy;
        key.z = base.z + dz;
        const long c = find_cell(s, &key);
        if (c < 0) continue;
        for (int idx = s->cell_start[c]; idx < s->cell_end[c]; ++idx) {
          const int j = s->entries[idx].particle;
          if (j == (int)i) continue;
          const float ddx = xi - px[j];
          const float ddy = yi - py[j];
          const float ddz = zi - pz[j];
          const float r2 = ddx * ddx + ddy * ddy + ddz * ddz;
          if (r2 < h2) {
            if (out) out[count] = j;
            ++count;
          }
        }
      }
  return count;
}

static void substep(oracle_sim* s) {
  const pbf_params* P = &s->p;
  const size_t n = s->n;
  if (n == 0) { /* core.cpp:122-125 */
    s->time += P->dt;
    return;
  }
  float *posx = s->st[0], *posy = s->st[1], *posz = s->st[2];
  float *velx = s->st[3], *vely = s->st[4], *velz = s->st[5];
  float* predx = s->scratch[PBF_SCRATCH_PRED_X];
  float* predy = s->scratch[PBF_SCRATCH_PRED_Y];
  float* predz = s->scratch[PBF_SCRATCH_PRED_Z];
  float* lambda = s->scratch[PBF_SCRATCH_LAMBDA];
  float* rho_a = s->scratch[PBF_SCRATCH_RHO];

  /* per-step constants, core.cpp:138-148 */
  const float dt = P->dt;
  const float h = P->h;
  const float h2 = h * h;
  const float min_r = 0.01f * h;
  const float min_r2 = min_r * min_r;
  const int scorr_enabled = P->enable_scorr && P->scorr_k != 0.0f;
  const float dq_coeff = (P->scorr_dq_coeff > 0.0f) ? P->scorr_dq_coeff : 0.3f;
  const float scorr_dq = dq_coeff * h;
  const float wdq = scorr_enabled ? poly6(scorr_dq * scorr_dq, h) : 0.0f;
  const float scorr_inv_wdq = (wdq > 1e-12f) ? (1.0f / wdq) : 0.0f;

  /* a3 predict, core.cpp:150-161 */
  {
    const float gx = P->external_force[0], gy = P->external_force[1], gz = P->external_force[2];
    long i;
#pragma omp parallel for
    for (i = 0; i < (long)n; ++i) {
      velx[i] += gx * dt;
      vely[i] += gy * dt;
      velz[i] += gz * dt;
      predx[i] = posx[i] + velx[i] * dt;
      predy[i] = posy[i] + vely[i] * dt;
      predz[i] = posz[i] + velz[i] * dt;
    }
  }

  /* a4/a5 cell keys + sort, core.cpp:164-183 */
  for (size_t i = 0; i < n; ++i) {
    s->entries[i].key = cell_of(predx[i], predy[i], predz[i], h);
    s->entries[i].particle = (int)i;
  }
  qsort(s->entries, n, sizeof(entry_t), entry_cmp);

  /* a6 cell table, core.cpp:185-203 */
  s->ncells = 0;
  {
    cell_t current = s->entries[0].key;
    int start = 0;
    for (size_t i = 1; i < n; ++i) {
      if (!cell_eq(&current, &s->entries[i].key)) {
        s->cell_key[s->ncells] = current;
        s->cell_start[s->ncells] = start;
        s->cell_end[s->ncells] = (int)i;
        s->ncells++;
        current = s->entries[i].key;
        start = (int)i;
      }
    }
    s->cell_key[s->ncells] = current;
    s->cell_start[s->ncells] = start;
    s->cell_end[s->ncells] = (int)n;
    s->ncells++;
  }

  /* a7 neighbour list, core.cpp:205-247 (count, inclusive scan, fill) */
  {
    long i;
#pragma omp parallel for schedule(dynamic, 256)
    for (i = 0; i < (long)n; ++i) s->nbr_prefix[i] = neighbours_of(s, (size_t)i, h, h2, NULL);
    size_t total = 0;
    for (size_t k = 0; k < n; ++k) {
      total += (size_t)s->nbr_prefix[k];
      s->nbr_prefix[k] = (int)total;
    }
    if (total > s->nbr_cap) {
      s->nbr_cap = total + total / 2 + 16;
      s->nbr_index = (int*)realloc(s->nbr_index, s->nbr_cap * sizeof(int));
    }
    s->nbr_count = total;
#pragma omp parallel for schedule(dynamic, 256)
    for (i = 0; i < (long)n; ++i) {
      const int begin = (i == 0) ? 0 : s->nbr_prefix[i - 1];
      neighbours_of(s, (size_t)i, h, h2, s->nbr_index + begin);
    }
  }

  /* core.cpp:270-274 */
  const float density = P->density;
  const float inv_density = 1.0f / density;
  const float mass = P->particle_mass;
  const float grad_scale = mass * inv_density;
  const float epsilon = P->epsilon;
  const int* nidx = s->nbr_index;
  const int* npre = s->nbr_prefix;

  for (int iter = 0; iter < P->solver_iterations; ++iter) { /* core.cpp:277 */
    long i;
    /* a8 lambda, core.cpp:281-329 */
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) {
      const int begin = (i == 0) ? 0 : npre[i - 1];
      const int end = npre[i];
      const float xi = predx[i], yi = predy[i], zi = predz[i];
      float rho = 0.0f, gsx = 0.0f, gsy = 0.0f, gsz = 0.0f, sum_grad2 = 0.0f;
      for (int k = begin; k < end; ++k) {
        const int j = nidx[k];
        const float dx = xi - predx[j];
        const float dy = yi - predy[j];
        const float dz = zi - predz[j];
        const float r2 = dx * dx + dy * dy + dz * dz;
        rho += poly6(r2, h);
        if (r2 < h2) {
          const float r = sqrtf(r2 < min_r2 ? min_r2 : r2);
          const float gf = spiky_factor(r, h);
          const float gx = gf * dx, gy = gf * dy, gz = gf * dz;
          gsx += gx;
          gsy += gy;
          gsz += gz;
          const float jx = -grad_scale * gx, jy = -grad_scale * gy, jz = -grad_scale * gz;
          sum_grad2 += jx * jx + jy * jy + jz * jz;
        }
      }
      rho += poly6(0.0f, h);
      rho *= mass;
      rho_a[i] = rho;
      const float C = rho * inv_density - 1.0f;
      const float ix = grad_scale * gsx, iy = grad_scale * gsy, iz = grad_scale * gsz;
      sum_grad2 += ix * ix + iy * iy + iz * iz;
      lambda[i] = -C / (sum_grad2 + epsilon);
    }

    /* a9 delta + s_corr + plane projection, core.cpp:334-398 */
    float* dlx = s->scratch[PBF_SCRATCH_DELTA_X];
    float* dly = s->scratch[PBF_SCRATCH_DELTA_Y];
    float* dlz = s->scratch[PBF_SCRATCH_DELTA_Z];
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) {
      const int begin = (i == 0) ? 0 : npre[i - 1];
      const int end = npre[i];
      const float xi = predx[i], yi = predy[i], zi = predz[i];
      const float li = lambda[i];
      float ax = 0.0f, ay = 0.0f, az = 0.0f;
      for (int k = begin; k < end; ++k) {
        const int j = nidx[k];
        const float dx = xi - predx[j];
        const float dy = yi - predy[j];
        const float dz = zi - predz[j];
        const float r2 = dx * dx + dy * dy + dz * dz;
        if (r2 < h2) {
          const float r = sqrtf(r2 < min_r2 ? min_r2 : r2);
          const float gf = spiky_factor(r, h);
          float sc = li + lambda[j];
          if (scorr_enabled && scorr_inv_wdq > 0.0f) {
            const float W = poly6(r2, h);
            const float ratio = W * scorr_inv_wdq;
            const float corr = -P->scorr_k * pow_ratio(ratio, P->scorr_n);
            sc += corr;
          }
          ax += sc * gf * dx;
          ay += sc * gf * dy;
          az += sc * gf * dz;
        }
      }
      ax *= inv_density;
      ay *= inv_density;
      az *= inv_density;
      if (s->nplanes > 0) { /* core.cpp:372-393 */
        float qx = xi + ax, qy = yi + ay, qz = zi + az;
        for (int p = 0; p < s->nplanes; ++p) {
          const float nx = s->pl_nx[p], ny = s->pl_ny[p], nz = s->pl_nz[p], d = s->pl_d[p];
          const float sd = nx * qx + ny * qy + nz * qz - d;
          const float pen = -sd;
          if (pen > 0.0f) {
            qx += nx * pen;
            qy += ny * pen;
            qz += nz * pen;
          }
        }
        ax = qx - xi;
        ay = qy - yi;
        az = qz - zi;
      }
      dlx[i] = ax;
      dly[i] = ay;
      dlz[i] = az;
    }

    /* a10 apply, core.cpp:400-407 */
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) {
      predx[i] += dlx[i];
      predy[i] += dly[i];
      predz[i] += dlz[i];
    }
  }

  /* a11 velocity update + commit, core.cpp:410-421 */
  {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) {
      velx[i] = (predx[i] - posx[i]) / dt;
      vely[i] = (predy[i] - posy[i]) / dt;
      velz[i] = (predz[i] - posz[i]) / dt;
      posx[i] = predx[i];
      posy[i] = predy[i];
      posz[i] = predz[i];
    }
  }

  /* a12 XSPH, core.cpp:423-466 */
  if (P->enable_xsph && P->visc_c != 0.0f) {
    float* dvx_a = s->scratch[PBF_SCRATCH_DV_X];
    float* dvy_a = s->scratch[PBF_SCRATCH_DV_Y];
    float* dvz_a = s->scratch[PBF_SCRATCH_DV_Z];
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) {
      const int begin = (i == 0) ? 0 : npre[i - 1];
      const int end = npre[i];
      const float vx = velx[i], vy = vely[i], vz = velz[i];
      float sx = 0.0f, sy = 0.0f, sz = 0.0f;
      for (int k = begin; k < end; ++k) {
        const int j = nidx[k];
        const float dx = posx[i] - posx[j];
        const float dy = posy[i] - posy[j];
        const float dz = posz[i] - posz[j];
        const float r2 = dx * dx + dy * dy + dz * dz;
        if (r2 < h2) {
          const float W = poly6(r2, h);
          const float inv_rho_j = (rho_a[j] > 0.0f) ? (mass / rho_a[j]) : 0.0f;
          sx += (velx[j] - vx) * W * inv_rho_j;
          sy += (vely[j] - vy) * W * inv_rho_j;
          sz += (velz[j] - vz) * W * inv_rho_j;
        }
      }
      dvx_a[i] = sx;
      dvy_a[i] = sy;
      dvz_a[i] = sz;
    }
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) {
      velx[i] += P->visc_c * dvx_a[i];
      vely[i] += P->visc_c * dvy_a[i];
      velz[i] += P->visc_c * dvz_a[i];
    }
  }

  /* a13 vorticity confinement, core.cpp:468-571 */
  if (P->enable_vorticity && P->vort_epsilon != 0.0f) {
    float* omx = s->scratch[PBF_SCRATCH_OMEGA_X];
    float* omy = s->scratch[PBF_SCRATCH_OMEGA_Y];
    float* omz = s->scratch[PBF_SCRATCH_OMEGA_Z];
    float* omm = s->scratch[PBF_SCRATCH_OMEGA_MAG];
    float* etx = s->scratch[PBF_SCRATCH_ETA_X];
    float* ety = s->scratch[PBF_SCRATCH_ETA_Y];
    float* etz = s->scratch[PBF_SCRATCH_ETA_Z];
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) { /* omega, core.cpp:472-508 */
      const int begin = (i == 0) ? 0 : npre[i - 1];
      const int end = npre[i];
      const float xi = posx[i], yi = posy[i], zi = posz[i];
      const float vx = velx[i], vy = vely[i], vz = velz[i];
      float ox = 0.0f, oy = 0.0f, oz = 0.0f;
      for (int k = begin; k < end; ++k) {
        const int j = nidx[k];
        const float dx = xi - posx[j];
        const float dy = yi - posy[j];
        const float dz = zi - posz[j];
        const float r2 = dx * dx + dy * dy + dz * dz;
        if (r2 < h2) {
          const float r = sqrtf(r2 < min_r2 ? min_r2 : r2);
          const float gf = spiky_factor(r, h);
          const float gx = gf * dx, gy = gf * dy, gz = gf * dz;
          const float ux = velx[j] - vx, uy = vely[j] - vy, uz = velz[j] - vz;
          ox += uy * gz - uz * gy;
          oy += uz * gx - ux * gz;
          oz += ux * gy - uy * gx;
        }
      }
      omx[i] = ox;
      omy[i] = oy;
      omz[i] = oz;
      omm[i] = sqrtf(ox * ox + oy * oy + oz * oz);
    }
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) { /* eta, core.cpp:512-543 */
      const int begin = (i == 0) ? 0 : npre[i - 1];
      const int end = npre[i];
      const float xi = posx[i], yi = posy[i], zi = posz[i];
      const float omi = omm[i];
      float ex = 0.0f, ey = 0.0f, ez = 0.0f;
      for (int k = begin; k < end; ++k) {
        const int j = nidx[k];
        const float dx = xi - posx[j];
        const float dy = yi - posy[j];
        const float dz = zi - posz[j];
        const float r2 = dx * dx + dy * dy + dz * dz;
        if (r2 < h2) {
          const float r = sqrtf(r2 < min_r2 ? min_r2 : r2);
          const float gf = spiky_factor(r, h);
          const float gx = gf * dx, gy = gf * dy, gz = gf * dz;
          const float coeff = omm[j] - omi;
          ex += coeff * gx;
          ey += coeff * gy;
          ez += coeff * gz;
        }
      }
      etx[i] = ex;
      ety[i] = ey;
      etz[i] = ez;
    }
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) { /* apply, core.cpp:547-570 */
      const float ex = etx[i], ey = ety[i], ez = etz[i];
      const float len = sqrtf(ex * ex + ey * ey + ez * ez);
      float nx = 0.0f, ny = 0.0f, nz = 0.0f;
      if (len > P->vort_norm_eps) {
        const float inv = 1.0f / len;
        nx = ex * inv;
        ny = ey * inv;
        nz = ez * inv;
      }
      const float ox = omx[i], oy = omy[i], oz = omz[i];
      const float fx = P->vort_epsilon * (ny * oz - nz * oy);
      const float fy = P->vort_epsilon * (nz * ox - nx * oz);
      const float fz = P->vort_epsilon * (nx * oy - ny * ox);
      velx[i] += dt * fx;
      vely[i] += dt * fy;
      velz[i] += dt * fz;
    }
  }

  /* a14 plane restitution / friction, core.cpp:573-612 */
  if ((P->plane_restitution > 0.0f || P->plane_friction > 0.0f) && s->nplanes > 0) {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; ++i) {
      float vx = velx[i], vy = vely[i], vz = velz[i];
      const float xi = posx[i], yi = posy[i], zi = posz[i];
      for (int p = 0; p < s->nplanes; ++p) {
        const float nx = s->pl_nx[p], ny = s->pl_ny[p], nz = s->pl_nz[p], d = s->pl_d[p];
        const float sd = nx * xi + ny * yi + nz * zi - d;
        if (sd <= 0.0f) {
          const float vn = nx * vx + ny * vy + nz * vz;
          float vn_new = vn;
          if (vn < 0.0f) vn_new = -P->plane_restitution * vn;
          const float tx = vx - vn * nx, ty = vy - vn * ny, tz = vz - vn * nz;
          const float scale = 1.0f - P->plane_friction;
          vx = tx * scale + vn_new * nx;
          vy = ty * scale + vn_new * ny;
          vz = tz * scale + vn_new * nz;
        }
      }
      velx[i] = vx;
      vely[i] = vy;
      velz[i] = vz;
    }
  }

  s->time += dt; /* a15, core.cpp:614 */
}

/* ---- exported API --------------------------------------------------------- */

const char* oracle_kind(void) { return "port"; }

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
  oracle_sim* s = (oracle_sim*)calloc(1, sizeof(oracle_sim));
  pbf_params* p = &s->p; /* defaults of fluid::Params, core.h:12-43 */
  p->dt = 1.0f / 60.0f;
  p->density = 6000.0f;
  p->particle_mass = 0.0f;
  p->h = 0.0f;
  p->particle_radius = 0.01f;
  p->epsilon = 600.0f;
  p->solver_iterations = 4;
  p->neighbor_reserve_factor = 1.5f;
  p->use_uniform_grid = 1;
  p->scorr_k = 0.00005f;
  p->scorr_n = 4;
  p->scorr_dq_coeff = 0.3f;
  p->visc_c = 0.0002f;
  p->vort_epsilon = 0.5f;
  p->vort_norm_eps = 1e-6f;
  p->external_force[1] = -9.8f;
  resize_particles(s, 0);
  return s;
}

void oracle_destroy(oracle_sim* s) {
  if (!s) return;
  for (int a = 0; a < 6; ++a) free(s->st[a]);
  for (int a = 0; a < PBF_SCRATCH_COUNT; ++a) free(s->scratch[a]);
  free(s->entries);
  free(s->cell_key);
  free(s->cell_start);
  free(s->cell_end);
  free(s->nbr_prefix);
  free(s->nbr_index);
  free(s->pl_nx);
  free(s->pl_ny);
  free(s->pl_nz);
  free(s->pl_d);
  free(s);
}

int oracle_load_scene(oracle_sim* s, const char* path, char* err, size_t errlen) {
  (void)s;
  (void)path;
  if (err && errlen) strncpy(err, "the port oracle has no scene loader", errlen - 1), err[errlen - 1] = 0;
  return -1;
}

int oracle_init_test_scene(oracle_sim* s) {
  (void)s;
  return -1;
}

void oracle_set_params(oracle_sim* s, const pbf_params* p) { s->p = *p; }
void oracle_get_params(const oracle_sim* s, pbf_params* p) { *p = s->p; }

void oracle_set_planes(oracle_sim* s, int count, const float* nx, const float* ny,
                       const float* nz, const float* d) {
  const size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(float);
  s->pl_nx = (float*)realloc(s->pl_nx, bytes);
  s->pl_ny = (float*)realloc(s->pl_ny, bytes);
  s->pl_nz = (float*)realloc(s->pl_nz, bytes);
  s->pl_d = (float*)realloc(s->pl_d, bytes);
  for (int i = 0; i < count; ++i) {
    s->pl_nx[i] = nx[i];
    s->pl_ny[i] = ny[i];
    s->pl_nz[i] = nz[i];
    s->pl_d[i] = d[i];
  }
  s->nplanes = count;
}

int oracle_plane_count(const oracle_sim* s) { return s->nplanes; }

void oracle_get_planes(const oracle_sim* s, float* nx, float* ny, float* nz, float* d) {
  for (int i = 0; i < s->nplanes; ++i) {
    nx[i] = s->pl_nx[i];
    ny[i] = s->pl_ny[i];
    nz[i] = s->pl_nz[i];
    d[i] = s->pl_d[i];
  }
}

void oracle_set_state(oracle_sim* s, size_t n, const float* px, const float* py,
                      const float* pz, const float* vx, const float* vy, const float* vz) {
  const float* src[6] = {px, py, pz, vx, vy, vz};
  resize_particles(s, n);
  for (int a = 0; a < 6; ++a) memcpy(s->st[a], src[a], n * sizeof(float));
}

size_t oracle_count(const oracle_sim* s) { return s->n; }

void oracle_get_state(const oracle_sim* s, float* px, float* py, float* pz, float* vx,
                      float* vy, float* vz) {
  float* dst[6] = {px, py, pz, vx, vy, vz};
  for (int a = 0; a < 6; ++a)
    if (dst[a]) memcpy(dst[a], s->st[a], s->n * sizeof(float));
}

float oracle_time(const oracle_sim* s) { return s->time; }
void oracle_set_time(oracle_sim* s, float t) { s->time = t; }

void oracle_step(oracle_sim* s, int nsteps) {
  for (int k = 0; k < nsteps; ++k) substep(s);
}

size_t oracle_ncells(const oracle_sim* s) { return s->ncells; }
size_t oracle_nneighbors(const oracle_sim* s) { return s->nbr_count; }

void oracle_get_grid(const oracle_sim* s, int32_t* ecx, int32_t* ecy, int32_t* ecz,
                     int32_t* eparticle, int32_t* cell_xyz, int32_t* cstart, int32_t* cend) {
  for (size_t i = 0; i < s->n; ++i) {
    if (ecx) ecx[i] = s->entries[i].key.x;
    if (ecy) ecy[i] = s->entries[i].key.y;
    if (ecz) ecz[i] = s->entries[i].key.z;
    if (eparticle) eparticle[i] = s->entries[i].particle;
  }
  for (size_t k = 0; k < s->ncells; ++k) {
    if (cell_xyz) {
      cell_xyz[3 * k + 0] = s->cell_key[k].x;
      cell_xyz[3 * k + 1] = s->cell_key[k].y;
      cell_xyz[3 * k + 2] = s->cell_key[k].z;
    }
    if (cstart) cstart[k] = s->cell_start[k];
    if (cend) cend[k] = s->cell_end[k];
  }
}

void oracle_get_neighbors(const oracle_sim* s, int32_t* prefix_sum, int32_t* indices) {
  if (prefix_sum) memcpy(prefix_sum, s->nbr_prefix, s->n * sizeof(int));
  if (indices && s->nbr_count) memcpy(indices, s->nbr_index, s->nbr_count * sizeof(int));
}

void oracle_get_scratch(const oracle_sim* s, int id, float* out) {
  if (id < 0 || id >= PBF_SCRATCH_COUNT || !out) return;
  memcpy(out, s->scratch[id], s->n * sizeof(float));
}

// pbf_capi.cu — the C ABI of include/pbf_b200.h: context management, the device-resident
// substep driver (stream launches or a replayed CUDA graph), overflow-safe batching, and the
// parity / measurement surfaces.  No CPU fallback exists anywhere in this file: every entry
// point either runs the CUDA path or fails with an error string.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pbf_context.h"

using namespace pbf;

namespace {
thread_local std::string g_error;  // context-less failures (pbf_create, pbf_device_count)
}

namespace pbf {
int fail(pbf_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->error = msg; else g_error = msg;
  return code;
}
}  // namespace pbf

namespace {

constexpr float kPi = 3.14159265358979323846f;  // core.cpp:10

// poly6_kernel on the host, same float expression order as core.cpp:35-46
float host_poly6(float r2, float h) {
  const float h2 = h * h;
  if (r2 > h2) return 0.0f;
  const float diff = h2 - r2;
  const float diff3 = diff * diff * diff;
  const float h4 = h2 * h2;
  const float h9 = h4 * h4 * h;
  const float coeff = 315.0f / (64.0f * kPi * h9);
  return coeff * diff3;
}

// Per-step constants in the oracle's float expression order (core.cpp:138-148, 270-274,
// and the coefficient expressions inside poly6_kernel / spiky_gradient_factor).
StepConsts make_consts(const pbf_params& p, int nplanes) {
  StepConsts c{};
  c.dt = p.dt;
  c.inv_dt = 1.0f / p.dt;
  const float h = p.h;
  c.h = h;
  c.h2 = h * h;
  c.inv_h = 1.0f / h;
  const float min_r = 0.01f * h;
  c.min_r2 = min_r * min_r;
  const float h2 = h * h;
  const float h4 = h2 * h2;
  const float h9 = h4 * h4 * h;
  c.poly6_coeff = 315.0f / (64.0f * kPi * h9);
  c.poly6_zero = host_poly6(0.0f, h);
  const float h6 = h2 * h2 * h2;
  c.spiky_coeff = -45.0f / (kPi * h6);
  c.inv_density = 1.0f / p.density;
  c.mass = p.particle_mass;
  c.grad_scale = p.particle_mass * c.inv_density;
  c.epsilon = p.epsilon;
  const bool scorr_enabled = p.enable_scorr && p.scorr_k != 0.0f;
  const float dq_coeff = (p.scorr_dq_coeff > 0.0f) ? p.scorr_dq_coeff : 0.3f;
  const float scorr_dq = dq_coeff * h;
  const float wdq = scorr_enabled ? host_poly6(scorr_dq * scorr_dq, h) : 0.0f;
  c.scorr_inv_wdq = (wdq > 1e-12f) ? (1.0f / wdq) : 0.0f;
  c.scorr_on = (scorr_enabled && c.scorr_inv_wdq > 0.0f) ? 1 : 0;
  c.scorr_negk = -p.scorr_k;
  c.scorr_n = p.scorr_n;
  c.visc_c = p.visc_c;
  c.vort_eps = p.vort_epsilon;
  c.vort_norm_eps = p.vort_norm_eps;
  c.restitution = p.plane_restitution;
  c.one_minus_friction = 1.0f - p.plane_friction;
  c.gdt_x = p.external_force[0] * p.dt;
  c.gdt_y = p.external_force[1] * p.dt;
  c.gdt_z = p.external_force[2] * p.dt;
  c.nplanes = nplanes;
  c.do_xsph = (p.enable_xsph && p.visc_c != 0.0f) ? 1 : 0;
  c.do_vort = (p.enable_vorticity && p.vort_epsilon != 0.0f) ? 1 : 0;
  c.do_rest = ((p.plane_restitution > 0.0f || p.plane_friction > 0.0f) && nplanes > 0) ? 1 : 0;
  c.sqrt_safe = (c.min_r2 >= 1e-30f && c.h2 <= 1e30f) ? 1 : 0;
  return c;
}

}  // namespace

namespace pbf {

void invalidate_graph(pbf_ctx* ctx) {
  if (ctx->graph_exec) {
    cudaGraphExecDestroy(ctx->graph_exec);
    ctx->graph_exec = nullptr;
  }
  if (ctx->host_graph) {
    cudaGraphExecDestroy(ctx->host_graph);
    ctx->host_graph = nullptr;
  }
}

// Per-particle buffers.  Arrays in ORIGINAL order hold `cap` owned particles; arrays in SORTED order
// additionally hold the ghosts of a slab (tot = cap + 2 * gcap; gcap == 0 on a single GPU).
// `keep` > 0 preserves that many elements of the persistent state across the growth.
int ensure_particles(pbf_ctx* ctx, size_t n, size_t keep) {
  const size_t want_tot = std::max<size_t>(n, 1024) + 2 * (size_t)ctx->slab.gcap;
  if (n <= ctx->cap && want_tot <= ctx->pred_a.n) return PBF_OK;
  const size_t cap = std::max<size_t>(std::max<size_t>(n, 1024), ctx->cap);
  const size_t tot = cap + 2 * (size_t)ctx->slab.gcap;
  invalidate_graph(ctx);
  PBF_CUDA(ctx, ctx->pos_o.grow_keep(cap, keep));
  PBF_CUDA(ctx, ctx->vel_o.grow_keep(cap, keep));
  PBF_CUDA(ctx, ctx->pos_bak.grow_keep(cap, keep));
  PBF_CUDA(ctx, ctx->vel_bak.grow_keep(cap, keep));
  PBF_CUDA(ctx, ctx->pred_o.reserve(cap));
  // + 2: the neighbour kernel may load (and ignore) the slot after the last particle, which is
  // padding or a slot no substep has written yet — zero-filled once so that the load is defined
  PBF_CUDA(ctx, ctx->pred_a.reserve(tot + 2));
  PBF_CUDA(ctx, ctx->pred_b.reserve(tot + 2));
  PBF_CUDA(ctx, cudaMemsetAsync(ctx->pred_a.p, 0, ctx->pred_a.n * sizeof(float4), ctx->stream));
  PBF_CUDA(ctx, cudaMemsetAsync(ctx->pred_b.p, 0, ctx->pred_b.n * sizeof(float4), ctx->stream));
  PBF_CUDA(ctx, ctx->pos_s.reserve(tot));
  PBF_CUDA(ctx, ctx->vel_a.reserve(tot));
  PBF_CUDA(ctx, ctx->vel_b.reserve(tot));
  PBF_CUDA(ctx, ctx->pv.reserve(tot));
  PBF_CUDA(ctx, ctx->omega.reserve(tot));
  PBF_CUDA(ctx, ctx->rho.reserve(tot));
  PBF_CUDA(ctx, ctx->keys0.reserve(cap));
  PBF_CUDA(ctx, ctx->keys1.reserve(cap));
  PBF_CUDA(ctx, ctx->vals0.reserve(cap));
  PBF_CUDA(ctx, ctx->vals1.reserve(cap));
  PBF_CUDA(ctx, ctx->nbr_count.reserve(tot));
  for (auto& b : ctx->soa) PBF_CUDA(ctx, b.reserve(cap));
  if (ctx->slab.enabled) {
    PBF_CUDA(ctx, ctx->slab.gid_o.grow_keep(cap, keep));
    PBF_CUDA(ctx, ctx->slab.gid_bak.grow_keep(cap, keep));
  }
  ctx->cap = cap;
  ctx->slab.tot_cap = ctx->slab.enabled ? tot : 0;
  return PBF_OK;
}

int ensure_tables(pbf_ctx* ctx) {
  if (ctx->cell_range.n < ctx->cell_cap) {
    invalidate_graph(ctx);
    PBF_CUDA(ctx, ctx->cell_range.reserve(ctx->cell_cap));
    PBF_CUDA(ctx, ctx->cell_count.reserve(ctx->cell_cap));
    PBF_CUDA(ctx, ctx->cell_excl.reserve(ctx->cell_cap));
    PBF_CUDA(ctx, ctx->cell_key.reserve(ctx->cell_cap));
    ctx->tables_dirty = true;
    PBF_CUDA(ctx, ctx->chunk_total.reserve(ctx->cell_cap / 2048 + 2));  // one chunk total per 2048 table cells
  }
  if (ctx->slot_id.n < ctx->cap) {
    invalidate_graph(ctx);
    PBF_CUDA(ctx, ctx->slot_id.reserve(ctx->cap));
  }
  const size_t slots = ctx->slab.enabled ? ctx->slab.tot_cap : ctx->cap;
  const size_t need = ((slots + 31) / 32) * (size_t)ctx->K * 32u;
  if (ctx->nbr_idx.n < need) {
    invalidate_graph(ctx);
    // The list is an ELL table: K = the largest neighbour count of any particle.  In the reference's
    // own divergence (vorticity on, SURVEY §0: ~10 000 particles in one cell) K reaches five digits;
    // say so instead of failing inside cudaMalloc.
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const size_t have = ctx->nbr_idx.n * sizeof(uint32_t);
      if (need * sizeof(uint32_t) > free_b + have) {
        char buf[320];
        std::snprintf(buf, sizeof(buf),
                      "pbf_step: the neighbour table needs %.1f GB (K = %d neighbours for %zu slots) but only %.1f GB of "
                      "device memory are free; the scene has collapsed into a few cells (state restored to the start of the batch)",
                      (double)need * 4e-9, ctx->K, slots, (double)(free_b + have) * 1e-9);
        return fail(ctx, PBF_E_CAPACITY, buf);
      }
    }
    PBF_CUDA(ctx, ctx->nbr_idx.reserve(need));
    // rows are only filled up to each particle's count; the debug surface copies whole rows
    PBF_CUDA(ctx, cudaMemsetAsync(ctx->nbr_idx.p, 0, need * sizeof(uint32_t), ctx->stream));
  }
  if (ctx->brick_on) {
    if (ctx->brick_cap < 1024) ctx->brick_cap = 1024;
    // a first guess from the particle count: ~6 particles per cell at rest density, bounding box
    // of a splash several times the fluid volume; the table grows on demand (kBrickGrow)
    const size_t guess = slots / (size_t)(pbf::kBrickX * pbf::kBrickY * pbf::kBrickZ) + 1024;
    if ((size_t)ctx->brick_cap < guess) ctx->brick_cap = (int)guess;
    if (ctx->bricks.n < (size_t)ctx->brick_cap) {
      invalidate_graph(ctx);
      PBF_CUDA(ctx, ctx->bricks.reserve((size_t)ctx->brick_cap));
    }
    if (!ctx->brick_ctl.p) {
      invalidate_graph(ctx);
      PBF_CUDA(ctx, ctx->brick_ctl.reserve(2));
    }
  }
  if (ctx->debug) {
    PBF_CUDA(ctx, ctx->dbg_lambda.reserve(slots));
    PBF_CUDA(ctx, ctx->dbg_rho.reserve(slots));
    PBF_CUDA(ctx, ctx->dbg_delta.reserve(slots));
    PBF_CUDA(ctx, ctx->dbg_dv.reserve(slots));
    PBF_CUDA(ctx, ctx->dbg_eta.reserve(slots));
  }
  return PBF_OK;
}

// ---- stage profiling --------------------------------------------------------
cudaEvent_t timer_event(StageTimer& t) {
  if (!t.pool.empty()) {
    cudaEvent_t e = t.pool.back();
    t.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void stage_mark(void* user, int stage, int begin) {
  pbf_ctx* ctx = static_cast<pbf_ctx*>(user);
  if (!ctx->profile) return;
  StageTimer& t = ctx->timer;
  cudaEvent_t e = timer_event(t);
  cudaEventRecord(e, ctx->stream);
  if (begin) {
    t.begin.push_back(e);
    t.stage.push_back(stage);
  } else {
    t.end.push_back(e);
  }
}

void timer_resolve(pbf_ctx* ctx) {
  StageTimer& t = ctx->timer;
  for (size_t k = 0; k < t.end.size(); ++k) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, t.begin[k], t.end[k]) == cudaSuccess) t.total_ms[t.stage[k]] += ms;
    t.pool.push_back(t.begin[k]);
    t.pool.push_back(t.end[k]);
  }
  t.begin.clear();
  t.end.clear();
  t.stage.clear();
}

// ---- one substep ------------------------------------------------------------
// Enqueues the kernels of one substep.  Returns the kernel count.
void fill_grid_buffers(pbf_ctx* ctx, GridBuffers& g) {
  g.desc = ctx->desc.p;
  g.status = ctx->status.p;
  g.keys[0] = ctx->keys0.p;
  g.keys[1] = ctx->keys1.p;
  g.vals[0] = ctx->vals0.p;
  g.vals[1] = ctx->vals1.p;
  g.chunk_total = ctx->chunk_total.p;
  g.cell_range = ctx->cell_range.p;
  g.cell_count = ctx->cell_count.p;
  g.cell_excl = ctx->cell_excl.p;
  g.slot_id = ctx->slot_id.p;
  g.cell_key = ctx->cell_key.p;
  g.cell_cap = ctx->cell_cap;
  g.bricks = ctx->brick_on ? ctx->bricks.p : nullptr;
  g.brick_cap = ctx->brick_on ? ctx->brick_cap : 0;
}

void fill_neighbor_list(pbf_ctx* ctx, NeighborList& nl) {
  nl.idx = ctx->nbr_idx.p;
  nl.count = ctx->nbr_count.p;
  nl.K = ctx->K;
  nl.bricks = ctx->brick_on ? ctx->bricks.p : nullptr;
  nl.desc = ctx->desc.p;
  nl.brick_cap = ctx->brick_on ? ctx->brick_cap : 0;
  nl.brick_ctl = ctx->brick_on ? ctx->brick_ctl.p : nullptr;
  nl.brick_persist = ctx->brick_persist;
}

void fill_solve_buffers(pbf_ctx* ctx, SolveBuffers& b) {
  b.pred[0] = ctx->pred_a.p;
  b.pred[1] = ctx->pred_b.p;
  b.pos_s = ctx->pos_s.p;
  b.vel[0] = ctx->vel_a.p;
  b.vel[1] = ctx->vel_b.p;
  b.pv = ctx->pv.p;
  b.omega = ctx->omega.p;
  b.rho = ctx->rho.p;
  b.planes = ctx->planes_dev.p;
  b.pos_o = ctx->pos_o.p;
  b.vel_o = ctx->vel_o.p;
  b.status = ctx->status.p;
  if (ctx->debug) {
    b.dbg.lambda = ctx->dbg_lambda.p;
    b.dbg.rho = ctx->dbg_rho.p;
    b.dbg.delta = ctx->dbg_delta.p;
    b.dbg.dv = ctx->dbg_dv.p;
    b.dbg.eta = ctx->dbg_eta.p;
  }
}

// phase 0 = the whole substep; 1 = up to and including the last delta pass (positions final);
// 2 = the tail (XSPH, vorticity, restitution, scatter to original order).
int enqueue_substep(pbf_ctx* ctx, int phase = 0) {
  const NRef n = nref((int)ctx->n);
  cudaStream_t s = ctx->stream;
  GridBuffers g{};
  fill_grid_buffers(ctx, g);
  NeighborList nl;
  fill_neighbor_list(ctx, nl);
  const StepConsts& c = ctx->consts;
  StageTimer& t = ctx->timer;
  int launches = 0, k;
  SolveBuffers b{};
  fill_solve_buffers(ctx, b);
  const int iters = ctx->params.solver_iterations;
  if (phase == 2) return launch_solve(b, nl, c, iters, n, ctx->mode == PBF_MODE_STRICT, s, stage_mark, ctx, 2);

  stage_mark(ctx, PBF_STAGE_PREDICT, 1);
  k = launch_predict(ctx->pos_o.p, ctx->vel_o.p, ctx->pred_o.p, c, g, n, false, s);
  stage_mark(ctx, PBF_STAGE_PREDICT, 0);
  t.launches[PBF_STAGE_PREDICT] += k; launches += k;

  int out = 0;
  stage_mark(ctx, PBF_STAGE_SORT, 1);
  k = launch_sort(ctx->pred_o.p, c, g, n, &out, s);
  stage_mark(ctx, PBF_STAGE_SORT, 0);
  t.launches[PBF_STAGE_SORT] += k; launches += k;
  ctx->sorted_buf = out;

  stage_mark(ctx, PBF_STAGE_CELLS, 1);
  k = launch_cells_reorder(ctx->pred_o.p, ctx->pos_o.p, ctx->pred_a.p, ctx->pos_s.p, nullptr, g, n, s);
  stage_mark(ctx, PBF_STAGE_CELLS, 0);
  t.launches[PBF_STAGE_CELLS] += k; launches += k;

  if (nl.bricks) {  // brick path: runs of every brick (cells stage), then the list from shared-memory tiles
    stage_mark(ctx, PBF_STAGE_CELLS, 1);
    k = launch_brick_table(g, s);
    stage_mark(ctx, PBF_STAGE_CELLS, 0);
    t.launches[PBF_STAGE_CELLS] += k; launches += k;
  }
  stage_mark(ctx, PBF_STAGE_NEIGHBORS, 1);
  k = nl.bricks ? launch_neighbors_brick(ctx->pred_a.p, c, g, nl, s) : launch_neighbors(ctx->pred_a.p, c, g, nl, n, s);
  stage_mark(ctx, PBF_STAGE_NEIGHBORS, 0);
  t.launches[PBF_STAGE_NEIGHBORS] += k; launches += k;

  k = launch_solve(b, nl, c, iters, n, ctx->mode == PBF_MODE_STRICT, s, stage_mark, ctx, phase);
  // attribute solver launches to their stages
  if (iters > 0) {
    t.launches[PBF_STAGE_LAMBDA] += iters;
    t.launches[PBF_STAGE_DELTA] += iters;
    if (c.do_xsph) t.launches[PBF_STAGE_XSPH] += 1;
    if (c.do_vort) { t.launches[PBF_STAGE_VORT_OMEGA] += 1; t.launches[PBF_STAGE_VORT_APPLY] += 1; }
  } else {
    t.launches[PBF_STAGE_FINALIZE] += 1;
  }
  launches += k;
  return launches;
}

}  // namespace pbf

namespace {

int run_substeps(pbf_ctx* ctx, int nsteps) {
  const bool graph = ctx->use_graph && !ctx->profile;
  for (int sidx = 0; sidx < nsteps; ++sidx) {
    if (graph) {
      if (!ctx->graph_exec) {
        cudaGraph_t gr = nullptr;
        PBF_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        uint64_t saved[PBF_STAGE_COUNT];
        std::memcpy(saved, ctx->timer.launches, sizeof(saved));
        ctx->graph_kernels = enqueue_substep(ctx);
        std::memcpy(ctx->timer.launches, saved, sizeof(saved));
        PBF_CUDA(ctx, cudaStreamEndCapture(ctx->stream, &gr));
        PBF_CUDA(ctx, cudaGraphInstantiate(&ctx->graph_exec, gr, 0));
        cudaGraphDestroy(gr);
      }
      PBF_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
      ctx->launch_count += (uint64_t)ctx->graph_kernels;
    } else {
      ctx->launch_count += (uint64_t)enqueue_substep(ctx);
    }
  }
  PBF_CUDA(ctx, cudaGetLastError());
  return PBF_OK;
}

}  // namespace

namespace pbf {

int reset_status(pbf_ctx* ctx) {
  StatusBlock z{};
  for (int a = 0; a < 3; ++a) { z.min_cell[a] = INT_MAX; z.max_cell[a] = INT_MIN; }
  *ctx->status_host = z;
  PBF_CUDA(ctx, cudaMemcpyAsync(ctx->status.p, ctx->status_host, sizeof(StatusBlock), cudaMemcpyHostToDevice, ctx->stream));
  // ticket / exit counter of the persistent brick kernels (they leave both at zero themselves)
  if (ctx->brick_ctl.p) PBF_CUDA(ctx, cudaMemsetAsync(ctx->brick_ctl.p, 0, 2 * sizeof(unsigned int), ctx->stream));
  // The per-cell counters are zero between substeps and a sparse table is wiped by the next
  // k_predict; only fresh allocations and a batch that failed half-way need a reset here.
  if (ctx->tables_dirty && ctx->cell_count.p) {
    PBF_CUDA(ctx, cudaMemsetAsync(ctx->cell_count.p, 0, ctx->cell_count.n * sizeof(uint32_t), ctx->stream));
    PBF_CUDA(ctx, cudaMemsetAsync(ctx->cell_key.p, 0xff, ctx->cell_key.n * sizeof(unsigned long long), ctx->stream));
    PBF_CUDA(ctx, cudaMemsetAsync(ctx->desc.p, 0, sizeof(GridDesc), ctx->stream));
    ctx->tables_dirty = false;
  }
  return PBF_OK;
}

}  // namespace pbf

namespace {

template <typename T>
int fetch(pbf_ctx* ctx, std::vector<T>& host, const T* dev, size_t count) {
  host.resize(count);
  if (count) PBF_CUDA(ctx, cudaMemcpy(host.data(), dev, count * sizeof(T), cudaMemcpyDeviceToHost));
  return PBF_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

void pbf_default_params(pbf_params* p) {  // fluid::Params defaults, core.h:12-43
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
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
}

int pbf_abi_version(void) { return PBF_ABI_VERSION; }

int pbf_device_count(const char** err) {  // cuda_stub.cu:740-762
  if (err) *err = nullptr;
  int count = 0;
  const cudaError_t e = cudaGetDeviceCount(&count);
  if (e == cudaErrorNoDevice) return 0;
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    g_error = cudaGetErrorString(e);
    return PBF_E_CUDA;
  }
  return count;
}

pbf_ctx* pbf_create(int device, size_t capacity) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    g_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
              " (this backend has no CPU fallback)";
    return nullptr;
  }
  if (device < 0 || device >= count) {
    g_error = "device index out of range";
    return nullptr;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return nullptr;
  }
  pbf_ctx* ctx = new pbf_ctx();
  ctx->device = device;
  pbf_default_params(&ctx->params);
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaMallocHost(reinterpret_cast<void**>(&ctx->status_host), sizeof(StatusBlock))) != cudaSuccess ||
      (e = ctx->desc.reserve(1)) != cudaSuccess || (e = ctx->status.reserve(1)) != cudaSuccess ||
      (e = ctx->planes_dev.reserve(kMaxPlanes)) != cudaSuccess) {
    g_error = std::string("pbf_create: ") + cudaGetErrorString(e);
    pbf_destroy(ctx);
    return nullptr;
  }
  ctx->stream = ctx->own_stream;
  if (const char* env = std::getenv("PBF_BRICK")) {  // 0 = global gather, 1 = persistent bricks, 2 = one CTA per brick
    ctx->brick_want = env[0] != '0';
    ctx->brick_persist = env[0] != '2';
  }
  if (const int brc = brick_setup()) {
    g_error = std::string("pbf_create: shared-memory opt-in of the brick kernels failed: ") + cudaGetErrorString((cudaError_t)brc);
    pbf_destroy(ctx);
    return nullptr;
  }
  if (capacity > 0 && ensure_particles(ctx, capacity, 0) != PBF_OK) {
    g_error = ctx->error;
    pbf_destroy(ctx);
    return nullptr;
  }
  return ctx;
}

void pbf_destroy(pbf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  invalidate_graph(ctx);
  ctx->pos_o.release(); ctx->vel_o.release(); ctx->pos_bak.release(); ctx->vel_bak.release();
  ctx->pred_o.release(); ctx->pred_a.release(); ctx->pred_b.release(); ctx->pos_s.release();
  ctx->vel_a.release(); ctx->vel_b.release(); ctx->pv.release(); ctx->omega.release(); ctx->rho.release();
  ctx->planes_dev.release(); ctx->desc.release(); ctx->status.release();
  ctx->keys0.release(); ctx->keys1.release(); ctx->vals0.release(); ctx->vals1.release();
  ctx->cell_count.release(); ctx->cell_excl.release(); ctx->slot_id.release(); ctx->cell_key.release();
  ctx->chunk_total.release(); ctx->cell_range.release(); ctx->nbr_idx.release(); ctx->nbr_count.release();
  ctx->bricks.release();
  ctx->brick_ctl.release();
  for (auto& b : ctx->soa) b.release();
  ctx->dbg_lambda.release(); ctx->dbg_rho.release(); ctx->dbg_delta.release();
  ctx->dbg_dv.release(); ctx->dbg_eta.release();
  slab_release(ctx);
  for (auto& sn : ctx->snap) {
    for (int a = 0; a < 3; ++a) {
      sn.dev[a].release();
      if (sn.host[a]) cudaFreeHost(sn.host[a]);
    }
    if (sn.ready) cudaEventDestroy(sn.ready);
    if (sn.done) cudaEventDestroy(sn.done);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (auto ev : ctx->timer.pool) cudaEventDestroy(ev);
  for (auto ev : ctx->timer.begin) cudaEventDestroy(ev);
  for (auto ev : ctx->timer.end) cudaEventDestroy(ev);
  if (ctx->status_host) cudaFreeHost(ctx->status_host);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* pbf_last_error(const pbf_ctx* ctx) { return ctx ? ctx->error.c_str() : g_error.c_str(); }

int pbf_set_params(pbf_ctx* ctx, const pbf_params* p) {
  if (!ctx || !p) return fail(ctx, PBF_E_INVALID, "pbf_set_params: null argument");
  if (!(p->h > 0.0f)) return fail(ctx, PBF_E_INVALID, "pbf_set_params: h must be > 0 (the uniform-grid path needs a cell size)");
  if (!p->use_uniform_grid)
    return fail(ctx, PBF_E_INVALID, "pbf_set_params: use_uniform_grid=false (the O(N^2) path, core.cpp:248-268) is out of scope");
  if (!(p->dt > 0.0f)) return fail(ctx, PBF_E_INVALID, "pbf_set_params: dt must be > 0");
  if (p->solver_iterations < 0) return fail(ctx, PBF_E_INVALID, "pbf_set_params: solver_iterations < 0");
  ctx->params = *p;
  ctx->consts = make_consts(ctx->params, (int)ctx->planes_host.size());
  invalidate_graph(ctx);
  return PBF_OK;
}

// 1 when STRICT results are bit-identical to the reference CPU path for the current parameters,
// 0 (with a reason) for the configurations where they are only close:
//   * s_corr with an exponent outside {2, 3, 4}: the reference calls std::pow (core.cpp:59-71), the
//     device powf — not the same bits;
//   * solver_iterations == 0 with XSPH or vorticity on: the reference still runs those passes on
//     whatever rho its scratch arrays hold from an earlier call; this backend only commits.
// (NaN inputs are a third case: the max(x, 0) clamps of the kernels return 0 where the reference's
// branches propagate NaN; no parameter check can see that.)
int pbf_strict_exact(const pbf_ctx* ctx, const char** why) {
  if (why) *why = nullptr;
  if (!ctx) return PBF_E_INVALID;
  const pbf_params& p = ctx->params;
  if (p.enable_scorr && (p.scorr_n < 2 || p.scorr_n > 4)) {
    if (why) *why = "s_corr exponent outside {2,3,4}: device powf vs host std::pow";
    return 0;
  }
  if (p.solver_iterations == 0 && (p.enable_xsph || p.enable_vorticity)) {
    if (why) *why = "solver_iterations == 0 with XSPH / vorticity: the reference runs them on stale rho, this backend only commits";
    return 0;
  }
  return 1;
}

int pbf_set_planes(pbf_ctx* ctx, int count, const float* nx, const float* ny, const float* nz, const float* d) {
  if (!ctx || count < 0 || (count > 0 && (!nx || !ny || !nz || !d)))
    return fail(ctx, PBF_E_INVALID, "pbf_set_planes: bad arguments");
  if (count > kMaxPlanes) return fail(ctx, PBF_E_INVALID, "pbf_set_planes: more than 64 planes");
  cudaSetDevice(ctx->device);
  ctx->planes_host.resize(count);
  for (int i = 0; i < count; ++i) ctx->planes_host[i] = make_float4(nx[i], ny[i], nz[i], d[i]);
  if (count > 0) {
    PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PBF_CUDA(ctx, cudaMemcpy(ctx->planes_dev.p, ctx->planes_host.data(), count * sizeof(float4), cudaMemcpyHostToDevice));
  }
  ctx->consts = make_consts(ctx->params, count);
  invalidate_graph(ctx);
  return PBF_OK;
}

int pbf_set_mode(pbf_ctx* ctx, int mode) {
  if (!ctx || (mode != PBF_MODE_STRICT && mode != PBF_MODE_FAST)) return fail(ctx, PBF_E_INVALID, "pbf_set_mode: bad mode");
  if (mode != ctx->mode) invalidate_graph(ctx);
  ctx->mode = mode;
  return PBF_OK;
}

int pbf_set_stream(pbf_ctx* ctx, void* cuda_stream) {
  if (!ctx) return PBF_E_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  invalidate_graph(ctx);
  return PBF_OK;
}

int pbf_set_graph(pbf_ctx* ctx, int enabled) {
  if (!ctx) return PBF_E_INVALID;
  ctx->use_graph = enabled != 0;
  return PBF_OK;
}

}  // extern "C"

namespace {
// wait: the host arrays may be reused on return (pbf_upload).  pbf_step_host keeps them alive for
// the whole call and everything that follows is ordered on the same stream, so it does not wait
// (one host round trip less per substep of the cuda_step contract).
int upload_state(pbf_ctx* ctx, size_t n, const float* px, const float* py, const float* pz, const float* vx,
                 const float* vy, const float* vz, bool wait) {
  if (!ctx) return PBF_E_INVALID;
  if (n > 0 && (!px || !py || !pz || !vx || !vy || !vz)) return fail(ctx, PBF_E_INVALID, "pbf_upload: null array");
  if (n > 0x7fffffffu - 64) return fail(ctx, PBF_E_INVALID, "pbf_upload: more than 2^31 particles in one slab");
  cudaSetDevice(ctx->device);
  if (ctx->slab.enabled) return fail(ctx, PBF_E_INVALID, "pbf_upload: this context is a slab; use pbf_slab_upload");
  int rc = ensure_particles(ctx, n, 0);
  if (rc != PBF_OK) return rc;
  if (n != ctx->n) invalidate_graph(ctx);
  ctx->n = n;
  if (n == 0) return PBF_OK;
  const float* src[6] = {px, py, pz, vx, vy, vz};
  for (int a = 0; a < 6; ++a)
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->soa[a].p, src[a], n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  const float* dsoa[6] = {ctx->soa[0].p, ctx->soa[1].p, ctx->soa[2].p, ctx->soa[3].p, ctx->soa[4].p, ctx->soa[5].p};
  ctx->launch_count += launch_pack_state(dsoa, ctx->pos_o.p, ctx->vel_o.p, (int)n, ctx->stream);
  if (wait) PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PBF_OK;
}
}  // namespace

extern "C" {

int pbf_upload(pbf_ctx* ctx, size_t n, const float* px, const float* py, const float* pz, const float* vx,
               const float* vy, const float* vz) {
  return upload_state(ctx, n, px, py, pz, vx, vy, vz, true);
}

int pbf_download(pbf_ctx* ctx, float* px, float* py, float* pz, float* vx, float* vy, float* vz) {
  if (!ctx) return PBF_E_INVALID;
  cudaSetDevice(ctx->device);
  const size_t n = ctx->n;
  if (n == 0) return PBF_OK;
  float* dst[6] = {px, py, pz, vx, vy, vz};
  float* dsoa[6];
  for (int a = 0; a < 6; ++a) dsoa[a] = dst[a] ? ctx->soa[a].p : nullptr;
  ctx->launch_count += launch_unpack_state(ctx->pos_o.p, ctx->vel_o.p, dsoa, (int)n, ctx->stream);
  for (int a = 0; a < 6; ++a)
    if (dst[a]) PBF_CUDA(ctx, cudaMemcpyAsync(dst[a], ctx->soa[a].p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PBF_OK;
}

int pbf_step(pbf_ctx* ctx, int nsteps) {
  if (!ctx || nsteps < 0) return fail(ctx, PBF_E_INVALID, "pbf_step: bad arguments");
  if (!(ctx->params.h > 0.0f)) return fail(ctx, PBF_E_INVALID, "pbf_step: parameters not set (h == 0)");
  cudaSetDevice(ctx->device);
  if (nsteps == 0) return PBF_OK;
  if (ctx->slab.enabled) return slab_step(ctx, nsteps);
  if (ctx->n == 0) {  // core.cpp:122-125: only time advances
    for (int s = 0; s < nsteps; ++s) ctx->time += ctx->params.dt;
    return PBF_OK;
  }
  const size_t n = ctx->n;
  // batch backup: a substep that overflows a device table is re-run after growing it
  PBF_CUDA(ctx, cudaMemcpyAsync(ctx->pos_bak.p, ctx->pos_o.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  PBF_CUDA(ctx, cudaMemcpyAsync(ctx->vel_bak.p, ctx->vel_o.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  // brick path or global-gather family for this batch (a failed brick batch is replayed without it)
  bool brick = ctx->brick_want && ctx->params.solver_iterations > 0;
  if (brick && ctx->brick_retry > 0) {
    ctx->brick_retry--;
    brick = false;
  }
  for (int attempt = 0; attempt < 32; ++attempt) {
    if (brick != ctx->brick_on) {
      ctx->brick_on = brick;
      invalidate_graph(ctx);
    }
    int rc = ensure_tables(ctx);
    if (rc != PBF_OK) return rc;
    if ((rc = reset_status(ctx)) != PBF_OK) return rc;
    if ((rc = run_substeps(ctx, nsteps)) != PBF_OK) return rc;
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->status_host, ctx->status.p, sizeof(StatusBlock), cudaMemcpyDeviceToHost, ctx->stream));
    PBF_CUDA(ctx, cudaMemcpyAsync(&ctx->last_desc, ctx->desc.p, sizeof(GridDesc), cudaMemcpyDeviceToHost, ctx->stream));
    PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->profile) timer_resolve(ctx);
    const StatusBlock st = *ctx->status_host;
    ctx->last_status = st;
    if (!st.grid_overflow && !st.nbr_overflow && !st.brick_overflow) {
      for (int s = 0; s < nsteps; ++s) ctx->time += ctx->params.dt;  // core.cpp:614
      ctx->last_brick = ctx->brick_on;
      return PBF_OK;
    }
    // grow and replay the batch from the backup: results never depend on table sizes
    ctx->batches_retried++;
    ctx->tables_dirty = true;
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->pos_o.p, ctx->pos_bak.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->vel_o.p, ctx->vel_bak.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    if (st.grid_overflow) {
      const unsigned long long max_cells = ((unsigned long long)st.max_cells_hi << 32) | st.max_cells_lo;
      if (max_cells > (1ull << 30)) {
        char buf[256];
        std::snprintf(buf, sizeof(buf),
                      "pbf_step: bounding grid needs %llu cells (> 2^30 dense-table limit); particle positions "
                      "have diverged or are non-finite (state restored to the start of the batch)",
                      max_cells);
        return fail(ctx, PBF_E_CAPACITY, buf);
      }
      uint32_t cap = ctx->cell_cap;
      while ((unsigned long long)cap < max_cells + max_cells / 4) cap <<= 1;
      ctx->cell_cap = cap;
      invalidate_graph(ctx);
    }
    if (st.brick_overflow & kBrickDisable) {
      // a tile larger than kTileCap records, or the sparse cell table: this batch runs on the
      // global-gather family; the brick path is tried again a few batches later
      brick = false;
      ctx->brick_retry = 16;
      ctx->brick_fallbacks++;
    } else if (st.brick_overflow & kBrickGrow) {
      ctx->brick_cap = (int)std::min<unsigned long long>((unsigned long long)st.max_bricks + st.max_bricks / 4 + 64, 1ull << 28);
      invalidate_graph(ctx);
    }
    if (st.nbr_overflow) {
      // the batch stops at the first overflowing substep, so the maximum seen is a lower bound of
      // what the rest of the batch needs: grow by half, not by a sliver
      const unsigned need = st.max_neighbors + st.max_neighbors / 2 + 16;
      ctx->K = (int)((need + 7u) & ~7u);
      invalidate_graph(ctx);
    }
  }
  return fail(ctx, PBF_E_CAPACITY, "pbf_step: device tables kept overflowing after 32 growth attempts");
}

// ---- asynchronous frame output (replaces the blocking read of positions at
// reference app/src/main.cpp:259-273) --------------------------------------------------------------
int pbf_snapshot_begin(pbf_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot > 1) return fail(ctx, PBF_E_INVALID, "pbf_snapshot_begin: bad arguments");
  if (ctx->slab.enabled) return fail(ctx, PBF_E_INVALID, "pbf_snapshot_begin: not available on a slab context");
  cudaSetDevice(ctx->device);
  pbf_ctx::Snapshot& sn = ctx->snap[slot];
  if (sn.pending) return fail(ctx, PBF_E_INVALID, "pbf_snapshot_begin: the slot still holds a snapshot nobody waited for");
  const size_t n = ctx->n;
  if (!ctx->copy_stream) PBF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  if (!sn.ready) {
    PBF_CUDA(ctx, cudaEventCreateWithFlags(&sn.ready, cudaEventDisableTiming));
    PBF_CUDA(ctx, cudaEventCreateWithFlags(&sn.done, cudaEventDisableTiming));
  }
  if (n > sn.host_cap) {
    for (int a = 0; a < 3; ++a) {
      if (sn.host[a]) cudaFreeHost(sn.host[a]);
      sn.host[a] = nullptr;
    }
    sn.host_cap = 0;
    for (int a = 0; a < 3; ++a) PBF_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&sn.host[a]), n * sizeof(float)));
    sn.host_cap = n;
  }
  sn.n = n;
  sn.time = ctx->time;
  if (n) {
    for (int a = 0; a < 3; ++a) PBF_CUDA(ctx, sn.dev[a].reserve(n));
    // positions of this moment into the slot's own device copy (stream order: after the last batch) ...
    float* dsoa[6] = {sn.dev[0].p, sn.dev[1].p, sn.dev[2].p, nullptr, nullptr, nullptr};
    ctx->launch_count += launch_unpack_state(ctx->pos_o.p, ctx->vel_o.p, dsoa, (int)n, ctx->stream);
    PBF_CUDA(ctx, cudaEventRecord(sn.ready, ctx->stream));
    // ... and from there to pinned host memory on the copy stream, under whatever the compute stream does next
    PBF_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, sn.ready, 0));
    for (int a = 0; a < 3; ++a)
      PBF_CUDA(ctx, cudaMemcpyAsync(sn.host[a], sn.dev[a].p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
  }
  PBF_CUDA(ctx, cudaEventRecord(sn.done, ctx->copy_stream));
  sn.pending = true;
  return PBF_OK;
}

int pbf_snapshot_wait(pbf_ctx* ctx, int slot, const float** px, const float** py, const float** pz, size_t* n,
                      float* time) {
  if (!ctx || slot < 0 || slot > 1) return fail(ctx, PBF_E_INVALID, "pbf_snapshot_wait: bad arguments");
  pbf_ctx::Snapshot& sn = ctx->snap[slot];
  if (!sn.pending) return fail(ctx, PBF_E_INVALID, "pbf_snapshot_wait: no snapshot was begun in this slot");
  cudaSetDevice(ctx->device);
  PBF_CUDA(ctx, cudaEventSynchronize(sn.done));
  sn.pending = false;
  if (px) *px = sn.host[0];
  if (py) *py = sn.host[1];
  if (pz) *pz = sn.host[2];
  if (n) *n = sn.n;
  if (time) *time = sn.time;
  return PBF_OK;
}

int pbf_host_register(pbf_ctx* ctx, void* ptr, size_t bytes) {
  if (!ctx || !ptr || bytes == 0) return fail(ctx, PBF_E_INVALID, "pbf_host_register: bad arguments");
  cudaSetDevice(ctx->device);
  PBF_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return PBF_OK;
}

int pbf_host_unregister(pbf_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return fail(ctx, PBF_E_INVALID, "pbf_host_unregister: bad arguments");
  cudaSetDevice(ctx->device);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PBF_CUDA(ctx, cudaHostUnregister(ptr));
  return PBF_OK;
}

}  // extern "C"

namespace {

bool is_pinned(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

// The cuda_step contract (reference cuda_stub.cu:764-1099: six H2D copies, one substep, six D2H copies)
// as ONE CUDA graph on page-locked host arrays:
//     6 x H2D -> pack -> backup -> predict .. last delta pass -+-> XSPH / vorticity -> unpack vel -> 3 x D2H vel -+-> status
//                                                               +-> scatter pos -> 3 x D2H pos (side branch) ------+
// The positions of a substep are final after the last delta pass, so their 12 bytes per particle
// cross PCIe while the tail passes still run.  The host pointers are graph parameters: the graph is
// re-captured when the caller passes other arrays (the reference application passes the same State
// every step).  Returns PBF_OK without having stepped when this path does not apply (pageable memory,
// no tail pass, brick kernels, profiling, graphs off: *result = kHostGraphNotUsed) or when the substep
// overflowed a device table (kHostGraphOverflowed: the uploaded state is back in place and the caller
// runs the plain path, which grows the table and replays).
enum HostGraphResult { kHostGraphNotUsed = 0, kHostGraphDone = 1, kHostGraphOverflowed = 2 };

int step_host_graph(pbf_ctx* ctx, size_t n, float* const host[6], HostGraphResult* result) {
  *result = kHostGraphNotUsed;
  const StepConsts& c = ctx->consts;
  if (!ctx->use_graph || ctx->profile || ctx->slab.enabled || ctx->brick_want || n == 0) return PBF_OK;
  if (ctx->params.solver_iterations <= 0 || (!c.do_xsph && !c.do_vort)) return PBF_OK;
  for (int a = 0; a < 6; ++a)
    if (!is_pinned(host[a])) return PBF_OK;
  int rc = ensure_particles(ctx, n, 0);
  if (rc != PBF_OK) return rc;
  if (n != ctx->n) invalidate_graph(ctx);
  ctx->n = n;
  if (ctx->brick_on) {
    ctx->brick_on = false;
    invalidate_graph(ctx);
  }
  if ((rc = ensure_tables(ctx)) != PBF_OK) return rc;
  if (!ctx->side_stream) {
    PBF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    PBF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    PBF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  }
  bool same = ctx->host_graph != nullptr;
  for (int a = 0; a < 6; ++a) same = same && ctx->host_graph_ptr[a] == host[a];
  if ((rc = reset_status(ctx)) != PBF_OK) return rc;  // outside the graph: it may have to wipe tables
  if (!same) {
    if (ctx->host_graph) {
      cudaGraphExecDestroy(ctx->host_graph);
      ctx->host_graph = nullptr;
    }
    cudaStream_t s = ctx->stream;
    cudaGraph_t gr = nullptr;
    uint64_t saved[PBF_STAGE_COUNT];
    std::memcpy(saved, ctx->timer.launches, sizeof(saved));
    PBF_CUDA(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int k = 0;
    bool ok = true;
    auto chk = [&](cudaError_t e) { ok = ok && e == cudaSuccess; };
    for (int a = 0; a < 6; ++a) chk(cudaMemcpyAsync(ctx->soa[a].p, host[a], n * sizeof(float), cudaMemcpyHostToDevice, s));
    const float* dsoa[6] = {ctx->soa[0].p, ctx->soa[1].p, ctx->soa[2].p, ctx->soa[3].p, ctx->soa[4].p, ctx->soa[5].p};
    k += launch_pack_state(dsoa, ctx->pos_o.p, ctx->vel_o.p, (int)n, s);
    chk(cudaMemcpyAsync(ctx->pos_bak.p, ctx->pos_o.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, s));
    chk(cudaMemcpyAsync(ctx->vel_bak.p, ctx->vel_o.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, s));
    k += enqueue_substep(ctx, 1);
    // side branch: final positions -> SoA staging -> host, under the tail passes
    chk(cudaEventRecord(ctx->ev_fork, s));
    chk(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    const float4* pos_final = (ctx->params.solver_iterations & 1) ? ctx->pred_b.p : ctx->pred_a.p;
    k += launch_scatter_positions(pos_final, ctx->pos_s.p, ctx->soa[0].p, ctx->soa[1].p, ctx->soa[2].p, (int)n, ctx->side_stream);
    for (int a = 0; a < 3; ++a)
      chk(cudaMemcpyAsync(host[a], ctx->soa[a].p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->side_stream));
    chk(cudaEventRecord(ctx->ev_join, ctx->side_stream));
    // main branch: the tail, then the velocities
    k += enqueue_substep(ctx, 2);
    float* vsoa[6] = {nullptr, nullptr, nullptr, ctx->soa[3].p, ctx->soa[4].p, ctx->soa[5].p};
    k += launch_unpack_state(ctx->pos_o.p, ctx->vel_o.p, vsoa, (int)n, s);
    for (int a = 3; a < 6; ++a)
      chk(cudaMemcpyAsync(host[a], ctx->soa[a].p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    chk(cudaMemcpyAsync(ctx->status_host, ctx->status.p, sizeof(StatusBlock), cudaMemcpyDeviceToHost, s));
    chk(cudaMemcpyAsync(&ctx->last_desc, ctx->desc.p, sizeof(GridDesc), cudaMemcpyDeviceToHost, s));
    chk(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    const cudaError_t ce = cudaStreamEndCapture(s, &gr);
    std::memcpy(ctx->timer.launches, saved, sizeof(saved));
    if (!ok || ce != cudaSuccess || !gr) {
      if (gr) cudaGraphDestroy(gr);
      cudaGetLastError();
      return fail(ctx, PBF_E_CUDA, std::string("pbf_step_host: capturing the contract graph failed: ") + cudaGetErrorString(ce));
    }
    const cudaError_t ie = cudaGraphInstantiate(&ctx->host_graph, gr, 0);
    cudaGraphDestroy(gr);
    if (ie != cudaSuccess) {
      ctx->host_graph = nullptr;
      return fail(ctx, PBF_E_CUDA, std::string("pbf_step_host: cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    }
    ctx->host_graph_kernels = k;
    for (int a = 0; a < 6; ++a) ctx->host_graph_ptr[a] = host[a];
  }
  PBF_CUDA(ctx, cudaGraphLaunch(ctx->host_graph, ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->launch_count += (uint64_t)ctx->host_graph_kernels;
  const StatusBlock st = *ctx->status_host;
  ctx->last_status = st;
  if (st.grid_overflow || st.nbr_overflow || st.brick_overflow) {
    // back to the uploaded state; the plain path below finds the same overflow, grows and replays
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->pos_o.p, ctx->pos_bak.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->vel_o.p, ctx->vel_bak.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->tables_dirty = true;
    *result = kHostGraphOverflowed;
    return PBF_OK;
  }
  ctx->time += ctx->params.dt;  // core.cpp:614
  ctx->last_brick = false;
  *result = kHostGraphDone;
  return PBF_OK;
}

}  // namespace

extern "C" {

int pbf_step_host(pbf_ctx* ctx, size_t n, float* px, float* py, float* pz, float* vx, float* vy, float* vz, int nsteps) {
  if (!ctx || nsteps < 0) return fail(ctx, PBF_E_INVALID, "pbf_step_host: bad arguments");
  if (n > 0 && (!px || !py || !pz || !vx || !vy || !vz)) return fail(ctx, PBF_E_INVALID, "pbf_step_host: null array");
  if (!(ctx->params.h > 0.0f)) return fail(ctx, PBF_E_INVALID, "pbf_step: parameters not set (h == 0)");
  cudaSetDevice(ctx->device);
  int rc;
  if (nsteps == 1 && !ctx->slab.enabled && n <= 0x7fffffffu - 64) {
    float* const host[6] = {px, py, pz, vx, vy, vz};
    HostGraphResult how = kHostGraphNotUsed;
    if ((rc = step_host_graph(ctx, n, host, &how)) != PBF_OK) return rc;
    if (how == kHostGraphDone) return PBF_OK;
    if (how == kHostGraphOverflowed) {
      // the contract graph uploaded the state and hit a table limit (state restored to the upload):
      // the plain path finds the same overflow, grows the table and replays
      if ((rc = pbf_step(ctx, nsteps)) != PBF_OK) return rc;
      return pbf_download(ctx, px, py, pz, vx, vy, vz);
    }
  }
  rc = upload_state(ctx, n, px, py, pz, vx, vy, vz, false);
  if (rc != PBF_OK) return rc;
  if ((rc = pbf_step(ctx, nsteps)) != PBF_OK) return rc;
  return pbf_download(ctx, px, py, pz, vx, vy, vz);
}

uint64_t pbf_batches_retried(const pbf_ctx* ctx) { return ctx ? ctx->batches_retried : 0; }
size_t pbf_count(const pbf_ctx* ctx) { return ctx ? ctx->n : 0; }
float pbf_time(const pbf_ctx* ctx) { return ctx ? ctx->time : 0.0f; }
int pbf_set_time(pbf_ctx* ctx, float t) {
  if (!ctx) return PBF_E_INVALID;
  ctx->time = t;
  return PBF_OK;
}

// ---------------------------------------------------------------- parity surface
int pbf_debug_enable(pbf_ctx* ctx, int enabled) {
  if (!ctx) return PBF_E_INVALID;
  if ((enabled != 0) != ctx->debug) invalidate_graph(ctx);
  ctx->debug = enabled != 0;
  return PBF_OK;
}

int pbf_debug_sizes(pbf_ctx* ctx, size_t* ncells, size_t* nneighbors) {
  if (!ctx) return PBF_E_INVALID;
  cudaSetDevice(ctx->device);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (nneighbors) {  // sum of the per-particle counts of the last substep
    std::vector<uint32_t> counts;
    int rc = fetch(ctx, counts, ctx->nbr_count.p, ctx->n);
    if (rc != PBF_OK) return rc;
    size_t total = 0;
    for (uint32_t c : counts) total += c;
    *nneighbors = total;
  }
  if (ncells) {
    size_t occupied = 0;
    if (ctx->n) {
      std::vector<int2> table;
      int rc = fetch(ctx, table, ctx->cell_range.p, (size_t)ctx->last_desc.ncells);
      if (rc != PBF_OK) return rc;
      for (const int2& r : table) occupied += (r.y > r.x) ? 1 : 0;
    }
    *ncells = occupied;
  }
  return PBF_OK;
}

int pbf_debug_grid(pbf_ctx* ctx, int32_t* ecx, int32_t* ecy, int32_t* ecz, int32_t* eparticle,
                   int32_t* cell_xyz, int32_t* cell_start, int32_t* cell_end) {
  if (!ctx) return PBF_E_INVALID;
  cudaSetDevice(ctx->device);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const size_t n = ctx->n;
  if (n == 0) return PBF_OK;
  const GridDesc d = ctx->last_desc;
  std::vector<uint32_t> keys, vals;
  int rc;
  if ((rc = fetch(ctx, keys, ctx->sorted_buf ? ctx->keys1.p : ctx->keys0.p, n)) != PBF_OK) return rc;
  if ((rc = fetch(ctx, vals, ctx->sorted_buf ? ctx->vals1.p : ctx->vals0.p, n)) != PBF_OK) return rc;
  // cell coordinates per device slot
  std::vector<int32_t> x(n), y(n), z(n);
  if (d.sparse) {
    std::vector<unsigned long long> cell_key;
    if ((rc = fetch(ctx, cell_key, ctx->cell_key.p, (size_t)d.ncells)) != PBF_OK) return rc;
    for (size_t i = 0; i < n; ++i) {
      const unsigned long long k = cell_key[keys[i]];
      x[i] = (int32_t)(k >> 42) + d.lo[0];
      y[i] = (int32_t)((k >> 21) & 0x1fffffu) + d.lo[1];
      z[i] = (int32_t)(k & 0x1fffffu) + d.lo[2];
    }
  } else {
    const uint32_t dy = (uint32_t)d.dim[1], dz = (uint32_t)d.dim[2];
    for (size_t i = 0; i < n; ++i) {
      z[i] = (int32_t)(keys[i] % dz) + d.lo[2];
      y[i] = (int32_t)((keys[i] / dz) % dy) + d.lo[1];
      x[i] = (int32_t)(keys[i] / (dz * dy)) + d.lo[0];
    }
  }
  // The reference order is lexicographic in (x, y, z), then particle id (core.cpp:12-21, 182).
  // A dense table stores the slots in that order already; a sparse one stores the cells in hash
  // order (each cell still contiguous and ordered by id), so the slots are re-ordered here.
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
  if (d.sparse)
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      if (x[a] != x[b]) return x[a] < x[b];
      if (y[a] != y[b]) return y[a] < y[b];
      return z[a] < z[b];
    });
  size_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t s = order[i];
    if (ecx) ecx[i] = x[s];
    if (ecy) ecy[i] = y[s];
    if (ecz) ecz[i] = z[s];
    if (eparticle) eparticle[i] = (int32_t)vals[s];
    const bool first = i == 0 || x[s] != x[order[i - 1]] || y[s] != y[order[i - 1]] || z[s] != z[order[i - 1]];
    if (first) {  // run-length table (core.cpp:185-203)
      if (k > 0 && cell_end) cell_end[k - 1] = (int32_t)i;
      if (cell_xyz) { cell_xyz[3 * k] = x[s]; cell_xyz[3 * k + 1] = y[s]; cell_xyz[3 * k + 2] = z[s]; }
      if (cell_start) cell_start[k] = (int32_t)i;
      ++k;
    }
  }
  if (k > 0 && cell_end) cell_end[k - 1] = (int32_t)n;
  return PBF_OK;
}

int pbf_debug_neighbors(pbf_ctx* ctx, int32_t* prefix_sum, int32_t* indices) {
  if (!ctx) return PBF_E_INVALID;
  cudaSetDevice(ctx->device);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const size_t n = ctx->n;
  if (n == 0) return PBF_OK;
  std::vector<uint32_t> vals, counts, list;
  int rc;
  if ((rc = fetch(ctx, vals, ctx->sorted_buf ? ctx->vals1.p : ctx->vals0.p, n)) != PBF_OK) return rc;
  if ((rc = fetch(ctx, counts, ctx->nbr_count.p, n)) != PBF_OK) return rc;

  const size_t K = (size_t)ctx->K;
  std::vector<uint32_t> slot_of(n);
  for (size_t s = 0; s < n; ++s) slot_of[vals[s]] = (uint32_t)s;
  if (ctx->last_brick) {
    // brick path: 16-bit entries are byte offsets into the tile of the particle's brick; translate
    // them back to sorted slots with the brick table of the last substep
    const GridDesc d = ctx->last_desc;
    std::vector<uint32_t> keys;
    std::vector<BrickRec> recs;
    std::vector<uint16_t> list16;
    if ((rc = fetch(ctx, keys, ctx->sorted_buf ? ctx->keys1.p : ctx->keys0.p, n)) != PBF_OK) return rc;
    if ((rc = fetch(ctx, recs, ctx->bricks.p, (size_t)d.nbricks)) != PBF_OK) return rc;
    if ((rc = fetch(ctx, list16, reinterpret_cast<const uint16_t*>(ctx->nbr_idx.p), ((n + 31) / 32) * K * 32)) != PBF_OK) return rc;
    const uint32_t dy = (uint32_t)d.dim[1], dz = (uint32_t)d.dim[2];
    size_t total = 0;
    for (size_t o = 0; o < n; ++o) {  // original particle order, like core.cpp:205
      const size_t s = slot_of[o];
      const uint32_t z = keys[s] % dz, y = (keys[s] / dz) % dy, x = keys[s] / (dz * dy);
      const size_t b = ((size_t)(x / kBrickX) * (size_t)d.bdim[1] + (size_t)(y / kBrickY)) * (size_t)d.bdim[2] + (size_t)(z / kBrickZ);
      const BrickRec& rec = recs[b];
      const uint16_t* row = list16.data() + (s >> 5) * K * 32 + (s & 31) * 4;
      for (uint32_t k = 0; k < counts[s]; ++k) {
        const int te = (int)(row[(size_t)(k >> 2) * 128 + (k & 3)] >> 4);
        int c = 0;
        while (c + 1 < kBrickCols && rec.col_base[c + 1] <= te) ++c;
        if (indices) indices[total] = (int32_t)vals[(size_t)(rec.col_start[c] + (te - rec.col_base[c]))];
        ++total;
      }
      if (prefix_sum) prefix_sum[o] = (int32_t)total;
    }
    return PBF_OK;
  }
  if ((rc = fetch(ctx, list, ctx->nbr_idx.p, ((n + 31) / 32) * K * 32)) != PBF_OK) return rc;
  size_t total = 0;
  for (size_t o = 0; o < n; ++o) {  // original particle order, like core.cpp:205
    const size_t s = slot_of[o];
    const uint32_t* row = list.data() + (s >> 5) * K * 32 + (s & 31) * 2;
    for (uint32_t k = 0; k < counts[s]; ++k) {
      if (indices) indices[total] = (int32_t)vals[row[(size_t)(k >> 1) * 64 + (k & 1)]];
      ++total;
    }
    if (prefix_sum) prefix_sum[o] = (int32_t)total;
  }
  return PBF_OK;
}

int pbf_debug_scratch(pbf_ctx* ctx, int id, float* out) {
  if (!ctx || !out || id < 0 || id >= PBF_SCRATCH_COUNT) return fail(ctx, PBF_E_INVALID, "pbf_debug_scratch: bad arguments");
  cudaSetDevice(ctx->device);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const size_t n = ctx->n;
  if (n == 0) return PBF_OK;
  int rc;
  if (id <= PBF_SCRATCH_PRED_Z) {  // scratch.pred_* ends the step equal to the committed positions
    std::vector<float4> pos;
    if ((rc = fetch(ctx, pos, ctx->pos_o.p, n)) != PBF_OK) return rc;
    for (size_t i = 0; i < n; ++i) out[i] = (&pos[i].x)[id - PBF_SCRATCH_PRED_X];
    return PBF_OK;
  }
  std::vector<uint32_t> vals;
  if ((rc = fetch(ctx, vals, ctx->sorted_buf ? ctx->vals1.p : ctx->vals0.p, n)) != PBF_OK) return rc;
  auto need_debug = [&]() { return fail(ctx, PBF_E_INVALID, "pbf_debug_scratch: call pbf_debug_enable(ctx, 1) before stepping"); };
  if (id == PBF_SCRATCH_LAMBDA || id == PBF_SCRATCH_RHO) {
    if (!ctx->debug) return need_debug();
    std::vector<float> v;
    if ((rc = fetch(ctx, v, id == PBF_SCRATCH_LAMBDA ? ctx->dbg_lambda.p : ctx->dbg_rho.p, n)) != PBF_OK) return rc;
    for (size_t s = 0; s < n; ++s) out[vals[s]] = v[s];
    return PBF_OK;
  }
  const float4* src = nullptr;
  int comp = 0;
  if (id >= PBF_SCRATCH_DELTA_X && id <= PBF_SCRATCH_DELTA_Z) { src = ctx->dbg_delta.p; comp = id - PBF_SCRATCH_DELTA_X; if (!ctx->debug) return need_debug(); }
  else if (id >= PBF_SCRATCH_DV_X && id <= PBF_SCRATCH_DV_Z) { src = ctx->dbg_dv.p; comp = id - PBF_SCRATCH_DV_X; if (!ctx->debug) return need_debug(); }
  else if (id >= PBF_SCRATCH_OMEGA_X && id <= PBF_SCRATCH_OMEGA_MAG) { src = ctx->omega.p; comp = id - PBF_SCRATCH_OMEGA_X; }
  else if (id >= PBF_SCRATCH_ETA_X && id <= PBF_SCRATCH_ETA_Z) { src = ctx->dbg_eta.p; comp = id - PBF_SCRATCH_ETA_X; if (!ctx->debug) return need_debug(); }
  std::vector<float4> v;
  if ((rc = fetch(ctx, v, src, n)) != PBF_OK) return rc;
  for (size_t s = 0; s < n; ++s) out[vals[s]] = (&v[s].x)[comp];
  return PBF_OK;
}

// ---------------------------------------------------------------- measurement surface
static const char* kStageNames[PBF_STAGE_COUNT] = {"predict", "sort", "cells", "neighbors", "lambda", "delta",
                                                   "xsph", "vort_omega", "vort_apply", "finalize", "exchange"};

const char* pbf_stage_name(int stage) { return (stage >= 0 && stage < PBF_STAGE_COUNT) ? kStageNames[stage] : "?"; }

int pbf_profile_enable(pbf_ctx* ctx, int enabled) {
  if (!ctx) return PBF_E_INVALID;
  ctx->profile = enabled != 0;
  return PBF_OK;
}

int pbf_profile_reset(pbf_ctx* ctx) {
  if (!ctx) return PBF_E_INVALID;
  for (int s = 0; s < PBF_STAGE_COUNT; ++s) { ctx->timer.total_ms[s] = 0; ctx->timer.launches[s] = 0; }
  return PBF_OK;
}

int pbf_profile_get(pbf_ctx* ctx, int stage, double* total_ms, uint64_t* launches) {
  if (!ctx || stage < 0 || stage >= PBF_STAGE_COUNT) return PBF_E_INVALID;
  if (total_ms) *total_ms = ctx->timer.total_ms[stage];
  if (launches) *launches = ctx->timer.launches[stage];
  return PBF_OK;
}

uint64_t pbf_launch_count(const pbf_ctx* ctx) { return ctx ? ctx->launch_count : 0; }

// Test hook (not in pbf_b200.h): shrink the device tables so the overflow -> grow -> replay
// path of pbf_step can be exercised on small inputs.
int pbf_debug_set_capacity(pbf_ctx* ctx, int K, uint32_t cell_cap) {
  if (!ctx || K < 2 || cell_cap < 8) return PBF_E_INVALID;
  ctx->K = (K + 1) & ~1;
  uint32_t cap = 8;  // a power of two: the sparse table masks with cell_cap - 1
  while (cap < cell_cap) cap <<= 1;
  ctx->cell_cap = cap;
  invalidate_graph(ctx);
  return PBF_OK;
}

// Kernel family of the passes that walk the neighbour list (same results bit for bit):
// PBF_BRICK_OFF = global-gather kernels (default: measured faster on B200),
// PBF_BRICK_PERSISTENT / PBF_BRICK_PER_CTA = shared-memory staged bricks.  Environment PBF_BRICK=0|1|2
// sets the initial choice of a context.
int pbf_set_brick(pbf_ctx* ctx, int mode) {
  if (!ctx || mode < PBF_BRICK_OFF || mode > PBF_BRICK_PER_CTA) return PBF_E_INVALID;
  const bool want = mode != PBF_BRICK_OFF, persist = mode != PBF_BRICK_PER_CTA;
  if (want && persist != ctx->brick_persist) invalidate_graph(ctx);
  ctx->brick_want = want;
  if (want) ctx->brick_persist = persist;
  ctx->brick_retry = 0;
  return PBF_OK;
}
// 1 if the last completed pbf_step batch ran on the brick path; *fallbacks (may be NULL) counts the
// batches that had to be replayed on the global-gather kernels, *max_tile the largest tile seen.
int pbf_brick_status(const pbf_ctx* ctx, uint64_t* fallbacks, uint32_t* max_tile) {
  if (!ctx) return PBF_E_INVALID;
  if (fallbacks) *fallbacks = ctx->brick_fallbacks;
  if (max_tile) *max_tile = ctx->last_status.max_tile;
  return ctx->last_brick ? 1 : 0;
}

// Test hook (not in pbf_b200.h): 1 if the last substep used the sparse (hashed) cell table.
int pbf_debug_grid_is_sparse(pbf_ctx* ctx) { return ctx ? ctx->last_desc.sparse : -1; }

// ---------------------------------------------------------------- slab decomposition
// Implemented in pbf_slab.cu.

}  // extern "C"

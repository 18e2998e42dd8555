// solve.cu — constraint-solve and post-solve kernels of the PBF substep for sm_100a:
//   a8  lambda                      (reference core/src/core.cpp:281-329)
//   a9  delta-p + s_corr + planes   (core.cpp:334-398)   fused with
//   a10 apply delta                 (core.cpp:400-407)   via the pred ping-pong buffer, and on the
//   a11 velocity update + commit    (core.cpp:410-421)   last iteration
//   a12 XSPH viscosity              (core.cpp:423-466)
//   a13 vorticity confinement       (core.cpp:468-571)   omega pass, eta+apply pass
//   a14 plane restitution/friction  (core.cpp:573-612)   fused into whichever pass is last,
//       together with the scatter back to original particle order.
//
// Every pass is one thread per sorted particle walking its neighbour list (built once per
// substep by k_neighbors, so the neighbour SET is the reference's: fixed at grid-build time,
// re-tested with r2 < h2 on current positions).  Two neighbours are processed per step with
// Blackwell's packed f32x2 instructions (FADD2/FMUL2/FFMA2): one 8-byte list load brings two
// indices (a coalesced 256-byte line pair per warp), each neighbour costs one 16-byte gather
// because everything a pass needs from particle j is packed into a single float4:
//   lambda pass   (pred.xyz, -)          delta pass   (pred.xyz, lambda_j)
//   XSPH          (pos.xyz, -) + (vel.xyz, m/rho_j)
//   omega pass    (pos.xyz, -) + (vel.xyz, -)          eta pass   (pos.xyz, |omega_j|)
// and the per-neighbour terms are then accumulated in list order with scalar adds, so the
// summation order is the reference's.
//
// STRICT = true: every operation is a correctly rounded IEEE binary32 op in the reference's
// expression order (bit-identical to the CPU path).  STRICT = false: FMA contraction and
// x*rsqrt(x); tolerance-gated.
#include "pbf_kernels.h"
#include "solve_passes.cuh"

namespace pbf {

namespace {

#ifndef PBF_SOLVE_BLOCK
#define PBF_SOLVE_BLOCK 128
#endif
#ifndef PBF_PAIR_UNROLL
#define PBF_PAIR_UNROLL 2
#endif
#ifndef PBF_SOLVE_MINBLOCKS
#define PBF_SOLVE_MINBLOCKS 1
#endif
#ifndef PBF_LIST_PREFETCH
#define PBF_LIST_PREFETCH 1
#endif
constexpr int kBlock = PBF_SOLVE_BLOCK;
constexpr int kPairUnroll = PBF_PAIR_UNROLL;  // neighbour PAIRS fetched per batch (4 independent gathers in flight)

// ---- pair iteration -----------------------------------------------------------------
// Walks the list of sorted slot i.  Full pairs run without any validity logic; an odd count
// ends with one half-valid pair.  body(a0, a1, v1): data of the two neighbours and validity of the second.
// List loads are streaming (ld.global.cs): every entry is used once per pass and must not
// evict the gathered particle arrays from L1/L2.
template <typename Body>
__device__ __forceinline__ void for_each_pair(const uint32_t* __restrict__ nbr_idx, int K, int i,
                                              uint32_t cnt, const float4* __restrict__ a4, Body&& body) {
  const uint2* row = reinterpret_cast<const uint2*>(nbr_idx + (size_t)(i >> 5) * (size_t)K * 32u) + (i & 31);
  const uint32_t npairs = cnt >> 1;
  uint32_t p = 0;
#if PBF_LIST_PREFETCH
  // the indices of batch p+1 are requested before batch p is gathered: the list streams from
  // DRAM/L2, and without this every batch pays list latency + gather latency back to back
  uint2 jn[kPairUnroll];
  if (kPairUnroll <= npairs) {
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) jn[u] = __ldcs(row + (size_t)u * 32u);
  }
  for (; p + kPairUnroll <= npairs; p += kPairUnroll) {
    uint2 j[kPairUnroll];
    float4 a0[kPairUnroll], a1[kPairUnroll];
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) j[u] = jn[u];
    if (p + 2 * kPairUnroll <= npairs) {
#pragma unroll
      for (int u = 0; u < kPairUnroll; ++u) jn[u] = __ldcs(row + (size_t)(p + kPairUnroll + u) * 32u);
    }
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) {
      a0[u] = a4[j[u].x];
      a1[u] = a4[j[u].y];
    }
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) body(a0[u], a1[u], true);
  }
#else
  for (; p + kPairUnroll <= npairs; p += kPairUnroll) {
    uint2 j[kPairUnroll];
    float4 a0[kPairUnroll], a1[kPairUnroll];
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) j[u] = __ldcs(row + (size_t)(p + u) * 32u);
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) {
      a0[u] = a4[j[u].x];
      a1[u] = a4[j[u].y];
    }
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) body(a0[u], a1[u], true);
  }
#endif
  for (; p < npairs; ++p) {
    const uint2 j = __ldcs(row + (size_t)p * 32u);
    const float4 a0 = a4[j.x], a1 = a4[j.y];
    body(a0, a1, true);
  }
  if (cnt & 1u) {
    const uint32_t j = __ldcs(reinterpret_cast<const uint32_t*>(row + (size_t)npairs * 32u));
    const float4 a0 = a4[j];
    body(a0, a4[i], false);
  }
}

// Same walk for the passes that need two float4 per neighbour; fetch(j, a, b) gathers them.
template <typename Fetch, typename Body>
__device__ __forceinline__ void for_each_pair2(const uint32_t* __restrict__ nbr_idx, int K, int i,
                                               uint32_t cnt, Fetch fetch, Body&& body) {
  const uint2* row = reinterpret_cast<const uint2*>(nbr_idx + (size_t)(i >> 5) * (size_t)K * 32u) + (i & 31);
  const uint32_t npairs = cnt >> 1;
  uint32_t p = 0;
#if PBF_LIST_PREFETCH
  uint2 jn[kPairUnroll];
  if (kPairUnroll <= npairs) {
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) jn[u] = __ldcs(row + (size_t)u * 32u);
  }
  for (; p + kPairUnroll <= npairs; p += kPairUnroll) {
    uint2 j[kPairUnroll];
    float4 a0[kPairUnroll], a1[kPairUnroll], b0[kPairUnroll], b1[kPairUnroll];
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) j[u] = jn[u];
    if (p + 2 * kPairUnroll <= npairs) {
#pragma unroll
      for (int u = 0; u < kPairUnroll; ++u) jn[u] = __ldcs(row + (size_t)(p + kPairUnroll + u) * 32u);
    }
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) {
      fetch(j[u].x, a0[u], b0[u]);
      fetch(j[u].y, a1[u], b1[u]);
    }
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) body(a0[u], a1[u], b0[u], b1[u], true);
  }
#else
  for (; p + kPairUnroll <= npairs; p += kPairUnroll) {
    uint2 j[kPairUnroll];
    float4 a0[kPairUnroll], a1[kPairUnroll], b0[kPairUnroll], b1[kPairUnroll];
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) j[u] = __ldcs(row + (size_t)(p + u) * 32u);
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) {
      fetch(j[u].x, a0[u], b0[u]);
      fetch(j[u].y, a1[u], b1[u]);
    }
#pragma unroll
    for (int u = 0; u < kPairUnroll; ++u) body(a0[u], a1[u], b0[u], b1[u], true);
  }
#endif
  for (; p < npairs; ++p) {
    const uint2 j = __ldcs(row + (size_t)p * 32u);
    float4 a0, a1, b0, b1;
    fetch(j.x, a0, b0);
    fetch(j.y, a1, b1);
    body(a0, a1, b0, b1, true);
  }
  if (cnt & 1u) {
    const uint32_t j = __ldcs(reinterpret_cast<const uint32_t*>(row + (size_t)npairs * 32u));
    float4 a0, a1, b0, b1;
    fetch(j, a0, b0);
    fetch((uint32_t)i, a1, b1);
    body(a0, a1, b0, b1, false);
  }
}

// ---------------------------------------------------------------- a8 lambda
template <bool S>
__global__ void __launch_bounds__(kBlock, PBF_SOLVE_MINBLOCKS)
k_lambda(float4* __restrict__ pred, const uint32_t* __restrict__ nbr_idx,
         const uint32_t* __restrict__ nbr_count, float* __restrict__ rho_out, StepConsts c,
         const StatusBlock* st, DebugPtrs dbg, Span span, int K, NRef nr) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int i = span.slot(blockIdx.x * blockDim.x + threadIdx.x, nr.get());
  if (i < 0) return;
  const float4 pi = pred[i];
  LambdaPass<S> acc(pi, c);
  for_each_pair(nbr_idx, K, i, nbr_count[i], pred, acc);
  float lambda, rho;
  acc.finish(lambda, rho);
  // only .w is written; concurrent readers of pred[i] use .xyz only in this pass
  reinterpret_cast<float*>(pred + i)[3] = lambda;
  rho_out[i] = rho;
  if (dbg.lambda) dbg.lambda[i] = lambda;
  if (dbg.rho) dbg.rho[i] = rho;
}

// ---------------------------------------------------------------- a9 + a10 (+ a11, a14)
// LAST: also velocity update / commit; is_final: additionally restitution + scatter.
// COMMON: see DeltaPass (solve_passes.cuh).
// MEASURED on B200 (fluid_million, settled): 76.7 -> 70.7 us per launch.  The same specialisation of
// k_lambda was slower (68.1 -> 72.6 us at the 56 registers of the generic kernel) and is not used.
template <bool S, bool LAST, bool COMMON>
__global__ void __launch_bounds__(kBlock, PBF_SOLVE_MINBLOCKS)
k_delta(const float4* __restrict__ pred_in, float4* __restrict__ pred_out,
        const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count,
        const float4* __restrict__ pos_s, const float* __restrict__ rho, float4* __restrict__ vel_out,
        PosVel* __restrict__ pv, const float4* __restrict__ planes, float4* __restrict__ pos_o,
        float4* __restrict__ vel_o, StepConsts c, const StatusBlock* st, DebugPtrs dbg, HaloOut halo, int is_final,
        int K, NRef nr) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pred_in[i];
  DeltaPass<S, COMMON> acc(pi, c);
  for_each_pair(nbr_idx, K, i, nbr_count[i], pred_in, acc);
  V3<FT<S>> np;
  float4 dlt;
  acc.finish(pi, planes, np, dlt);
  delta_store<S, LAST>(i, np, dlt, pred_out, pos_s, rho, vel_out, pv, planes, pos_o, vel_o, c, dbg, halo, is_final);
}

// ---------------------------------------------------------------- a12 XSPH
template <bool S>
__global__ void __launch_bounds__(kBlock, PBF_SOLVE_MINBLOCKS)
k_xsph(const float4* __restrict__ pos, const float4* __restrict__ vel_in, const PosVel* __restrict__ pv,
       float4* __restrict__ vel_out, const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count,
       const float4* __restrict__ pos_s, const float4* __restrict__ planes, float4* __restrict__ pos_o,
       float4* __restrict__ vel_o, StepConsts c, const StatusBlock* st, DebugPtrs dbg, HaloOut halo, int is_final,
       int K, NRef nr) {
  pdl_wait();
  using F = FT<S>;
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#if PBF_XSPH_PV
  const float4 pi = pv[i].p;
  const float4 vi = pv[i].v;
  auto fetch = [&](uint32_t j, float4& a, float4& b) {
    const PosVel r = ld_posvel(pv + j);
    a = r.p;
    b = r.v;
  };
#else
  const float4 pi = pos[i];
  const float4 vi = vel_in[i];
  auto fetch = [&](uint32_t j, float4& a, float4& b) {
    a = pos[j];
    b = vel_in[j];
  };
#endif
  XsphPass<S> acc(pi, vi, c);
  for_each_pair2(nbr_idx, K, i, nbr_count[i], fetch, acc);
  if (dbg.dv) dbg.dv[i] = make_float4(acc.sx, acc.sy, acc.sz, 0.0f);
  const V3<F> v = acc.finish(vi);
  if (is_final) {
    finalize_particle<F>(pi, v, __float_as_uint(pos_s[i].w), c, planes, pos_o, vel_o);
  } else {
    const float4 vo = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), vi.w);
    vel_out[i] = vo;
    halo.put(i, vo);
  }
}

// ---------------------------------------------------------------- a13 vorticity, pass 1
template <bool S>
__global__ void __launch_bounds__(kBlock, PBF_SOLVE_MINBLOCKS)
k_vort_omega(float4* __restrict__ pos, const float4* __restrict__ vel, float4* __restrict__ omega,
             const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count, StepConsts c,
             const StatusBlock* st, int K, NRef nr) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pos[i];
  const float4 vi = vel[i];
  auto fetch = [&](uint32_t j, float4& a, float4& b) {
    a = pos[j];
    b = vel[j];
  };
  OmegaPass<S> acc(pi, vi, c);
  for_each_pair2(nbr_idx, K, i, nbr_count[i], fetch, acc);
  const float4 om = acc.finish();
  omega[i] = om;
  // |omega_i| rides in pos[i].w for the eta pass; readers of pos[] use .xyz only here
  reinterpret_cast<float*>(pos + i)[3] = om.w;
}

// ---------------------------------------------------------------- a13 pass 2 + apply (+ a14)
template <bool S>
__global__ void __launch_bounds__(kBlock, PBF_SOLVE_MINBLOCKS)
k_vort_apply(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ omega,
             const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count,
             const float4* __restrict__ pos_s, const float4* __restrict__ planes,
             float4* __restrict__ pos_o, float4* __restrict__ vel_o, StepConsts c, const StatusBlock* st,
             DebugPtrs dbg, int K, NRef nr) {
  pdl_wait();
  using F = FT<S>;
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pos[i];
  EtaPass<S> acc(pi, c);
  for_each_pair(nbr_idx, K, i, nbr_count[i], pos, acc);
  if (dbg.eta) dbg.eta[i] = make_float4(acc.ex, acc.ey, acc.ez, 0.0f);
  const V3<F> v = acc.finish(omega[i], vel[i]);
  finalize_particle<F>(pi, v, __float_as_uint(pos_s[i].w), c, planes, pos_o, vel_o);
}

// solver_iterations == 0: core.cpp:277 never runs, pred is committed as predicted.
template <bool S>
__global__ void __launch_bounds__(kBlock, PBF_SOLVE_MINBLOCKS)
k_commit_only(const float4* __restrict__ pred, const float4* __restrict__ pos_s,
              const float4* __restrict__ planes, float4* __restrict__ pos_o, float4* __restrict__ vel_o,
              StepConsts c, const StatusBlock* st, NRef nr) {
  pdl_wait();
  using F = FT<S>;
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 np = pred[i];
  const float4 p0 = pos_s[i];
  V3<F> v;
  v.x = Arith<F>::div_dt(F(np.x) - F(p0.x), c.dt, c.inv_dt);
  v.y = Arith<F>::div_dt(F(np.y) - F(p0.y), c.dt, c.inv_dt);
  v.z = Arith<F>::div_dt(F(np.z) - F(p0.z), c.dt, c.inv_dt);
  finalize_particle<F>(np, v, __float_as_uint(p0.w), c, planes, pos_o, vel_o);
}

}  // namespace

// ---- per-pass launchers (the slab driver interleaves them with halo exchanges) ----------------
static inline int blocks_for(NRef n) { return (n.n + kBlock - 1) / kBlock; }

// Where the velocity update leaves (pos, vel, m/rho) for the pass after it: the 32-byte records
// when that pass is XSPH, vel[0] otherwise (vorticity without XSPH).
PosVel* xsph_record(const SolveBuffers& b, const StepConsts& c) {
#if PBF_XSPH_PV
  return c.do_xsph ? b.pv : nullptr;  // (the brick family stages pos and vel[0] as two tiles instead)
#else
  return nullptr;
#endif
}

// The specialised (COMMON) delta kernels apply: see k_delta.  PBF_SOLVE_COMMON=0 builds without them.
#ifndef PBF_SOLVE_COMMON
#define PBF_SOLVE_COMMON 1
#endif
static inline bool common_case(const StepConsts& c) {
  return PBF_SOLVE_COMMON && c.sqrt_safe != 0 && (c.scorr_on == 0 || c.scorr_n == 4);
}

int launch_lambda(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, NRef n,
                  bool strict, cudaStream_t s, Span span) {
  if (nl.bricks) return launch_lambda_brick(b, nl, c, cur, strict, s);
  // n.n bounds the thread count; with a span it is the caller's bound for that part
  if (strict)
    PBF_LAUNCH(k_lambda<true>, blocks_for(n), kBlock, s, b.pred[cur], nl.idx, nl.count, b.rho, c, b.status, b.dbg, span,
               nl.K, n);
  else
    PBF_LAUNCH(k_lambda<false>, blocks_for(n), kBlock, s, b.pred[cur], nl.idx, nl.count, b.rho, c, b.status, b.dbg, span,
               nl.K, n);
  return 1;
}

template <bool S, bool COMMON>
static void delta_impl(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool last,
                       bool is_final, NRef n, cudaStream_t s) {
  if (last)
    PBF_LAUNCH((k_delta<S, true, COMMON>), blocks_for(n), kBlock, s, b.pred[cur], b.pred[cur ^ 1], nl.idx, nl.count, b.pos_s,
               b.rho, b.vel[0], xsph_record(b, c), b.planes, b.pos_o, b.vel_o, c, b.status, b.dbg, b.halo,
               is_final ? 1 : 0, nl.K, n);
  else
    PBF_LAUNCH((k_delta<S, false, COMMON>), blocks_for(n), kBlock, s, b.pred[cur], b.pred[cur ^ 1], nl.idx, nl.count, b.pos_s,
               b.rho, b.vel[0], (PosVel*)nullptr, b.planes, b.pos_o, b.vel_o, c, b.status, b.dbg, b.halo, 0, nl.K, n);
}

int launch_delta(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool last,
                 bool is_final, NRef n, bool strict, cudaStream_t s) {
  if (nl.bricks) return launch_delta_brick(b, nl, c, cur, last, is_final, strict, s);
  if (strict && common_case(c)) delta_impl<true, true>(b, nl, c, cur, last, is_final, n, s);
  else if (strict) delta_impl<true, false>(b, nl, c, cur, last, is_final, n, s);
  else delta_impl<false, false>(b, nl, c, cur, last, is_final, n, s);
  return 1;
}

int launch_xsph(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, bool is_final,
                NRef n, bool strict, cudaStream_t s) {
  if (nl.bricks) return launch_xsph_brick(b, nl, c, pos, is_final, strict, s);
  if (strict)
    PBF_LAUNCH(k_xsph<true>, blocks_for(n), kBlock, s, pos, b.vel[0], b.pv, b.vel[1], nl.idx, nl.count, b.pos_s, b.planes,
                                                 b.pos_o, b.vel_o, c, b.status, b.dbg, b.halo, is_final ? 1 : 0, nl.K, n);
  else
    PBF_LAUNCH(k_xsph<false>, blocks_for(n), kBlock, s, pos, b.vel[0], b.pv, b.vel[1], nl.idx, nl.count, b.pos_s, b.planes,
                                                  b.pos_o, b.vel_o, c, b.status, b.dbg, b.halo, is_final ? 1 : 0, nl.K, n);
  return 1;
}

int launch_vort_omega(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                      NRef n, bool strict, cudaStream_t s) {
  if (nl.bricks) return launch_vort_omega_brick(b, nl, c, pos, vcur, strict, s);
  if (strict)
    PBF_LAUNCH(k_vort_omega<true>, blocks_for(n), kBlock, s, pos, b.vel[vcur], b.omega, nl.idx, nl.count, c, b.status, nl.K, n);
  else
    PBF_LAUNCH(k_vort_omega<false>, blocks_for(n), kBlock, s, pos, b.vel[vcur], b.omega, nl.idx, nl.count, c, b.status, nl.K, n);
  return 1;
}

int launch_vort_apply(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                      NRef n, bool strict, cudaStream_t s) {
  if (nl.bricks) return launch_vort_apply_brick(b, nl, c, pos, vcur, strict, s);
  if (strict)
    PBF_LAUNCH(k_vort_apply<true>, blocks_for(n), kBlock, s, pos, b.vel[vcur], b.omega, nl.idx, nl.count, b.pos_s, b.planes,
                                                       b.pos_o, b.vel_o, c, b.status, b.dbg, nl.K, n);
  else
    PBF_LAUNCH(k_vort_apply<false>, blocks_for(n), kBlock, s, pos, b.vel[vcur], b.omega, nl.idx, nl.count, b.pos_s, b.planes,
                                                        b.pos_o, b.vel_o, c, b.status, b.dbg, nl.K, n);
  return 1;
}

int launch_commit_only(const SolveBuffers& b, const StepConsts& c, NRef n, bool strict, cudaStream_t s) {
  // rho/lambda are never computed when solver_iterations == 0 (the reference reads its stale
  // scratch); XSPH would need rho, so only the plain commit is supported.
  StepConsts c0 = c;
  c0.do_xsph = 0;
  c0.do_vort = 0;
  if (strict)
    PBF_LAUNCH(k_commit_only<true>, blocks_for(n), kBlock, s, b.pred[0], b.pos_s, b.planes, b.pos_o, b.vel_o, c0, b.status, n);
  else
    PBF_LAUNCH(k_commit_only<false>, blocks_for(n), kBlock, s, b.pred[0], b.pos_s, b.planes, b.pos_o, b.vel_o, c0, b.status, n);
  return 1;
}

// a8..a14 of one substep on a single GPU.  phase 0 = everything; 1 = the iteration loop only (after
// it the positions of the substep are final: pred[iterations & 1], sorted order); 2 = the tail only
// (XSPH, vorticity, restitution + scatter) — the cuda_step contract path downloads the positions
// while the tail still runs (pbf_capi.cu).
int launch_solve(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int iterations, NRef n,
                 bool strict, cudaStream_t s, StageCallback cb, void* user, int phase) {
  if (iterations <= 0) return phase == 2 ? 0 : launch_commit_only(b, c, n, strict, s);
  int launches = 0;
  int cur = 0;
  auto stage = [&](int id, int begin) { if (cb) cb(user, id, begin); };
  const bool tail_xsph = c.do_xsph != 0, tail_vort = c.do_vort != 0;
  const bool final_in_delta = !tail_xsph && !tail_vort;
  if (phase != 2) {
    for (int it = 0; it < iterations; ++it) {
      const bool last = (it == iterations - 1);
      stage(4, 1);
      launches += launch_lambda(b, nl, c, cur, n, strict, s);
      stage(4, 0);
      stage(5, 1);
      launches += launch_delta(b, nl, c, cur, last, last && final_in_delta, n, strict, s);
      stage(5, 0);
      cur ^= 1;
    }
  } else {
    cur = iterations & 1;
  }
  if (phase == 1) return launches;
  float4* pos = b.pred[cur];  // committed positions, sorted order
  int vcur = 0;
  if (tail_xsph) {
    stage(6, 1);
    launches += launch_xsph(b, nl, c, pos, !tail_vort, n, strict, s);
    stage(6, 0);
    vcur = 1;
  }
  if (tail_vort) {
    stage(7, 1);
    launches += launch_vort_omega(b, nl, c, pos, vcur, n, strict, s);
    stage(7, 0);
    stage(8, 1);
    launches += launch_vort_apply(b, nl, c, pos, vcur, n, strict, s);
    stage(8, 0);
  }
  return launches;
}

}  // namespace pbf

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
#include <type_traits>

#include "pbf_kernels.h"

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

template <bool S> using FT = typename std::conditional<S, sfloat, float>::type;
template <typename F> struct V3 { F x, y, z; };

// ---- pair iteration -----------------------------------------------------------------
// Walks the list of sorted slot i.  Full pairs run without any validity logic; an odd count
// ends with one half-valid pair.  body(a0, a1, v1): data of the two neighbours and validity of the second.
// List loads are streaming (ld.global.cs): every entry is used once per pass and must not
// evict the gathered particle arrays from L1/L2.
template <typename Body>
__device__ __forceinline__ void for_each_pair(const uint32_t* __restrict__ nbr_idx, int K, int i,
                                              uint32_t cnt, const float4* __restrict__ a4, Body body) {
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
                                               uint32_t cnt, Fetch fetch, Body body) {
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

// ---- per-neighbour geometry -----------------------------------------------------------
// Instruction-count driven layout (the passes are issue-bound): the x and y
// components of ONE neighbour are packed into an f32x2 (they already sit in an aligned register
// pair after the 16-byte gather, so no moves are needed), z stays scalar; the scalar chains that
// depend only on r2 (poly6, sqrt, spiky) are then evaluated 2-wide over TWO neighbours.
template <bool S> struct A1;
template <> struct A1<true> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};
template <> struct A1<false> {
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
};

struct NGeom {
  f2 dxy;    // (xi - xj, yi - yj)
  float dz;  // zi - zj
  float r2;  // (dx*dx + dy*dy) + dz*dz (core.cpp:299)
};

template <bool S>
__device__ __forceinline__ NGeom ngeom(f2 pxy, float pz, float4 a) {
  using M = M2<S>;
  using A = A1<S>;
  NGeom g;
  g.dxy = M::sub(pxy, make_float2(a.x, a.y));
  g.dz = A::sub(pz, a.z);
  const f2 sq = M::mul(g.dxy, g.dxy);
  g.r2 = A::add(A::add(sq.x, sq.y), A::mul(g.dz, g.dz));
  return g;
}

// poly6_kernel (core.cpp:35-46) on two r2 values: coeff * ((diff*diff)*diff), 0 if r2 > h2.
// CLAMP: the "0 if r2 > h2" branch as max(diff, 0) — coeff * 0 is the same +0 the branch returns
// (and r2 == h2 gives coeff * 0 in the reference as well).  Callers that only use the value under
// r2 < h2 skip the clamp.
template <bool S, bool CLAMP>
__device__ __forceinline__ f2 poly6_2(f2 r2, const StepConsts& c) {
  using M = M2<S>;
  f2 diff = M::sub(bcast(c.h2), r2);
  if (CLAMP) diff = make_float2(fmaxf(diff.x, 0.0f), fmaxf(diff.y, 0.0f));
  return M::mul(bcast(c.poly6_coeff), M::mul(M::mul(diff, diff), diff));
}

// spiky_gradient_factor(sqrt(max(r2, min_r2))) (core.cpp:48-57, 303-304) on two r2 values:
// (coeff * diff) * diff with diff = h - r.  Only used under r2 < h2; the reference's "0 if r > h"
// can then only trigger through rounding at r2 ~ h2, and max(diff, 0) reproduces it up to the
// sign of a zero that is added to an accumulator (x + -0 == x + +0 for every x but -0, and the
// accumulators start at +0 and can never become -0).
// SAFE: the host guarantees c.sqrt_safe (the launcher picked the specialised kernel), so the
// choice between the two sqrt paths is not re-made for every neighbour pair.
template <bool S, bool SAFE = false>
__device__ __forceinline__ f2 spiky_2(f2 r2, const StepConsts& c) {
  using M = M2<S>;
  const f2 rc = make_float2(fmaxf(r2.x, c.min_r2), fmaxf(r2.y, c.min_r2));
  const f2 r = M::sqrt(rc, SAFE || c.sqrt_safe != 0);
  f2 diff = M::sub(bcast(c.h), r);
  diff = make_float2(fmaxf(diff.x, 0.0f), fmaxf(diff.y, 0.0f));
  return M::mul(M::mul(bcast(c.spiky_coeff), diff), diff);
}

// pow_ratio_n (core.cpp:59-71)
template <bool S>
__device__ __forceinline__ f2 pow_ratio_2(f2 ratio, int n) {
  using M = M2<S>;
  if (n == 2) return M::mul(ratio, ratio);
  if (n == 3) return M::mul(M::mul(ratio, ratio), ratio);
  if (n == 4) {
    const f2 r2 = M::mul(ratio, ratio);
    return M::mul(r2, r2);
  }
  return make_float2(powf(ratio.x, (float)n), powf(ratio.y, (float)n));  // not bit-pinned: no shipped scene reaches it
}

// a14 + scatter: restitution/friction on the committed position (core.cpp:579-610), then the
// particle goes back to its original slot (State stays in original order, core.h:121-132).
template <typename F>
__device__ __forceinline__ void finalize_particle(float4 pos, V3<F> v, uint32_t orig, const StepConsts& c,
                                                  const float4* __restrict__ planes,
                                                  float4* __restrict__ pos_o, float4* __restrict__ vel_o) {
  if (c.do_rest) {
    const F px(pos.x), py(pos.y), pz(pos.z);
    for (int p = 0; p < c.nplanes; ++p) {
      const float4 pl = planes[p];
      const F nx(pl.x), ny(pl.y), nz(pl.z), d(pl.w);
      const F sd = nx * px + ny * py + nz * pz - d;
      if (sd <= F(0.0f)) {
        const F vn = nx * v.x + ny * v.y + nz * v.z;
        F vn_new = vn;
        if (vn < F(0.0f)) vn_new = F(-c.restitution) * vn;
        const F tx = v.x - vn * nx, ty = v.y - vn * ny, tz = v.z - vn * nz;
        const F scale(c.one_minus_friction);
        v.x = tx * scale + vn_new * nx;
        v.y = ty * scale + vn_new * ny;
        v.z = tz * scale + vn_new * nz;
      }
    }
  }
  pos_o[orig] = make_float4(pos.x, pos.y, pos.z, 0.0f);
  vel_o[orig] = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), 0.0f);
}

// ---------------------------------------------------------------- a8 lambda
template <bool S>
__global__ void __launch_bounds__(kBlock, PBF_SOLVE_MINBLOCKS)
k_lambda(float4* __restrict__ pred, const uint32_t* __restrict__ nbr_idx,
         const uint32_t* __restrict__ nbr_count, float* __restrict__ rho_out, StepConsts c,
         const StatusBlock* st, DebugPtrs dbg, Span span, int K, NRef nr) {
  pdl_wait();
  using M = M2<S>;
  using F = FT<S>;
  if (batch_failed(st)) return;
  const int i = span.slot(blockIdx.x * blockDim.x + threadIdx.x, nr.get());
  if (i < 0) return;
  const float4 pi = pred[i];
  float rho = 0.0f, gsx = 0.0f, gsy = 0.0f, gsz = 0.0f, sum_grad2 = 0.0f;
  using A = A1<S>;
  const f2 pxy = make_float2(pi.x, pi.y);
  const float neg_scale = -c.grad_scale;
  for_each_pair(nbr_idx, K, i, nbr_count[i], pred, [&](float4 a0, float4 a1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pi.z, a0), g1 = ngeom<S>(pxy, pi.z, a1);
    const f2 r2 = make_float2(g0.r2, g1.r2);
    f2 w = poly6_2<S, true>(r2, c);                        // rho += poly6(r2) (core.cpp:300)
    f2 gf = spiky_2<S>(r2, c);
    // Straight-line code instead of two divergent branches: a neighbour that fails r2 < h2
    // (core.cpp:302) gets grad_factor = 0, so every term it adds below is a +-0 — a no-op on
    // accumulators that start at +0 (they can never hold -0).  Same for the odd tail slot.
    gf.x = (g0.r2 < c.h2) ? gf.x : 0.0f;
    gf.y = (v1 && g1.r2 < c.h2) ? gf.y : 0.0f;
    if (!v1) w.y = 0.0f;
    {
      const f2 gxy = M::mul(g0.dxy, bcast(gf.x));
      const float gz = A::mul(gf.x, g0.dz);
      const f2 jxy = M::mul(bcast(neg_scale), gxy);
      const float jz = A::mul(neg_scale, gz);
      const f2 jj = M::mul(jxy, jxy);
      const float t = A::add(A::add(jj.x, jj.y), A::mul(jz, jz));
      rho = M::adds(rho, w.x);
      gsx = M::adds(gsx, gxy.x);
      gsy = M::adds(gsy, gxy.y);
      gsz = M::adds(gsz, gz);
      sum_grad2 = M::adds(sum_grad2, t);
    }
    {
      const f2 gxy = M::mul(g1.dxy, bcast(gf.y));
      const float gz = A::mul(gf.y, g1.dz);
      const f2 jxy = M::mul(bcast(neg_scale), gxy);
      const float jz = A::mul(neg_scale, gz);
      const f2 jj = M::mul(jxy, jxy);
      const float t = A::add(A::add(jj.x, jj.y), A::mul(jz, jz));
      rho = M::adds(rho, w.y);
      gsx = M::adds(gsx, gxy.x);
      gsy = M::adds(gsy, gxy.y);
      gsz = M::adds(gsz, gz);
      sum_grad2 = M::adds(sum_grad2, t);
    }
  });
  // core.cpp:319-328
  F rho_f(rho), sg(sum_grad2);
  rho_f += F(c.poly6_zero);
  rho_f *= F(c.mass);
  const F C = rho_f * F(c.inv_density) - F(1.0f);
  const F grad_scale(c.grad_scale);
  const F ix = grad_scale * F(gsx), iy = grad_scale * F(gsy), iz = grad_scale * F(gsz);
  sg += ix * ix + iy * iy + iz * iz;
  const F lambda = -C / (sg + F(c.epsilon));
  // only .w is written; concurrent readers of pred[i] use .xyz only in this pass
  reinterpret_cast<float*>(pred + i)[3] = Arith<F>::val(lambda);
  rho_out[i] = Arith<F>::val(rho_f);
  if (dbg.lambda) dbg.lambda[i] = Arith<F>::val(lambda);
  if (dbg.rho) dbg.rho[i] = Arith<F>::val(rho_f);
}

// ---------------------------------------------------------------- a9 + a10 (+ a11, a14)
// LAST: also velocity update / commit; is_final: additionally restitution + scatter.
// COMMON: specialised for what every shipped scene uses — min_r2 and h2 inside the fast-path range
// of the 2-wide sqrt, and s_corr off or with exponent 4 (core.h:33) — so that neither the sqrt path
// nor the exponent is selected per neighbour pair and the powf fallback of pow_ratio_2 is not part
// of the loop body (664 instead of 1464 instructions).  Same arithmetic; the launcher decides.
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
  using M = M2<S>;
  using F = FT<S>;
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pred_in[i];
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  using A = A1<S>;
  const f2 pxy = make_float2(pi.x, pi.y);
  for_each_pair(nbr_idx, K, i, nbr_count[i], pred_in, [&](float4 a0, float4 a1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pi.z, a0), g1 = ngeom<S>(pxy, pi.z, a1);
    const f2 r2 = make_float2(g0.r2, g1.r2);
    const f2 gf = spiky_2<S, COMMON>(r2, c);
    f2 s = make_float2(A::add(pi.w, a0.w), A::add(pi.w, a1.w));  // lambda_i + lambda_j (core.cpp:355)
    if (c.scorr_on) {                                      // core.cpp:356-361
      const f2 w = poly6_2<S, false>(r2, c);
      const f2 ratio = M::mul(w, bcast(c.scorr_inv_wdq));
      const f2 corr = M::mul(bcast(c.scorr_negk), pow_ratio_2<S>(ratio, COMMON ? 4 : c.scorr_n));
      s = M::addp(s, corr);
    }
    f2 sg = M::mul(s, gf);                                 // (s * grad_factor) * d (core.cpp:362-364)
    sg.x = (g0.r2 < c.h2) ? sg.x : 0.0f;                   // outside h: the terms below are +-0 (no-ops)
    sg.y = (v1 && g1.r2 < c.h2) ? sg.y : 0.0f;
    const f2 t0 = M::mul(g0.dxy, bcast(sg.x)), t1 = M::mul(g1.dxy, bcast(sg.y));
    sx = M::adds(sx, t0.x);
    sy = M::adds(sy, t0.y);
    sz = M::adds(sz, A::mul(sg.x, g0.dz));
    sx = M::adds(sx, t1.x);
    sy = M::adds(sy, t1.y);
    sz = M::adds(sz, A::mul(sg.y, g1.dz));
  });
  const F xi(pi.x), yi(pi.y), zi(pi.z);
  F ax(sx), ay(sy), az(sz);
  ax *= F(c.inv_density);
  ay *= F(c.inv_density);
  az *= F(c.inv_density);
  if (c.nplanes > 0) {  // sequential projection, each plane sees the previous push (core.cpp:372-393)
    F qx = xi + ax, qy = yi + ay, qz = zi + az;
    for (int p = 0; p < c.nplanes; ++p) {
      const float4 pl = planes[p];
      const F nx(pl.x), ny(pl.y), nz(pl.z), d(pl.w);
      const F sd = nx * qx + ny * qy + nz * qz - d;
      const F pen = -sd;
      if (pen > F(0.0f)) {
        qx += nx * pen;
        qy += ny * pen;
        qz += nz * pen;
      }
    }
    ax = qx - xi;
    ay = qy - yi;
    az = qz - zi;
  }
  if (dbg.delta) dbg.delta[i] = make_float4(Arith<F>::val(ax), Arith<F>::val(ay), Arith<F>::val(az), 0.0f);
  // Jacobi apply (core.cpp:403-407): pred += delta, into the other buffer
  const F nx_ = xi + ax, ny_ = yi + ay, nz_ = zi + az;
  const float4 np = make_float4(Arith<F>::val(nx_), Arith<F>::val(ny_), Arith<F>::val(nz_), 0.0f);
  pred_out[i] = np;
  halo.put(i, np);
  if (LAST) {  // core.cpp:414-420
    const float4 p0 = pos_s[i];
    V3<F> v;
    v.x = Arith<F>::div_dt(nx_ - F(p0.x), c.dt, c.inv_dt);
    v.y = Arith<F>::div_dt(ny_ - F(p0.y), c.dt, c.inv_dt);
    v.z = Arith<F>::div_dt(nz_ - F(p0.z), c.dt, c.inv_dt);
    if (is_final) {
      finalize_particle<F>(np, v, __float_as_uint(p0.w), c, planes, pos_o, vel_o);
    } else {
      const F r(rho[i]);
      const F inv_rho = (r > F(0.0f)) ? (F(c.mass) / r) : F(0.0f);  // core.cpp:447-448
      const float4 v4 = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), Arith<F>::val(inv_rho));
      if (pv) {  // XSPH follows: it reads position and velocity of a neighbour as one 32-byte record
        pv[i].p = np;
        pv[i].v = v4;
      } else {
        vel_out[i] = v4;
      }
    }
  }
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
  using M = M2<S>;
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
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  using A = A1<S>;
  const f2 pxy = make_float2(pi.x, pi.y), vxy = make_float2(vi.x, vi.y);
  for_each_pair2(nbr_idx, K, i, nbr_count[i], fetch,
                 [&](float4 a0, float4 a1, float4 b0, float4 b1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pi.z, a0), g1 = ngeom<S>(pxy, pi.z, a1);
    f2 w = poly6_2<S, false>(make_float2(g0.r2, g1.r2), c);
    w.x = (g0.r2 < c.h2) ? w.x : 0.0f;                     // outside h: the terms below are +-0 (no-ops)
    w.y = (v1 && g1.r2 < c.h2) ? w.y : 0.0f;
    // ((v_j - v_i) * W) * inv_rho_j (core.cpp:449-451)
    const f2 t0 = M::mul(M::mul(M::sub(make_float2(b0.x, b0.y), vxy), bcast(w.x)), bcast(b0.w));
    const f2 t1 = M::mul(M::mul(M::sub(make_float2(b1.x, b1.y), vxy), bcast(w.y)), bcast(b1.w));
    sx = M::adds(sx, t0.x);
    sy = M::adds(sy, t0.y);
    sz = M::adds(sz, A::mul(A::mul(A::sub(b0.z, vi.z), w.x), b0.w));
    sx = M::adds(sx, t1.x);
    sy = M::adds(sy, t1.y);
    sz = M::adds(sz, A::mul(A::mul(A::sub(b1.z, vi.z), w.y), b1.w));
  });
  if (dbg.dv) dbg.dv[i] = make_float4(sx, sy, sz, 0.0f);
  V3<F> v;  // core.cpp:461-465
  v.x = F(vi.x) + F(c.visc_c) * F(sx);
  v.y = F(vi.y) + F(c.visc_c) * F(sy);
  v.z = F(vi.z) + F(c.visc_c) * F(sz);
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
  using M = M2<S>;
  using F = FT<S>;
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pos[i];
  const float4 vi = vel[i];
  float ox = 0.0f, oy = 0.0f, oz = 0.0f;
  using A = A1<S>;
  const f2 pxy = make_float2(pi.x, pi.y), vxy = make_float2(vi.x, vi.y);
  auto one = [&](const NGeom& g, float gf, float4 b) {     // core.cpp:493-504
    const f2 gxy = M::mul(g.dxy, bcast(gf));
    const float gz = A::mul(gf, g.dz);
    const f2 uxy = M::sub(make_float2(b.x, b.y), vxy);
    const float uz = A::sub(b.z, vi.z);
    const float tx = A::sub(A::mul(uxy.y, gz), A::mul(uz, gxy.y));   // core.cpp:499-501
    const float ty = A::sub(A::mul(uz, gxy.x), A::mul(uxy.x, gz));
    const float tz = A::sub(A::mul(uxy.x, gxy.y), A::mul(uxy.y, gxy.x));
    ox = M::adds(ox, tx);
    oy = M::adds(oy, ty);
    oz = M::adds(oz, tz);
  };
  auto fetch = [&](uint32_t j, float4& a, float4& b) {
    a = pos[j];
    b = vel[j];
  };
  for_each_pair2(nbr_idx, K, i, nbr_count[i], fetch,
                 [&](float4 a0, float4 a1, float4 b0, float4 b1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pi.z, a0), g1 = ngeom<S>(pxy, pi.z, a1);
    f2 gf = spiky_2<S>(make_float2(g0.r2, g1.r2), c);
    gf.x = (g0.r2 < c.h2) ? gf.x : 0.0f;                   // outside h: every term is +-0 (a no-op)
    gf.y = (v1 && g1.r2 < c.h2) ? gf.y : 0.0f;
    one(g0, gf.x, b0);
    one(g1, gf.y, b1);
  });
  const F fx(ox), fy(oy), fz(oz);
  float m = __fsqrt_rn(Arith<F>::val(fx * fx + fy * fy + fz * fz));  // core.cpp:507
  omega[i] = make_float4(ox, oy, oz, m);
  // |omega_i| rides in pos[i].w for the eta pass; readers of pos[] use .xyz only here
  reinterpret_cast<float*>(pos + i)[3] = m;
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
  using M = M2<S>;
  using F = FT<S>;
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pos[i];
  float ex = 0.0f, ey = 0.0f, ez = 0.0f;
  using A = A1<S>;
  const f2 pxy = make_float2(pi.x, pi.y);
  auto one = [&](const NGeom& g, float gf, float wj) {     // core.cpp:528-538
    const f2 gxy = M::mul(g.dxy, bcast(gf));
    const float gz = A::mul(gf, g.dz);
    const float coeff = A::sub(wj, pi.w);                  // |omega_j| - |omega_i| (core.cpp:534)
    const f2 txy = M::mul(bcast(coeff), gxy);
    ex = M::adds(ex, txy.x);
    ey = M::adds(ey, txy.y);
    ez = M::adds(ez, A::mul(coeff, gz));
  };
  for_each_pair(nbr_idx, K, i, nbr_count[i], pos, [&](float4 a0, float4 a1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pi.z, a0), g1 = ngeom<S>(pxy, pi.z, a1);
    f2 gf = spiky_2<S>(make_float2(g0.r2, g1.r2), c);
    gf.x = (g0.r2 < c.h2) ? gf.x : 0.0f;                   // outside h: every term is +-0 (a no-op)
    gf.y = (v1 && g1.r2 < c.h2) ? gf.y : 0.0f;
    one(g0, gf.x, a0.w);
    one(g1, gf.y, a1.w);
  });
  if (dbg.eta) dbg.eta[i] = make_float4(ex, ey, ez, 0.0f);
  // core.cpp:547-570
  const F fex(ex), fey(ey), fez(ez);
  const F len(__fsqrt_rn(Arith<F>::val(fex * fex + fey * fey + fez * fez)));
  F nx(0.0f), ny(0.0f), nz(0.0f);
  if (len > F(c.vort_norm_eps)) {
    const F inv = F(1.0f) / len;
    nx = fex * inv;
    ny = fey * inv;
    nz = fez * inv;
  }
  const float4 om = omega[i];
  const F ox(om.x), oy(om.y), oz(om.z);
  const F fx = F(c.vort_eps) * (ny * oz - nz * oy);
  const F fy = F(c.vort_eps) * (nz * ox - nx * oz);
  const F fz = F(c.vort_eps) * (nx * oy - ny * ox);
  const float4 vi = vel[i];
  V3<F> v;
  v.x = F(vi.x) + F(c.dt) * fx;
  v.y = F(vi.y) + F(c.dt) * fy;
  v.z = F(vi.z) + F(c.dt) * fz;
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
  return c.do_xsph ? b.pv : nullptr;
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
  if (strict && common_case(c)) delta_impl<true, true>(b, nl, c, cur, last, is_final, n, s);
  else if (strict) delta_impl<true, false>(b, nl, c, cur, last, is_final, n, s);
  else delta_impl<false, false>(b, nl, c, cur, last, is_final, n, s);
  return 1;
}

int launch_xsph(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, bool is_final,
                NRef n, bool strict, cudaStream_t s) {
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
  if (strict)
    PBF_LAUNCH(k_vort_omega<true>, blocks_for(n), kBlock, s, pos, b.vel[vcur], b.omega, nl.idx, nl.count, c, b.status, nl.K, n);
  else
    PBF_LAUNCH(k_vort_omega<false>, blocks_for(n), kBlock, s, pos, b.vel[vcur], b.omega, nl.idx, nl.count, c, b.status, nl.K, n);
  return 1;
}

int launch_vort_apply(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                      NRef n, bool strict, cudaStream_t s) {
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

// a8..a14 of one substep on a single GPU.
int launch_solve(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int iterations, NRef n,
                 bool strict, cudaStream_t s, StageCallback cb, void* user) {
  if (iterations <= 0) return launch_commit_only(b, c, n, strict, s);
  int launches = 0;
  int cur = 0;
  auto stage = [&](int id, int begin) { if (cb) cb(user, id, begin); };
  const bool tail_xsph = c.do_xsph != 0, tail_vort = c.do_vort != 0;
  const bool final_in_delta = !tail_xsph && !tail_vort;
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

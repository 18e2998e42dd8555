// solve_passes.cuh — the per-particle bodies of the constraint-solve and post-solve passes, shared
// by the two kernel families that run them:
//   * solve.cu  — one thread per sorted slot, neighbours gathered from global memory through L1
//                 (32-bit list entries = sorted slots); every configuration, slabs, sparse table;
//   * brick.cu  — one CTA per brick of grid cells, the brick's halo staged into shared memory by
//                 cp.async.bulk (TMA), neighbours gathered with LDS.128 (16-bit tile-relative entries).
// A pass is a functor that consumes neighbour PAIRS (two neighbours per call, Blackwell packed
// f32x2 arithmetic) in list order, plus a finish() that reproduces the reference's epilogue:
//   a8  lambda                      (reference core/src/core.cpp:281-329)
//   a9  delta-p + s_corr + planes   (core.cpp:334-398), a10 apply (core.cpp:400-407),
//   a11 velocity update + commit    (core.cpp:410-421)
//   a12 XSPH viscosity              (core.cpp:423-466)
//   a13 vorticity confinement       (core.cpp:468-571)
//   a14 plane restitution/friction  (core.cpp:573-612)
// STRICT = true: every operation is a correctly rounded IEEE binary32 op in the reference's
// expression order (bit-identical to the CPU path).  STRICT = false: FMA contraction and
// x*rsqrt(x); tolerance-gated.
#pragma once

#include <type_traits>

#include "pbf_device.cuh"

namespace pbf {

template <bool S> using FT = typename std::conditional<S, sfloat, float>::type;
template <typename F> struct V3 { F x, y, z; };

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
// Neighbour data: (pred.xyz, -).
// SAFE: the launcher guarantees c.sqrt_safe (see DeltaPass::COMMON).
template <bool S, bool SAFE = false>
struct LambdaPass {
  using M = M2<S>;
  using A = A1<S>;
  using F = FT<S>;
  const StepConsts& c;
  f2 pxy;
  float pz, neg_scale;
  float rho = 0.0f, gsx = 0.0f, gsy = 0.0f, gsz = 0.0f, sum_grad2 = 0.0f;
  __device__ __forceinline__ LambdaPass(float4 pi, const StepConsts& c_)
      : c(c_), pxy(make_float2(pi.x, pi.y)), pz(pi.z), neg_scale(-c_.grad_scale) {}
  __device__ __forceinline__ void one(const NGeom& g, float gf, float w) {
    const f2 gxy = M::mul(g.dxy, bcast(gf));
    const float gz = A::mul(gf, g.dz);
    const f2 jxy = M::mul(bcast(neg_scale), gxy);
    const float jz = A::mul(neg_scale, gz);
    const f2 jj = M::mul(jxy, jxy);
    const float t = A::add(A::add(jj.x, jj.y), A::mul(jz, jz));
    rho = M::adds(rho, w);
    gsx = M::adds(gsx, gxy.x);
    gsy = M::adds(gsy, gxy.y);
    gsz = M::adds(gsz, gz);
    sum_grad2 = M::adds(sum_grad2, t);
  }
  __device__ __forceinline__ void operator()(float4 a0, float4 a1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pz, a0), g1 = ngeom<S>(pxy, pz, a1);
    const f2 r2 = make_float2(g0.r2, g1.r2);
    f2 w = poly6_2<S, true>(r2, c);                        // rho += poly6(r2) (core.cpp:300)
    f2 gf = spiky_2<S, SAFE>(r2, c);
    // Straight-line code instead of two divergent branches: a neighbour that fails r2 < h2
    // (core.cpp:302) gets grad_factor = 0, so every term it adds below is a +-0 — a no-op on
    // accumulators that start at +0 (they can never hold -0).  Same for the odd tail slot.
    gf.x = (g0.r2 < c.h2) ? gf.x : 0.0f;
    gf.y = (v1 && g1.r2 < c.h2) ? gf.y : 0.0f;
    if (!v1) w.y = 0.0f;
    one(g0, gf.x, w.x);
    one(g1, gf.y, w.y);
  }
  // core.cpp:319-328
  __device__ __forceinline__ void finish(float& lambda, float& rho_out) const {
    F rho_f(rho), sg(sum_grad2);
    rho_f += F(c.poly6_zero);
    rho_f *= F(c.mass);
    const F C = rho_f * F(c.inv_density) - F(1.0f);
    const F grad_scale(c.grad_scale);
    const F ix = grad_scale * F(gsx), iy = grad_scale * F(gsy), iz = grad_scale * F(gsz);
    sg += ix * ix + iy * iy + iz * iz;
    const F l = -C / (sg + F(c.epsilon));
    lambda = Arith<F>::val(l);
    rho_out = Arith<F>::val(rho_f);
  }
};

// ---------------------------------------------------------------- a9 (+ a10, a11, a14)
// Neighbour data: (pred.xyz, lambda_j).
// COMMON: specialised for what every shipped scene uses — min_r2 and h2 inside the fast-path range
// of the 2-wide sqrt, and s_corr off or with exponent 4 (core.h:33) — so that neither the sqrt path
// nor the exponent is selected per neighbour pair and the powf fallback of pow_ratio_2 is not part
// of the loop body (664 instead of 1464 instructions).  Same arithmetic; the launcher decides.
template <bool S, bool COMMON>
struct DeltaPass {
  using M = M2<S>;
  using A = A1<S>;
  using F = FT<S>;
  const StepConsts& c;
  f2 pxy;
  float pz, li;
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  __device__ __forceinline__ DeltaPass(float4 pi, const StepConsts& c_)
      : c(c_), pxy(make_float2(pi.x, pi.y)), pz(pi.z), li(pi.w) {}
  __device__ __forceinline__ void operator()(float4 a0, float4 a1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pz, a0), g1 = ngeom<S>(pxy, pz, a1);
    const f2 r2 = make_float2(g0.r2, g1.r2);
    const f2 gf = spiky_2<S, COMMON>(r2, c);
    f2 s = make_float2(A::add(li, a0.w), A::add(li, a1.w));  // lambda_i + lambda_j (core.cpp:355)
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
  }
  // scale, sequential plane projection (each plane sees the previous push, core.cpp:372-393),
  // Jacobi apply (core.cpp:403-407).  Returns the new predicted position; delta through `d`.
  __device__ __forceinline__ void finish(float4 pi, const float4* __restrict__ planes, V3<F>& np, float4& d) const {
    const F xi(pi.x), yi(pi.y), zi(pi.z);
    F ax(sx), ay(sy), az(sz);
    ax *= F(c.inv_density);
    ay *= F(c.inv_density);
    az *= F(c.inv_density);
    if (c.nplanes > 0) {
      F qx = xi + ax, qy = yi + ay, qz = zi + az;
      for (int p = 0; p < c.nplanes; ++p) {
        const float4 pl = planes[p];
        const F nx(pl.x), ny(pl.y), nz(pl.z), dd(pl.w);
        const F sd = nx * qx + ny * qy + nz * qz - dd;
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
    d = make_float4(Arith<F>::val(ax), Arith<F>::val(ay), Arith<F>::val(az), 0.0f);
    np.x = xi + ax;
    np.y = yi + ay;
    np.z = zi + az;
  }
};

// Epilogue of a delta pass for slot i: store the new prediction (+ halo), and on the LAST iteration
// the velocity update / commit (core.cpp:414-420), optionally the final restitution + scatter.
template <bool S, bool LAST>
__device__ __forceinline__ void delta_store(int i, V3<FT<S>> npf, float4 dlt, float4* __restrict__ pred_out,
                                            const float4* __restrict__ pos_s, const float* __restrict__ rho,
                                            float4* __restrict__ vel_out, PosVel* __restrict__ pv,
                                            const float4* __restrict__ planes, float4* __restrict__ pos_o,
                                            float4* __restrict__ vel_o, const StepConsts& c, const DebugPtrs& dbg,
                                            const HaloOut& halo, int is_final) {
  using F = FT<S>;
  if (dbg.delta) dbg.delta[i] = dlt;
  const float4 np = make_float4(Arith<F>::val(npf.x), Arith<F>::val(npf.y), Arith<F>::val(npf.z), 0.0f);
  pred_out[i] = np;
  halo.put(i, np);
  if (LAST) {
    const float4 p0 = pos_s[i];
    V3<F> v;
    v.x = Arith<F>::div_dt(npf.x - F(p0.x), c.dt, c.inv_dt);
    v.y = Arith<F>::div_dt(npf.y - F(p0.y), c.dt, c.inv_dt);
    v.z = Arith<F>::div_dt(npf.z - F(p0.z), c.dt, c.inv_dt);
    if (is_final) {
      finalize_particle<F>(np, v, __float_as_uint(p0.w), c, planes, pos_o, vel_o);
    } else {
      const F r(rho[i]);
      const F inv_rho = (r > F(0.0f)) ? (F(c.mass) / r) : F(0.0f);  // core.cpp:447-448
      const float4 v4 = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), Arith<F>::val(inv_rho));
      if (pv) {  // XSPH (global-gather family) reads position and velocity of a neighbour as one 32-byte record
        pv[i].p = np;
        pv[i].v = v4;
      } else {
        vel_out[i] = v4;
      }
    }
  }
}

// ---------------------------------------------------------------- a12 XSPH
// Neighbour data: (pos.xyz, -) and (vel.xyz, m/rho_j).
template <bool S>
struct XsphPass {
  using M = M2<S>;
  using A = A1<S>;
  using F = FT<S>;
  const StepConsts& c;
  f2 pxy, vxy;
  float pz, vz;
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  __device__ __forceinline__ XsphPass(float4 pi, float4 vi, const StepConsts& c_)
      : c(c_), pxy(make_float2(pi.x, pi.y)), vxy(make_float2(vi.x, vi.y)), pz(pi.z), vz(vi.z) {}
  __device__ __forceinline__ void operator()(float4 a0, float4 a1, float4 b0, float4 b1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pz, a0), g1 = ngeom<S>(pxy, pz, a1);
    f2 w = poly6_2<S, false>(make_float2(g0.r2, g1.r2), c);
    w.x = (g0.r2 < c.h2) ? w.x : 0.0f;                     // outside h: the terms below are +-0 (no-ops)
    w.y = (v1 && g1.r2 < c.h2) ? w.y : 0.0f;
    // ((v_j - v_i) * W) * inv_rho_j (core.cpp:449-451)
    const f2 t0 = M::mul(M::mul(M::sub(make_float2(b0.x, b0.y), vxy), bcast(w.x)), bcast(b0.w));
    const f2 t1 = M::mul(M::mul(M::sub(make_float2(b1.x, b1.y), vxy), bcast(w.y)), bcast(b1.w));
    sx = M::adds(sx, t0.x);
    sy = M::adds(sy, t0.y);
    sz = M::adds(sz, A::mul(A::mul(A::sub(b0.z, vz), w.x), b0.w));
    sx = M::adds(sx, t1.x);
    sy = M::adds(sy, t1.y);
    sz = M::adds(sz, A::mul(A::mul(A::sub(b1.z, vz), w.y), b1.w));
  }
  // core.cpp:461-465
  __device__ __forceinline__ V3<F> finish(float4 vi) const {
    V3<F> v;
    v.x = F(vi.x) + F(c.visc_c) * F(sx);
    v.y = F(vi.y) + F(c.visc_c) * F(sy);
    v.z = F(vi.z) + F(c.visc_c) * F(sz);
    return v;
  }
};

// ---------------------------------------------------------------- a13 vorticity, pass 1
// Neighbour data: (pos.xyz, -) and (vel.xyz, -).
template <bool S>
struct OmegaPass {
  using M = M2<S>;
  using A = A1<S>;
  using F = FT<S>;
  const StepConsts& c;
  f2 pxy, vxy;
  float pz, vz;
  float ox = 0.0f, oy = 0.0f, oz = 0.0f;
  __device__ __forceinline__ OmegaPass(float4 pi, float4 vi, const StepConsts& c_)
      : c(c_), pxy(make_float2(pi.x, pi.y)), vxy(make_float2(vi.x, vi.y)), pz(pi.z), vz(vi.z) {}
  __device__ __forceinline__ void one(const NGeom& g, float gf, float4 b) {     // core.cpp:493-504
    const f2 gxy = M::mul(g.dxy, bcast(gf));
    const float gz = A::mul(gf, g.dz);
    const f2 uxy = M::sub(make_float2(b.x, b.y), vxy);
    const float uz = A::sub(b.z, vz);
    const float tx = A::sub(A::mul(uxy.y, gz), A::mul(uz, gxy.y));   // core.cpp:499-501
    const float ty = A::sub(A::mul(uz, gxy.x), A::mul(uxy.x, gz));
    const float tz = A::sub(A::mul(uxy.x, gxy.y), A::mul(uxy.y, gxy.x));
    ox = M::adds(ox, tx);
    oy = M::adds(oy, ty);
    oz = M::adds(oz, tz);
  }
  __device__ __forceinline__ void operator()(float4 a0, float4 a1, float4 b0, float4 b1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pz, a0), g1 = ngeom<S>(pxy, pz, a1);
    f2 gf = spiky_2<S>(make_float2(g0.r2, g1.r2), c);
    gf.x = (g0.r2 < c.h2) ? gf.x : 0.0f;                   // outside h: every term is +-0 (a no-op)
    gf.y = (v1 && g1.r2 < c.h2) ? gf.y : 0.0f;
    one(g0, gf.x, b0);
    one(g1, gf.y, b1);
  }
  // (omega xyz, |omega|) (core.cpp:507)
  __device__ __forceinline__ float4 finish() const {
    const F fx(ox), fy(oy), fz(oz);
    const float m = __fsqrt_rn(Arith<F>::val(fx * fx + fy * fy + fz * fz));
    return make_float4(ox, oy, oz, m);
  }
};

// ---------------------------------------------------------------- a13 pass 2 + apply
// Neighbour data: (pos.xyz, |omega_j|).
template <bool S>
struct EtaPass {
  using M = M2<S>;
  using A = A1<S>;
  using F = FT<S>;
  const StepConsts& c;
  f2 pxy;
  float pz, wi;
  float ex = 0.0f, ey = 0.0f, ez = 0.0f;
  __device__ __forceinline__ EtaPass(float4 pi, const StepConsts& c_)
      : c(c_), pxy(make_float2(pi.x, pi.y)), pz(pi.z), wi(pi.w) {}
  __device__ __forceinline__ void one(const NGeom& g, float gf, float wj) {     // core.cpp:528-538
    const f2 gxy = M::mul(g.dxy, bcast(gf));
    const float gz = A::mul(gf, g.dz);
    const float coeff = A::sub(wj, wi);                    // |omega_j| - |omega_i| (core.cpp:534)
    const f2 txy = M::mul(bcast(coeff), gxy);
    ex = M::adds(ex, txy.x);
    ey = M::adds(ey, txy.y);
    ez = M::adds(ez, A::mul(coeff, gz));
  }
  __device__ __forceinline__ void operator()(float4 a0, float4 a1, bool v1) {
    const NGeom g0 = ngeom<S>(pxy, pz, a0), g1 = ngeom<S>(pxy, pz, a1);
    f2 gf = spiky_2<S>(make_float2(g0.r2, g1.r2), c);
    gf.x = (g0.r2 < c.h2) ? gf.x : 0.0f;                   // outside h: every term is +-0 (a no-op)
    gf.y = (v1 && g1.r2 < c.h2) ? gf.y : 0.0f;
    one(g0, gf.x, a0.w);
    one(g1, gf.y, a1.w);
  }
  // core.cpp:547-570: v += dt * eps * (N x omega)
  __device__ __forceinline__ V3<F> finish(float4 om, float4 vi) const {
    const F fex(ex), fey(ey), fez(ez);
    const F len(__fsqrt_rn(Arith<F>::val(fex * fex + fey * fey + fez * fez)));
    F nx(0.0f), ny(0.0f), nz(0.0f);
    if (len > F(c.vort_norm_eps)) {
      const F inv = F(1.0f) / len;
      nx = fex * inv;
      ny = fey * inv;
      nz = fez * inv;
    }
    const F ox(om.x), oy(om.y), oz(om.z);
    const F fx = F(c.vort_eps) * (ny * oz - nz * oy);
    const F fy = F(c.vort_eps) * (nz * ox - nx * oz);
    const F fz = F(c.vort_eps) * (nx * oy - ny * ox);
    V3<F> v;
    v.x = F(vi.x) + F(c.dt) * fx;
    v.y = F(vi.y) + F(c.dt) * fy;
    v.z = F(vi.z) + F(c.dt) * fz;
    return v;
  }
};

}  // namespace pbf

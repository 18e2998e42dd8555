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
// re-tested with r2 < h2 on current positions).  The k-th list entry of a warp is one
// coalesced 128-byte line; each neighbour costs one 16-byte gather because everything a pass
// needs from particle j is packed into a single float4:
//   lambda pass   (pred.xyz, -)          delta pass   (pred.xyz, lambda_j)
//   XSPH          (pos.xyz, -) + (vel.xyz, m/rho_j)
//   omega pass    (pos.xyz, -) + (vel.xyz, -)          eta pass   (pos.xyz, |omega_j|)
//
// Templated on the arithmetic type F: sfloat = STRICT (bit-identical to the CPU reference,
// see pbf_device.cuh) or float = FAST.  The code below is written once, in the reference's
// expression order.
#include "pbf_kernels.h"

namespace pbf {

namespace {

constexpr int kBlock = 128;
constexpr int kUnroll = 4;  // list entries fetched per batch (independent gathers in flight)

template <typename F> struct V3 { F x, y, z; };

__device__ __forceinline__ bool batch_failed(const StatusBlock* st) {
  return (st->grid_overflow | st->nbr_overflow) != 0;
}

// poly6_kernel (core.cpp:35-46): 0 if r2 > h2 else coeff * (h2-r2)^3
template <typename F>
__device__ __forceinline__ F poly6(F r2, const StepConsts& c) {
  if (r2 > F(c.h2)) return F(0.0f);
  const F diff = F(c.h2) - r2;
  const F diff3 = diff * diff * diff;
  return F(c.poly6_coeff) * diff3;
}

// spiky_gradient_factor (core.cpp:48-57): 0 if r > h else coeff * (h-r)^2
template <typename F>
__device__ __forceinline__ F spiky(F r, const StepConsts& c) {
  if (r > F(c.h)) return F(0.0f);
  const F diff = F(c.h) - r;
  return F(c.spiky_coeff) * diff * diff;
}

// pow_ratio_n (core.cpp:59-71)
template <typename F>
__device__ __forceinline__ F pow_ratio(F ratio, int n) {
  if (n == 2) return ratio * ratio;
  if (n == 3) return ratio * ratio * ratio;
  if (n == 4) {
    const F r2 = ratio * ratio;
    return r2 * r2;
  }
  return F(powf(Arith<F>::val(ratio), (float)n));  // not bit-pinned: no shipped scene reaches it
}

// clamp + sqrt of core.cpp:303
template <typename F>
__device__ __forceinline__ F clamped_r(F r2, const StepConsts& c) {
  return Arith<F>::sqrt(r2 < F(c.min_r2) ? F(c.min_r2) : r2);
}

// Walks the neighbour list of sorted particle i; body(j, data_j...) is called in list order.
// Gathers for kUnroll entries are issued before any of them is consumed.
template <typename Body>
__device__ __forceinline__ void for_each_neighbor(const uint32_t* __restrict__ nbr_idx, int K, int i,
                                                  uint32_t cnt, const float4* __restrict__ a4, Body body) {
  const uint32_t* row = nbr_idx + (size_t)(i >> 5) * (size_t)K * 32u + (uint32_t)(i & 31);
  for (uint32_t k = 0; k < cnt; k += kUnroll) {
    uint32_t j[kUnroll];
    float4 a[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) j[u] = (k + u < cnt) ? row[(size_t)(k + u) * 32u] : (uint32_t)i;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) a[u] = a4[j[u]];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (k + u < cnt) body(j[u], a[u]);
  }
}

template <typename Body>
__device__ __forceinline__ void for_each_neighbor2(const uint32_t* __restrict__ nbr_idx, int K, int i,
                                                   uint32_t cnt, const float4* __restrict__ a4,
                                                   const float4* __restrict__ b4, Body body) {
  const uint32_t* row = nbr_idx + (size_t)(i >> 5) * (size_t)K * 32u + (uint32_t)(i & 31);
  for (uint32_t k = 0; k < cnt; k += kUnroll) {
    uint32_t j[kUnroll];
    float4 a[kUnroll], b[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) j[u] = (k + u < cnt) ? row[(size_t)(k + u) * 32u] : (uint32_t)i;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      a[u] = a4[j[u]];
      b[u] = b4[j[u]];
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (k + u < cnt) body(j[u], a[u], b[u]);
  }
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
template <typename F>
__global__ void __launch_bounds__(kBlock)
k_lambda(float4* __restrict__ pred, const uint32_t* __restrict__ nbr_idx,
         const uint32_t* __restrict__ nbr_count, float* __restrict__ rho_out, StepConsts c,
         const StatusBlock* st, DebugPtrs dbg, int K, int n) {
  if (batch_failed(st)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pred[i];
  const F xi(pi.x), yi(pi.y), zi(pi.z);
  F rho(0.0f), gsx(0.0f), gsy(0.0f), gsz(0.0f), sum_grad2(0.0f);
  const F grad_scale(c.grad_scale);
  for_each_neighbor(nbr_idx, K, i, nbr_count[i], pred, [&](uint32_t, float4 pj) {
    const F dx = xi - F(pj.x), dy = yi - F(pj.y), dz = zi - F(pj.z);
    const F r2 = dx * dx + dy * dy + dz * dz;
    rho += poly6(r2, c);
    if (r2 < F(c.h2)) {
      const F gf = spiky(clamped_r(r2, c), c);
      const F gx = gf * dx, gy = gf * dy, gz = gf * dz;
      gsx += gx;
      gsy += gy;
      gsz += gz;
      const F jx = -grad_scale * gx, jy = -grad_scale * gy, jz = -grad_scale * gz;
      sum_grad2 += jx * jx + jy * jy + jz * jz;
    }
  });
  rho += F(c.poly6_zero);
  rho *= F(c.mass);
  const F C = rho * F(c.inv_density) - F(1.0f);
  const F ix = grad_scale * gsx, iy = grad_scale * gsy, iz = grad_scale * gsz;
  sum_grad2 += ix * ix + iy * iy + iz * iz;
  const F lambda = -C / (sum_grad2 + F(c.epsilon));
  // only .w is written; concurrent readers of pred[i] use .xyz only in this pass
  reinterpret_cast<float*>(pred + i)[3] = Arith<F>::val(lambda);
  rho_out[i] = Arith<F>::val(rho);
  if (dbg.lambda) dbg.lambda[i] = Arith<F>::val(lambda);
  if (dbg.rho) dbg.rho[i] = Arith<F>::val(rho);
}

// ---------------------------------------------------------------- a9 + a10 (+ a11, a14)
// LAST: also velocity update / commit; FINAL: additionally restitution + scatter.
template <typename F, bool LAST>
__global__ void __launch_bounds__(kBlock)
k_delta(const float4* __restrict__ pred_in, float4* __restrict__ pred_out,
        const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count,
        const float4* __restrict__ pos_s, const float* __restrict__ rho, float4* __restrict__ vel_out,
        const float4* __restrict__ planes, float4* __restrict__ pos_o, float4* __restrict__ vel_o,
        StepConsts c, const StatusBlock* st, DebugPtrs dbg, int is_final, int K, int n) {
  if (batch_failed(st)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pred_in[i];
  const F xi(pi.x), yi(pi.y), zi(pi.z), li(pi.w);
  F ax(0.0f), ay(0.0f), az(0.0f);
  for_each_neighbor(nbr_idx, K, i, nbr_count[i], pred_in, [&](uint32_t, float4 pj) {
    const F dx = xi - F(pj.x), dy = yi - F(pj.y), dz = zi - F(pj.z);
    const F r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < F(c.h2)) {
      const F gf = spiky(clamped_r(r2, c), c);
      F s = li + F(pj.w);
      if (c.scorr_on) {
        const F W = poly6(r2, c);
        const F ratio = W * F(c.scorr_inv_wdq);
        const F corr = F(c.scorr_negk) * pow_ratio(ratio, c.scorr_n);
        s += corr;
      }
      ax += s * gf * dx;
      ay += s * gf * dy;
      az += s * gf * dz;
    }
  });
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
      vel_out[i] = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), Arith<F>::val(inv_rho));
    }
  }
}

// ---------------------------------------------------------------- a12 XSPH
template <typename F>
__global__ void __launch_bounds__(kBlock)
k_xsph(const float4* __restrict__ pos, const float4* __restrict__ vel_in, float4* __restrict__ vel_out,
       const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count,
       const float4* __restrict__ pos_s, const float4* __restrict__ planes, float4* __restrict__ pos_o,
       float4* __restrict__ vel_o, StepConsts c, const StatusBlock* st, DebugPtrs dbg, int is_final,
       int K, int n) {
  if (batch_failed(st)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pos[i];
  const float4 vi = vel_in[i];
  const F xi(pi.x), yi(pi.y), zi(pi.z), vx(vi.x), vy(vi.y), vz(vi.z);
  F sx(0.0f), sy(0.0f), sz(0.0f);
  for_each_neighbor2(nbr_idx, K, i, nbr_count[i], pos, vel_in, [&](uint32_t, float4 pj, float4 vj) {
    const F dx = xi - F(pj.x), dy = yi - F(pj.y), dz = zi - F(pj.z);
    const F r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < F(c.h2)) {
      const F W = poly6(r2, c);
      const F inv_rho_j(vj.w);
      sx += (F(vj.x) - vx) * W * inv_rho_j;
      sy += (F(vj.y) - vy) * W * inv_rho_j;
      sz += (F(vj.z) - vz) * W * inv_rho_j;
    }
  });
  if (dbg.dv) dbg.dv[i] = make_float4(Arith<F>::val(sx), Arith<F>::val(sy), Arith<F>::val(sz), 0.0f);
  V3<F> v;  // core.cpp:461-465
  v.x = vx + F(c.visc_c) * sx;
  v.y = vy + F(c.visc_c) * sy;
  v.z = vz + F(c.visc_c) * sz;
  if (is_final) {
    finalize_particle<F>(pi, v, __float_as_uint(pos_s[i].w), c, planes, pos_o, vel_o);
  } else {
    vel_out[i] = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), vi.w);
  }
}

// ---------------------------------------------------------------- a13 vorticity, pass 1
template <typename F>
__global__ void __launch_bounds__(kBlock)
k_vort_omega(float4* __restrict__ pos, const float4* __restrict__ vel, float4* __restrict__ omega,
             const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count, StepConsts c,
             const StatusBlock* st, int K, int n) {
  if (batch_failed(st)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pos[i];
  const float4 vi = vel[i];
  const F xi(pi.x), yi(pi.y), zi(pi.z), vx(vi.x), vy(vi.y), vz(vi.z);
  F ox(0.0f), oy(0.0f), oz(0.0f);
  for_each_neighbor2(nbr_idx, K, i, nbr_count[i], pos, vel, [&](uint32_t, float4 pj, float4 vj) {
    const F dx = xi - F(pj.x), dy = yi - F(pj.y), dz = zi - F(pj.z);
    const F r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < F(c.h2)) {
      const F gf = spiky(clamped_r(r2, c), c);
      const F gx = gf * dx, gy = gf * dy, gz = gf * dz;
      const F ux = F(vj.x) - vx, uy = F(vj.y) - vy, uz = F(vj.z) - vz;
      ox += uy * gz - uz * gy;
      oy += uz * gx - ux * gz;
      oz += ux * gy - uy * gx;
    }
  });
  const F mag = Arith<F>::sqrt(ox * ox + oy * oy + oz * oz);  // core.cpp:507
  float m = Arith<F>::val(mag);
  if (!Arith<F>::strict && !(m == m)) m = 0.0f;  // x*rsqrt(x) at x == 0
  omega[i] = make_float4(Arith<F>::val(ox), Arith<F>::val(oy), Arith<F>::val(oz), m);
  // |omega_i| rides in pos[i].w for the eta pass; readers of pos[] use .xyz only here
  reinterpret_cast<float*>(pos + i)[3] = m;
}

// ---------------------------------------------------------------- a13 pass 2 + apply (+ a14)
template <typename F>
__global__ void __launch_bounds__(kBlock)
k_vort_apply(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ omega,
             const uint32_t* __restrict__ nbr_idx, const uint32_t* __restrict__ nbr_count,
             const float4* __restrict__ pos_s, const float4* __restrict__ planes,
             float4* __restrict__ pos_o, float4* __restrict__ vel_o, StepConsts c, const StatusBlock* st,
             DebugPtrs dbg, int K, int n) {
  if (batch_failed(st)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pi = pos[i];
  const F xi(pi.x), yi(pi.y), zi(pi.z), omi(pi.w);
  F ex(0.0f), ey(0.0f), ez(0.0f);
  for_each_neighbor(nbr_idx, K, i, nbr_count[i], pos, [&](uint32_t, float4 pj) {
    const F dx = xi - F(pj.x), dy = yi - F(pj.y), dz = zi - F(pj.z);
    const F r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < F(c.h2)) {
      const F gf = spiky(clamped_r(r2, c), c);
      const F gx = gf * dx, gy = gf * dy, gz = gf * dz;
      const F coeff = F(pj.w) - omi;
      ex += coeff * gx;
      ey += coeff * gy;
      ez += coeff * gz;
    }
  });
  if (dbg.eta) dbg.eta[i] = make_float4(Arith<F>::val(ex), Arith<F>::val(ey), Arith<F>::val(ez), 0.0f);
  // core.cpp:547-570
  const F len = Arith<F>::sqrt(ex * ex + ey * ey + ez * ez);
  F nx(0.0f), ny(0.0f), nz(0.0f);
  if (len > F(c.vort_norm_eps)) {
    const F inv = F(1.0f) / len;
    nx = ex * inv;
    ny = ey * inv;
    nz = ez * inv;
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

template <typename F>
int solve_impl(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int iterations, int n,
               cudaStream_t s, StageCallback cb, void* user) {
  const int blocks = (n + kBlock - 1) / kBlock;
  int launches = 0;
  int cur = 0;
  auto stage = [&](int id, int begin) { if (cb) cb(user, id, begin); };
  const bool tail_xsph = c.do_xsph != 0, tail_vort = c.do_vort != 0;
  const int final_in_delta = (!tail_xsph && !tail_vort) ? 1 : 0;
  // 0 iterations: the predicted positions are committed unchanged; run the LAST delta variant
  // on an empty list so the velocity update / commit still happens (core.cpp:277 loop skipped).
  for (int it = 0; it < iterations; ++it) {
    const bool last = (it == iterations - 1);
    stage(4, 1);
    k_lambda<F><<<blocks, kBlock, 0, s>>>(b.pred[cur], nl.idx, nl.count, b.rho, c, b.status, b.dbg, nl.K, n);
    stage(4, 0);
    stage(5, 1);
    if (last)
      k_delta<F, true><<<blocks, kBlock, 0, s>>>(b.pred[cur], b.pred[cur ^ 1], nl.idx, nl.count, b.pos_s, b.rho,
                                                 b.vel[0], b.planes, b.pos_o, b.vel_o, c, b.status, b.dbg,
                                                 final_in_delta, nl.K, n);
    else
      k_delta<F, false><<<blocks, kBlock, 0, s>>>(b.pred[cur], b.pred[cur ^ 1], nl.idx, nl.count, b.pos_s, b.rho,
                                                  b.vel[0], b.planes, b.pos_o, b.vel_o, c, b.status, b.dbg, 0,
                                                  nl.K, n);
    stage(5, 0);
    cur ^= 1;
    launches += 2;
  }
  float4* pos = b.pred[cur];  // committed positions, sorted order
  int vcur = 0;
  if (tail_xsph) {
    stage(6, 1);
    k_xsph<F><<<blocks, kBlock, 0, s>>>(pos, b.vel[0], b.vel[1], nl.idx, nl.count, b.pos_s, b.planes, b.pos_o,
                                        b.vel_o, c, b.status, b.dbg, tail_vort ? 0 : 1, nl.K, n);
    stage(6, 0);
    vcur = 1;
    ++launches;
  }
  if (tail_vort) {
    stage(7, 1);
    k_vort_omega<F><<<blocks, kBlock, 0, s>>>(pos, b.vel[vcur], b.omega, nl.idx, nl.count, c, b.status, nl.K, n);
    stage(7, 0);
    stage(8, 1);
    k_vort_apply<F><<<blocks, kBlock, 0, s>>>(pos, b.vel[vcur], b.omega, nl.idx, nl.count, b.pos_s, b.planes,
                                              b.pos_o, b.vel_o, c, b.status, b.dbg, nl.K, n);
    stage(8, 0);
    launches += 2;
  }
  return launches;
}

// solver_iterations == 0: core.cpp:277 never runs, pred is committed as predicted.
template <typename F>
__global__ void __launch_bounds__(kBlock)
k_commit_only(const float4* __restrict__ pred, const float4* __restrict__ pos_s, float4* __restrict__ vel_out,
              const float4* __restrict__ planes, float4* __restrict__ pos_o, float4* __restrict__ vel_o,
              StepConsts c, const StatusBlock* st, int is_final, int n) {
  if (batch_failed(st)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 np = pred[i];
  const float4 p0 = pos_s[i];
  V3<F> v;
  v.x = Arith<F>::div_dt(F(np.x) - F(p0.x), c.dt, c.inv_dt);
  v.y = Arith<F>::div_dt(F(np.y) - F(p0.y), c.dt, c.inv_dt);
  v.z = Arith<F>::div_dt(F(np.z) - F(p0.z), c.dt, c.inv_dt);
  if (is_final)
    finalize_particle<F>(np, v, __float_as_uint(p0.w), c, planes, pos_o, vel_o);
  else
    vel_out[i] = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), 0.0f);
}

}  // namespace

int launch_solve(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int iterations, int n,
                 bool strict, cudaStream_t s, StageCallback cb, void* cb_user) {
  if (iterations <= 0) {
    // rho/lambda are never computed in this case (the reference reads its stale scratch);
    // XSPH would need rho, so only the plain commit is supported.
    const int blocks = (n + kBlock - 1) / kBlock;
    StepConsts c0 = c;
    c0.do_xsph = 0;
    c0.do_vort = 0;
    if (strict)
      k_commit_only<sfloat><<<blocks, kBlock, 0, s>>>(b.pred[0], b.pos_s, b.vel[0], b.planes, b.pos_o, b.vel_o, c0,
                                                      b.status, 1, n);
    else
      k_commit_only<float><<<blocks, kBlock, 0, s>>>(b.pred[0], b.pos_s, b.vel[0], b.planes, b.pos_o, b.vel_o, c0,
                                                     b.status, 1, n);
    return 1;
  }
  return strict ? solve_impl<sfloat>(b, nl, c, iterations, n, s, cb, cb_user)
                : solve_impl<float>(b, nl, c, iterations, n, s, cb, cb_user);
}

}  // namespace pbf

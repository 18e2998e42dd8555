// slab.cu — device side of the x-slab decomposition.
//
// The reference has no multi-GPU path; what pins this file is the single-GPU result: the
// reference's sort order is x-major (cell_key_less, reference core/src/core.cpp:12-21), so the
// particles of the x-cells [cut_lo, cut_hi) are one contiguous range of the global sorted order.
// Everything here preserves two invariants that make the slab result BIT-IDENTICAL to one GPU:
//   (1) owned particles are stored in ascending global id, so the stable radix sort breaks ties
//       inside a cell exactly like the reference (core.cpp:182);
//   (2) ghosts are appended in the sender's sorted order, so ghost cells list their particles in
//       ascending id as well.
// All counts live in device memory (SlabCounts); messages have a fixed capacity and carry their
// element count in a header, so a batch of substeps never synchronises with the host.
#include "pbf_kernels.h"

namespace pbf {

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kInvalidKey = 0xffffffffu;

// Header of a counted message: .x = element count, .y = 1 when the sender's batch has already
// failed (its kernels are no-ops, so the payload is stale and the receiver must stop too).
__device__ __forceinline__ int hdr_count(const float4* msg) { return (int)__float_as_uint(msg[0].x); }
__device__ __forceinline__ bool hdr_failed(const float4* msg) { return __float_as_uint(msg[0].y) != 0u; }
__device__ __forceinline__ float4 hdr_make(int count) { return make_float4(__uint_as_float((uint32_t)count), 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 hdr_fail() { return make_float4(0.f, __uint_as_float(1u), 0.f, 0.f); }

// 0 = stays, 1 = goes to the left neighbour, 2 = goes to the right neighbour
__device__ __forceinline__ int slab_class(float x, float inv_h, int cut_lo, int cut_hi) {
  const int cx = cell_coord(x, inv_h);
  return cx < cut_lo ? 1 : (cx >= cut_hi ? 2 : 0);
}

// ---------------------------------------------------------------- migration: split
__global__ void __launch_bounds__(kThreads)
k_slab_count(const float4* __restrict__ pred_o, const SlabCounts* __restrict__ counts, uint32_t* __restrict__ blk_cnt,
             const StatusBlock* st, float inv_h, int cut_lo, int cut_hi, int nblocks) {
  __shared__ uint32_t wc[3][kThreads / 32];
  if (batch_failed(st)) return;
  const int n = counts->n_own;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cls = (i < n) ? slab_class(pred_o[i].x, inv_h, cut_lo, cut_hi) : 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint32_t m = __ballot_sync(0xffffffffu, cls == c);
    if (lane == 0) wc[c][warp] = __popc(m);
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += wc[threadIdx.x][w];
    blk_cnt[threadIdx.x * nblocks + blockIdx.x] = t;
  }
}

// Exclusive scan of the three per-block counter rows (one block), totals -> SlabCounts + headers.
__global__ void __launch_bounds__(1024)
k_slab_scan(uint32_t* __restrict__ blk_cnt, SlabCounts* __restrict__ counts, StatusBlock* st, float4* send_l,
            float4* send_r, int nblocks, int mcap) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_sh;
  if (batch_failed(st)) {
    if (threadIdx.x == 0) send_l[0] = send_r[0] = hdr_fail();
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < 3; ++c) {
    uint32_t* row = blk_cnt + (size_t)c * nblocks;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
      const int i = base + threadIdx.x;
      const uint32_t v = (i < nblocks) ? row[i] : 0u;
      uint32_t incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) warp_sums[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += t;
        }
        warp_sums[lane] = wi - w;
      }
      __syncthreads();
      const uint32_t excl = carry_sh + warp_sums[warp] + (incl - v);
      if (i < nblocks) row[i] = excl;
      __syncthreads();
      if (threadIdx.x == 1023) carry_sh = excl + v;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const int total = (int)carry_sh;
      if (c == 0) {
        counts->n_keep = total;
      } else {
        counts->n_send[c - 1] = total;
        if (total > mcap) st->mig_overflow = 1;
        atomicMax(&st->max_send, (unsigned int)total);
        float4* msg = (c == 1) ? send_l : send_r;
        // an overflowing message is never packed (the batch is replayed): tell the receiver so
        msg[0] = total > mcap ? hdr_fail() : hdr_make(total);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads)
k_slab_scatter(const float4* __restrict__ pos_o, const float4* __restrict__ pred_o, const uint32_t* __restrict__ gid_o,
               const SlabCounts* __restrict__ counts, const uint32_t* __restrict__ blk_cnt, float4* __restrict__ keep_pos,
               float4* __restrict__ keep_pred, float4* __restrict__ send_l, float4* __restrict__ send_r,
               const StatusBlock* st, float inv_h, int cut_lo, int cut_hi, int nblocks, int mcap) {
  __shared__ uint32_t wc[3][kThreads / 32];
  if (batch_failed(st)) return;
  const int n = counts->n_own;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = p;
  int cls = 3;
  if (i < n) {
    p = pos_o[i];
    q = pred_o[i];
    cls = slab_class(q.x, inv_h, cut_lo, cut_hi);
  }
  uint32_t rank_in_warp = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint32_t m = __ballot_sync(0xffffffffu, cls == c);
    if (lane == 0) wc[c][warp] = __popc(m);
    if (cls == c) rank_in_warp = __popc(m & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (cls == 3) return;
  uint32_t dst = blk_cnt[cls * nblocks + blockIdx.x] + rank_in_warp;
  for (int w = 0; w < warp; ++w) dst += wc[cls][w];
  p.w = __uint_as_float(gid_o[i]);
  q.w = 0.0f;
  if (cls == 0) {
    keep_pos[dst] = p;
    keep_pred[dst] = q;
  } else if (dst < (uint32_t)mcap) {
    float4* msg = (cls == 1) ? send_l : send_r;
    msg[1 + dst] = p;
    msg[1 + mcap + dst] = q;
  }
}

// ---------------------------------------------------------------- migration: merge
__device__ __forceinline__ int lower_bound_gid(const float4* __restrict__ a, int n, uint32_t gid) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__float_as_uint(a[mid].w) < gid) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Three sequences, each ascending in global id: kept (A), from the left (B), from the right (C).
// Destination of an element = its index in its own sequence + the number of smaller ids in the
// other two.  On the last hop the cell bounds of the merged set are taken for the grid.
__global__ void __launch_bounds__(kThreads)
k_slab_merge(const float4* __restrict__ keep_pos, const float4* __restrict__ keep_pred,
             const float4* __restrict__ recv_l, const float4* __restrict__ recv_r, float4* __restrict__ pos_o,
             float4* __restrict__ pred_o, uint32_t* __restrict__ gid_o, SlabCounts* __restrict__ counts,
             StatusBlock* st, float inv_h, int cut_lo, int cut_hi, int cap, int mcap, int last_hop) {
  if (batch_failed(st)) return;
  if (hdr_failed(recv_l) || hdr_failed(recv_r)) {
    st->peer_failed = 1;
    return;
  }
  const int nA = counts->n_keep;
  const int nB = hdr_count(recv_l), nC = hdr_count(recv_r);
  const int total = nA + nB + nC;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t == 0) {
    atomicMax(&st->max_own, (unsigned int)total);
    if (total > cap) st->own_overflow = 1; else counts->n_own = total;
  }
  if (total > cap) return;
  const float4* B = recv_l + 1;
  const float4* C = recv_r + 1;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX};
  int hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  if (t < total) {
    float4 p, q;
    int dst;
    if (t < nA) {
      p = keep_pos[t];
      q = keep_pred[t];
      dst = t;
      if (nB | nC) {
        const uint32_t gid = __float_as_uint(p.w);
        dst += lower_bound_gid(B, nB, gid) + lower_bound_gid(C, nC, gid);
      }
    } else if (t - nA < nB) {
      const int k = t - nA;
      p = B[k];
      q = B[mcap + k];
      const uint32_t gid = __float_as_uint(p.w);
      dst = k + lower_bound_gid(keep_pos, nA, gid) + lower_bound_gid(C, nC, gid);
    } else {
      const int k = t - nA - nB;
      p = C[k];
      q = C[mcap + k];
      const uint32_t gid = __float_as_uint(p.w);
      dst = k + lower_bound_gid(keep_pos, nA, gid) + lower_bound_gid(B, nB, gid);
    }
    gid_o[dst] = __float_as_uint(p.w);
    pos_o[dst] = make_float4(p.x, p.y, p.z, 0.0f);
    pred_o[dst] = make_float4(q.x, q.y, q.z, 0.0f);
    if (last_hop) {
      const int cx = cell_coord(q.x, inv_h), cy = cell_coord(q.y, inv_h), cz = cell_coord(q.z, inv_h);
      lo[0] = hi[0] = cx;
      lo[1] = hi[1] = cy;
      lo[2] = hi[2] = cz;
      if (cx < cut_lo || cx >= cut_hi) st->far_migrant = 1;  // needs another hop: the batch is replayed
    }
  }
  if (!last_hop) return;
  // block-level reduction, then six pre-checked atomics per block (see k_predict)
  __shared__ int s_lo[3][kThreads / 32], s_hi[3][kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int wlo = __reduce_min_sync(0xffffffffu, lo[a]);
    const int whi = __reduce_max_sync(0xffffffffu, hi[a]);
    if (lane == 0) {
      s_lo[a][warp] = wlo;
      s_hi[a][warp] = whi;
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    int blo = INT_MAX, bhi = INT_MIN;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      blo = min(blo, s_lo[a][w]);
      bhi = max(bhi, s_hi[a][w]);
    }
    if (blo < *(volatile int*)&st->min_cell[a]) atomicMin(&st->min_cell[a], blo);
    if (bhi > *(volatile int*)&st->max_cell[a]) atomicMax(&st->max_cell[a], bhi);
  }
}

// ---------------------------------------------------------------- ghost build
__device__ __forceinline__ int lower_bound_key(const uint32_t* __restrict__ keys, int n, uint32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The two x-layers next to a cut are a prefix (left) / suffix (right) of the sorted owned slots.
__global__ void k_slab_bounds(const uint32_t* __restrict__ keys, const GridDesc* __restrict__ desc,
                              SlabCounts* __restrict__ counts, StatusBlock* st, float4* send_l, float4* send_r,
                              int cut_lo, int cut_hi, int gcap) {
  if (batch_failed(st)) {
    send_l[0] = send_r[0] = hdr_fail();
    return;
  }
  const int n = counts->n_own;
  const GridDesc d = *desc;
  const uint32_t layer = (uint32_t)d.dim[1] * (uint32_t)d.dim[2];
  int b0 = 0, b1 = 0;
  if (cut_lo != INT_MIN) {
    const long long xi = (long long)cut_lo + 2 - d.lo[0];
    b0 = xi <= 0 ? 0 : (xi >= d.dim[0] ? n : lower_bound_key(keys, n, (uint32_t)xi * layer));
  }
  if (cut_hi != INT_MAX) {
    const long long xi = (long long)cut_hi - 2 - d.lo[0];
    b1 = xi <= 0 ? n : (xi >= d.dim[0] ? 0 : n - lower_bound_key(keys, n, (uint32_t)xi * layer));
  }
  const int worst = b0 > b1 ? b0 : b1;
  atomicMax(&st->max_ghost, (unsigned int)worst);
  if (worst > gcap) {  // never packed (the batch is replayed): tell the receivers so
    st->ghost_overflow = 1;
    counts->b[0] = counts->b[1] = 0;
    send_l[0] = send_r[0] = hdr_fail();
    return;
  }
  counts->b[0] = b0;
  counts->b[1] = b1;
  send_l[0] = hdr_make(b0);
  send_r[0] = hdr_make(b1);
}

__global__ void __launch_bounds__(kThreads)
k_slab_ghost_pack(const float4* __restrict__ pred_s, const float4* __restrict__ pos_s,
                  const SlabCounts* __restrict__ counts, float4* __restrict__ send_l, float4* __restrict__ send_r,
                  const StatusBlock* st, int gcap) {
  if (batch_failed(st)) return;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int side = t / gcap, k = t - side * gcap;
  if (side > 1 || k >= counts->b[side]) return;
  const int src = side == 0 ? k : counts->n_own - counts->b[1] + k;
  float4* msg = side == 0 ? send_l : send_r;
  const float4 q = pred_s[src];
  msg[1 + k] = make_float4(q.x, q.y, q.z, 0.0f);
  msg[1 + gcap + k] = pos_s[src];
}

__device__ __forceinline__ uint32_t ghost_key(float4 q, float inv_h, const GridDesc& d) {
  const long long rx = (long long)cell_coord(q.x, inv_h) - d.lo[0];
  const long long ry = (long long)cell_coord(q.y, inv_h) - d.lo[1];
  const long long rz = (long long)cell_coord(q.z, inv_h) - d.lo[2];
  if (rx < 0 || rx >= d.dim[0] || ry < 0 || ry >= d.dim[1] || rz < 0 || rz >= d.dim[2]) return kInvalidKey;
  return ((uint32_t)rx * (uint32_t)d.dim[1] + (uint32_t)ry) * (uint32_t)d.dim[2] + (uint32_t)rz;
}

// Ghost k of a message goes to slot n_own + [ghosts from the left] + k.  Ghosts whose cell lies
// outside this slab's table can not be neighbours of an owned particle and get no table entry.
__global__ void __launch_bounds__(kThreads)
k_slab_ghost_unpack(const float4* __restrict__ recv_l, const float4* __restrict__ recv_r, float4* __restrict__ pred_s,
                    float4* __restrict__ pos_s, int2* __restrict__ cell_range, const GridDesc* __restrict__ desc,
                    SlabCounts* __restrict__ counts, StatusBlock* st, float inv_h, int gcap, int tot_cap) {
  if (batch_failed(st)) return;
  if (hdr_failed(recv_l) || hdr_failed(recv_r)) {
    st->peer_failed = 1;
    return;
  }
  const int n_own = counts->n_own;
  const int g0 = hdr_count(recv_l), g1 = hdr_count(recv_r);
  const int tot = n_own + g0 + g1;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t == 0) {
    counts->n_ghost[0] = g0;
    counts->n_ghost[1] = g1;
    counts->n_tot = tot > tot_cap ? n_own : tot;
    atomicMax(&st->max_own, (unsigned int)tot);
    if (tot > tot_cap) st->own_overflow = 1;
  }
  if (tot > tot_cap) return;
  const int side = t / gcap, k = t - side * gcap;
  if (side > 1) return;
  const int ng = side == 0 ? g0 : g1;
  if (k >= ng) return;
  const float4* msg = side == 0 ? recv_l : recv_r;
  const int slot = n_own + (side == 0 ? 0 : g0) + k;
  const float4 q = msg[1 + k];
  const float4 p = msg[1 + gcap + k];
  pred_s[slot] = make_float4(q.x, q.y, q.z, 0.0f);
  pos_s[slot] = make_float4(p.x, p.y, p.z, __uint_as_float(0xffffffffu));
  const GridDesc d = *desc;
  const uint32_t key = ghost_key(q, inv_h, d);
  if (key == kInvalidKey) return;
  const bool first = (k == 0) || ghost_key(msg[k], inv_h, d) != key;          // msg[1 + (k-1)]
  const bool last = (k == ng - 1) || ghost_key(msg[2 + k], inv_h, d) != key;  // msg[1 + (k+1)]
  if (first) cell_range[key].x = slot;
  if (last) cell_range[key].y = slot + 1;
}

// ---------------------------------------------------------------- per-iteration halo refresh
// (the outgoing side is HaloOut::put inside the delta / xsph kernels, kernels/solve.cu)
__global__ void __launch_bounds__(kThreads)
k_slab_halo_unpack(const float4* __restrict__ recv_l, const float4* __restrict__ recv_r, float4* __restrict__ arr,
                   const SlabCounts* __restrict__ counts, const StatusBlock* st, int gcap) {
  if (batch_failed(st)) return;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int side = t / gcap, k = t - side * gcap;
  if (side > 1 || k >= counts->n_ghost[side]) return;
  const int slot = counts->n_own + (side == 0 ? 0 : counts->n_ghost[0]) + k;
  arr[slot] = (side == 0 ? recv_l : recv_r)[k];
}

// velocity update of the ghosts (reference core.cpp:414-420 applied to a remote particle): the
// owner computes exactly the same expression, so no message is needed.
template <bool S>
__global__ void __launch_bounds__(kThreads)
k_slab_ghost_vel(const float4* __restrict__ pred, const float4* __restrict__ pos_s, const float* __restrict__ rho,
                 float4* __restrict__ vel, const SlabCounts* __restrict__ counts, const StatusBlock* st, StepConsts c) {
  if (batch_failed(st)) return;
  const int i = counts->n_own + blockIdx.x * kThreads + threadIdx.x;
  if (i >= counts->n_tot) return;
  const float4 q = pred[i];
  const float4 p = pos_s[i];
  float vx, vy, vz, inv_rho;
  const float r = rho[i];
  if (S) {
    vx = __fdiv_rn(__fsub_rn(q.x, p.x), c.dt);
    vy = __fdiv_rn(__fsub_rn(q.y, p.y), c.dt);
    vz = __fdiv_rn(__fsub_rn(q.z, p.z), c.dt);
    inv_rho = r > 0.0f ? __fdiv_rn(c.mass, r) : 0.0f;
  } else {
    vx = (q.x - p.x) * c.inv_dt;
    vy = (q.y - p.y) * c.inv_dt;
    vz = (q.z - p.z) * c.inv_dt;
    inv_rho = r > 0.0f ? c.mass / r : 0.0f;
  }
  vel[i] = make_float4(vx, vy, vz, inv_rho);
}

inline int grid_for(int n) { return (n + kThreads - 1) / kThreads; }

}  // namespace

// ================================================================== launchers
int launch_slab_split(const float4* pos_o, const float4* pred_o, const SlabBuffers& sb, const StepConsts& c,
                      cudaStream_t s) {
  const int nb = grid_for(sb.cap);
  k_slab_count<<<nb, kThreads, 0, s>>>(pred_o, sb.counts, sb.blk_cnt, sb.status, c.inv_h, sb.cut_lo, sb.cut_hi, nb);
  k_slab_scan<<<1, 1024, 0, s>>>(sb.blk_cnt, sb.counts, sb.status, sb.send[0], sb.send[1], nb, sb.mcap);
  k_slab_scatter<<<nb, kThreads, 0, s>>>(pos_o, pred_o, sb.gid_o, sb.counts, sb.blk_cnt, sb.keep_pos, sb.keep_pred,
                                         sb.send[0], sb.send[1], sb.status, c.inv_h, sb.cut_lo, sb.cut_hi, nb, sb.mcap);
  return 3;
}

int launch_slab_merge(float4* pos_o, float4* pred_o, const SlabBuffers& sb, const StepConsts& c, bool last_hop,
                      cudaStream_t s) {
  k_slab_merge<<<grid_for(sb.cap + 2 * sb.mcap), kThreads, 0, s>>>(
      sb.keep_pos, sb.keep_pred, sb.recv[0], sb.recv[1], pos_o, pred_o, sb.gid_o, sb.counts, sb.status, c.inv_h,
      sb.cut_lo, sb.cut_hi, sb.cap, sb.mcap, last_hop ? 1 : 0);
  return 1;
}

int launch_slab_ghost_pack(const uint32_t* keys_sorted, const float4* pred_s, const float4* pos_s,
                           const GridBuffers& g, const SlabBuffers& sb, cudaStream_t s) {
  k_slab_bounds<<<1, 1, 0, s>>>(keys_sorted, g.desc, sb.counts, sb.status, sb.send[0], sb.send[1], sb.cut_lo,
                               sb.cut_hi, sb.gcap);
  k_slab_ghost_pack<<<grid_for(2 * sb.gcap), kThreads, 0, s>>>(pred_s, pos_s, sb.counts, sb.send[0], sb.send[1],
                                                              sb.status, sb.gcap);
  return 2;
}

int launch_slab_ghost_unpack(float4* pred_s, float4* pos_s, const GridBuffers& g, const SlabBuffers& sb,
                             const StepConsts& c, cudaStream_t s) {
  k_slab_ghost_unpack<<<grid_for(2 * sb.gcap), kThreads, 0, s>>>(sb.recv[0], sb.recv[1], pred_s, pos_s, g.cell_range,
                                                                g.desc, sb.counts, sb.status, c.inv_h, sb.gcap,
                                                                sb.tot_cap);
  return 1;
}

int launch_slab_halo_unpack(float4* arr, const SlabBuffers& sb, cudaStream_t s) {
  k_slab_halo_unpack<<<grid_for(2 * sb.gcap), kThreads, 0, s>>>(sb.recv[0], sb.recv[1], arr, sb.counts, sb.status,
                                                               sb.gcap);
  return 1;
}

int launch_slab_ghost_vel(const float4* pred_final, const float4* pos_s, const float* rho, float4* vel,
                          const SlabBuffers& sb, const StepConsts& c, bool strict, cudaStream_t s) {
  if (strict)
    k_slab_ghost_vel<true><<<grid_for(2 * sb.gcap), kThreads, 0, s>>>(pred_final, pos_s, rho, vel, sb.counts, sb.status, c);
  else
    k_slab_ghost_vel<false><<<grid_for(2 * sb.gcap), kThreads, 0, s>>>(pred_final, pos_s, rho, vel, sb.counts, sb.status, c);
  return 1;
}

}  // namespace pbf

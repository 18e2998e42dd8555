// slab.cu — device side of the x-slab decomposition.
//
// The reference has no multi-GPU path; what pins this file is the single-GPU result: the
// reference's sort order is x-major (cell_key_less, reference core/src/core.cpp:12-21), so the
// particles of the x-cells [cut_lo, cut_hi) are one contiguous range of the global sorted order.
// Two things make the slab result BIT-IDENTICAL to one GPU:
//   (1) the counting sort orders the members of a cell by GLOBAL particle id (k_cell_order), which
//       is the reference's tie-break (core.cpp:182) whatever order a slab stores its particles in;
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

// ---------------------------------------------------------------- migration
// The order in which a slab stores its particles is irrelevant (cells are ordered by global id in
// k_cell_order), so migration moves O(migrants) data: leavers are copied into the messages and
// their slots recorded; the vacated slots below the new end are refilled with the stayers of the
// tail; arrivals are appended.  Only k_slab_leave touches every particle (one 16-byte read).
__global__ void __launch_bounds__(kThreads)
k_slab_leave(const float4* __restrict__ pos_o, const float4* __restrict__ pred_o, const uint32_t* __restrict__ gid_o,
             SlabCounts* __restrict__ counts, uint32_t* __restrict__ holes, float4* __restrict__ send_l,
             float4* __restrict__ send_r, const StatusBlock* st, float inv_h, int mcap) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= counts->n_own) return;
  const float4 q = pred_o[i];
  const int cls = slab_class(q.x, inv_h, counts->cut_lo, counts->cut_hi);
  if (cls == 0) return;
  const int k = atomicAdd(&counts->n_send[cls - 1], 1);
  if (k < mcap) {
    float4* msg = (cls == 1) ? send_l : send_r;
    const float4 p = pos_o[i];
    msg[1 + k] = make_float4(p.x, p.y, p.z, __uint_as_float(gid_o[i]));
    msg[1 + mcap + k] = make_float4(q.x, q.y, q.z, 0.0f);
  }
  const int h = atomicAdd(&counts->n_holes, 1);
  if (h < 2 * mcap) holes[h] = (uint32_t)i;
}

// One block: message headers, then holes below the new end <- stayers of the tail.
__global__ void __launch_bounds__(1024)
k_slab_refill(float4* __restrict__ pos_o, float4* __restrict__ pred_o, uint32_t* __restrict__ gid_o,
              SlabCounts* __restrict__ counts, uint32_t* __restrict__ holes, float4* send_l, float4* send_r,
              StatusBlock* st, float inv_h, int mcap) {
  pdl_wait();
  __shared__ int s_a, s_b;
  const int cut_lo = counts->cut_lo, cut_hi = counts->cut_hi;
  if (batch_failed(st)) {
    if (threadIdx.x == 0) send_l[0] = send_r[0] = hdr_fail();
    return;
  }
  const int n = counts->n_own, L = counts->n_holes;
  const int nl = counts->n_send[0], nr = counts->n_send[1];
  const bool overflow = nl > mcap || nr > mcap;
  if (threadIdx.x == 0) {
    atomicMax(&st->max_send, (unsigned int)(nl > nr ? nl : nr));
    if (overflow) st->mig_overflow = 1;
    // an overflowing message is never complete (the batch is replayed): tell the receiver so
    send_l[0] = overflow ? hdr_fail() : hdr_make(nl);
    send_r[0] = overflow ? hdr_fail() : hdr_make(nr);
    s_a = s_b = 0;
  }
  if (overflow) return;
  __syncthreads();
  const int base = n - L;  // owned particles that stay
  uint32_t* list_a = holes + 2 * mcap;  // vacated slots below base
  uint32_t* list_b = holes + 4 * mcap;  // stayers at or above base
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const int h = (int)holes[t];
    if (h < base) list_a[atomicAdd(&s_a, 1)] = (uint32_t)h;
    const int i = base + t;
    if (slab_class(pred_o[i].x, inv_h, cut_lo, cut_hi) == 0) list_b[atomicAdd(&s_b, 1)] = (uint32_t)i;
  }
  __syncthreads();
  const int moves = s_a;  // == s_b: L slots vanish, every stayer above base has a hole below it
  for (int t = threadIdx.x; t < moves; t += blockDim.x) {
    const uint32_t dst = list_a[t], src = list_b[t];
    pos_o[dst] = pos_o[src];
    pred_o[dst] = pred_o[src];
    gid_o[dst] = gid_o[src];
  }
  if (threadIdx.x == 0) counts->n_keep = base;
}

// Arrivals are appended after the particles that stayed.
__global__ void __launch_bounds__(kThreads)
k_slab_arrive(const float4* __restrict__ recv_l, const float4* __restrict__ recv_r, float4* __restrict__ pos_o,
              float4* __restrict__ pred_o, uint32_t* __restrict__ gid_o, SlabCounts* __restrict__ counts,
              StatusBlock* st, float inv_h, int cap, int mcap, int last_hop, PeerSync sync) {
  pdl_wait();
  peer_sync(sync, st);
  if (batch_failed(st)) return;
  const int cut_lo = counts->cut_lo, cut_hi = counts->cut_hi;
  if (hdr_failed(recv_l) || hdr_failed(recv_r)) {
    st->peer_failed = 1;
    return;
  }
  const int base = counts->n_keep;
  const int nB = hdr_count(recv_l), nC = hdr_count(recv_r);
  const int total = base + nB + nC;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t == 0) {
    atomicMax(&st->max_own, (unsigned int)total);
    if (total > cap) st->own_overflow = 1; else counts->n_own = total;
    counts->n_send[0] = counts->n_send[1] = counts->n_holes = 0;  // for the next hop / substep
  }
  if (total > cap) return;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX};
  int hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  if (t < nB + nC) {
    const float4* msg = t < nB ? recv_l : recv_r;
    const int k = t < nB ? t : t - nB;
    const float4 p = msg[1 + k], q = msg[1 + mcap + k];
    const int dst = base + t;
    gid_o[dst] = __float_as_uint(p.w);
    pos_o[dst] = make_float4(p.x, p.y, p.z, 0.0f);
    pred_o[dst] = make_float4(q.x, q.y, q.z, 0.0f);
    const int cx = cell_coord(q.x, inv_h), cy = cell_coord(q.y, inv_h), cz = cell_coord(q.z, inv_h);
    lo[0] = hi[0] = cx;
    lo[1] = hi[1] = cy;
    lo[2] = hi[2] = cz;
    if (last_hop && (cx < cut_lo || cx >= cut_hi)) st->far_migrant = 1;  // needs another hop: the batch is replayed
  }
  // the bounds of the particles that were here before come from k_predict (a superset: it also saw
  // the leavers); the arrivals of every hop extend them
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int wlo = __reduce_min_sync(0xffffffffu, lo[a]);
    const int whi = __reduce_max_sync(0xffffffffu, hi[a]);
    if ((threadIdx.x & 31) == 0 && wlo != INT_MAX) {
      if (wlo < *(volatile int*)&st->min_cell[a]) atomicMin(&st->min_cell[a], wlo);
      if (whi > *(volatile int*)&st->max_cell[a]) atomicMax(&st->max_cell[a], whi);
    }
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
                              int gcap) {
  pdl_wait();
  const int cut_lo = counts->cut_lo, cut_hi = counts->cut_hi;
  if (batch_failed(st)) {
    send_l[0] = send_r[0] = hdr_fail();
    return;
  }
  const int n = counts->n_own;
  const GridDesc d = *desc;
  const uint32_t layer = (uint32_t)d.dim[1] * (uint32_t)d.dim[2];
  int b0 = 0, b1 = 0, c0 = 0, c1 = 0;
  // owned sorted slots with x-cell < x: a prefix of the sorted order
  auto below = [&](long long x) -> int {
    const long long xi = x - d.lo[0];
    return xi <= 0 ? 0 : (xi >= d.dim[0] ? n : lower_bound_key(keys, n, (uint32_t)xi * layer));
  };
  if (cut_lo != INT_MIN) {
    b0 = below((long long)cut_lo + 2);
    c0 = below((long long)cut_lo + 1);
  }
  if (cut_hi != INT_MAX) {
    b1 = n - below((long long)cut_hi - 2);
    c1 = n - below((long long)cut_hi - 1);
  }
  if (c0 + c1 > n) c1 = n - c0;  // a slab whose first and last layer overlap: everything is boundary
  counts->c1[0] = c0;
  counts->c1[1] = c1;
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
  pdl_wait();
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
                    SlabCounts* __restrict__ counts, StatusBlock* st, float inv_h, int gcap, int tot_cap, PeerSync sync) {
  pdl_wait();
  peer_sync(sync, st);
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
                   const SlabCounts* __restrict__ counts, StatusBlock* st, int gcap, PeerSync sync) {
  pdl_wait();
  peer_sync(sync, st);
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
                 float4* __restrict__ vel, PosVel* __restrict__ pv, const SlabCounts* __restrict__ counts,
                 const StatusBlock* st, StepConsts c) {
  pdl_wait();
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
  const float4 v4 = make_float4(vx, vy, vz, inv_rho);
  if (pv) {  // XSPH follows and gathers (pos, vel) records, see k_delta
    pv[i].p = make_float4(q.x, q.y, q.z, 0.0f);
    pv[i].v = v4;
  } else {
    vel[i] = v4;
  }
}

// ---------------------------------------------------------------- re-balancing support
__global__ void k_slab_report(const SlabCounts* __restrict__ counts, StatusBlock* st, int rank) {
  pdl_wait();
  if (rank < kMaxSlabs) st->own_by_rank[rank] = (unsigned int)counts->n_own;
}

__global__ void __launch_bounds__(kThreads)
k_slab_xrange(const float4* __restrict__ pos_o, const float4* __restrict__ vel_o, const SlabCounts* __restrict__ counts,
              float inv_h, float lookahead, int* out) {
  pdl_wait();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  int lo = INT_MAX, hi = INT_MIN;
  // where the particle will be `lookahead` seconds from now if it keeps its velocity: the cuts are
  // planned for the middle of the batch they will serve, not for the moment of planning
  if (i < counts->n_own) lo = hi = cell_coord(fmaf(vel_o[i].x, lookahead, pos_o[i].x), inv_h);
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0 && lo != INT_MAX) {
    atomicMin(&out[0], lo);
    atomicMax(&out[1], hi);
  }
}

__global__ void __launch_bounds__(kThreads)
k_slab_xhist(const float4* __restrict__ pos_o, const float4* __restrict__ vel_o, const SlabCounts* __restrict__ counts,
             float inv_h, float lookahead, int x_min, int layers, unsigned long long* __restrict__ hist) {
  pdl_wait();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= counts->n_own) return;
  const long long l = (long long)cell_coord(fmaf(vel_o[i].x, lookahead, pos_o[i].x), inv_h) - x_min;
  if (l >= 0 && l < layers) atomicAdd(&hist[l], 1ull);
}

inline int grid_for(int n) { return (n + kThreads - 1) / kThreads; }

// The exchange that precedes consumer kernel `kind`, when the transport fused it (SlabBuffers::sync).
inline PeerSync sync_for(const SlabBuffers& sb, int kind) {
  PeerSync ps = sb.sync;
  if (ps.mine) {
    ps.arrivals = &sb.counts->sync_arrivals[kind];
    ps.done = &sb.counts->sync_done[kind];
  }
  return ps;
}

}  // namespace

// ================================================================== launchers
int launch_slab_split(float4* pos_o, float4* pred_o, const SlabBuffers& sb, const StepConsts& c, cudaStream_t s) {
  PBF_LAUNCH(k_slab_leave, grid_for(sb.cap), kThreads, s, pos_o, pred_o, sb.gid_o, sb.counts, sb.holes, sb.send[0],
                                                    sb.send[1], sb.status, c.inv_h, sb.mcap);
  PBF_LAUNCH(k_slab_refill, 1, 1024, s, pos_o, pred_o, sb.gid_o, sb.counts, sb.holes, sb.send[0], sb.send[1], sb.status,
                                  c.inv_h, sb.mcap);
  return 2;
}

int launch_slab_merge(float4* pos_o, float4* pred_o, const SlabBuffers& sb, const StepConsts& c, bool last_hop,
                      cudaStream_t s) {
  PBF_LAUNCH(k_slab_arrive, grid_for(2 * sb.mcap), kThreads, s, sb.recv[0], sb.recv[1], pos_o, pred_o, sb.gid_o, sb.counts,
                                                          sb.status, c.inv_h, sb.cap, sb.mcap,
                                                          last_hop ? 1 : 0, sync_for(sb, 0));
  return 1;
}

int launch_slab_ghost_pack(const uint32_t* keys_sorted, const float4* pred_s, const float4* pos_s,
                           const GridBuffers& g, const SlabBuffers& sb, cudaStream_t s) {
  PBF_LAUNCH(k_slab_bounds, 1, 1, s, keys_sorted, g.desc, sb.counts, sb.status, sb.send[0], sb.send[1], sb.gcap);
  PBF_LAUNCH(k_slab_ghost_pack, grid_for(2 * sb.gcap), kThreads, s, pred_s, pos_s, sb.counts, sb.send[0], sb.send[1],
                                                              sb.status, sb.gcap);
  return 2;
}

int launch_slab_ghost_unpack(float4* pred_s, float4* pos_s, const GridBuffers& g, const SlabBuffers& sb,
                             const StepConsts& c, cudaStream_t s) {
  PBF_LAUNCH(k_slab_ghost_unpack, grid_for(2 * sb.gcap), kThreads, s, sb.recv[0], sb.recv[1], pred_s, pos_s, g.cell_range,
                                                                g.desc, sb.counts, sb.status, c.inv_h, sb.gcap,
                                                                sb.tot_cap, sync_for(sb, 1));
  return 1;
}

int launch_slab_halo_unpack(float4* arr, const SlabBuffers& sb, cudaStream_t s) {
  PBF_LAUNCH(k_slab_halo_unpack, grid_for(2 * sb.gcap), kThreads, s, sb.recv[0], sb.recv[1], arr, sb.counts, sb.status,
                                                               sb.gcap, sync_for(sb, 2));
  return 1;
}

int launch_slab_report(const SlabBuffers& sb, int rank, cudaStream_t s) {
  PBF_LAUNCH(k_slab_report, 1, 1, s, sb.counts, sb.status, rank);
  return 1;
}

int launch_slab_xrange(const float4* pos_o, const float4* vel_o, const SlabBuffers& sb, const StepConsts& c, float lookahead,
                       int* out_min_max, cudaStream_t s) {
  PBF_LAUNCH(k_slab_xrange, grid_for(sb.cap), kThreads, s, pos_o, vel_o, sb.counts, c.inv_h, lookahead, out_min_max);
  return 1;
}

int launch_slab_xhist(const float4* pos_o, const float4* vel_o, const SlabBuffers& sb, const StepConsts& c, float lookahead,
                      int x_min, int layers, unsigned long long* hist, cudaStream_t s) {
  PBF_LAUNCH(k_slab_xhist, grid_for(sb.cap), kThreads, s, pos_o, vel_o, sb.counts, c.inv_h, lookahead, x_min, layers, hist);
  return 1;
}

int launch_slab_ghost_vel(const float4* pred_final, const float4* pos_s, const float* rho, float4* vel, PosVel* pv,
                          const SlabBuffers& sb, const StepConsts& c, bool strict, cudaStream_t s) {
  if (strict)
    PBF_LAUNCH(k_slab_ghost_vel<true>, grid_for(2 * sb.gcap), kThreads, s, pred_final, pos_s, rho, vel, pv, sb.counts, sb.status, c);
  else
    PBF_LAUNCH(k_slab_ghost_vel<false>, grid_for(2 * sb.gcap), kThreads, s, pred_final, pos_s, rho, vel, pv, sb.counts, sb.status, c);
  return 1;
}

}  // namespace pbf

// grid.cu — neighbour-grid kernels of the PBF substep for sm_100a:
//   a3  predict / integrate            (reference core/src/core.cpp:150-161)
//   a4  cell coordinates and keys      (core.cpp:28-34, 164-171)
//   a5  sort by (x, y, z, particle id) (core.cpp:12-21, 173-183)  -> counting sort on a dense cell key
//   a6  cell start/end table           (core.cpp:185-203)         -> the exclusive scan of that sort
//   a7  neighbour list                 (core.cpp:205-247)         -> warp-interleaved ELL list
//
// Integer results (keys, sorted order, cell table, neighbour sets) are bit-exact with
// the reference in BOTH arithmetic modes: everything that feeds them uses explicit
// round-to-nearest intrinsics, never contracted.
//
// Key design points
//   * dense key = ((x-x0)*ny + (y-y0))*nz + (z-z0) over the per-substep bounding box of
//     occupied cells: order-isomorphic to the reference's lexicographic (x,y,z) compare and
//     small (~19 bits), so cell counts + exclusive scan + placement, with the members of a cell
//     ranked by particle id, reproduce the reference's (key, particle) total order exactly.
//   * the table is (start,end) per dense cell, padded by one empty layer so the 27-cell
//     stencil needs no bounds checks; a box too large for it (the reference's vorticity blow-up)
//     switches the substep to a sparse table: same arrays, hashed cell coordinates.
//   * nothing here synchronises with the host: bounds, cell counts, the dense/sparse decision
//     and overflow flags live in device memory (GridDesc / StatusBlock).
#include "pbf_kernels.h"
#include "neighbor_test.cuh"

namespace pbf {

namespace {

constexpr int kThreads = 256;
// ---------------------------------------------------------------- state (de)interleave
__global__ void __launch_bounds__(kThreads)
k_pack_state(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
             const float* __restrict__ vx, const float* __restrict__ vy, const float* __restrict__ vz,
             float4* __restrict__ pos_o, float4* __restrict__ vel_o, int n) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pos_o[i] = make_float4(px[i], py[i], pz[i], 0.0f);
  vel_o[i] = make_float4(vx[i], vy[i], vz[i], 0.0f);
}

__global__ void __launch_bounds__(kThreads)
k_unpack_state(const float4* __restrict__ pos_o, const float4* __restrict__ vel_o,
               float* __restrict__ px, float* __restrict__ py, float* __restrict__ pz,
               float* __restrict__ vx, float* __restrict__ vy, float* __restrict__ vz, int n) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pos_o[i];
  const float4 v = vel_o[i];
  if (px) px[i] = p.x;
  if (py) py[i] = p.y;
  if (pz) pz[i] = p.z;
  if (vx) vx[i] = v.x;
  if (vy) vy[i] = v.y;
  if (vz) vz[i] = v.z;
}

// ---------------------------------------------------------------- a3 + a4 bounds
// vel += g*dt; pred = pos + vel*dt (core.cpp:155-160), then min/max of the cell
// coordinates.  Grid-stride so that only a few thousand warps touch the six atomics.
__device__ void grid_finalize_body(GridDesc* desc, StatusBlock* st, uint32_t cell_cap, int pad, NRef nr, int allow_sparse,
                                   int brick_cap);

// finalize != 0 (one GPU): the block that delivers its bounds last also writes the table descriptor
// of the substep (what k_grid_finalize does as a kernel of its own for slabs, after migration).
__global__ void __launch_bounds__(kThreads)
k_predict(const float4* __restrict__ pos_o, const float4* __restrict__ vel_o,
          float4* __restrict__ pred_o, StepConsts c, StatusBlock* st, NRef nr, int do_bounds,
          GridDesc* desc, unsigned long long* __restrict__ cell_key, uint32_t cell_cap, int finalize, int brick_cap) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int n = nr.get();
  // the sparse cell table of the PREVIOUS substep (if it was sparse) is wiped here, by the first
  // kernel of the substep, so that the dense path never pays for it
  if (desc->sparse)
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < desc->ncells; c += gridDim.x * blockDim.x)
      cell_key[c] = kEmptyCell;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX};
  int hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pos_o[i];
    const float4 v = vel_o[i];
    const float vx = __fadd_rn(v.x, c.gdt_x);
    const float vy = __fadd_rn(v.y, c.gdt_y);
    const float vz = __fadd_rn(v.z, c.gdt_z);
    const float qx = __fadd_rn(p.x, __fmul_rn(vx, c.dt));
    const float qy = __fadd_rn(p.y, __fmul_rn(vy, c.dt));
    const float qz = __fadd_rn(p.z, __fmul_rn(vz, c.dt));
    pred_o[i] = make_float4(qx, qy, qz, 0.0f);
    const int cx = cell_coord(qx, c.inv_h), cy = cell_coord(qy, c.inv_h), cz = cell_coord(qz, c.inv_h);
    lo[0] = min(lo[0], cx); hi[0] = max(hi[0], cx);
    lo[1] = min(lo[1], cy); hi[1] = max(hi[1], cy);
    lo[2] = min(lo[2], cz); hi[2] = max(hi[2], cz);
  }
  if (!do_bounds) return;
  // block-level reduction first: six pre-checked atomics per BLOCK.  (Per-warp atomics made every
  // warp poll the same L2 line; with one particle per thread that serialised the whole kernel.)
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
    // plain loads first: after the first few blocks almost nobody needs the atomic
    if (blo < *(volatile int*)&st->min_cell[a]) atomicMin(&st->min_cell[a], blo);
    if (bhi > *(volatile int*)&st->max_cell[a]) atomicMax(&st->max_cell[a], bhi);
    __threadfence();
  }
  if (!finalize) return;
  // last block out: every block's bounds (and its reads of the previous substep's descriptor) are done
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&desc->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    grid_finalize_body(desc, st, cell_cap, 1, nr, 1, brick_cap);
    desc->ticket = 0;
  }
}

// `pad` empty layers surround the occupied cells: 1 so the 27-cell stencil of every particle stays
// inside the table; 2 in slab mode, where first-layer ghosts run their own stencil.
// When the box would need more cells than the table holds — the reference diverges with vorticity
// on, SURVEY §0 — the substep switches to a SPARSE table: the same arrays, indexed by a hash of
// the packed cell coordinates (`allow_sparse`; not in slab mode, whose ghost layers rely on the
// x-major order of the dense key).
__global__ void k_grid_finalize(GridDesc* desc, StatusBlock* st, uint32_t cell_cap, int pad, NRef nr, int allow_sparse,
                                int brick_cap) {
  pdl_wait();
  if (batch_failed(st)) return;
  grid_finalize_body(desc, st, cell_cap, pad, nr, allow_sparse, brick_cap);
}

// One thread.  The bounds were accumulated with atomics (possibly by other blocks of the running
// kernel): they are read through volatile loads, never from a stale L1 line.
__device__ void grid_finalize_body(GridDesc* desc, StatusBlock* st, uint32_t cell_cap, int pad, NRef nr, int allow_sparse,
                                   int brick_cap) {
  unsigned long long cells = 1;
  bool bad = false;
  int mn[3], mx[3];
  for (int a = 0; a < 3; ++a) {
    mn[a] = *(volatile int*)&st->min_cell[a];
    mx[a] = *(volatile int*)&st->max_cell[a];
  }
  // a slab that currently owns no particle (everything migrated away) keeps a minimal table
  const bool empty = mn[0] == INT_MAX && mx[0] == INT_MIN;
  for (int a = 0; a < 3; ++a) {
    if (empty) mn[a] = mx[a] = 0;
    const long long lo = (long long)mn[a] - pad;
    const long long hi = (long long)mx[a] + pad;
    const long long dim = hi - lo + 1;
    if (mn[a] == INT_MIN || mx[a] == INT_MIN || dim <= 0 || dim > 0x7fffffffLL) bad = true;
    desc->lo[a] = (int)lo;
    desc->hi[a] = (int)hi;
    desc->dim[a] = (int)dim;
    if (!bad) {
      if (cells > 0xffffffffffffULL / (unsigned long long)dim) bad = true; else cells *= (unsigned long long)dim;
    }
    st->min_cell[a] = INT_MAX;
    st->max_cell[a] = INT_MIN;
  }
  if (bad) cells = 0xffffffffffffULL;
  // sparse alternative: a hash table at load factor <= 1/2 over cell_cap (a power of two) slots
  const unsigned long long n = (unsigned long long)nr.get();
  const bool packable = allow_sparse && !bad && desc->dim[0] < (1 << 21) && desc->dim[1] < (1 << 21) && desc->dim[2] < (1 << 21);
  const bool dense = cells <= (unsigned long long)cell_cap;
  const bool sparse = !dense && packable && 2 * n <= (unsigned long long)cell_cap;
  // what the host must provide if neither fits: the smaller of the two tables
  unsigned long long need = cells;
  if (!dense && packable && 2 * n < need) need = 2 * n;
  const unsigned long long seen = ((unsigned long long)st->max_cells_hi << 32) | st->max_cells_lo;
  if (need > seen) {
    st->max_cells_hi = (unsigned int)(need >> 32);
    st->max_cells_lo = (unsigned int)need;
  }
  const int overflow = (dense || sparse) ? 0 : 1;
  desc->ncells = overflow ? 0u : (dense ? (uint32_t)cells : cell_cap);
  desc->sparse = sparse ? 1 : 0;
  desc->overflow = overflow;
  if (overflow) st->grid_overflow = 1;
  // brick path (brick.cu): bricks tile the dense table; a sparse table has no z-runs to stage
  desc->nbricks = 0;
  if (brick_cap > 0 && !overflow) {
    if (!dense) {
      st->brick_overflow |= kBrickDisable;
    } else {
      const unsigned long long bx = (unsigned)(desc->dim[0] + kBrickX - 1) / kBrickX, by = (unsigned)(desc->dim[1] + kBrickY - 1) / kBrickY,
                               bz = (unsigned)(desc->dim[2] + kBrickZ - 1) / kBrickZ;
      const unsigned long long nb = bx * by * bz;
      if (nb > st->max_bricks) st->max_bricks = nb > 0xffffffffull ? 0xffffffffu : (unsigned int)nb;
      if (nb > (unsigned long long)brick_cap) {
        st->brick_overflow |= kBrickGrow;
      } else {
        desc->nbricks = (int)nb;
        desc->bdim[0] = (int)bx;
        desc->bdim[1] = (int)by;
        desc->bdim[2] = (int)bz;
      }
    }
  }
}

// ---------------------------------------------------------------- a4 cell keys
__device__ __forceinline__ uint32_t dense_key(float4 q, float inv_h, const GridDesc& d) {
  const int cx = cell_coord(q.x, inv_h) - d.lo[0];
  const int cy = cell_coord(q.y, inv_h) - d.lo[1];
  const int cz = cell_coord(q.z, inv_h) - d.lo[2];
  return ((uint32_t)cx * (uint32_t)d.dim[1] + (uint32_t)cy) * (uint32_t)d.dim[2] + (uint32_t)cz;
}

// sparse table: slot of a cell, inserting it if new (linear probing; load factor <= 1/2)
__device__ __forceinline__ uint32_t sparse_insert(unsigned long long* __restrict__ cell_key, uint32_t mask,
                                                  unsigned long long key) {
  uint32_t slot = hash_cell(key) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&cell_key[slot], kEmptyCell, key);
    if (prev == kEmptyCell || prev == key) return slot;
    slot = (slot + 1) & mask;
  }
}

// (start, end) of the cell at bbox-relative coordinates (rx, ry, rz); empty if the cell does not exist
__device__ __forceinline__ int2 sparse_lookup(const unsigned long long* __restrict__ cell_key,
                                              const int2* __restrict__ cell_range, uint32_t mask, int rx, int ry, int rz) {
  const unsigned long long key = pack_cell(rx, ry, rz);
  uint32_t slot = hash_cell(key) & mask;
  for (;;) {
    const unsigned long long k = cell_key[slot];
    if (k == key) return cell_range[slot];
    if (k == kEmptyCell) return make_int2(0, 0);
    slot = (slot + 1) & mask;
  }
}

// Two-level exclusive scan helpers: a block scans one chunk of kScanChunk entries and publishes
// the chunk total; k_cell_ranges sums the totals of the chunks before its own.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;  // 2048 entries per block

// ---------------------------------------------------------------- a5 + a6 as a counting sort
// The key is a dense cell index, so the sort of core.cpp:173-183 plus the run-length table of
// core.cpp:185-203 is a counting sort: count per cell, exclusive scan = cell starts (= the table),
// place.  The reference order inside a cell is ascending particle id (core.cpp:182); atomics hand
// out arrival slots in arbitrary order, so the last kernel ranks every particle among the (few)
// members of its cell by id — O(occupancy) reads per particle — and writes the final slot together
// with the gathered positions.  5 launches; the 3-pass stable radix sort + run-length table this
// replaced took 14 (same bits, 105 -> 55 us at 1 M particles).  In slab mode "particle id" is the
// GLOBAL id (gid): the storage order of a slab's particles is then irrelevant, which is what lets
// migration fill holes instead of re-packing the slab (kernels/slab.cu).
__global__ void __launch_bounds__(kThreads)
k_cell_count(const float4* __restrict__ pred_o, uint32_t* __restrict__ keys, uint32_t* __restrict__ arrival,
             uint32_t* __restrict__ cell_count, float inv_h, const GridDesc* __restrict__ desc,
             unsigned long long* __restrict__ cell_key, const StatusBlock* st, NRef nr) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k;
  if (desc->sparse) {
    const float4 q = pred_o[i];
    k = sparse_insert(cell_key, desc->ncells - 1u,
                      pack_cell((long long)cell_coord(q.x, inv_h) - desc->lo[0], (long long)cell_coord(q.y, inv_h) - desc->lo[1],
                                (long long)cell_coord(q.z, inv_h) - desc->lo[2]));
  } else {
    k = dense_key(pred_o[i], inv_h, *desc);
  }
  keys[i] = k;
  arrival[i] = atomicAdd(&cell_count[k], 1u);
}

// chunk-local exclusive scan of the cell counts (counts stay), chunk totals for the second level
__global__ void __launch_bounds__(kScanThreads)
k_cell_scan(const uint32_t* __restrict__ cell_count, uint32_t* __restrict__ cell_excl,
            uint32_t* __restrict__ chunk_total, const GridDesc* __restrict__ desc, const StatusBlock* st) {
  pdl_wait();
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  if (batch_failed(st)) return;
  const int m = (int)desc->ncells;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  if (blockIdx.x * kScanChunk >= m) {  // beyond the table: contributes nothing to the second level
    if (threadIdx.x == 0) chunk_total[blockIdx.x] = 0;
    return;
  }
  uint32_t v[kScanItems];
  uint32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (i0 + k < m) ? cell_count[i0 + k] : 0u;
    tsum += v[k];
  }
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t warp_excl = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const uint32_t ws = warp_sums[w];
    if (w < warp) warp_excl += ws;
    total += ws;
  }
  uint32_t excl = warp_excl + (incl - tsum);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (i0 + k < m) cell_excl[i0 + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == 0) chunk_total[blockIdx.x] = total;
}

// (start, end) of every table cell; the counters are zeroed for the next substep.  One block per scan
// chunk; the second scan level is folded in: the block sums the totals of the chunks before its own
// (<= cell_cap / 2048 values from L2) instead of a separate one-block kernel turning them into a prefix.
__global__ void __launch_bounds__(kThreads)
k_cell_ranges(uint32_t* __restrict__ cell_count, const uint32_t* __restrict__ cell_excl,
              const uint32_t* __restrict__ chunk_total, int2* __restrict__ cell_range,
              const GridDesc* __restrict__ desc, const StatusBlock* st) {
  pdl_wait();
  if (batch_failed(st)) return;
  const uint32_t ncells = desc->ncells;
  const uint32_t first = blockIdx.x * (uint32_t)kScanChunk;
  if (first >= ncells) return;
  __shared__ uint32_t warp_sums[kThreads / 32];
  __shared__ uint32_t s_prefix;
  uint32_t part = 0;
  for (uint32_t k = threadIdx.x; k < blockIdx.x; k += kThreads) part += chunk_total[k];
  part = __reduce_add_sync(0xffffffffu, part);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) sum += warp_sums[w];
    s_prefix = sum;
  }
  __syncthreads();
  const uint32_t prefix = s_prefix;
  const uint32_t last = min(first + (uint32_t)kScanChunk, ncells);
  for (uint32_t c = first + threadIdx.x; c < last; c += kThreads) {
    const uint32_t cnt = cell_count[c];
    const uint32_t start = cell_excl[c] + prefix;
    cell_range[c] = make_int2((int)start, (int)(start + cnt));
    if (cnt) cell_count[c] = 0;
  }
}

__global__ void __launch_bounds__(kThreads)
k_cell_place(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ arrival,
             const int2* __restrict__ cell_range, uint32_t* __restrict__ slot_id, const StatusBlock* st, NRef nr) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  slot_id[(uint32_t)cell_range[keys[i]].x + arrival[i]] = (uint32_t)i;
}

// Thread per (unordered) slot: final slot = cell start + number of cell members with a smaller id.
__global__ void __launch_bounds__(kThreads)
k_cell_order(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ slot_id,
             const int2* __restrict__ cell_range, const float4* __restrict__ pred_o,
             const float4* __restrict__ pos_o, uint32_t* __restrict__ keys_sorted, uint32_t* __restrict__ vals_sorted,
             float4* __restrict__ pred_s, float4* __restrict__ pos_s, const uint32_t* __restrict__ gid,
             const StatusBlock* st, NRef nr) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t id = slot_id[s];
  const uint32_t key = keys[id];
  const int2 r = cell_range[key];
  int rank = 0;
  if (gid) {  // slab: the reference's particle id is the global one
    const uint32_t g = gid[id];
    for (int t = r.x; t < r.y; ++t) rank += gid[slot_id[t]] < g ? 1 : 0;
  } else {
    for (int t = r.x; t < r.y; ++t) rank += slot_id[t] < id ? 1 : 0;
  }
  const int dst = r.x + rank;
  const float4 q = pred_o[id];
  const float4 p = pos_o[id];
  keys_sorted[dst] = key;
  vals_sorted[dst] = id;
  pred_s[dst] = make_float4(q.x, q.y, q.z, 0.0f);
  pos_s[dst] = make_float4(p.x, p.y, p.z, __uint_as_float(id));
}

// ---------------------------------------------------------------- a7 neighbour list
// Thread per sorted particle; candidates in the reference order: dz, dy, dx with dx
// innermost (core.cpp:211-213), ascending slot (== ascending particle id) inside a cell.
// Two candidates are tested per step with packed f32x2 arithmetic (FADD2 / FMUL2; the
// sums of products use scalar adds, see pbf_device.cuh).  Entry k of particle i lives at
// idx[(i/32)*K*32 + (k/2)*64 + (i%32)*2 + k%2]: a solver pass reads two entries per lane
// with one 8-byte load, 256 contiguous bytes per warp.
// The kernel is issue-bound (73 % issue-active, 19 of 32 lanes busy on average: lanes of a warp
// sit in ~5 different cells whose stencil cells hold different numbers of particles).  Three
// restructurings that remove the divergence (flattened walk, warp per cell, tests in memory order
// + bitmasks) were measured and are slower, see DESIGN.md §4.
//
// Two per-cell loops exist.  neighbors_cell (PBF_NBR_MASK=0, kept for A/B runs) tests a pair of
// candidates and stores the hits inside the same loop; neighbors_cell_mask (the default) first
// collects the hits of a cell in a bit mask and then emits them.  Same list, entry for entry.
#ifndef PBF_NBR_MASK
#define PBF_NBR_MASK 1
#endif
#ifndef PBF_NBR_HOIST
#define PBF_NBR_HOIST 1
#endif
#ifndef PBF_NBR_MASK_UNROLL
#define PBF_NBR_MASK_UNROLL 1
#endif
#if !PBF_NBR_MASK
struct NbrEmit {
  uint32_t* out;
  uint32_t cnt, off;  // entries so far; element offset of entry `cnt`: (cnt / 2) * 64 + cnt % 2
  uint32_t K;
  __device__ __forceinline__ void put(bool hit, int j) {
    if (hit && cnt < K) __stcs(out + off, (uint32_t)j);
    off += hit ? ((cnt & 1u) ? 63u : 1u) : 0u;
    cnt += hit ? 1u : 0u;
  }
};

template <bool CENTER>
__device__ __forceinline__ void neighbors_cell(const float4* __restrict__ pred_s, int2 range, float4 pi, f2 pxy, int i,
                                               float h2, NbrEmit& e) {
  for (int j = range.x; j < range.y; j += 2) {
    const bool v1 = (j + 1) < range.y;
    const float4 a0 = pred_s[j];
    const float4 a1 = pred_s[v1 ? j + 1 : j];
    // x and y of ONE candidate share an f32x2 (they sit in an aligned register pair after
    // the 16-byte load: no packing moves), z is scalar; same roundings as (dx*dx + dy*dy) + dz*dz
    const f2 d0 = __fadd2_rn(pxy, make_float2(-a0.x, -a0.y));
    const f2 d1 = __fadd2_rn(pxy, make_float2(-a1.x, -a1.y));
    const float z0 = __fsub_rn(pi.z, a0.z), z1 = __fsub_rn(pi.z, a1.z);
    const f2 q0 = __fmul2_rn(d0, d0), q1 = __fmul2_rn(d1, d1);
    const float r2a = __fadd_rn(__fadd_rn(q0.x, q0.y), __fmul_rn(z0, z0));
    const float r2b = __fadd_rn(__fadd_rn(q1.x, q1.y), __fmul_rn(z1, z1));
    bool h0 = r2a < h2, h1 = v1 && (r2b < h2);  // core.cpp:231-240
    if (CENTER) {
      h0 = h0 && (j != i);
      h1 = h1 && (j + 1 != i);
    }
    e.put(h0, j);
    e.put(h1, j + 1);
  }
}
#endif

// The test loop lives in neighbor_test.cuh (shared with the brick kernel).  List cursor of the mask
// variant: a pointer to the next entry instead of an element offset (two instructions per hit for
// the address instead of four).
struct NbrCursor {
  uint32_t* p;    // entry `cnt` of this lane's list
  uint32_t cnt;
  uint32_t K;
  __device__ __forceinline__ void put(int j) {
    const uint32_t odd = cnt & 1u;
    if (cnt < K) __stcs(p, (uint32_t)j);
    p += 1u + 62u * odd;  // entry k at (k / 2) * 64 + k % 2
    cnt += 1u;
  }
};

// (an explicit minimum of 1 block per SM is not the same as none: ptxas then takes 52 registers
// instead of 40)
#ifdef PBF_NBR_MINBLOCKS
__global__ void __launch_bounds__(128, PBF_NBR_MINBLOCKS)
#else
__global__ void __launch_bounds__(128)
#endif
k_neighbors(const float4* __restrict__ pred_s, const int2* __restrict__ cell_range,
            const GridDesc* __restrict__ desc, const unsigned long long* __restrict__ cell_key,
            uint32_t* __restrict__ nbr_idx, uint32_t* __restrict__ nbr_count, StatusBlock* st, float inv_h, float h2,
            int K, NRef nr) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int n = nr.get();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cnt = 0;
  bool active = i < n;
  int cx = 0, cy = 0, cz = 0;
  float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
  const int dimy = desc->dim[1], dimz = desc->dim[2];
  if (active) {
    pi = pred_s[i];
    // Owned particles always sit at least one layer inside the table.  Ghosts of the outermost
    // layer (or outside the table) can not touch an owned particle: they get an empty list.
    const long long rx = (long long)cell_coord(pi.x, inv_h) - desc->lo[0];
    const long long ry = (long long)cell_coord(pi.y, inv_h) - desc->lo[1];
    const long long rz = (long long)cell_coord(pi.z, inv_h) - desc->lo[2];
    active = rx >= 1 && rx <= desc->dim[0] - 2 && ry >= 1 && ry <= dimy - 2 && rz >= 1 && rz <= dimz - 2;
    cx = (int)rx; cy = (int)ry; cz = (int)rz;
    if (!active) nbr_count[i] = 0;
  }
  if (active) {
#if PBF_NBR_MASK
    NbrCursor e;
    e.p = nbr_idx + (size_t)(i >> 5) * (size_t)K * 32u + (uint32_t)(i & 31) * 2u;
    e.cnt = 0;
    e.K = (uint32_t)K;
#if PBF_NBR_HOIST
    // keep h2 in a register: ptxas otherwise re-loads it from the constant bank in every step of
    // the test loop (1 of 27 issue slots).  blockIdx.x >> 31 is 0 and x + 0.0f == x for x > 0.
    // (The same trick on the pred_s pointer makes ptxas rebuild the address with 4 instructions.)
    const float4* const ps = pred_s;
    const float h2r = __fadd_rn(h2, __uint_as_float(blockIdx.x >> 31));
#else
    const float4* const ps = pred_s;
    const float h2r = h2;
#endif
#else
    NbrEmit e;
    e.out = nbr_idx + (size_t)(i >> 5) * (size_t)K * 32u + (uint32_t)(i & 31) * 2u;
    e.cnt = 0;
    e.off = 0;
    e.K = (uint32_t)K;
#endif
    const uint32_t xstride = (uint32_t)dimy * (uint32_t)dimz;
    const bool sparse = desc->sparse != 0;
    const uint32_t mask = desc->ncells - 1u;
    const f2 pxy = make_float2(pi.x, pi.y);
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll 1
      for (int dy = -1; dy <= 1; ++dy) {
        int2 r0, r1, r2;
        if (sparse) {
          r0 = sparse_lookup(cell_key, cell_range, mask, cx - 1, cy + dy, cz + dz);
          r1 = sparse_lookup(cell_key, cell_range, mask, cx, cy + dy, cz + dz);
          r2 = sparse_lookup(cell_key, cell_range, mask, cx + 1, cy + dy, cz + dz);
        } else {
          const uint32_t row = ((uint32_t)(cx - 1) * (uint32_t)dimy + (uint32_t)(cy + dy)) * (uint32_t)dimz +
                               (uint32_t)(cz + dz);
          r0 = cell_range[row];
          r1 = cell_range[row + xstride];
          r2 = cell_range[row + 2u * xstride];
        }
#if PBF_NBR_MASK
        neighbors_cell_mask<false>(ps, r0, pi.z, pxy, i, h2r, e);
        if (dz == 0 && dy == 0)
          neighbors_cell_mask<true>(ps, r1, pi.z, pxy, i, h2r, e);
        else
          neighbors_cell_mask<false>(ps, r1, pi.z, pxy, i, h2r, e);
        neighbors_cell_mask<false>(ps, r2, pi.z, pxy, i, h2r, e);
#else
        neighbors_cell<false>(pred_s, r0, pi, pxy, i, h2, e);
        if (dz == 0 && dy == 0)
          neighbors_cell<true>(pred_s, r1, pi, pxy, i, h2, e);
        else
          neighbors_cell<false>(pred_s, r1, pi, pxy, i, h2, e);
        neighbors_cell<false>(pred_s, r2, pi, pxy, i, h2, e);
#endif
      }
    cnt = e.cnt;
    nbr_count[i] = cnt < (uint32_t)K ? cnt : (uint32_t)K;
  }
  // batch statistic: max count (to size K)
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) {
    if (wmax > *(volatile unsigned int*)&st->max_neighbors) atomicMax(&st->max_neighbors, wmax);
    if (wmax > (uint32_t)K) st->nbr_overflow = 1;
  }
}

inline int grid_for(int n, int threads) { return (n + threads - 1) / threads; }

}  // namespace

// Final positions of a substep straight from the sorted order into the download staging (the
// cuda_step contract path starts their D2H while XSPH / vorticity still run).
__global__ void k_scatter_positions(const float4* __restrict__ pos_sorted, const float4* __restrict__ pos_s,
                                    float* __restrict__ x, float* __restrict__ y, float* __restrict__ z, int n) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pos_sorted[i];
  const uint32_t orig = __float_as_uint(pos_s[i].w);
  x[orig] = p.x;
  y[orig] = p.y;
  z[orig] = p.z;
}

// ================================================================== launchers
int launch_scatter_positions(const float4* pos_sorted, const float4* pos_s, float* x, float* y, float* z, int n,
                             cudaStream_t s) {
  if (n <= 0) return 0;
  PBF_LAUNCH(k_scatter_positions, grid_for(n, kThreads), kThreads, s, pos_sorted, pos_s, x, y, z, n);
  return 1;
}

int launch_pack_state(const float* const soa[6], float4* pos_o, float4* vel_o, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  PBF_LAUNCH(k_pack_state, grid_for(n, kThreads), kThreads, s, soa[0], soa[1], soa[2], soa[3], soa[4], soa[5],
                                                         pos_o, vel_o, n);
  return 1;
}

int launch_unpack_state(const float4* pos_o, const float4* vel_o, float* const soa[6], int n, cudaStream_t s) {
  if (n <= 0) return 0;
  PBF_LAUNCH(k_unpack_state, grid_for(n, kThreads), kThreads, s, pos_o, vel_o, soa[0], soa[1], soa[2], soa[3],
                                                           soa[4], soa[5], n);
  return 1;
}

int launch_predict(float4* pos_o, float4* vel_o, float4* pred_o, const StepConsts& c,
                   const GridBuffers& g, NRef n, bool slab, cudaStream_t s) {
  // one particle per thread up to 16 waves of 148 SMs x 8 blocks: enough loads in flight to
  // stream at HBM speed, few enough warps that the six bound atomics stay cheap
  int blocks = grid_for(n.n, kThreads);
  if (blocks > 148 * 8 * 16) blocks = 148 * 8 * 16;
  // The table descriptor follows as a one-thread kernel.  (Writing it from the last block of
  // k_predict — `finalize` = 1, an atomic ticket per block — was MEASURED on B200: the predict stage
  // went from 11.6 to 24.4 us at 1 M particles, because every block then fences its 256 pred stores
  // before it takes the ticket; the kernel launch it saves costs 3 us.  So it stays off.)
  PBF_LAUNCH(k_predict, blocks, kThreads, s, pos_o, vel_o, pred_o, c, g.status, n, 1, g.desc, g.cell_key, g.cell_cap,
                                      0, 0);
  if (slab) return 1;  // migrants extend the bounds; the table descriptor follows (launch_grid_finalize)
  PBF_LAUNCH(k_grid_finalize, 1, 1, s, g.desc, g.status, g.cell_cap, 1, n, 1, g.bricks ? g.brick_cap : 0);
  return 2;
}

int launch_grid_finalize(const GridBuffers& g, int pad, NRef n, cudaStream_t s) {
  PBF_LAUNCH(k_grid_finalize, 1, 1, s, g.desc, g.status, g.cell_cap, pad, n, 0, 0);
  return 1;
}

int launch_sort(const float4* pred_o, const StepConsts& c, const GridBuffers& g, NRef n, int* out,
                cudaStream_t s) {
  // keys[0] = key per particle, vals[0] = arrival slot inside its cell; the ordered result
  // (keys[1], vals[1]) is written by launch_cells_reorder
  PBF_LAUNCH(k_cell_count, grid_for(n.n, kThreads), kThreads, s, pred_o, g.keys[0], g.vals[0], g.cell_count, c.inv_h,
             g.desc, g.cell_key, g.status, n);
  const int nchunks = (int)((g.cell_cap + kScanChunk - 1) / kScanChunk);
  PBF_LAUNCH(k_cell_scan, nchunks, kScanThreads, s, g.cell_count, g.cell_excl, g.chunk_total, g.desc, g.status);
  PBF_LAUNCH(k_cell_ranges, nchunks, kThreads, s, g.cell_count, g.cell_excl, g.chunk_total, g.cell_range, g.desc,
                                          g.status);
  PBF_LAUNCH(k_cell_place, grid_for(n.n, kThreads), kThreads, s, g.keys[0], g.vals[0], g.cell_range, g.slot_id,
                                                           g.status, n);
  *out = 1;
  return 4;
}

int launch_cells_reorder(const float4* pred_o, const float4* pos_o, float4* pred_s, float4* pos_s,
                         const uint32_t* gid, const GridBuffers& g, NRef n, cudaStream_t s) {
  PBF_LAUNCH(k_cell_order, grid_for(n.n, kThreads), kThreads, s, g.keys[0], g.slot_id, g.cell_range, pred_o, pos_o,
                                                           g.keys[1], g.vals[1], pred_s, pos_s, gid, g.status, n);
  return 1;
}

int launch_neighbors(const float4* pred_s, const StepConsts& c, const GridBuffers& g,
                     const NeighborList& nl, NRef n, cudaStream_t s) {
  PBF_LAUNCH(k_neighbors, grid_for(n.n, 128), 128, s, pred_s, g.cell_range, g.desc, g.cell_key, nl.idx, nl.count,
             g.status, c.inv_h, c.h2, nl.K, n);
  return 1;
}

}  // namespace pbf

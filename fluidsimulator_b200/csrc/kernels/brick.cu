// brick.cu — the CTA-brick family of the PBF substep for sm_100a: neighbour build and every pass that
// walks the neighbour list, as ONE CTA PER BRICK of grid cells with the brick's neighbourhood staged in
// shared memory (north_star: "neighbor gathers staged through shared memory").
//
//   a7  neighbour list              (reference core/src/core.cpp:205-247)   k_neighbors_brick
//   a8  lambda                      (core.cpp:281-329)                      k_lambda_brick
//   a9  delta-p (+ a10, a11, a14)   (core.cpp:334-421)                      k_delta_brick
//   a12 XSPH                        (core.cpp:423-466)                      k_xsph_brick
//   a13 vorticity                   (core.cpp:468-571)                      k_vort_omega_brick, k_vort_apply_brick
//
// Why: the global-gather family (solve.cu) is bound by the L1 tag stage — a warp-wide 16-byte gather
// touches ~14.4 different 128-byte lines (profiles/ncu_r02a_*: l1tex 72 % at 66 % issue, XSPH 87 % at
// 29 %) — and re-streams a 4-byte list entry per neighbour per pass.  Here
//   * the dense cell table is tiled by bricks of kBrickX x kBrickY z-columns x kBrickZ cells.  In the
//     x-major sorted order a z-column is a contiguous run, so a brick plus one halo cell layer is
//     (kBrickX+2)(kBrickY+2) contiguous runs: k_brick_table records them once per substep and every
//     pass copies them into its tile with one cp.async.bulk (TMA, SASS UBLKCP) per run, completed on
//     one mbarrier.  L2 -> SM traffic per pass drops from ~30 scattered 16-byte gathers per particle
//     to ~3 coalesced records per particle;
//   * gathers are LDS.128 at tile + entry, where a list entry is the 16-bit BYTE OFFSET of the
//     neighbour's record in the tile (record index * 16 <= 65520): half the list bytes, one ALU
//     instruction per neighbour for the address.  tools/analysis/brick_sim.py counts 7.9 shared-memory
//     wavefronts per warp gather on the reference's own lists against 14.4 L1 lines;
//   * the summation order is untouched (only the address space of an entry changes), so STRICT stays
//     bit-identical to the reference.
// A substep whose tiles do not fit (kTileCap records), whose brick table is too small, or that runs
// on the sparse cell table raises StatusBlock::brick_overflow; pbf_step restores the batch, grows
// the table or switches that batch to the global-gather family, and replays (pbf_capi.cu).
#include "neighbor_test.cuh"
#include "pbf_kernels.h"
#include "solve_passes.cuh"

namespace pbf {

namespace {

#ifndef PBF_BRICK_THREADS
#define PBF_BRICK_THREADS 256
#endif
#ifndef PBF_BRICK_THREADS2
#define PBF_BRICK_THREADS2 256  // passes with two tiles (XSPH, omega)
#endif
#ifndef PBF_BRICK_MINBLOCKS
#define PBF_BRICK_MINBLOCKS 1
#endif
constexpr int kBT = PBF_BRICK_THREADS;
constexpr int kBT2 = PBF_BRICK_THREADS2;
// + 1: the neighbour test loads the record after a cell's last candidate unconditionally
constexpr size_t kTileBytes = (size_t)(kTileCap + 1) * sizeof(float4);
constexpr int kCrX = kBrickX + 2, kCrY = kBrickY + 2, kCrZ = kBrickZ + 2;

// ---- TMA (cp.async.bulk) + mbarrier ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
// global -> shared bulk copy; 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds128(const unsigned char* tile, uint32_t off) {
  return *reinterpret_cast<const float4*>(tile + off);
}

// ---- brick table --------------------------------------------------------------------------------
// One block per brick, thread c < kBrickCols per halo column: the column's cells z0-1 .. z0+kBrickZ
// are consecutive table entries, their particles one contiguous run of sorted slots (checked: a
// column whose non-empty cells do not add up to the run disables the brick path for the batch).
__global__ void __launch_bounds__(64)
k_brick_table(const int2* __restrict__ cell_range, const GridDesc* __restrict__ desc, BrickRec* __restrict__ bricks,
              StatusBlock* st) {
  pdl_wait();
  if (batch_failed(st)) return;
  const int b = blockIdx.x;
  if (b >= desc->nbricks) return;
  __shared__ int s_start[kBrickCols], s_len[kBrickCols], s_base[kBrickCols + 1];
  __shared__ int s_ostart[kBrickOwn], s_olen[kBrickOwn], s_opre[kBrickOwn + 1];
  __shared__ int s_bad;
  const int dimx = desc->dim[0], dimy = desc->dim[1], dimz = desc->dim[2];
  const int nbz = desc->bdim[2], nby = desc->bdim[1];
  const int bk = b % nbz, bj = (b / nbz) % nby, bi = b / (nbz * nby);
  const int x0 = bi * kBrickX, y0 = bj * kBrickY, z0 = bk * kBrickZ;
  const int t = threadIdx.x;
  if (t == 0) s_bad = 0;
  __syncthreads();
  if (t < kBrickCols) {
    const int ix = t / kCrY, iy = t % kCrY;
    const int X = x0 - 1 + ix, Y = y0 - 1 + iy;
    int start = 0, len = 0, ostart = 0, olen = 0;
    if (X >= 0 && X < dimx && Y >= 0 && Y < dimy) {
      const int2* col = cell_range + ((size_t)X * (size_t)dimy + (size_t)Y) * (size_t)dimz;
      const int zlo = max(z0 - 1, 0), zhi = min(z0 + kBrickZ, dimz - 1);
      const int olo = z0, ohi = min(z0 + kBrickZ, dimz) - 1;
      int first = -1, last = 0, sum = 0, ofirst = -1, olast = 0;
      for (int z = zlo; z <= zhi; ++z) {
        const int2 r = col[z];
        if (r.y > r.x) {
          if (first < 0) first = r.x;
          if (r.x != first + sum) s_bad = 1;  // the cells of a column must follow each other
          last = r.y;
          sum += r.y - r.x;
          if (z >= olo && z <= ohi) {
            if (ofirst < 0) ofirst = r.x;
            olast = r.y;
          }
        }
      }
      if (first >= 0) {
        start = first;
        len = last - first;
        if (len != sum) s_bad = 1;
      }
      if (ofirst >= 0) {
        ostart = ofirst;
        olen = olast - ofirst;
      }
    }
    s_start[t] = start;
    s_len[t] = len;
    const bool owned = ix >= 1 && ix <= kBrickX && iy >= 1 && iy <= kBrickY;
    if (owned) {
      const int r = (ix - 1) * kBrickY + (iy - 1);
      s_ostart[r] = ostart;
      s_olen[r] = olen;
    }
  }
  __syncthreads();
  if (t == 0) {
    int acc = 0;
    for (int c = 0; c < kBrickCols; ++c) {
      s_base[c] = acc;
      acc += s_len[c];
    }
    s_base[kBrickCols] = acc;
    int oacc = 0;
    for (int r = 0; r < kBrickOwn; ++r) {
      s_opre[r] = oacc;
      oacc += s_olen[r];
    }
    s_opre[kBrickOwn] = oacc;
    if (oacc > 0) {  // only tiles that a CTA will stage count
      if ((unsigned)acc > *(volatile unsigned int*)&st->max_tile) atomicMax(&st->max_tile, (unsigned)acc);
      if (acc > kTileCap || s_bad) atomicOr(&st->brick_overflow, kBrickDisable);
    }
  }
  __syncthreads();
  BrickRec* rec = bricks + b;
  if (t < kBrickCols) {
    rec->col_start[t] = s_start[t];
    rec->col_base[t] = s_base[t];
  }
  if (t < kBrickOwn) {
    rec->own_start[t] = s_ostart[t];
    rec->own_prefix[t] = s_opre[t];
  }
  if (t == 0) {
    rec->col_base[kBrickCols] = s_base[kBrickCols];
    rec->own_prefix[kBrickOwn] = s_opre[kBrickOwn];
    rec->tile_n = s_base[kBrickCols];
    rec->own_n = s_opre[kBrickOwn];
  }
}

// ---- per-CTA prologue ---------------------------------------------------------------------------
struct BrickShared {
  BrickRec rec;
  unsigned long long bar;
};

// Loads the brick record and arms the mbarrier.  false (for the whole CTA) when there is nothing to do.
__device__ __forceinline__ bool brick_begin(BrickShared& sh, const BrickRec* __restrict__ bricks,
                                            const GridDesc* __restrict__ desc, const StatusBlock* st) {
  if (batch_failed(st)) return false;
  const int b = blockIdx.x;
  if (b >= desc->nbricks) return false;
  if (bricks[b].own_n == 0) return false;
  const int* src = reinterpret_cast<const int*>(bricks + b);
  int* dst = reinterpret_cast<int*>(&sh.rec);
  for (int w = threadIdx.x; w < (int)(sizeof(BrickRec) / sizeof(int)); w += blockDim.x) dst[w] = src[w];
  if (threadIdx.x == 0) mbar_init(&sh.bar, 1);
  __syncthreads();
  return true;
}

// One bulk copy per halo column, issued by kBrickCols threads at once.
__device__ __forceinline__ void brick_stage(BrickShared& sh, unsigned char* tile, const float4* __restrict__ src) {
  if (threadIdx.x < kBrickCols) {
    const int c = threadIdx.x;
    const int base = sh.rec.col_base[c], len = sh.rec.col_base[c + 1] - base;
    if (len > 0) bulk_g2s(tile + (size_t)base * 16u, src + sh.rec.col_start[c], (uint32_t)len * 16u, &sh.bar);
  }
}

// Owned particle number t of the brick: its sorted slot and the byte offset of its record in the tile.
struct Owned {
  int i;
  uint32_t off;
  int r;  // owned column
};
__device__ __forceinline__ Owned brick_owned(const BrickRec& rec, int t) {
  int r = 0;
#pragma unroll
  for (int s = kBrickOwn / 2; s >= 1; s >>= 1)
    if (rec.own_prefix[r + s] <= t) r += s;
  Owned o;
  o.r = r;
  o.i = rec.own_start[r] + (t - rec.own_prefix[r]);
  const int col = (r / kBrickY + 1) * kCrY + (r % kBrickY) + 1;
  o.off = (uint32_t)(rec.col_base[col] + (o.i - rec.col_start[col])) * 16u;
  return o;
}

// ---- list walk ----------------------------------------------------------------------------------
// Entry k of slot i: ((uint16_t*)idx)[(i/32)*K*32 + (k/4)*128 + (i%32)*4 + k%4] — one 8-byte load per
// lane brings four neighbours, 256 contiguous bytes per warp of consecutive slots.  Streaming loads
// (ld.global.cs), the next quad requested before the current one is gathered.
__device__ __forceinline__ const uint2* list_row(const uint16_t* __restrict__ nbr, int K, int i) {
  return reinterpret_cast<const uint2*>(nbr + (size_t)(i >> 5) * (size_t)K * 32u) + (i & 31);
}

// body(a0, a1, v1): two neighbours and the validity of the second; `self` pads an odd tail (finite
// values whose terms the passes zero out).
template <typename Body>
__device__ __forceinline__ void tile_pairs(const uint2* __restrict__ row, uint32_t cnt, uint2 first,
                                           const unsigned char* tile, float4 self, Body&& body) {
  const uint32_t nquads = cnt >> 2, nq_all = (cnt + 3u) >> 2;
  uint2 jn = first;
  for (uint32_t q = 0; q < nquads; ++q) {
    const uint2 j = jn;
    if (q + 1 < nq_all) jn = __ldcs(row + (size_t)(q + 1) * 32u);
    const float4 a0 = lds128(tile, j.x & 0xffffu), a1 = lds128(tile, j.x >> 16);
    const float4 a2 = lds128(tile, j.y & 0xffffu), a3 = lds128(tile, j.y >> 16);
    body(a0, a1, true);
    body(a2, a3, true);
  }
  const uint32_t rem = cnt & 3u;
  if (rem) {
    const float4 a0 = lds128(tile, jn.x & 0xffffu);
    const float4 a1 = rem >= 2 ? lds128(tile, jn.x >> 16) : self;
    body(a0, a1, rem >= 2);
    if (rem == 3) body(lds128(tile, jn.y & 0xffffu), self, false);
  }
}

// Two arrays per neighbour: tiles A and B share the layout (same byte offsets).
template <typename Body>
__device__ __forceinline__ void tile_pairs2(const uint2* __restrict__ row, uint32_t cnt, uint2 first,
                                            const unsigned char* ta, const unsigned char* tb, float4 self_a,
                                            float4 self_b, Body&& body) {
  const uint32_t nquads = cnt >> 2, nq_all = (cnt + 3u) >> 2;
  uint2 jn = first;
  for (uint32_t q = 0; q < nquads; ++q) {
    const uint2 j = jn;
    if (q + 1 < nq_all) jn = __ldcs(row + (size_t)(q + 1) * 32u);
    const uint32_t o0 = j.x & 0xffffu, o1 = j.x >> 16, o2 = j.y & 0xffffu, o3 = j.y >> 16;
    const float4 a0 = lds128(ta, o0), a1 = lds128(ta, o1), b0 = lds128(tb, o0), b1 = lds128(tb, o1);
    const float4 a2 = lds128(ta, o2), a3 = lds128(ta, o3), b2 = lds128(tb, o2), b3 = lds128(tb, o3);
    body(a0, a1, b0, b1, true);
    body(a2, a3, b2, b3, true);
  }
  const uint32_t rem = cnt & 3u;
  if (rem) {
    const uint32_t o0 = jn.x & 0xffffu, o1 = jn.x >> 16, o2 = jn.y & 0xffffu;
    const float4 a0 = lds128(ta, o0), b0 = lds128(tb, o0);
    const float4 a1 = rem >= 2 ? lds128(ta, o1) : self_a, b1 = rem >= 2 ? lds128(tb, o1) : self_b;
    body(a0, a1, b0, b1, rem >= 2);
    if (rem == 3) body(lds128(ta, o2), self_a, lds128(tb, o2), self_b, false);
  }
}

// Walks the owned particles of the brick, kThreads at a time.  The list row, count and first quad
// of a thread's NEXT particle are requested before its current one is processed, so the dependent
// chain count -> list -> gather is not exposed once per round.
template <int kThreads, typename Fn>
__device__ __forceinline__ void for_each_owned(const BrickRec& rec, const uint16_t* __restrict__ nbr,
                                               const uint32_t* __restrict__ nbr_count, int K, Fn&& fn) {
  const int own_n = rec.own_n;
  int t = threadIdx.x;
  if (t >= own_n) return;
  Owned o = brick_owned(rec, t);
  const uint2* row = list_row(nbr, K, o.i);
  uint32_t cnt = nbr_count[o.i];
  uint2 first = __ldcs(row);
  for (;;) {
    const int tn = t + kThreads;
    const bool more = tn < own_n;
    Owned on = o;
    const uint2* rown = row;
    uint32_t cntn = 0;
    uint2 firstn = make_uint2(0u, 0u);
    if (more) {
      on = brick_owned(rec, tn);
      rown = list_row(nbr, K, on.i);
      cntn = nbr_count[on.i];
      firstn = __ldcs(rown);
    }
    fn(o, row, cnt, first);
    if (!more) break;
    o = on;
    row = rown;
    cnt = cntn;
    first = firstn;
    t = tn;
  }
}

// ---------------------------------------------------------------- a7 neighbour list
struct NbrCursor16 {
  uint16_t* p;  // entry `cnt` of this lane's list
  uint32_t cnt;
  uint32_t K;
  __device__ __forceinline__ void put(int j) {  // j = record index in the tile
    if (cnt < K) asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)(j << 4)) : "memory");
    p += ((cnt & 3u) == 3u) ? 125 : 1;  // entry k at (k / 4) * 128 + k % 4
    cnt += 1u;
  }
};

// Thread per owned particle; candidates in the reference order — dz, dy, dx with dx innermost
// (core.cpp:211-213), ascending slot inside a cell — read from the shared-memory tile; the cell
// table of the brick's halo is translated into tile coordinates once per CTA.  Same test loop as
// grid.cu (neighbor_test.cuh), 16-bit entries.
__global__ void __launch_bounds__(kBT, PBF_BRICK_MINBLOCKS)
k_neighbors_brick(const float4* __restrict__ pred_s, const int2* __restrict__ cell_range,
                  const GridDesc* __restrict__ desc, const BrickRec* __restrict__ bricks, uint16_t* __restrict__ nbr,
                  uint32_t* __restrict__ nbr_count, StatusBlock* st, float inv_h, float h2, int K) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ BrickShared sh;
  __shared__ int2 cr[kCrX * kCrY * kCrZ];
  pdl_wait();
  if (!brick_begin(sh, bricks, desc, st)) return;
  brick_stage(sh, dyn, pred_s);
  if (threadIdx.x == 0) mbar_arrive_expect_tx(&sh.bar, (uint32_t)sh.rec.tile_n * 16u);
  const int dimx = desc->dim[0], dimy = desc->dim[1], dimz = desc->dim[2];
  const int nbz = desc->bdim[2], nby = desc->bdim[1];
  const int b = blockIdx.x;
  const int bk = b % nbz, bj = (b / nbz) % nby, bi = b / (nbz * nby);
  const int x0 = bi * kBrickX, y0 = bj * kBrickY, z0 = bk * kBrickZ;
  for (int cc = threadIdx.x; cc < kCrX * kCrY * kCrZ; cc += kBT) {
    const int iz = cc % kCrZ, iy = (cc / kCrZ) % kCrY, ix = cc / (kCrZ * kCrY);
    const int X = x0 - 1 + ix, Y = y0 - 1 + iy, Z = z0 - 1 + iz;
    int2 r = make_int2(0, 0);
    if (X >= 0 && X < dimx && Y >= 0 && Y < dimy && Z >= 0 && Z < dimz) {
      const int2 g = cell_range[((size_t)X * (size_t)dimy + (size_t)Y) * (size_t)dimz + (size_t)Z];
      if (g.y > g.x) {
        const int col = ix * kCrY + iy;
        const int shift = sh.rec.col_base[col] - sh.rec.col_start[col];
        r = make_int2(g.x + shift, g.y + shift);
      }
    }
    cr[cc] = r;
  }
  __syncthreads();
  mbar_wait(&sh.bar, 0);
  const float4* tile = reinterpret_cast<const float4*>(dyn);
  const int own_n = sh.rec.own_n;
  const int lo_z = desc->lo[2];
  // keep h2 in a register (see k_neighbors, grid.cu)
  const float h2r = __fadd_rn(h2, __uint_as_float(blockIdx.x >> 31));
  for (int base = 0; base < own_n; base += kBT) {  // uniform trip count: the warp reduction below
    const int t = base + threadIdx.x;
    uint32_t cnt = 0;
    if (t < own_n) {
      const Owned o = brick_owned(sh.rec, t);
      const int ti = (int)(o.off >> 4);
      const float4 pi = tile[ti];
      const int ix = o.r / kBrickY + 1, iy = o.r % kBrickY + 1;
      const int iz = cell_coord(pi.z, inv_h) - lo_z - (z0 - 1);
      const int X = x0 - 1 + ix, Y = y0 - 1 + iy, Z = z0 - 1 + iz;
      // Owned particles sit at least one layer inside the table; particles of its outermost layer
      // (slab ghosts) can not touch an owned particle: they get an empty list.
      const bool active = X >= 1 && X <= dimx - 2 && Y >= 1 && Y <= dimy - 2 && Z >= 1 && Z <= dimz - 2 &&
                          iz >= 1 && iz <= kBrickZ;
      if (active) {
        NbrCursor16 e;
        e.p = nbr + (size_t)(o.i >> 5) * (size_t)K * 32u + (uint32_t)(o.i & 31) * 4u;
        e.cnt = 0;
        e.K = (uint32_t)K;
        const f2 pxy = make_float2(pi.x, pi.y);
#pragma unroll 1
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll 1
          for (int dy = -1; dy <= 1; ++dy) {
            const int c0 = ((ix - 1) * kCrY + (iy + dy)) * kCrZ + (iz + dz);
            const int2 r0 = cr[c0], r1 = cr[c0 + kCrY * kCrZ], r2 = cr[c0 + 2 * kCrY * kCrZ];
            neighbors_cell_mask<false>(tile, r0, pi.z, pxy, ti, h2r, e);
            if (dz == 0 && dy == 0)
              neighbors_cell_mask<true>(tile, r1, pi.z, pxy, ti, h2r, e);
            else
              neighbors_cell_mask<false>(tile, r1, pi.z, pxy, ti, h2r, e);
            neighbors_cell_mask<false>(tile, r2, pi.z, pxy, ti, h2r, e);
          }
        cnt = e.cnt;
      }
      nbr_count[o.i] = cnt < (uint32_t)K ? cnt : (uint32_t)K;
    }
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) {
      if (wmax > *(volatile unsigned int*)&st->max_neighbors) atomicMax(&st->max_neighbors, wmax);
      if (wmax > (uint32_t)K) st->nbr_overflow = 1;
    }
  }
}

// ================================================================== the passes that walk the list
// A pass is an Op: which arrays it stages (kTiles tiles with the same layout) and what it does for
// one owned particle.  Two drivers run an Op:
//   k_brick_once     one CTA per brick (brick v1): stage, wait, walk the owned particles kBT at a time;
//   k_brick_persist  persistent CTAs (brick v2): a ring of kPS tile slots per CTA filled by cp.async.bulk
//                    while the warps work on earlier slots; bricks are handed out through a global
//                    ticket, the owned particles of a brick 32 at a time through a shared-memory
//                    counter, so a warp never waits for the slowest warp of its CTA, the staging of
//                    brick k+1.. overlaps the arithmetic of brick k, and there is no per-brick tail.
template <bool S, bool SAFE>
struct LambdaOp {  // a8
  static constexpr int kTiles = 1;
  float4* pred;
  float* rho_out;
  StepConsts c;
  DebugPtrs dbg;
  __device__ __forceinline__ const float4* source(int) const { return pred; }
  __device__ __forceinline__ void particle(const Owned& o, const uint2* row, uint32_t cnt, uint2 first,
                                           const unsigned char* ta, const unsigned char*) const {
    const float4 pi = lds128(ta, o.off);
    LambdaPass<S, SAFE> acc(pi, c);
    tile_pairs(row, cnt, first, ta, pi, acc);
    float lambda, rho;
    acc.finish(lambda, rho);
    // only .w is written; the tiles other CTAs stage from pred use .xyz only in this pass
    reinterpret_cast<float*>(pred + o.i)[3] = lambda;
    rho_out[o.i] = rho;
    if (dbg.lambda) dbg.lambda[o.i] = lambda;
    if (dbg.rho) dbg.rho[o.i] = rho;
  }
};

template <bool S, bool LAST, bool COMMON>
struct DeltaOp {  // a9 + a10 (+ a11, a14)
  static constexpr int kTiles = 1;
  const float4* pred_in;
  float4* pred_out;
  const float4* pos_s;
  const float* rho;
  float4* vel_out;
  const float4* planes;
  float4* pos_o;
  float4* vel_o;
  StepConsts c;
  DebugPtrs dbg;
  HaloOut halo;
  int is_final;
  __device__ __forceinline__ const float4* source(int) const { return pred_in; }
  __device__ __forceinline__ void particle(const Owned& o, const uint2* row, uint32_t cnt, uint2 first,
                                           const unsigned char* ta, const unsigned char*) const {
    const float4 pi = lds128(ta, o.off);
    DeltaPass<S, COMMON> acc(pi, c);
    tile_pairs(row, cnt, first, ta, pi, acc);
    V3<FT<S>> np;
    float4 dlt;
    acc.finish(pi, planes, np, dlt);
    // the brick family's XSPH stages positions and velocities as two tiles: no 32-byte records
    delta_store<S, LAST>(o.i, np, dlt, pred_out, pos_s, rho, vel_out, (PosVel*)nullptr, planes, pos_o, vel_o, c, dbg,
                         halo, is_final);
  }
};

template <bool S>
struct XsphOp {  // a12
  static constexpr int kTiles = 2;
  const float4* pos;
  const float4* vel_in;
  float4* vel_out;
  const float4* pos_s;
  const float4* planes;
  float4* pos_o;
  float4* vel_o;
  StepConsts c;
  DebugPtrs dbg;
  HaloOut halo;
  int is_final;
  __device__ __forceinline__ const float4* source(int k) const { return k ? vel_in : pos; }
  __device__ __forceinline__ void particle(const Owned& o, const uint2* row, uint32_t cnt, uint2 first,
                                           const unsigned char* ta, const unsigned char* tb) const {
    using F = FT<S>;
    const float4 pi = lds128(ta, o.off), vi = lds128(tb, o.off);
    XsphPass<S> acc(pi, vi, c);
    tile_pairs2(row, cnt, first, ta, tb, pi, vi, acc);
    if (dbg.dv) dbg.dv[o.i] = make_float4(acc.sx, acc.sy, acc.sz, 0.0f);
    const V3<F> v = acc.finish(vi);
    if (is_final) {
      finalize_particle<F>(pi, v, __float_as_uint(pos_s[o.i].w), c, planes, pos_o, vel_o);
    } else {
      const float4 vo = make_float4(Arith<F>::val(v.x), Arith<F>::val(v.y), Arith<F>::val(v.z), vi.w);
      vel_out[o.i] = vo;
      halo.put(o.i, vo);
    }
  }
};

template <bool S>
struct OmegaOp {  // a13, pass 1
  static constexpr int kTiles = 2;
  float4* pos;
  const float4* vel;
  float4* omega;
  StepConsts c;
  __device__ __forceinline__ const float4* source(int k) const { return k ? vel : pos; }
  __device__ __forceinline__ void particle(const Owned& o, const uint2* row, uint32_t cnt, uint2 first,
                                           const unsigned char* ta, const unsigned char* tb) const {
    const float4 pi = lds128(ta, o.off), vi = lds128(tb, o.off);
    OmegaPass<S> acc(pi, vi, c);
    tile_pairs2(row, cnt, first, ta, tb, pi, vi, acc);
    const float4 om = acc.finish();
    omega[o.i] = om;
    // |omega_i| rides in pos[i].w for the eta pass; the tiles staged from pos use .xyz only here
    reinterpret_cast<float*>(pos + o.i)[3] = om.w;
  }
};

template <bool S>
struct EtaOp {  // a13 pass 2 + apply (+ a14)
  static constexpr int kTiles = 1;
  const float4* pos;
  const float4* vel;
  const float4* omega;
  const float4* pos_s;
  const float4* planes;
  float4* pos_o;
  float4* vel_o;
  StepConsts c;
  DebugPtrs dbg;
  __device__ __forceinline__ const float4* source(int) const { return pos; }
  __device__ __forceinline__ void particle(const Owned& o, const uint2* row, uint32_t cnt, uint2 first,
                                           const unsigned char* ta, const unsigned char*) const {
    using F = FT<S>;
    const float4 pi = lds128(ta, o.off);  // (pos xyz, |omega_i|)
    EtaPass<S> acc(pi, c);
    tile_pairs(row, cnt, first, ta, pi, acc);
    if (dbg.eta) dbg.eta[o.i] = make_float4(acc.ex, acc.ey, acc.ez, 0.0f);
    const V3<F> v = acc.finish(omega[o.i], vel[o.i]);
    finalize_particle<F>(pi, v, __float_as_uint(pos_s[o.i].w), c, planes, pos_o, vel_o);
  }
};

// ---- driver 1: one CTA per brick --------------------------------------------------------------
template <typename Op>
__global__ void __launch_bounds__(kBT, PBF_BRICK_MINBLOCKS)
k_brick_once(const __grid_constant__ Op op, const uint16_t* __restrict__ nbr, const uint32_t* __restrict__ nbr_count,
             const BrickRec* __restrict__ bricks, const GridDesc* __restrict__ desc, const StatusBlock* st, int K) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ BrickShared sh;
  pdl_wait();
  if (!brick_begin(sh, bricks, desc, st)) return;
  unsigned char* ta = dyn;
  unsigned char* tb = dyn + (Op::kTiles > 1 ? kTileBytes : 0);
#pragma unroll
  for (int k = 0; k < Op::kTiles; ++k) brick_stage(sh, dyn + (size_t)k * kTileBytes, op.source(k));
  if (threadIdx.x == 0) mbar_arrive_expect_tx(&sh.bar, (uint32_t)sh.rec.tile_n * 16u * Op::kTiles);
  mbar_wait(&sh.bar, 0);
  for_each_owned<kBT>(sh.rec, nbr, nbr_count, K, [&](const Owned& o, const uint2* row, uint32_t cnt, uint2 first) {
    op.particle(o, row, cnt, first, ta, tb);
  });
}

// ---- driver 2: persistent CTAs, ring of tile slots ---------------------------------------------
#ifndef PBF_PERSIST_THREADS
#define PBF_PERSIST_THREADS 512   // one-tile passes: 2 CTAs per SM
#endif
#ifndef PBF_PERSIST_THREADS2
#define PBF_PERSIST_THREADS2 768   // two-tile passes (XSPH, omega): 1 CTA per SM, up to 85 registers
#endif
#ifndef PBF_PERSIST_SLOTS
#define PBF_PERSIST_SLOTS 3
#endif
constexpr int kPS = PBF_PERSIST_SLOTS;

struct PSlot {
  BrickRec rec;
  unsigned long long full;  // mbarrier: record and tile(s) of the slot are in place (or the ring ran dry)
  int brick;                // brick index, -1 = no more work
  int next;                 // next owned particle of the brick to hand out
  int left;                 // warps that are done with the slot
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One warp (all lanes) puts the next non-empty brick into slot `sl`: ticket, record, bulk copies.
template <typename Op>
__device__ __forceinline__ void persist_refill(const Op& op, PSlot& sl, unsigned char* tiles,
                                               const BrickRec* __restrict__ bricks, int nbricks, unsigned int* ticket,
                                               int lane) {
  int b = -1;
  if (lane == 0) {
    for (;;) {
      const unsigned int t = atomicAdd(ticket, 1u);
      if (t >= (unsigned int)nbricks) break;
      if (bricks[t].own_n > 0) { b = (int)t; break; }
    }
  }
  b = __shfl_sync(0xffffffffu, b, 0);
  if (b < 0) {
    if (lane == 0) {
      sl.brick = -1;
      mbar_arrive(&sl.full);
    }
    return;
  }
  const int* src = reinterpret_cast<const int*>(bricks + b);
  int* dst = reinterpret_cast<int*>(&sl.rec);
  for (int w = lane; w < (int)(sizeof(BrickRec) / sizeof(int)); w += 32) dst[w] = src[w];
  if (lane == 0) {
    sl.brick = b;
    sl.next = 0;
    sl.left = 0;
  }
  __syncwarp();
  // the slot's previous tile was read through the generic proxy; the bulk copies write through the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (lane == 0) mbar_arrive_expect_tx(&sl.full, (uint32_t)sl.rec.tile_n * 16u * Op::kTiles);
  for (int c = lane; c < kBrickCols; c += 32) {
    const int base = sl.rec.col_base[c], len = sl.rec.col_base[c + 1] - base;
    if (len > 0) {
#pragma unroll
      for (int k = 0; k < Op::kTiles; ++k)
        bulk_g2s(tiles + (size_t)k * kTileBytes + (size_t)base * 16u, op.source(k) + sl.rec.col_start[c],
                 (uint32_t)len * 16u, &sl.full);
    }
  }
}

// ctl[0] = brick ticket, ctl[1] = CTAs that have finished; the last CTA out resets both, so every
// launch (they are stream-ordered) starts from zero without a memset in the graph.
template <typename Op, int kPT, int MINB>
__global__ void __launch_bounds__(kPT, MINB)
k_brick_persist(const __grid_constant__ Op op, const uint16_t* __restrict__ nbr, const uint32_t* __restrict__ nbr_count,
                const BrickRec* __restrict__ bricks, const GridDesc* __restrict__ desc, const StatusBlock* st,
                unsigned int* ctl, int K) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ PSlot slots[kPS];
  constexpr size_t kSlotBytes = (size_t)Op::kTiles * kTileBytes;
  constexpr int kPW = kPT / 32;
  pdl_wait();
  if (batch_failed(st)) return;
  const int nbricks = desc->nbricks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kPS) mbar_init(&slots[threadIdx.x].full, 1);
  __syncthreads();
  // one warp fills the ring IN ORDER: tickets of a CTA must grow with the slot sequence, so that the
  // first empty slot a warp meets means that every later one is empty too
  if (warp == 0) {
#pragma unroll 1
    for (int s = 0; s < kPS; ++s) persist_refill(op, slots[s], dyn + (size_t)s * kSlotBytes, bricks, nbricks, ctl, lane);
  }
  for (int it = 0;; ++it) {
    const int s = it % kPS;
    PSlot& sl = slots[s];
    mbar_wait(&sl.full, (uint32_t)(it / kPS) & 1u);
    if (sl.brick < 0) break;
    const unsigned char* ta = dyn + (size_t)s * kSlotBytes;
    const unsigned char* tb = ta + (Op::kTiles > 1 ? kTileBytes : 0);
    const int own_n = sl.rec.own_n;
    // chunks of 32 owned particles; the list row, count and first quad of the NEXT chunk are
    // requested before the current one is processed (the chain ticket -> count -> list -> gather)
    int base = 0;
    if (lane == 0) base = atomicAdd(&sl.next, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    bool have = base + lane < own_n;
    Owned o{};
    const uint2* row = nullptr;
    uint32_t cnt = 0;
    uint2 first = make_uint2(0u, 0u);
    if (have) {
      o = brick_owned(sl.rec, base + lane);
      row = list_row(nbr, K, o.i);
      cnt = nbr_count[o.i];
      first = __ldcs(row);
    }
    while (base < own_n) {
      int nbase = 0;
      if (lane == 0) nbase = atomicAdd(&sl.next, 32);
      nbase = __shfl_sync(0xffffffffu, nbase, 0);
      const bool haven = nbase + lane < own_n;
      Owned on{};
      const uint2* rown = nullptr;
      uint32_t cntn = 0;
      uint2 firstn = make_uint2(0u, 0u);
      if (haven) {
        on = brick_owned(sl.rec, nbase + lane);
        rown = list_row(nbr, K, on.i);
        cntn = nbr_count[on.i];
        firstn = __ldcs(rown);
      }
      if (have) op.particle(o, row, cnt, first, ta, tb);
      base = nbase;
      have = haven;
      o = on;
      row = rown;
      cnt = cntn;
      first = firstn;
    }
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence_block();
      last = atomicAdd(&sl.left, 1) == kPW - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence_block();
      persist_refill(op, sl, dyn + (size_t)s * kSlotBytes, bricks, nbricks, ctl, lane);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(ctl + 1, gridDim.x - 1) == gridDim.x - 1) {
      __threadfence();
      ctl[0] = 0u;
    }
  }
}

// <<<grid, block, smem, stream>>> with the optional programmatic-serialization attribute of PBF_LAUNCH
template <typename... KArgs, typename... Args>
inline void launch_smem(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PBF_PDL ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename Kern>
cudaError_t opt_in(Kern kernel, size_t smem) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

static inline bool common_case(const StepConsts& c) {
  return c.sqrt_safe != 0 && (c.scorr_on == 0 || c.scorr_n == 4);
}

}  // namespace

// ================================================================== launchers
namespace {

int g_sms = 0;
bool g_ring_fits = true;  // a ring of kPS slots fits one SM twice (one-tile passes run two CTAs per SM)

template <typename Op>
cudaError_t opt_in_op() {
  constexpr size_t once = (size_t)Op::kTiles * kTileBytes, ring = (size_t)kPS * once;
  cudaError_t e = opt_in(k_brick_once<Op>, once);
  if (e != cudaSuccess) return e;
  if (!g_ring_fits) return e;
  if constexpr (Op::kTiles == 1) return opt_in(k_brick_persist<Op, PBF_PERSIST_THREADS, 2>, ring);
  else return opt_in(k_brick_persist<Op, PBF_PERSIST_THREADS2, 1>, ring);
}

// Runs one pass over every brick with the driver selected at start-up.
template <typename Op>
void run_pass(const Op& op, const NeighborList& nl, const StatusBlock* st, cudaStream_t s) {
  const uint16_t* idx = reinterpret_cast<const uint16_t*>(nl.idx);
  if (nl.brick_persist && g_ring_fits) {
    constexpr size_t ring = (size_t)kPS * Op::kTiles * kTileBytes;
    if constexpr (Op::kTiles == 1)
      launch_smem(k_brick_persist<Op, PBF_PERSIST_THREADS, 2>, 2 * g_sms, PBF_PERSIST_THREADS, ring, s, op, idx, nl.count,
                  nl.bricks, nl.desc, st, nl.brick_ctl, nl.K);
    else
      launch_smem(k_brick_persist<Op, PBF_PERSIST_THREADS2, 1>, g_sms, PBF_PERSIST_THREADS2, ring, s, op, idx, nl.count,
                  nl.bricks, nl.desc, st, nl.brick_ctl, nl.K);
  } else {
    launch_smem(k_brick_once<Op>, nl.brick_cap, Op::kTiles == 1 ? kBT : kBT2, (size_t)Op::kTiles * kTileBytes, s, op, idx,
                nl.count, nl.bricks, nl.desc, st, nl.K);
  }
}

}  // namespace

int brick_setup() {
  if ((size_t)kPS * kTileBytes * 2 > 220u * 1024u) g_ring_fits = false;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chk(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  chk(opt_in(k_neighbors_brick, kTileBytes));
  chk(opt_in_op<LambdaOp<true, true>>());
  chk(opt_in_op<LambdaOp<true, false>>());
  chk(opt_in_op<LambdaOp<false, false>>());
  chk(opt_in_op<DeltaOp<true, false, true>>());
  chk(opt_in_op<DeltaOp<true, true, true>>());
  chk(opt_in_op<DeltaOp<true, false, false>>());
  chk(opt_in_op<DeltaOp<true, true, false>>());
  chk(opt_in_op<DeltaOp<false, false, false>>());
  chk(opt_in_op<DeltaOp<false, true, false>>());
  chk(opt_in_op<XsphOp<true>>());
  chk(opt_in_op<XsphOp<false>>());
  chk(opt_in_op<OmegaOp<true>>());
  chk(opt_in_op<OmegaOp<false>>());
  chk(opt_in_op<EtaOp<true>>());
  chk(opt_in_op<EtaOp<false>>());
  return e == cudaSuccess ? 0 : (int)e;
}

int launch_brick_table(const GridBuffers& g, cudaStream_t s) {
  PBF_LAUNCH(k_brick_table, g.brick_cap, 64, s, g.cell_range, g.desc, g.bricks, g.status);
  return 1;
}

int launch_neighbors_brick(const float4* pred_s, const StepConsts& c, const GridBuffers& g, const NeighborList& nl,
                           cudaStream_t s) {
  launch_smem(k_neighbors_brick, nl.brick_cap, kBT, kTileBytes, s, pred_s, g.cell_range, g.desc, nl.bricks,
              reinterpret_cast<uint16_t*>(nl.idx), nl.count, g.status, c.inv_h, c.h2, nl.K);
  return 1;
}

int launch_lambda_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool strict,
                        cudaStream_t s) {
  if (strict && c.sqrt_safe)
    run_pass(LambdaOp<true, true>{b.pred[cur], b.rho, c, b.dbg}, nl, b.status, s);
  else if (strict)
    run_pass(LambdaOp<true, false>{b.pred[cur], b.rho, c, b.dbg}, nl, b.status, s);
  else
    run_pass(LambdaOp<false, false>{b.pred[cur], b.rho, c, b.dbg}, nl, b.status, s);
  return 1;
}

template <bool S, bool COMMON>
static void delta_brick_impl(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool last,
                             bool is_final, cudaStream_t s) {
  if (last)
    run_pass(DeltaOp<S, true, COMMON>{b.pred[cur], b.pred[cur ^ 1], b.pos_s, b.rho, b.vel[0], b.planes, b.pos_o, b.vel_o, c,
                                      b.dbg, b.halo, is_final ? 1 : 0},
             nl, b.status, s);
  else
    run_pass(DeltaOp<S, false, COMMON>{b.pred[cur], b.pred[cur ^ 1], b.pos_s, b.rho, b.vel[0], b.planes, b.pos_o, b.vel_o, c,
                                       b.dbg, b.halo, 0},
             nl, b.status, s);
}

int launch_delta_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, int cur, bool last,
                       bool is_final, bool strict, cudaStream_t s) {
  if (strict && common_case(c)) delta_brick_impl<true, true>(b, nl, c, cur, last, is_final, s);
  else if (strict) delta_brick_impl<true, false>(b, nl, c, cur, last, is_final, s);
  else delta_brick_impl<false, false>(b, nl, c, cur, last, is_final, s);
  return 1;
}

int launch_xsph_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, bool is_final,
                      bool strict, cudaStream_t s) {
  if (strict)
    run_pass(XsphOp<true>{pos, b.vel[0], b.vel[1], b.pos_s, b.planes, b.pos_o, b.vel_o, c, b.dbg, b.halo, is_final ? 1 : 0},
             nl, b.status, s);
  else
    run_pass(XsphOp<false>{pos, b.vel[0], b.vel[1], b.pos_s, b.planes, b.pos_o, b.vel_o, c, b.dbg, b.halo, is_final ? 1 : 0},
             nl, b.status, s);
  return 1;
}

int launch_vort_omega_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                            bool strict, cudaStream_t s) {
  if (strict) run_pass(OmegaOp<true>{pos, b.vel[vcur], b.omega, c}, nl, b.status, s);
  else run_pass(OmegaOp<false>{pos, b.vel[vcur], b.omega, c}, nl, b.status, s);
  return 1;
}

int launch_vort_apply_brick(const SolveBuffers& b, const NeighborList& nl, const StepConsts& c, float4* pos, int vcur,
                            bool strict, cudaStream_t s) {
  if (strict)
    run_pass(EtaOp<true>{pos, b.vel[vcur], b.omega, b.pos_s, b.planes, b.pos_o, b.vel_o, c, b.dbg}, nl, b.status, s);
  else
    run_pass(EtaOp<false>{pos, b.vel[vcur], b.omega, b.pos_s, b.planes, b.pos_o, b.vel_o, c, b.dbg}, nl, b.status, s);
  return 1;
}

}  // namespace pbf

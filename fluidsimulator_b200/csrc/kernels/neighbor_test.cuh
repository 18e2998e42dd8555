// neighbor_test.cuh — the candidate test loop of the neighbour build (reference core.cpp:205-247),
// shared by the global-memory kernel (grid.cu: candidates through L1, 32-bit entries) and the brick
// kernel (brick.cu: candidates from the shared-memory tile, 16-bit tile-relative entries).
#pragma once

#include "pbf_device.cuh"

namespace pbf {

#ifndef PBF_NBR_MASK_UNROLL
#define PBF_NBR_MASK_UNROLL 1
#endif
// PBF_NBR_MASK=1 (default): the candidate test and the list store are separated.  In the loop above
// some lane of the warp has a hit in almost every step, so the whole warp walks through both store
// blocks (62 instructions per candidate pair, 40 of them bookkeeping and control flow; SASS of
// r01e).  Here the test loop only collects a bit per candidate (<= 32 candidates per chunk of a
// cell: 25 instructions per pair) and has no hit-dependent control flow; the hits are then emitted
// from the mask in ascending slot order (17 instructions per hit) — the order of the loop above.
// The second candidate of a step is loaded unconditionally: the slot after the last particle is
// padding (ensure_particles) and its bit is masked.
// MEASURED on B200 (fluid_million, settled, profiles/ab_r01g_*.txt): 254 -> 228 us, bit-identical
// state and lists; unrolling the test loop 2x / 4x (48 / 56 registers) gives the gain back.
constexpr int kNbrMaskUnroll = PBF_NBR_MASK_UNROLL;  // candidate PAIRS per unrolled step of the test loop

// `ps` is the candidate array (sorted slots in global memory, or the tile in shared memory), `range`
// the cell's [start, end) in it, `i` the index of the particle itself in the same array; Cursor::put(j)
// appends candidate j to the list.
template <bool CENTER, class Cursor>
__device__ __forceinline__ void neighbors_cell_mask(const float4* __restrict__ pred_s, int2 range, float pz, f2 pxy,
                                                    int i, float h2, Cursor& e) {
#pragma unroll 1
  for (int base = range.x; base < range.y; base += 32) {
    const int end = min(range.y, base + 32);
    uint32_t m = 0;
#pragma unroll kNbrMaskUnroll
    for (int j = base; j < end; j += 2) {
      const float4 a0 = pred_s[j];
      const float4 a1 = pred_s[j + 1];
      const f2 d0 = __fadd2_rn(pxy, make_float2(-a0.x, -a0.y));
      const f2 d1 = __fadd2_rn(pxy, make_float2(-a1.x, -a1.y));
      const float z0 = __fsub_rn(pz, a0.z), z1 = __fsub_rn(pz, a1.z);
      const f2 q0 = __fmul2_rn(d0, d0), q1 = __fmul2_rn(d1, d1);
      const float r2a = __fadd_rn(__fadd_rn(q0.x, q0.y), __fmul_rn(z0, z0));
      const float r2b = __fadd_rn(__fadd_rn(q1.x, q1.y), __fmul_rn(z1, z1));
      // hits are shifted in from the top, two per step (core.cpp:231-240: strict r2 < h2)
      m = (m >> 2) | ((r2a < h2) ? 0x40000000u : 0u) | ((r2b < h2) ? 0x80000000u : 0u);
    }
    const int lim = 32 - (end - base);            // 0 .. 31
    m >>= lim & ~1;                               // candidate base + t at bit t (an odd chunk ran one slot over)
    m &= 0xffffffffu >> lim;                      // ... whose bit is dropped here
    if (CENTER) {
      const uint32_t self = (uint32_t)(i - base);
      if (self < 32u) m &= ~(1u << self);
    }
#pragma unroll 1
    while (m) {
      const int t = __ffs((int)m) - 1;
      m &= m - 1u;
      e.put(base + t);
    }
  }
}


}  // namespace pbf

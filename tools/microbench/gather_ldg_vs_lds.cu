// gather_ldg_vs_lds.cu — micro-benchmark behind DESIGN.md §9: does staging a warp's stencil tile in
// shared memory and gathering neighbours with LDS.128 beat gathering them with LDG.128 through L1?
//
// Model of one solver pass (k_lambda / k_delta): a warp owns 32 consecutive sorted particles (5 cells
// of one z-column); its neighbours live in 9 contiguous runs of ~7 cells (the 3x3 columns around it).
// Step k of lane l reads the 16-byte record of its k-th neighbour: stencil cell sc = k*27/STEPS in the
// reference's (dz, dy, dx) order, run r = (dy, dx), position inside the run = the lane's cell + dz,
// plus a random member of that cell.  The particle array is shared (1 M records, 16 MB, L2-resident;
// run (dy, dx) of warp w starts at 32 w + 350 dy + 19250 dx, so neighbouring warps reuse each other's
// lines as in the real passes).  Variant A gathers from global memory (L1/L2), variant B copies the
// 9 runs into shared memory first (coalesced loads) and gathers with LDS.128, variant C lets the TMA
// engine copy them (cp.async.bulk, one copy per run).  `work` dependent
// FMAs per neighbour stand in for the ~30 instructions of the real passes.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_ldg_vs_lds gather_ldg_vs_lds.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int kSteps = 32;           // neighbours per particle
constexpr int kRun = 50;             // slots per run (7 cells x 6.4 + slack)
constexpr int kTile = 9 * kRun;      // 450 slots = 7.2 KB
constexpr int kTilePad = 456;        // multiple of 8 slots
constexpr int kBlock = 128;

constexpr int kPad = 20000, kN = 1000000;
__host__ __device__ inline int run_start(int w, int r) { return kPad + 32 * w + (r / 3 - 1) * 350 + (r % 3 - 1) * 19250 - 9; }

template <int WORK>
__device__ __forceinline__ float consume(float4 v, float acc) {
  float t = v.x * v.y + v.z;
#pragma unroll
  for (int w = 0; w < WORK; ++w) t = fmaf(t, 1.0001f, v.w);
  return acc + t;
}

template <int WORK>
__global__ void __launch_bounds__(kBlock) k_ldg(const float4* __restrict__ tiles, const uint32_t* __restrict__ idx,
                                               float* out, int nwarps) {
  const int gw = (blockIdx.x * kBlock + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= nwarps) return;
  const float4* tile = tiles;  // entries are global slots
  const uint2* row = reinterpret_cast<const uint2*>(idx + (size_t)gw * kSteps * 32) + lane;
  float acc = 0.f;
  for (int k = 0; k < kSteps / 2; k += 2) {
    const uint2 e0 = __ldcs(row + (size_t)k * 32), e1 = __ldcs(row + (size_t)(k + 1) * 32);
    const float4 a = tile[e0.x], b = tile[e0.y], c = tile[e1.x], d = tile[e1.y];
    acc = consume<WORK>(a, acc); acc = consume<WORK>(b, acc);
    acc = consume<WORK>(c, acc); acc = consume<WORK>(d, acc);
  }
  out[blockIdx.x * kBlock + threadIdx.x] = acc;
}

template <int WORK>
__global__ void __launch_bounds__(kBlock) k_lds(const float4* __restrict__ tiles, const uint32_t* __restrict__ idx,
                                               float* out, int nwarps) {
  __shared__ float4 sm[kBlock / 32][kTilePad];
  const int gw = (blockIdx.x * kBlock + threadIdx.x) >> 5, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (gw >= nwarps) return;
  for (int r = 0; r < 9; ++r) {
    const float4* src = tiles + run_start(gw, r);
    for (int t = lane; t < kRun; t += 32) sm[w][r * kRun + t] = src[t];
  }
  __syncwarp();
  const uint2* row = reinterpret_cast<const uint2*>(idx + (size_t)gw * kSteps * 32) + lane;
  float acc = 0.f;
  for (int k = 0; k < kSteps / 2; k += 2) {
    const uint2 e0 = __ldcs(row + (size_t)k * 32), e1 = __ldcs(row + (size_t)(k + 1) * 32);
    const float4 a = sm[w][e0.x], b = sm[w][e0.y], c = sm[w][e1.x], d = sm[w][e1.y];
    acc = consume<WORK>(a, acc); acc = consume<WORK>(b, acc);
    acc = consume<WORK>(c, acc); acc = consume<WORK>(d, acc);
  }
  out[blockIdx.x * kBlock + threadIdx.x] = acc;
}

// Variant C: the tile arrives as 9 bulk copies (cp.async.bulk, the TMA engine), one per run, issued by
// lanes 0..8 in one instruction and completed on a per-warp mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int WORK>
__global__ void __launch_bounds__(kBlock) k_tma(const float4* __restrict__ tiles, const uint32_t* __restrict__ idx,
                                               float* out, int nwarps) {
  __shared__ __align__(128) float4 sm[kBlock / 32][kTilePad];
  __shared__ __align__(8) unsigned long long bars[kBlock / 32];
  const int gw = (blockIdx.x * kBlock + threadIdx.x) >> 5, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (gw >= nwarps) return;
  const uint32_t bar = smem_u32(&bars[w]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(kTile * sizeof(float4))) : "memory");
  }
  __syncwarp();
  if (lane < 9) {
    const uint32_t dst = smem_u32(&sm[w][lane * kRun]);
    const float4* src = tiles + run_start(gw, lane);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"((uint32_t)(kRun * sizeof(float4))), "r"(bar) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar) : "memory");
  }
  const uint2* row = reinterpret_cast<const uint2*>(idx + (size_t)gw * kSteps * 32) + lane;
  float acc = 0.f;
  for (int k = 0; k < kSteps / 2; k += 2) {
    const uint2 e0 = __ldcs(row + (size_t)k * 32), e1 = __ldcs(row + (size_t)(k + 1) * 32);
    const float4 a = sm[w][e0.x], b = sm[w][e0.y], c = sm[w][e1.x], d = sm[w][e1.y];
    acc = consume<WORK>(a, acc); acc = consume<WORK>(b, acc);
    acc = consume<WORK>(c, acc); acc = consume<WORK>(d, acc);
  }
  out[blockIdx.x * kBlock + threadIdx.x] = acc;
}

template <typename K>
float time_it(K kernel, int blocks, const float4* tiles, const uint32_t* idx, float* out, int nwarps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) kernel<<<blocks, kBlock>>>(tiles, idx, out, nwarps);
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) kernel<<<blocks, kBlock>>>(tiles, idx, out, nwarps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 20.f * 1e3f;
}

int main() {
  const int nwarps = kN / 32;
  std::vector<float4> pred((size_t)kN + 2 * kPad + 64);
  for (auto& v : pred) v = make_float4(drand48(), drand48(), drand48(), 1e-3f);
  std::vector<uint32_t> idx_g((size_t)nwarps * kSteps * 32), idx_s(idx_g.size());
  for (int w = 0; w < nwarps; ++w)
    for (int k = 0; k < kSteps; ++k)
      for (int l = 0; l < 32; ++l) {
        int kk = k + (int)(drand48() * 3.0) - 1;          // lanes drift by a list position or two
        if (kk < 0) kk = 0;
        if (kk >= kSteps) kk = kSteps - 1;
        const int sc = kk * 27 / kSteps, dz = sc / 9, dy = (sc / 3) % 3, dx = sc % 3;
        const int cell = (int)(l / 6.4) + dz;             // 0..6 inside the run
        int rel = (int)(cell * 6.4 + drand48() * 6.4);
        if (rel >= kRun) rel = kRun - 1;
        const int r = dy * 3 + dx;
        // list layout of the real passes: entry k of lane l at (k/2)*64 + l*2 + k%2
        const size_t at = (size_t)w * kSteps * 32 + (k / 2) * 64 + l * 2 + (k & 1);
        idx_g[at] = (uint32_t)(run_start(w, r) + rel);
        idx_s[at] = (uint32_t)(r * kRun + rel);
      }
  float4* d_pred; uint32_t *d_g, *d_s; float* d_out;
  cudaMalloc(&d_pred, pred.size() * sizeof(float4));
  cudaMalloc(&d_g, idx_g.size() * sizeof(uint32_t));
  cudaMalloc(&d_s, idx_s.size() * sizeof(uint32_t));
  cudaMalloc(&d_out, (size_t)nwarps * 32 * sizeof(float));
  cudaMemcpy(d_pred, pred.data(), pred.size() * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemcpy(d_g, idx_g.data(), idx_g.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
  cudaMemcpy(d_s, idx_s.data(), idx_s.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
  const int blocks = (nwarps * 32 + kBlock - 1) / kBlock;
  std::printf("gathers per launch: %d x %d x 32; shared-memory tile %d B per warp\n", nwarps, kSteps, (int)(kTilePad * sizeof(float4)));
  std::printf("work  ldg_us  lds_us  tma_us\n");
#define ROW(W) std::printf("%4d %7.1f %7.1f %7.1f\n", W, time_it(k_ldg<W>, blocks, d_pred, d_g, d_out, nwarps), \
                           time_it(k_lds<W>, blocks, d_pred, d_s, d_out, nwarps), time_it(k_tma<W>, blocks, d_pred, d_s, d_out, nwarps));
  ROW(0) ROW(8) ROW(16) ROW(30)
  cudaError_t e = cudaDeviceSynchronize();
  std::printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

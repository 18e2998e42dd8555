// gather_ldg_vs_lds.cu — micro-benchmark behind DESIGN.md §9: does staging a warp's stencil tile in
// shared memory and gathering neighbours with LDS.128 beat gathering them with LDG.128 through L1?
//
// Model of one solver pass (k_lambda / k_delta): a warp owns 32 consecutive sorted particles (5 cells
// of one z-column); its neighbours live in 9 contiguous runs of ~7 cells (the 3x3 columns around it).
// Step k of lane l reads the 16-byte record of its k-th neighbour: stencil cell sc = k*27/STEPS in the
// reference's (dz, dy, dx) order, run r = (dy, dx), position inside the run = the lane's cell + dz,
// plus a random member of that cell.  Variant A gathers from the global tile (L1/L2), variant B
// copies the tile into shared memory first (coalesced) and gathers with LDS.128.  `work` dependent
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
  const float4* tile = tiles + (size_t)gw * kTilePad;
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
  const float4* tile = tiles + (size_t)gw * kTilePad;
  for (int t = lane; t < kTilePad; t += 32) sm[w][t] = tile[t];
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
  const int nwarps = 31250;  // 1 M particles
  std::vector<float4> tiles((size_t)nwarps * kTilePad);
  for (auto& v : tiles) v = make_float4(drand48(), drand48(), drand48(), 1e-3f);
  std::vector<uint32_t> idx((size_t)nwarps * kSteps * 32);
  for (int w = 0; w < nwarps; ++w)
    for (int k = 0; k < kSteps; ++k)
      for (int l = 0; l < 32; ++l) {
        int kk = k + (int)(drand48() * 3.0) - 1;          // lanes drift by a list position or two
        if (kk < 0) kk = 0;
        if (kk >= kSteps) kk = kSteps - 1;
        const int sc = kk * 27 / kSteps, dz = sc / 9, dy = (sc / 3) % 3, dx = sc % 3;
        const int cell = (int)(l / 6.4) + dz;             // 0..6 inside the run
        int slot = (dy * 3 + dx) * kRun + (int)(cell * 6.4 + drand48() * 6.4);
        if (slot >= kTile) slot = kTile - 1;
        // list layout of the real passes: entry k of lane l at (k/2)*64 + l*2 + k%2
        idx[(size_t)w * kSteps * 32 + (k / 2) * 64 + l * 2 + (k & 1)] = (uint32_t)slot;
      }
  float4* d_tiles; uint32_t* d_idx; float* d_out;
  cudaMalloc(&d_tiles, tiles.size() * sizeof(float4));
  cudaMalloc(&d_idx, idx.size() * sizeof(uint32_t));
  cudaMalloc(&d_out, (size_t)nwarps * 32 * sizeof(float));
  cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemcpy(d_idx, idx.data(), idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
  const int blocks = (nwarps * 32 + kBlock - 1) / kBlock;
  std::printf("gathers per launch: %d x %d x 32; tile %d B per warp\n", nwarps, kSteps, (int)(kTilePad * sizeof(float4)));
  std::printf("work  ldg_us  lds_us\n");
  std::printf("%4d %7.1f %7.1f\n", 0, time_it(k_ldg<0>, blocks, d_tiles, d_idx, d_out, nwarps), time_it(k_lds<0>, blocks, d_tiles, d_idx, d_out, nwarps));
  std::printf("%4d %7.1f %7.1f\n", 8, time_it(k_ldg<8>, blocks, d_tiles, d_idx, d_out, nwarps), time_it(k_lds<8>, blocks, d_tiles, d_idx, d_out, nwarps));
  std::printf("%4d %7.1f %7.1f\n", 16, time_it(k_ldg<16>, blocks, d_tiles, d_idx, d_out, nwarps), time_it(k_lds<16>, blocks, d_tiles, d_idx, d_out, nwarps));
  std::printf("%4d %7.1f %7.1f\n", 30, time_it(k_ldg<30>, blocks, d_tiles, d_idx, d_out, nwarps), time_it(k_lds<30>, blocks, d_tiles, d_idx, d_out, nwarps));
  cudaError_t e = cudaDeviceSynchronize();
  std::printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

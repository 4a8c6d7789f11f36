// Microbenchmark (sm_100a): throughput of the TMA reduce path, cp.reduce.async.bulk ... .add.f32 (shared -> global, the adds
// done in L2), at the tile size a fused attention backward would need for its dQ partials: one 64 x 64 fp32 tile (16 KB) per
// (128-key, 64-query) tile pair and SM.  Patterns:
//   0  every CTA adds into its own tiles (no two CTAs ever touch the same address)
//   1  the attention pattern: the 15 key-tile CTAs of a (batch, head) walk the same 29 query tiles at the same time
//   2  plain cp.async.bulk stores of the same tiles (the no-add reference)
// One elected thread issues; DEPTH bulk groups in flight per CTA.  Prints microseconds, GB/s and cycles per tile and SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_reduce tools/ubench/tma_reduce.cu && tools/ubench/tma_reduce
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TILE_BYTES = 64 * 64 * 4;

template <int PATTERN, int DEPTH>
__global__ void k(float* g, int tiles_per_cta, int nq, int ctas_per_bh, size_t total_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* s = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < DEPTH * TILE_BYTES / 4; i += blockDim.x) s[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t sa = uint32_t(__cvta_generic_to_shared(smem));
    for (int it = 0; it < tiles_per_cta; ++it) {
      size_t tile;
      if (PATTERN == 1) tile = size_t(blockIdx.x / ctas_per_bh) * nq + size_t(it % nq);
      else tile = (size_t(blockIdx.x) * tiles_per_cta + it);
      tile %= total_tiles;
      const float* dst = g + tile * (TILE_BYTES / 4);
      const uint32_t src = sa + uint32_t(it % DEPTH) * TILE_BYTES;
      if (PATTERN == 2)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(TILE_BYTES) : "memory");
      else
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(TILE_BYTES)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");  // the slot reused next is free again
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <int PATTERN, int DEPTH>
void run(const char* name, float* g, size_t total_tiles, int sms, int clock_khz) {
  const int tiles_per_cta = 29 * 8;  // eight (batch, head) rounds of 29 query tiles
  cudaFuncSetAttribute(k<PATTERN, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * TILE_BYTES);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(g, 0, total_tiles * TILE_BYTES);
    cudaEventRecord(e0);
    k<PATTERN, DEPTH><<<sms, 128, DEPTH * TILE_BYTES>>>(g, tiles_per_cta, 29, 15, total_tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 1) {
      const double bytes = double(sms) * tiles_per_cta * TILE_BYTES;
      printf("%-44s depth %d: %8.1f us  %7.1f GB/s  %7.0f cycles per tile and SM (at %d MHz)  %s\n", name, DEPTH, ms * 1e3,
             bytes / ms / 1e6, ms * 1e-3 * clock_khz * 1e3 / tiles_per_cta, clock_khz / 1000,
             cudaGetLastError() == cudaSuccess ? "" : cudaGetErrorString(cudaGetLastError()));
    }
  }
  float h[4];
  cudaMemcpy(h, g, sizeof(h), cudaMemcpyDeviceToHost);
  printf("    first element after the run: %.0f\n", h[0]);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount;
  const size_t total_tiles = size_t(sms) * 29 * 8;  // 550 MB: larger than L2
  float* g;
  cudaMalloc(&g, total_tiles * TILE_BYTES);
  printf("%s, %d SMs, 16 KB fp32 tiles, %zu tiles (%.0f MB)\n", prop.name, sms, total_tiles, total_tiles * TILE_BYTES / 1e6);
  run<2, 2>("bulk store (no add), own tiles", g, total_tiles, sms, clock_khz);
  run<0, 1>("bulk reduce add.f32, own tiles", g, total_tiles, sms, clock_khz);
  run<0, 2>("bulk reduce add.f32, own tiles", g, total_tiles, sms, clock_khz);
  run<0, 4>("bulk reduce add.f32, own tiles", g, total_tiles, sms, clock_khz);
  run<1, 2>("bulk reduce add.f32, 15 CTAs share 29 tiles", g, total_tiles, sms, clock_khz);
  run<1, 4>("bulk reduce add.f32, 15 CTAs share 29 tiles", g, total_tiles, sms, clock_khz);
  return 0;
}

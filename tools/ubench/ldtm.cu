// Microbenchmark: tcgen05.ld (LDTM) throughput per SM vs number of warps and shape (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../repurpose_b200/csrc/ptx.cuh"
using namespace rp;
template <int X>  // columns per load: 32 or 16
__global__ void k(long long* cyc, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[32];
    if (X == 32) { tmem_ld32(tb + (it & 7) * 32, v); }
    else { tmem_ld16(tb + (it & 15) * 16, v); }
    tmem_ld_wait();
    #pragma unroll
    for (int i = 0; i < X; i += 8) acc ^= v[i];
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}
template <int X>
__global__ void k4(long long* cyc, uint32_t* sink, int iters) {  // 4 loads in flight per wait
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[128];
    tmem_ld32(tb + 0, v); tmem_ld32(tb + 32, v + 32); tmem_ld32(tb + 64, v + 64); tmem_ld32(tb + 96, v + 96);
    tmem_ld_wait();
    #pragma unroll
    for (int i = 0; i < 128; i += 16) acc ^= v[i];
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}
int main() {
  long long* cyc; uint32_t* sink; cudaMalloc(&cyc, 64); cudaMalloc(&sink, 4096);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    k<32><<<1, 32 * warps>>>(cyc, sink, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double bytes = double(iters) * warps * 32 * 32 * 4;
    printf("x32 serial  warps=%2d: %.1f cycles per LDTM per warp, %.1f B/clk/SM\n", warps, double(c) / iters, bytes / c);
    k4<32><<<1, 32 * warps>>>(cyc, sink, iters); cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    bytes = double(iters) * warps * 128 * 32 * 4;
    printf("x32 4-deep  warps=%2d: %.1f cycles per 4 LDTM per warp, %.1f B/clk/SM\n", warps, double(c) / iters, bytes / c);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

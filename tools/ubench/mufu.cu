// Microbenchmark: MUFU.EX2 issue cadence per warp vs. warps per SM sub-partition (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template<int ILP>
__global__ void k(float* out, long long* cyc, int iters){
  float v[ILP];
  #pragma unroll
  for(int i=0;i<ILP;++i) v[i]=threadIdx.x*1e-3f+i;
  __syncthreads();
  long long t0=clock64();
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<ILP;++i) v[i]=ex2(v[i]*0.0001f);   // FMUL + MUFU, independent chains
  }
  long long t1=clock64();
  float s=0; 
  #pragma unroll
  for(int i=0;i<ILP;++i) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
int main(){
  float* out; long long* cyc; cudaMalloc(&out,1<<20); cudaMalloc(&cyc,1024);
  for(int warps=1; warps<=16; warps*=2){
    const int iters=200;
    k<16><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    k<16><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
    double per_warp = double(c)/(iters*16);
    int per_smsp = (warps+3)/4;
    printf("warps/CTA=%2d (per SMSP %d): %.2f cycles per MUFU per warp -> %.2f cycles per MUFU per SMSP\n", warps, per_smsp, per_warp, per_warp/per_smsp);
  }
  return 0;
}

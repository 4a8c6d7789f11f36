// Microbenchmark: which ops share the XU pipe with MUFU.EX2 and at what rate (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ uint32_t ex2b(uint32_t x){uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;":"=r"(y):"r"(x)); return y;}
__device__ __forceinline__ uint32_t cvtb(float lo,float hi){uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;":"=r"(r):"f"(hi),"f"(lo)); return r;}
template<int MODE>
__global__ void k(uint32_t* out,long long* cyc,int iters){
  uint32_t v[16]; float f[16];
  #pragma unroll
  for(int i=0;i<16;++i){ v[i]=0x3c003c00u+threadIdx.x+i; f[i]=threadIdx.x*1e-3f+i; }
  __syncthreads();
  long long t0=clock64();
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<16;++i){
      if(MODE==0) f[i]=ex2(f[i]);                       // MUFU.EX2 fp32
      if(MODE==1) v[i]=ex2b(v[i]);                      // MUFU.EX2.BF16x2
      if(MODE==2) v[i]=cvtb(f[i], __uint_as_float(v[i]));   // F2FP only
      if(MODE==3){ f[i]=ex2(f[i]); v[i]=cvtb(f[i], f[(i+1)&15]); }  // MUFU + F2FP
      if(MODE==4){ v[i]=__byte_perm(v[i], v[(i+1)&15], 0x7632)+1; } // PRMT baseline
    }
  }
  long long t1=clock64();
  uint32_t s=0;
  #pragma unroll
  for(int i=0;i<16;++i) s^=v[i]^__float_as_uint(f[i]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0) cyc[0]=t1-t0;
}
template<int MODE> void run(const char* name,uint32_t*out,long long*cyc){
  for(int warps: {4,16}){
    const int iters=500;
    k<MODE><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    k<MODE><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
    printf("%-28s warps/SMSP=%d: %.2f cycles per loop-op per SMSP\n",name,warps/4,double(c)/(iters*16)/(warps/4.0));
  }
}
int main(){ uint32_t*out; long long*cyc; cudaMalloc(&out,1<<20); cudaMalloc(&cyc,64);
  run<0>("MUFU.EX2 f32",out,cyc); run<1>("MUFU.EX2.BF16x2",out,cyc); run<2>("F2FP bf16x2",out,cyc); run<3>("MUFU f32 + F2FP",out,cyc); run<4>("PRMT+IADD",out,cyc);
  return 0; }

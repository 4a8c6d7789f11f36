// Microbenchmark (sm_100a): does MUFU.EX2 share an issue / dispatch port with other pipes?  Each loop
// iteration issues 16 independent MUFU.EX2 plus NOTHER independent ops of one class; if the classes
// overlap, cycles/iteration stays at the MUFU cost (16 x 8), if they share a port it is the sum.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float fmx(float a,float b){float y; asm volatile("max.f32 %0, %1, %2;":"=f"(y):"f"(a),"f"(b)); return y;}
__device__ __forceinline__ float fmx3(float a,float b,float c){float y; asm volatile("max.f32 %0, %1, %2, %3;":"=f"(y):"f"(a),"f"(b),"f"(c)); return y;}
__device__ __forceinline__ float ffma(float a,float b,float c){float y; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(y):"f"(a),"f"(b),"f"(c)); return y;}
__device__ __forceinline__ uint32_t cvtb(float lo,float hi){uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;":"=r"(r):"f"(hi),"f"(lo)); return r;}
__device__ __forceinline__ uint32_t lop(uint32_t a,uint32_t b){uint32_t r; asm volatile("xor.b32 %0, %1, %2;":"=r"(r):"r"(a),"r"(b)); return r;}
__device__ __forceinline__ uint32_t iadd(uint32_t a,uint32_t b){uint32_t r; asm volatile("add.u32 %0, %1, %2;":"=r"(r):"r"(a),"r"(b)); return r;}
__device__ __forceinline__ uint32_t imad(uint32_t a,uint32_t b,uint32_t c){uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;":"=r"(r):"r"(a),"r"(b),"r"(c)); return r;}
__device__ __forceinline__ unsigned long long add2(unsigned long long a,unsigned long long b){unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a,unsigned long long b,unsigned long long c){unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ uint32_t shl(uint32_t a){uint32_t r; asm volatile("shl.b32 %0, %1, 23;":"=r"(r):"r"(a)); return r;}
__device__ __forceinline__ uint32_t hmax2(uint32_t a,uint32_t b){uint32_t r; asm volatile("max.bf16x2 %0, %1, %2;":"=r"(r):"r"(a),"r"(b)); return r;}
__device__ __forceinline__ uint32_t hfma2(uint32_t a,uint32_t b,uint32_t c){uint32_t r; asm volatile("fma.rn.bf16x2 %0, %1, %2, %3;":"=r"(r):"r"(a),"r"(b),"r"(c)); return r;}

// MODE: class of the other op; NMUFU, NOTHER per iteration
template<int MODE,int NMUFU,int NOTHER>
__global__ void k(uint32_t* out,long long* cyc,int iters){
  float f[16]; float g[16]; uint32_t v[16]; unsigned long long w[8];
  #pragma unroll
  for(int i=0;i<16;++i){ f[i]=threadIdx.x*1e-3f+i; g[i]=threadIdx.x*2e-3f-i; v[i]=threadIdx.x*77u+i; }
  #pragma unroll
  for(int i=0;i<8;++i) w[i]=(unsigned long long)(threadIdx.x+i)*0x3f8000013f800001ull;
  __syncthreads();
  long long t0=clock64();
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<16;++i){
      if(i<NMUFU) f[i]=ex2(f[i]);
      if(i<NOTHER){
        if(MODE==1) g[i]=fmx(g[i],g[(i+5)&15]);
        if(MODE==2) g[i]=ffma(g[i],1.0001f,g[(i+5)&15]);
        if(MODE==3) v[i]=cvtb(g[i],g[(i+5)&15])^v[i];
        if(MODE==4) v[i]=lop(v[i],v[(i+5)&15]);
        if(MODE==5) w[i&7]=add2(w[i&7],w[(i+3)&7]);
        if(MODE==6) g[i]=fmx3(g[i],g[(i+5)&15],g[(i+9)&15]);
        if(MODE==7) v[i]=iadd(v[i],v[(i+5)&15]);
        if(MODE==8) v[i]=imad(v[i],v[(i+5)&15],v[(i+9)&15]);
        if(MODE==9) w[i&7]=fma2(w[i&7],w[(i+3)&7],w[(i+5)&7]);
        if(MODE==10) v[i]=shl(v[i])+v[(i+5)&15];
        if(MODE==11) v[i]=hmax2(v[i],v[(i+5)&15]);
        if(MODE==12) v[i]=hfma2(v[i],v[(i+5)&15],v[(i+9)&15]);
      }
    }
  }
  long long t1=clock64();
  uint32_t s=0;
  #pragma unroll
  for(int i=0;i<16;++i) s^=v[i]^__float_as_uint(f[i])^__float_as_uint(g[i]);
  #pragma unroll
  for(int i=0;i<8;++i) s^=uint32_t(w[i])^uint32_t(w[i]>>32);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0) cyc[0]=t1-t0;
}
template<int MODE,int NM,int NO> void run(const char* name,uint32_t*out,long long*cyc){
  printf("%-34s", name);
  for(int warps: {4,8,12}){
    const int iters=400;
    k<MODE,NM,NO><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    k<MODE,NM,NO><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
    printf("  w/SMSP=%d: %7.1f cyc/iter/warp", warps/4, double(c)/iters/(warps/4.0));
  }
  printf("\n");
}
#define BOTH(M,name) run<M,0,16>(name " alone (16)",out,cyc); run<M,16,16>("16 MUFU + 16 " name,out,cyc);
int main(){ uint32_t*out; long long*cyc; cudaMalloc(&out,1<<20); cudaMalloc(&cyc,64);
  printf("cycles per loop iteration, divided by warps per SMSP (16 MUFU alone = 128 if the pipe is saturated)\n");
  run<0,16,0>("16 MUFU.EX2 alone",out,cyc);
  BOTH(1,"FMNMX") BOTH(6,"FMNMX3") BOTH(2,"FFMA") BOTH(3,"F2FP+LOP") BOTH(4,"LOP3") BOTH(7,"IADD") BOTH(8,"IMAD")
  BOTH(5,"FADD2") BOTH(9,"FFMA2") BOTH(10,"SHL+IADD") BOTH(11,"HMNMX2.BF16") BOTH(12,"HFMA2.BF16")
  return 0; }

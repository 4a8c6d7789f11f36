// Microbenchmark: cost of the FMA-pipe exp2 emulation (per pair) vs MUFU, and a 1:3 / 2:2 mix.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi){unsigned long long r; asm("mov.b64 %0, {%1, %2};":"=l"(r):"f"(lo),"f"(hi)); return r;}
__device__ __forceinline__ void unpack2(unsigned long long v,float&lo,float&hi){asm("mov.b64 {%0, %1}, %2;":"=f"(lo),"=f"(hi):"l"(v));}
__device__ __forceinline__ unsigned long long add2(unsigned long long a,unsigned long long b){unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a,unsigned long long b,unsigned long long c){unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ void emu2(unsigned long long x2,float&p0,float&p1){
  float x0,x1; unpack2(x2,x0,x1); x0=fmaxf(x0,-126.f); x1=fmaxf(x1,-126.f);
  unsigned long long xc=pack2(x0,x1), magic=pack2(12582912.f,12582912.f), nm=pack2(-12582912.f,-12582912.f);
  unsigned long long t=add2(xc,magic), n=add2(t,nm), f=fma2(n,pack2(-1.f,-1.f),xc);
  unsigned long long p=fma2(pack2(0.0551716536f,0.0551716536f),f,pack2(0.2426111251f,0.2426111251f));
  p=fma2(p,f,pack2(0.6932609677f,0.6932609677f)); p=fma2(p,f,pack2(0.9999280572f,0.9999280572f));
  float t0,t1,q0,q1; unpack2(t,t0,t1); unpack2(p,q0,q1);
  p0=__uint_as_float((__float_as_uint(t0)<<23)+__float_as_uint(q0)); p1=__uint_as_float((__float_as_uint(t1)<<23)+__float_as_uint(q1));
}
template<int EMU>  // EMU of every 4 pairs emulated
__global__ void k(float* out,long long* cyc,int iters){
  float v[32];
  #pragma unroll
  for(int i=0;i<32;++i) v[i]=-(threadIdx.x*1e-3f+i*0.1f);
  unsigned long long negm=pack2(-0.25f,-0.25f);
  __syncthreads();
  long long t0=clock64();
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int c=0;c<16;++c){
      unsigned long long x2=add2(pack2(v[2*c],v[2*c+1]),negm);
      float p0,p1;
      if((c&3)<EMU) emu2(x2,p0,p1); else { float x0,x1; unpack2(x2,x0,x1); p0=ex2(x0); p1=ex2(x1);} 
      v[2*c]=-p0; v[2*c+1]=-p1;
    }
  }
  long long t1=clock64();
  float s=0;
  #pragma unroll
  for(int i=0;i<32;++i) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int EMU> void run(float*out,long long*cyc){
  for(int warps=4;warps<=16;warps*=2){
    const int iters=200;
    k<EMU><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    k<EMU><<<1,32*warps>>>(out,cyc,iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
    printf("EMU=%d/4 warps/SMSP=%d: %.1f cycles per pair per SMSP (per element %.2f)\n",EMU,warps/4,double(c)/(iters*16)/(warps/4.0),double(c)/(iters*32)/(warps/4.0));
  }
}
int main(){ float*out; long long*cyc; cudaMalloc(&out,1<<20); cudaMalloc(&cyc,1024);
  run<0>(out,cyc); run<1>(out,cyc); run<2>(out,cyc); run<3>(out,cyc); run<4>(out,cyc); return 0; }

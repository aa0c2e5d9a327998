// Microbenchmark: issue/throughput of packed fma.rn.f32x2 vs scalar FFMA on sm_100a, alone and mixed with ALU/SHFL work.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template<int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s){
  float a[8]; u64 p[8]; unsigned q[8];
  #pragma unroll
  for(int i=0;i<8;i++){ a[i]=threadIdx.x*0.001f+i; float2 t=make_float2(a[i],a[i]+1.f); p[i]=*reinterpret_cast<u64*>(&t); q[i]=threadIdx.x+i; }
  float2 m2=make_float2(s,s); u64 m=*reinterpret_cast<u64*>(&m2);
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<8;i++){
      if(MODE==0) a[i]=fma1(a[i],s,a[i]);                  // 8 FFMA
      if(MODE==1) p[i]=fma2(p[i],m,p[i]);                  // 8 FFMA2 (16 fma)
      if(MODE==2){ a[i]=fma1(a[i],s,a[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i+1)&7]), "r"(it)); }  // 8 FFMA + 8 LOP3
      if(MODE==3){ p[i]=fma2(p[i],m,p[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i+1)&7]), "r"(it)); }  // 8 FFMA2 + 8 LOP3
      if(MODE==4){ p[i]=fma2(p[i],m,p[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i+1)&7]), "r"(it));
                   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[(i+3)&7]) : "r"(q[(i+2)&7]), "r"(it)); }                  // 8 FFMA2 + 16 LOP3
      if(MODE==5){ p[i]=fma2(p[i],m,p[i]); q[i]=__shfl_xor_sync(0xffffffffu,q[i],1); }                                                // 8 FFMA2 + 8 SHFL
      if(MODE==6){ a[i]=fma1(a[i],s,a[i]); q[i]=__shfl_xor_sync(0xffffffffu,q[i],1); }                                                // 8 FFMA + 8 SHFL
    }
  }
  float r=0; unsigned qq=0;
  #pragma unroll
  for(int i=0;i<8;i++){ float2 t=*reinterpret_cast<float2*>(&p[i]); r+=a[i]+t.x+t.y; qq^=q[i]; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=r+qq;
}
template<int MODE> void run(const char* name, int fma_per_iter, int other_per_iter){
  float* out; cudaMalloc(&out, 148*8*256*4);
  int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148*8,256>>>(out,100,1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148*8,256>>>(out,iters,1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double thr=148.0*8*256; double fma=thr*iters*fma_per_iter; double inst=thr/32*iters*(8+other_per_iter);
  printf("%-24s %.3f ms  %.1f TFLOP/s fp32  %.2f warp-inst/ns total (%.2f per SM per clk @1.965GHz)\n", name, ms, 2*fma/ms*1e-9, inst/ms*1e-6, inst/ms*1e-6/148/1.965);
  cudaFree(out);
}
int main(){
  run<0>("8 FFMA",8,0); run<1>("8 FFMA2",16,0); run<2>("8 FFMA + 8 LOP3",8,8); run<3>("8 FFMA2 + 8 LOP3",16,8);
  run<4>("8 FFMA2 + 16 LOP3",16,16); run<5>("8 FFMA2 + 8 SHFL",16,8); run<6>("8 FFMA + 8 SHFL",8,8);
  return 0;
}

// gather_probe.cu -- microbenchmark: random 8-byte (or 4-byte) gathers over a footprint of F MiB.
// Answers two design questions for the x gathers of SpMV on B200: (1) the gather rate when x is
// L2-resident (the ceiling of any CSR-like kernel on uniform-random matrices) and (2) the footprint at
// which the L2 stops holding x.  Development tool, not part of the library.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long z){z+=0x9E3779B97F4A7C15ull;z=(z^(z>>30))*0xBF58476D1CE4E5B9ull;z=(z^(z>>27))*0x94D049BB133111EBull;return z^(z>>31);}
template<typename T,int MODE> // MODE 0: plain ld.nc ; 1: evict_last hint
__global__ void gather(const T* __restrict__ x, unsigned long long n, int iters, T* out){
  unsigned long long t=blockIdx.x*(unsigned long long)blockDim.x+threadIdx.x; T s=0;
  uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;":"=l"(pol));
  unsigned long long h=mix(t);
  #pragma unroll 8
  for(int i=0;i<iters;++i){ h=h*6364136223846793005ull+1442695040888963407ull; unsigned long long idx=(h>>20)%n;
    T v; if(MODE==0) v=__ldg(x+idx); else { if(sizeof(T)==8) asm volatile("ld.global.nc.L2::cache_hint.f64 %0,[%1],%2;":"=d"(*(double*)&v):"l"(x+idx),"l"(pol)); else asm volatile("ld.global.nc.L2::cache_hint.f32 %0,[%1],%2;":"=f"(*(float*)&v):"l"(x+idx),"l"(pol)); }
    s+=v; }
  if(s==(T)123.456) out[0]=s;
}
// SpMV-like: streams col (4B) + val (8B) coalesced with evict-first, gathers x[col]
__global__ void spmv_like(const int* __restrict__ col,const double* __restrict__ val,const double* __restrict__ x,long long nnz,double* out){
  long long i=blockIdx.x*(long long)blockDim.x+threadIdx.x; long long stride=(long long)gridDim.x*blockDim.x; double s=0;
  for(;i<nnz;i+=stride){ int c; double v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0,[%1];":"=r"(c):"l"(col+i)); asm volatile("ld.global.nc.L1::no_allocate.f64 %0,[%1];":"=d"(v):"l"(val+i)); s+=v*__ldg(x+c);} if(s==123.456) out[0]=s;}
__global__ void fillcol(int* col,long long nnz,unsigned n){long long i=blockIdx.x*(long long)blockDim.x+threadIdx.x; if(i<nnz) col[i]=(int)((mix(i)>>11)%n);}
int main(){
  const size_t maxb=512ull<<20; double* x; cudaMalloc(&x,maxb); cudaMemset(x,0,maxb); double* out; cudaMalloc(&out,64);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int mibs[]={4,8,16,24,32,40,48,56,64,80,96,112,128,192,256,512};
  printf("pure gathers, 148*16 CTAs x 256 thr, 256 gathers/thread\n%8s %14s %14s %14s %14s\n","MiB","f64 Gg/s","f64+last Gg/s","f32 Gg/s","f32+last Gg/s");
  for(int mb:mibs){ float ms[4]; for(int v=0;v<4;++v){ unsigned long long n=((size_t)mb<<20)/((v<2)?8:4); int grid=148*16, it=256;
      for(int rep=0;rep<2;++rep){ cudaEventRecord(e0);
        if(v==0) gather<double,0><<<grid,256>>>(x,n,it,out); else if(v==1) gather<double,1><<<grid,256>>>(x,n,it,out);
        else if(v==2) gather<float,0><<<grid,256>>>((float*)x,n,it,(float*)out); else gather<float,1><<<grid,256>>>((float*)x,n,it,(float*)out);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[v],e0,e1);} }
    double g=148.0*16*256*256/1e6; printf("%8d %14.1f %14.1f %14.1f %14.1f\n",mb,g/ms[0],g/ms[1],g/ms[2],g/ms[3]); }
  // spmv-like with streams
  long long nnz=1ll<<28; int* col; double* val; cudaMalloc(&col,nnz*4); cudaMalloc(&val,nnz*8); cudaMemset(val,0,nnz*8);
  printf("spmv-like stream(12B/nnz)+gather, nnz=2^28\n%8s %12s %12s\n","x MiB","ms","Gnnz/s");
  for(int mb:mibs){ unsigned n=((size_t)mb<<20)/8; fillcol<<<(unsigned)((nnz+255)/256),256>>>(col,nnz,n); float ms=0;
    for(int rep=0;rep<2;++rep){cudaEventRecord(e0); spmv_like<<<148*32,256>>>(col,val,x,nnz,out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);} 
    printf("%8d %12.3f %12.1f\n",mb,ms,nnz/ms/1e6);}
  printf("err=%s\n",cudaGetErrorString(cudaGetLastError())); return 0; }

// gather_probe2.cu -- follow-up microbenchmarks to gather_probe.cu (development tool, not part of the library):
//  (a) random 8-byte gathers over SMALL footprints (16 KiB .. 4 MiB): is the ~267 Gg/s ceiling an L1TEX
//      tag-stage limit (then L1-resident footprints do not help) or an L2 limit?
//  (b) the same gathers through the texture path (tex1Dfetch<int2>): does TEX serve more distinct lines/clk?
//  (c) random 8-byte reads from shared memory (the only on-chip alternative);
//  (d) flushed device copy at the size of C1 (84 MB in+out): what a tiny problem can reach at all.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long z){z+=0x9E3779B97F4A7C15ull;z=(z^(z>>30))*0xBF58476D1CE4E5B9ull;z=(z^(z>>27))*0x94D049BB133111EBull;return z^(z>>31);}

__global__ void gather_mask(const double* __restrict__ x, unsigned mask, int iters, double* out){
  unsigned t=blockIdx.x*blockDim.x+threadIdx.x; double s=0; unsigned h=(unsigned)mix(t);
  #pragma unroll 8
  for(int i=0;i<iters;++i){ h=h*1664525u+1013904223u; unsigned idx=(h>>7)&mask; s+=__ldg(x+idx); }
  if(s==123.456) out[0]=s;
}
__global__ void gather_tex(cudaTextureObject_t tex, unsigned mask, int iters, double* out){
  unsigned t=blockIdx.x*blockDim.x+threadIdx.x; double s=0; unsigned h=(unsigned)mix(t);
  #pragma unroll 8
  for(int i=0;i<iters;++i){ h=h*1664525u+1013904223u; unsigned idx=(h>>7)&mask; int2 v=tex1Dfetch<int2>(tex,(int)idx); s+=__hiloint2double(v.y,v.x); }
  if(s==123.456) out[0]=s;
}
// pairs of lanes read the same 16-byte-aligned pair (2 gathers per sector touch): coalescing inside one instruction
__global__ void gather_pairs(const double* __restrict__ x, unsigned mask, int iters, double* out){
  unsigned t=blockIdx.x*blockDim.x+threadIdx.x; double s=0; unsigned h=(unsigned)mix(t>>1);
  #pragma unroll 8
  for(int i=0;i<iters;++i){ h=h*1664525u+1013904223u; unsigned idx=(((h>>7)&mask)&~1u)|(t&1u); s+=__ldg(x+idx); }
  if(s==123.456) out[0]=s;
}
__global__ void gather_smem(const double* __restrict__ x, int words, int iters, double* out){
  extern __shared__ double sx[];
  for(int i=threadIdx.x;i<words;i+=blockDim.x) sx[i]=x[i];
  __syncthreads();
  unsigned t=blockIdx.x*blockDim.x+threadIdx.x; double s=0; unsigned h=(unsigned)mix(t); unsigned mask=words-1;
  #pragma unroll 8
  for(int i=0;i<iters;++i){ h=h*1664525u+1013904223u; s+=sx[(h>>7)&mask]; }
  if(s==123.456) out[0]=s;
}
__global__ void copy_kernel(const double4* __restrict__ a, double4* __restrict__ b, long long n){
  long long i=blockIdx.x*(long long)blockDim.x+threadIdx.x; if(i<n) b[i]=a[i];
}
__global__ void touch(const double4* __restrict__ a, long long n, double* out){
  long long i=blockIdx.x*(long long)blockDim.x+threadIdx.x; double s=0; long long st=(long long)gridDim.x*blockDim.x;
  for(;i<n;i+=st){double4 v=a[i]; s+=v.x+v.w;} if(s==123.456) out[0]=s;
}
int main(){
  const size_t maxb=512ull<<20; double* x; cudaMalloc(&x,maxb); cudaMemset(x,0,maxb); double* out; cudaMalloc(&out,64);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaResourceDesc rd; memset(&rd,0,sizeof(rd)); rd.resType=cudaResourceTypeLinear; rd.res.linear.devPtr=x; rd.res.linear.desc=cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes=128ull<<20;
  cudaTextureDesc td; memset(&td,0,sizeof(td)); td.readMode=cudaReadModeElementType; cudaTextureObject_t tex=0;
  cudaError_t te=cudaCreateTextureObject(&tex,&rd,&td,nullptr); printf("texture object: %s\n",cudaGetErrorString(te));
  const int grid=148*16, it=512; const double g=148.0*16*256*it/1e6;
  long long kibs[]={16,32,64,128,256,512,1024,4096,32768,65536};
  printf("random 8-byte gathers, %d CTAs x 256 thr, %d gathers/thread (power-of-two footprints)\n%10s %12s %12s %14s\n",grid,it,"KiB","ldg Gg/s","tex Gg/s","ldg-pairs Gg/s");
  for(long long kb:kibs){ unsigned mask=(unsigned)((kb<<10)/8-1); float ms[3]={0,0,0};
    for(int v=0;v<3;++v) for(int rep=0;rep<3;++rep){ cudaEventRecord(e0);
      if(v==0) gather_mask<<<grid,256>>>(x,mask,it,out); else if(v==1){ if(te==cudaSuccess) gather_tex<<<grid,256>>>(tex,mask,it,out);} else gather_pairs<<<grid,256>>>(x,mask,it,out);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[v],e0,e1);}
    printf("%10lld %12.1f %12.1f %14.1f\n",kb,g/ms[0],te==cudaSuccess?g/ms[1]:0.0,g/ms[2]); }
  printf("random 8-byte reads from shared memory (148*4 CTAs x 512 thr, %d reads/thread)\n%10s %12s\n",2048,"KiB","Gg/s");
  for(int kb:{16,64,128}){ int words=(kb<<10)/8; cudaFuncSetAttribute(gather_smem,cudaFuncAttributeMaxDynamicSharedMemorySize,kb<<10); float ms=0;
    for(int rep=0;rep<3;++rep){ cudaEventRecord(e0); gather_smem<<<148*4,512,kb<<10>>>(x,words,2048,out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);}
    printf("%10d %12.1f\n",kb,148.0*4*512*2048/1e6/ms); }
  // flushed copies of small sizes
  double4 *a,*b; size_t big=1ull<<30; cudaMalloc(&a,big); cudaMalloc(&b,big); cudaMemset(a,0,big); cudaMemset(b,0,big);
  double4* fl; cudaMalloc(&fl,512ull<<20); cudaMemset(fl,0,512ull<<20);
  printf("device copy, L2 flushed by a 512 MiB read sweep before each run (bytes = read+write)\n%12s %10s %10s\n","MB moved","us","GB/s");
  for(double mb:{21.0,42.0,84.0,168.0,336.0,1342.0}){ long long n=(long long)(mb*1e6/2/32); float best=1e9;
    for(int rep=0;rep<5;++rep){ touch<<<148*8,256>>>(fl,(512ll<<20)/32,out); cudaEventRecord(e0); copy_kernel<<<(unsigned)((n+255)/256),256>>>(a,b,n); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    printf("%12.1f %10.2f %10.1f\n",mb,best*1e3,n*64.0/best/1e6); }
  printf("err=%s\n",cudaGetErrorString(cudaGetLastError())); return 0; }

// die_probe.cu -- does the L2 behave as TWO caches of ~63 MiB for random gathers because every die keeps its own
// copy of the lines its SMs touch?  If so, a kernel whose SMs only gather from "their" half of x would see the
// whole 126 MiB.  (1) map SMs to dies by L2-hit latency to one 2 KiB chunk (near die ~234 cycles, far die ~262);
// (2) random 8-byte gathers over F MiB, once with every CTA roaming all of x and once with the CTAs of die d
// confined to half d of x.  Development tool, not part of the library.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long mix(unsigned long long z){z+=0x9E3779B97F4A7C15ull;z=(z^(z>>30))*0xBF58476D1CE4E5B9ull;z=(z^(z>>27))*0x94D049BB133111EBull;return z^(z>>31);}

// one thread per CTA chases a pointer ring inside a 2 KiB chunk with L1 bypassed
__global__ void latency(const unsigned *ring, int hops, unsigned *sm_of_cta, unsigned *cycles)
{
    if (threadIdx.x) return;
    unsigned idx = 0;
    for (int i = 0; i < 64; ++i) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(idx) : "l"(ring + idx));  // warm L2
    long long t0 = clock64();
    for (int i = 0; i < hops; ++i) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(idx) : "l"(ring + idx));
    long long t1 = clock64();
    sm_of_cta[blockIdx.x] = smid();
    cycles[blockIdx.x] = (unsigned)((t1 - t0) / hops) + (idx == 12345u);
}

// MODE 0: every CTA gathers from all n elements; MODE 1: the CTAs of die d from half d
template <int MODE>
__global__ void gather(const double *__restrict__ x, unsigned long long n, int iters, const unsigned char *__restrict__ die_of_sm, double *out)
{
    unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long h = mix(t);
    unsigned long long lo = 0, span = n;
    if (MODE == 1) { span = n / 2; lo = die_of_sm[smid()] ? span : 0; }
    double s = 0;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        h = h * 6364136223846793005ull + 1442695040888963407ull;
        s += __ldg(x + lo + (h >> 20) % span);
    }
    if (s == 123.456) out[0] = s;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    // (1) die map
    unsigned h_ring[512];
    for (int i = 0; i < 512; ++i) h_ring[i] = (i * 37 + 11) % 512;  // a permutation ring over one 2 KiB chunk
    unsigned *ring; cudaMalloc(&ring, 1 << 20); cudaMemcpy(ring, h_ring, sizeof(h_ring), cudaMemcpyHostToDevice);
    const int ctas = sms * 4;
    unsigned *d_sm, *d_cyc; cudaMalloc(&d_sm, ctas * 4); cudaMalloc(&d_cyc, ctas * 4);
    latency<<<ctas, 32>>>(ring, 2000, d_sm, d_cyc);
    std::vector<unsigned> smv(ctas), cyc(ctas);
    cudaMemcpy(smv.data(), d_sm, ctas * 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc.data(), d_cyc, ctas * 4, cudaMemcpyDeviceToHost);
    std::vector<unsigned> lat(256, 0);
    for (int i = 0; i < ctas; ++i) if (smv[i] < 256) lat[smv[i]] = std::max(lat[smv[i]], cyc[i]);
    std::vector<unsigned> seen;
    for (int s = 0; s < 256; ++s) if (lat[s]) seen.push_back(lat[s]);
    std::sort(seen.begin(), seen.end());
    const unsigned cut = seen.empty() ? 0 : (seen.front() + seen.back()) / 2;
    std::vector<unsigned char> die(256, 0);
    int n0 = 0, n1 = 0;
    for (int s = 0; s < 256; ++s) if (lat[s]) { die[s] = lat[s] > cut; (die[s] ? n1 : n0)++; }
    printf("SMs %d, seen %zu; L2-hit latency to one 2 KiB chunk: min %u max %u cycles, cut %u -> near die %d SMs, far die %d SMs\n",
           sms, seen.size(), seen.empty() ? 0 : seen.front(), seen.empty() ? 0 : seen.back(), cut, n0, n1);
    printf("latencies by SM id:");
    for (int s = 0; s < 256; ++s) if (lat[s]) printf(" %u", lat[s]);
    printf("\n");
    unsigned char *d_die; cudaMalloc(&d_die, 256); cudaMemcpy(d_die, die.data(), 256, cudaMemcpyHostToDevice);
    // (2) gathers
    const size_t maxb = 512ull << 20; double *x; cudaMalloc(&x, maxb); cudaMemset(x, 0, maxb); double *out; cudaMalloc(&out, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 16, it = 256; const double g = (double)grid * 256 * it / 1e6;
    printf("random 8-byte gathers, %d CTAs x 256 thr, %d per thread\n%8s %16s %22s\n", grid, it, "MiB", "roam all Gg/s", "die-affine halves Gg/s");
    for (int mb : {32, 48, 64, 80, 96, 112, 128, 160, 192, 256}) {
        const unsigned long long n = ((size_t)mb << 20) / 8; float ms[2] = {0, 0};
        for (int v = 0; v < 2; ++v) for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) gather<0><<<grid, 256>>>(x, n, it, d_die, out); else gather<1><<<grid, 256>>>(x, n, it, d_die, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[v], e0, e1);
        }
        printf("%8d %16.1f %22.1f\n", mb, g / ms[0], g / ms[1]);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// common.cuh -- device state, error latch, and the sm_100a load/store + warp primitives shared by all
// kernel families of libspmv_b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#define SPMV_B200_NO_HOST_HEADERS 1
#include "../../include/spmv_b200.h"

namespace sb {

// ------------------------------------------------------------------------------------------------
// error latch (the public API returns void, reference include/spmv.h)
// ------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
bool cuda_ok(cudaError_t e, const char *what, const char *file, int line);
#define SB_CUDA(expr) ::sb::cuda_ok((expr), #expr, __FILE__, __LINE__)
#define SB_TRY(expr) do { if (!SB_CUDA(expr)) return false; } while (0)

void count_launch(int n = 1);

constexpr int kThreads = 256;       // CTA size of every spmv-path kernel
constexpr int kWarpsPerCta = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kPipeChunks = 8;      // row chunks / x pieces of the pipelined host-pointer path
constexpr int kMaxPieces = 64;      // >= kMaxBands (csr_kernels.cuh)
constexpr int kMaxPeers = 8;        // extra y destinations of the fused SpMV + all-gather

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Device-resident state hanging off spmv_Handle::extraHandle.
// ------------------------------------------------------------------------------------------------
struct DeviceState {
    uint32_t magic = 0x5b200a11u;
    // The handle owns staging buffers, partial-sum arrays and events (the reference's handle is read-only in spmv()
    // for most methods, SURVEY 8b "Threading"): calls on ONE handle from several threads are serialised here, so
    // that their launches reach the stream call after call instead of interleaved
    std::mutex mu;
    int device = 0;
    cudaStream_t stream = nullptr;  // legacy default stream unless spmv_b200_set_stream
    int m = 0, n = 0, nnz = 0;
    int vsize = 8;                  // 8 = fp64, 4 = fp32
    int requested = 0;              // SPMV_METHODS asked for at create
    int auto_method = -1;           // option "auto": the SPMV_METHODS value create picked instead (-1: not used)
    int kernel = SPMV_B200_KERNEL_NONE;
    bool ok = false;                // false => spmv() is a no-op
    bool has_empty_rows = false;
    int dev_sms = 0;
    long long dev_l2 = 0;
    int layout_fallbacks = 0;       // optional layouts (band copy, band segments, row bins) that could not be built
    int n_peers = 0;                // spmv_b200_set_y_peers
    void *peers[kMaxPeers] = {};

    // CSR on the device (owned upload, or the caller's device arrays adopted in place)
    int *rowptr = nullptr, *col = nullptr;
    void *val = nullptr;
    bool owns_csr = false;
    bool released_csr = false;      // the upload was freed after a re-laid-out copy took over

    // pageable host x / y that keep coming back (the reference's drivers reuse one X and one Y for every call) are
    // page-locked in place after the second sighting, so that the copies run at PCIe speed instead of being
    // staged by the driver.  Opt-in (option "pin_host" / SPMV_B200_PIN_HOST=1): the caller must not free such a
    // buffer while the handle lives.  Undone at clear / destroy or when the caller switches buffers
    struct PinSlot { const void *ptr = nullptr; size_t bytes = 0; int seen = 0; bool registered = false; };
    PinSlot pin[2];
    bool pin_host = false;
    // staging for host-side x / y; the pipelined host path (CSR-vector kernel) copies x in pieces on s_in,
    // computes row chunks as soon as their prefix of x has arrived, and returns finished chunks of y on s_out
    void *x_stage = nullptr, *y_stage = nullptr;
    bool pipeline = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kMaxPieces] = {}, ev_out[kPipeChunks] = {}, ev_start = nullptr, ev_x = nullptr;
    int chunk_xmax[kPipeChunks] = {};

    // Method_Parallel
    int tpr = 0;
    // Active matrix view used by every layout builder and kernel: the CSR itself, or (x_bands > 1) its
    // band-major copy, in which virtual row b*m + r holds the entries of row r whose column lies in band
    // b = col / band_cols.  Kernels then write the virtual y (v_y) and band_reduce_kernel folds it.
    int a_m = 0;
    int *a_rowptr = nullptr, *a_col = nullptr;
    void *a_val = nullptr;
    int x_bands = 1, band_cols = 0;
    double far_fraction = -1.0;     // share of entries far from the diagonal (automatic band decision)
    int *v_rowptr = nullptr, *v_col = nullptr;
    void *v_val = nullptr, *v_y = nullptr;
    // Method_Balanced: row blocks; ref_splitter mirrors the reference with the caller's nthreads
    int parts = 0;
    int *splitter = nullptr, *ref_splitter = nullptr;
    int ref_T = 0;
    // Method_Balanced2 / Method_Balanced_Yid: tiles + carries
    int tiles = 0, tile_items = 0;
    int *tile_rows = nullptr;       // nnz-split: row containing the first nnz of each tile [tiles+1]
    int2 *merge_coords = nullptr;   // merge-path: (row, nnz) start of each tile [tiles+1]
    void *carry_val = nullptr;      // [tiles] partial sum that belongs to a row started earlier
    int *carry_row = nullptr;       // [tiles] that row, or -1
    void *carry2_val = nullptr;     // level-2 carry list of the two-level fix-up: 2 entries per group of 64 tiles
    int *carry2_row = nullptr;
    // Method_SellCSigma
    int sigma = 0, banner = 0, slices = 0, sell_variant = 0;
    long long padded = 0;
    int *sell_perm = nullptr, *sell_width = nullptr, *sell_full = nullptr, *sell_col = nullptr;
    long long *sell_slice_ptr = nullptr;
    void *sell_val = nullptr;
    // Method_Parallel on short-row matrices with hubs: rows binned by length class (<= 8, <= 32, <= 128), one
    // launch per bin with 1 / 4 / 16 lanes per row; longer rows are on the long-row list
    bool binned = false;
    int *bin_list = nullptr;        // row ids, bin after bin, ascending inside a bin
    int bin_ptr[4] = {};
    // hyper-sparse column bands as band segments (band_seg.cuh): x_bands = K but the active view stays the CSR
    int coo_bands = 0, coo_tiles = 0;  // (option / info key names kept from round 1: "coo_bands")
    int seg_ptr[kMaxPieces + 1] = {};  // first slot of every band in seg_col / seg_val (aligned to 4), [K] = end
    int seg_cnt[kMaxPieces] = {};      // entries of every band
    int seg_tile0[kMaxPieces + 1] = {};  // first tile of every band
    long long seg_total = 0;           // segments = (band, row) pairs with at least one entry
    int seg_groups = 0;                // 32-row groups of the merge pass
    int seg_ctas = 0;                  // persistent CTAs per SM of pass 1
    int seg_grid = 0;                  // persistent CTAs of pass 1 (device-wide occupancy, set at the first launch)
    bool seg_cross = false;            // some segment crosses a tile boundary: the carry fix-up is needed
    bool seg_mask64 = false;           // K > 32: 64-bit row masks
    int *seg_col = nullptr;            // column index | segment-end bit 31
    void *seg_ent_val = nullptr;       // values, band-major
    void *seg_mask = nullptr;          // per row: bands in which it has a segment
    int4 *seg_tile_ent = nullptr;      // per tile: (first slot, entries, L2-prefetch duty: first x element, elements)
    int *seg_tile_seg0 = nullptr;      // per 256-entry chunk of a tile (8 per tile): index of its first segment sum
    int *seg_gbase = nullptr;          // [seg_groups][K]: position of a row group's first segment sum in band b's list
    int *seg_ticket = nullptr;         // ticket counter of pass 1's dynamic tile schedule
    void *seg_sums = nullptr;          // [seg_total] segment sums, band-major (written by pass 1, read by pass 2)
    // long rows / row tails left over by the main kernel of Method_Parallel and Method_SellCSigma (long_rows.cuh)
    int long_thr = 0x7fffffff;      // CSR-vector kernels skip rows longer than this
    int lr_rows = 0, lr_segs = 0;
    bool lr_accumulate = false;     // true: the main kernel already wrote the head of the row (SELL)
    int *lr_row = nullptr, *lr_start = nullptr, *lr_seg_ptr = nullptr, *lr_seg_row = nullptr;
    void *lr_partial = nullptr;
    // Method_CSR5SPMV (omega = 32)
    int c5_sigma = 0, c5_p = 0, c5_bit_y = 0, c5_bit_ss = 0, c5_num_offsets = 0, c5_tail_start = 0;
    uint32_t *c5_tile_ptr = nullptr, *c5_tile_desc = nullptr;
    int *c5_off_ptr = nullptr, *c5_off = nullptr, *c5_col = nullptr;
    void *c5_val = nullptr;
};

// ------------------------------------------------------------------------------------------------
// sm_100a memory primitives.
//  * matrix streams (ColIdx / Val) are read exactly once per SpMV: non-coherent path, L2 evict-first (the tile
//    kernels also keep them out of L1; band_seg.cuh stages them with bulk copies instead);
//  * x is the only re-used operand: read-only path with an L2 evict-last policy so the streams do
//    not push it out of the 126 MB L2.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---- gathers from x (evict-last) ----
__device__ __forceinline__ double ldg_x(const double *p, uint64_t pol)
{
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_x(const float *p, uint64_t pol)
{
    float r;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}

// ---- scalar streaming loads ----
__device__ __forceinline__ int ldg_stream(const int *p)
{
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double *p)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// ---- scalar streaming loads with an explicit L2 policy (evict-first for the matrix streams) ----
__device__ __forceinline__ int ldg_stream(const int *p, uint64_t pol)
{
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float *p, uint64_t pol)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double *p, uint64_t pol)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}

// ---- scalar matrix-stream loads that DO allocate in L1 (L2 evict-first) ----
__device__ __forceinline__ int ldg_cached(const int *p, uint64_t pol)
{
    int r;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_cached(const float *p, uint64_t pol)
{
    float r;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ldg_cached(const double *p, uint64_t pol)
{
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}

// ---- y stores: written once, never re-read by the kernel ----
__device__ __forceinline__ void stg_y(double *p, double v)
{
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void stg_y(float *p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- fused SpMV + all-gather: extra destinations of y (peer GPUs' next-x, mapped through CUDA IPC) ----
template <typename T>
struct PeerList {
    int n;
    T *p[kMaxPeers];
};
template <bool PEERS, typename T>
__device__ __forceinline__ void store_y(T *y, const PeerList<T> &peers, long long row, T v)
{
    stg_y(y + row, v);
    if (PEERS) {  // NVLink peer stores, overlapped with the SpMV (separate instantiation: the plain kernels pay nothing)
        // constant indices only: a dynamically indexed kernel-parameter array would be copied to local
        // memory by every thread (measured: +20-35 % on the L1TEX-bound kernels)
#pragma unroll
        for (int i = 0; i < kMaxPeers; ++i)
            if (i < peers.n) stg_y(peers.p[i] + row, v);
    }
}

__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }

// sub-warp butterfly sum over `width` lanes (power of two); every lane of the group gets the total
template <typename T>
__device__ __forceinline__ T group_sum(T v, int width)
{
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <typename T, int WIDTH>
__device__ __forceinline__ T group_sum_c(T v)
{
#pragma unroll
    for (int o = WIDTH >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// number of entries of a[0..size) that are <= key: the reference's
// binary_search_right_boundary_kernel (parallel_balanced_spmv.c:17-37), usable on host and device.
__host__ __device__ __forceinline__ int right_boundary(const int *a, int key, int size)
{
    int start = 0, stop = size - 1;
    while (stop >= start) {
        const int median = (int)(((long long)stop + start) >> 1);
        if (key >= a[median]) start = median + 1; else stop = median - 1;
    }
    return start;
}
#endif  // __CUDACC__

}  // namespace sb

// band_coo.cuh -- column bands for matrices whose bands are hyper-sparse.
//
// The band-major copy of csr_kernels.cuh stores band b of row r as virtual row b*m + r of a CSR.  When x is so
// large that the number of bands K approaches the mean row length (BASELINE.json config C5: 16 non-zeros per
// row, x = 2 GiB => K = 46, 0.35 non-zeros per row and band) the K*m+1 virtual row pointers outweigh the
// matrix.  Here a band is instead a COO list (row, col, val) sorted by row -- the entries of the CSR in their
// original order, stably bucketed by col / band_cols.  One launch per band, bands in ascending order:
//   * x gathers of a band stay inside an L2-resident slice of x (the point of banding);
//   * a CTA takes a tile of 2048 consecutive entries: coalesced loads, products into shared memory, then every
//     thread walks 8 consecutive products and closes the row segments that end inside its stretch; the open
//     segment at a thread's end travels to the following threads through a block-wide scan-by-key;
//   * a segment that ends inside the tile is added to y by exactly one thread (plain read-modify-write: within a
//     launch no other thread touches that row, earlier bands are earlier launches); the segment still open at
//     the tile's end goes to carry_val[tile] and carry_fixup_kernel adds the band's carries in tile order right
//     after the band's launch (per band: the same row may carry in several bands).
// No atomics, fixed order: bitwise reproducible.  The reference has no counterpart (its x lives in one NUMA
// domain's DRAM behind the CPU caches); this is the GPU answer to "x does not fit the last-level cache".
#pragma once
#include "common.cuh"
#include "tile_kernels.cuh"

namespace sb {

constexpr int kCooIpt = 8;
constexpr int kCooTile = kThreads * kCooIpt;

// row id of every CSR entry + its band (the sort key)
__global__ void coo_expand_kernel(int m, int band_cols, int bands, const int *__restrict__ rowptr,
                                  const int *__restrict__ col, int *__restrict__ ent_row,
                                  unsigned char *__restrict__ ent_band)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    for (int j = rowptr[r]; j < rowptr[r + 1]; ++j) {
        int b = col[j] / band_cols;
        b = b < 0 ? 0 : (b >= bands ? bands - 1 : b);
        ent_row[j] = r;
        ent_band[j] = (unsigned char)b;
    }
}

template <typename T>
__global__ void coo_gather_kernel(int nnz, const int *__restrict__ order, const int *__restrict__ ent_row,
                                  const int *__restrict__ col, const T *__restrict__ val, int *__restrict__ brow,
                                  int *__restrict__ bcol, T *__restrict__ bval)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int j = order[i];
    brow[i] = ent_row[j];
    bcol[i] = col[j];
    bval[i] = val[j];
}

__global__ void coo_iota_kernel(int n, int *__restrict__ a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

// band_ptr[b] = first sorted entry whose band is >= b
__global__ void coo_band_ptr_kernel(int nnz, int bands, const unsigned char *__restrict__ sorted_band,
                                    int *__restrict__ band_ptr)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > bands) return;
    int lo = 0, hi = nnz;
    while (lo < hi) {
        const int mid = (int)(((long long)lo + hi) >> 1);
        if ((int)sorted_band[mid] < b) lo = mid + 1; else hi = mid;
    }
    band_ptr[b] = lo;
}

// entries [e0, e1) of one band; tile index of this CTA in the carry arrays = tile_base + blockIdx.x
template <typename T>
__global__ void __launch_bounds__(kThreads)
band_coo_kernel(int e0, int e1, int tile_base, const int *__restrict__ brow, const int *__restrict__ bcol,
                const T *__restrict__ bval, const T *__restrict__ x, T *__restrict__ y,
                T *__restrict__ carry_val, int *__restrict__ carry_row)
{
    __shared__ int s_row[kCooTile + kCooTile / 8 + 2];
    __shared__ T s_prod[kCooTile + kCooTile / 8 + 2];
    __shared__ int s_k[kWarpsPerCta];
    __shared__ T s_v[kWarpsPerCta];

    const uint64_t pl = policy_evict_last(), pf = policy_evict_first();
    const int tid = threadIdx.x;
    const int t0 = e0 + blockIdx.x * kCooTile;
    const int cnt = min(kCooTile, e1 - t0);
#pragma unroll
    for (int k = 0; k < kCooIpt; ++k) {
        const int i = tid + k * kThreads;
        if (i < cnt) s_row[pad8(i)] = ldg_stream(brow + t0 + i, pf);
    }
    tile_products<T, kCooIpt>(tid, t0, cnt, bcol, bval, x, pl, s_prod);
    // the row that follows the tile inside this band (-1: the band ends here, the last segment is closed)
    if (tid == 0) s_row[pad8(cnt)] = (t0 + cnt < e1) ? brow[t0 + cnt] : -1;
    __syncthreads();

    const int b = tid * kCooIpt;
    int out_row[kCooIpt];
    T out_val[kCooIpt];
    T run = 0;
    int open_key = -2 - tid;  // "no open segment": a key no neighbour shares
#pragma unroll
    for (int k = 0; k < kCooIpt; ++k) {
        out_row[k] = -1;
        out_val[k] = 0;
        const int i = b + k;
        if (i < cnt) {
            const int r = s_row[pad8(i)];
            run += s_prod[pad8(i)];
            if (s_row[pad8(i + 1)] != r) {  // the segment of row r ends here
                out_row[k] = r;
                out_val[k] = run;
                run = 0;
                open_key = -2 - tid;
            } else {
                open_key = r;               // still open (so far)
            }
        }
    }
    // y += segment sums.  The rows a thread closes are distinct and nobody else touches them in this launch, so
    // all reads of y are issued up front -- 8 independent loads that fly while the block-wide scan runs -- and
    // the writes follow at the end.
    T old[kCooIpt];
#pragma unroll
    for (int k = 0; k < kCooIpt; ++k) old[k] = out_row[k] >= 0 ? y[out_row[k]] : (T)0;  // default caching: neighbours share sectors
    // (evict-first / no-allocate hints on this sweep measured 30 % slower: the 4 rows of a sector are touched by
    // different lanes at different times and would be fetched from DRAM again)

    int ex_key, agg_key;
    T ex_val, agg_val;
    block_scan_by_key<T>(open_key, run, ex_key, ex_val, agg_key, agg_val, s_k, s_v);

    bool first = true;
#pragma unroll
    for (int k = 0; k < kCooIpt; ++k) {
        if (out_row[k] >= 0) {
            T v = out_val[k];
            if (first && ex_key == out_row[k]) v = ex_val + v;  // the part that sits in the preceding threads
            first = false;
            stg_y(y + out_row[k], old[k] + v);
        }
    }
    if (tid == 0) {
        carry_row[tile_base + blockIdx.x] = agg_key >= 0 ? agg_key : -1;  // row open at the tile's end
        carry_val[tile_base + blockIdx.x] = agg_val;
    }
}

}  // namespace sb

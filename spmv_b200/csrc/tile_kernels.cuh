// tile_kernels.cuh -- kernels whose unit of work is a fixed-size tile of the non-zero stream:
//   merge_path_kernel   Method_Balanced2     tiles of (rows + nnz) merge items
//   nnz_split_kernel    Method_Balanced_Yid  tiles of nnz only
//   carry_fixup_kernel  the serial `Y[Yid[t]] += Ysum[t]` pass of the reference, in tile order
// No floating-point atomics anywhere: partial sums of rows that span tiles go to carry_val[tile] and
// are added by carry_fixup_kernel in ascending tile order, so results are bitwise reproducible.
#pragma once
#include "common.cuh"

namespace sb {

// smem index padding: one extra slot every 8 so that threads walking 8-element stretches hit
// distinct banks (stride 9 elements)
__device__ __forceinline__ int pad8(int i) { return i + (i >> 3); }

// Products of a tile's slice of the non-zero stream into shared memory: element i = tid + k*kThreads of the
// tile (coalesced).  Three separate passes -- all ColIdx / Val loads, then all x gathers, then the stores -- so
// that a thread has its 2*IPT stream loads and then its IPT gathers in flight together instead of one dependent
// round trip per element (the inline-asm loads keep their program order).
template <typename T, int IPT>
__device__ __forceinline__ void tile_products(int tid, int first, int count, const int *__restrict__ col,
                                              const T *__restrict__ val, const T *__restrict__ x, uint64_t pl,
                                              T *__restrict__ s_prod)
{
    int c[IPT];
    T v[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int i = tid + k * kThreads;
        c[k] = i < count ? ldg_stream(col + first + i) : -1;
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int i = tid + k * kThreads;
        v[k] = i < count ? ldg_stream(val + first + i) : (T)0;
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k)
        if (c[k] >= 0) v[k] = v[k] * ldg_x(x + c[k], pl);
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int i = tid + k * kThreads;
        if (i < count) s_prod[pad8(i)] = v[k];
    }
}

// ------------------------------------------------------------------------------------------------
// Precomputation (handle construction)
// ------------------------------------------------------------------------------------------------

// nnz-split: tile_rows[t] = row that contains non-zero t*tile_nnz -- the a9 formula with stride =
// tile_nnz (reference parallel_balanced2_spmv.c:41-53; the Yid variant parallel_balanced_Yid_spmv.c
// :16-53 derives the same rows through lower_bound).
__global__ void tile_rows_kernel(int tiles, int tile_nnz, int nnz, int m, const int *__restrict__ rowptr,
                                 int *__restrict__ tile_rows)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > tiles) return;
    long long b = (long long)t * tile_nnz;
    if (b > nnz) b = nnz;
    tile_rows[t] = right_boundary(rowptr, (int)b, m + 1) - 1;
}

// merge-path: start coordinate (row, nnz) of every tile on the diagonal t*items of the merge of the
// row-end offsets RowPtr[1..m] with the non-zero indices 0..nnz-1.
__global__ void merge_coords_kernel(int tiles, int items, int nnz, int m, const int *__restrict__ rowptr,
                                    int2 *__restrict__ coords)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > tiles) return;
    long long d = (long long)t * items;
    const long long total = (long long)m + nnz;
    if (d > total) d = total;
    long long lo = d > nnz ? d - nnz : 0, hi = d < m ? d : m;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)rowptr[mid + 1] <= d - mid - 1) lo = mid + 1; else hi = mid;
    }
    coords[t] = make_int2((int)lo, (int)(d - lo));
}

// ------------------------------------------------------------------------------------------------
// Block-wide exclusive "sum by key" scan over (key, val) pairs with non-decreasing keys:
// prefix(i) = sum of val over the maximal run of threads j < i that ends at i-1 and shares its key.
// Fixed evaluation order.  Also returns the block aggregate (inclusive value of the last thread).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void block_scan_by_key(int key, T val, int &ex_key, T &ex_val, int &agg_key,
                                                  T &agg_val, int *s_k, T *s_v)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T v = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int k2 = __shfl_up_sync(kFull, key, o);
        const T v2 = __shfl_up_sync(kFull, v, o);
        if (lane >= o && k2 == key) v = v2 + v;
    }
    if (lane == 31) { s_k[warp] = key; s_v[warp] = v; }
    __syncthreads();
    int pk = -1;
    T pv = 0;
    for (int w = 0; w < warp; ++w) {
        pv = (s_k[w] == pk) ? pv + s_v[w] : s_v[w];
        pk = s_k[w];
    }
    if (warp > 0 && pk == key) v = pv + v;
    int ek = __shfl_up_sync(kFull, key, 1);
    T ev = __shfl_up_sync(kFull, v, 1);
    if (lane == 0) { ek = pk; ev = pv; }
    ex_key = ek;
    ex_val = ev;
    int ak = -1;
    T av = 0;
    for (int w = 0; w < kWarpsPerCta; ++w) {
        av = (s_k[w] == ak) ? av + s_v[w] : s_v[w];
        ak = s_k[w];
    }
    agg_key = ak;
    agg_val = av;
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Method_Balanced2 on the GPU: merge-path.  Replaces spmv_parallel_balanced2_cpp_{d,s} (reference
// src/src_spmv/parallel_balanced2_spmv.c:211-359).  The reference splits rows among threads by nnz
// and, when one row spans several threads, lets each compute a slice into Ysum[t] and adds
// `Y[Yid[t]] += Ysum[t]` serially (:277-282).  Here every CTA gets exactly kThreads*IPT merge items
// (row ends + non-zeros), every thread IPT of them, so neither long rows nor runs of empty rows can
// unbalance anything; the serial add survives as carry_fixup_kernel.
// ------------------------------------------------------------------------------------------------
template <typename T, int IPT>
__global__ void __launch_bounds__(kThreads)
merge_path_kernel(int m, int nnz, const int2 *__restrict__ coords, const int *__restrict__ rowptr,
                  const int *__restrict__ col, const T *__restrict__ val, const T *__restrict__ x,
                  T *__restrict__ y, T *__restrict__ carry_val, int *__restrict__ carry_row)
{
    constexpr int ITEMS = kThreads * IPT;
    __shared__ int s_rowend[ITEMS + 1];
    __shared__ T s_prod[ITEMS + ITEMS / 8 + 1];
    __shared__ int s_k[kWarpsPerCta];
    __shared__ T s_v[kWarpsPerCta];

    const uint64_t pl = policy_evict_last();
    const int tid = threadIdx.x;
    const int2 c0 = coords[blockIdx.x], c1 = coords[blockIdx.x + 1];
    const int row0 = c0.x, nz0 = c0.y;
    const int tile_rows = c1.x - row0, tile_nz = c1.y - nz0;

    // row-end offsets of the rows that end in this tile, plus the one still open at its end
    for (int i = tid; i <= tile_rows; i += kThreads) {
        const int r = row0 + 1 + i;
        s_rowend[i] = rowptr[r < m ? r : m];
    }
    // products, coalesced over the tile's slice of the non-zero stream
    tile_products<T, IPT>(tid, nz0, tile_nz, col, val, x, pl, s_prod);
    __syncthreads();

    // this thread's start on the tile-local merge path
    const int total = tile_rows + tile_nz;
    int diag = tid * IPT;
    if (diag > total) diag = total;
    int lo = diag > tile_nz ? diag - tile_nz : 0, hi = diag < tile_rows ? diag : tile_rows;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_rowend[mid] <= nz0 + (diag - mid - 1)) lo = mid + 1; else hi = mid;
    }
    int row_i = lo, nz_j = diag - lo;
    int cnt = total - diag;
    if (cnt > IPT) cnt = IPT;

    int out_row[IPT];
    T out_val[IPT];
    T running = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        out_row[k] = -1;
        out_val[k] = 0;
        if (k < cnt) {
            if (nz0 + nz_j < s_rowend[row_i]) {
                running += s_prod[pad8(nz_j)];
                ++nz_j;
            } else {
                out_row[k] = row_i;
                out_val[k] = running;
                running = 0;
                ++row_i;
            }
        }
    }

    // carry of the row still open at this thread's end -> the thread that ends it
    int ex_key, agg_key;
    T ex_val, agg_val;
    block_scan_by_key<T>(row_i, running, ex_key, ex_val, agg_key, agg_val, s_k, s_v);

    bool first = true;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        if (out_row[k] >= 0) {
            T v = out_val[k];
            if (first && ex_key == out_row[k]) v = ex_val + v;
            first = false;
            stg_y(y + row0 + out_row[k], v);
        }
    }
    if (tid == 0) {
        const int r = row0 + agg_key;  // row open at the end of the tile
        carry_row[blockIdx.x] = (r < m) ? r : -1;
        carry_val[blockIdx.x] = agg_val;
    }
}

// rows r0 .. r_last of one equal-nnz tile, TPR lanes per row, products taken from shared memory
template <typename T, int TPR>
__device__ __forceinline__ void nnz_split_rows(int t, int r0, int r_last, int tile_start, int tile_end, int lane, int warp,
                                               const int *__restrict__ rowptr, const T *__restrict__ s_prod,
                                               T *__restrict__ y, T *__restrict__ carry_val, int *__restrict__ carry_row)
{
    constexpr int groups_per_warp = 32 / TPR;
    constexpr int stride_rows = kWarpsPerCta * groups_per_warp;
    const int sub = lane / TPR, sl = lane & (TPR - 1);
    for (int base = r0; base <= r_last; base += stride_rows) {  // warp-uniform trip count
        const int row = base + warp * groups_per_warp + sub;
        const bool valid = row <= r_last;
        int rs = 0, re = 0;
        if (valid) { rs = rowptr[row]; re = rowptr[row + 1]; }
        const int lo = max(rs, tile_start), hi = min(re, tile_end);
        // blocked: a row much longer than the tile's mean is walked by few lanes, and a plain chain of hundreds of
        // adds would exceed the 8*eps*sum|a x| bound (seen on R-MAT at full size: 12.8 eps)
        T sum = 0, blk = 0;
        int in_blk = 0;
        for (int j = lo + sl; j < hi; j += TPR) {
            blk += s_prod[pad8(j - tile_start)];
            if (++in_blk == 32) { sum += blk; blk = 0; in_blk = 0; }
        }
        sum += blk;
        sum = group_sum_c<T, TPR>(sum);
        if (valid && sl == 0) {
            if (rs < tile_start) {
                // continues a row begun in an earlier tile (only row == r0 can): carry it
                carry_val[t] = sum;
                carry_row[t] = row;
            } else if (rs < tile_end || re == rs) {
                // row starts in this tile (complete, or its first slice), or is empty
                if (re > rs || rs >= tile_start) stg_y(y + row, sum);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Method_Balanced_Yid on the GPU: equal-nnz tiles.  Replaces spmv_parallel_balanced_Yid_cpp_{d,s}
// (reference src/src_spmv/parallel_balanced_Yid_spmv.c:97-225): tile t owns non-zeros
// [t*TILE, (t+1)*TILE); rows wholly inside are written directly, the part of the first row that
// started in an earlier tile goes to carry_val[t] (the reference's begin_val, added serially at
// :151-156), a last row that continues past the tile is written directly and completed by the
// carries of the following tiles.  Rows are reduced by sub-warps sized from the tile's mean row
// length.  Tile 0 starts at row 0 and the last tile runs to row m-1, so leading/trailing empty rows
// get their zero (the reference skips them).
// ------------------------------------------------------------------------------------------------
template <typename T, int IPT>
__global__ void __launch_bounds__(kThreads)
nnz_split_kernel(int m, int nnz, const int *__restrict__ tile_rows, const int *__restrict__ rowptr,
                 const int *__restrict__ col, const T *__restrict__ val, const T *__restrict__ x,
                 T *__restrict__ y, T *__restrict__ carry_val, int *__restrict__ carry_row)
{
    constexpr int TILE = kThreads * IPT;
    __shared__ T s_prod[TILE + TILE / 8 + 1];

    const uint64_t pl = policy_evict_last();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;
    const int tile_start = t * TILE;  // tiles*TILE < 2^31 + TILE is guaranteed by nnz < 2^31 - TILE at create
    const int tile_end = min(tile_start + TILE, nnz);
    const int r0 = (t == 0) ? 0 : tile_rows[t];
    int r_last = tile_rows[t + 1];
    if (r_last > m - 1) r_last = m - 1;

    tile_products<T, IPT>(tid, tile_start, tile_end - tile_start, col, val, x, pl, s_prod);
    if (tid == 0) carry_row[t] = -1;
    __syncthreads();

    const int nrows = r_last - r0 + 1;
    const int avg = (tile_end - tile_start + nrows - 1) / max(nrows, 1);
    // lanes per row from the tile's mean row length (CTA-uniform), ~4 products per lane; compile-time bodies so
    // that the butterfly has a fixed depth and the row walk carries no loop over a run-time lane count
#define SB_ROWS(N) nnz_split_rows<T, N>(t, r0, r_last, tile_start, tile_end, lane, warp, rowptr, s_prod, y, carry_val, carry_row)
    if (avg <= 4) SB_ROWS(1);
    else if (avg <= 8) SB_ROWS(2);
    else if (avg <= 16) SB_ROWS(4);
    else if (avg <= 32) SB_ROWS(8);
    else if (avg <= 64) SB_ROWS(16);
    else SB_ROWS(32);
#undef SB_ROWS
}

// ------------------------------------------------------------------------------------------------
// The reference's serial tail `for tid: Y[Yid[tid]] += Ysum[tid]` (parallel_balanced2_spmv.c:277-282,
// parallel_balanced_Yid_spmv.c:151-156, and the CSR5 calibrator csr5_spmv_avx2.h:320-335), made
// parallel across rows but kept sequential -- ascending tile order -- within a row.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void carry_fixup_kernel(int tiles, const int *__restrict__ carry_row,
                                   const T *__restrict__ carry_val, T *__restrict__ y)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tiles) return;
    const int row = carry_row[t];
    if (row < 0) return;
    if (t > 0 && carry_row[t - 1] == row) return;  // not the first tile carrying into this row
    T acc = carry_val[t];
    for (int u = t + 1; u < tiles && carry_row[u] == row; ++u) acc += carry_val[u];
    y[row] += acc;
}

// Two-level variant for long carry chains (a hub row of 2.5 M non-zeros is 4900 CSR5 tiles: one thread walking
// that chain costs ~70 us and piles up rounding error).  Level 1: a warp takes kCarryGroup = 32 consecutive tiles,
// one per lane (coalesced), and sums the runs of equal rows with a segmented shuffle scan; runs that lie wholly
// inside the group are added to y at once, the group's first and last run -- which may continue in the
// neighbouring groups -- go to a list of 2 entries per group, on which carry_fixup_kernel then runs as before
// (chains 32x shorter).  Same order every time: bitwise reproducible.
constexpr int kCarryGroup = 32;

template <typename T>
__global__ void __launch_bounds__(kThreads)
carry_group_kernel(int tiles, const int *__restrict__ carry_row, const T *__restrict__ carry_val,
                   T *__restrict__ y, int *__restrict__ g_row, T *__restrict__ g_val)
{
    const int g = (int)(((long long)blockIdx.x * kThreads + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if ((long long)g * kCarryGroup >= tiles) return;
    const int u = g * kCarryGroup + lane;
    const int r = u < tiles ? carry_row[u] : -1;
    T v = (u < tiles && r >= 0) ? carry_val[u] : (T)0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // inclusive sum over the run of equal rows ending at this lane
        const int r2 = __shfl_up_sync(kFull, r, o);
        const T v2 = __shfl_up_sync(kFull, v, o);
        if (lane >= o && r2 == r) v = v2 + v;
    }
    const int next = __shfl_down_sync(kFull, r, 1);
    const bool run_end = lane == 31 || next != r;
    const int first_row = __shfl_sync(kFull, r, 0), last_row = __shfl_sync(kFull, r, 31);
    // last lane of the FIRST run: rows are contiguous, so it is the last lane before the first mismatch with lane 0
    const unsigned same_as_first = __ballot_sync(kFull, r == first_row);
    const int first_end = (same_as_first == kFull) ? 31 : (__ffs(~same_as_first) - 2);
    const T head_val = __shfl_sync(kFull, v, first_end);
    const T last_val = __shfl_sync(kFull, v, 31);
    if (run_end && r >= 0 && lane > first_end && r != last_row) y[r] += v;  // a run wholly inside the group
    if (lane == 0) {
        const bool one_run = first_end == 31;
        g_row[2 * g] = first_row;
        g_val[2 * g] = first_row >= 0 ? head_val : (T)0;
        // a group that is ONE run keeps its row in both slots (value once): the level-2 chain has no gap
        g_row[2 * g + 1] = last_row;
        g_val[2 * g + 1] = (one_run || last_row < 0) ? (T)0 : last_val;
    }
}

}  // namespace sb

// csr5.cuh -- Method_CSR5SPMV on the GPU: CSR5 with omega = 32 (one warp per tile), sigma <= 16.
//
// Builder: replaces generate_partition_pointer / generate_partition_descriptor /
// generate_partition_descriptor_offset / aosoa_transpose (reference
// src/src_spmv/csr5_avx2/avx2/format_avx2.h:7-425, driven by anonymouslib_avx2.h:112-242).  The bit
// packing of tile_ptr (bit 31 = "tile spans an empty row") and tile_desc (y_offset | scansum_offset |
// sigma bit-flags, MSB first) is the reference's, evaluated at omega = 32, so it can be diffed against
// the oracle's (omega, sigma)-parametric restatement, which in turn is proven bit-equal to the
// reference at its own (4, 16).  Unlike the reference the transpose goes into DEVICE COPIES; the
// caller's ColIdx / Val are never modified (the reference transposes them in place and back on
// destroy, format_avx2.h:381-395, anonymouslib_avx2.h:94-106).
//
// SpMV: replaces spmv_csr5_compute_kernel / partition_fast_track / spmv_csr5_calibrate_kernel /
// spmv_csr5_tail_partition_kernel (csr5_spmv_avx2.h:7-410).  Differences by design:
//   * the cross-lane step is a segmented shuffle scan, not "prefix-sum then subtract"
//     (csr5_spmv_avx2.h:262-283): no cancellation error, per-row error bound holds;
//   * the first segment of a tile (a row begun in an earlier tile) is never `+=`-ed into y by the
//     tile; it goes to carry_val[tile] and carry_fixup_kernel adds carries in tile order
//     (the reference's calibrator, generalised from per-thread to per-tile): no atomics, bitwise
//     reproducible;
//   * y of empty rows is written (0), which the reference omits (csr5_spmv.cpp:50).
#pragma once
#include "common.cuh"

namespace sb {

constexpr int kC5Omega = 32;

// a16 s1 (format_avx2.h:7-25)
__global__ void c5_tile_ptr_kernel(int p, int sigma, int nnz, int m, const int *__restrict__ rowptr,
                                   uint32_t *__restrict__ tile_ptr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > p) return;
    long long b = (long long)t * sigma * kC5Omega;
    if (b > nnz) b = nnz;
    tile_ptr[t] = (uint32_t)(right_boundary(rowptr, (int)b, m + 1) - 1);
}

// a16 s2 (format_avx2.h:27-57): mark tiles whose row span contains an empty row.  Reads the clean
// values written by s1 from `src` and writes `dst`, so that neighbours never see a half-updated word.
__global__ void c5_tile_dirty_kernel(int p, int m, const int *__restrict__ rowptr,
                                     const uint32_t *__restrict__ src, uint32_t *__restrict__ dst)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > p) return;
    uint32_t start = src[t];
    if (t < p) {
        const uint32_t stop = src[t + 1];
        if (start != stop) {
            for (uint32_t r = start; r <= stop && r < (uint32_t)m; ++r)
                if (rowptr[r] == rowptr[r + 1]) { start |= 0x80000000u; break; }
        }
    }
    dst[t] = start;
}

// a17 s1 + s2 (format_avx2.h:80-217), one warp per tile, lane = CSR5 lane.  num_packet == 1 is
// guaranteed by sigma <= 32 - bit_all.
__global__ void c5_tile_desc_kernel(int p, int sigma, int bit_y, int bit_ss, const int *__restrict__ rowptr,
                                    const uint32_t *__restrict__ tile_ptr, uint32_t *__restrict__ desc,
                                    int *__restrict__ off_cnt)
{
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= p) return;
    const int bit_all = bit_y + bit_ss;
    if (t == p - 1) {  // the tail tile carries no descriptor (reference loops stop at p-1)
        desc[(long long)t * kC5Omega + lane] = 0;
        if (lane == 0) off_cnt[t] = 0;
        return;
    }
    const uint32_t tp = tile_ptr[t];
    const int row_start = (int)(tp & 0x7FFFFFFFu), row_stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
    const bool dirty = (tp >> 31) & 1;
    const int tile_nnz = kC5Omega * sigma;
    // s1: bit (ptr % sigma) of lane (ptr / sigma) % omega for every row start inside the tile
    uint32_t word = 0;
    for (int rbase = row_start; rbase <= row_stop; rbase += 32) {  // warp-uniform trip count
        const int rid = rbase + lane;
        const int ptr = rid <= row_stop ? rowptr[rid] : -1;
        const bool in = ptr >= 0 && (ptr / tile_nnz == t);
        const int lx = (ptr / sigma) % kC5Omega;
        const uint32_t bit = 1u << (31 - (ptr % sigma + bit_all));
        // route the bit to lane lx
        for (int src = 0; src < 32; ++src) {
            const int s_in = __shfl_sync(kFull, (int)in, src);
            const int s_lx = __shfl_sync(kFull, lx, src);
            const uint32_t s_bit = __shfl_sync(kFull, bit, src);
            if (s_in && s_lx == lane) word |= s_bit;
        }
    }
    if (row_start == row_stop) {  // fast-track tile: s2 skips it (format_avx2.h:151-152)
        desc[(long long)t * kC5Omega + lane] = word;
        if (lane == 0) off_cnt[t] = 0;
        return;
    }
    // s2: segments per lane
    const uint32_t flags = (word << bit_all) | ((lane == 0) ? 0x80000000u : 0u);
    const int first = (flags >> 31) & 1;
    const uint32_t rest = (sigma >= 32) ? (flags & 0x7FFFFFFFu)
                                        : (flags & 0x7FFFFFFFu & ~((1u << (32 - sigma)) - 1u));
    const int stop = __popc(rest);
    const int present = (first | (stop > 0)) ? 1 : 0;
    int segn = stop - (first ? 0 : 1) + present;
    if (segn < 0) segn = 0;
    int incl = segn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    const int excl = incl - segn;
    const int total = __shfl_sync(kFull, incl, 31);
    const uint32_t pmask = __ballot_sync(kFull, present);
    int scansum = 0;
    if (present) {
        const uint32_t above = (lane == 31) ? 0u : (pmask >> (lane + 1));
        scansum = above ? (__ffs(above) - 1) : (31 - lane);
    }
    const int y_offset = lane ? excl - 1 : 0;
    word |= (uint32_t)y_offset << (32 - bit_y);
    word |= (uint32_t)scansum << (32 - bit_all);
    desc[(long long)t * kC5Omega + lane] = word;
    if (lane == 0) off_cnt[t] = dirty ? total : 0;
}

// a17 offsets (format_avx2.h:256-323): true y index (relative to row_start+1) of every segment start
__global__ void c5_desc_offset_kernel(int p, int sigma, int bit_y, int bit_ss, const int *__restrict__ rowptr,
                                      const uint32_t *__restrict__ tile_ptr, const uint32_t *__restrict__ desc,
                                      const int *__restrict__ off_ptr, int *__restrict__ off)
{
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= p - 1) return;
    const uint32_t tp = tile_ptr[t];
    if (!((tp >> 31) & 1)) return;
    const int row_start = (int)(tp & 0x7FFFFFFFu), row_stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
    const int bit_all = bit_y + bit_ss;
    const uint32_t d = desc[(long long)t * kC5Omega + lane];
    int y_offset = (int)(d >> (32 - bit_y));
    const uint32_t flags = d << bit_all;
    const int base = off_ptr[t];
    for (int i = 0; i < sigma; ++i) {
        if (!((flags >> (31 - i)) & 1)) continue;
        if (i == 0 && lane == 0) continue;  // the reference's `local_bit && lane_id` (format_avx2.h:292)
        const int idx = t * kC5Omega * sigma + lane * sigma + i;
        off[base + y_offset] = right_boundary(rowptr + row_start + 1, idx, row_stop - row_start) - 1;
        ++y_offset;
    }
}

// a18 (format_avx2.h:347-425) into copies: element x*sigma + y of a full, non-fast-track tile moves
// to y*omega + x; everything else (fast-track tiles, the tail tile) is copied unchanged.
template <typename T>
__global__ void c5_transpose_kernel(int nnz, int sigma, int p, const uint32_t *__restrict__ tile_ptr,
                                    const int *__restrict__ col, const T *__restrict__ val,
                                    int *__restrict__ tcol, T *__restrict__ tval)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int tile_nnz = kC5Omega * sigma;
    const int t = (int)(i / tile_nnz);
    long long dst = i;
    if (t < p - 1 && tile_ptr[t] != tile_ptr[t + 1]) {
        const int idx = (int)(i - (long long)t * tile_nnz);
        const int yy = idx % sigma, xx = idx / sigma;
        dst = (long long)t * tile_nnz + (long long)yy * kC5Omega + xx;
    }
    tcol[dst] = col[i];
    tval[dst] = val[i];
}

// ------------------------------------------------------------------------------------------------
// SpMV over the full tiles 0 .. p-2
// ------------------------------------------------------------------------------------------------
template <typename T, int SIGMA>
__global__ void __launch_bounds__(kThreads)
csr5_kernel(int p, int bit_y, int bit_ss, const uint32_t *__restrict__ tile_ptr,
            const uint32_t *__restrict__ desc, const int *__restrict__ off_ptr, const int *__restrict__ off,
            const int *__restrict__ tcol, const T *__restrict__ tval, const T *__restrict__ x,
            T *__restrict__ y, T *__restrict__ carry_val, int *__restrict__ carry_row)
{
    const uint64_t pl = policy_evict_last(), pf = policy_evict_first();
    const int t = (int)(((long long)blockIdx.x * kThreads + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= p - 1) return;
    const uint32_t tp = tile_ptr[t];
    const int row_start = (int)(tp & 0x7FFFFFFFu), row_stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
    const bool dirty = (tp >> 31) & 1;
    const int bit_all = bit_y + bit_ss;
    const uint32_t d = desc[(long long)t * kC5Omega + lane];
    const uint32_t flags = d << bit_all;  // bit 31-i = "element i of this lane starts a row"
    const bool tile_starts_row = (__shfl_sync(kFull, flags, 0) >> 31) & 1;
    const long long base = (long long)t * kC5Omega * SIGMA;

    // all 2*SIGMA coalesced streaming loads first ...
    int c[SIGMA];
    T v[SIGMA];
#pragma unroll
    for (int i = 0; i < SIGMA; ++i) c[i] = ldg_stream(tcol + base + i * kC5Omega + lane, pf);
#pragma unroll
    for (int i = 0; i < SIGMA; ++i) v[i] = ldg_stream(tval + base + i * kC5Omega + lane, pf);

    // ... and ALL gathers before the first use: v[i] becomes the product a_ij * x_j.  The segmented sums below
    // store finished rows as they go; with the gathers inside that loop every row end would expose one full
    // gather latency (the store needs the running sum, the next gather is issued behind the store).
#pragma unroll
    for (int i = 0; i < SIGMA; ++i) v[i] = v[i] * ldg_x(x + c[i], pl);

    if (row_start == row_stop) {
        // fast track (csr5_spmv_avx2.h:7-49): the whole tile lies inside one row
        T sum = 0;
#pragma unroll
        for (int i = 0; i < SIGMA; ++i) sum += v[i];
        sum = group_sum_c<T, 32>(sum);
        if (lane == 0) {
            if (tile_starts_row) { stg_y(y + row_start, sum); carry_row[t] = -1; }
            else { carry_val[t] = sum; carry_row[t] = row_start; }
        }
        return;
    }

    const int *offs = dirty ? off + off_ptr[t] : nullptr;
    int next_idx = (int)(d >> (32 - bit_y));  // index of this lane's first segment among the tile's
    T sum = 0, first_sum = 0;
    bool seen = false;
    int cur_row = row_start;  // row of the segment currently open in this lane (valid once seen)
#pragma unroll
    for (int i = 0; i < SIGMA; ++i) {
        if ((flags >> (31 - i)) & 1) {
            if (seen) stg_y(y + cur_row, sum);  // a row that starts and ends inside this lane
            else first_sum = sum;
            seen = true;
            sum = 0;
            if (lane == 0 && i == 0) {
                cur_row = row_start;  // the tile's first element starts row_start itself
            } else {
                cur_row = row_start + 1 + (dirty ? offs[next_idx] : next_idx);
                ++next_idx;
            }
        }
        sum += v[i];
    }
    if (!seen) { first_sum = sum; sum = 0; }

    // G[l] = first_sum[l] + first_sum[l+1] + ... up to and including the next lane that has a flag:
    // the part of a segment that lies in the lanes after the one where it started.
    T g = first_sum;
    bool done = seen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T g2 = __shfl_down_sync(kFull, g, o);
        const int d2 = __shfl_down_sync(kFull, (int)done, o);
        if (lane + o < 32 && !done) { g = g + g2; done = d2; }
    }
    T tail = __shfl_down_sync(kFull, g, 1);
    if (lane == 31) tail = 0;
    if (seen) stg_y(y + cur_row, sum + tail);  // segment open at the end of this lane
    if (lane == 0) {
        // whatever precedes the tile's first row start continues a row begun in an earlier tile
        if (tile_starts_row) carry_row[t] = -1;
        else { carry_val[t] = g; carry_row[t] = row_start; }
    }
}

// the last (partial or full) tile as CSR rows (csr5_spmv_avx2.h:337-366), one warp per row
template <typename T>
__global__ void __launch_bounds__(kThreads)
csr5_tail_kernel(int m, int tail_start, int tail_nz0, int tile, const int *__restrict__ rowptr,
                 const int *__restrict__ tcol, const T *__restrict__ tval, const T *__restrict__ x,
                 T *__restrict__ y, T *__restrict__ carry_val, int *__restrict__ carry_row)
{
    const uint64_t pl = policy_evict_last();
    const long long w = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long row_l = tail_start + w;
    if (row_l >= m) return;
    const int row = (int)row_l;
    const int rs = rowptr[row], re = rowptr[row + 1];
    const int s = max(rs, tail_nz0);
    T sum = 0;
    for (int j = s + lane; j < re; j += 32) sum = fma_t(tval[j], ldg_x(x + tcol[j], pl), sum);
    sum = group_sum_c<T, 32>(sum);
    if (lane == 0) {
        if (row == tail_start && rs < tail_nz0) {
            carry_val[tile] = sum;
            carry_row[tile] = row;
        } else {
            stg_y(y + row, sum);
            if (row == tail_start) carry_row[tile] = -1;
        }
    }
}

}  // namespace sb

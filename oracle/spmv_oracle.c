/*
 * spmv_oracle.c -- CPU restatement of the reference's SpMV hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed from the
 * product library (libspmv_b200.so); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it, and there only as the checker / the CPU baseline.
 *
 * Parity pin: the reference stores no golden vectors, so this restatement is pinned
 * against the UNMODIFIED reference compiled from /root/reference (oracle/_ref/libmv_l2.so, see
 * oracle/Makefile) -- bit-exact on y for Method_Serial (fp64 and fp32) and bit-exact on every
 * structure (splitters, Yid tables, SELL permutation, CSR5 tile_ptr / tile_desc / offsets) -- by
 * tests/test_oracle_vs_reference.py, and against the fixtures that script-generated outputs of that
 * .so left in tests/golden/ (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Compiled with -ffp-contract=off so that every fused multiply-add below is an explicit fma()/fmaf()
 * exactly where the reference's gcc -O3 -mfma build has one.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * a5  Dot_Product_Avx2_d  (src/src_spmv/inner_spmv.h:232-286)
 * 4 accumulator lanes striped over the row (lane l takes elements l, l+4, ...), each an FMA chain
 * starting from +0.0; horizontal add (l0+l1)+(l2+l3) only when at least one full group of 4 exists;
 * then len%4 trailing elements folded in sequentially (gcc contracts `result += a*b` to an FMA).
 * ---------------------------------------------------------------------------------------------- */
static double row_dot_d(int len, const int *idx, const double *val, const double *x)
{
    double lane[4] = {0.0, 0.0, 0.0, 0.0};
    const int groups = len / 4, rem = len % 4;
    for (int g = 0; g < groups; ++g)
        for (int l = 0; l < 4; ++l)
            lane[l] = fma(val[4 * g + l], x[idx[4 * g + l]], lane[l]);
    double result = 0.0;
    if (groups)
        result = (lane[0] + lane[1]) + (lane[2] + lane[3]);
    for (int j = 0; j < rem; ++j)
        result = fma(val[4 * groups + j], x[idx[4 * groups + j]], result);
    return result;
}

/* a5  Dot_Product_Avx2_s  (src/src_spmv/inner_spmv.h:288-354)
 * 8 lanes; tree ((l0+l4)+(l2+l6)) + ((l1+l5)+(l3+l7)); len%8 trailing elements sequential. */
static float row_dot_s(int len, const int *idx, const float *val, const float *x)
{
    float lane[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int groups = len / 8, rem = len % 8;
    for (int g = 0; g < groups; ++g)
        for (int l = 0; l < 8; ++l)
            lane[l] = fmaf(val[8 * g + l], x[idx[8 * g + l]], lane[l]);
    float result = 0.f;
    if (groups) {
        const float q0 = lane[0] + lane[4], q1 = lane[1] + lane[5];
        const float q2 = lane[2] + lane[6], q3 = lane[3] + lane[7];
        result = (q0 + q2) + (q1 + q3);
    }
    /* Remainder (inner_spmv.h:346-348, `result += *matValPtr++ * x[*colIndPtr++]`).  The pinned build
     * (oracle/Makefile: gcc 13.3 -O3 -mavx2 -mfma) vectorises this loop 4-wide as an IN-ORDER
     * reduction: when len%8 >= 4 the first four products are formed by an unfused vmulps and added
     * to `result` one after another; whatever is left (<4) goes through the scalar epilogue, which
     * gcc contracts to FMAs.  (The fp64 remainder is at most 3 long and stays scalar/FMA.)
     * Observed with tests/test_oracle_vs_reference.py::test_serial_rowlen_sweep. */
    int j = 0;
    if (rem >= 4) {
        float p[4];
        for (int l = 0; l < 4; ++l)
            p[l] = val[8 * groups + l] * x[idx[8 * groups + l]];
        for (int l = 0; l < 4; ++l)
            result = result + p[l];
        j = 4;
    }
    for (; j < rem; ++j)
        result = fmaf(val[8 * groups + j], x[idx[8 * groups + j]], result);
    return result;
}

/* a6  spmv_serial_cpp_d / _s  (src/src_spmv/serial_spmv.c:9-37): y[i] = dot(row i), one thread. */
void oracle_spmv_serial_d(int m, const int *rowptr, const int *col, const double *val,
                          const double *x, double *y)
{
    for (int i = 0; i < m; ++i)
        y[i] = row_dot_d(rowptr[i + 1] - rowptr[i], col + rowptr[i], val + rowptr[i], x);
}

void oracle_spmv_serial_s(int m, const int *rowptr, const int *col, const float *val,
                          const float *x, float *y)
{
    for (int i = 0; i < m; ++i)
        y[i] = row_dot_s(rowptr[i + 1] - rowptr[i], col + rowptr[i], val + rowptr[i], x);
}

/* Same rows under an OpenMP team: the CPU baseline a7 (src/src_spmv/parallel_spmv.c:5-34) when the
 * reference .so itself is not available on the box. */
void oracle_spmv_parallel_d(int m, const int *rowptr, const int *col, const double *val,
                            const double *x, double *y)
{
#pragma omp parallel for
    for (int i = 0; i < m; ++i)
        y[i] = row_dot_d(rowptr[i + 1] - rowptr[i], col + rowptr[i], val + rowptr[i], x);
}

void oracle_spmv_parallel_s(int m, const int *rowptr, const int *col, const float *val,
                            const float *x, float *y)
{
#pragma omp parallel for
    for (int i = 0; i < m; ++i)
        y[i] = row_dot_s(rowptr[i + 1] - rowptr[i], col + rowptr[i], val + rowptr[i], x);
}

/* The scalar CSR-order golden of the sample driver (src/samples/test_spmv.c:204-207); gcc contracts
 * `Y[i] += Val[j]*X[ColIdx[j]]` to an FMA as well, kept explicit here. */
void oracle_spmv_scalar_golden_d(int m, const int *rowptr, const int *col, const double *val,
                                 const double *x, double *y)
{
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            s = fma(val[j], x[col[j]], s);
        y[i] = s;
    }
}

/* Correctly-rounded-for-all-practical-purposes y: products and sums in x87 extended precision (64-bit
 * mantissa) for fp64, in double for fp32.  Not a reference function: it separates the reference's own
 * rounding error on very long rows (its 4/8-lane chains accumulate ~sqrt(len)*eps) from the GPU's when
 * the per-row bound of north_star is checked on rows of 10^3..10^6 non-zeros. */
void oracle_spmv_exact_d(int m, const int *rowptr, const int *col, const double *val,
                         const double *x, double *y)
{
#pragma omp parallel for
    for (int i = 0; i < m; ++i) {
        long double a = 0.0L;
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            a += (long double)val[j] * (long double)x[col[j]];
        y[i] = (double)a;
    }
}

void oracle_spmv_exact_s(int m, const int *rowptr, const int *col, const float *val,
                         const float *x, float *y)
{
#pragma omp parallel for
    for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            a += (double)val[j] * (double)x[col[j]];
        y[i] = (float)a;
    }
}

/* Per-row magnitude sum  S_i = sum_j |a_ij * x_j|  in double: the tolerance of north_star is
 * |y - y_ref| <= 8 * eps * S_i. */
void oracle_row_abs_sum_d(int m, const int *rowptr, const int *col, const double *val,
                          const double *x, double *s)
{
#pragma omp parallel for
    for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            a += fabs(val[j] * x[col[j]]);
        s[i] = a;
    }
}

void oracle_row_abs_sum_s(int m, const int *rowptr, const int *col, const float *val,
                          const float *x, double *s)
{
#pragma omp parallel for
    for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            a += fabs((double)val[j] * (double)x[col[j]]);
        s[i] = a;
    }
}

/* ------------------------------------------------------------------------------------------------
 * a8  binary_search_right_boundary_kernel  (src/src_spmv/parallel_balanced_spmv.c:17-37, template
 * twin csr5_avx2/avx2/utils_avx2.h:23-44): number of entries of row_pointer[0..size) that are <= key.
 * ---------------------------------------------------------------------------------------------- */
int oracle_right_boundary(const int *row_pointer, int key, int size)
{
    int start = 0, stop = size - 1;
    while (stop >= start) {
        const int median = (stop + start) / 2;
        if (key >= row_pointer[median])
            start = median + 1;
        else
            stop = median - 1;
    }
    return start;
}

/* std::lower_bound as wrapped by the reference (src/src_spmv/csr5_spmv.cpp:54-56). */
static int lower_bound_int(const int *a, int n, int key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* a9  init_csrSplitter_balanced2  (src/src_spmv/parallel_balanced2_spmv.c:41-53).
 * splitter[t] = right_boundary(RowPtr, min(t*ceil(nnz/T), nnz), m+1) - 1, t = 0..T. */
void oracle_splitter_balanced2(int T, int nnz, int m, const int *rowptr, int *splitter)
{
    const int stride = (int)(((long long)nnz + T - 1) / T);
    for (int t = 0; t <= T; ++t) {
        long long b = (long long)t * stride;
        if (b > nnz) b = nnz;
        splitter[t] = oracle_right_boundary(rowptr, (int)b, m + 1) - 1;
    }
}

/* a10  the Yid scan of parallel_balanced2_get_handle (parallel_balanced2_spmv.c:72-90):
 * Yid[t] = splitter[t] when partition t owns no whole row and splitter[t] != m, else -1.
 * Returns 1 when every Yid is -1 (the handle is then demoted to Method_Balanced). */
int oracle_balanced2_yid(int T, int m, const int *splitter, int *yid)
{
    int use_balanced = 1;
    for (int t = 0; t < T; ++t) {
        if (splitter[t + 1] - splitter[t] == 0 && splitter[t] != m) {
            yid[t] = splitter[t];
            use_balanced = 0;
        } else {
            yid[t] = -1;
        }
    }
    return use_balanced;
}

/* a12  init_splitter_balancedYid  (src/src_spmv/parallel_balanced_Yid_spmv.c:16-53).
 * Thread i owns nnz [stride*i, min(stride*(i+1), nnz)); l,r = lower_bound(RowPtr, begin/end). */
void oracle_splitter_yid(int T, int nnz, int m, const int *rowptr, int *splitter /*2T*/, int *type,
                         int *brow, int *begin_idx, int *erow, int *end_idx)
{
    const int stride = (int)(((long long)nnz + T - 1) / T);
    for (int i = 1; i <= T; ++i) {
        const long long b64 = (long long)stride * (i - 1);
        long long e64 = b64 + stride;
        if (e64 > nnz) e64 = nnz;
        const int b = (int)b64, e = (int)e64;
        const int l = lower_bound_int(rowptr, m + 1, b);
        const int r = lower_bound_int(rowptr, m + 1, e);
        splitter[2 * i - 2] = l;
        splitter[2 * i - 1] = r - 1;
        brow[i - 1] = l - 1;
        erow[i - 1] = r - 1;
        begin_idx[i - 1] = b;
        end_idx[i - 1] = e;
        type[i - 1] = (l == r) ? 0 : 1;
    }
}

/* ------------------------------------------------------------------------------------------------
 * a13  SELL-C-sigma structure  (src/src_spmv/sell_C_Sigma_spmv.c:132-247).
 * Rows [0, banner) with banner = sigma*floor(m/sigma) are sorted inside each sigma-window by
 * (row length ascending, row index ascending) -- cmp() at :132-138 (note the `return -p`).
 * perm[i] is the original row that lands at sorted position i (Sigma_Block.RowIndex, :104-106).
 * Returns banner.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int len, row; } len_row_t;

static int cmp_len_row(const void *a, const void *b)
{
    const len_row_t *p = (const len_row_t *)a, *q = (const len_row_t *)b;
    if (p->len != q->len) return p->len < q->len ? -1 : 1;
    return p->row - q->row;
}

int oracle_sell_perm(int m, const int *rowptr, int sigma, int *perm)
{
    if (sigma <= 0) return 0;
    const int nwin = m / sigma, banner = nwin * sigma;
#pragma omp parallel for
    for (int w = 0; w < nwin; ++w) {
        len_row_t *tmp = (len_row_t *)malloc(sizeof(len_row_t) * (size_t)sigma);
        for (int i = 0; i < sigma; ++i) {
            const int r = w * sigma + i;
            tmp[i].len = rowptr[r + 1] - rowptr[r];
            tmp[i].row = r;
        }
        qsort(tmp, (size_t)sigma, sizeof(len_row_t), cmp_len_row);
        for (int i = 0; i < sigma; ++i) perm[w * sigma + i] = tmp[i].row;
        free(tmp);
    }
    return banner;
}

/* Per-chunk widths for chunk height C over the sorted rows (spmv_Sigma_Blocks_init, :61-83 and
 * :108-123): width[k] = max row length in chunk k (the `ld` differences), full[k] = min row length. */
void oracle_sell_chunks(const int *rowptr, const int *perm, int banner, int C, int *width, int *full)
{
    const int nchunk = banner / C;
    for (int k = 0; k < nchunk; ++k) {
        int mx = 0, mn = 1000000000;
        for (int i = 0; i < C; ++i) {
            const int r = perm[k * C + i];
            const int len = rowptr[r + 1] - rowptr[r];
            if (len > mx) mx = len;
            if (len < mn) mn = len;
        }
        width[k] = mx;
        full[k] = mn;
    }
}

/* ------------------------------------------------------------------------------------------------
 * a15-a17  CSR5 structure for arbitrary (omega, sigma)  (src/src_spmv/csr5_avx2/anonymouslib_avx2.h
 * :112-242, csr5_avx2/avx2/format_avx2.h:7-345).  The reference fixes omega=4, sigma=16
 * (common_avx2.h:12-13); this restatement takes them as arguments so that the (4,16) instance can be
 * proven bit-equal to the reference and the (32,sigma) instance used as the KAT for the GPU builder.
 * ---------------------------------------------------------------------------------------------- */

/* anonymouslib_avx2.h:124-146: bit widths, packets per lane, number of tiles. */
void oracle_csr5_params(int omega, int sigma, int nnz, int *bit_y_offset, int *bit_scansum_offset,
                        int *num_packet, int *p)
{
    int base = 2, by = 1;
    while (base < omega * sigma) { base *= 2; by++; }
    int bs = 1;
    base = 2;
    while (base < omega) { base *= 2; bs++; }
    *bit_y_offset = by;
    *bit_scansum_offset = bs;
    *num_packet = (by + bs + sigma + 31) / 32;
    *p = (int)(((long long)nnz + (long long)omega * sigma - 1) / ((long long)omega * sigma));
}

/* a16  generate_partition_pointer s1+s2  (format_avx2.h:7-78).  tile_ptr has p+1 entries; bit 31
 * marks a tile whose row span [start..stop] contains an empty row.  NOTE the reference's s2 loop reads
 * row_pointer[stop+1] with stop possibly == m (format_avx2.h:44-45, an out-of-bounds read); the
 * restatement treats that comparison as false, which is what the reference's result is whenever the
 * word behind RowPtr differs from RowPtr[m]. */
void oracle_csr5_tile_ptr(int omega, int sigma, int p, int m, int nnz, const int *rowptr,
                          uint32_t *tile_ptr)
{
    for (int t = 0; t <= p; ++t) {
        long long b = (long long)t * sigma * omega;
        if (b > nnz) b = nnz;
        tile_ptr[t] = (uint32_t)(oracle_right_boundary(rowptr, (int)b, m + 1) - 1);
    }
    for (int t = 0; t < p; ++t) {
        const uint32_t start = tile_ptr[t] & 0x7FFFFFFFu, stop = tile_ptr[t + 1] & 0x7FFFFFFFu;
        if (start == stop) continue;
        int dirty = 0;
        for (uint32_t r = start; r <= stop; ++r) {
            if ((int)r + 1 > m) break; /* guarded out-of-bounds read, see NOTE above */
            if (rowptr[r] == rowptr[r + 1]) { dirty = 1; break; }
        }
        if (dirty) tile_ptr[t] = start | 0x80000000u;
    }
}

/* a17  generate_partition_descriptor s1+s2 and the scan of the offset pointer
 * (format_avx2.h:80-254).  desc has p*omega*num_packet words (pre-zeroed here); off_ptr has p+1
 * entries.  Returns num_offsets. */
int oracle_csr5_tile_desc(int omega, int sigma, int p, int m, int bit_y_offset,
                          int bit_scansum_offset, int num_packet, const int *rowptr,
                          const uint32_t *tile_ptr, uint32_t *desc, int *off_ptr)
{
    const int bit_all = bit_y_offset + bit_scansum_offset;
    memset(desc, 0, sizeof(uint32_t) * (size_t)p * omega * num_packet);
    memset(off_ptr, 0, sizeof(int) * ((size_t)p + 1));
    (void)m;
    /* s1 (format_avx2.h:80-117): one bit per row start that falls inside the tile. */
    for (int t = 0; t < p - 1; ++t) {
        const int row_start = (int)(tile_ptr[t] & 0x7FFFFFFFu);
        const int row_stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
        for (int rid = row_start; rid <= row_stop; ++rid) {
            const int ptr = rowptr[rid];
            const int pid = ptr / (omega * sigma);
            if (pid == t) {
                const int lx = (ptr / sigma) % omega;
                const int glid = ptr % sigma + bit_all;
                const int ly = glid / 32, llid = glid % 32;
                desc[(size_t)pid * omega * num_packet + (size_t)ly * omega + lx] |= 1u << (31 - llid);
            }
        }
    }
    /* s2 (format_avx2.h:119-217): per-lane segment counts -> y_offset, scansum_offset. */
    int *segn = (int *)malloc(sizeof(int) * ((size_t)omega + 1));
    int *present = (int *)malloc(sizeof(int) * ((size_t)omega + 1));
    for (int t = 0; t < p - 1; ++t) {
        const int with_empty = (tile_ptr[t] >> 31) & 1;
        const int row_start = (int)(tile_ptr[t] & 0x7FFFFFFFu);
        const int row_stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
        if (row_start == row_stop) continue;
        uint32_t *d = desc + (size_t)t * omega * num_packet;
        for (int lane = 0; lane < omega; ++lane) {
            int start, stop = 0, pres = !lane, ly = 0;
            uint32_t bitflag = (d[lane] << bit_all) | ((uint32_t)pres << 31);
            start = !((bitflag >> 31) & 1);
            pres |= (bitflag >> 31) & 1;
            for (int i = 1; i < sigma; ++i) {
                if ((!ly && i == 32 - bit_all) || (ly && (i - (32 - bit_all)) % 32 == 0)) {
                    ly++;
                    bitflag = d[(size_t)ly * omega + lane];
                }
                const int norm_i = !ly ? i : i - (32 - bit_all);
                stop += (bitflag >> (31 - norm_i % 32)) & 1;
                pres |= (bitflag >> (31 - norm_i % 32)) & 1;
            }
            int s = stop - start + pres;
            segn[lane] = s > 0 ? s : 0;
            present[lane] = pres;
        }
        segn[omega] = 0;
        present[omega] = 1; /* sentinel: the reference's while loop stops at next1 == omega */
        /* scan_single: exclusive scan over omega+1 entries (utils_avx2.h:69-82) */
        int run = 0;
        for (int i = 0; i <= omega; ++i) { const int v = segn[i]; segn[i] = run; run += v; }
        if (with_empty) {
            off_ptr[t] = segn[omega];
            off_ptr[p] += segn[omega];
        }
        for (int lane = 0; lane < omega; ++lane) {
            int y_offset = segn[lane], scansum = 0, next1 = lane + 1;
            if (present[lane])
                while (next1 < omega && !present[next1]) { scansum++; next1++; }
            y_offset = lane ? y_offset - 1 : 0;
            d[lane] |= (uint32_t)y_offset << (32 - bit_y_offset);
            d[lane] |= (uint32_t)scansum << (32 - bit_all);
        }
    }
    free(segn);
    free(present);
    /* format_avx2.h:243-246: exclusive scan of the per-tile offset counts when any exist. */
    if (off_ptr[p]) {
        int run = 0;
        for (int i = 0; i <= p; ++i) { const int v = off_ptr[i]; off_ptr[i] = run; run += v; }
    }
    return off_ptr[p];
}

/* a17  generate_partition_descriptor_offset_kernel (format_avx2.h:256-323): for tiles with empty
 * rows, the true y index (relative to row_start+1) of every segment that starts in the tile. */
void oracle_csr5_desc_offset(int omega, int sigma, int p, int bit_y_offset, int bit_scansum_offset,
                             int num_packet, const int *rowptr, const uint32_t *tile_ptr,
                             const uint32_t *desc, const int *off_ptr, int *off)
{
    const int bit_all = bit_y_offset + bit_scansum_offset, bit_bitflag = 32 - bit_all;
    for (int t = 0; t < p - 1; ++t) {
        if (!((tile_ptr[t] >> 31) & 1)) continue;
        const int row_start = (int)(tile_ptr[t] & 0x7FFFFFFFu);
        const int row_stop = (int)(tile_ptr[t + 1] & 0x7FFFFFFFu);
        const uint32_t *d = desc + (size_t)t * omega * num_packet;
        for (int lane = 0; lane < omega; ++lane) {
            int ly = 0;
            uint32_t w = d[lane];
            int y_offset = (int)(w >> (32 - bit_y_offset));
            w <<= bit_all;
            if (!lane) w |= 0x80000000u;
            if (((w >> 31) & 1) && lane) {
                const int idx = t * omega * sigma + lane * sigma;
                off[off_ptr[t] + y_offset] =
                    oracle_right_boundary(rowptr + row_start + 1, idx, row_stop - row_start) - 1;
                y_offset++;
            }
            for (int i = 1; i < sigma; ++i) {
                if ((!ly && i == bit_bitflag) || (ly && !(31 & (i - bit_bitflag)))) {
                    ly++;
                    w = d[(size_t)ly * omega + lane];
                }
                const int norm_i = 31 & (!ly ? i : i - bit_bitflag);
                if ((w >> (31 - norm_i)) & 1) {
                    const int idx = t * omega * sigma + lane * sigma + i;
                    off[off_ptr[t] + y_offset] =
                        oracle_right_boundary(rowptr + row_start + 1, idx, row_stop - row_start) - 1;
                    y_offset++;
                }
            }
        }
    }
}

/* a18  aosoa_transpose R2C (format_avx2.h:347-425) into a COPY (the reference works in place):
 * inside every full tile that is not a fast-track tile (tile_ptr[t] != tile_ptr[t+1], raw words
 * compared, :366), element x*sigma + y moves to y*omega + x. */
void oracle_csr5_transpose_i32(int omega, int sigma, int nnz, const uint32_t *tile_ptr,
                               const int *src, int *dst)
{
    const int tile = omega * sigma;
    const int full = (int)(((long long)nnz + tile - 1) / tile) - 1;
    memcpy(dst, src, sizeof(int) * (size_t)nnz);
    for (int t = 0; t < full; ++t) {
        if (tile_ptr[t] == tile_ptr[t + 1]) continue;
        for (int idx = 0; idx < tile; ++idx) {
            const int y = idx % sigma, x = idx / sigma;
            dst[(size_t)t * tile + (size_t)y * omega + x] = src[(size_t)t * tile + idx];
        }
    }
}

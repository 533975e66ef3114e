// spmv_b200.cu -- the C-ABI of libspmv_b200.so: the four drop-in entry points of include/spmv.h, the
// exported name tables, and the extensions of include/spmv_b200.h.  Handle construction uploads (or
// adopts) the CSR arrays once and builds the device layout of the requested SPMV_METHODS value;
// spmv() launches the matching sm_100a kernel family.  There is no CPU compute path in this file.
#include <cstdarg>
#include <atomic>
#include <mutex>
#include <string>
#include <map>
#include <cmath>
#include <vector>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "csr_kernels.cuh"
#include "tile_kernels.cuh"
#include "sell.cuh"
#include "csr5.cuh"
#include "long_rows.cuh"
#include "band_seg.cuh"

namespace sb {

// ------------------------------------------------------------------------------------------------
// error latch, launch counter, options
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[spmv_b200] %s\n", g_err);
}

bool cuda_ok(cudaError_t e, const char *what, const char *file, int line)
{
    if (e == cudaSuccess) return true;
    set_error("CUDA error %s (%s) at %s:%d in `%s`", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
    return false;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

struct Options {
    std::mutex mu;
    std::map<std::string, long long> v{{"sell_sigma", 256}, {"csr5_sigma", 0}, {"block_nnz", 512},
                                       {"tile_items", 8},   {"tpr", 0},         {"x_bands", 0},
                                       {"force_merge", 0}, {"sell_cap", 1024},
                                       {"long_thr", 0},    {"pipeline", 1},   {"sell_variant", -1},
                                       {"seg_bands", 0},    {"seg_prefetch", 1}, {"seg_ctas", 2},  {"auto", 0},     {"reorder", 0},    {"row_bins", 1},
                                       {"pin_host", 1}};
    std::map<std::string, bool> user_set;
};
static Options &options()
{
    static Options o;
    return o;
}
static long long opt(const char *key)
{
    Options &o = options();
    std::lock_guard<std::mutex> g(o.mu);
    auto it = o.v.find(key);
    if (it == o.v.end()) return -1;
    if (!o.user_set[key]) {  // environment: SPMV_B200_<KEY>
        std::string env = "SPMV_B200_";
        for (const char *p = key; *p; ++p) env.push_back((char)toupper(*p));
        if (const char *e = getenv(env.c_str())) return atoll(e);
    }
    return it->second;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
static bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

template <typename U>
static bool dmalloc(U **p, size_t count)
{
    *p = nullptr;
    return SB_CUDA(cudaMalloc((void **)p, (count ? count : 1) * sizeof(U)));
}

static void dfree(void *p) { if (p) cudaFree(p); }

static inline int blocks_for(long long threads) { return (int)((threads + kThreads - 1) / kThreads); }

struct DeviceGuard {  // library code is device-agnostic: run on the handle's device, then restore
    int prev = -1, want;
    explicit DeviceGuard(int dev) : want(dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != want) cudaSetDevice(want);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
};

static bool read_flag(int *d_flag, cudaStream_t s, int *out)
{
    SB_TRY(cudaMemcpyAsync(out, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    SB_TRY(cudaStreamSynchronize(s));
    return true;
}

template <typename V>
static bool exclusive_scan(const V *in, V *out, int count, cudaStream_t s)
{
    size_t bytes = 0;
    SB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, s));
    void *tmp = nullptr;
    SB_TRY(cudaMalloc(&tmp, bytes ? bytes : 1));
    const bool ok = SB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, count, s)) &&
                    SB_CUDA(cudaStreamSynchronize(s));
    cudaFree(tmp);
    return ok;
}

// ------------------------------------------------------------------------------------------------
// layout builders (handle construction, a2)
// ------------------------------------------------------------------------------------------------
static void free_layouts(DeviceState *st)
{
    dfree(st->v_rowptr); st->v_rowptr = nullptr;
    dfree(st->v_col); st->v_col = nullptr;
    dfree(st->v_val); st->v_val = nullptr;
    dfree(st->v_y); st->v_y = nullptr;
    dfree(st->splitter); st->splitter = nullptr;
    dfree(st->ref_splitter); st->ref_splitter = nullptr;
    dfree(st->tile_rows); st->tile_rows = nullptr;
    dfree(st->merge_coords); st->merge_coords = nullptr;
    dfree(st->carry_val); st->carry_val = nullptr;
    dfree(st->carry_row); st->carry_row = nullptr;
    dfree(st->carry2_val); st->carry2_val = nullptr;
    dfree(st->carry2_row); st->carry2_row = nullptr;
    dfree(st->sell_perm); st->sell_perm = nullptr;
    dfree(st->sell_width); st->sell_width = nullptr;
    dfree(st->sell_full); st->sell_full = nullptr;
    dfree(st->sell_col); st->sell_col = nullptr;
    dfree(st->sell_slice_ptr); st->sell_slice_ptr = nullptr;
    dfree(st->sell_val); st->sell_val = nullptr;
    dfree(st->c5_tile_ptr); st->c5_tile_ptr = nullptr;
    dfree(st->c5_tile_desc); st->c5_tile_desc = nullptr;
    dfree(st->c5_off_ptr); st->c5_off_ptr = nullptr;
    dfree(st->c5_off); st->c5_off = nullptr;
    dfree(st->c5_col); st->c5_col = nullptr;
    dfree(st->c5_val); st->c5_val = nullptr;
    dfree(st->bin_list); st->bin_list = nullptr;
    dfree(st->seg_col); st->seg_col = nullptr;
    dfree(st->seg_ent_val); st->seg_ent_val = nullptr;
    dfree(st->seg_mask); st->seg_mask = nullptr;
    dfree(st->seg_tile_ent); st->seg_tile_ent = nullptr;
    dfree(st->seg_tile_seg0); st->seg_tile_seg0 = nullptr;
    dfree(st->seg_gbase); st->seg_gbase = nullptr;
    dfree(st->seg_ticket); st->seg_ticket = nullptr;
    dfree(st->seg_sums); st->seg_sums = nullptr;
    dfree(st->lr_row); st->lr_row = nullptr;
    dfree(st->lr_start); st->lr_start = nullptr;
    dfree(st->lr_seg_ptr); st->lr_seg_ptr = nullptr;
    dfree(st->lr_seg_row); st->lr_seg_row = nullptr;
    dfree(st->lr_partial); st->lr_partial = nullptr;
    dfree(st->x_stage); st->x_stage = nullptr;
    dfree(st->y_stage); st->y_stage = nullptr;
}

static void unpin(DeviceState::PinSlot &slot)
{
    if (slot.registered && cudaHostUnregister(const_cast<void *>(slot.ptr)) != cudaSuccess) cudaGetLastError();
    slot = DeviceState::PinSlot();
}

// see DeviceState::pin
static void note_host_buffer(DeviceState *st, int which, const void *p, size_t bytes)
{
    DeviceState::PinSlot &slot = st->pin[which];
    if (slot.ptr != p || slot.bytes != bytes) {
        unpin(slot);
        slot.ptr = p;
        slot.bytes = bytes;
        slot.seen = 1;
        return;
    }
    if (slot.registered || slot.seen < 0 || bytes < (1u << 20)) return;
    if (++slot.seen < 2) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); slot.seen = -1; return; }
    if (a.type != cudaMemoryTypeUnregistered) { slot.seen = -1; return; }  // already pinned by the caller
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) slot.registered = true;
    else { cudaGetLastError(); slot.seen = -1; }  // e.g. registered through another handle: leave it alone
}

static void free_state(DeviceState *st)
{
    if (!st) return;
    DeviceGuard g(st->device);
    unpin(st->pin[0]);
    unpin(st->pin[1]);
    free_layouts(st);
    for (int i = 0; i < kMaxPieces; ++i) if (st->ev_in[i]) cudaEventDestroy(st->ev_in[i]);
    for (int i = 0; i < kPipeChunks; ++i) if (st->ev_out[i]) cudaEventDestroy(st->ev_out[i]);
    if (st->ev_start) cudaEventDestroy(st->ev_start);
    if (st->ev_x) cudaEventDestroy(st->ev_x);
    if (st->s_in) cudaStreamDestroy(st->s_in);
    if (st->s_out) cudaStreamDestroy(st->s_out);
    if (st->owns_csr) { dfree(st->rowptr); dfree(st->col); dfree(st->val); }
    st->magic = 0;
    delete st;
}

// Device facts the layout decisions need.  (The persisting-L2 carve-out, cudaLimitMaxL2FetchGranularity and an
// access-policy window over x were options in round 1; none of them moved any measured configuration -- C2 then, the
// C5 shard in round 2 -- and they are gone.)
static void apply_device_limits(DeviceState *st)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, st->device) != cudaSuccess) { cudaGetLastError(); return; }
    st->dev_sms = prop.multiProcessorCount;
    st->dev_l2 = prop.l2CacheSize;
    st->pin_host = opt("pin_host") != 0;
}

static int pick_tpr(long long nnz, int m)
{
    const long long forced = opt("tpr");
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16 || forced == 32) return (int)forced;
    const double mean = m > 0 ? (double)nnz / m : 0.0;
    int tpr = 1;
    // measured on B200 (scripts/sweep.py): ~7 entries per lane is the sweet spot -- fewer lanes per row
    // means more independent x gathers in flight per thread (mean 5 -> 1, 10.7 -> 2, 27 -> 4, 32 -> 8)
    while (tpr < 32 && 7.0 * tpr < mean) tpr <<= 1;
    return tpr;
}

// Decide the number of column bands and, when > 1, build the band-major copy and make it the active view.
template <typename T>
static bool build_band_major(DeviceState *st)
{
    st->a_m = st->m; st->a_rowptr = st->rowptr; st->a_col = st->col; st->a_val = st->val;
    if (st->m == 0 || st->nnz == 0) return true;
    long long bands = opt("x_bands");  // 0 = automatic, 1 = off, >= 2 forced
    const double xbytes = (double)st->n * st->vsize;
    const double usable = st->dev_l2 > 0 ? 0.5 * (double)st->dev_l2 : 63.0 * 1048576.0;  // random-access reach
    {
        // locality probe: the share of entries further than 1/8 of the L2 reach from the (scaled) diagonal.
        // > 25 % = gather-dominated (bound by L2 requests), else diagonal-local (bound by DRAM).  It decides
        // whether banding can pay (below) and which SELL kernel flavour runs (build_sell).
        unsigned long long *far = nullptr, h_far = 0;
        bool probed = dmalloc(&far, 1) && SB_CUDA(cudaMemsetAsync(far, 0, sizeof(*far), st->stream));
        if (probed) {
            const int halfwidth = (int)(0.125 * usable / st->vsize);
            band_locality_kernel<<<blocks_for(st->m), kThreads, 0, st->stream>>>(
                st->m, (double)st->n / st->m, halfwidth, st->rowptr, st->col, far);
            probed = SB_CUDA(cudaMemcpyAsync(&h_far, far, sizeof(h_far), cudaMemcpyDeviceToHost, st->stream)) &&
                     SB_CUDA(cudaStreamSynchronize(st->stream));
        }
        dfree(far);
        if (!probed) return false;  // a 8-byte allocation or a trivial kernel failed: the device is unusable
        st->far_fraction = (double)h_far / (double)st->nnz;
    }
    if (bands == 0) {
        bands = 1;
        // x up to ~1.2x the reach still gathers at >= 85 % of the L2 rate (profiles/r01_gather_probe.txt) and
        // banding costs 15-20 % (virtual row pointers, partial y): R-MAT s24 fp32 (x = 64 MiB) measured 7-15 %
        // FASTER unbanded, uniform fp64 (x = 128 MiB) 1.9x faster banded.  Banding only pays when the accesses
        // are NOT already diagonal-local.
        if (xbytes > 1.2 * usable && st->far_fraction > 0.25) {
            const long long k = (long long)ceil(xbytes / (0.75 * usable));
            if (k <= kMaxBands && (double)st->nnz / ((double)k * st->m) >= 4.0) bands = k;
            else if (opt("seg_bands") == 0) {
                // hyper-sparse bands (< 4 entries per virtual row) would cost more in row pointers than they save:
                // band segments (band_seg.cuh).  More bands = a smaller slice of x = fewer L2 misses of the gathers,
                // but more and shorter segments for the merge pass.  Measured on the C5 shard (x = 2 GiB, 16 per
                // row; scripts/c5_sweep.sh): 32 bands 5.24 ms, 40: 4.94, 44: 4.88, 48: 4.78, 52: 4.92, 56: 4.96,
                // 64: 5.19 -- a slice of about 2/3 of the reach
                const long long kc = (long long)ceil(xbytes / (0.68 * usable));
                st->coo_bands = (int)(kc < 2 ? 2 : (kc > kSegMaxBands ? kSegMaxBands : kc));
            }
        }
    }
    if (opt("seg_bands") >= 2) { st->coo_bands = (int)(opt("seg_bands") > kSegMaxBands ? kSegMaxBands : opt("seg_bands")); bands = 1; }
    if (bands <= 1) return true;
    if (bands > kMaxBands) bands = kMaxBands;
    if ((long long)st->m * bands > 0x7fffffffLL - 8192) return true;
    const int K = (int)bands, m = st->m;
    const int band_cols = (int)(((long long)st->n + K - 1) / K);
    const size_t vm = (size_t)m * K;
    int *counts = nullptr;
    // The band-major copy is a speed-up only: when it cannot be allocated (or built) the handle keeps running the
    // plain CSR view it already holds -- never fail a handle over an optional layout.
    auto give_up = [&]() {
        cudaGetLastError();
        dfree(counts);
        dfree(st->v_rowptr); st->v_rowptr = nullptr;
        dfree(st->v_col); st->v_col = nullptr;
        dfree(st->v_val); st->v_val = nullptr;
        dfree(st->v_y); st->v_y = nullptr;
        spmv_b200_clear_error();
        st->layout_fallbacks++;
        return true;
    };
    if (!dmalloc(&counts, vm + 1) || !dmalloc(&st->v_rowptr, vm + 1 + 8) || !dmalloc(&st->v_col, (size_t)st->nnz + 8) ||
        !SB_CUDA(cudaMalloc(&st->v_val, ((size_t)st->nnz + 8) * sizeof(T))) || !SB_CUDA(cudaMalloc(&st->v_y, vm * sizeof(T))))
        return give_up();
    bool ok = SB_CUDA(cudaMemsetAsync(counts + vm, 0, sizeof(int), st->stream));
    if (ok) {
        band_count_kernel<<<blocks_for(m), kThreads, 0, st->stream>>>(m, K, band_cols, st->rowptr, st->col, counts);
        ok = SB_CUDA(cudaGetLastError()) && exclusive_scan(counts, st->v_rowptr, (int)vm + 1, st->stream);
    }
    if (ok) {
        band_scatter_kernel<T><<<blocks_for(m), kThreads, 0, st->stream>>>(m, K, band_cols, st->rowptr, st->col, (const T *)st->val,
                                                                           st->v_rowptr, st->v_col, (T *)st->v_val);
        ok = SB_CUDA(cudaGetLastError()) && SB_CUDA(cudaStreamSynchronize(st->stream));
    }
    if (!ok) return give_up();
    dfree(counts);
    st->x_bands = K;
    st->band_cols = band_cols;
    st->a_m = (int)vm; st->a_rowptr = st->v_rowptr; st->a_col = st->v_col; st->a_val = st->v_val;
    return true;
}

// a9 on the device for `parts` partitions; *starved = some partition owns no whole row (a10)
static bool build_splitter(DeviceState *st, int m, const int *rowptr, int parts, int **out, int *starved)
{
    if (!dmalloc(out, (size_t)parts + 1)) return false;
    splitter_kernel<<<blocks_for(parts + 1), kThreads, 0, st->stream>>>(parts, st->nnz, m, rowptr, *out);
    SB_TRY(cudaGetLastError());
    if (starved) {
        int *flag = nullptr;
        if (!dmalloc(&flag, 1)) return false;
        SB_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st->stream));
        splitter_starved_kernel<<<blocks_for(parts), kThreads, 0, st->stream>>>(parts, m, *out, flag);
        const bool ok = read_flag(flag, st->stream, starved);
        dfree(flag);
        if (!ok) return false;
    }
    return true;
}

// carry arrays of the tile kernels: one (row, value) per tile + the level-2 list of the two-level fix-up
static bool alloc_carries(DeviceState *st, int tiles)
{
    const size_t t = tiles > 0 ? (size_t)tiles : 1;
    const size_t g2 = 2 * ((t + kCarryGroup - 1) / kCarryGroup);
    return SB_CUDA(cudaMalloc(&st->carry_val, t * st->vsize)) && dmalloc(&st->carry_row, t) &&
           SB_CUDA(cudaMalloc(&st->carry2_val, g2 * st->vsize)) && dmalloc(&st->carry2_row, g2);
}

// y[row] += the carries of `tiles` consecutive tiles, in tile order; long lists in two levels
template <typename T>
static void launch_carry_fixup(DeviceState *st, cudaStream_t s, int tiles, const int *carry_row, const T *carry_val, T *y)
{
    if (tiles <= 0) return;
    if (tiles < 1024) {
        carry_fixup_kernel<T><<<blocks_for(tiles), kThreads, 0, s>>>(tiles, carry_row, carry_val, y);
        count_launch();
        return;
    }
    const int groups = (tiles + kCarryGroup - 1) / kCarryGroup;
    carry_group_kernel<T><<<blocks_for((long long)groups * 32), kThreads, 0, s>>>(tiles, carry_row, carry_val, y, st->carry2_row, (T *)st->carry2_val);
    carry_fixup_kernel<T><<<blocks_for(2 * groups), kThreads, 0, s>>>(2 * groups, st->carry2_row, (const T *)st->carry2_val, y);
    count_launch(2);
}

// Segment list for the entries the main kernel does not cover: covered[r] leading entries of row r are done
// by the main kernel, the rest by long_seg_kernel / long_final_kernel (long_rows.cuh).  Frees `covered`.
static bool build_long_rows(DeviceState *st, int *covered, bool accumulate)
{
    const int m = st->a_m;
    int *cnt_r = nullptr, *cnt_s = nullptr, *scan_r = nullptr, *scan_s = nullptr;
    bool ok = dmalloc(&cnt_r, (size_t)m + 1) && dmalloc(&cnt_s, (size_t)m + 1) && dmalloc(&scan_r, (size_t)m + 1) &&
              dmalloc(&scan_s, (size_t)m + 1);
    if (ok) {
        long_count_kernel<<<blocks_for((long long)m + 1), kThreads, 0, st->stream>>>(m, st->a_rowptr, covered, cnt_r, cnt_s);
        ok = SB_CUDA(cudaGetLastError()) && exclusive_scan(cnt_r, scan_r, m + 1, st->stream) &&
             exclusive_scan(cnt_s, scan_s, m + 1, st->stream);
    }
    int totals[2] = {0, 0};
    if (ok) ok = SB_CUDA(cudaMemcpy(&totals[0], scan_r + m, sizeof(int), cudaMemcpyDeviceToHost)) &&
                 SB_CUDA(cudaMemcpy(&totals[1], scan_s + m, sizeof(int), cudaMemcpyDeviceToHost));
    st->lr_rows = totals[0];
    st->lr_segs = totals[1];
    st->lr_accumulate = accumulate;
    if (ok && st->lr_rows > 0) {
        ok = dmalloc(&st->lr_row, (size_t)st->lr_rows) && dmalloc(&st->lr_start, (size_t)st->lr_rows) &&
             dmalloc(&st->lr_seg_ptr, (size_t)st->lr_rows + 1) && dmalloc(&st->lr_seg_row, (size_t)st->lr_segs) &&
             SB_CUDA(cudaMalloc(&st->lr_partial, (size_t)st->lr_segs * st->vsize));
        if (ok) {
            long_fill_kernel<<<blocks_for((long long)m + 1), kThreads, 0, st->stream>>>(
                m, st->a_rowptr, covered, scan_r, scan_s, st->lr_row, st->lr_start, st->lr_seg_ptr, st->lr_seg_row);
            ok = SB_CUDA(cudaGetLastError()) && SB_CUDA(cudaStreamSynchronize(st->stream));
        }
    }
    dfree(cnt_r); dfree(cnt_s); dfree(scan_r); dfree(scan_s); dfree(covered);
    return ok;
}

// Method_Parallel: rows longer than ~256 entries per lane would keep one lane group busy long after the rest
// of the grid has drained; they go to the long-row path as a whole.
static void free_long_rows(DeviceState *st)
{
    dfree(st->lr_row); st->lr_row = nullptr;
    dfree(st->lr_start); st->lr_start = nullptr;
    dfree(st->lr_seg_ptr); st->lr_seg_ptr = nullptr;
    dfree(st->lr_seg_row); st->lr_seg_row = nullptr;
    dfree(st->lr_partial); st->lr_partial = nullptr;
    st->lr_rows = st->lr_segs = 0;
}

static bool build_long_rows_at(DeviceState *st, int thr)
{
    st->long_thr = thr;
    int *covered = nullptr;
    if (!dmalloc(&covered, (size_t)st->a_m + 1)) return false;
    long_cover_threshold_kernel<<<blocks_for(st->a_m), kThreads, 0, st->stream>>>(st->a_m, 0, thr, st->a_rowptr, covered);
    SB_TRY(cudaGetLastError());
    return build_long_rows(st, covered, /*accumulate=*/false);
}

static bool build_long_rows_threshold(DeviceState *st)
{
    long long thr = opt("long_thr");  // 0 = automatic, < 0 = off
    if (thr < 0) { st->long_thr = 0x7fffffff; return true; }
    const bool automatic = thr == 0;
    if (automatic) { thr = 256LL * st->tpr; if (thr < 512) thr = 512; if (thr > 4096) thr = 4096; }
    if (!build_long_rows_at(st, (int)thr)) return false;
    // A short-row matrix WITH hub rows (power-law graphs): one lane-group size for all rows leaves most lanes
    // waiting for the longest row of their warp.  Bin the rows by length class instead -- stable radix sort of the
    // row ids by bin -- and give every bin its own launch; rows beyond 128 entries go to the long-row path.
    // (C3: 1.26 ms with one lane-group size, 1.20 ms re-split at 128 entries, see DESIGN.md for the binned figure)
    if (automatic && st->lr_rows > 0 && st->tpr <= 4 && opt("tpr") == 0 && opt("row_bins") != 0) {
        free_long_rows(st);
        if (!build_long_rows_at(st, 128)) return false;
        const int m = st->a_m;
        unsigned char *key_in = nullptr, *key_out = nullptr;
        int *ids = nullptr, *d_ptr = nullptr;
        void *tmp = nullptr;
        size_t tmp_bytes = 0;
        bool ok = dmalloc(&key_in, (size_t)m) && dmalloc(&key_out, (size_t)m) && dmalloc(&ids, (size_t)m) &&
                  dmalloc(&st->bin_list, (size_t)m) && dmalloc(&d_ptr, 5);
        if (ok) {
            row_bin_kernel<<<blocks_for(m), kThreads, 0, st->stream>>>(m, st->a_rowptr, key_in, ids);
            ok = SB_CUDA(cudaGetLastError()) &&
                 SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out, ids, st->bin_list, m, 0, 2, st->stream)) &&
                 SB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1)) &&
                 SB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, ids, st->bin_list, m, 0, 2, st->stream));
        }
        int h_ptr[5] = {0, 0, 0, 0, 0};
        if (ok) {
            sorted_key_ptr_kernel<<<1, kThreads, 0, st->stream>>>(m, 4, key_out, d_ptr);  // first sorted row of every bin
            ok = SB_CUDA(cudaGetLastError()) &&
                 SB_CUDA(cudaMemcpyAsync(h_ptr, d_ptr, sizeof(h_ptr), cudaMemcpyDeviceToHost, st->stream)) &&
                 SB_CUDA(cudaStreamSynchronize(st->stream));
        }
        dfree(key_in); dfree(key_out); dfree(ids); dfree(d_ptr); dfree(tmp);
        if (!ok) return false;
        for (int b = 0; b < 4; ++b) st->bin_ptr[b] = h_ptr[b];  // bin b = list[bin_ptr[b], bin_ptr[b+1]); bin 3 is unused
        st->binned = true;
    }
    return true;
}

// Host-pointer pipeline of the CSR-vector kernel (option "pipeline", default on): which prefix of x each of
// the kPipeChunks row chunks needs (unbanded view; a band needs exactly its own slice of x).
static bool build_pipeline(DeviceState *st)
{
    st->pipeline = false;
    if (opt("pipeline") == 0 || st->lr_rows > 0 || st->m < (1 << 16)) return true;
    if (st->x_bands > 1) { st->pipeline = st->x_bands <= kMaxPieces; return true; }
    int *d = nullptr;
    if (!dmalloc(&d, kPipeChunks)) return false;
    SB_TRY(cudaMemsetAsync(d, 0, kPipeChunks * sizeof(int), st->stream));
    chunk_xmax_kernel<<<blocks_for(st->m), kThreads, 0, st->stream>>>(st->m, kPipeChunks, st->rowptr, st->col, d);
    const bool ok = SB_CUDA(cudaGetLastError()) &&
                    SB_CUDA(cudaMemcpyAsync(st->chunk_xmax, d, kPipeChunks * sizeof(int), cudaMemcpyDeviceToHost, st->stream)) &&
                    SB_CUDA(cudaStreamSynchronize(st->stream));
    dfree(d);
    st->pipeline = ok;
    return ok;
}

static bool build_tiles(DeviceState *st, bool merge)
{
    int ipt = (int)opt("tile_items");
    if (ipt != 4 && ipt != 8) ipt = 8;
    st->tile_items = ipt;
    const long long per_tile = (long long)kThreads * ipt;
    const long long total = merge ? (long long)st->a_m + st->nnz : (long long)st->nnz;
    st->tiles = (int)((total + per_tile - 1) / per_tile);
    if (st->tiles < 1) st->tiles = 1;
    if (merge) {
        if (!dmalloc(&st->merge_coords, (size_t)st->tiles + 1)) return false;
        merge_coords_kernel<<<blocks_for(st->tiles + 1), kThreads, 0, st->stream>>>(
            st->tiles, (int)per_tile, st->nnz, st->a_m, st->a_rowptr, st->merge_coords);
    } else {
        if (!dmalloc(&st->tile_rows, (size_t)st->tiles + 1)) return false;
        tile_rows_kernel<<<blocks_for(st->tiles + 1), kThreads, 0, st->stream>>>(
            st->tiles, (int)per_tile, st->nnz, st->a_m, st->a_rowptr, st->tile_rows);
    }
    SB_TRY(cudaGetLastError());
    if (!alloc_carries(st, st->tiles)) return false;
    st->kernel = merge ? SPMV_B200_KERNEL_MERGE_PATH : SPMV_B200_KERNEL_NNZ_SPLIT;
    return true;
}

template <typename T>
static bool build_sell(DeviceState *st)
{
    long long sigma = opt("sell_sigma");
    if (sigma < kSellC) sigma = kSellC;
    if (sigma > kSellMaxSigma) sigma = kSellMaxSigma;
    sigma = (sigma / kSellC) * kSellC;
    st->sigma = (int)sigma;
    const int windows = st->a_m / st->sigma;
    st->banner = windows * st->sigma;  // reference sell_C_Sigma_spmv.c:148-156
    st->slices = st->banner / kSellC;
    st->tpr = pick_tpr(st->nnz, st->a_m);  // for the CSR tail rows [banner, m)
    st->kernel = SPMV_B200_KERNEL_SELL;
    // HBM-bound (diagonal-local) matrices want every warp slot filled: 4 columns per step in 32 registers
    // (C4: 0.86 -> 0.98 of the measured peak); gather-bound ones run best with 8 columns per step (C2: 0.43 vs 0.39)
    st->sell_variant = (int)opt("sell_variant");
    if (st->sell_variant != 0 && st->sell_variant != 2) st->sell_variant = (st->far_fraction > 0.25) ? 0 : 2;
    // covered[r]: entries of row r stored in its slice (sell_width_kernel); tail rows [banner, m) stay CSR,
    // except hub rows, which csr_tail_kernel zeroes and the long-row path then adds as a whole
    long long cap_opt = opt("sell_cap");  // 0 = off (the reference's widths), else the widest slice allowed
    const int cap = cap_opt <= 0 ? 0 : (int)(cap_opt < 32 ? 32 : (cap_opt > (1 << 20) ? (1 << 20) : cap_opt));
    int *covered = nullptr;
    if (cap) {
        if (!dmalloc(&covered, (size_t)st->a_m + 1)) return false;
        st->long_thr = 4096;
        long_cover_threshold_kernel<<<blocks_for(st->a_m), kThreads, 0, st->stream>>>(st->a_m, st->banner, st->long_thr, st->a_rowptr, covered);
        SB_TRY(cudaGetLastError());
    }
    if (st->banner == 0) return cap ? build_long_rows(st, covered, /*accumulate=*/true) : true;
    int pow2 = 1;
    while (pow2 < st->sigma) pow2 <<= 1;
    if (!dmalloc(&st->sell_perm, (size_t)st->banner)) return false;
    sell_sort_kernel<<<windows, kThreads, (size_t)pow2 * sizeof(unsigned long long), st->stream>>>(
        st->sigma, pow2, st->a_rowptr, st->sell_perm);
    SB_TRY(cudaGetLastError());
    long long *count = nullptr;
    if (!dmalloc(&st->sell_width, (size_t)st->slices) || !dmalloc(&st->sell_full, (size_t)st->slices) ||
        !dmalloc(&count, (size_t)st->slices + 1) || !dmalloc(&st->sell_slice_ptr, (size_t)st->slices + 1))
        return false;
    SB_TRY(cudaMemsetAsync(count, 0, ((size_t)st->slices + 1) * sizeof(long long), st->stream));
    sell_width_kernel<<<blocks_for((long long)st->slices * 32), kThreads, 0, st->stream>>>(
        st->slices, cap, st->a_rowptr, st->sell_perm, st->sell_width, st->sell_full, count, covered);
    SB_TRY(cudaGetLastError());
    if (cap && !build_long_rows(st, covered, /*accumulate=*/true)) return false;
    const bool ok = exclusive_scan(count, st->sell_slice_ptr, st->slices + 1, st->stream);
    dfree(count);
    if (!ok) return false;
    SB_TRY(cudaMemcpy(&st->padded, st->sell_slice_ptr + st->slices, sizeof(long long), cudaMemcpyDeviceToHost));
    if (!dmalloc(&st->sell_col, (size_t)st->padded)) return false;
    if (!SB_CUDA(cudaMalloc(&st->sell_val, (size_t)(st->padded ? st->padded : 1) * sizeof(T)))) return false;
    sell_fill_kernel<T><<<blocks_for((long long)st->slices * 32), kThreads, 0, st->stream>>>(
        st->slices, st->a_rowptr, st->a_col, (const T *)st->a_val, st->sell_perm, st->sell_slice_ptr, st->sell_col,
        (T *)st->sell_val);
    SB_TRY(cudaGetLastError());
    return true;
}

template <typename T>
static bool build_csr5(DeviceState *st)
{
    int sigma = (int)opt("csr5_sigma");  // 0 = automatic: 16, or 8 on small matrices (twice the tiles to fill the GPU)
    if (sigma != 4 && sigma != 8 && sigma != 16) sigma = st->nnz < (1 << 25) ? 8 : 16;
    st->c5_sigma = sigma;
    // anonymouslib_avx2.h:124-146 at omega = 32
    int base = 2, by = 1;
    while (base < kC5Omega * sigma) { base *= 2; by++; }
    int bs = 1;
    base = 2;
    while (base < kC5Omega) { base *= 2; bs++; }
    st->c5_bit_y = by;
    st->c5_bit_ss = bs;  // by + bs + sigma <= 32 for sigma <= 16: one descriptor word per lane
    const int tile_nnz = kC5Omega * sigma;
    const int p = (int)(((long long)st->nnz + tile_nnz - 1) / tile_nnz);
    st->c5_p = p;
    st->kernel = SPMV_B200_KERNEL_CSR5;
    st->tiles = p;
    if (p == 0) return true;
    uint32_t *raw = nullptr;
    int *off_cnt = nullptr;
    if (!dmalloc(&raw, (size_t)p + 1) || !dmalloc(&st->c5_tile_ptr, (size_t)p + 1) ||
        !dmalloc(&st->c5_tile_desc, (size_t)p * kC5Omega) || !dmalloc(&off_cnt, (size_t)p + 1) ||
        !dmalloc(&st->c5_off_ptr, (size_t)p + 1))
        return false;
    c5_tile_ptr_kernel<<<blocks_for(p + 1), kThreads, 0, st->stream>>>(p, sigma, st->nnz, st->a_m, st->a_rowptr, raw);
    c5_tile_dirty_kernel<<<blocks_for(p + 1), kThreads, 0, st->stream>>>(p, st->a_m, st->a_rowptr, raw, st->c5_tile_ptr);
    SB_TRY(cudaMemsetAsync(off_cnt, 0, ((size_t)p + 1) * sizeof(int), st->stream));
    c5_tile_desc_kernel<<<blocks_for((long long)p * 32), kThreads, 0, st->stream>>>(
        p, sigma, by, bs, st->a_rowptr, st->c5_tile_ptr, st->c5_tile_desc, off_cnt);
    SB_TRY(cudaGetLastError());
    bool ok = exclusive_scan(off_cnt, st->c5_off_ptr, p + 1, st->stream);
    dfree(raw);
    dfree(off_cnt);
    if (!ok) return false;
    uint32_t tail_word = 0;
    SB_TRY(cudaMemcpy(&st->c5_num_offsets, st->c5_off_ptr + p, sizeof(int), cudaMemcpyDeviceToHost));
    SB_TRY(cudaMemcpy(&tail_word, st->c5_tile_ptr + (p - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
    st->c5_tail_start = (int)(tail_word & 0x7FFFFFFFu);  // anonymouslib_avx2.h:177
    if (st->c5_num_offsets > 0) {
        if (!dmalloc(&st->c5_off, (size_t)st->c5_num_offsets)) return false;
        SB_TRY(cudaMemsetAsync(st->c5_off, 0, (size_t)st->c5_num_offsets * sizeof(int), st->stream));
        c5_desc_offset_kernel<<<blocks_for((long long)p * 32), kThreads, 0, st->stream>>>(
            p, sigma, by, bs, st->a_rowptr, st->c5_tile_ptr, st->c5_tile_desc, st->c5_off_ptr, st->c5_off);
        SB_TRY(cudaGetLastError());
    }
    if (!dmalloc(&st->c5_col, (size_t)st->nnz)) return false;
    if (!SB_CUDA(cudaMalloc(&st->c5_val, (size_t)st->nnz * sizeof(T)))) return false;
    c5_transpose_kernel<T><<<blocks_for(st->nnz), kThreads, 0, st->stream>>>(
        st->nnz, sigma, p, st->c5_tile_ptr, st->a_col, (const T *)st->a_val, st->c5_col, (T *)st->c5_val);
    SB_TRY(cudaGetLastError());
    return alloc_carries(st, p);
}

// Band segments (band_seg.cuh): stable bucketing of the CSR entries by col / band_cols, segment-end bits, row masks,
// tile table and the per-block positions of the merge pass.  All temporaries are released on every exit path.
template <typename T, typename MaskT>
static bool build_band_seg_t(DeviceState *st)
{
    const int K = st->coo_bands, nnz = st->nnz, m = st->m;
    st->band_cols = (int)(((long long)st->n + K - 1) / K);
    st->seg_mask64 = sizeof(MaskT) == 8;
    int *ent_row = nullptr, *idx_in = nullptr, *idx_out = nullptr, *d_tab = nullptr, *tile_segs = nullptr, *blk_cnt = nullptr, *blk_scan = nullptr, *d_cross = nullptr;
    unsigned char *key_in = nullptr, *key_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int h_ptr[kMaxPieces + 1] = {};
    auto release = [&]() { dfree(ent_row); dfree(idx_in); dfree(idx_out); dfree(d_tab); dfree(tile_segs); dfree(blk_cnt); dfree(blk_scan); dfree(d_cross);
                           dfree(key_in); dfree(key_out); dfree(tmp); };
    bool ok = dmalloc(&ent_row, (size_t)nnz) && dmalloc(&idx_in, (size_t)nnz) && dmalloc(&idx_out, (size_t)nnz) &&
              dmalloc(&key_in, (size_t)nnz) && dmalloc(&key_out, (size_t)nnz) && dmalloc(&d_tab, 4 * (size_t)(K + 1)) &&
              dmalloc(&d_cross, 1) && SB_CUDA(cudaMalloc(&st->seg_mask, ((size_t)m + 1) * sizeof(MaskT)));
    if (ok) {
        bseg_expand_kernel<MaskT><<<blocks_for(m), kThreads, 0, st->stream>>>(m, st->band_cols, K, st->rowptr, st->col, ent_row, key_in,
                                                                              (MaskT *)st->seg_mask);
        bseg_iota_kernel<<<blocks_for(nnz), kThreads, 0, st->stream>>>(nnz, idx_in);
        int bits = 1;
        while ((1 << bits) < K) ++bits;
        ok = SB_CUDA(cudaGetLastError()) &&
             SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out, idx_in, idx_out, nnz, 0, bits, st->stream)) &&
             SB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1)) &&
             SB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, idx_in, idx_out, nnz, 0, bits, st->stream));
    }
    if (ok) {
        sorted_key_ptr_kernel<<<1, kThreads, 0, st->stream>>>(nnz, K, key_out, d_tab);
        ok = SB_CUDA(cudaGetLastError()) &&
             SB_CUDA(cudaMemcpyAsync(h_ptr, d_tab, ((size_t)K + 1) * sizeof(int), cudaMemcpyDeviceToHost, st->stream)) &&
             SB_CUDA(cudaStreamSynchronize(st->stream));
    }
    if (!ok) { release(); return false; }
    // band b occupies slots [seg_ptr[b], seg_ptr[b] + seg_cnt[b]); starts aligned to 4 entries (16 bytes of ColIdx)
    long long slot = 0;
    int tiles = 0;
    for (int b = 0; b < K; ++b) {
        st->seg_ptr[b] = (int)slot;
        st->seg_cnt[b] = h_ptr[b + 1] - h_ptr[b];
        st->seg_tile0[b] = tiles;
        tiles += ceil_div(st->seg_cnt[b], kSegTile);
        slot = (slot + st->seg_cnt[b] + 3) & ~3LL;
    }
    st->seg_ptr[K] = (int)slot;
    st->seg_tile0[K] = tiles;
    st->coo_tiles = st->tiles = tiles;
    st->seg_groups = ceil_div(m, 32);
    const size_t gk = (size_t)K * st->seg_groups;
    const size_t slots = (size_t)slot + 8;
    int h_tab[4 * (kMaxPieces + 1)];
    for (int b = 0; b <= K; ++b) {
        h_tab[b] = h_ptr[b];
        h_tab[(K + 1) + b] = st->seg_ptr[b];
        h_tab[2 * (K + 1) + b] = b < K ? st->seg_cnt[b] : 0;
        h_tab[3 * (K + 1) + b] = st->seg_tile0[b];
    }
    ok = dmalloc(&st->seg_col, slots) && SB_CUDA(cudaMalloc(&st->seg_ent_val, slots * sizeof(T))) &&
         dmalloc(&st->seg_tile_ent, (size_t)tiles + 1) && dmalloc(&st->seg_tile_seg0, (size_t)tiles * kSegChunks + 4) &&
         dmalloc(&tile_segs, (size_t)tiles * kSegChunks + 4) && dmalloc(&blk_cnt, gk + 1) && dmalloc(&blk_scan, gk + 1) &&
         dmalloc(&st->seg_gbase, gk + 1) &&
         SB_CUDA(cudaMemcpyAsync(d_tab, h_tab, 4 * (size_t)(K + 1) * sizeof(int), cudaMemcpyHostToDevice, st->stream)) &&
         SB_CUDA(cudaMemsetAsync(st->seg_col, 0, slots * sizeof(int), st->stream)) &&
         SB_CUDA(cudaMemsetAsync(st->seg_ent_val, 0, slots * sizeof(T), st->stream)) &&
         SB_CUDA(cudaMemsetAsync(tile_segs, 0, ((size_t)tiles * kSegChunks + 4) * sizeof(int), st->stream)) &&
         SB_CUDA(cudaMemsetAsync(blk_cnt, 0, (gk + 1) * sizeof(int), st->stream)) &&
         SB_CUDA(cudaMemsetAsync(d_cross, 0, sizeof(int), st->stream));
    int h_cross = 0, h_segs[2] = {0, 0};
    if (ok) {
        bseg_gather_kernel<T><<<blocks_for(nnz), kThreads, 0, st->stream>>>(nnz, idx_out, key_out, ent_row, st->col, (const T *)st->val, d_tab,
                                                                           d_tab + (K + 1), st->seg_col, (T *)st->seg_ent_val);
        bseg_tile_kernel<<<blocks_for((long long)tiles * 32), kThreads, 0, st->stream>>>(tiles, K, st->band_cols, st->n, st->vsize, d_tab + 3 * (K + 1), d_tab + (K + 1), d_tab + 2 * (K + 1),
                                                                                         st->seg_col, st->seg_tile_ent, tile_segs, d_cross);
        bseg_group_count_kernel<MaskT><<<blocks_for((long long)st->seg_groups * 32), kThreads, 0, st->stream>>>(m, K, st->seg_groups, (const MaskT *)st->seg_mask, blk_cnt);
        ok = SB_CUDA(cudaGetLastError()) && exclusive_scan(tile_segs, st->seg_tile_seg0, tiles * kSegChunks + 1, st->stream) &&
             gk + 1 < 0x7fffffffULL && exclusive_scan(blk_cnt, blk_scan, (int)gk + 1, st->stream);
        if (ok) {
            bseg_transpose_kernel<<<blocks_for((long long)gk), kThreads, 0, st->stream>>>(K, st->seg_groups, blk_scan, st->seg_gbase);
            ok = SB_CUDA(cudaGetLastError());
        }
        ok = ok &&
             SB_CUDA(cudaMemcpy(&h_segs[0], st->seg_tile_seg0 + (size_t)tiles * kSegChunks, sizeof(int), cudaMemcpyDeviceToHost)) &&
             SB_CUDA(cudaMemcpy(&h_segs[1], blk_scan + gk, sizeof(int), cudaMemcpyDeviceToHost)) &&
             SB_CUDA(cudaMemcpy(&h_cross, d_cross, sizeof(int), cudaMemcpyDeviceToHost));
    }
    release();
    if (!ok) return false;
    if (h_segs[0] != h_segs[1]) { set_error("band segments: %d segment ends but %d (row, band) pairs", h_segs[0], h_segs[1]); return false; }
    st->seg_total = h_segs[0];
    st->seg_cross = h_cross != 0;
    if (!SB_CUDA(cudaMalloc(&st->seg_sums, ((size_t)st->seg_total + 8) * sizeof(T))) || !dmalloc(&st->seg_ticket, 1)) return false;
    if (!alloc_carries(st, tiles * kSegChunks)) return false;  // one carry slot per 256-entry chunk
    st->x_bands = K;
    st->kernel = SPMV_B200_KERNEL_BAND_SEG;
    return true;
}

template <typename T>
static bool build_band_seg(DeviceState *st)
{
    return st->coo_bands <= 32 ? build_band_seg_t<T, uint32_t>(st) : build_band_seg_t<T, unsigned long long>(st);
}

template <typename T>
static bool build_method(DeviceState *st, spmv_Handle *h, int method)
{
    if (st->coo_bands > 1 && method != Method_Serial) {
        if (build_band_seg<T>(st)) {
            if (method == Method_Balanced || method == Method_Balanced2) {
                // keep what a client reads from the handle: the reference's Balanced <-> Balanced2 rule (a10)
                st->ref_T = (int)(h->nthreads ? (h->nthreads > (1u << 24) ? (1u << 24) : h->nthreads) : 1);
                int ref_starved = 0;
                if (!build_splitter(st, st->m, st->rowptr, st->ref_T, &st->ref_splitter, &ref_starved)) return false;
                h->spmvMethod = ref_starved ? Method_Balanced2 : Method_Balanced;
            }
            return true;
        }
        // an optional layout: fall back to the method's own kernel on the plain CSR view
        cudaGetLastError();
        free_layouts(st);
        spmv_b200_clear_error();
        st->coo_bands = 0;
        st->x_bands = 1;
        st->kernel = SPMV_B200_KERNEL_NONE;
        st->layout_fallbacks++;
    }
    switch (method) {
    case Method_Serial:
        st->kernel = SPMV_B200_KERNEL_CSR_REFORDER;
        return true;
    case Method_Parallel:
        st->tpr = pick_tpr(st->nnz, st->a_m);
        st->kernel = SPMV_B200_KERNEL_CSR_VECTOR;
        if (!build_long_rows_threshold(st)) {  // bins / long-row list are speed-ups: plain CSR-vector still works
            cudaGetLastError();
            free_long_rows(st);
            dfree(st->bin_list); st->bin_list = nullptr;
            st->binned = false;
            st->long_thr = 0x7fffffff;
            spmv_b200_clear_error();
            st->layout_fallbacks++;
        }
        if (!build_pipeline(st)) { cudaGetLastError(); st->pipeline = false; spmv_b200_clear_error(); }
        return true;
    case Method_Balanced:
    case Method_Balanced2: {
        // mirror of the reference's demotion / promotion rule with the CALLER's nthreads (a10,
        // parallel_balanced2_spmv.c:72-94) for clients that read handle->spmvMethod
        st->ref_T = (int)(h->nthreads ? (h->nthreads > (1u << 24) ? (1u << 24) : h->nthreads) : 1);
        int ref_starved = 0;
        if (!build_splitter(st, st->m, st->rowptr, st->ref_T, &st->ref_splitter, &ref_starved)) return false;
        h->spmvMethod = ref_starved ? Method_Balanced2 : Method_Balanced;
        // the GPU geometry: row blocks of ~block_nnz non-zeros, one warp each
        long long block_nnz = opt("block_nnz");
        if (block_nnz < 32) block_nnz = 32;
        st->parts = (int)(((long long)st->nnz + block_nnz - 1) / block_nnz);
        if (st->parts < 1) st->parts = 1;
        int starved = 0;
        if (!build_splitter(st, st->a_m, st->a_rowptr, st->parts, &st->splitter, &starved)) return false;
        // The reference's rule (parallel_balanced2_spmv.c:87-94), applied to the GPU's own partition: a
        // starved block (some row spans more than a block) => merge-path; otherwise both methods run the
        // row-block kernel ("Balanced2 demoted to Balanced").  Option force_merge=1 keeps merge-path for
        // every Method_Balanced2 handle.
        if (starved || (method == Method_Balanced2 && opt("force_merge") != 0)) return build_tiles(st, /*merge=*/true);
        st->kernel = SPMV_B200_KERNEL_ROW_BLOCKS;
        return true;
    }
    case Method_Balanced_Yid:
        return build_tiles(st, /*merge=*/false);
    case Method_SellCSigma:
        return build_sell<T>(st);
    case Method_CSR5SPMV:
        return build_csr5<T>(st);
    default:
        st->kernel = SPMV_B200_KERNEL_CSR_REFORDER;
        return true;
    }
}

// "Matrix inspect and choose best method to run" (the empty heading of the reference's README.md:222), inside create:
// option "auto" = 1 lets every create whose Function is not Method_Serial pick the SPMV_METHODS value whose GPU layout
// measured fastest on matrices of that shape (profiles/r02a_bench_c2_n1.json, fraction of the measured HBM peak):
//   * more than a quarter of the non-zeros in rows much longer than the mean (power-law graphs): Method_CSR5SPMV
//     (C3: CSR5 0.390, Parallel 0.326, merge-path 0.270);
//   * short rows (mean <= 8) whose gathers are diagonal-local: Method_Parallel (C1: Parallel 0.605, SELL 0.569);
//   * everything else: Method_SellCSigma (C2: SELL 0.434, Parallel 0.398; C4: SELL 0.981, Parallel 0.883);
//   * fewer than 8192 rows: Method_Parallel (nothing to amortise a layout build).
// The statistics are the ones create gathers on the device anyway (far_fraction of the locality probe) plus one pass
// over RowPtr.  Method_Serial is never overridden: it promises the reference's bits.
static int auto_pick_method(DeviceState *st)
{
    if (st->m < 8192 || st->nnz <= 0) return Method_Parallel;
    const double mean = (double)st->nnz / st->m;
    const int cut = (int)(4.0 * mean) + 16;
    unsigned long long *heavy = nullptr, h_heavy = 0;
    bool ok = dmalloc(&heavy, 1) && SB_CUDA(cudaMemsetAsync(heavy, 0, sizeof(*heavy), st->stream));
    if (ok) {
        heavy_rows_kernel<<<blocks_for(st->m), kThreads, 0, st->stream>>>(st->m, cut, st->rowptr, heavy);
        ok = SB_CUDA(cudaGetLastError()) &&
             SB_CUDA(cudaMemcpyAsync(&h_heavy, heavy, sizeof(h_heavy), cudaMemcpyDeviceToHost, st->stream)) &&
             SB_CUDA(cudaStreamSynchronize(st->stream));
    }
    dfree(heavy);
    if (!ok) { cudaGetLastError(); spmv_b200_clear_error(); return Method_Parallel; }
    if (4.0 * (double)h_heavy > (double)st->nnz) return Method_CSR5SPMV;
    if (mean <= 8.0 && st->far_fraction <= 0.25) return Method_Parallel;
    return Method_SellCSigma;
}

static bool build_state(DeviceState *st, spmv_Handle *h, int m, int n, int *RowPtr, int *ColIdx, void *Val,
                        int method)
{
    st->requested = method;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: libspmv_b200 has no CPU fallback");
        return false;
    }
    SB_TRY(cudaGetDevice(&st->device));
    apply_device_limits(st);
    if (m < 0 || n < 0 || (m > 0 && (!RowPtr))) { set_error("invalid arguments (m=%d n=%d RowPtr=%p)", m, n, (void *)RowPtr); return false; }
    st->m = m;
    st->n = n;
    st->vsize = (h->data_size == sizeof(double)) ? 8 : 4;  // reference serial_spmv.c:48-54
    st->requested = method;

    const bool dev_csr = is_device_ptr(RowPtr);
    int ends[2] = {0, 0};
    if (dev_csr) {
        if (m > 0) {
            SB_TRY(cudaMemcpy(&ends[0], RowPtr, sizeof(int), cudaMemcpyDeviceToHost));
            SB_TRY(cudaMemcpy(&ends[1], RowPtr + m, sizeof(int), cudaMemcpyDeviceToHost));
        }
    } else if (m > 0) {
        ends[0] = RowPtr[0];
        ends[1] = RowPtr[m];
    }
    if (ends[0] != 0 || ends[1] < 0) { set_error("RowPtr[0] must be 0 and RowPtr[m] >= 0 (got %d, %d)", ends[0], ends[1]); return false; }
    if ((long long)ends[1] > 2147483647LL - 8192) { set_error("nnz too large for 32-bit tiles"); return false; }
    st->nnz = ends[1];
    if (st->nnz > 0 && (!ColIdx || !Val)) { set_error("ColIdx / Matrix_Val are NULL"); return false; }

    if (dev_csr) {
        // device-resident CSR: adopted in place (borrowed, like the reference borrows host arrays)
        if (st->nnz > 0 && (!is_device_ptr(ColIdx) || !is_device_ptr(Val))) { set_error("RowPtr is a device pointer but ColIdx / Matrix_Val are not"); return false; }
        st->rowptr = RowPtr;
        st->col = ColIdx;
        st->val = Val;
        st->owns_csr = false;
    } else {
        // upload once (32 bytes of slack behind every array)
        st->owns_csr = true;
        if (!dmalloc(&st->rowptr, (size_t)m + 1 + 8) || !dmalloc(&st->col, (size_t)st->nnz + 8)) return false;
        if (!SB_CUDA(cudaMalloc(&st->val, ((size_t)st->nnz + 8) * st->vsize))) return false;
        if (m > 0) SB_TRY(cudaMemcpy(st->rowptr, RowPtr, ((size_t)m + 1) * sizeof(int), cudaMemcpyHostToDevice));
        else SB_TRY(cudaMemset(st->rowptr, 0, sizeof(int)));
        if (st->nnz > 0) {
            SB_TRY(cudaMemcpy(st->col, ColIdx, (size_t)st->nnz * sizeof(int), cudaMemcpyHostToDevice));
            SB_TRY(cudaMemcpy(st->val, Val, (size_t)st->nnz * st->vsize, cudaMemcpyHostToDevice));
        }
    }

    st->a_m = st->m; st->a_rowptr = st->rowptr; st->a_col = st->col; st->a_val = st->val;
    if (method != Method_Serial) {  // Method_Serial keeps the reference's exact summation order: never banded
        const bool okb = st->vsize == 8 ? build_band_major<double>(st) : build_band_major<float>(st);
        if (!okb) return false;
    }
    if (m > 0) {
        int *flag = nullptr;
        if (!dmalloc(&flag, 1)) return false;
        SB_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st->stream));
        empty_rows_kernel<<<blocks_for(st->a_m), kThreads, 0, st->stream>>>(st->a_m, st->a_rowptr, flag);
        int e = 0;
        const bool ok = read_flag(flag, st->stream, &e);
        dfree(flag);
        if (!ok) return false;
        st->has_empty_rows = e != 0;
    }
    if (m == 0 || st->nnz == 0) {  // nothing to lay out: spmv() only has zeros to write
        st->kernel = SPMV_B200_KERNEL_NONE;
        return true;
    }
    st->auto_method = -1;
    if (opt("auto") != 0 && method != Method_Serial) {
        method = st->auto_method = auto_pick_method(st);
        if (st->auto_method == Method_SellCSigma || st->auto_method == Method_CSR5SPMV || st->auto_method == Method_Parallel)
            h->spmvMethod = (SPMV_METHODS)st->auto_method;  // what actually runs, for clients that read the field
    }
    const bool ok = st->vsize == 8 ? build_method<double>(st, h, method) : build_method<float>(st, h, method);
    if (!ok) return false;
    SB_TRY(cudaStreamSynchronize(st->stream));
    // Our own upload of the CSR is dead weight once every kernel of the handle reads a re-laid-out copy (the
    // band-major copy or the COO bands): give those gigabytes back (an adopted device CSR is the caller's).
    if (st->owns_csr && (st->kernel == SPMV_B200_KERNEL_BAND_SEG || st->a_rowptr != st->rowptr)) {
        dfree(st->rowptr); dfree(st->col); dfree(st->val);
        st->rowptr = st->col = nullptr;
        st->val = nullptr;
        st->owns_csr = false;
        st->released_csr = true;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// launch dispatch (a3: the reference's spmv_functions[] table, common.c:85-94)
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC>
static void launch_vector(DeviceState *st, int tpr, int row0, int row1, const T *x, T *y, const PeerList<T> &peers, int fuse_bands)
{
    const int grid = blocks_for((long long)(row1 - row0) * tpr);
    if (grid <= 0) return;
#define SB_ARGS row0, row1, st->nnz, st->long_thr, st->a_rowptr, st->a_col, (const T *)st->a_val, x, y, peers, fuse_bands, st->m, (const T *)st->v_y
#define SB_CASE(N) case N: \
        if (fuse_bands > 0) csr_vector_kernel<T, N, VEC, false, true><<<grid, kThreads, 0, st->stream>>>(SB_ARGS); \
        else if (peers.n > 0) csr_vector_kernel<T, N, VEC, true, false><<<grid, kThreads, 0, st->stream>>>(SB_ARGS); \
        else csr_vector_kernel<T, N, VEC, false, false><<<grid, kThreads, 0, st->stream>>>(SB_ARGS); \
        break;
    switch (tpr) { SB_CASE(1) SB_CASE(2) SB_CASE(4) SB_CASE(8) SB_CASE(16) default: SB_CASE(32) }
#undef SB_CASE
#undef SB_ARGS
    count_launch();
}

// rows [row0, row1) of the active view; fuse_bands > 0: they belong to the last band and y is the FINAL y
template <typename T>
static void launch_vector_mode(DeviceState *st, int row0, int row1, const T *x, T *y, const PeerList<T> &peers, int fuse_bands = 0)
{
    launch_vector<T, 4>(st, st->tpr, row0, row1, x, y, peers, fuse_bands);
}

// CSR-vector over rows [row0, m) only (the CSR tail of SELL)
template <typename T>
__global__ void __launch_bounds__(kThreads)
csr_tail_kernel(int row0, int m, int long_thr, const int *__restrict__ rowptr, const int *__restrict__ col,
                const T *__restrict__ val, const T *__restrict__ x, T *__restrict__ y)
{
    const uint64_t pl = policy_evict_last();
    const long long w = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row0 + w >= m) return;
    const int row = (int)(row0 + w);
    const int start = rowptr[row];
    int end = rowptr[row + 1];
    if (end - start > long_thr) end = start;  // hub row: y = 0 here, the long-row path adds the whole row
    T sum = 0;
    for (int j = start + lane; j < end; j += 32) sum = fma_t(val[j], ldg_x(x + col[j], pl), sum);
    sum = group_sum_c<T, 32>(sum);
    if (lane == 0) y[row] = sum;
}

static int opt_cached_seg_prefetch()
{
    static const int v = (int)opt("seg_prefetch");  // read once: spmv() is the hot path
    return v;
}

// band segments, pass 1 over the tiles of bands [b0, b0 + count)
template <typename T>
static bool launch_seg_bands(DeviceState *st, int b0, int count, const T *x)
{
    const int t0 = st->seg_tile0[b0], t1 = st->seg_tile0[b0 + count];
    if (t1 <= t0) return true;
    if (st->seg_grid == 0) {  // persistent CTAs: option seg_ctas per SM (shared memory AND registers allow 2 or 3)
        st->seg_ctas = (int)opt("seg_ctas");
        if (st->seg_ctas != 3) st->seg_ctas = 2;
        const void *fn = st->seg_ctas == 3 ? (const void *)bseg_kernel<T, 3> : (const void *)bseg_kernel<T, 2>;
        SB_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bseg_smem_bytes<T>()));
        // leave the rest of the unified array to L1: the gathers need lines for their outstanding misses
        const int carve = (int)((st->seg_ctas * (bseg_smem_bytes<T>() + 2048) * 100 + 228 * 1024 - 1) / (228 * 1024));
        SB_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve > 100 ? 100 : carve));
        st->seg_grid = st->seg_ctas * (st->dev_sms > 0 ? st->dev_sms : 1);
    }
    const int grid = t1 - t0 < st->seg_grid ? t1 - t0 : st->seg_grid;
    // ticket counter of the dynamic tile schedule: tickets 0 .. 3*grid-1 are implicit (every CTA's first three tiles)
    set_int_kernel<<<1, 1, 0, st->stream>>>(st->seg_ticket, kSegStages * grid);
    // L2 prefetch of the next band's x slice pays only when two slices fit the part of L2 random gathers can use
    // (measured on the C5 shard: 64 MiB slices 5.78 -> 6.13 ms with it, 32 MiB slices 5.81 -> 5.76 ms and DRAM reads
    // 14.0 -> 9.9 GB); option seg_prefetch: 0 = never, 1 = automatic, 2 = always
    const int pf_opt = opt_cached_seg_prefetch();
    const bool two_fit = 2.0 * (double)st->band_cols * st->vsize <= 0.53 * (double)st->dev_l2;
    const int pf_on = (((uintptr_t)x & 15u) == 0 && (pf_opt == 2 || (pf_opt == 1 && two_fit))) ? 1 : 0;
#define SB_SEG(MINB) bseg_kernel<T, MINB><<<grid, kSegThreads, bseg_smem_bytes<T>(), st->stream>>>(t0, t1, pf_on, st->seg_ticket, st->seg_tile_ent, st->seg_tile_seg0, \
        st->seg_col, (const T *)st->seg_ent_val, x, (T *)st->seg_sums, (T *)st->carry_val, st->carry_row)
    if (st->seg_ctas == 3) SB_SEG(3); else SB_SEG(2);
#undef SB_SEG
    count_launch();
    return SB_CUDA(cudaGetLastError());
}

// band segments: carries of segments that cross a tile boundary, then pass 2 (y written once, peers included)
template <typename T>
static bool launch_seg_finish(DeviceState *st, T *y_out)
{
    cudaStream_t s = st->stream;
    if (st->seg_cross) launch_carry_fixup<T>(st, s, st->tiles * kSegChunks, st->carry_row, (const T *)st->carry_val, (T *)st->seg_sums);
    PeerList<T> pr;
    pr.n = st->n_peers;
    for (int i = 0; i < kMaxPeers; ++i) pr.p[i] = (T *)st->peers[i];
#define SB_MERGE(M, P) bseg_merge_kernel<T, M, P><<<blocks_for((long long)st->seg_groups * 32), kThreads, 0, s>>>(0, st->m, st->coo_bands, (const M *)st->seg_mask, \
                                                                                     st->seg_gbase, (const T *)st->seg_sums, y_out, pr)
    if (st->seg_mask64) { if (pr.n > 0) SB_MERGE(unsigned long long, true); else SB_MERGE(unsigned long long, false); }
    else { if (pr.n > 0) SB_MERGE(uint32_t, true); else SB_MERGE(uint32_t, false); }
#undef SB_MERGE
    count_launch();
    return SB_CUDA(cudaGetLastError());
}

template <typename T>
static bool launch(DeviceState *st, const T *x, T *y_out)
{
    cudaStream_t s = st->stream;
    if (st->m == 0) return true;
    if (st->kernel == SPMV_B200_KERNEL_NONE) {  // nnz == 0: y = 0
        return SB_CUDA(cudaMemsetAsync(y_out, 0, (size_t)st->m * sizeof(T), s));  // all-zero bits = +0.0
    }
    if (st->kernel == SPMV_B200_KERNEL_BAND_SEG) {
        // pass 1: segment sums of every tile (one launch, tiles in band order); carries of segments that cross a
        // tile boundary; pass 2: every row adds its segment sums in band order and writes y once
        return launch_seg_bands<T>(st, 0, st->coo_bands, x) && launch_seg_finish<T>(st, y_out);
    }
    // the active view: the CSR itself, or its band-major copy writing the virtual y
    const int m = st->a_m;
    const bool banded = st->x_bands > 1;
    T *y = banded ? (T *)st->v_y : y_out;
    const T *val = (const T *)st->a_val;
    // extra y destinations (fused all-gather): written by whichever kernel produces the FINAL y
    PeerList<T> peers, none;
    none.n = 0;
    peers.n = st->n_peers;
    for (int i = 0; i < kMaxPeers; ++i) { peers.p[i] = (T *)st->peers[i]; none.p[i] = nullptr; }
    const PeerList<T> &direct = banded ? none : peers;  // kernels write virtual y when banded
    bool scattered = peers.n == 0;
    switch (st->kernel) {
    case SPMV_B200_KERNEL_CSR_REFORDER: {
        constexpr int L = sizeof(T) == 8 ? 4 : 8;
        csr_reforder_kernel<T><<<blocks_for((long long)m * L), kThreads, 0, s>>>(m, st->a_rowptr, st->a_col, val, x, y);
        count_launch();
        break;
    }
    case SPMV_B200_KERNEL_CSR_VECTOR:
        // (one launch over all bands + band_reduce measured faster than two launches with the reduce fused into
        // the last band: 2.59 vs 2.64 ms on C2; the fused form is used by the pipelined host path, where the
        // last band runs in row chunks anyway)
        if (st->binned) {
            static const int lanes[3] = {1, 4, 16};
            for (int b = 0; b < 3; ++b) {
                const int count = st->bin_ptr[b + 1] - st->bin_ptr[b];
                if (count <= 0) continue;
                const int grid = blocks_for((long long)count * lanes[b]);
                const int *list = st->bin_list + st->bin_ptr[b];
#define SB_BIN(N) if (direct.n > 0) csr_vector_list_kernel<T, N, true><<<grid, kThreads, 0, s>>>(count, st->nnz, list, st->a_rowptr, st->a_col, val, x, y, direct); \
                  else csr_vector_list_kernel<T, N, false><<<grid, kThreads, 0, s>>>(count, st->nnz, list, st->a_rowptr, st->a_col, val, x, y, direct)
                if (b == 0) { SB_BIN(1); } else if (b == 1) { SB_BIN(4); } else { SB_BIN(16); }
#undef SB_BIN
                count_launch();
            }
        } else {
            launch_vector_mode<T>(st, 0, m, x, y, direct);
        }
        scattered = scattered || !banded;
        break;
    case SPMV_B200_KERNEL_ROW_BLOCKS: {
        const int grid = blocks_for((long long)st->parts * 32);
#define SB_RB(V, P) row_block_kernel<T, V, P><<<grid, kThreads, 0, s>>>(st->parts, st->nnz, st->splitter, st->a_rowptr, st->a_col, val, x, y, direct)
        if (direct.n > 0) SB_RB(4, true); else SB_RB(4, false);
#undef SB_RB
        scattered = scattered || !banded;
        count_launch();
        break;
    }
    case SPMV_B200_KERNEL_MERGE_PATH: {
        if (st->tile_items == 4)
            merge_path_kernel<T, 4><<<st->tiles, kThreads, 0, s>>>(m, st->nnz, st->merge_coords, st->a_rowptr, st->a_col, val, x, y, (T *)st->carry_val, st->carry_row);
        else
            merge_path_kernel<T, 8><<<st->tiles, kThreads, 0, s>>>(m, st->nnz, st->merge_coords, st->a_rowptr, st->a_col, val, x, y, (T *)st->carry_val, st->carry_row);
        launch_carry_fixup<T>(st, s, st->tiles, st->carry_row, (const T *)st->carry_val, y);
        count_launch();
        break;
    }
    case SPMV_B200_KERNEL_NNZ_SPLIT: {
        if (st->tile_items == 4)
            nnz_split_kernel<T, 4><<<st->tiles, kThreads, 0, s>>>(m, st->nnz, st->tile_rows, st->a_rowptr, st->a_col, val, x, y, (T *)st->carry_val, st->carry_row);
        else
            nnz_split_kernel<T, 8><<<st->tiles, kThreads, 0, s>>>(m, st->nnz, st->tile_rows, st->a_rowptr, st->a_col, val, x, y, (T *)st->carry_val, st->carry_row);
        launch_carry_fixup<T>(st, s, st->tiles, st->carry_row, (const T *)st->carry_val, y);
        count_launch();
        break;
    }
    case SPMV_B200_KERNEL_SELL: {
        if (st->slices > 0) {
            const PeerList<T> &sp = (st->banner == m) ? direct : none;
            const int sgrid = blocks_for((long long)st->slices * 32);
#define SB_SELL(P, U, MINB) sell_kernel<T, P, U, MINB><<<sgrid, kThreads, 0, s>>>( \
                st->slices, st->sell_slice_ptr, st->sell_full, st->sell_perm, st->sell_col, (const T *)st->sell_val, x, y, sp)
            if (sp.n > 0) SB_SELL(true, 8, 6);
            else if (st->sell_variant == 2) SB_SELL(false, 4, 8);
            else SB_SELL(false, 8, 6);
#undef SB_SELL
            scattered = scattered || (!banded && st->banner == m);
            count_launch();
        }
        if (st->banner < m) {
            csr_tail_kernel<T><<<blocks_for((long long)(m - st->banner) * 32), kThreads, 0, s>>>(st->banner, m, st->long_thr, st->a_rowptr, st->a_col, val, x, y);
            count_launch();
        }
        break;
    }
    case SPMV_B200_KERNEL_CSR5: {
        const int p = st->c5_p;
        if (st->has_empty_rows) {
            if (!SB_CUDA(cudaMemsetAsync(y, 0, (size_t)m * sizeof(T), s))) return false;  // all-zero bits = +0.0
        }
        const T *tval = (const T *)st->c5_val;
        if (p > 1) {
            const int grid = blocks_for((long long)(p - 1) * 32);
#define SB_C5(SG) csr5_kernel<T, SG><<<grid, kThreads, 0, s>>>(p, st->c5_bit_y, st->c5_bit_ss, st->c5_tile_ptr, st->c5_tile_desc, st->c5_off_ptr, st->c5_off, st->c5_col, tval, x, y, (T *)st->carry_val, st->carry_row)
            if (st->c5_sigma == 4) SB_C5(4); else if (st->c5_sigma == 8) SB_C5(8); else SB_C5(16);
#undef SB_C5
            count_launch();
        }
        const int tail_nz0 = (p - 1) * kC5Omega * st->c5_sigma;
        csr5_tail_kernel<T><<<blocks_for((long long)(m - st->c5_tail_start) * 32), kThreads, 0, s>>>(
            m, st->c5_tail_start, tail_nz0, p - 1, st->a_rowptr, st->c5_col, tval, x, y, (T *)st->carry_val, st->carry_row);
        launch_carry_fixup<T>(st, s, p, st->carry_row, (const T *)st->carry_val, y);
        count_launch();
        break;
    }
    default:
        set_error("handle has no kernel (%d)", st->kernel);
        return false;
    }
    if (st->lr_rows > 0) {  // hub rows / SELL overflow: segment sums, then one ordered add per row
        const int single = direct.n == 0;  // with peer destinations every final value goes through long_final_kernel
        long_seg_kernel<T><<<blocks_for((long long)st->lr_segs * 32), kThreads, 0, s>>>(
            st->lr_segs, st->lr_seg_row, st->lr_row, st->lr_start, st->lr_seg_ptr, st->a_rowptr, st->a_col, val, x, (T *)st->lr_partial,
            y, single, st->lr_accumulate);
        const int fgrid = blocks_for((long long)st->lr_rows * 32);
        if (direct.n > 0) long_final_kernel<T, true><<<fgrid, kThreads, 0, s>>>(st->lr_rows, st->lr_accumulate, single, st->lr_row, st->lr_seg_ptr, (const T *)st->lr_partial, y, direct);
        else long_final_kernel<T, false><<<fgrid, kThreads, 0, s>>>(st->lr_rows, st->lr_accumulate, single, st->lr_row, st->lr_seg_ptr, (const T *)st->lr_partial, y, direct);
        count_launch(2);
    }
    if (banded) {
        if (peers.n > 0) band_reduce_kernel<T, true><<<blocks_for(st->m), kThreads, 0, s>>>(0, st->m, st->m, st->x_bands, (const T *)st->v_y, y_out, peers);
        else band_reduce_kernel<T, false><<<blocks_for(st->m), kThreads, 0, s>>>(0, st->m, st->m, st->x_bands, (const T *)st->v_y, y_out, peers);
        scattered = true;
        count_launch();
    }
    if (!scattered) {  // kernel family without a fused epilogue: one stream-ordered copy to the peers
        peer_copy_kernel<T><<<blocks_for(st->m), kThreads, 0, s>>>(st->m, y_out, peers);
        count_launch();
    }
    return SB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Host x and host y through the CSR-vector kernel, pipelined over PCIe: x goes up in pieces on s_in, the
// compute stream starts every band / row chunk as soon as the part of x it gathers from has arrived, and
// finished chunks of y go back on s_out while later chunks are still being computed.  Same kernels, same
// per-row arithmetic as the one-shot path: the bits of y do not depend on which path ran.
// ------------------------------------------------------------------------------------------------
template <typename T>
static bool run_pipelined(DeviceState *st, const T *hx, T *hy)
{
    if (!st->s_in) {
        SB_TRY(cudaStreamCreateWithFlags(&st->s_in, cudaStreamNonBlocking));
        SB_TRY(cudaStreamCreateWithFlags(&st->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxPieces; ++i) SB_TRY(cudaEventCreateWithFlags(&st->ev_in[i], cudaEventDisableTiming));
        for (int i = 0; i < kPipeChunks; ++i) SB_TRY(cudaEventCreateWithFlags(&st->ev_out[i], cudaEventDisableTiming));
        SB_TRY(cudaEventCreateWithFlags(&st->ev_start, cudaEventDisableTiming));
    }
    cudaStream_t s = st->stream;
    T *xd = (T *)st->x_stage, *yd = (T *)st->y_stage;
    const int m = st->m, n = st->n;
    const bool banded = st->x_bands > 1;
    const int pieces = banded ? st->x_bands : kPipeChunks;
    PeerList<T> none;
    none.n = 0;
    for (int i = 0; i < kMaxPeers; ++i) none.p[i] = nullptr;
    // whatever the caller queued on the compute stream before this call stays ahead of the copies
    SB_TRY(cudaEventRecord(st->ev_start, s));
    SB_TRY(cudaStreamWaitEvent(st->s_in, st->ev_start, 0));
    SB_TRY(cudaStreamWaitEvent(st->s_out, st->ev_start, 0));
    auto piece_lo = [&](int p) -> long long { return banded ? (long long)p * st->band_cols : (long long)n * p / pieces; };
    auto piece_hi = [&](int p) -> long long {
        long long hi = banded ? (long long)(p + 1) * st->band_cols : (long long)n * (p + 1) / pieces;
        return (p == pieces - 1 || hi > n) ? n : hi;
    };
    for (int p = 0; p < pieces; ++p) {
        const long long lo = piece_lo(p), hi = piece_hi(p);
        if (hi > lo) SB_TRY(cudaMemcpyAsync(xd + lo, hx + lo, (size_t)(hi - lo) * sizeof(T), cudaMemcpyHostToDevice, st->s_in));
        SB_TRY(cudaEventRecord(st->ev_in[p], st->s_in));
    }
    int waited = -1;
    auto need_piece = [&](int p) -> bool {
        for (; waited < p; ++waited) SB_TRY(cudaStreamWaitEvent(s, st->ev_in[waited + 1], 0));
        return true;
    };
    auto chunk_out = [&](int c, int r0, int r1) -> bool {
        SB_TRY(cudaEventRecord(st->ev_out[c], s));
        SB_TRY(cudaStreamWaitEvent(st->s_out, st->ev_out[c], 0));
        if (r1 > r0) SB_TRY(cudaMemcpyAsync(hy + r0, yd + r0, (size_t)(r1 - r0) * sizeof(T), cudaMemcpyDeviceToHost, st->s_out));
        return true;
    };
    if (banded) {
        const int K = st->x_bands;
        for (int b = 0; b + 1 < K; ++b) {
            if (!need_piece(b)) return false;
            launch_vector_mode<T>(st, b * m, (b + 1) * m, xd, (T *)st->v_y, none);
        }
        if (!need_piece(K - 1)) return false;
        for (int c = 0; c < kPipeChunks; ++c) {
            const int r0 = (int)((long long)m * c / kPipeChunks), r1 = (int)((long long)m * (c + 1) / kPipeChunks);
            launch_vector_mode<T>(st, (K - 1) * m + r0, (K - 1) * m + r1, xd, yd, none, K - 1);
            if (!chunk_out(c, r0, r1)) return false;
        }
    } else {
        for (int c = 0; c < kPipeChunks; ++c) {
            const int r0 = (int)((long long)m * c / kPipeChunks), r1 = (int)((long long)m * (c + 1) / kPipeChunks);
            int p = 0;
            while (p < pieces - 1 && piece_hi(p) < st->chunk_xmax[c]) ++p;
            if (!need_piece(p)) return false;
            launch_vector_mode<T>(st, r0, r1, xd, yd, none);
            if (!chunk_out(c, r0, r1)) return false;
        }
    }
    SB_TRY(cudaGetLastError());
    SB_TRY(cudaStreamSynchronize(st->s_out));  // every chunk of y is home (and so every kernel has run)
    SB_TRY(cudaStreamSynchronize(st->s_in));   // x may be modified by the caller from here on
    return true;
}

static DeviceState *state_of(const spmv_Handle *h)
{
    if (!h || !h->extraHandle) return nullptr;
    DeviceState *st = (DeviceState *)h->extraHandle;
    return st->magic == 0x5b200a11u ? st : nullptr;
}

static void handle_init(spmv_Handle *h)  // reference gemv_Handle_init, common.c:18-29
{
    h->spmvMethod = Method_Serial;
    h->nthreads = 0;
    h->extraHandle = nullptr;
    h->RowPtr = nullptr;
    h->ColIdx = nullptr;
    h->Matrix_Val = nullptr;
    h->Y_temp = nullptr;
    h->index = nullptr;
    h->Level_3_opt_used = 0;
}

}  // namespace sb

using namespace sb;

// ================================================================================================
// exported data (reference common.c:306-339; same strings, same order)
// ================================================================================================
extern "C" {

const char *funcNames[] = {
    "Method_Serial_VECTOR_NONE", "Method_Serial_VECTOR_AVX2", "Method_Serial_VECTOR_AVX512",
    "Method_Parallel_VECTOR_NONE", "Method_Parallel_VECTOR_AVX2", "Method_Parallel_VECTOR_AVX512",
    "Method_Balanced_VECTOR_NONE", "Method_Balanced_VECTOR_AVX2", "Method_Balanced_VECTOR_AVX512",
    "Method_Balanced2_VECTOR_NONE", "Method_Balanced2_VECTOR_AVX2", "Method_Balanced2_VECTOR_AVX512",
    "Method_SellCSigma_VECTOR_NONE", "Method_SellCSigma_VECTOR_AVX2", "Method_SellCSigma_VECTOR_AVX512",
    "Method_Csr5Spmv_VECTOR_NONE", "Method_Csr5Spmv_VECTOR_AVX2", "Method_Csr5Spmv_VECTOR_AVX512"};
const char *Methods_names[] = {"Method_Serial", "Method_Parallel", "Method_Balanced", "Method_Balanced2",
                               "Method_BalancedYid", "Method_SellCSigma", "Method_Csr5Spmv"};
const char *Vectorized_names[] = {"VECTOR_NONE", "VECTOR_AVX2", "VECTOR_AVX512"};

// ================================================================================================
// the four drop-in entry points
// ================================================================================================
// everything create does to an initialised public struct (shared by create and spmv_b200_update_values)
static void create_into(spmv_Handle *h, BASIC_INT_TYPE m, BASIC_INT_TYPE n, BASIC_INT_TYPE *RowPtr, BASIC_INT_TYPE *ColIdx,
                        void *Matrix_Val, BASIC_SIZE_TYPE nthreads, int method, BASIC_SIZE_TYPE size, VECTORIZED_WAY vectorizedWay)
{
    if (method < (int)Method_Serial || method >= (int)Method_Total_Size) method = Method_Serial;  // common.c:136
    h->nthreads = nthreads;  // handle_init_common_parameters, common.c:74-83
    h->vectorizedWay = vectorizedWay;
    h->data_size = size;
    h->spmvMethod = (SPMV_METHODS)method;
    h->RowPtr = RowPtr;  // borrowed, never written (common.c:157-159)
    h->ColIdx = ColIdx;
    h->Matrix_Val = Matrix_Val;
    // (fp32 Method_CSR5SPMV: the reference silently runs SELL and stores Method_SellCSigma in the handle,
    // common.c:177-180; here it is a real fp32 CSR5 and the public field keeps Method_CSR5SPMV -- a documented
    // deviation, include/spmv.h)
    DeviceState *st = new DeviceState();
    h->extraHandle = st;
    // Option "reorder": the reference's level-3 hook (common.c:144-156, compiled out upstream).  Same condition as
    // there (square, > 8096 rows, > 100000 non-zeros), HOST arrays only; the device layout is then built from
    // A' = P A P^T and the permutation is handed to the caller in handle->index (reorder.cu).
    std::vector<int> p_rowptr, p_col;
    std::vector<double> p_val;  // (storage only: 8-byte aligned, holds fp32 or fp64 values)
    if (opt("reorder") != 0 && m == n && m > 8096 && RowPtr && ColIdx && Matrix_Val && !is_device_ptr(RowPtr) &&
        RowPtr[0] == 0 && RowPtr[m] > 100000) {
        const size_t nnz = (size_t)RowPtr[m];
        int *index = (int *)malloc(((size_t)m + 1) * sizeof(int));
        bool done = false;
        if (index && spmv_b200_reorder(m, RowPtr, ColIdx, index) == 0) {
            p_rowptr.resize((size_t)m + 1);
            p_col.resize(nnz);
            p_val.resize(nnz);
            done = spmv_b200_permute_csr(m, RowPtr, ColIdx, Matrix_Val, size, index, p_rowptr.data(), p_col.data(), p_val.data()) == 0;
        }
        if (done) {
            index[m] = m;
            h->index = index;           // freed by spmv_clear_handle
            h->Level_3_opt_used = 1;
            st->ok = build_state(st, h, m, n, p_rowptr.data(), p_col.data(), p_val.data(), method);
            return;
        }
        free(index);
    }
    st->ok = build_state(st, h, m, n, RowPtr, ColIdx, Matrix_Val, method);
}

void spmv_create_handle_all_in_one(spmv_Handle_t *Handle, BASIC_INT_TYPE m, BASIC_INT_TYPE n,
                                   BASIC_INT_TYPE *RowPtr, BASIC_INT_TYPE *ColIdx, void *Matrix_Val,
                                   BASIC_SIZE_TYPE nthreads, SPMV_METHODS Function, BASIC_SIZE_TYPE size,
                                   VECTORIZED_WAY vectorizedWay, const char *MtxToken)
{
    (void)MtxToken;  // only keys the reference's compiled-out METIS cache (common.c:152-154)
    if (!Handle) return;
    spmv_Handle *h = (spmv_Handle *)malloc(sizeof(spmv_Handle));  // freed by spmv_destory_handle
    *Handle = h;
    if (!h) return;
    handle_init(h);
    create_into(h, m, n, RowPtr, ColIdx, Matrix_Val, nthreads, (int)Function, size, vectorizedWay);
}

void spmv(const spmv_Handle_t handle, BASIC_INT_TYPE m, const BASIC_INT_TYPE *RowPtr,
          const BASIC_INT_TYPE *ColIdx, const void *Matrix_Val, const void *Vector_Val_X, void *Vector_Val_Y)
{
    (void)m; (void)RowPtr; (void)ColIdx; (void)Matrix_Val;  // the handle owns the device copies
    DeviceState *st = state_of(handle);  // NULL handle: silent return (common.c:285)
    if (!st || !st->ok || !Vector_Val_Y || (!Vector_Val_X && st->n > 0)) return;
    std::lock_guard<std::mutex> lock(st->mu);
    DeviceGuard guard(st->device);
    const bool x_dev = is_device_ptr(Vector_Val_X), y_dev = is_device_ptr(Vector_Val_Y);
    const void *xd = Vector_Val_X;
    void *yd = Vector_Val_Y;
    const size_t xb = (size_t)st->n * st->vsize, yb = (size_t)st->m * st->vsize;
    if (st->pin_host) {
        if (!x_dev && xb) note_host_buffer(st, 0, Vector_Val_X, xb);
        if (!y_dev && yb) note_host_buffer(st, 1, Vector_Val_Y, yb);
    }
    if (!x_dev && !st->x_stage && !SB_CUDA(cudaMalloc(&st->x_stage, xb ? xb : 1))) return;
    if (!y_dev && !st->y_stage && !SB_CUDA(cudaMalloc(&st->y_stage, yb ? yb : 1))) return;
    if (!x_dev && !y_dev && st->pipeline && st->n_peers == 0 && st->kernel == SPMV_B200_KERNEL_CSR_VECTOR && xb && yb) {
        if (st->vsize == 8) run_pipelined<double>(st, (const double *)Vector_Val_X, (double *)Vector_Val_Y);
        else run_pipelined<float>(st, (const float *)Vector_Val_X, (float *)Vector_Val_Y);
        return;
    }
    if (!x_dev) {
        if (xb && !SB_CUDA(cudaMemcpyAsync(st->x_stage, Vector_Val_X, xb, cudaMemcpyHostToDevice, st->stream))) return;
        xd = st->x_stage;
        if (y_dev && xb) {
            // host x, device y: the call returns without a stream sync, but the caller may reuse X at once (a pinned
            // X makes this copy truly asynchronous): wait for the copy alone
            if (!st->ev_x) { if (!SB_CUDA(cudaEventCreateWithFlags(&st->ev_x, cudaEventDisableTiming))) return; }
            if (!SB_CUDA(cudaEventRecord(st->ev_x, st->stream)) || !SB_CUDA(cudaEventSynchronize(st->ev_x))) return;
        }
    }
    if (!y_dev) yd = st->y_stage;
    const bool ok = st->vsize == 8 ? launch<double>(st, (const double *)xd, (double *)yd)
                                   : launch<float>(st, (const float *)xd, (float *)yd);
    if (!ok) return;
    if (!y_dev) {
        if (yb && !SB_CUDA(cudaMemcpyAsync(Vector_Val_Y, yd, yb, cudaMemcpyDeviceToHost, st->stream))) return;
        SB_CUDA(cudaStreamSynchronize(st->stream));  // host y is complete on return, as in the reference
    }
}

void spmv_clear_handle(spmv_Handle_t this_handle)  // reference gemv_Handle_clear, common.c:31-52,69-71
{
    if (!this_handle) return;
    free_state(state_of(this_handle));
    if (this_handle->Level_3_opt_used && this_handle->index) free(this_handle->index);  // ours (option "reorder")
    handle_init(this_handle);
}

void spmv_destory_handle(spmv_Handle_t this_handle)  // reference common.c:54-61
{
    if (!this_handle) return;
    spmv_clear_handle(this_handle);
    free(this_handle);
}

// The reference re-reads the caller's Matrix_Val on every spmv() (Serial / Parallel / Balanced*); here the values are
// part of the device layout.  A client that changes values on a fixed pattern calls this instead of destroy + create:
// the handle is rebuilt in place from its own borrowed RowPtr / ColIdx, the values given here (NULL: the Matrix_Val
// it was created with, re-read), the same method, precision, nthreads and stream.  Cost: one create.
int spmv_b200_update_values(spmv_Handle_t handle, void *Matrix_Val)
{
    DeviceState *st = state_of(handle);
    if (!st || !st->ok) return -1;  // (a handle whose create failed has nothing to refresh)
    const int m = st->m, n = st->n, method = st->requested;
    cudaStream_t stream = st->stream;
    BASIC_INT_TYPE *rp = handle->RowPtr, *ci = handle->ColIdx;
    void *va = Matrix_Val ? Matrix_Val : handle->Matrix_Val;
    const BASIC_SIZE_TYPE nthreads = handle->nthreads, size = handle->data_size;
    const VECTORIZED_WAY vw = handle->vectorizedWay;
    spmv_clear_handle(handle);
    create_into(handle, m, n, rp, ci, va, nthreads, method, size, vw);
    DeviceState *fresh = state_of(handle);
    if (fresh) fresh->stream = stream;
    return fresh && fresh->ok ? 0 : -1;
}

// ================================================================================================
// extensions (include/spmv_b200.h)
// ================================================================================================
int spmv_b200_version(void) { return SPMV_B200_VERSION; }
const char *spmv_b200_last_error(void) { return g_err; }
void spmv_b200_clear_error(void) { g_err[0] = 0; }
unsigned long long spmv_b200_launch_count(void) { return g_launches.load(); }

void spmv_b200_set_stream(spmv_Handle_t handle, void *cuda_stream)
{
    if (DeviceState *st = state_of(handle)) {
        std::lock_guard<std::mutex> lock(st->mu);
        st->stream = (cudaStream_t)cuda_stream;
    }
}

void spmv_b200_sync(spmv_Handle_t handle)
{
    if (DeviceState *st = state_of(handle)) {
        DeviceGuard g(st->device);
        SB_CUDA(cudaStreamSynchronize(st->stream));
    }
}

int spmv_b200_set_y_peers(spmv_Handle_t handle, int count, void *const *device_ptrs)
{
    DeviceState *st = state_of(handle);
    if (!st || count < 0 || count > kMaxPeers || (count > 0 && !device_ptrs)) return -1;
    std::lock_guard<std::mutex> lock(st->mu);
    st->n_peers = count;
    for (int i = 0; i < kMaxPeers; ++i) st->peers[i] = i < count ? device_ptrs[i] : nullptr;
    return 0;
}

// ---- column stages: the contribution of a range of column bands, then the fold ----------------------
static bool stageable(const DeviceState *st)
{
    if (!st || !st->ok || st->x_bands <= 1) return false;
    if (st->kernel == SPMV_B200_KERNEL_BAND_SEG) return true;
    return st->kernel == SPMV_B200_KERNEL_CSR_VECTOR && !st->binned && st->lr_rows == 0;
}

int spmv_b200_bands(spmv_Handle_t handle)
{
    DeviceState *st = state_of(handle);
    if (!st || !st->ok) return -1;
    return stageable(st) ? st->x_bands : 1;
}

int spmv_b200_band_columns(spmv_Handle_t handle, int band, long long *col_lo, long long *col_hi)
{
    DeviceState *st = state_of(handle);
    if (!st || !st->ok || band < 0) return -1;
    const int K = stageable(st) ? st->x_bands : 1;
    if (band >= K) return -1;
    const long long lo = K == 1 ? 0 : (long long)band * st->band_cols;
    long long hi = K == 1 ? st->n : lo + st->band_cols;
    if (band == K - 1 || hi > st->n) hi = st->n;
    if (col_lo) *col_lo = lo < st->n ? lo : st->n;
    if (col_hi) *col_hi = hi;
    return 0;
}

int spmv_b200_spmv_bands(spmv_Handle_t handle, int band_first, int band_count, const void *x_device)
{
    DeviceState *st = state_of(handle);
    if (!stageable(st) || band_first < 0 || band_count < 0 || band_first + band_count > st->x_bands) return -1;
    if (band_count == 0) return 0;
    if (!is_device_ptr(x_device)) { set_error("spmv_b200_spmv_bands: x must be a device pointer"); return -1; }
    std::lock_guard<std::mutex> lock(st->mu);
    DeviceGuard guard(st->device);
    bool ok;
    if (st->kernel == SPMV_B200_KERNEL_BAND_SEG) {
        ok = st->vsize == 8 ? launch_seg_bands<double>(st, band_first, band_count, (const double *)x_device)
                            : launch_seg_bands<float>(st, band_first, band_count, (const float *)x_device);
    } else {
        const int r0 = band_first * st->m, r1 = (band_first + band_count) * st->m;  // virtual rows of these bands
        if (st->vsize == 8) {
            PeerList<double> none;
            none.n = 0;
            for (int i = 0; i < kMaxPeers; ++i) none.p[i] = nullptr;
            launch_vector_mode<double>(st, r0, r1, (const double *)x_device, (double *)st->v_y, none);
        } else {
            PeerList<float> none;
            none.n = 0;
            for (int i = 0; i < kMaxPeers; ++i) none.p[i] = nullptr;
            launch_vector_mode<float>(st, r0, r1, (const float *)x_device, (float *)st->v_y, none);
        }
        ok = SB_CUDA(cudaGetLastError());
    }
    return ok ? 0 : -1;
}

}  // extern "C"

template <typename T>
static bool finish_banded(DeviceState *st, T *y)
{
    PeerList<T> peers;
    peers.n = st->n_peers;
    for (int i = 0; i < kMaxPeers; ++i) peers.p[i] = (T *)st->peers[i];
    if (peers.n > 0) band_reduce_kernel<T, true><<<blocks_for(st->m), kThreads, 0, st->stream>>>(0, st->m, st->m, st->x_bands, (const T *)st->v_y, y, peers);
    else band_reduce_kernel<T, false><<<blocks_for(st->m), kThreads, 0, st->stream>>>(0, st->m, st->m, st->x_bands, (const T *)st->v_y, y, peers);
    count_launch();
    return SB_CUDA(cudaGetLastError());
}

extern "C" {

int spmv_b200_spmv_finish(spmv_Handle_t handle, void *y_device)
{
    DeviceState *st = state_of(handle);
    if (!stageable(st)) return -1;
    if (!is_device_ptr(y_device)) { set_error("spmv_b200_spmv_finish: y must be a device pointer"); return -1; }
    std::lock_guard<std::mutex> lock(st->mu);
    DeviceGuard guard(st->device);
    bool ok;
    if (st->kernel == SPMV_B200_KERNEL_BAND_SEG)
        ok = st->vsize == 8 ? launch_seg_finish<double>(st, (double *)y_device) : launch_seg_finish<float>(st, (float *)y_device);
    else
        ok = st->vsize == 8 ? finish_banded<double>(st, (double *)y_device) : finish_banded<float>(st, (float *)y_device);
    return ok ? 0 : -1;
}

// ---- stream-ordered building blocks of a copy-engine exchange (multigpu.py: CopyEnginePowerMethod) ----
int spmv_b200_memcpy_async(void *dst, const void *src, size_t bytes, void *cuda_stream)
{
    if (!bytes) return 0;
    return SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream)) ? 0 : -1;
}

typedef int (*StreamValueFn)(cudaStream_t, unsigned long long /*CUdeviceptr*/, unsigned, unsigned);
static StreamValueFn driver_fn(const char *name)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        set_error("driver entry point %s is not available", name);
        return nullptr;
    }
    return (StreamValueFn)fn;
}

int spmv_b200_stream_write32(void *cuda_stream, void *device_ptr, unsigned value)
{
    static StreamValueFn fn = driver_fn("cuStreamWriteValue32");
    if (!fn || !device_ptr) return -1;
    const int rc = fn((cudaStream_t)cuda_stream, (unsigned long long)(uintptr_t)device_ptr, value, 0 /*CU_STREAM_WRITE_VALUE_DEFAULT*/);
    if (rc != 0) { set_error("cuStreamWriteValue32 failed (%d)", rc); return -1; }
    return 0;
}

int spmv_b200_stream_wait32_geq(void *cuda_stream, void *device_ptr, unsigned value)
{
    static StreamValueFn fn = driver_fn("cuStreamWaitValue32");
    if (!fn || !device_ptr) return -1;
    const int rc = fn((cudaStream_t)cuda_stream, (unsigned long long)(uintptr_t)device_ptr, value, 0 /*CU_STREAM_WAIT_VALUE_GEQ*/);
    if (rc != 0) { set_error("cuStreamWaitValue32 failed (%d)", rc); return -1; }
    return 0;
}

int spmv_b200_ipc_export(const void *device_ptr, void *handle_out_64)
{
    if (!device_ptr || !handle_out_64) return -1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (!SB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(device_ptr)))) return -1;
    memcpy(handle_out_64, &h, sizeof(h));
    return 0;
}

void *spmv_b200_ipc_open(const void *handle_64)
{
    if (!handle_64) return nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, sizeof(h));
    void *p = nullptr;
    if (!SB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess))) return nullptr;
    return p;
}

int spmv_b200_ipc_close(void *opened_ptr)
{
    if (!opened_ptr) return -1;
    return SB_CUDA(cudaIpcCloseMemHandle(opened_ptr)) ? 0 : -1;
}

int spmv_b200_set_option(const char *key, long long value)
{
    Options &o = options();
    std::lock_guard<std::mutex> g(o.mu);
    auto it = o.v.find(key ? key : "");
    if (it == o.v.end()) return -1;
    it->second = value;
    o.user_set[key] = true;
    return 0;
}

long long spmv_b200_get_option(const char *key) { return key ? opt(key) : -1; }

long long spmv_b200_info(spmv_Handle_t handle, const char *key)
{
    DeviceState *st = state_of(handle);
    if (!st || !key) return -1;
    const std::string k(key);
    if (k == "kernel") return st->kernel;
    if (k == "requested") return st->requested;
    if (k == "ok") return st->ok;
    if (k == "m") return st->m;
    if (k == "n") return st->n;
    if (k == "nnz") return st->nnz;
    if (k == "tpr") return st->tpr;
    if (k == "parts") return st->parts;
    if (k == "ref_parts") return st->ref_T;
    if (k == "tiles") return st->tiles;
    if (k == "tile_items") return st->tile_items;
    if (k == "sigma") return st->sigma;
    if (k == "banner") return st->banner;
    if (k == "slices") return st->slices;
    if (k == "padded_nnz") return st->padded;
    if (k == "csr5_p") return st->c5_p;
    if (k == "csr5_sigma") return st->c5_sigma;
    if (k == "csr5_bit_y_offset") return st->c5_bit_y;
    if (k == "csr5_bit_scansum_offset") return st->c5_bit_ss;
    if (k == "csr5_num_offsets") return st->c5_num_offsets;
    if (k == "csr5_tail_start") return st->c5_tail_start;
    if (k == "pipeline") return st->pipeline;
    if (k == "values_snapshotted") return 1;
    if (k == "binned") return st->binned;
    if (k == "pinned_host_buffers") return (int)st->pin[0].registered + (int)st->pin[1].registered;
    if (k == "long_rows") return st->lr_rows;
    if (k == "long_segs") return st->lr_segs;
    if (k == "long_thr") return st->long_thr;
    if (k == "device") return st->device;
    if (k == "has_empty_rows") return st->has_empty_rows;
    if (k == "x_bands") return st->x_bands;
    if (k == "auto_method") return st->auto_method;
    if (k == "seg_bands") return st->coo_bands;
    if (k == "segments") return st->seg_total;
    if (k == "seg_cross") return st->seg_cross;
    if (k == "layout_fallbacks") return st->layout_fallbacks;
    if (k == "band_cols") return st->band_cols;
    if (k == "far_permille") return (long long)(st->far_fraction * 1000.0);
    if (k == "active_rows") return st->a_m;
    if (k == "owns_csr") return st->owns_csr;
    if (k == "released_csr") return st->released_csr;
    if (k == "dev_l2_bytes") return st->dev_l2;
    return -1;
}

long long spmv_b200_structure(spmv_Handle_t handle, const char *name, void *dst, size_t dst_bytes)
{
    DeviceState *st = state_of(handle);
    if (!st || !name) return -1;
    const std::string k(name);
    const void *src = nullptr;
    size_t bytes = 0;
    if (k == "splitter") { src = st->splitter; bytes = ((size_t)st->parts + 1) * 4; }
    else if (k == "ref_splitter") { src = st->ref_splitter; bytes = ((size_t)st->ref_T + 1) * 4; }
    else if (k == "tile_rows") { src = st->tile_rows; bytes = ((size_t)st->tiles + 1) * 4; }
    else if (k == "merge_coords") { src = st->merge_coords; bytes = ((size_t)st->tiles + 1) * 8; }
    else if (k == "sell_perm") { src = st->sell_perm; bytes = (size_t)st->banner * 4; }
    else if (k == "sell_width") { src = st->sell_width; bytes = (size_t)st->slices * 4; }
    else if (k == "sell_full") { src = st->sell_full; bytes = (size_t)st->slices * 4; }
    else if (k == "sell_slice_ptr") { src = st->sell_slice_ptr; bytes = ((size_t)st->slices + 1) * 8; }
    else if (k == "sell_col") { src = st->sell_col; bytes = (size_t)st->padded * 4; }
    else if (k == "sell_val") { src = st->sell_val; bytes = (size_t)st->padded * st->vsize; }
    else if (k == "csr5_tile_ptr") { src = st->c5_tile_ptr; bytes = ((size_t)st->c5_p + 1) * 4; }
    else if (k == "csr5_tile_desc") { src = st->c5_tile_desc; bytes = (size_t)st->c5_p * kC5Omega * 4; }
    else if (k == "csr5_offset_ptr") { src = st->c5_off_ptr; bytes = ((size_t)st->c5_p + 1) * 4; }
    else if (k == "csr5_offsets") { src = st->c5_off; bytes = (size_t)st->c5_num_offsets * 4; }
    else if (k == "csr5_col") { src = st->c5_col; bytes = (size_t)st->nnz * 4; }
    else if (k == "csr5_val") { src = st->c5_val; bytes = (size_t)st->nnz * st->vsize; }
    else if (k == "seg_col") { src = st->seg_col; bytes = st->seg_col ? (size_t)st->seg_ptr[st->coo_bands] * 4 : 0; }
    else if (k == "seg_mask") { src = st->seg_mask; bytes = st->seg_mask ? (size_t)st->m * (st->seg_mask64 ? 8 : 4) : 0; }
    else if (k == "seg_gbase") { src = st->seg_gbase; bytes = st->seg_gbase ? (size_t)st->coo_bands * st->seg_groups * 4 : 0; }
    else if (k == "seg_chunk_seg0") { src = st->seg_tile_seg0; bytes = st->seg_tile_seg0 ? ((size_t)st->tiles * kSegChunks + 1) * 4 : 0; }
    else if (k == "seg_ptr" || k == "seg_cnt") {
        const int *tab = k == "seg_ptr" ? st->seg_ptr : st->seg_cnt;
        const size_t need = (size_t)(st->coo_bands + (k == "seg_ptr" ? 1 : 0)) * 4;
        if (!dst) return st->coo_bands > 1 ? (long long)need : 0;
        if (st->coo_bands <= 1 || dst_bytes < need) return -1;
        memcpy(dst, tab, need);
        return (long long)need;
    }
    else if (k == "band_rowptr") { src = st->v_rowptr; bytes = st->v_rowptr ? ((size_t)st->a_m + 1) * 4 : 0; }
    else if (k == "band_col") { src = st->v_col; bytes = st->v_col ? (size_t)st->nnz * 4 : 0; }
    else return -1;
    if (!src) bytes = 0;
    if (!dst || bytes == 0) return (long long)bytes;
    if (dst_bytes < bytes) { set_error("structure %s needs %zu bytes, got %zu", name, bytes, dst_bytes); return -1; }
    DeviceGuard g(st->device);
    if (!SB_CUDA(cudaStreamSynchronize(st->stream))) return -1;
    if (!SB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost))) return -1;
    return (long long)bytes;
}

int spmv_b200_partition_rows(const int *RowPtr, int m, int parts, int *splitter_out)
{
    if (!RowPtr || !splitter_out || m < 0 || parts < 1) return -1;
    const int nnz = RowPtr[m] - RowPtr[0];
    const long long stride = ((long long)nnz + parts - 1) / parts;
    for (int g = 0; g <= parts; ++g) {
        long long b = (long long)g * stride;
        if (b > nnz) b = nnz;
        splitter_out[g] = right_boundary(RowPtr, (int)b, m + 1) - 1;
    }
    return 0;
}

int spmv_b200_recommend_method(int m, const int *RowPtr)
{
    if (!RowPtr || m < 0) return -1;
    if (m < 8192) return Method_Parallel;
    const long long nnz = (long long)RowPtr[m] - RowPtr[0];
    if (nnz <= 0) return Method_Parallel;
    const double mean = (double)nnz / m;
    const long long cut = (long long)(4.0 * mean) + 16;
    long long heavy = 0;  // non-zeros in rows much longer than the mean
    for (int r = 0; r < m; ++r) {
        const long long len = (long long)RowPtr[r + 1] - RowPtr[r];
        if (len > cut) heavy += len;
    }
    return (4 * heavy > nnz) ? Method_CSR5SPMV : Method_SellCSigma;
}

void *spmv_b200_malloc(size_t bytes)
{
    void *p = nullptr;
    if (!SB_CUDA(cudaMalloc(&p, bytes ? bytes : 1))) return nullptr;
    return p;
}

void spmv_b200_free(void *device_ptr) { dfree(device_ptr); }

int spmv_b200_memcpy(void *dst, const void *src, size_t bytes, int kind)
{
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    return SB_CUDA(cudaMemcpy(dst, src, bytes, k)) ? 0 : -1;
}

}  // extern "C"

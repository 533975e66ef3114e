"""Shared matrix cases (same generator parameters as tests/golden/make_golden.py)."""
import numpy as np

from spmv_b200 import matrices as M

GOLDEN_CASES = {
    "lap48": lambda: M.laplacian2d(48),
    "uni32": lambda: M.uniform_random(1500, 1500, 32, seed=M.SEED_C2),
    "uni5r": lambda: M.uniform_random(700, 300, 5, seed=99),
    "rmat10": lambda: M.rmat(10, 16, dtype=np.float64),
    "st27_9": lambda: M.stencil27(9),
    "skew": lambda: M.skewed(3000, 3000, max_len=1500),
    "longrow0": lambda: M.from_row_lengths([40, 1, 1, 1, 1], 50),
    "empties": lambda: M.from_row_lengths([0, 0, 0, 7, 0, 3, 0, 0, 13, 0], 20),
}
SPLIT_T = (1, 2, 7, 8, 64, 500)
SELL_NT = (1, 3, 11)

# extra shapes for the GPU parity tests: ragged, empty, extreme
EXTRA_CASES = {
    "one_row": lambda: M.from_row_lengths([17], 40),
    "one_long_row": lambda: M.from_row_lengths([20000], 5000),
    "all_empty": lambda: M.from_row_lengths([0] * 100, 10),
    "lead_trail_empty": lambda: M.from_row_lengths([0] * 70 + [3, 900, 0, 0, 5] + [0] * 90, 1000),
    "tiny_m3": lambda: M.from_row_lengths([2, 0, 1], 3),
    "len_sweep": lambda: M.from_row_lengths(list(range(0, 70)) * 3, 500),
    "hub": lambda: M.from_row_lengths([3] * 500 + [6000] + [2] * 700 + [2500, 0, 0, 1] + [4] * 300, 8000),
    "uni16": lambda: M.uniform_random(5000, 7000, 16, seed=5),
    "lap100x37": lambda: M.laplacian2d(100, 37),
    "rmat12": lambda: M.rmat(12, 8, dtype=np.float64),
    "exact2048": lambda: M.from_row_lengths([2048] * 8 + [1024] * 4, 3000),
}


def all_cases():
    d = dict(GOLDEN_CASES)
    d.update(EXTRA_CASES)
    return d


def expected_band_segments(a, bands, with_values=False):
    """numpy restatement of the band-segment layout (band_seg.cuh): entries stably bucketed by col // band_cols,
    band starts aligned to 4 slots, bit 31 of the column index marks the last entry of a (band, row) run."""
    bc = -(-a.n // bands)
    rows = np.repeat(np.arange(a.m, dtype=np.int64), np.diff(a.rowptr))
    band = np.minimum(a.col // bc, bands - 1)
    order = np.argsort(band, kind="stable")
    cnt = np.bincount(band, minlength=bands).astype(np.int64)
    ptr = np.zeros(bands + 1, dtype=np.int64)
    for b in range(bands):
        ptr[b + 1] = (ptr[b] + cnt[b] + 3) & ~3
    col = np.zeros(int(ptr[bands]), dtype=np.uint32)
    srows, sband = rows[order], band[order]
    last = np.ones(len(order), dtype=bool)
    if len(order) > 1:
        last[:-1] = (srows[1:] != srows[:-1]) | (sband[1:] != sband[:-1])
    sorted_start = np.concatenate([[0], np.cumsum(cnt)])
    for b in range(bands):
        lo, hi = int(sorted_start[b]), int(sorted_start[b + 1])
        col[int(ptr[b]):int(ptr[b]) + hi - lo] = a.col[order[lo:hi]].astype(np.uint32) | (last[lo:hi].astype(np.uint32) << 31)
    mask = np.zeros(a.m, dtype=np.uint64)
    np.bitwise_or.at(mask, rows, np.uint64(1) << band.astype(np.uint64))
    if with_values:
        val = np.zeros(int(ptr[bands]), dtype=a.val.dtype)
        for b in range(bands):
            lo, hi = int(sorted_start[b]), int(sorted_start[b + 1])
            val[int(ptr[b]):int(ptr[b]) + hi - lo] = a.val[order[lo:hi]]
        return bc, ptr.astype(np.int32), cnt.astype(np.int32), col, mask, int(last.sum()), val
    return bc, ptr.astype(np.int32), cnt.astype(np.int32), col, mask, int(last.sum())

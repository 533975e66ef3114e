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

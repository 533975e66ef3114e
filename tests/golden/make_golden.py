"""Generate tests/golden/reference_golden.npz from the UNMODIFIED reference (oracle/_ref/libmv_l2.so).

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these are outputs of the reference
itself on small deterministic inputs (spmv_b200.matrices generators, parameters recorded by name).
The fixture pins oracle/spmv_oracle.c on machines where the reference cannot be built (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from spmv_b200 import matrices as M  # noqa: E402

CASES = {
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


def main():
    R = O.Reference()
    out = {}
    for name, make in CASES.items():
        A = make()
        for dt, tag in ((np.float64, "d"), (np.float32, "s")):
            a = A.astype(dt)
            x = M.make_x(a.n, 1234, dt)
            out[f"{name}/y_serial_{tag}"] = R.serial(a.rowptr, a.col, a.val, x)
        for T in SPLIT_T:
            s, yid, meth = R.splitter(A.rowptr, T)
            out[f"{name}/splitter_T{T}"] = s
            out[f"{name}/yid_T{T}"] = yid
            out[f"{name}/method_T{T}"] = np.array([meth], np.int32)
            if A.m >= 64:  # the reference's Yid builder reads out of bounds on tiny inputs
                for k, v in R.splitter_yid(A.rowptr, T).items():
                    out[f"{name}/yidsplit_{k}_T{T}"] = v
        for nt in SELL_NT:
            sigma, banner, perm, widths, full = R.sell(A.rowptr, A.col, A.val, nt)
            out[f"{name}/sell_nt{nt}_sigma"] = np.array([sigma, banner], np.int32)
            out[f"{name}/sell_nt{nt}_perm"] = perm
            out[f"{name}/sell_nt{nt}_width"] = widths
        r5 = R.csr5(A.rowptr, A.col, A.val.astype(np.float64))
        for k in ("tile_ptr", "tile_desc", "offset_ptr", "offsets", "col_t"):
            out[f"{name}/csr5_{k}"] = r5[k]
        out[f"{name}/csr5_scalars"] = np.array([r5[k] for k in ("p", "bit_y_offset", "bit_scansum_offset",
                                                                 "num_packet", "tail_start")], np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

"""Matrix-Market ingestion and the reference driver's binary cache.

Mirrors the data formats on the input side of the path in the reference's sample driver:
  * ``read_mtx``  = mmio_allinone (reference src/samples/mmio_highlevel.h:325-491): coordinate file -> CSR,
    1-based -> 0-based, entries kept in FILE ORDER inside each row (no sorting), symmetric / hermitian
    files expanded by mirroring every off-diagonal entry, pattern files get value 1.0, complex files keep
    the real part, integer files are converted;
  * ``save_bin`` / ``read_bin`` = mmio_save_as_bin / mmio_read_from_bin (:531-584): raw
    ``int m, n, nnz; int rowptr[m+1]; int colidx[nnz]; double val[nnz]`` in ``mtx_cache/<name>.bin`` where
    <name> is the path with '/', '\\' and ' ' replaced by '_'.
Checked against files written by the reference's own driver in tests/test_mtx.py.
"""
from __future__ import annotations

import os

import numpy as np

from .matrices import CSR


def _banner(line: str):
    t = line.strip().lower().split()
    if len(t) < 5 or t[0] != "%%matrixmarket" or t[1] != "matrix" or t[2] != "coordinate":
        raise ValueError("not a Matrix-Market coordinate file: " + line.strip())
    return t[3], t[4]  # field, symmetry


def read_mtx(path: str, dtype=np.float64) -> CSR:
    with open(path, "r") as f:
        field, sym = _banner(f.readline())
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, nnz_file = (int(v) for v in line.split()[:3])
        body = np.loadtxt(f, ndmin=2, dtype=np.float64) if nnz_file else np.zeros((0, 3))
    if body.shape[0] != nnz_file:
        raise ValueError(f"{path}: expected {nnz_file} entries, found {body.shape[0]}")
    row = body[:, 0].astype(np.int64) - 1
    col = body[:, 1].astype(np.int64) - 1
    if field == "pattern":
        val = np.ones(nnz_file)
    else:  # real / integer / complex (real part, as the reference keeps only fval)
        val = body[:, 2].astype(np.float64)
    idx = np.arange(nnz_file, dtype=np.int64)
    if sym in ("symmetric", "hermitian"):
        off = row != col
        row, col, val, idx = (np.concatenate([row, col[off]]), np.concatenate([col, row[off]]),
                              np.concatenate([val, val[off]]), np.concatenate([idx, idx[off]]))
    order = np.lexsort((idx, row))  # by row, file order inside the row (the mirror of entry i stays at i)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=m), out=rowptr[1:])
    return CSR(m, n, rowptr.astype(np.int32), col[order].astype(np.int32), val[order].astype(dtype),
               os.path.basename(path))


def write_mtx(path: str, A: CSR, symmetric: bool = False) -> None:
    """General (or, for a structurally symmetric A, lower-triangular 'symmetric') real coordinate file."""
    rows = np.repeat(np.arange(A.m, dtype=np.int64), np.diff(A.rowptr))
    cols = A.col.astype(np.int64)
    vals = A.val.astype(np.float64)
    if symmetric:
        keep = rows >= cols
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real " + ("symmetric" if symmetric else "general") + "\n")
        f.write(f"{A.m} {A.n} {len(vals)}\n")
        np.savetxt(f, np.column_stack([rows + 1, cols + 1, vals]), fmt="%d %d %.17g")


def cache_path(mtx_path: str, root: str = ".") -> str:
    name = mtx_path.replace("/", "_").replace("\\", "_").replace(" ", "_")
    return os.path.join(root, "mtx_cache", name + ".bin")


def save_bin(A: CSR, mtx_path: str, root: str = ".") -> str:
    p = cache_path(mtx_path, root)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "wb") as f:
        np.array([A.m, A.n, A.nnz], dtype=np.int32).tofile(f)
        A.rowptr.astype(np.int32).tofile(f)
        A.col.astype(np.int32).tofile(f)
        A.val.astype(np.float64).tofile(f)
    return p


def read_bin(path: str) -> CSR:
    with open(path, "rb") as f:
        m, n, nnz = (int(v) for v in np.fromfile(f, dtype=np.int32, count=3))
        rowptr = np.fromfile(f, dtype=np.int32, count=m + 1)
        col = np.fromfile(f, dtype=np.int32, count=nnz)
        val = np.fromfile(f, dtype=np.float64, count=nnz)
    if len(rowptr) != m + 1 or len(col) != nnz or len(val) != nnz:
        raise ValueError(path + ": truncated cache file")
    return CSR(m, n, rowptr, col, val, os.path.basename(path))

"""Deterministic synthetic CSR generators (host / numpy side).

The five BASELINE.json matrix families, specified in SURVEY.md 8(d).  Every generator is a pure
function of (shape, seed) built on a counter-based hash (splitmix64 of ``seed + counter``), so that the
CUDA generators in csrc/generators.cu produce the SAME matrix bit for bit on the device (checked by
tests/test_gpu_generators.py) and no 6 GB matrix ever has to cross PCIe for a benchmark.

These play the role of the reference driver's input stage (src/samples/test_spmv.c:158-209 loads a .mtx,
overwrites the values with ``rand()%8*0.125`` and sets X=1); ``values="eighths"`` / ``x="ones"`` mirror
that protocol, the default values are the ones SURVEY.md 8(d) fixes per family.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

U64 = np.uint64
_GOLDEN = U64(0x9E3779B97F4A7C15)
_M1 = U64(0xBF58476D1CE4E5B9)
_M2 = U64(0x94D049BB133111EB)
VAL_SALT = U64(0x5A17ED0000000000)
X_SALT = U64(0x0C0FFEE000000000)

SEED_C2 = 0x5EED0002
SEED_C3 = 0x5EED0003
SEED_C5 = 0x5EED0005


def splitmix64(z):
    """One splitmix64 output for counter(s) ``z`` (uint64, wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(z, dtype=U64) + _GOLDEN).astype(U64)
        z = ((z ^ (z >> U64(30))) * _M1).astype(U64)
        z = ((z ^ (z >> U64(27))) * _M2).astype(U64)
        return (z ^ (z >> U64(31))).astype(U64)


@dataclass
class CSR:
    m: int
    n: int
    rowptr: np.ndarray  # int32 [m+1]
    col: np.ndarray     # int32 [nnz]
    val: np.ndarray     # float64 / float32 [nnz]
    name: str = ""

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])

    def min_bytes(self) -> int:
        """B_min of BASELINE.md: nnz*(val+idx) + (m+1)*idx + m*val + n*val."""
        v = self.val.dtype.itemsize
        return self.nnz * (v + 4) + (self.m + 1) * 4 + self.m * v + self.n * v

    def astype(self, dtype) -> "CSR":
        return CSR(self.m, self.n, self.rowptr, self.col, self.val.astype(dtype), self.name)


def min_bytes(m: int, n: int, nnz: int, vsize: int) -> int:
    return nnz * (vsize + 4) + (m + 1) * 4 + m * vsize + n * vsize


# --------------------------------------------------------------------------------------------
# values / x
# --------------------------------------------------------------------------------------------
def hashed_values(nnz: int, seed: int, dtype, eighths: bool = False, offset: int = 0) -> np.ndarray:
    """Position-keyed values.  default: ((h & 7) + 1) / 8  in {0.125 .. 1.0};
    ``eighths``: (h & 7) / 8 in {0 .. 0.875}, the reference driver's rand()%8*0.125 (test_spmv.c:200)."""
    with np.errstate(over="ignore"):
        h = splitmix64(U64(seed) + VAL_SALT + np.arange(offset, offset + nnz, dtype=U64))
    k = (h & U64(7)).astype(np.float64)
    return ((k if eighths else k + 1.0) * 0.125).astype(dtype)


def make_x(n: int, seed: int, dtype, kind: str = "hashed") -> np.ndarray:
    """x_j = 0.5 + (h(j) mod 1000)/1000 or all ones (test_spmv.c:201-202)."""
    if kind == "ones":
        return np.ones(n, dtype=dtype)
    with np.errstate(over="ignore"):
        h = splitmix64(U64(seed) + X_SALT + np.arange(n, dtype=U64))
    return (0.5 + (h % U64(1000)).astype(np.float64) / 1000.0).astype(dtype)


# --------------------------------------------------------------------------------------------
# C1: 5-point 2-D Laplacian, C4: 27-point 3-D stencil  (row-major grid order, Dirichlet truncation)
# --------------------------------------------------------------------------------------------
def _stencil(dims, offsets, diag, dtype, name):
    dims = tuple(int(d) for d in dims)
    m = int(np.prod(dims))
    coords = np.unravel_index(np.arange(m, dtype=np.int64), dims)
    cols, ok = [], []
    for off in offsets:  # offsets are listed in ascending linear order -> columns ascending per row
        valid = np.ones(m, dtype=bool)
        lin = np.zeros(m, dtype=np.int64)
        for c, o, d in zip(coords, off, dims):
            cc = c + o
            valid &= (cc >= 0) & (cc < d)
            lin = lin * d + cc
        cols.append(lin)
        ok.append(valid)
    cols, ok = np.stack(cols, 1), np.stack(ok, 1)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(ok.sum(1), out=rowptr[1:])
    is_diag = np.array([all(o == 0 for o in off) for off in offsets])
    vals = np.where(is_diag[None, :], float(diag), -1.0)
    vals = np.broadcast_to(vals, ok.shape)
    return CSR(m, m, rowptr.astype(np.int32), cols[ok].astype(np.int32), vals[ok].astype(dtype), name)


def laplacian2d(nx: int, ny: int | None = None, dtype=np.float64) -> CSR:
    """C1 family.  Grid point (i, j) -> row i*ny + j; entries (i-1,j),(i,j-1),(i,j),(i,j+1),(i+1,j)."""
    ny = nx if ny is None else ny
    offs = [(-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)]
    return _stencil((nx, ny), offs, 4.0, dtype, f"laplacian2d_{nx}x{ny}")


def stencil27(nx: int, ny: int | None = None, nz: int | None = None, dtype=np.float64) -> CSR:
    """C4 family.  Lexicographic order, truncated at the faces; diag 26, off-diagonal -1."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
    return _stencil((nx, ny, nz), offs, 26.0, dtype, f"stencil27_{nx}x{ny}x{nz}")


# --------------------------------------------------------------------------------------------
# C2 / C5: uniform random, exactly k entries per row, sorted, duplicates kept
# --------------------------------------------------------------------------------------------
def uniform_random(m: int, n: int, k: int, seed: int = SEED_C2, dtype=np.float64, row0: int = 0,
                   eighths: bool = False) -> CSR:
    """Rows [row0, row0+m) of the global matrix: column of (row r, slot j) is
    (splitmix64(seed + r*k + j) >> 11) mod n, sorted ascending within the row; the value at sorted
    position j of row r is keyed by the global position r*k + j."""
    with np.errstate(over="ignore"):
        ctr = (U64(row0) * U64(k) + np.arange(m * k, dtype=U64)).astype(U64)
        h = splitmix64(U64(seed) + ctr)
    col = ((h >> U64(11)) % U64(n)).astype(np.int32).reshape(m, k)
    col.sort(axis=1)
    rowptr = (np.arange(m + 1, dtype=np.int64) * k).astype(np.int32)
    val = hashed_values(m * k, seed, dtype, eighths, offset=row0 * k)
    return CSR(m, n, rowptr, col.reshape(-1), val, f"uniform_{m}x{n}_k{k}")


# --------------------------------------------------------------------------------------------
# C3: R-MAT (Graph500 a,b,c,d = .57,.19,.19,.05), duplicates kept, rows sorted by column
# --------------------------------------------------------------------------------------------
def rmat(scale: int, edge_factor: int = 16, seed: int = SEED_C3, dtype=np.float32,
         a: float = 0.57, b: float = 0.19, c: float = 0.19) -> CSR:
    m = 1 << scale
    e = m * edge_factor
    row = np.zeros(e, dtype=np.int64)
    colv = np.zeros(e, dtype=np.int64)
    eid = np.arange(e, dtype=U64)
    for lvl in range(scale):
        with np.errstate(over="ignore"):
            h = splitmix64(U64(seed) + eid * U64(scale) + U64(lvl))
        r = (h >> U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        rb = (r >= a + b).astype(np.int64)                     # quadrants c,d -> lower half (row bit)
        cb = (((r >= a) & (r < a + b)) | (r >= a + b + c)).astype(np.int64)  # quadrants b,d -> right half
        row = (row << 1) | rb
        colv = (colv << 1) | cb
    key = np.sort((row << 32) | colv)
    row, colv = key >> 32, key & 0xFFFFFFFF
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=m), out=rowptr[1:])
    val = hashed_values(e, seed, dtype)
    return CSR(m, m, rowptr.astype(np.int32), colv.astype(np.int32), val, f"rmat_s{scale}_ef{edge_factor}")


# --------------------------------------------------------------------------------------------
# small irregular shapes for edge-case tests (not BASELINE configs)
# --------------------------------------------------------------------------------------------
def from_row_lengths(lengths, n: int, seed: int = 7, dtype=np.float64, name="custom") -> CSR:
    """Arbitrary row-length profile; columns hashed, sorted within the row."""
    lengths = np.asarray(lengths, dtype=np.int64)
    m = len(lengths)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lengths, out=rowptr[1:])
    nnz = int(rowptr[-1])
    with np.errstate(over="ignore"):
        h = splitmix64(U64(seed) + np.arange(nnz, dtype=U64))
    col = ((h >> U64(11)) % U64(max(n, 1))).astype(np.int64)
    rows = np.repeat(np.arange(m, dtype=np.int64), lengths)
    order = np.lexsort((col, rows))
    col = col[order].astype(np.int32)
    val = hashed_values(nnz, seed, dtype)
    return CSR(m, n, rowptr.astype(np.int32), col, val, name)


def skewed(m: int, n: int, seed: int = 11, dtype=np.float64, max_len: int = 5000, empty_frac: float = 0.3) -> CSR:
    """Power-law-ish row lengths with a share of empty rows and a few very long rows."""
    with np.errstate(over="ignore"):
        h = splitmix64(U64(seed) + U64(0xABCDEF) + np.arange(m, dtype=U64))
    u = (h >> U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    lens = np.minimum((1.0 / np.maximum(u, 1e-9)) ** 0.9, max_len).astype(np.int64)
    lens[u > 1.0 - empty_frac] = 0
    return from_row_lengths(lens, n, seed, dtype, name=f"skewed_{m}x{n}")

"""Matrix-Market ingestion and the `mtx_cache/*.bin` format, pinned against files
written by the reference's own sample driver (oracle/_ref/test_spmv_ref = src/samples/test_spmv.c built
unmodified against the reference library: it parses the .mtx with mmio_allinone and saves the CSR with
mmio_save_as_bin, reference src/samples/mmio_highlevel.h:325-491,531-584)."""
import os
import subprocess

import numpy as np
import pytest

from spmv_b200 import matrices as M, mtx
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_raw(path, banner, m, n, lines):
    with open(path, "w") as f:
        f.write(banner + "\n% a comment line\n")
        f.write(f"{m} {n} {len(lines)}\n")
        f.write("\n".join(lines) + "\n")


def _shuffled_general(path):
    """General real matrix whose entries are NOT sorted in the file (CSR keeps file order inside a row)."""
    A = M.uniform_random(40, 55, 6, seed=3)
    rows = np.repeat(np.arange(A.m), np.diff(A.rowptr))
    perm = np.random.default_rng(7).permutation(A.nnz)
    lines = [f"{rows[k] + 1} {A.col[k] + 1} {A.val[k]:.17g}" for k in perm]
    _write_raw(path, "%%MatrixMarket matrix coordinate real general", A.m, A.n, lines)


def _symmetric_real(path):
    mtx.write_mtx(path, M.laplacian2d(12), symmetric=True)


def _pattern_symmetric(path):
    rng = np.random.default_rng(11)
    pairs = {(int(max(a, b)), int(min(a, b))) for a, b in rng.integers(0, 30, size=(120, 2))}
    lines = [f"{r + 1} {c + 1}" for r, c in sorted(pairs, key=lambda p: (p[1], p[0]))]
    _write_raw(path, "%%MatrixMarket matrix coordinate pattern symmetric", 30, 30, lines)


def _integer_general(path):
    rng = np.random.default_rng(13)
    lines = [f"{rng.integers(1, 21)} {rng.integers(1, 18)} {rng.integers(-9, 10)}" for _ in range(90)]
    _write_raw(path, "%%MatrixMarket matrix coordinate integer general", 20, 17, lines)


def _complex_hermitian(path):
    rng = np.random.default_rng(17)
    lines = []
    for r in range(15):
        for c in range(r + 1):
            if rng.random() < 0.3:
                lines.append(f"{r + 1} {c + 1} {rng.standard_normal():.17g} {rng.standard_normal():.17g}")
    _write_raw(path, "%%MatrixMarket matrix coordinate complex hermitian", 15, 15, lines)


CASES = {"general_shuffled": _shuffled_general, "symmetric_real": _symmetric_real,
         "pattern_symmetric": _pattern_symmetric, "integer_general": _integer_general,
         "complex_hermitian": _complex_hermitian}


@pytest.mark.skipif(not os.path.exists(O.DRIVER_REF), reason="oracle/_ref/test_spmv_ref not built")
@pytest.mark.parametrize("name", list(CASES))
def test_read_mtx_equals_reference_mmio(tmp_path, port, name):
    """read_mtx == what the reference driver parsed (bit for bit: rowptr, colidx in file order, values)."""
    path = "m.mtx"
    CASES[name](str(tmp_path / path))
    os.makedirs(tmp_path / "mtx_cache")
    r = subprocess.run([O.DRIVER_REF, path, "1", "1"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    ref = mtx.read_bin(mtx.cache_path(path, str(tmp_path)))
    ours = mtx.read_mtx(str(tmp_path / path))
    assert (ours.m, ours.n, ours.nnz) == (ref.m, ref.n, ref.nnz)
    assert np.array_equal(ours.rowptr, ref.rowptr)
    assert np.array_equal(ours.col, ref.col)
    assert np.array_equal(ours.val.view(np.uint64), ref.val.view(np.uint64))
    # the CSV of the reference run: 6 methods x 1 thread count, error column 0 (eighths with x = 1 are exact)
    rows = [l.split(",") for l in r.stdout.strip().splitlines()]
    assert [x[1] for x in rows] == ["Method_Parallel", "Method_Balanced", "Method_Balanced2", "Method_BalancedYid",
                                    "Method_SellCSigma", "Method_Csr5Spmv"]
    assert all(int(x[4]) == ref.nnz for x in rows)


def test_bin_cache_round_trip(tmp_path):
    A = M.skewed(300, 200, max_len=90)
    p = mtx.save_bin(A, "dir with space/a b.mtx", str(tmp_path))
    assert os.path.basename(p) == "dir_with_space_a_b.mtx.bin"  # '/', '\\', ' ' -> '_' (mmio_highlevel.h:536-541)
    B = mtx.read_bin(p)
    assert (B.m, B.n, B.nnz) == (A.m, A.n, A.nnz)
    assert np.array_equal(B.rowptr, A.rowptr) and np.array_equal(B.col, A.col) and np.array_equal(B.val, A.val)
    with open(p, "rb") as f:
        raw = f.read()
    assert len(raw) == 12 + 4 * (A.m + 1) + 12 * A.nnz
    with open(p, "wb") as f:
        f.write(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        mtx.read_bin(p)


def test_write_read_mtx_round_trip(tmp_path):
    A = M.rmat(8, 6, dtype=np.float64)
    mtx.write_mtx(str(tmp_path / "g.mtx"), A)
    B = mtx.read_mtx(str(tmp_path / "g.mtx"))
    assert np.array_equal(B.rowptr, A.rowptr) and np.array_equal(B.col, A.col) and np.array_equal(B.val, A.val)
    with pytest.raises(ValueError, match="Matrix-Market"):
        (tmp_path / "bad.mtx").write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
        mtx.read_mtx(str(tmp_path / "bad.mtx"))

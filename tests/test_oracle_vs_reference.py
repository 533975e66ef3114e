"""Pin the CPU restatement (oracle/spmv_oracle.c) against the UNMODIFIED reference compiled from
/root/reference (oracle/_ref/libmv_l2.so) and against the committed fixtures made from it.  CPU only."""
import numpy as np
import pytest

from conftest import bits_equal
from cases import GOLDEN_CASES, SPLIT_T, SELL_NT
from spmv_b200 import matrices as M


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_port_matches_golden_fixture(port, golden, name):
    A = GOLDEN_CASES[name]()
    for dt, tag in ((np.float64, "d"), (np.float32, "s")):
        a = A.astype(dt)
        x = M.make_x(a.n, 1234, dt)
        assert bits_equal(port.spmv_serial(a.rowptr, a.col, a.val, x), golden[f"{name}/y_serial_{tag}"])
    for T in SPLIT_T:
        s = port.splitter(A.rowptr, T)
        assert np.array_equal(s, golden[f"{name}/splitter_T{T}"])
        use_bal, yid = port.balanced2_yid(s, A.m)
        assert np.array_equal(yid, golden[f"{name}/yid_T{T}"])
        assert use_bal == (int(golden[f"{name}/method_T{T}"][0]) == 2)
        if A.m >= 64:
            for k, v in port.splitter_yid(A.rowptr, T).items():
                assert np.array_equal(v, golden[f"{name}/yidsplit_{k}_T{T}"]), (T, k)
    for nt in SELL_NT:
        sigma, banner = golden[f"{name}/sell_nt{nt}_sigma"]
        perm = port.sell_perm(A.rowptr, int(sigma))
        assert len(perm) == banner and np.array_equal(perm, golden[f"{name}/sell_nt{nt}_perm"])
        if banner:
            w, _ = port.sell_chunks(A.rowptr, perm, 4)
            assert np.array_equal(w, golden[f"{name}/sell_nt{nt}_width"])
    p5 = port.csr5(A.rowptr, 4, 16, A.col)
    sc = golden[f"{name}/csr5_scalars"]
    assert [p5[k] for k in ("p", "bit_y_offset", "bit_scansum_offset", "num_packet", "tail_start")] == list(sc)
    for k in ("tile_ptr", "tile_desc", "offset_ptr", "col_t"):
        assert np.array_equal(p5[k], golden[f"{name}/csr5_{k}"]), k
    _check_offsets(p5["offsets"], golden[f"{name}/csr5_offsets"], p5["offset_ptr"])


def _check_offsets(mine, theirs, off_ptr):
    """The reference never writes the last offset slot of a dirty tile (it counts the forced first bit of
    lane 0 but skips it when filling, format_avx2.h:292) -- that slot is uninitialised malloc memory there
    and 0 here; every other slot must agree."""
    assert len(mine) == len(theirs)
    keep = np.ones(len(mine), bool)
    last = off_ptr[1:][np.diff(off_ptr) > 0] - 1
    keep[last] = False
    assert np.array_equal(mine[keep], theirs[keep])


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_serial_rowlen_sweep(port, ref, dt):
    """Every row length 0..40 with signed random data: bit-exact y, including the compiled remainder
    behaviour documented in oracle/spmv_oracle.c."""
    rng = np.random.default_rng(3)
    for L in range(0, 41):
        A = M.from_row_lengths([L] * 64, 200, dtype=dt)
        A.val[:] = rng.standard_normal(A.nnz).astype(dt)
        x = rng.standard_normal(200).astype(dt)
        assert bits_equal(ref.serial(A.rowptr, A.col, A.val, x), port.spmv_serial(A.rowptr, A.col, A.val, x)), L


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_port_matches_live_reference(port, ref, name):
    A = GOLDEN_CASES[name]()
    for dt in (np.float64, np.float32):
        a = A.astype(dt)
        x = M.make_x(a.n, 77, dt)
        assert bits_equal(ref.serial(a.rowptr, a.col, a.val, x), port.spmv_serial(a.rowptr, a.col, a.val, x))
    for T in (1, 3, 16, 100, 1000):
        s, yid, meth = ref.splitter(A.rowptr, T)
        sp = port.splitter(A.rowptr, T)
        assert np.array_equal(s, sp)
        ub, yp = port.balanced2_yid(sp, A.m)
        assert np.array_equal(yid, yp) and ub == (meth == 2)
    if A.m >= 256:
        # sigma = 4*floor(m/nthreads/4) (common.c:139-140): whatever window the reference picks, same perm
        sigma, banner, perm, w, _ = ref.sell(A.rowptr, A.col, A.val, A.m // 256)
        assert 256 <= sigma < 512 and banner == sigma * (A.m // sigma)
        assert np.array_equal(perm, port.sell_perm(A.rowptr, sigma))
    r5 = ref.csr5(A.rowptr, A.col, A.val.astype(np.float64))
    p5 = port.csr5(A.rowptr, 4, 16, A.col)
    for k in ("tile_ptr", "tile_desc", "offset_ptr", "col_t"):
        assert np.array_equal(r5[k], p5[k]), k
    _check_offsets(p5["offsets"], r5["offsets"], p5["offset_ptr"])


def test_reference_methods_agree_with_serial(ref):
    """The reference's own parallel methods (fp64) against its Method_Serial on a regular matrix -- the
    sanity anchor for using them as the CPU timing baseline."""
    A = M.laplacian2d(64)
    x = M.make_x(A.n, 5, np.float64)
    y0 = ref.serial(A.rowptr, A.col, A.val, x)
    s = np.abs(A.val).max() * 5 * np.abs(x).max()
    for method in range(1, 7):
        h = ref.create(A.m, A.n, A.rowptr, A.col, A.val, 4, method)
        y = ref.spmv(h, x)
        h.destroy()
        assert np.abs(y - y0).max() <= 8 * np.finfo(np.float64).eps * s, method


def test_scalar_golden_protocol(port):
    """The sample driver's protocol (test_spmv.c:199-207): eighths values, X = 1 -> every sum is exact, so
    Method_Serial order and scalar CSR order agree bit for bit."""
    A = M.uniform_random(2000, 2000, 32, eighths=True)
    x = M.make_x(A.n, 0, np.float64, kind="ones")
    assert bits_equal(port.spmv_serial(A.rowptr, A.col, A.val, x), port.scalar_golden(A.rowptr, A.col, A.val, x))

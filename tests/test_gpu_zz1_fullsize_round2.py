"""Full-size parity cases added in round 2 (same checker as tests/test_gpu_fullsize.py).  Kept near the end of the GPU
suite (file name) because they were written after the round's GPU budget was spent and have not run on a GPU yet."""
import numpy as np
import pytest

from conftest import bits_equal
from spmv_b200 import api, matrices as M

pytestmark = pytest.mark.gpu


def test_full_size_parity_on_the_c5_shard(libpath):
    """One GPU's share of C5 at 8 GPUs: rows [0, 2^25) of the 2^28-column matrix, 2^29 non-zeros, x = 2 GiB -- the one
    shape on which the AUTOMATIC band-segment layout (48 column bands, 64-bit row masks) and 2^28-wide column indices
    are exercised; every method against the torch fp64 evaluation + reproducibility + homogeneity."""
    from test_gpu_fullsize import check_full_size
    check_full_size("c5shard", lambda: (api.gen_uniform(1 << 25, 1 << 28, 16, M.SEED_C5, 0, False, 8), M.SEED_C5))


def test_full_size_c1_against_the_live_reference(libpath, port):
    """BASELINE.json configs[0] at its full size (5-point Laplacian on a 1024 x 1024 grid, fp64): Method_Serial on
    the GPU is bit-identical to the compiled reference's Method_Serial; every other method meets the per-row bound
    against it and against the extended-precision sum; fp32 likewise."""
    from oracle import oracle as O
    a64 = M.laplacian2d(1024)
    for dt in (np.float64, np.float32):
        a = a64.astype(dt)
        x = M.make_x(a.n, 1, dt)
        if O.have_reference():
            y_ref = O.Reference().serial(a.rowptr, a.col, a.val, x)
        else:
            y_ref = port.spmv_serial(a.rowptr, a.col, a.val, x)
        y_ex = port.spmv_exact(a.rowptr, a.col, a.val, x)
        eps = np.finfo(dt).eps
        tol = 8 * eps * port.row_abs_sum(a.rowptr, a.col, a.val, x)
        for method in range(7):
            h = api.Handle(a.m, a.n, a.rowptr, a.col, a.val, method, nthreads=8)
            y = np.full(a.m, np.nan, dtype=dt)
            h.spmv(x, y)
            tag = f"c1/{dt.__name__}/{api.METHOD_NAMES[method]}[{h.kernel}]"
            if method == api.Method_Serial:
                assert bits_equal(y, y_ref), tag
            assert (np.abs(y.astype(np.float64) - y_ref.astype(np.float64)) <= tol).all(), tag
            assert (np.abs(y.astype(np.float64) - y_ex.astype(np.float64)) <= tol + 0.5 * eps * np.abs(y_ex)).all(), tag
            y2 = np.full(a.m, np.nan, dtype=dt)
            h.spmv(x, y2)
            assert bits_equal(y, y2), tag
            h.destroy()

"""the drop-in acceptance run.  The reference's sample driver src/samples/test_spmv.c,
compiled UNMODIFIED against include/spmv.h + libspmv_b200.so (oracle/Makefile `drivers`; the binary is
built in the development container and travels to the GPU box), is run on Matrix-Market files next to the
same driver linked against the reference library.  Both must print the reference's CSV schema
(matrix,method,vectorized,threads,nnz,err,pre_ms,avg_ms,GFLOPS_avg,GFLOPS_best -- test_spmv.c:146-149) with
the same method rows and nnz and a zero error column (the driver overwrites the values with exact eighths
and uses x = 1, test_spmv.c:199-207, so every summation order gives the same bits)."""
import os
import subprocess

import numpy as np
import pytest

from spmv_b200 import matrices as M, mtx
from oracle import oracle as O

pytestmark = pytest.mark.gpu

METHOD_ROWS = ["Method_Parallel", "Method_Balanced", "Method_Balanced2", "Method_BalancedYid",
               "Method_SellCSigma", "Method_Csr5Spmv"]


def _run(exe, cwd, path, t0, t1):
    r = subprocess.run([exe, path, str(t0), str(t1)], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = [l.split(",") for l in r.stdout.strip().splitlines() if l.count(",") == 9]
    return rows, r.stderr


@pytest.fixture(scope="module")
def drivers(libpath):
    if os.path.isdir("/root/reference"):
        O.build(ref=True)
    if not (os.path.exists(O.DRIVER_B200) and os.path.exists(O.DRIVER_REF)):
        pytest.skip("oracle/_ref/test_spmv_{b200,ref} not built (run `make -C oracle drivers` where /root/reference exists)")
    return O.DRIVER_B200, O.DRIVER_REF


@pytest.mark.parametrize("name,make,sym", [
    ("lap40", lambda: M.laplacian2d(40), True),
    ("uni", lambda: M.uniform_random(3000, 3000, 12, seed=21), False),
    ("skew", lambda: M.skewed(2500, 2500, max_len=1200), False),
    ("empties", lambda: M.from_row_lengths([0, 0, 5, 0, 9, 1, 0, 0, 33, 0, 0, 2] * 40, 500), False),
])
def test_unmodified_sample_driver_runs_on_the_gpu_library(tmp_path, drivers, name, make, sym):
    ours_exe, ref_exe = drivers
    A = make()
    path = name + ".mtx"
    mtx.write_mtx(str(tmp_path / path), A, symmetric=sym)
    os.makedirs(tmp_path / "mtx_cache")
    ref_rows, _ = _run(ref_exe, tmp_path, path, 1, 2)       # also writes mtx_cache/<name>.bin ...
    our_rows, err = _run(ours_exe, tmp_path, path, 1, 2)    # ... which our run then loads (mmio_read_from_bin)
    assert "[spmv_b200]" not in err, err
    assert len(our_rows) == len(ref_rows) == 12            # 6 methods x threads {1, 2}
    for o, r in zip(our_rows, ref_rows):
        assert o[0] == r[0] == path
        assert o[1] == r[1] and o[1] in METHOD_ROWS        # Methods_names[] strings
        assert o[2] == r[2] == "VECTOR_AVX2"               # Vectorized_names[] strings
        assert o[3] == r[3]                                 # threads column
        assert int(o[4]) == int(r[4]) == A.nnz              # nnz column
        assert float(o[5]) == 0.0, o                        # err column: exact on the GPU ...
        for v in o[6:]:
            assert np.isfinite(float(v))
    # ... and on the reference wherever the reference itself is sound (its CSR5 leaves empty rows unwritten
    # and Balanced2 double-counts a long first row)
    if name in ("lap40", "uni"):
        assert all(float(r[5]) == 0.0 for r in ref_rows)
    assert [o[1] for o in our_rows[::2]] == METHOD_ROWS

"""Automatic method selection inside create (option "auto"; the empty README.md:222 heading of the
reference): the rule is deterministic in the matrix statistics, the result keeps parity, and on the full-size
BASELINE.json configurations the pick is (close to) the fastest of the methods it chooses between.  Runs last: the
timing assertion is the only one in the suite that depends on measured speed."""
import numpy as np
import pytest

from conftest import bits_equal
from spmv_b200 import api, matrices as M

pytestmark = pytest.mark.gpu


def _auto_handle(a, method, **kw):
    api.set_option("auto", 1)
    try:
        return api.Handle(a.m, a.n, a.rowptr, a.col, a.val, method, **kw)
    finally:
        api.set_option("auto", 0)


def test_auto_pick_follows_the_rule_and_keeps_parity(libpath, port):
    cases = [
        ("small", M.laplacian2d(48), api.Method_Parallel),                                  # < 8192 rows
        ("short_local", M.laplacian2d(128), api.Method_Parallel),                           # mean 5, diagonal-local
        ("regular_32", M.uniform_random(20000, 20000, 32, seed=3), api.Method_SellCSigma),  # mean 32
        ("power_law", M.from_row_lengths([2] * 20000 + [30000], 40000), api.Method_CSR5SPMV),  # 43 % of nnz in one row
    ]
    for name, a, want in cases:
        x = M.make_x(a.n, 7, np.float64)
        y_ex = port.spmv_exact(a.rowptr, a.col, a.val, x)
        tol = 8 * np.finfo(np.float64).eps * port.row_abs_sum(a.rowptr, a.col, a.val, x) + 0.5 * np.finfo(np.float64).eps * np.abs(y_ex)
        for asked in (api.Method_Parallel, api.Method_Balanced2, api.Method_CSR5SPMV):
            h = _auto_handle(a, asked)
            assert h.info("auto_method") == want and h.struct.spmvMethod == want, (name, asked, h.info("auto_method"))
            y = np.full(a.m, np.nan)
            h.spmv(x, y)
            assert (np.abs(y - y_ex) <= tol).all(), (name, asked)
            h.destroy()
        # Method_Serial promises the reference's bits and is never overridden
        h = _auto_handle(a, api.Method_Serial)
        assert h.info("auto_method") == -1 and h.kernel == "csr_reforder" and h.struct.spmvMethod == api.Method_Serial
        y = np.full(a.m, np.nan)
        h.spmv(x, y)
        assert bits_equal(y, port.spmv_serial(a.rowptr, a.col, a.val, x)), name
        h.destroy()
    # without the option nothing is overridden
    a = cases[2][1]
    h = api.Handle(a.m, a.n, a.rowptr, a.col, a.val, api.Method_Parallel)
    assert h.info("auto_method") == -1 and h.kernel == "csr_vector"
    h.destroy()


FULL = {
    "c2": lambda: (api.gen_uniform(1 << 24, 1 << 24, 32, M.SEED_C2, 0, False, 8), M.SEED_C2, api.Method_SellCSigma),
    "c3": lambda: (api.gen_rmat(24, 16, M.SEED_C3, 4), M.SEED_C3, api.Method_CSR5SPMV),
    "c4": lambda: (api.gen_stencil27(256, 256, 256, 8), 4, api.Method_SellCSigma),
}


@pytest.mark.parametrize("name", list(FULL))
def test_auto_pick_is_close_to_the_fastest_method_at_full_size(libpath, name):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a B200-class device")
    A, seed, want = FULL[name]()
    tdt = torch.float64 if A.size == 8 else torch.float32
    x = torch.empty(A.n, dtype=tdt, device="cuda")
    api.gen_x(x, A.n, seed, False, A.size)
    y = torch.empty(A.m, dtype=tdt, device="cuda")

    def ms_per_call(h):
        for _ in range(5):
            h.spmv(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            h.spmv(x, y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    api.set_option("auto", 1)
    try:
        h = A.handle(api.Method_Parallel)
    finally:
        api.set_option("auto", 0)
    picked = h.info("auto_method")
    assert picked == want, (name, picked)
    t_pick = ms_per_call(h)
    h.destroy()
    times = {}
    for method in (api.Method_Parallel, api.Method_SellCSigma, api.Method_CSR5SPMV):
        h = A.handle(method)
        times[method] = ms_per_call(h)
        h.destroy()
    A.destroy()
    best = min(times.values())
    # measured in round 2 (profiles/r02a_bench_c2_n1.json): the pick IS the fastest of the three on C2, C3 and C4, by
    # 8 % or more; 15 % of slack absorbs run-to-run noise
    assert t_pick <= 1.15 * best, (name, api.METHOD_NAMES[picked], t_pick, {api.METHOD_NAMES[k]: v for k, v in times.items()})

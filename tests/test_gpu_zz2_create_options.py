"""The two create-time options that sit above the hot path.

Option "auto": automatic method selection inside create (the empty README.md:222 heading of the reference) -- the rule
is deterministic in the matrix statistics, the result keeps parity, and on the full-size BASELINE.json configurations
the pick is (close to) the fastest of the methods it chooses between.
Option "reorder": the reference's compiled-out level-3 hook (common.c:144-156) -- create fills handle->index with a
reverse Cuthill-McKee permutation and builds the layout of P A P^T; the caller follows the reference's protocol
(test_spmv.c:95-101,130-137), which the UNMODIFIED sample driver does by itself.

Runs last (file name): these tests were written after round 2's GPU budget was spent and have not run on a GPU yet;
the timing assertion is the only one in the suite that depends on measured speed."""
import ctypes as C
import os
import subprocess

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


# ------------------------------------------------------------------------------------------------
# option "reorder"
# ------------------------------------------------------------------------------------------------
def _scrambled_laplacian(g, seed):
    A = M.laplacian2d(g)
    q = np.random.default_rng(seed).permutation(A.m).astype(np.int32)
    rp, ci, va = api.permute_csr(A.rowptr, A.col, A.val, q)
    return M.CSR(A.m, A.n, rp, ci, va, f"lap{g}_scrambled")


def _reorder_handle(a, method):
    api.set_option("reorder", 1)
    try:
        return api.Handle(a.m, a.n, a.rowptr, a.col, a.val, method, nthreads=4)
    finally:
        api.set_option("reorder", 0)


def test_reorder_option_fills_handle_index_and_the_reference_protocol_gives_y(libpath, port):
    a = _scrambled_laplacian(160, 11)                      # 25 600 rows, 127 360 non-zeros: above the reference's thresholds
    want = api.reorder(a.rowptr, a.col)
    x = M.make_x(a.n, 4, np.float64)
    y_ref = port.spmv_serial(a.rowptr, a.col, a.val, x)
    tol = 8 * np.finfo(np.float64).eps * port.row_abs_sum(a.rowptr, a.col, a.val, x)
    for method in (api.Method_Serial, api.Method_Parallel, api.Method_Balanced2, api.Method_SellCSigma, api.Method_CSR5SPMV):
        h = _reorder_handle(a, method)
        s = h.struct
        assert s.Level_3_opt_used == 1 and s.index, api.METHOD_NAMES[method]
        index = np.ctypeslib.as_array(C.cast(s.index, C.POINTER(C.c_int)), shape=(a.m,)).copy()
        assert np.array_equal(index, want)
        xx = x[index]                                       # test_spmv.c:95-98
        yy = np.full(a.m, np.nan)
        h.spmv(xx, yy)
        y = np.empty_like(yy)
        y[index] = yy                                       # test_spmv.c:130-133
        assert (np.abs(y - y_ref) <= tol).all(), api.METHOD_NAMES[method]
        h.destroy()
    # below the thresholds, or without the option: the handle is the plain one
    b = M.laplacian2d(48)
    h = _reorder_handle(b, api.Method_Parallel)
    assert h.struct.Level_3_opt_used == 0 and not h.struct.index
    h.destroy()
    h = api.Handle(a.m, a.n, a.rowptr, a.col, a.val, api.Method_Parallel)
    assert h.struct.Level_3_opt_used == 0 and not h.struct.index
    h.destroy()


def test_unmodified_sample_driver_takes_its_index_branch(libpath, tmp_path):
    """SPMV_B200_REORDER=1 in the environment of the reference's UNMODIFIED test_spmv.c: the driver sees handle->index,
    permutes x, scatters y back and compares with its own golden vector -- the error column must stay 0."""
    from oracle import oracle as O
    from spmv_b200 import mtx
    if os.path.isdir("/root/reference"):
        O.build(ref=True)
    if not (os.path.exists(O.DRIVER_B200) and os.path.exists(O.DRIVER_REF)):
        pytest.skip("oracle/_ref/test_spmv_{b200,ref} not built")
    a = _scrambled_laplacian(160, 5)
    mtx.write_mtx(str(tmp_path / "scr.mtx"), a, symmetric=False)
    os.makedirs(tmp_path / "mtx_cache")
    subprocess.run([O.DRIVER_REF, "scr.mtx", "1", "1"], cwd=tmp_path, capture_output=True, text=True, timeout=600, check=True)
    env = dict(os.environ, SPMV_B200_REORDER="1")
    r = subprocess.run([O.DRIVER_B200, "scr.mtx", "2", "2"], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = [l.split(",") for l in r.stdout.strip().splitlines() if l.count(",") == 9]
    assert len(rows) == 6 and "[spmv_b200]" not in r.stderr, (r.stdout, r.stderr)
    for o in rows:
        assert int(o[4]) == a.nnz and float(o[5]) == 0.0, o


# ------------------------------------------------------------------------------------------------
# spmv_b200_update_values: new values on a fixed pattern
# ------------------------------------------------------------------------------------------------
def test_update_values_rebuilds_the_layout_in_place(libpath, port):
    """The values are part of the device layout (spmv() ignores Matrix_Val, unlike the reference); after changing them
    the client calls spmv_b200_update_values().  Checked for layouts that copy the values (SELL, CSR5, band-major,
    band segments) and for the plain CSR kernels."""
    a = M.uniform_random(6000, 6000, 24, seed=9)
    x = M.make_x(a.n, 3, np.float64)
    new_val = (a.val * 0.5 + 0.25).copy()
    want = port.spmv_exact(a.rowptr, a.col, new_val, x)
    tol = 8 * np.finfo(np.float64).eps * port.row_abs_sum(a.rowptr, a.col, new_val, x) + 0.5 * np.finfo(np.float64).eps * np.abs(want)
    for method, opt in ((api.Method_Parallel, None), (api.Method_SellCSigma, None), (api.Method_CSR5SPMV, None),
                        (api.Method_Parallel, ("x_bands", 3)), (api.Method_Balanced2, ("seg_bands", 5))):
        if opt:
            api.set_option(*opt)
        try:
            h = api.Handle(a.m, a.n, a.rowptr, a.col, a.val.copy(), method)
            kernel = h.kernel
            y0 = np.full(a.m, np.nan)
            h.spmv(x, y0)
            h.update_values(new_val)                      # (the option is still set: the rebuilt layout is the same kind)
        finally:
            if opt:
                api.set_option(opt[0], 0)
        assert h.kernel == kernel, (method, opt)
        y = np.full(a.m, np.nan)
        h.spmv(x, y)
        assert (np.abs(y - want) <= tol).all(), (api.METHOD_NAMES[method], opt)
        assert not np.array_equal(y, y0)
        h.destroy()

"""SURVEY.md 8(f)-4, the host half: the permutation create stores in handle->index with option "reorder" (reverse
Cuthill-McKee) and the permuted matrix A' = P A P^T, checked on the CPU: valid permutation, bandwidth actually
shrinks on a scrambled mesh, A' is exactly the permuted operator, and the reference's calling protocol
(x'[i] = x[index[i]], y[index[i]] = y'[i]; src/samples/test_spmv.c:95-101,130-137) reproduces y = A x."""
import numpy as np
import pytest
import scipy.sparse as sp

from spmv_b200 import api, matrices as M


def _scramble(A, seed):
    """B = Q A Q^T for a random permutation Q (a mesh whose numbering has been destroyed)."""
    rng = np.random.default_rng(seed)
    q = rng.permutation(A.m).astype(np.int32)
    rp, ci, va = api.permute_csr(A.rowptr, A.col, A.val, q)
    return M.CSR(A.m, A.n, rp, ci, va, A.name + "_scrambled")


def _bandwidth(a):
    rows = np.repeat(np.arange(a.m), np.diff(a.rowptr))
    return int(np.abs(rows - a.col).max()) if a.nnz else 0


def test_permute_csr_is_the_permuted_operator(libpath):
    A = M.uniform_random(300, 300, 7, seed=5)            # duplicates inside rows included
    rng = np.random.default_rng(1)
    index = rng.permutation(A.m).astype(np.int32)
    rp, ci, va = api.permute_csr(A.rowptr, A.col, A.val, index)
    S = sp.csr_matrix((A.val, A.col, A.rowptr), shape=(A.m, A.n))       # (sums duplicates)
    S2 = sp.csr_matrix((va, ci, rp), shape=(A.m, A.n))
    P = sp.csr_matrix((np.ones(A.m), (np.arange(A.m), index)), shape=(A.m, A.m))   # row i of P picks row index[i]
    assert abs(S2 - P @ S @ P.T).max() == 0.0
    for i in range(A.m):                                  # columns ascending inside every row
        assert (np.diff(ci[rp[i]:rp[i + 1]]) >= 0).all()
    assert np.array_equal(np.diff(rp), np.diff(A.rowptr)[index])
    # not a permutation / not square: refused
    bad = index.copy()
    bad[0] = bad[1]
    with pytest.raises(ValueError):
        api.permute_csr(A.rowptr, A.col, A.val, bad)
    R = M.uniform_random(50, 80, 3, seed=2)
    with pytest.raises(ValueError):
        api.permute_csr(R.rowptr, R.col, R.val, np.arange(50, dtype=np.int32))


@pytest.mark.parametrize("make", [lambda: M.laplacian2d(64), lambda: M.stencil27(12)])
def test_rcm_restores_the_locality_of_a_scrambled_mesh(libpath, make):
    A = make()
    B = _scramble(A, 7)
    index = api.reorder(B.rowptr, B.col)
    assert np.array_equal(np.sort(index), np.arange(B.m))
    rp, ci, va = api.permute_csr(B.rowptr, B.col, B.val, index)
    C = M.CSR(B.m, B.n, rp, ci, va)
    bw_natural, bw_scrambled, bw_rcm = _bandwidth(A), _bandwidth(B), _bandwidth(C)
    assert bw_scrambled > 10 * bw_natural                 # the scrambling really destroyed the numbering
    assert bw_rcm <= 3 * bw_natural, (bw_natural, bw_scrambled, bw_rcm)


def test_reference_protocol_on_the_reordered_matrix_reproduces_y(libpath, port):
    A = _scramble(M.laplacian2d(40), 3)
    index = api.reorder(A.rowptr, A.col)
    rp, ci, va = api.permute_csr(A.rowptr, A.col, A.val, index)
    x = M.make_x(A.n, 9, np.float64)
    y_ref = port.spmv_serial(A.rowptr, A.col, A.val, x)
    xx = x[index]                                          # test_spmv.c:95-98
    yy = port.spmv_serial(rp, ci, va, xx)
    y = np.empty_like(yy)
    y[index] = yy                                          # test_spmv.c:130-133
    tol = 8 * np.finfo(np.float64).eps * port.row_abs_sum(A.rowptr, A.col, A.val, x)
    assert (np.abs(y - y_ref) <= tol).all()


def test_reorder_handles_unsymmetric_patterns_empty_rows_and_components(libpath):
    A = M.from_row_lengths([0, 3, 0, 0, 5, 1, 0, 2] * 50, 400)     # unsymmetric, many empty rows, disconnected
    index = api.reorder(A.rowptr, A.col)
    assert np.array_equal(np.sort(index), np.arange(A.m))
    rp, ci, va = api.permute_csr(A.rowptr, A.col, A.val, index)
    assert rp[-1] == A.nnz and np.array_equal(np.sort(va), np.sort(A.val))
    Z = M.from_row_lengths([0] * 10, 10)
    assert np.array_equal(np.sort(api.reorder(Z.rowptr, Z.col)), np.arange(10))


def test_create_with_option_reorder_fills_the_public_fields_even_without_a_gpu(libpath):
    """The host half of create runs before anything touches the device: on a box without a GPU the handle ends up
    unusable (`ok` == 0, no CPU fallback), but handle->index / Level_3_opt_used are already what the reference's
    protocol expects, and clear / destroy release the permutation."""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("the GPU half is covered by tests/test_gpu_zz2_create_options.py")
    a = _scramble(M.laplacian2d(160), 11)                 # 25 600 rows, 127 360 non-zeros: above the reference's thresholds
    want = api.reorder(a.rowptr, a.col)
    api.set_option("reorder", 1)
    try:
        h = api.spmv_create_handle_all_in_one(a.m, a.n, a.rowptr, a.col, a.val, 4, api.Method_Parallel, 8)
        small = M.laplacian2d(48)
        h2 = api.spmv_create_handle_all_in_one(small.m, small.n, small.rowptr, small.col, small.val, 4, api.Method_Parallel, 8)
    finally:
        api.set_option("reorder", 0)
    api.clear_error()
    s = h.contents
    assert api.lib().spmv_b200_info(h, b"ok") == 0
    assert s.Level_3_opt_used == 1 and s.index
    index = np.ctypeslib.as_array(C.cast(s.index, C.POINTER(C.c_int)), shape=(a.m + 1,)).copy()
    assert np.array_equal(index[:a.m], want) and index[a.m] == a.m
    assert h2.contents.Level_3_opt_used == 0 and not h2.contents.index     # below the thresholds: untouched
    api.spmv_clear_handle(h)
    assert h.contents.Level_3_opt_used == 0 and not h.contents.index
    api.spmv_destory_handle(h)
    api.spmv_destory_handle(h2)
    api.clear_error()


def test_update_values_on_an_unusable_handle_fails_cleanly(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("the working case is covered by tests/test_gpu_zz2_create_options.py")
    a = M.laplacian2d(20)
    h = api.spmv_create_handle_all_in_one(a.m, a.n, a.rowptr, a.col, a.val, 1, api.Method_SellCSigma, 8)
    assert api.lib().spmv_b200_update_values(h, a.val.ctypes.data) == -1      # no device: nothing to refresh, nothing touched
    assert api.lib().spmv_b200_info(h, b"requested") == api.Method_SellCSigma and api.lib().spmv_b200_info(h, b"ok") == 0
    assert api.lib().spmv_b200_update_values(None, None) == -1
    api.spmv_destory_handle(h)
    api.clear_error()

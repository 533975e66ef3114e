"""Parity at BASELINE.json's FULL sizes, where the CPU oracle would take minutes: every SPMV_METHODS kernel on
the device-generated C2 / C3 / C4 matrices against an independent plain-PyTorch evaluation of the same
operator (per-row segment sums of val * x[col] in the value type), with the north-star bound
|y - y_ref| <= 8*eps*sum_j|a_ij x_j| per row, plus size-independent properties: bitwise reproducibility,
exact homogeneity under a power-of-two scaling of x (every rounding commutes with *2), and -- because the small
cases prove Method_Serial bit-identical to the reference -- agreement of every method with Method_Serial."""
import numpy as np
import pytest

from spmv_b200 import api, matrices as M

pytestmark = pytest.mark.gpu


class _DevView:
    """Zero-copy torch view of a raw device pointer (through __cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(torch, ptr, n, typestr):
    return torch.as_tensor(_DevView(ptr, n, typestr), device="cuda")


def _row_sums(torch, data, lens):
    try:
        return torch.segment_reduce(data, "sum", lengths=lens, unsafe=True)
    except Exception:  # older builds: scatter-add (atomics; accurate to far below the tolerance)
        rows = torch.repeat_interleave(torch.arange(len(lens), device=data.device), lens)
        return torch.zeros(len(lens), dtype=data.dtype, device=data.device).index_add_(0, rows, data)


CONFIGS = {
    "c2": lambda: (api.gen_uniform(1 << 24, 1 << 24, 32, M.SEED_C2, 0, False, 8), M.SEED_C2),
    "c3": lambda: (api.gen_rmat(24, 16, M.SEED_C3, 4), M.SEED_C3),
    "c4": lambda: (api.gen_stencil27(256, 256, 256, 8), 4),
}


def check_full_size(name, make):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a B200-class device (full-size matrices + torch temporaries)")
    A, seed = make()
    tdt = torch.float64 if A.size == 8 else torch.float32
    vstr = "<f8" if A.size == 8 else "<f4"
    eps = float(torch.finfo(tdt).eps)
    x = torch.empty(A.n, dtype=tdt, device="cuda")
    api.gen_x(x, A.n, seed, False, A.size)
    rp = _view(torch, A.RowPtr, A.m + 1, "<i4")
    col = _view(torch, A.ColIdx, A.nnz, "<i4")
    val = _view(torch, A.Val, A.nnz, vstr)
    lens = (rp[1:] - rp[:-1]).long()
    prod = val.double() * x[col.long()].double()            # exact products (fp32 inputs) / fp64 products
    y_ref = _row_sums(torch, prod, lens)                     # fp64 accumulation: the reference value
    S = _row_sums(torch, prod.abs_(), lens)
    del prod
    torch.cuda.empty_cache()
    tol = 8 * eps * S + 0.5 * eps * y_ref.abs()              # + half an ulp for rounding the sum to the value type
    long_rows = lens > 1024                                  # reference-order chains there carry ~sqrt(len)*eps (R-MAT hubs)
    y_serial = None
    for method in range(7):
        h = A.handle(method)
        y = torch.full((A.m,), float("nan"), dtype=tdt, device="cuda")
        h.spmv(x, y)
        h.sync()
        tag = f"{name}/{api.METHOD_NAMES[method]}[{h.kernel}]"
        if name == "c5shard" and method != api.Method_Serial:
            assert h.kernel == "band_seg" and h.info("seg_bands") == 48 and h.info("layout_fallbacks") == 0, tag
        err = (y.double() - y_ref).abs()
        bad = err > tol
        if method == api.Method_Serial:
            bad &= ~long_rows
            y_serial = y.clone()
        assert not bool(bad.any()), f"{tag}: {int(bad.sum())} rows beyond 8*eps*sum|a x|, worst {float((err / (S + 1e-300)).max() / eps):.1f} eps"
        # agreement with Method_Serial (bit-identical to the reference's own output, proven on the small cases)
        d = (y.double() - y_serial.double()).abs()
        ok = (d <= 2 * tol) | long_rows
        assert bool(ok.all()), f"{tag}: differs from Method_Serial beyond the bound"
        # bitwise reproducible
        y2 = torch.empty_like(y)
        h.spmv(x, y2)
        h.sync()
        assert torch.equal(y, y2), tag
        # exact homogeneity: A(2x) == 2 A(x) bit for bit
        h.spmv(x * 2, y2)
        h.sync()
        assert torch.equal(y2, y * 2), tag
        h.destroy()
    A.destroy()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_size_parity_against_torch_and_properties(libpath, name):
    check_full_size(name, CONFIGS[name])

"""The drop-in boundary, checked without a GPU: the C-ABI library loads, exports every symbol that
include/spmv.h and include/spmv_b200.h declare, the public struct has the reference's layout, the name
tables carry the reference's strings, and the reference's own sample driver compiles and links against
our headers + library unchanged."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cases import GOLDEN_CASES, SPLIT_T
from spmv_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
GCC = "/usr/bin/gcc"


def test_library_loads_and_exports_everything(libpath):
    L = C.CDLL(libpath)
    for name in api.EXPORTED_FUNCTIONS + api.EXPORTED_DATA:
        assert hasattr(L, name), name
    assert L.spmv_b200_version() == 200


def test_every_header_declaration_is_in_the_export_list():
    declared = set()
    for h in ("spmv.h", "spmv_b200.h"):
        src = open(os.path.join(INC, h)).read()
        declared |= set(re.findall(r"SPMV_B200_API[^;(]*?\b(\w+)\s*\(", src))
    assert declared == set(api.EXPORTED_FUNCTIONS)


def test_name_tables_match_reference_strings(libpath):
    L = C.CDLL(libpath)
    methods = (C.c_char_p * 7).in_dll(L, "Methods_names")
    assert [m.decode() for m in methods] == api.METHOD_NAMES  # reference common.c:325-334
    vec = (C.c_char_p * 3).in_dll(L, "Vectorized_names")
    assert [v.decode() for v in vec] == ["VECTOR_NONE", "VECTOR_AVX2", "VECTOR_AVX512"]
    fn = (C.c_char_p * 18).in_dll(L, "funcNames")
    assert fn[0] == b"Method_Serial_VECTOR_NONE" and fn[17] == b"Method_Csr5Spmv_VECTOR_AVX512"


def test_handle_struct_layout_matches_c(tmp_path):
    """sizeof / offsetof from a C compile of OUR header == the ctypes mirror == the reference header."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "spmv.h"
int main(void){ printf("%zu %zu %zu %zu %zu %d %d\n", sizeof(spmv_Handle), offsetof(spmv_Handle,data_size),
  offsetof(spmv_Handle,RowPtr), offsetof(spmv_Handle,extraHandle), offsetof(spmv_Handle,index),
  (int)Method_Total_Size, (int)Method_Numa); return 0; }'''
    outs = []
    incs = [INC] + (["/root/reference/include"] if os.path.isdir("/root/reference/include") else [])
    for inc in incs:
        src = tmp_path / "t.c"
        src.write_text(prog)
        exe = tmp_path / "t"
        subprocess.run([GCC, "-mavx2", "-fopenmp", "-I", inc, str(src), "-o", str(exe)], check=True)
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert all(o == outs[0] for o in outs)
    size, off_ds, off_rp, off_ex, off_idx, total, numa = map(int, outs[0])
    H = api.spmv_Handle
    assert size == C.sizeof(H) == 80
    assert (off_ds, off_rp, off_ex, off_idx) == (H.data_size.offset, H.RowPtr.offset, H.extraHandle.offset, H.index.offset)
    assert (total, numa) == (api.Method_Total_Size, api.Method_Numa) == (7, 8)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/samples"), reason="reference tree not present")
def test_reference_sample_driver_builds_against_our_library(tmp_path, libpath):
    """src/samples/test_spmv.c, UNMODIFIED, compiles against include/spmv.h and links
    against libspmv_b200.so (running it needs a GPU: tests/test_gpu_dropin.py)."""
    exe = tmp_path / "test_spmv_b200"
    cmd = [GCC, "-O2", "-fopenmp", "-mavx2", "-w", "-I", INC, "-I", "/root/reference/src/samples",
           "/root/reference/src/samples/test_spmv.c", "-o", str(exe),
           "-L", os.path.dirname(libpath), "-lspmv_b200", "-lm", "-Wl,-rpath," + os.path.dirname(libpath)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert exe.exists()


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_partition_rows_is_the_reference_splitter(libpath, golden, port, name):
    """The multi-GPU row partition is host code in the library: bit-exact against a9 outputs of the
    reference (fixture) and of the port."""
    A = GOLDEN_CASES[name]()
    for T in SPLIT_T:
        s = api.partition_rows(A.rowptr, T)
        assert np.array_equal(s, golden[f"{name}/splitter_T{T}"])
        assert np.array_equal(s, port.splitter(A.rowptr, T))


def test_recommend_method_by_row_length_distribution(libpath):
    """regular matrices -> SELL, power-law -> CSR5, small -> Parallel (host code)."""
    from spmv_b200 import matrices as M
    assert api.recommend_method(M.laplacian2d(128).rowptr) == api.Method_SellCSigma
    assert api.recommend_method(M.uniform_random(20000, 20000, 32, seed=1).rowptr) == api.Method_SellCSigma
    assert api.recommend_method(M.stencil27(24).rowptr) == api.Method_SellCSigma
    assert api.recommend_method(M.rmat(15, 16, dtype=np.float32).rowptr) == api.Method_CSR5SPMV
    assert api.recommend_method(M.skewed(20000, 20000, max_len=9000).rowptr) == api.Method_CSR5SPMV
    assert api.recommend_method(M.laplacian2d(48).rowptr) == api.Method_Parallel
    assert api.recommend_method(np.zeros(9000, dtype=np.int32)) == api.Method_Parallel  # no non-zeros


def test_no_device_is_a_loud_error_not_a_fallback(libpath):
    """Without a GPU, create must fail with a message and spmv must leave y untouched."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    A = GOLDEN_CASES["lap48"]()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        api.Handle(A.m, A.n, A.rowptr, A.col, A.val, api.Method_Parallel)
    h = api.spmv_create_handle_all_in_one(A.m, A.n, A.rowptr, A.col, A.val, 1, api.Method_Parallel, 8)
    y = np.full(A.m, 123.0)
    api.spmv(h, A.m, A.rowptr, A.col, A.val, np.ones(A.n), y)
    assert (y == 123.0).all()
    api.spmv_destory_handle(h)
    api.spmv(None, A.m, A.rowptr, A.col, A.val, np.ones(A.n), y)  # NULL handle: silent no-op (common.c:285)


def test_extension_header_is_plain_c_and_the_new_entry_points_are_null_safe(libpath, tmp_path):
    """include/spmv_b200.h compiles as pedantic C11 (a C client of the reference is C), links against the library,
    and the round-2 entry points (band-staged SpMV, copy-engine exchange primitives) refuse NULL arguments instead
    of crashing -- with or without a GPU."""
    prog = r'''
#include <stdio.h>
#include "spmv_b200.h"
int main(void) {
    long long lo = -1, hi = -1;
    spmv_Handle_t h = NULL;
    int r = 0;
    r |= (spmv_b200_bands(h) != -1) << 0;
    r |= (spmv_b200_band_columns(h, 0, &lo, &hi) != -1) << 1;
    r |= (spmv_b200_spmv_bands(h, 0, 1, NULL) != -1) << 2;
    r |= (spmv_b200_spmv_finish(h, NULL) != -1) << 3;
    r |= (spmv_b200_memcpy_async(NULL, NULL, 0, NULL) != 0) << 4;
    r |= (spmv_b200_stream_write32(NULL, NULL, 1u) != -1) << 5;
    r |= (spmv_b200_stream_wait32_geq(NULL, NULL, 1u) != -1) << 6;
    r |= (spmv_b200_set_y_peers(h, 0, NULL) != -1) << 7;
    printf("%d %d\n", spmv_b200_version(), r);
    return 0;
}'''
    src = tmp_path / "ext.c"
    src.write_text(prog)
    exe = tmp_path / "ext"
    libdir = os.path.dirname(libpath)
    subprocess.run([GCC, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{INC}", str(src), f"-L{libdir}", "-lspmv_b200",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["200", "0"], out

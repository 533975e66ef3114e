"""ctypes front-end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY (see spmv_oracle.c header).

* ``Port``      -- oracle/liboracle.so, our plain-C restatement (oracle/spmv_oracle.c).
* ``Reference`` -- oracle/_ref/libmv_l2.so, the UNMODIFIED reference compiled by oracle/Makefile from
                   /root/reference (present in the build container; the built .so travels to the GPU
                   box, the sources do not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmv_l2.so")

# SPMV_METHODS / VECTORIZED_WAY of reference include/spmv_Defines.h:18-35
METHOD_SERIAL, METHOD_PARALLEL, METHOD_BALANCED, METHOD_BALANCED2, METHOD_BALANCED_YID, \
    METHOD_SELLCSIGMA, METHOD_CSR5, METHOD_TOTAL = range(8)
VECTOR_NONE, VECTOR_AVX2, VECTOR_AVX512 = range(3)

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the port (always) and the reference (when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
        subprocess.run(["make", "-s", "-C", HERE, "drivers"], check=True)


DRIVER_REF = os.path.join(HERE, "_ref", "test_spmv_ref")
DRIVER_B200 = os.path.join(HERE, "_ref", "test_spmv_b200")


def have_reference() -> bool:
    return os.path.exists(REF_SO)


class Port:
    """The restatement.  All arrays are numpy; outputs are returned."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = self.lib = C.CDLL(PORT_SO)
        for name, vp in (("d", _f64p), ("s", _f32p)):
            for fn in ("oracle_spmv_serial_", "oracle_spmv_parallel_", "oracle_spmv_exact_"):
                f = getattr(L, fn + name)
                f.argtypes = [C.c_int, _i32p, _i32p, vp, vp, vp]
                f.restype = None
            f = getattr(L, "oracle_row_abs_sum_" + name)
            f.argtypes = [C.c_int, _i32p, _i32p, vp, vp, _f64p]
            f.restype = None
        L.oracle_spmv_scalar_golden_d.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.oracle_right_boundary.argtypes = [_i32p, C.c_int, C.c_int]
        L.oracle_right_boundary.restype = C.c_int
        L.oracle_splitter_balanced2.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p]
        L.oracle_balanced2_yid.argtypes = [C.c_int, C.c_int, _i32p, _i32p]
        L.oracle_balanced2_yid.restype = C.c_int
        L.oracle_splitter_yid.argtypes = [C.c_int, C.c_int, C.c_int, _i32p] + [_i32p] * 6
        L.oracle_sell_perm.argtypes = [C.c_int, _i32p, C.c_int, _i32p]
        L.oracle_sell_perm.restype = C.c_int
        L.oracle_sell_chunks.argtypes = [_i32p, _i32p, C.c_int, C.c_int, _i32p, _i32p]
        L.oracle_csr5_params.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 4
        L.oracle_csr5_tile_ptr.argtypes = [C.c_int] * 5 + [_i32p, _u32p]
        L.oracle_csr5_tile_desc.argtypes = [C.c_int] * 7 + [_i32p, _u32p, _u32p, _i32p]
        L.oracle_csr5_tile_desc.restype = C.c_int
        L.oracle_csr5_desc_offset.argtypes = [C.c_int] * 6 + [_i32p, _u32p, _u32p, _i32p, _i32p]
        L.oracle_csr5_transpose_i32.argtypes = [C.c_int] * 3 + [_u32p, _i32p, _i32p]

    # -- y ------------------------------------------------------------------------------------
    def spmv_serial(self, rowptr, col, val, x, parallel=False):
        m = len(rowptr) - 1
        y = np.empty(m, dtype=val.dtype)
        sfx = "d" if val.dtype == np.float64 else "s"
        fn = "oracle_spmv_parallel_" if parallel else "oracle_spmv_serial_"
        getattr(self.lib, fn + sfx)(m, rowptr, _pad(col), _pad(val), x, y)
        return y

    def spmv_exact(self, rowptr, col, val, x):
        m = len(rowptr) - 1
        y = np.empty(m, dtype=val.dtype)
        sfx = "d" if val.dtype == np.float64 else "s"
        getattr(self.lib, "oracle_spmv_exact_" + sfx)(m, rowptr, _pad(col), _pad(val), x, y)
        return y

    def scalar_golden(self, rowptr, col, val, x):
        y = np.empty(len(rowptr) - 1, dtype=np.float64)
        self.lib.oracle_spmv_scalar_golden_d(len(y), rowptr, _pad(col), _pad(val), x, y)
        return y

    def row_abs_sum(self, rowptr, col, val, x):
        m = len(rowptr) - 1
        s = np.empty(m, dtype=np.float64)
        sfx = "d" if val.dtype == np.float64 else "s"
        getattr(self.lib, "oracle_row_abs_sum_" + sfx)(m, rowptr, _pad(col), _pad(val), x, s)
        return s

    # -- structures ---------------------------------------------------------------------------
    def splitter(self, rowptr, T):
        m = len(rowptr) - 1
        out = np.empty(T + 1, dtype=np.int32)
        self.lib.oracle_splitter_balanced2(T, int(rowptr[m] - rowptr[0]), m, rowptr, out)
        return out

    def balanced2_yid(self, splitter, m):
        T = len(splitter) - 1
        yid = np.empty(T, dtype=np.int32)
        use_balanced = self.lib.oracle_balanced2_yid(T, m, splitter, yid)
        return bool(use_balanced), yid

    def splitter_yid(self, rowptr, T):
        m = len(rowptr) - 1
        a = [np.empty(2 * T, dtype=np.int32)] + [np.empty(T, dtype=np.int32) for _ in range(5)]
        self.lib.oracle_splitter_yid(T, int(rowptr[m] - rowptr[0]), m, rowptr, *a)
        return dict(zip(("splitter", "Type", "Brow", "beginIdx", "Erow", "endIdx"), a))

    def sell_perm(self, rowptr, sigma):
        m = len(rowptr) - 1
        perm = np.empty(max(m, 1), dtype=np.int32)
        banner = self.lib.oracle_sell_perm(m, rowptr, sigma, perm)
        return perm[:banner].copy()

    def sell_chunks(self, rowptr, perm, Cc):
        n = len(perm) // Cc
        w = np.empty(max(n, 1), dtype=np.int32)
        f = np.empty(max(n, 1), dtype=np.int32)
        self.lib.oracle_sell_chunks(rowptr, perm if len(perm) else np.zeros(1, np.int32), len(perm), Cc, w, f)
        return w[:n].copy(), f[:n].copy()

    def csr5(self, rowptr, omega, sigma, col=None):
        """tile_ptr, tile_desc, offset pointer, offsets (+ transposed col) for (omega, sigma)."""
        m = len(rowptr) - 1
        nnz = int(rowptr[m] - rowptr[0])
        by, bs, npk, p = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.oracle_csr5_params(omega, sigma, nnz, by, bs, npk, p)
        by, bs, npk, p = by.value, bs.value, npk.value, p.value
        tile_ptr = np.zeros(p + 1, dtype=np.uint32)
        desc = np.zeros(max(p * omega * npk, 1), dtype=np.uint32)
        off_ptr = np.zeros(p + 1, dtype=np.int32)
        out = dict(p=p, bit_y_offset=by, bit_scansum_offset=bs, num_packet=npk)
        if p == 0:
            out.update(tile_ptr=tile_ptr, tile_desc=desc[:0], offset_ptr=off_ptr, offsets=np.zeros(0, np.int32),
                       tail_start=0)
            return out
        self.lib.oracle_csr5_tile_ptr(omega, sigma, p, m, nnz, rowptr, tile_ptr)
        nof = self.lib.oracle_csr5_tile_desc(omega, sigma, p, m, by, bs, npk, rowptr, tile_ptr, desc, off_ptr)
        off = np.zeros(max(nof, 1), dtype=np.int32)
        if nof:
            self.lib.oracle_csr5_desc_offset(omega, sigma, p, by, bs, npk, rowptr, tile_ptr, desc, off_ptr, off)
        out.update(tile_ptr=tile_ptr, tile_desc=desc[:p * omega * npk], offset_ptr=off_ptr, offsets=off[:nof],
                   tail_start=int(tile_ptr[p - 1] & 0x7FFFFFFF))
        if col is not None:
            t = np.empty(max(nnz, 1), dtype=np.int32)
            self.lib.oracle_csr5_transpose_i32(omega, sigma, nnz, tile_ptr, _pad(col), t)
            out["col_t"] = t[:nnz]
        return out


def _pad(a):
    """ndpointer rejects zero-length arrays from some numpy builds; hand it one element."""
    return a if len(a) else np.zeros(1, dtype=a.dtype)


# ---------------------------------------------------------------------------------------------
# The compiled reference.  Private structs mirrored from the reference sources (cited).
# ---------------------------------------------------------------------------------------------
class _Handle(C.Structure):  # include/spmv_Defines.h:44-70
    _fields_ = [("spmvMethod", C.c_int), ("data_size", C.c_ulong), ("nthreads", C.c_ulong),
                ("vectorizedWay", C.c_int), ("Level_3_opt_used", C.c_int),
                ("RowPtr", C.c_void_p), ("ColIdx", C.c_void_p), ("index", C.c_void_p),
                ("Matrix_Val", C.c_void_p), ("Y_temp", C.c_void_p), ("extraHandle", C.c_void_p)]


class _BalancedEnv(C.Structure):  # src/src_spmv/parallel_balanced2_spmv.c:9-19
    _fields_ = [(n, C.POINTER(C.c_int)) for n in
                ("csrSplitter", "Yid", "Apinter", "Start1", "End1", "Start2", "End2", "Bpinter", "label")]


class _YidEnv(C.Structure):  # src/src_spmv/parallel_balanced_Yid_spmv.c:4-13
    _fields_ = [(n, C.POINTER(C.c_int)) for n in ("Brow", "beginIdx", "Erow", "endIdx", "splitter", "Type")]


class _SigmaBlock(C.Structure):  # src/src_spmv/sell_C_Sigma_spmv.c:19-29
    _fields_ = [("C", C.c_int), ("times", C.c_int), ("ld", C.POINTER(C.c_int)), ("full", C.POINTER(C.c_int)),
                ("ColIndex", C.POINTER(C.c_int)), ("RowIndex", C.POINTER(C.c_int)), ("total", C.c_int),
                ("ValT", C.c_void_p)]


class _SigmaEnv(C.Structure):  # src/src_spmv/sell_C_Sigma_spmv.c:39-43
    _fields_ = [("Sigma", C.c_int), ("C", C.c_int), ("banner", C.c_int), ("sigmaBlock", C.POINTER(_SigmaBlock))]


class _Csr5Handle(C.Structure):  # src/src_spmv/csr5_avx2/anonymouslib_avx2.h:39-63 (no virtuals: plain layout)
    _fields_ = [("_format", C.c_int), ("_m", C.c_int), ("_n", C.c_int), ("_nnz", C.c_int),
                ("_csr_row_pointer", C.POINTER(C.c_int)), ("_csr_column_index", C.POINTER(C.c_int)),
                ("_csr_value", C.POINTER(C.c_double)),
                ("_csr5_sigma", C.c_int), ("_bit_y_offset", C.c_int), ("_bit_scansum_offset", C.c_int),
                ("_num_packet", C.c_int), ("_tail_partition_start", C.c_int), ("_p", C.c_int),
                ("_csr5_partition_pointer", C.POINTER(C.c_uint)), ("_csr5_partition_descriptor", C.POINTER(C.c_uint)),
                ("_num_offsets", C.c_int), ("_csr5_partition_descriptor_offset_pointer", C.POINTER(C.c_int)),
                ("_csr5_partition_descriptor_offset", C.POINTER(C.c_int)), ("_temp_calibrator", C.POINTER(C.c_double)),
                ("_x", C.POINTER(C.c_double))]


class RefHandle:
    """Owns one reference handle plus the arrays it borrows (the reference never copies them)."""

    def __init__(self, ref, ptr, keep):
        self.ref, self.ptr, self.keep = ref, ptr, keep

    @property
    def s(self):
        return self.ptr.contents

    def destroy(self):
        if self.ptr:
            self.ref.lib.spmv_destory_handle(self.ptr)
            self.ptr = None

    def __del__(self):
        self.destroy()


class Reference:
    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
        L = self.lib = C.CDLL(REF_SO)  # RTLD_LOCAL: same symbol names as libspmv_b200.so
        self.omp = C.CDLL("libgomp.so.1")
        HP = C.POINTER(_Handle)
        L.spmv_create_handle_all_in_one.argtypes = [C.POINTER(HP), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_ulong, C.c_int, C.c_ulong, C.c_int, C.c_char_p]
        L.spmv_create_handle_all_in_one.restype = None
        L.spmv.argtypes = [HP, C.c_int] + [C.c_void_p] * 5
        L.spmv.restype = None
        L.spmv_destory_handle.argtypes = [HP]
        L.spmv_destory_handle.restype = None
        self.HP = HP

    def max_threads(self) -> int:
        return int(self.omp.omp_get_max_threads())

    def set_threads(self, t: int) -> None:
        self.omp.omp_set_num_threads(int(t))

    def create(self, m, n, rowptr, col, val, nthreads, method, vec=VECTOR_AVX2, copy=True) -> RefHandle:
        """spmv_create_handle_all_in_one (src/src_spmv/common.c:123-190).  ``copy`` gives the handle its
        own ColIdx/Val (the reference's CSR5 transposes the caller's arrays in place)."""
        self.set_threads(nthreads)  # protocol of src/samples/test_spmv.c:87-91
        rp = np.ascontiguousarray(rowptr, dtype=np.int32)
        ci = np.array(col, dtype=np.int32, copy=True) if copy else col
        va = np.array(val, copy=True) if copy else val
        # one int of slack behind RowPtr: the reference's CSR5 builder reads RowPtr[m+1] (format_avx2.h:44-45)
        rp_s = np.empty(len(rp) + 1, dtype=np.int32)
        rp_s[:-1] = rp
        rp_s[-1] = -1
        h = self.HP()
        self.lib.spmv_create_handle_all_in_one(C.byref(h), m, n, rp_s.ctypes.data, ci.ctypes.data, va.ctypes.data,
                                               nthreads, method, va.dtype.itemsize, vec, b"oracle")
        return RefHandle(self, h, (rp_s, ci, va))

    def spmv(self, h: RefHandle, x, y=None):
        rp, ci, va = h.keep
        m = len(rp) - 2
        if y is None:
            y = np.zeros(m, dtype=va.dtype)
        self.lib.spmv(h.ptr, m, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, x.ctypes.data, y.ctypes.data)
        return y

    def serial(self, rowptr, col, val, x):
        """y of Method_Serial: THE parity target (src/src_spmv/serial_spmv.c:9-55)."""
        m = len(rowptr) - 1
        h = self.create(m, len(x), rowptr, col, val, 1, METHOD_SERIAL, VECTOR_NONE, copy=False)
        y = self.spmv(h, x)
        h.destroy()
        return y

    # -- structure dumps ----------------------------------------------------------------------
    def splitter(self, rowptr, T):
        """csrSplitter[T+1] of init_csrSplitter_balanced2 via a Method_Balanced2 handle."""
        m = len(rowptr) - 1
        z = np.zeros(max(int(rowptr[m]), 1), dtype=np.float64)
        h = self.create(m, m, rowptr, np.zeros(len(z), np.int32), z, T, METHOD_BALANCED2, copy=False)
        if h.s.spmvMethod == METHOD_BALANCED:
            arr = C.cast(h.s.extraHandle, C.POINTER(C.c_int))
            out, yid = np.array(arr[:T + 1], dtype=np.int32), np.full(T, -1, np.int32)
        else:
            env = C.cast(h.s.extraHandle, C.POINTER(_BalancedEnv)).contents
            out = np.array(env.csrSplitter[:T + 1], dtype=np.int32)
            yid = np.array(env.Yid[:T], dtype=np.int32)
        method = h.s.spmvMethod
        h.destroy()
        return out, yid, method

    def splitter_yid(self, rowptr, T):
        m = len(rowptr) - 1
        z = np.zeros(max(int(rowptr[m]), 1), dtype=np.float64)
        h = self.create(m, m, rowptr, np.zeros(len(z), np.int32), z, T, METHOD_BALANCED_YID, copy=False)
        env = C.cast(h.s.extraHandle, C.POINTER(_YidEnv)).contents
        out = {k: np.array(getattr(env, k)[:(2 * T if k == "splitter" else T)], dtype=np.int32)
               for k in ("splitter", "Type", "Brow", "beginIdx", "Erow", "endIdx")}
        h.destroy()
        return out

    def sell(self, rowptr, col, val, nthreads):
        """(sigma, banner, perm, widths(C=4), full(C=4)) of a Method_SellCSigma handle."""
        m = len(rowptr) - 1
        h = self.create(m, m, rowptr, col, val, nthreads, METHOD_SELLCSIGMA)
        env = C.cast(h.s.extraHandle, C.POINTER(_SigmaEnv)).contents
        sigma, banner = env.Sigma, env.banner
        perm, widths, full = [], [], []
        for b in range(banner // sigma if sigma else 0):
            blk = env.sigmaBlock[b]
            if not blk.ld:  # all rows of the window empty: RowIndex is NULL (sell_C_Sigma_spmv.c:84-92)
                perm.extend(range(b * sigma, (b + 1) * sigma))
                widths.extend([0] * blk.times)
                full.extend([0] * blk.times)
                continue
            perm.extend(blk.RowIndex[:sigma])
            ld = blk.ld[:blk.times + 1]
            widths.extend(int(ld[i + 1] - ld[i]) for i in range(blk.times))
            full.extend(blk.full[:blk.times])
        h.destroy()
        return sigma, banner, np.array(perm, np.int32), np.array(widths, np.int32), np.array(full, np.int32)

    def csr5(self, rowptr, col, val, nthreads=1):
        """Private members of anonymouslibHandle at the reference's (omega=4, sigma=16)."""
        m = len(rowptr) - 1
        h = self.create(m, m, rowptr, col, np.asarray(val, dtype=np.float64), nthreads, METHOD_CSR5)
        a = C.cast(h.s.extraHandle, C.POINTER(_Csr5Handle)).contents
        p, npk = a._p, a._num_packet
        out = dict(p=p, bit_y_offset=a._bit_y_offset, bit_scansum_offset=a._bit_scansum_offset, num_packet=npk,
                   tail_start=a._tail_partition_start,
                   tile_ptr=np.array(a._csr5_partition_pointer[:p + 1], dtype=np.uint32),
                   tile_desc=np.array(a._csr5_partition_descriptor[:p * 4 * npk], dtype=np.uint32),
                   offset_ptr=np.array(a._csr5_partition_descriptor_offset_pointer[:p + 1], dtype=np.int32),
                   offsets=np.array(a._csr5_partition_descriptor_offset[:a._num_offsets], dtype=np.int32)
                   if a._num_offsets else np.zeros(0, np.int32),
                   col_t=np.array(h.keep[1], copy=True))
        h.destroy()
        return out

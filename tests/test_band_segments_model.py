"""The band-segment layout (spmv_b200/csrc/band_seg.cuh) as a CPU model: from the numpy restatement of the layout that
the GPU structure test pins bit for bit (tests/cases.py::expected_band_segments: slots, segment-end bits, band masks),
y is computed the way the two GPU passes do -- per band the sums of maximal runs of entries up to an end bit (pass 1,
incl. the tile / chunk carries), then per row the sum of its segments in band order, found through the mask ranks
(pass 2) -- and compared with the oracle.  Shows that the layout alone determines y = A x (no row ids are stored)."""
import numpy as np
import pytest

from cases import all_cases, expected_band_segments
from spmv_b200 import matrices as M

CASES = all_cases()
CHUNK = 256  # entries per consumer warp of pass 1: runs that cross a chunk are completed by a carry


def bandseg_spmv(a, bands, x):
    bc, ptr, cnt, col, mask, nseg, val = expected_band_segments(a, bands, with_values=True)
    seg_lists = []
    for b in range(bands):
        lo, n = int(ptr[b]), int(cnt[b])
        c = col[lo:lo + n]
        end = (c >> np.uint32(31)).astype(bool)
        cols = (c & np.uint32(0x7FFFFFFF)).astype(np.int64)
        assert n == 0 or ((cols // bc).clip(max=bands - 1) == b).all()      # every entry of the band gathers from its slice
        prod = val[lo:lo + n].astype(np.float64) * x[cols].astype(np.float64)
        # pass 1, chunk by chunk: closed runs are written, the open run at a chunk's end is carried into the segment
        # that closes later (carry fix-up), exactly one sum per end bit
        sums, carry = [], 0.0
        for c0 in range(0, n, CHUNK):
            run = 0.0
            first = True
            for i in range(c0, min(c0 + CHUNK, n)):
                run += prod[i]
                if end[i]:
                    sums.append(run + (carry if first else 0.0))
                    if first:
                        carry, first = 0.0, False
                    run = 0.0
            carry = carry + run if first else run
        assert n == 0 or end[n - 1]                                             # a band never ends inside a segment
        seg_lists.append(np.array(sums))
    assert sum(len(s) for s in seg_lists) == nseg
    # pass 2: row r's segment in band b is the rank(r)-th of the band's list, rank = rows < r with bit b set
    y = np.zeros(a.m)
    for b in range(bands):
        rows = np.nonzero((mask >> np.uint64(b)) & np.uint64(1))[0]
        assert len(rows) == len(seg_lists[b])
        y[rows] += seg_lists[b]                                                 # (bands ascending: the order of the GPU fold)
    return y


@pytest.mark.parametrize("bands", [2, 7, 40])
@pytest.mark.parametrize("name", ["uni32", "uni5r", "skew", "hub", "lead_trail_empty", "lap48", "one_long_row", "tiny_m3", "empties", "exact2048"])
def test_the_layout_alone_gives_y(port, name, bands):
    a = CASES[name]()
    x = M.make_x(a.n, 5, np.float64)
    y = bandseg_spmv(a, bands, x)
    y_ex = port.spmv_exact(a.rowptr, a.col, a.val, x)
    tol = 8 * np.finfo(np.float64).eps * port.row_abs_sum(a.rowptr, a.col, a.val, x) + 0.5 * np.finfo(np.float64).eps * np.abs(y_ex)
    assert (np.abs(y - y_ex) <= tol).all()

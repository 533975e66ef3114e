"""Row-sharded SpMV across GPUs and the y -> x power-method loop.

Rows are partitioned by equal nnz with the reference's own splitter formula
(init_csrSplitter_balanced2, reference src/src_spmv/parallel_balanced2_spmv.c:41-53, nthreads = number
of GPUs), x is replicated, and every rank computes its y slice through the unchanged C API with NO
communication.  Only the iterated loop x <- A x needs an exchange: the y slices are all-gathered into the
next x (NCCL over NVLink on GPUs; gloo in the CPU tests), timed separately from the SpMV.

The local SpMV is injected as a callable so that this host logic is testable on CPU ranks (gloo) with the
oracle standing in for the CUDA library; on GPUs it is `Handle.spmv` of spmv_b200.api.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from . import api
from .matrices import CSR


def equal_nnz_partition(rowptr: np.ndarray, parts: int) -> np.ndarray:
    """splitter[g], g = 0..parts; rank g owns rows [splitter[g], splitter[g+1]).  splitter[0] is forced to
    0 so that leading empty rows belong to rank 0 (the reference's formula skips them)."""
    s = api.partition_rows(rowptr, parts)
    s[0] = 0
    return s


def local_shard(A: CSR, splitter: Sequence[int], rank: int) -> CSR:
    """Rows [splitter[rank], splitter[rank+1]) as a CSR with RowPtr rebased to 0 and GLOBAL column indices."""
    lo, hi = int(splitter[rank]), int(splitter[rank + 1])
    a, b = int(A.rowptr[lo]), int(A.rowptr[hi])
    return CSR(hi - lo, A.n, (A.rowptr[lo:hi + 1] - a).astype(np.int32), A.col[a:b].copy(), A.val[a:b].copy(),
               f"{A.name}[{lo}:{hi}]")


class PowerMethod:
    """x <- A x, `iters` times, un-normalised (the benchmark matrices keep 50 iterations inside fp range).

    spmv_local(x_full, y_slice) must write this rank's rows of A @ x_full into y_slice (a view into the
    next x).  `splitter` is the global row partition; `x0` the replicated start vector (torch tensor).
    """

    def __init__(self, spmv_local: Callable, splitter: Sequence[int], x0, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.spmv_local = spmv_local
        self.splitter = [int(v) for v in splitter]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert len(self.splitter) == self.world + 1 and self.splitter[-1] == x0.numel(), "square matrix expected"
        self.x = [x0.clone(), torch.empty_like(x0)]
        sizes = {self.splitter[g + 1] - self.splitter[g] for g in range(self.world)}
        self.equal = len(sizes) == 1
        self.cuda = x0.is_cuda

    def _allgather(self, xn):
        dist, lo, hi = self.dist, self.splitter[self.rank], self.splitter[self.rank + 1]
        if self.world == 1:
            return
        if self.equal:
            dist.all_gather_into_tensor(xn, xn[lo:hi], group=self.group)  # in place: slice g lands at its row offset
        else:  # unequal row counts (equal nnz does not mean equal rows): one broadcast per owner
            for g in range(self.world):
                dist.broadcast(xn[self.splitter[g]:self.splitter[g + 1]], src=g, group=self.group)

    def run(self, iters: int):
        """Returns (x_final, spmv_ms_per_iter, comm_ms_per_iter); times are device times on CUDA."""
        torch = self.torch
        lo, hi = self.splitter[self.rank], self.splitter[self.rank + 1]
        t_spmv = t_comm = 0.0
        if self.cuda:
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
        else:
            import time
        cur = 0
        for it in range(iters):
            xc, xn = self.x[cur], self.x[1 - cur]
            if self.cuda:
                ev[it][0].record()
                self.spmv_local(xc, xn[lo:hi])
                ev[it][1].record()
                self._allgather(xn)
                ev[it][2].record()
            else:
                t0 = time.perf_counter()
                self.spmv_local(xc, xn[lo:hi])
                t1 = time.perf_counter()
                self._allgather(xn)
                t2 = time.perf_counter()
                t_spmv += (t1 - t0) * 1e3
                t_comm += (t2 - t1) * 1e3
            cur = 1 - cur
        if self.cuda:
            torch.cuda.synchronize()
            t_spmv = sum(e[0].elapsed_time(e[1]) for e in ev)
            t_comm = sum(e[1].elapsed_time(e[2]) for e in ev)
        return self.x[cur], t_spmv / max(iters, 1), t_comm / max(iters, 1)


class FusedPowerMethod:
    """x <- A x with the all-gather fused into the SpMV: the kernel that produces y stores every value into
    this rank's next x AND, over NVLink peer mappings (CUDA IPC), into every other rank's next x at the
    same row offset (spmv_b200_set_y_peers).  No collective moves data; one 1-element all-reduce per
    iteration orders the ranks before the next x is read.

    `handle` is this rank's spmv_b200.api.Handle; `splitter` the global row partition; `x0` a CUDA tensor.
    The two x buffers are plain cudaMalloc allocations (IPC-exportable), not torch storage.
    """

    def __init__(self, handle, splitter: Sequence[int], x0, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.h = torch, dist, group, handle
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.splitter = [int(v) for v in splitter]
        self.n, self.item = x0.numel(), x0.element_size()
        assert self.splitter[-1] == self.n and len(self.splitter) == self.world + 1
        nbytes = self.n * self.item
        self.buf = [api.device_malloc(nbytes), api.device_malloc(nbytes)]
        api.device_memcpy(self.buf[0], x0, nbytes, 2)
        self.opened = []
        mine = [api.ipc_export(b) for b in self.buf]
        if self.world > 1:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
        else:
            everyone = [mine]
        lo = self.splitter[self.rank] * self.item
        self.peer_dst = [[], []]  # per buffer parity: where my slice starts inside every OTHER rank's buffer
        for q in range(self.world):
            if q == self.rank:
                continue
            for parity in (0, 1):
                base = api.ipc_open(everyone[q][parity])
                self.opened.append(base)
                self.peer_dst[parity].append(base + lo)
        self.flag = torch.zeros(1, device=x0.device)
        self.dtype, self.device = x0.dtype, x0.device

    def run(self, iters: int):
        torch, dist = self.torch, self.dist
        lo = self.splitter[self.rank] * self.item
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
        cur = 0
        for it in range(iters):
            nxt = 1 - cur
            self.h.set_y_peers(self.peer_dst[nxt])
            ev[it][0].record()
            self.h.spmv(self.buf[cur], self.buf[nxt] + lo)  # local slice + peer stores in one kernel
            ev[it][1].record()
            if self.world > 1:
                dist.all_reduce(self.flag, group=self.group)  # ordering only: everyone's stores have landed
            ev[it][2].record()
            cur = nxt
        torch.cuda.synchronize()
        self.h.set_y_peers([])
        self.cur = cur
        t_spmv = sum(e[0].elapsed_time(e[1]) for e in ev) / max(iters, 1)
        t_sync = sum(e[1].elapsed_time(e[2]) for e in ev) / max(iters, 1)
        return t_spmv, t_sync

    def result(self):
        x = self.torch.empty(self.n, dtype=self.dtype, device=self.device)
        api.device_memcpy(x, self.buf[self.cur], self.n * self.item, 2)
        return x

    def close(self):
        for p in self.opened:
            api.ipc_close(p)
        self.opened = []
        for b in self.buf:
            api.device_free(b)
        self.buf = []

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


# -------------------------------------------------------------------------------------------------------------
# Pipelined exchange: the all-gather of iteration k hides behind the SpMV of iteration k+1
# -------------------------------------------------------------------------------------------------------------
def ring_offsets(world: int):
    """Offsets d_1 .. d_{N-1} of the exchange schedule: in step j every rank r sends its slice to rank (r - d_j) and
    receives the slice of rank (r + d_j) -- a permutation per step, so no GPU's NVLink ingress is oversubscribed.
    The order +1, -1, +2, -2, ... brings the NEIGHBOURING slices first: column bands are contiguous column ranges,
    so the bands around a rank's own slice complete early whichever side they extend to."""
    out = [0]
    for k in range(1, world):
        d = (k + 1) // 2
        out.append(d if k % 2 == 1 else -d)
    return [d % world for d in out]


def band_ready_step(col_lo: int, col_hi: int, splitter: Sequence[int], rank: int) -> int:
    """Exchange step after which x[col_lo, col_hi) is complete on `rank`: slice g of the new x is produced by rank g
    (rows = columns of a square matrix) and arrives in the step j with (rank + d_j) mod N == g of ring_offsets
    (step 0 = the rank's own slice, already in place).  A band is ready once every owner overlapping its columns has
    arrived."""
    world = len(splitter) - 1
    offs = ring_offsets(world)
    step_of = {(rank + d) % world: j for j, d in enumerate(offs)}
    step = 0
    for g in range(world):
        lo, hi = int(splitter[g]), int(splitter[g + 1])
        if hi > lo and lo < col_hi and hi > col_lo:
            step = max(step, step_of[g])
    return step


class PipelinedPowerMethod:
    """x <- A x on N ranks with the y -> x exchange OVERLAPPED with the next SpMV.

    The layouts this library builds for matrices whose x does not fit L2 are banded by columns, and band b of the
    next SpMV only reads x[col_lo_b, col_hi_b) -- the y slices of the few ranks that own those rows.  So instead of
    "SpMV, then all-gather, then SpMV" the loop is software-pipelined:

      * exchange k runs on a communication stream as N-1 steps; in step j every rank sends its fresh y slice to
        rank (r - d_j) and receives the slice of rank (r + d_j), d = +1, -1, +2, -2, ... (ring_offsets): one
        permutation per step, so no GPU's NVLink ingress is oversubscribed, and slices arrive in a known order,
        nearest neighbours first;
      * SpMV k+1 starts with the bands that only need the rank's OWN slice (already in place) and launches every
        other band as soon as the steps it depends on have completed (stream waits on events; the host never
        blocks), through spmv_b200_spmv_bands / spmv_b200_spmv_finish -- the fold adds the band partials in band
        order whatever order they were computed in, so the result is bit-identical to the plain loop.

    Each rank may hold its rows as SEVERAL handles (row sub-blocks): the reference ABI is int32, so a shard with
    2^31 or more non-zeros has to be presented as more than one matrix.  `parts` = [(handle, row_lo, row_hi)] with
    rows relative to the rank's first row; handles expose bands() / band_columns() / spmv_bands() / spmv_finish()
    / spmv() (spmv_b200.api.Handle, or a CPU stand-in in the gloo tests).

    Transport: NCCL point-to-point (torch.distributed.batch_isend_irecv) over NVLink -- plumbing; what makes the
    overlap possible is the band-staged kernel interface.
    """

    def __init__(self, parts, splitter: Sequence[int], x0, group=None, overlap: bool = True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.splitter = [int(v) for v in splitter]
        assert len(self.splitter) == self.world + 1 and self.splitter[-1] == x0.numel(), "square matrix expected"
        self.parts = list(parts)
        self.lo, self.hi = self.splitter[self.rank], self.splitter[self.rank + 1]
        assert self.parts and self.parts[0][1] == 0 and self.parts[-1][2] == self.hi - self.lo
        self.x = [x0.clone(), x0.clone()]
        self.cuda = x0.is_cuda
        self.overlap = overlap
        # schedule[j] = [(part index, first band, band count)] to launch once exchange step j has completed
        self.schedule = [[] for _ in range(self.world)]
        self.whole = []  # parts that cannot be staged: plain spmv() after the last step
        for pi, (h, r0, r1) in enumerate(self.parts):
            K = h.bands()
            if K <= 1:
                self.whole.append(pi)
                continue
            steps = [band_ready_step(*h.band_columns(b), self.splitter, self.rank) if overlap else self.world - 1
                     for b in range(K)]
            b = 0
            while b < K:  # runs of consecutive bands that become ready in the same step: one launch each
                e = b
                while e + 1 < K and steps[e + 1] == steps[b]:
                    e += 1
                self.schedule[steps[b]].append((pi, b, e - b + 1))
                b = e + 1
        if self.cuda:
            self.comm_stream = torch.cuda.Stream()
            self.ev_recv = [[torch.cuda.Event() for _ in range(self.world)] for _ in range(2)]
            self.ev_y = [torch.cuda.Event() for _ in range(2)]
        self.offsets = ring_offsets(self.world)

    # ---- one exchange: N-1 ring steps on the communication stream; returns nothing, completion is in the events ----
    def _exchange(self, buf, parity):
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return
        mine = buf[self.lo:self.hi]
        if self.cuda:
            self.ev_y[parity].record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(self.ev_y[parity])
                for j in range(1, self.world):
                    d = self.offsets[j]
                    dst, src = (self.rank - d) % self.world, (self.rank + d) % self.world
                    ops = []
                    if mine.numel():
                        ops.append(dist.P2POp(dist.isend, mine, dst, group=self.group))
                    theirs = buf[self.splitter[src]:self.splitter[src + 1]]
                    if theirs.numel():
                        ops.append(dist.P2POp(dist.irecv, theirs, src, group=self.group))
                    for w in (dist.batch_isend_irecv(ops) if ops else []):
                        w.wait()  # orders the communication stream after the step; the host does not block
                    self.ev_recv[parity][j].record(self.comm_stream)
        else:
            for j in range(1, self.world):
                d = self.offsets[j]
                dst, src = (self.rank - d) % self.world, (self.rank + d) % self.world
                ops = []
                if mine.numel():
                    ops.append(dist.P2POp(dist.isend, mine, dst, group=self.group))
                theirs = buf[self.splitter[src]:self.splitter[src + 1]]
                if theirs.numel():
                    ops.append(dist.P2POp(dist.irecv, theirs, src, group=self.group))
                for w in (dist.batch_isend_irecv(ops) if ops else []):
                    w.wait()

    def _spmv(self, xc, xn, wait_parity):
        """One SpMV in band stages; stage j waits for step j of the exchange that is filling xc (if any)."""
        torch = self.torch
        cur = torch.cuda.current_stream() if self.cuda else None
        for j in range(self.world):
            if j > 0 and wait_parity is not None and self.cuda:
                cur.wait_event(self.ev_recv[wait_parity][j])
            for pi, b0, cnt in self.schedule[j]:
                self.parts[pi][0].spmv_bands(b0, cnt, xc)
        for pi in self.whole:
            h, r0, r1 = self.parts[pi]
            h.spmv(xc, xn[self.lo + r0:self.lo + r1])
        for pi, (h, r0, r1) in enumerate(self.parts):
            if pi not in self.whole:
                h.spmv_finish(xn[self.lo + r0:self.lo + r1])

    def run(self, iters: int, exchange: bool = True):
        """`iters` iterations; returns (x_final, ms per iteration).  exchange=False times the staged SpMV alone (x is
        not advanced between ranks: a pure compute measurement)."""
        torch = self.torch
        if self.cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
        else:
            import time
            t0 = time.perf_counter()
        cur = 0
        wait_parity = None
        for _ in range(iters):
            xc, xn = self.x[cur], self.x[1 - cur]
            self._spmv(xc, xn, wait_parity)
            if exchange:
                self._exchange(xn, 1 - cur)
                wait_parity = 1 - cur
            cur = 1 - cur
        if self.cuda:
            if exchange and self.world > 1 and iters > 0:
                torch.cuda.current_stream().wait_event(self.ev_recv[cur][self.world - 1])  # the last x is complete
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        else:
            ms = (time.perf_counter() - t0) * 1e3
        self.cur = cur
        return self.x[cur], ms / max(iters, 1)

    def exchange_only(self, iters: int):
        """The exchange alone, back to back (ms per exchange): what the overlap has to hide."""
        torch = self.torch
        if self.world == 1:
            return 0.0
        if self.cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
        else:
            import time
            t0 = time.perf_counter()
        scratch = self.x[1 - getattr(self, "cur", 0)]
        for _ in range(iters):
            self._exchange(scratch, 0)
            if self.cuda:
                torch.cuda.current_stream().wait_event(self.ev_recv[0][self.world - 1])
        if self.cuda:
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / max(iters, 1)
        return (time.perf_counter() - t0) * 1e3 / max(iters, 1)


class CopyEnginePowerMethod:
    """The pipelined loop of PipelinedPowerMethod with the exchange done by the COPY ENGINES over NVLink peer
    mappings instead of NCCL kernels: a persistent SpMV kernel owns every SM, so a communication kernel launched
    next to it waits for SMs (measured on 2 GPUs: NCCL send/recv steps overlapped only partly, and ran at 367 GB/s);
    DMA copies need none.

    Per iteration k, rank r, after the fold kernel has written its slice of x_{k+1} into its own buffer:
      communication stream, for every step j (ring_offsets): wait until the destination has finished READING the
      target buffer two iterations ago (flag free[dst] >= k, written by dst into r's memory after its fold of
      iteration k-1), copy the slice into the destination's buffer (cudaMemcpyAsync on the CUDA-IPC mapping: a peer
      DMA), then write the arrival flag arrive[r] = k+1 into the destination's memory (cuStreamWriteValue32: ordered
      after the copy);
      compute stream of iteration k+1: before the bands that need owner s, wait for arrive[s] >= k+1
      (cuStreamWaitValue32 on local memory).
    No collective, no kernel, no host synchronisation inside the loop.  Same arithmetic as the plain loop: bitwise
    equal results (asserted by bench.py).
    """

    def __init__(self, parts, splitter: Sequence[int], x0, group=None, lanes: int = 1):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.splitter = [int(v) for v in splitter]
        self.parts = list(parts)
        self.n, self.item = x0.numel(), x0.element_size()
        # `lanes` exchange steps are in flight at a time, each on its own stream (its own copy engine): step j runs on
        # stream (j-1) mod lanes.  Measured on 8 GPUs (C5, 268 MB slices): 1 lane 4.41 ms per exchange, 2: 4.69,
        # 3: 4.96, 7: 5.18 -- the steps are permutations, so one copy per GPU at a time already fills every link
        self.lanes = max(1, int(lanes))
        self.lo, self.hi = self.splitter[self.rank], self.splitter[self.rank + 1]
        self.dtype, self.device = x0.dtype, x0.device
        W = self.world
        nbytes = self.n * self.item
        self.buf = [api.device_malloc(nbytes), api.device_malloc(nbytes)]
        for b in self.buf:
            api.device_memcpy(b, x0, nbytes, 2)
        self.flags = api.device_malloc(2 * W * 4)  # arrive[W], free[W]
        zero = torch.zeros(2 * W, dtype=torch.int32, device=x0.device)
        api.device_memcpy(self.flags, zero, 2 * W * 4, 2)
        torch.cuda.synchronize()
        mine = [api.ipc_export(self.buf[0]), api.ipc_export(self.buf[1]), api.ipc_export(self.flags)]
        if W > 1:
            everyone = [None] * W
            dist.all_gather_object(everyone, mine, group=group)
        else:
            everyone = [mine]
        self.opened, self.peer_buf, self.peer_flags = [], {}, {}
        for q in range(W):
            if q == self.rank:
                continue
            ptrs = [api.ipc_open(hd) for hd in everyone[q]]
            self.opened += ptrs
            self.peer_buf[q] = ptrs[:2]
            self.peer_flags[q] = ptrs[2]
        self.offsets = ring_offsets(W)
        self.step_of = {(self.rank + d) % W: j for j, d in enumerate(self.offsets)}
        self.schedule = [[] for _ in range(W)]
        self.whole = []
        for pi, (h, r0, r1) in enumerate(self.parts):
            K = h.bands()
            if K <= 1:
                self.whole.append(pi)
                continue
            steps = [band_ready_step(*h.band_columns(b), self.splitter, self.rank) for b in range(K)]
            b = 0
            while b < K:
                e = b
                while e + 1 < K and steps[e + 1] == steps[b]:
                    e += 1
                self.schedule[steps[b]].append((pi, b, e - b + 1))
                b = e + 1
        self.comm = torch.cuda.Stream()
        self.lane = [self.comm] + [torch.cuda.Stream() for _ in range(self.lanes - 1)]
        self.ev_y = [torch.cuda.Event() for _ in range(2)]
        self.it = 0      # iterations completed so far (flags are monotonic across run() calls)
        self.cur = 0
        # flag writes into PEER memory: cuStreamWriteValue32 on the IPC mapping where the driver allows it, else a local
        # write into a scratch word followed by a 4-byte peer copy (both stream-ordered)
        self.scratch = api.device_malloc(4 * 2 * W)
        self.direct_flags = True
        if W > 1:
            dist.barrier(group=group)  # every rank has mapped every buffer before anyone writes
            try:
                q = next(iter(self.peer_flags))
                api.stream_write32(self.comm.cuda_stream, self.peer_flags[q] + 4 * self.rank, 0)
                self.comm.synchronize()
            except RuntimeError:
                api.clear_error()
                self.direct_flags = False
            dist.barrier(group=group)

    def _flag_to_peer(self, stream_h, q, index, value):
        dst = self.peer_flags[q] + 4 * index
        if self.direct_flags:
            api.stream_write32(stream_h, dst, value)
        else:
            word = self.scratch + 4 * (index % (2 * self.world))
            api.stream_write32(stream_h, word, value)
            api.memcpy_async(dst, word, 4, stream_h)

    def _arrive(self, owner):  # address of arrive[owner] in MY flags
        return self.flags + 4 * owner

    def run(self, iters: int, exchange: bool = True, compute: bool = True):
        """ms per iteration.  exchange=False: the band-staged SpMV alone; compute=False: the exchange alone (the same
        copies and flags, nothing to hide behind)."""
        torch, W = self.torch, self.world
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if W > 1:
            self.dist.barrier(group=self.group)
        cs = torch.cuda.current_stream()
        cs_h, comm_h = cs.cuda_stream, self.comm.cuda_stream
        e0.record()
        for _ in range(iters):
            k = self.it
            xc, xn = self.buf[self.cur], self.buf[1 - self.cur]
            for j in range(W):
                if j > 0 and exchange and k > 0:
                    owner = (self.rank + self.offsets[j]) % W
                    api.stream_wait32_geq(cs_h, self._arrive(owner), k)  # owner's slice of x_k is in xc
                if compute:
                    for pi, b0, cnt in self.schedule[j]:
                        self.parts[pi][0].spmv_bands(b0, cnt, xc)
            if compute:
                for pi in self.whole:  # (every arrival flag has been waited for above)
                    h, r0, r1 = self.parts[pi]
                    h.spmv(xc, xn + (self.lo + r0) * self.item)
                for pi, (h, r0, r1) in enumerate(self.parts):
                    if pi not in self.whole:
                        h.spmv_finish(xn + (self.lo + r0) * self.item)
            if exchange and W > 1:
                self.ev_y[self.cur].record(cs)
                for st in self.lane:
                    st.wait_event(self.ev_y[self.cur])
                # (communication stream, i.e. after the fold kernel) I have finished reading buffer `cur` for
                # iteration k: tell everyone -- they overwrite it with their slices of x_{k+2}
                for q in self.peer_flags:
                    self._flag_to_peer(comm_h, q, W + self.rank, k + 1)
                nbytes = (self.hi - self.lo) * self.item
                for j in range(1, W):
                    dst = (self.rank - self.offsets[j]) % W
                    # dst must be done with iteration k-1 (the last reader of ITS buffer 1-cur) before I overwrite it
                    lane_h = self.lane[(j - 1) % self.lanes].cuda_stream  # steps j, j + lanes, ... share a stream: in order
                    if k > 0:
                        api.stream_wait32_geq(lane_h, self.flags + 4 * (W + dst), k)
                    api.memcpy_async(self.peer_buf[dst][1 - self.cur] + self.lo * self.item, xn + self.lo * self.item, nbytes, lane_h)
                    self._flag_to_peer(lane_h, dst, self.rank, k + 1)
            self.cur = 1 - self.cur
            if exchange:
                self.it += 1
        if exchange and W > 1 and iters > 0:
            for j in range(1, W):  # the last x is complete here
                api.stream_wait32_geq(cs_h, self._arrive((self.rank + self.offsets[j]) % W), self.it)
        e1.record()
        torch.cuda.synchronize()
        for st in self.lane:
            st.synchronize()
        if not exchange and iters % 2 == 1:
            self.cur = 1 - self.cur  # a compute-only measurement leaves the loop where it was
        if W > 1:
            self.dist.barrier(group=self.group)
        return e0.elapsed_time(e1) / max(iters, 1)

    def result(self):
        x = self.torch.empty(self.n, dtype=self.dtype, device=self.device)
        api.device_memcpy(x, self.buf[self.cur], self.n * self.item, 2)
        return x

    def close(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        for p in self.opened:
            api.ipc_close(p)
        self.opened = []
        for b in self.buf + [self.flags, self.scratch]:
            api.device_free(b)
        self.buf = []

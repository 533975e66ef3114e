"""The N>1 host logic on CPU: world_size-2 (and 3) gloo ranks, equal-nnz row shards, replicated x, in-place
all-gather of the y slices into the next x.  The oracle stands in for the CUDA library as the local SpMV."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_file, case, iters, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as O
    from spmv_b200 import matrices as M, multigpu as G
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_file)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = O.Port()
    A = M.uniform_random(600, 600, 8, seed=3) if case == "uniform" else M.skewed(500, 500, max_len=200, seed=5)
    A.val *= 0.2  # keep 5 un-normalised iterations tame
    split = G.equal_nnz_partition(A.rowptr, world)
    mine = G.local_shard(A, split, rank)

    def spmv_local(x_full, y_slice):
        y_slice.copy_(torch.from_numpy(P.spmv_serial(mine.rowptr, mine.col, mine.val, x_full.numpy())))

    x0 = torch.from_numpy(M.make_x(A.n, 1, np.float64))
    pm = G.PowerMethod(spmv_local, split, x0)
    x, _, _ = pm.run(iters)
    # single-process truth
    xs = x0.numpy().copy()
    for _ in range(iters):
        xs = P.spmv_serial(A.rowptr, A.col, A.val, xs)
    ok = np.array_equal(x.numpy().view(np.uint8), xs.view(np.uint8))
    q.put((rank, bool(ok), [int(v) for v in split], bool(pm.equal)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "uniform"), (2, "skewed"), (3, "skewed")])
def test_sharded_power_method_matches_single_process(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 3 + (7 if case == "skewed" else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    splits = {tuple(s) for _, _, s, _ in res}
    assert len(splits) == 1
    if case == "uniform":
        assert all(eq for *_, eq in res)      # equal nnz <=> equal rows: plain all-gather path
    else:
        assert not any(eq for *_, eq in res)  # unequal slices: broadcast path


def test_partition_and_shards_cover_the_matrix():
    sys.path.insert(0, ROOT)
    from spmv_b200 import matrices as M, multigpu as G
    A = M.skewed(2000, 2000, max_len=900)
    for parts in (1, 2, 4, 8):
        s = G.equal_nnz_partition(A.rowptr, parts)
        assert s[0] == 0 and s[-1] == A.m and (np.diff(s) >= 0).all()
        shards = [G.local_shard(A, s, g) for g in range(parts)]
        assert sum(sh.nnz for sh in shards) == A.nnz
        assert np.array_equal(np.concatenate([sh.col for sh in shards]), A.col)
        nnz = np.array([sh.nnz for sh in shards])
        assert nnz.max() <= A.nnz / parts + np.diff(A.rowptr).max()  # balanced up to one row


# -------------------------------------------------------------------------------------------------------------
# pipelined exchange (PipelinedPowerMethod): band-staged SpMV + ring exchange, CPU ranks over gloo
# -------------------------------------------------------------------------------------------------------------
class _BandedCpuPart:
    """CPU stand-in for a banded spmv_b200 handle: rows [r0, r1) of a shard, K column bands, partial sums per band
    folded in band order (the contract of spmv_b200_spmv_bands / spmv_b200_spmv_finish)."""

    def __init__(self, shard, K, n):
        import scipy.sparse as sp
        self.K, self.n = K, n
        self.bc = -(-n // K)
        A = sp.csr_matrix((shard.val, shard.col, shard.rowptr), shape=(shard.m, n))
        self.blocks = [A[:, b * self.bc:min((b + 1) * self.bc, n)].tocsr() for b in range(K)]
        self.partial = [None] * K
        self.log = []

    def bands(self):
        return self.K

    def band_columns(self, b):
        return b * self.bc, min((b + 1) * self.bc, self.n)

    def spmv_bands(self, b0, cnt, x):
        for b in range(b0, b0 + cnt):
            lo, hi = self.band_columns(b)
            xs = x.numpy()[lo:hi]
            assert not np.isnan(xs).any(), f"band {b} launched before its slice of x had arrived"
            self.partial[b] = self.blocks[b] @ xs
            self.log.append(b)

    def spmv_finish(self, y):
        acc = np.zeros(len(self.partial[0]))
        for b in range(self.K):
            acc = acc + self.partial[b]
        y.copy_(torch.from_numpy(acc))
        self.partial = [None] * self.K

    def spmv(self, x, y):
        self.spmv_bands(0, self.K, x)
        self.spmv_finish(y)


def _pipe_worker(rank, world, port, K, nparts, q):
    sys.path.insert(0, ROOT)
    from spmv_b200 import matrices as M, multigpu as G
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = M.skewed(700, 700, max_len=150, seed=9)
    A.val *= 0.1
    split = G.equal_nnz_partition(A.rowptr, world)
    mine = G.local_shard(A, split, rank)
    # the rank's rows as `nparts` row sub-blocks (a shard with >= 2^31 non-zeros has to be several int32 handles)
    cuts = [mine.m * i // nparts for i in range(nparts + 1)]
    parts = []
    for i in range(nparts):
        sub = G.local_shard(mine, cuts, i)
        parts.append((_BandedCpuPart(sub, K, A.n), cuts[i], cuts[i + 1]))
    x0 = torch.from_numpy(M.make_x(A.n, 1, np.float64))
    outs = []
    for overlap in (True, False):
        pm = G.PipelinedPowerMethod(parts, split, x0, overlap=overlap)
        # poison everything that is not this rank's own slice in the buffer the first exchange fills: a band that is
        # scheduled before its owners' step would read NaN and trip the assert in spmv_bands
        x, _ = pm.run(4)
        outs.append(x.numpy().copy())
        if overlap:
            sched = [[(pi, b0, c) for pi, b0, c in step] for step in pm.schedule]
    # single-process truth with the same banded arithmetic
    whole = _BandedCpuPart(A, K, A.n)
    xs = x0.clone()
    for _ in range(4):
        y = torch.empty_like(xs)
        whole.spmv(xs, y)
        xs = y
    same = np.array_equal(outs[0].view(np.uint8), outs[1].view(np.uint8))
    close = np.allclose(outs[0], xs.numpy(), rtol=1e-12, atol=0)
    q.put((rank, bool(same), bool(close), sched))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,K,nparts", [(2, 4, 1), (2, 3, 2), (3, 7, 1)])
def test_pipelined_power_method_on_gloo(world, K, nparts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world * 5 + K
    procs = [ctx.Process(target=_pipe_worker, args=(r, world, port, K, nparts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, close, sched in res:
        assert same, f"rank {rank}: overlapped schedule changed the bits"
        assert close, f"rank {rank}: differs from the single-process loop"
        assert sum(c for step in sched for _, _, c in step) == K * nparts  # every band of every part exactly once
        assert sum(len(step) for step in sched[:-1]) >= 1                   # something starts before the last step


def test_band_ready_step_covers_exactly_the_owners_of_a_band():
    sys.path.insert(0, ROOT)
    from spmv_b200 import multigpu as G
    split = [0, 100, 250, 250, 400]  # rank 2 owns nothing
    for rank in range(4):
        for lo, hi in [(0, 50), (90, 110), (240, 260), (0, 400), (399, 400), (100, 250)]:
            j = G.band_ready_step(lo, hi, split, rank)
            offs = G.ring_offsets(4)
            assert sorted(offs) == [0, 1, 2, 3] and offs[:3] == [0, 1, 3]   # own, +1, -1, then +2
            arrived = {(rank + offs[s]) % 4 for s in range(j + 1)}
            covered = set()
            for g in arrived:
                covered |= set(range(max(lo, split[g]), min(hi, split[g + 1])))
            assert covered == set(range(lo, hi)), (rank, lo, hi, j)
            if j > 0:  # and not a step earlier
                early = {(rank + offs[s]) % 4 for s in range(j)}
                cov = set()
                for g in early:
                    cov |= set(range(max(lo, split[g]), min(hi, split[g + 1])))
                assert cov != set(range(lo, hi))

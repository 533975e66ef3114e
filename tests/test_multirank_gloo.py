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

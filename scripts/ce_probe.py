"""8-GPU probe of the copy-engine exchange (C5 or C2 shards): the full loop, the SpMV alone and the exchange alone for
several numbers of lanes.  Run it under `timeout` (e.g. `timeout 300 python -m torch.distributed.run ... ce_probe.py 1,2 c5`):
a rank that hangs holds the whole box until gpurun's own limit."""
import os, sys, time, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spmv_b200 import api, build, matrices as M, multigpu as G
build.build()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = sys.argv[2] if len(sys.argv) > 2 else "c5"
n, k = (1 << 28, 16) if wl == "c5" else (1 << 24, 32)
mp = n // world
split = [g * mp for g in range(world + 1)]
if wl != "c5":
    api.set_option("x_bands", world)
A = api.gen_uniform(mp, n, k, M.SEED_C5, rank * mp, False, 8)
h = A.handle(1)
if wl == "c5":
    A.destroy()
x0 = torch.empty(n, dtype=torch.float64, device="cuda")
api.gen_x(x0, n, M.SEED_C5, False, 8)
out = {}
for lanes in [int(v) for v in sys.argv[1].split(",")]:
    cp = G.CopyEnginePowerMethod([(h, 0, mp)], split, x0, lanes=lanes)
    cp.run(2)
    t_loop = cp.run(20)
    t_spmv = cp.run(10, exchange=False)
    t_x = cp.run(10, compute=False)
    t = torch.tensor([t_loop, t_spmv, t_x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cp.close()
    out[lanes] = [round(float(v), 3) for v in t.tolist()]
    if rank == 0:
        print(f"{wl} lanes {lanes}: loop {out[lanes][0]} ms, spmv alone {out[lanes][1]}, exchange alone {out[lanes][2]}", flush=True)
if rank == 0:
    json.dump(out, open(f"gpurun_out/ce_probe_{wl}.json", "w"))
h.destroy()
dist.destroy_process_group()

#!/bin/bash
B="python bench.py --workload c5shard --steps 10 --warmup 3 --no-cpu --power-iters 0 --also ''"
rm -f gpurun_out/c5_probe2.txt
python - >> gpurun_out/c5_probe2.txt 2>&1 <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persistMax", getattr(p, "persisting_l2_cache_max_size", None), "window", getattr(p,"access_policy_max_window_size",None))
PY
for p in 0 -1; do for k in 32 44 64; do
  echo "== l2_persist=$p seg_bands=$k" >> gpurun_out/c5_probe2.txt
  SPMV_B200_L2_PERSIST=$p SPMV_B200_SEG_BANDS=$k eval $B 2>> gpurun_out/c5_probe2.txt > /dev/null
done; done
grep "rank 0\|==\|L2" gpurun_out/c5_probe2.txt

#!/bin/bash
export SPMV_B200_SEG_BANDS=64
B="python bench.py --workload c5shard --steps 4 --warmup 3 --no-cpu --power-iters 0 --also ''"
for c in 2 3; do
SPMV_B200_SEG_CTAS=$c eval $B 2>&1 | grep "rank 0"
done
eval $B > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bseg_kernel' -s 2 -c 1 -o gpurun_out/r02_c5shard_bseg python bench.py --workload c5shard --steps 4 --warmup 3 --no-cpu --power-iters 0 --also '' > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log

#!/bin/bash
B="python bench.py --workload c5shard --steps 10 --warmup 3 --no-cpu --power-iters 0 --also ''"
rm -f gpurun_out/c5_sweep.txt
for p in 1 2; do for k in 40 44 48 52 56; do
  echo "== seg_prefetch=$p seg_bands=$k" >> gpurun_out/c5_sweep.txt
  SPMV_B200_SEG_PREFETCH=$p SPMV_B200_SEG_BANDS=$k eval $B 2>> gpurun_out/c5_sweep.txt > /dev/null
done; done
grep "rank 0\|==" gpurun_out/c5_sweep.txt

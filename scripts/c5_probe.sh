#!/bin/bash
# C5-shard probe: tests of the band-segment layout, option sweep, per-kernel device times
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "band_segment or band_staged" > gpurun_out/t_seg.log 2>&1; tail -5 gpurun_out/t_seg.log
B="python bench.py --workload c5shard --steps 10 --warmup 3 --no-cpu --power-iters 0 --also ''"
rm -f gpurun_out/c5_probe.txt
for k in 32 48 64; do
  echo "== seg_bands=$k" >> gpurun_out/c5_probe.txt
  SPMV_B200_SEG_BANDS=$k eval $B 2>> gpurun_out/c5_probe.txt > /dev/null
done
for k in 32 64; do
SPMV_B200_SEG_BANDS=$k eval $B > gpurun_out/plain.log 2>&1 &&
SPMV_B200_SEG_BANDS=$k ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_active --clock-control none -k regex:'bseg|carry' -s 12 -c 4 --csv --log-file gpurun_out/c5_launches_$k.csv python bench.py --workload c5shard --steps 10 --warmup 3 --no-cpu --power-iters 0 --also '' > gpurun_out/ncu.log 2>&1
done
grep "rank 0\|==" gpurun_out/c5_probe.txt

#!/bin/bash
# Bench + ncu evidence for the round: plain bench, launch list of the same command, full capture of the
# dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --power-iters 0"
timeout 900 python bench.py --steps ${STEPS:-50} --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -4 gpurun_out/bench.err
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:csr_vector -s 3 -c 2 -o gpurun_out/top_kernel $B > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log

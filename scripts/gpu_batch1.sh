#!/bin/bash
# round-1 batch: probes, per-config sweeps, ncu captures of every kernel family on C1/C3/C4 (condensed on
# the box: gpurun copies back at most 64 MiB)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 120 ./scripts/gather_probe2 > gpurun_out/probe2.txt 2>&1; echo "probe2 rc=$?"
for w in c1 c4 c3 c2; do
  fl=""; [ $w = c1 ] && fl="--flush"
  timeout 400 python scripts/sweep.py --workload $w $fl --steps 30 > gpurun_out/sweep_$w.txt 2>&1; echo "sweep $w rc=$?"
done
for w in c1 c4 c3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_vector|sell_kernel|csr5_kernel|merge_path|row_block|nnz_split|reforder" -c 7 -o gpurun_out/r01b_${w}_kernels -f \
     python scripts/sweep.py --workload $w --profile-only 1 > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
  ncu -i gpurun_out/r01b_${w}_kernels.ncu-rep --page raw --csv > gpurun_out/r01b_${w}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r01b_${w}_kernels.ncu-rep --page details > gpurun_out/r01b_${w}_details.txt 2>/dev/null
  rm -f gpurun_out/r01b_${w}_kernels.ncu-rep
done
du -sh gpurun_out

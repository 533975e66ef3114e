#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for w in c4 c2; do timeout 400 python scripts/sweep.py --workload $w --steps 30 --variants "parallel" > gpurun_out/sweep13_$w.txt 2>&1; grep -v "^# " gpurun_out/sweep13_$w.txt; done
timeout 400 python scripts/sweep.py --workload c1 --flush --steps 30 --variants "parallel" > gpurun_out/sweep13_c1.txt 2>&1; grep -v "^# " gpurun_out/sweep13_c1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"band_coo_kernel" -s 3 -c 2 -o gpurun_out/r01d_c5_band_coo -f \
     python scripts/sweep.py --workload c5shard --profile-only 1 --variants "parallel,coo_bands=32" > gpurun_out/ncu_c5.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r01d_c5_band_coo.ncu-rep --page raw --csv > gpurun_out/r01d_c5_band_coo_raw.csv 2>/dev/null
ncu -i gpurun_out/r01d_c5_band_coo.ncu-rep --page details > gpurun_out/r01d_c5_band_coo_details.txt 2>/dev/null
ncu -i gpurun_out/r01d_c5_band_coo.ncu-rep --page source --csv > gpurun_out/r01d_c5_band_coo_source.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep

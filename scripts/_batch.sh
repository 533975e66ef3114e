#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 400 python scripts/sweep.py --workload c3 --steps 30 --variants "parallel;parallel,long_thr=2048;parallel,long_thr=4096;parallel,long_thr=512;balanced_yid" > gpurun_out/sweep25_c3.txt 2>&1; grep -v "^# device" gpurun_out/sweep25_c3.txt
C="python scripts/sweep.py --workload c3 --profile-only 1 --variants parallel"
$C > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"long_seg_kernel" -c 1 -o gpurun_out/r01e_c3_long_seg -f $C > gpurun_out/ncu_c3_long.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r01e_c3_long_seg.ncu-rep --page details > gpurun_out/r01e_c3_long_seg_details.txt 2>/dev/null
ncu -i gpurun_out/r01e_c3_long_seg.ncu-rep --page raw --csv > gpurun_out/r01e_c3_long_seg_raw.csv 2>/dev/null

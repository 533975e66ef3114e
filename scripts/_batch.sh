#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log
V="balanced_yid;balanced2,force_merge=1;balanced"
for w in c4 c2 c3; do timeout 400 python scripts/sweep.py --workload $w --steps 30 --variants "$V" > gpurun_out/sweep21_$w.txt 2>&1; grep -v "^# device" gpurun_out/sweep21_$w.txt; done
timeout 400 python scripts/sweep.py --workload c1 --flush --steps 30 --variants "$V" > gpurun_out/sweep21_c1.txt 2>&1; grep -v "^# device" gpurun_out/sweep21_c1.txt
timeout 600 python scripts/sweep.py --workload c5shard --steps 10 --variants "parallel" > gpurun_out/sweep21_c5.txt 2>&1; grep -v "^# device" gpurun_out/sweep21_c5.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 400 python scripts/sweep.py --workload c3 --steps 30 --variants "csr5;balanced2;balanced_yid" > gpurun_out/sweep28_c3.txt 2>&1; grep -v "^# device" gpurun_out/sweep28_c3.txt
timeout 600 python scripts/sweep.py --workload c5shard --steps 10 --variants "parallel" > gpurun_out/sweep28_c5.txt 2>&1; grep -v "^# device" gpurun_out/sweep28_c5.txt

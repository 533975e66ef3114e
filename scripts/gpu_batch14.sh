#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python scripts/sweep.py --workload c5shard --steps 10 --variants "parallel;parallel,coo_bands=32;parallel,coo_bands=40" > gpurun_out/sweep14_c5.txt 2>&1; grep -v "^# device" gpurun_out/sweep14_c5.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "coo_band" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest.log

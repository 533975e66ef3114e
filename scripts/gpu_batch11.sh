#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "coo_band" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 600 python scripts/sweep.py --workload c5shard --steps 10 --variants "parallel,coo_bands=-1;parallel;sell;parallel,coo_bands=32;parallel,coo_bands=64" > gpurun_out/sweep11_c5.txt 2>&1; grep -v "^# device" gpurun_out/sweep11_c5.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "pipelined" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 400 python scripts/sweep.py --workload c2 --steps 30 --variants "parallel;parallel,x_bands=2;parallel,x_bands=4;parallel,x_bands=2,tpr=4;sell,x_bands=2;sell;sell,x_bands=4;csr5,x_bands=2;balanced,x_bands=2;parallel,tpr=1" > gpurun_out/sweep5_c2.txt 2>&1; grep -v "^# device" gpurun_out/sweep5_c2.txt
timeout 400 python scripts/sweep.py --workload c3 --steps 30 --variants "csr5;csr5,x_bands=1;balanced2,x_bands=1;parallel,x_bands=1;sell,x_bands=1;csr5,x_bands=1,csr5_sigma=8" > gpurun_out/sweep5_c3.txt 2>&1; grep -v "^# device" gpurun_out/sweep5_c3.txt

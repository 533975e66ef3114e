#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log
timeout 400 python scripts/sweep.py --workload c4 --steps 30 --variants "parallel;parallel,tpr=2;parallel,tpr=8;balanced;balanced,rb_auto=1;serial" > gpurun_out/sweep8_c4.txt 2>&1; grep -v "^# device" gpurun_out/sweep8_c4.txt
timeout 400 python scripts/sweep.py --workload c1 --flush --steps 30 --variants "parallel;parallel,tpr=2;balanced;balanced,rb_auto=1;serial" > gpurun_out/sweep8_c1.txt 2>&1; grep -v "^# device" gpurun_out/sweep8_c1.txt
timeout 400 python scripts/sweep.py --workload c2 --steps 30 --variants "parallel;parallel,tpr=1;parallel,tpr=4;balanced;balanced,rb_auto=1;parallel,x_bands=2;parallel,x_bands=4" > gpurun_out/sweep8_c2.txt 2>&1; grep -v "^# device" gpurun_out/sweep8_c2.txt
timeout 400 python scripts/sweep.py --workload c3 --steps 30 --variants "parallel;parallel,tpr=1;parallel,tpr=4;serial" > gpurun_out/sweep8_c3.txt 2>&1; grep -v "^# device" gpurun_out/sweep8_c3.txt

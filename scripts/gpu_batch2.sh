#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
V4="parallel;parallel,vec=2;parallel,vec=0;parallel,vec=2,tpr=2;parallel,vec=2,tpr=8;parallel,vec=1,tpr=8;balanced;balanced,vec=2;balanced,vec=2,block_nnz=1024;serial;sell"
V1="parallel;parallel,vec=2;parallel,vec=0;parallel,vec=2,tpr=2;parallel,vec=1,tpr=2;balanced;balanced,vec=2;sell;sell,sell_sigma=32;sell,sell_sigma=1024"
V2="parallel;parallel,vec=2;parallel,vec=2,tpr=4;parallel,vec=2,tpr=1;parallel,vec=1,tpr=4;balanced,vec=2;sell"
timeout 400 python scripts/sweep.py --workload c4 --steps 30 --variants "$V4" > gpurun_out/sweep2_c4.txt 2>&1; echo "sweep c4 rc=$?"
timeout 400 python scripts/sweep.py --workload c1 --flush --steps 30 --variants "$V1" > gpurun_out/sweep2_c1.txt 2>&1; echo "sweep c1 rc=$?"
timeout 400 python scripts/sweep.py --workload c2 --steps 30 --variants "$V2" > gpurun_out/sweep2_c2.txt 2>&1; echo "sweep c2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_dropin.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_dropin.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_dropin.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
V="parallel;parallel,vec=3;parallel,vec=3,tpr=2;sell;sell,sell_variant=1;sell,sell_variant=2;sell,sell_variant=3;sell,sell_variant=4;balanced;balanced,block_nnz=128;balanced,block_nnz=256;csr5,csr5_sigma=8"
timeout 400 python scripts/sweep.py --workload c4 --steps 30 --variants "$V" > gpurun_out/sweep6_c4.txt 2>&1; grep -v "^# device" gpurun_out/sweep6_c4.txt
timeout 400 python scripts/sweep.py --workload c1 --flush --steps 30 --variants "$V" > gpurun_out/sweep6_c1.txt 2>&1; grep -v "^# device" gpurun_out/sweep6_c1.txt
timeout 400 python scripts/sweep.py --workload c2 --steps 30 --variants "parallel;parallel,vec=3;sell;sell,sell_variant=1;sell,sell_variant=2;sell,sell_variant=3;sell,sell_variant=4" > gpurun_out/sweep6_c2.txt 2>&1; grep -v "^# device" gpurun_out/sweep6_c2.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python scripts/sweep.py --workload c5shard --steps 10 --variants "parallel,coo_bands=32;parallel,coo_bands=32,l2_persist=-1,x_window=1;parallel,coo_bands=40,l2_persist=-1,x_window=1;parallel,coo_bands=64,l2_persist=-1,x_window=1" > gpurun_out/sweep15_c5.txt 2>&1; grep -v "^# device" gpurun_out/sweep15_c5.txt

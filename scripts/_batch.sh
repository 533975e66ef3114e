#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 400 python scripts/sweep.py --workload c3 --steps 30 --variants "balanced2;balanced2,tile_items=16;balanced2,tile_items=4;balanced_yid;balanced_yid,tile_items=16;balanced_yid,tile_items=4" > gpurun_out/sweep30_c3.txt 2>&1; grep -v "^# device" gpurun_out/sweep30_c3.txt
timeout 400 python scripts/sweep.py --workload c4 --steps 30 --variants "balanced_yid;balanced_yid,tile_items=16;balanced_yid,tile_items=4" > gpurun_out/sweep30_c4.txt 2>&1; grep -v "^# device" gpurun_out/sweep30_c4.txt
timeout 400 python scripts/sweep.py --workload c2 --steps 30 --variants "balanced_yid;balanced_yid,tile_items=16;balanced_yid,tile_items=4" > gpurun_out/sweep30_c2.txt 2>&1; grep -v "^# device" gpurun_out/sweep30_c2.txt

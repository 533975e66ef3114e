#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest.log
for w in c3 c1 c4 c2; do
  fl=""; [ $w = c1 ] && fl="--flush"
  timeout 400 python scripts/sweep.py --workload $w $fl --steps 30 > gpurun_out/sweep3_$w.txt 2>&1; echo "sweep $w rc=$?"
  grep -v "^# device" gpurun_out/sweep3_$w.txt
done

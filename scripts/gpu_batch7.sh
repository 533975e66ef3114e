#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
V="parallel;parallel,vec=4;balanced;balanced,rb_auto=1;sell;csr5;balanced_yid"
for w in c4 c1 c2 c3; do
  fl=""; [ $w = c1 ] && fl="--flush"
  timeout 400 python scripts/sweep.py --workload $w $fl --steps 30 --variants "$V" > gpurun_out/sweep7_$w.txt 2>&1; grep -v "^# device" gpurun_out/sweep7_$w.txt
done
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log

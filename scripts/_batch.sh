#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log
for w in c1 c2 c3 c4; do
  fl=""; [ $w = c1 ] && fl="--flush"
  timeout 400 python scripts/sweep.py --workload $w $fl --steps 30 > gpurun_out/sweepF_$w.txt 2>&1; grep -v "^# device" gpurun_out/sweepF_$w.txt
done
timeout 600 python scripts/sweep.py --workload c5shard --steps 10 --variants "parallel;parallel,coo_bands=-1" > gpurun_out/sweepF_c5.txt 2>&1; grep -v "^# device" gpurun_out/sweepF_c5.txt
for w in c1 c3 c4; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --power-iters 0 > gpurun_out/benchF_$w.json 2> gpurun_out/benchF_$w.err; echo "bench $w rc=$?"
done

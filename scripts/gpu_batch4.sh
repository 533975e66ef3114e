#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest.log
timeout 400 python scripts/sweep.py --workload c3 --steps 30 --variants "sell;parallel;csr5" > gpurun_out/sweep4_c3.txt 2>&1; grep -v "^# device" gpurun_out/sweep4_c3.txt
for w in c2 c4 c1; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --power-iters 0 > gpurun_out/bench4_$w.json 2> gpurun_out/bench4_$w.err; echo "bench $w rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench4_$w.json").read().strip().splitlines()[-1])
print("$w", "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "e2e", d["e2e"], "launches", d["gpu_launches"], d.get("l2_warm"))
PY
done
SPMV_B200_PIPELINE=0 timeout 600 python bench.py --workload c2 --steps 10 --warmup 5 --no-cpu --power-iters 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 nopipe e2e', d['e2e'])"

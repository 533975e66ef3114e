#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log
timeout 400 python scripts/sweep.py --workload c2 --steps 30 --variants "parallel;sell;balanced" > gpurun_out/sweep9_c2.txt 2>&1; grep -v "^# device" gpurun_out/sweep9_c2.txt
timeout 600 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu --power-iters 5 > gpurun_out/bench9_c2.json 2> gpurun_out/bench9_c2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench9_c2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench9_c2.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
PY

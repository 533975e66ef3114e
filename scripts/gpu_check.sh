#!/bin/bash
# Full GPU check: build, parity tests, smoke, bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "nproc=$(nproc)" >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|Socket|Thread|Core" >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -8 gpurun_out/bench.err; cat gpurun_out/bench.json

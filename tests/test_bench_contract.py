"""The bench contract, as far as it can be checked without a GPU: `bench.py --impl reference` (the reference's own CPU
path on the host cores, here on the 64x smaller --small matrices) prints exactly ONE JSON line on stdout with the keys
the driver reads, and the GPU arm's helpers (ring schedule of the exchange, keyed traffic table) behave."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--small", "--steps", "5", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "spmv_gflops_fp64_csr" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["dtype"] == "f64" and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["same_config"] is True and "workload" in line["config"]


def test_traffic_table_is_keyed_by_workload_and_kernel():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.traffic_for("c2", "csr_vector") == 7292376984
    assert bench.traffic_for("c1", "csr_vector") not in (None, bench.traffic_for("c2", "csr_vector"))
    assert bench.traffic_for("c4", "sell") > 5e9 and bench.traffic_for("c3", "csr5") > 2e9
    assert bench.traffic_for("c1", "no_such_kernel") is None
    table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert all("/" in k for k in table if not k.startswith("_"))

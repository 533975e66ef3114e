#!/usr/bin/env python
"""Timing sweep over (method, option) variants on one device-resident workload.  Development tool:
prints a table to stdout / JSON lines to gpurun_out/sweep.jsonl.  Not the contract bench (bench.py)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from spmv_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--variants", default="parallel;balanced;balanced2;balanced_yid;sell;csr5;serial",
                    help="';'-separated list of method[,opt=val...]")
    ap.add_argument("--profile-only", type=int, default=0, help="run each variant N times without timing (for ncu)")
    ap.add_argument("--flush", action="store_true")
    args = ap.parse_args()
    A, n, vsize, desc, seed, _ = bench.make_workload(args.workload, args.small)
    tdt = torch.float64 if vsize == 8 else torch.float32
    x = torch.empty(n, dtype=tdt, device="cuda")
    api.gen_x(x, n, seed, False, vsize)
    y = torch.zeros(A.m, dtype=tdt, device="cuda")
    bmin, flops = A.min_bytes(), 2.0 * A.nnz
    peak, _ = bench.measured_peak()
    l2 = torch.cuda.get_device_properties(0).L2_cache_size
    fb = torch.zeros(max(2 * l2, 1 << 28), dtype=torch.uint8, device="cuda") if args.flush else None
    print(f"# {desc}: m={A.m} nnz={A.nnz} B_min={bmin / 1e9:.3f} GB, roofline {bmin / peak / 1e6:.3f} ms @ {peak} GB/s")
    yref = None
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "a")
    for var in args.variants.split(";"):
        parts = var.split(",")
        method = parts[0]
        defaults = {}
        for kv in parts[1:]:
            k, v = kv.split("=")
            defaults[k] = api.get_option(k)
            api.set_option(k, int(v))
        try:
            h = A.handle(bench.METHODS[method])
        except RuntimeError as e:
            print(f"{var:40s} FAILED {e}")
            continue
        finally:
            for k, v in defaults.items():
                api.set_option(k, v)
        if yref is None:
            print("# device: L2 %d MiB" % (h.info("dev_l2_bytes") >> 20))
        if args.profile_only:
            for _ in range(args.profile_only):
                h.spmv(x, y)
            torch.cuda.synchronize()
            h.destroy()
            continue
        ms = bench.time_steps(lambda: h.spmv(x, y), args.steps, 5, torch, None, (lambda: fb.max()) if fb is not None else None) / args.steps
        if yref is None:
            yref = y.clone()
            dev = 0.0
        else:
            dev = float((y - yref).abs().max() / yref.abs().max())
        rec = {"workload": args.workload, "variant": var, "kernel": h.kernel, "ms": ms, "gflops": flops / ms / 1e6,
               "gbs": bmin / ms / 1e6, "frac": bmin / ms / 1e6 / peak, "max_rel_dev_vs_first": dev,
               "tpr": h.info("tpr"), "bands": h.info("x_bands")}
        print(f"{var:40s} [{h.kernel:12s}] {ms:8.4f} ms {rec['gflops']:9.1f} GF/s {rec['gbs']:8.1f} GB/s  frac {rec['frac']:.3f}  dev {dev:.1e}")
        out.write(json.dumps(rec) + "\n")
        h.destroy()
    A.destroy()


if __name__ == "__main__":
    main()

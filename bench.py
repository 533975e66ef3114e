#!/usr/bin/env python
"""bench.py -- SpMV GFLOP/s + effective HBM GB/s on the BASELINE.json configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--method parallel]
    python bench.py --impl reference ...      # the reference's own CPU path on the host cores

A "step" is one spmv() over the whole (per-rank) matrix.  N=1 runs BASELINE.json configs[1]
(uniform-random 2^24 x 2^24, 32 nnz/row, fp64: "C2"); N>1 is weak scaling of the same shard: the global
matrix has N*2^24 rows over the same 2^24 columns, every rank owns 2^24 rows (equal nnz = the reference's
splitter), x (128 MiB) is replicated, no collective on the data path.  The y->x power-method loop is a
separate leg on the SQUARE C2 matrix row-sharded over the N GPUs (strong scaling, 2^24/N rows each): once
with an in-place NCCL all-gather timed apart from the SpMV ("power_method") and once with the all-gather
fused into the SpMV epilogue as NVLink peer stores ("power_method_fused").

Other workloads (`--workload c1|c3|c4|c5|c5shard`) time the remaining BASELINE.json configurations the same way;
`c5shard` is one GPU's share of C5 at 8 GPUs (2^25 rows x 2^28 columns) on a single GPU.  `--also` names further
methods timed after the primary one (default: balanced2, sell); their numbers land in "methods".

One JSON line on stdout (rank 0).  `value` = whole-job GFLOP/s with everything resident in HBM;
`e2e` = the same metric through the C-ABI with HOST x / y (pinned), H2D + kernel + D2H inside the timed
region; `roofline` = algorithmic bytes (B_min) / kernel time vs the measured HBM copy peak.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METHODS = {"serial": 0, "parallel": 1, "balanced": 2, "balanced2": 3, "balanced_yid": 4, "sell": 5, "csr5": 6}
LOG2_ROWS_C2 = 24
SAMPLE_LOG2_ROWS = 21  # CPU sample: the first 2^21 rows of C2 (1/8 of the matrix), full-length x


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Native libraries write to file descriptor 1 behind Python's back
# (NCCL prints its version banner there), so descriptor 1 is pointed at stderr for the whole run and the result
# line goes to a private duplicate of the original stdout.
RESULT = None


def claim_stdout():
    global RESULT
    if RESULT is None:
        sys.stdout.flush()
        RESULT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = RESULT if RESULT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def make_workload(name: str, rank: int, world: int, small: bool):
    """Device-resident matrix of this rank + (n_global, dtype size, description)."""
    from spmv_b200 import api, matrices as M
    sh = 6 if small else 0  # --small: 64x fewer rows, for debugging the script itself
    if name == "c2":
        # weak scaling: every GPU gets a full C2's worth of rows (2^24 x 32 nnz) of the (N*2^24) x 2^24
        # matrix; x (128 MiB) is the same replicated vector at every N, so per-GPU work is identical
        rows = 1 << (LOG2_ROWS_C2 - sh)
        n = rows
        A = api.gen_uniform(rows, n, 32, M.SEED_C2, rank * rows, False, 8)
        desc = f"C2 uniform-random {rows * world}x{n}, 32 nnz/row, fp64 CSR ({rows} rows/GPU)"
        return A, n, 8, desc, M.SEED_C2
    if name == "c5":
        rows = 1 << (25 - sh)  # 2^25 rows per GPU: exactly BASELINE.json's C5 (2^28 rows) at 8 GPUs
        n = rows * world
        A = api.gen_uniform(rows, n, 16, M.SEED_C5, rank * rows, False, 8)
        desc = f"C5-family uniform-random {n}x{n}, 16 nnz/row, fp64 CSR ({rows} rows/GPU)"
        return A, n, 8, desc, M.SEED_C5
    if name == "c5shard":
        # exactly one GPU's share of C5 at 8 GPUs, on ONE GPU: 2^25 rows of the 2^28-column matrix (x = 2 GiB)
        rows = 1 << (25 - sh)
        n = rows * 8
        A = api.gen_uniform(rows, n, 16, M.SEED_C5, rank * rows, False, 8)
        return A, n, 8, f"C5 shard: rows [0, {rows}) of uniform-random {n}x{n}, 16 nnz/row, fp64 CSR", M.SEED_C5
    if world != 1:
        raise SystemExit(f"workload {name} is single-GPU")
    if name == "c1":
        g = 1024 >> (sh // 2)
        A = api.gen_laplacian2d(g, g, 8)
        return A, A.n, 8, f"C1 5-point 2-D Laplacian {g}x{g} grid, fp64 CSR", 1
    if name == "c3":
        A = api.gen_rmat(24 - sh, 16, M.SEED_C3, 4)
        return A, A.n, 4, f"C3 R-MAT scale {24 - sh} edge-factor 16 (duplicates kept), fp32 CSR", M.SEED_C3
    if name == "c4":
        g = 256 >> (sh // 3)
        A = api.gen_stencil27(g, g, g, 8)
        return A, A.n, 8, f"C4 27-point stencil {g}^3, fp64 CSR", 4
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU, polled through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log("clock sampler unavailable:", e)

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (device copy, of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def traffic_for(kernel: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# CPU side: the reference (oracle/_ref) or the port, on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample_matrix(small: bool):
    from spmv_b200 import matrices as M
    rows = 1 << (SAMPLE_LOG2_ROWS - (6 if small else 0))
    n = 1 << (LOG2_ROWS_C2 - (6 if small else 0))
    A = M.uniform_random(rows, n, 32, seed=M.SEED_C2)
    x = M.make_x(n, M.SEED_C2, np.float64)
    return A, x, f"rows [0, 2^{int(np.log2(rows))}) of C2 ({A.nnz} nnz, 1/{(n // rows)} of the matrix), full-length x"


def cpu_time_reference(steps: int, warmup: int, small: bool, budget_s: float = 25.0):
    """Best OpenMP+AVX2 method of the reference on all host threads (protocol of
    src/samples/test_spmv.c:87-124: create with nthreads = team size, warm-up, timed spmv() calls)."""
    from oracle import oracle as O
    A, x, sample = cpu_sample_matrix(small)
    flops = 2.0 * A.nnz
    if O.have_reference():
        R = O.Reference()
        # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: override it explicitly,
        # the reference sizes its team from omp_set_num_threads, src/samples/test_spmv.c:88)
        T = max(R.max_threads(), len(os.sched_getaffinity(0)))
        best = None
        # spin the OpenMP team up first: the first ~second of parallel regions in a fresh process runs
        # 100x slow on these hosts (thread creation + cgroup ramp-up) and would bias the method choice
        h = R.create(A.m, A.n, A.rowptr, A.col, A.val, T, 1)
        y = np.zeros(A.m)
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            R.spmv(h, x, y)
        h.destroy()
        for method in (1, 2, 3, 4, 5, 6):
            h = R.create(A.m, A.n, A.rowptr, A.col, A.val, T, method)
            y = np.zeros(A.m)
            R.spmv(h, x, y)
            reps, dt = 3, 1e9
            for _ in range(reps):
                t0 = time.perf_counter()
                R.spmv(h, x, y)
                dt = min(dt, time.perf_counter() - t0)
            log(f"  cpu reference {O.Reference.__name__} method {method}: {dt * 1e3:.2f} ms")
            if best is None or dt < best[1]:
                if best:
                    best[2].destroy()
                best = (method, dt, h)
            else:
                h.destroy()
        method, dt, h = best
        steps = max(3, min(steps, int(budget_s / max(dt, 1e-4))))
        y = np.zeros(A.m)
        for _ in range(warmup):
            R.spmv(h, x, y)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            R.spmv(h, x, y)
            times.append(time.perf_counter() - t0)
        h.destroy()
        kind, cores, name = "reference", T, ["", "Method_Parallel", "Method_Balanced", "Method_Balanced2",
                                             "Method_BalancedYid", "Method_SellCSigma", "Method_Csr5Spmv"][method]
    else:
        P = O.Port()
        T = len(os.sched_getaffinity(0))
        C_omp = __import__("ctypes").CDLL("libgomp.so.1")
        C_omp.omp_set_num_threads(T)
        y = P.spmv_serial(A.rowptr, A.col, A.val, x, parallel=True)
        t0 = time.perf_counter()
        P.spmv_serial(A.rowptr, A.col, A.val, x, parallel=True)
        dt = time.perf_counter() - t0
        steps = max(3, min(steps, int(budget_s / max(dt, 1e-4))))
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            P.spmv_serial(A.rowptr, A.col, A.val, x, parallel=True)
            times.append(time.perf_counter() - t0)
        kind, cores, name = "port", T, "oracle_spmv_parallel_d"
    avg = float(np.mean(times))
    return {"value": flops / avg / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": sample + f"; best method {name}; {steps} timed calls, avg {avg * 1e3:.2f} ms, best {min(times) * 1e3:.2f} ms",
            "best_value": flops / min(times) / 1e9}, steps, avg


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs and prints; the others exit 0 without work
    cb, steps, avg = cpu_time_reference(args.steps, args.warmup, args.small, budget_s=60.0)
    line = {"metric": "spmv_gflops_fp64_csr", "value": cb["value"], "unit": "GFLOP/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": avg * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 uniform-random 16777216x16777216, 32 nnz/row, fp64 CSR (CPU: bounded sample)",
                       "sample": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def time_steps(fn, steps, warmup, torch, dist, flush=None):
    """W untimed + K timed steps, CUDA events on the launching stream, barrier + sync on both sides,
    max over ranks.  With `flush`, L2 is overwritten before every timed step and steps are timed one by one."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        ms = 0.0
        for _ in range(steps):
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    torch.cuda.synchronize()
    return ms


def run_gpu_arm(args):
    import torch
    from spmv_b200 import api, build
    build.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # belt and braces: see RESULT below
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    dev = torch.device("cuda", local)

    A, n, vsize, desc, seed = make_workload(args.workload, rank, world, args.small)
    tdt = torch.float64 if vsize == 8 else torch.float32
    m, nnz = A.m, A.nnz
    bmin = A.min_bytes()
    flops = 2.0 * nnz
    x = torch.empty(n, dtype=tdt, device=dev)
    api.gen_x(x, n, seed, False, vsize)
    y = torch.zeros(m, dtype=tdt, device=dev)
    l2_bytes = torch.cuda.get_device_properties(local).L2_cache_size
    fits_l2 = bmin < 2 * l2_bytes
    flush_buf = torch.zeros(max(2 * l2_bytes, 1 << 28), dtype=torch.uint8, device=dev) if fits_l2 else None
    # read-only flush: evicts the working set without leaving dirty lines to be written back under the timer
    flush = (lambda: flush_buf.max()) if fits_l2 else None

    primary = args.method
    names = [primary] + [mname for mname in args.also.split(",") if mname and mname != primary]
    results, handles = {}, {}
    sampler = None
    launches = 0
    for mname in names:
        t0 = time.perf_counter()
        h = A.handle(METHODS[mname])
        torch.cuda.synchronize()
        create_ms = (time.perf_counter() - t0) * 1e3
        handles[mname] = h
        fn = lambda h=h: h.spmv(x, y)  # noqa: E731
        if mname == primary:
            sampler = ClockSampler(local)
            for _ in range(args.warmup):
                fn()
            torch.cuda.synchronize()
            sampler.start()
            l0 = api.launch_count()
            ms = time_steps(fn, args.steps, 0, torch, dist, flush)
            launches = api.launch_count() - l0
            clocks = sampler.result()
        else:
            ms = time_steps(fn, args.steps, args.warmup, torch, dist, flush)
        per = ms / args.steps
        results[mname] = {"kernel": h.kernel, "ms_per_step": per, "gflops": world * flops / per / 1e6,
                          "gbs_effective_per_gpu": bmin / per / 1e6, "create_ms": create_ms}
        log(f"[rank {rank}] {mname:13s} [{h.kernel}] {per:.4f} ms/step  {world * flops / per / 1e6:9.1f} GFLOP/s  "
            f"{bmin / per / 1e6:8.1f} GB/s eff/GPU  (create {create_ms:.1f} ms)")

    # L2-warm figure for matrices that fit L2 (C1): reported next to the flushed headline
    warm = None
    if fits_l2:
        ms = time_steps(lambda: handles[primary].spmv(x, y), args.steps, args.warmup, torch, dist, None)
        warm = {"ms_per_step": ms / args.steps, "gflops": world * flops / (ms / args.steps) / 1e6}

    # ---- e2e: host (pinned) x and y through the same C-ABI call ----
    hx = torch.empty(n, dtype=tdt, pin_memory=True)
    hx.copy_(x)
    hy = torch.empty(m, dtype=tdt, pin_memory=True)
    h = handles[primary]
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        h.spmv(hx, hy)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.spmv(hx, hy)  # H2D x, kernel, D2H y, stream sync -- all inside
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_per = e2e_ms / e2e_steps
    y_chk = torch.empty_like(y)
    h.spmv(x, y_chk)
    torch.cuda.synchronize()
    assert torch.equal(y_chk.cpu(), hy), "host-pointer path and device-pointer path disagree"

    # ---- power method on the SQUARE matrix of the workload, row-sharded over the ranks (strong scaling) ----
    power = fused = None
    if args.power_iters > 0 and args.workload in ("c2", "c5"):
        from spmv_b200 import multigpu as G
        if world == 1 or args.workload == "c5":
            Ap, hp, mp = A, handles[primary], m          # the main shard already is a slice of a square matrix
        else:
            mp = n // world                               # rows [rank*n/N, (rank+1)*n/N) of the SAME C2 matrix
            Ap = api.gen_uniform(mp, n, 32, seed, rank * mp, False, vsize)
            hp = Ap.handle(METHODS[primary])
        split = [g * mp for g in range(world + 1)]
        xs = x * (1.0 / 16.0)  # un-normalised loop: keep 50 iterations far inside the fp64 range
        pm = G.PowerMethod(lambda xf, ys: hp.spmv(xf, ys), split, xs)
        pm.run(2)
        x_nccl, t_spmv, t_comm = pm.run(args.power_iters)
        if dist is not None:
            t = torch.tensor([t_spmv, t_comm], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_spmv, t_comm = (float(v) for v in t.tolist())
        power = {"iters": args.power_iters, "rows_per_gpu": mp, "kernel": hp.kernel, "spmv_ms_per_iter": t_spmv,
                 "allgather_ms_per_iter": t_comm, "allgather_bytes_recv_per_gpu": (world - 1) * mp * vsize,
                 "collective": "ncclAllGather in place (torch.distributed)" if world > 1 else "none (1 GPU)",
                 "normalised": False}
        try:
            fp = G.FusedPowerMethod(hp, split, xs)
            fp.run(2)
            fp2 = G.FusedPowerMethod(hp, split, xs)
            t_f, t_sync = fp2.run(args.power_iters + 2)
            same = bool(torch.equal(fp2.result(), x_nccl))
            if dist is not None:
                t = torch.tensor([t_f, t_sync, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_f, t_sync, bad = (float(v) for v in t.tolist())
                same = bad == 0.0
            fused = {"iters": args.power_iters + 2, "spmv_plus_peer_stores_ms_per_iter": t_f, "rank_sync_ms_per_iter": t_sync,
                     "peer_bytes_sent_per_gpu": (world - 1) * mp * vsize, "bitwise_equal_to_nccl_loop": same,
                     "transport": "st.global to CUDA-IPC peer mappings over NVLink from the SpMV epilogue"}
            fp.close()
            fp2.close()
        except Exception as e:  # IPC may be unavailable in some containers: the NCCL leg stands
            log("fused power method unavailable:", repr(e))
        if hp is not handles[primary]:
            hp.destroy()
            Ap.destroy()

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only, bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.workload == "c2":
        try:
            cpu, _, _ = cpu_time_reference(20, 3, args.small)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the GPU numbers stand on their own
            log("cpu baseline failed:", repr(e))

    if rank == 0:
        r = results[primary]
        per = r["ms_per_step"]
        peak, peak_src = measured_peak()
        achieved = bmin / per / 1e6  # GB/s, algorithmic bytes of ONE launch / its average duration
        line = {
            "metric": "spmv_gflops_fp64_csr" if vsize == 8 else "spmv_gflops_fp32_csr",
            "value": r["gflops"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if vsize == 8 else "f32", "data": "synthetic",
            "config": {"workload": desc, "method": api.METHOD_NAMES[METHODS[primary]], "kernel": r["kernel"],
                       "m_per_gpu": m, "n": n, "nnz_per_gpu": nnz, "min_bytes_per_gpu": bmin,
                       "l2": ("L2 flushed (%d MiB read of a scratch buffer) before every timed step" % (flush_buf.numel() >> 20)) if fits_l2
                       else "inputs larger than L2 (%.1f GB per step vs %d MiB L2); no flush" % (bmin / 1e9, l2_bytes >> 20),
                       "timing": "CUDA events on the launch stream around K back-to-back spmv() calls, max over ranks"},
            "gbs_effective": world * achieved,
            "frac_of_8TBps_nominal": achieved / 8000.0,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_for(r["kernel"]), "kernel": r["kernel"], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bmin},
            "methods": results,
            "e2e": {"value": world * flops / e2e_per / 1e6, "unit": "GFLOP/s", "ms_per_step": e2e_per,
                    "h2d_bytes_per_step": n * vsize, "d2h_bytes_per_step": m * vsize, "steps": e2e_steps,
                    "api": "spmv() of include/spmv.h with pinned HOST x and y"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if warm:
            line["l2_warm"] = warm
        if power:
            line["power_method"] = power
        if fused:
            line["power_method_fused"] = fused
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    for h in handles.values():
        h.destroy()
    A.destroy()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "c5shard"])
    ap.add_argument("--method", default="parallel", choices=list(METHODS))
    ap.add_argument("--also", default="balanced2,sell", help="comma list of further methods timed after the primary")
    ap.add_argument("--power-iters", type=int, default=50)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--small", action="store_true", help="64x smaller matrices (script debugging only; not a bench)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

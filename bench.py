#!/usr/bin/env python
"""bench.py -- NMF iterations/sec on synthetic sparse counts (BASELINE.json: 30k genes x 1M cells,
5 % density, run_nmf k=32) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One *step* is one ALS iteration (H update, scale, W update, scale, cor: reference
src/singlet.cpp:647-664). N > 1 is launched by torchrun, one rank per GPU: cells are sharded for the
H update, genes for the W update (singlet_b200/sharded.py); the total problem is fixed, so the
scaling is "strong". Rank 0 prints ONE JSON line.

* ``value``: iterations/sec with A/At resident in HBM (generated on the device, bit-identical to the
  numpy generator), K steps timed with CUDA events between barriers, max over ranks.
* ``e2e``: the same metric through the host-facing C ABI call ``sgl_nmf`` (what the Rcpp glue binds)
  with HOST dgCMatrix buffers: upload of A and At, K iterations, download of w/d/h all inside the
  timed region.
* ``roofline``: the dominant kernel (the tiled SpMM of the H update and W update), algorithmic bytes
  per launch (8*nnz + 4*(ncol+1) + 4*k*nrow + 4*k*ncol) / its CUDA-event duration
  measured live on the launching stream, against MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the reference's own OpenMP implementation (oracle/_ref, the
  reference's functions compiled from /root/reference; else the oracle port) on the host cores, on a
  bounded column sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on
    "c3": dict(m=30000, n=1000000, density=0.05, k=32, name="synthetic 30k genes x 1M cells, 5% density, run_nmf k=32"),
    # BASELINE.json configs[4]: atlas scale, needs several GPUs (4.2e9 non-zeros: beyond one dgCMatrix's int32 pointers)
    "c5": dict(m=35000, n=4000000, density=0.03, k=64, name="synthetic 35k genes x 4M cells, 3% density, k=64 (atlas scale)"),
    # smaller stand-ins for quick checks (never the default)
    "c5shard": dict(m=35000, n=500000, density=0.03, k=64, name="one eighth of config 5 (35k x 500k, 3%, k=64; tuning only)"),
    "mid": dict(m=30000, n=100000, density=0.05, k=32, name="MID 30k x 100k (not a bench config)"),
    "mini": dict(m=3000, n=20000, density=0.05, k=32, name="MINI 3k x 20k (not a bench config)"),
    "c4shape": dict(m=20000, n=250000, density=0.08, k=16, name="synthetic 20k x 250k, 8% density, k=16 (shape of config 4)"),
}
L1, L2 = 0.01, 0.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample-cells", type=int, default=0, help="cells in the CPU sample (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(config_name, world):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if not captured for
    this configuration)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as fh:
            t = json.load(fh)
        if t.get("config") == config_name and t.get("n_gpus") == world:
            mean_read = 0.5 * (t["dram_bytes_read_per_launch"] + t["w_update_launch"]["dram_bytes_read"])
            mean_write = 0.5 * (t["dram_bytes_write_per_launch"] + t["w_update_launch"]["dram_bytes_write"])
            return mean_read + mean_write
    except Exception:
        pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_leg(cfg, steps, warmup, sample_cells, backend=None):
    """Time the reference's CPU implementation on a bounded column sample: `sample_cells` cells of the
    same synthetic matrix (all genes). Returns the cpu_baseline dict; value is extrapolated to the
    full cell count (every term of an iteration but the m NNLS solves of the W update is
    proportional to the number of cells)."""
    import scipy.sparse as sp

    from oracle.pyoracle import Oracle, have_reference
    from singlet_b200 import synth

    kind = "reference" if have_reference() else "port"
    orc = Oracle(kind)
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every host core explicitly
    cores = max(orc.max_threads(), os.cpu_count() or 1)
    m, n, dens, k = cfg["m"], cfg["n"], cfg["density"], cfg["k"]
    if sample_cells <= 0:
        # ~1.5e6 non-zeros per core and per iteration keeps one iteration in the seconds range
        sample_cells = int(min(n, max(2000, cores * 1.2e6 / (dens * m))))
    if backend is not None:
        h = backend.synth(m, n, dens, synth.DATA_SEED, 0, 0, sample_cells, synth.values_table(m, dens))
        p, i, x, _, _ = backend.matrix_to_host(h)
    else:
        p, i, x = synth.synth_csc(m, n, dens, synth.DATA_SEED, 0, sample_cells)
    A = sp.csc_matrix((x, i, p), shape=(m, sample_cells))
    At = A.T.tocsc()
    At.sort_indices()
    w0 = synth.w_init(k, m)
    times = []
    for rep in range(max(1, min(warmup, 1)) + max(1, min(steps, 3))):
        t0 = time.perf_counter()
        orc.nmf(A, At, w0, tol=0.0, maxit=1, L1=(L1, L1), L2=(L2, L2), threads=cores)
        times.append(time.perf_counter() - t0)
    per_iter = float(np.median(times[1:])) if len(times) > 1 else times[0]
    its_sample = 1.0 / per_iter
    scale = sample_cells / float(n)
    try:
        with open("/proc/cpuinfo") as fh:
            cpu_model = [ln.split(":", 1)[1].strip() for ln in fh if ln.startswith("model name")][0]
    except Exception:
        cpu_model = "unknown"
    return {"value": its_sample * scale, "unit": "iterations/s", "cores": cores, "kind": kind, "cpu_model": cpu_model,
            "sample": f"{m} genes x {sample_cells} cells of the same synthetic matrix ({A.nnz} non-zeros), one c_nmf iteration "
                      f"= {per_iter:.3f} s on {cores} threads; value = sample it/s x {scale:.4g} (time per iteration is "
                      f"proportional to the cell count)",
            "sample_iterations_per_s": its_sample}


# ------------------------------------------------------------------------------------------------
_JSON_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but C libraries print there too (NCCL prints "NCCL version ..." under
    NCCL_DEBUG=VERSION): keep a private duplicate of the real stdout for the JSON line and point fd 1 at stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(obj) + "\n").encode())


def main():
    args = parse()
    _claim_stdout()
    cfg = CONFIGS[args.config]
    m, n, dens, k = cfg["m"], cfg["n"], cfg["density"], cfg["k"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    base_cfg = {"workload": cfg["name"], "m_genes": m, "n_cells": n, "density": dens, "k": k, "L1": L1, "L2": L2,
                "tol": 0.0, "parallelism": f"cells sharded over {world} GPU(s); W-update RHS reduce-scattered, W all-gathered (layout B)",
                "cache_policy": "inputs larger than L2 (A and At streams are GBs per half-iteration)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        backend = None
        try:
            import torch

            if torch.cuda.is_available():
                from singlet_b200.sharded import CudaBackend

                backend = CudaBackend(0)
        except Exception:
            backend = None
        cb = cpu_reference_leg(cfg, args.steps, args.warmup, args.cpu_sample_cells, backend)
        v = cb["value"]
        out = {"impl": "reference", "metric": "nmf_iterations_per_sec", "value": v, "unit": "iterations/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": base_cfg, "cpu_baseline": cb,
               "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(out)
        return 0

    import torch
    import torch.distributed as dist

    from singlet_b200 import synth
    from singlet_b200.sharded import CudaBackend, ShardedNMF, shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    be = CudaBackend(local_rank)
    table = synth.values_table(m, dens)
    c0, c1, _ = shard_bounds(n, world, rank)
    g0, g1, _ = shard_bounds(m, world, rank)
    t_gen = time.perf_counter()
    # layout "B" (sharded.py): this rank's cells for the H update, and the transpose of the SAME cell block
    # (local cells x all genes) for its partial W-update right-hand sides
    A_sh = be.synth(m, n, dens, synth.DATA_SEED, 0, c0, c1 - c0, table)
    At_sh = be.synth_block(m, n, dens, synth.DATA_SEED, 1, 0, m, c0, c1 - c0, table)
    be.synchronize()
    t_gen = time.perf_counter() - t_gen
    nnz_A, nnz_At = be.matrix_info(A_sh)[2], be.matrix_info(At_sh)[2]
    fit = ShardedNMF(be, m, n, k, A_sh, At_sh, rank, world, group, layout="B")
    fit.set_w(synth.w_init(k, m))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        fit.iteration(L1, L1, L2, L2)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    be.profile(True)
    be.profile_read()
    launches0 = be.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    tol = None
    for _ in range(args.steps):
        tol = fit.iteration(L1, L1, L2, L2)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    prof = be.profile_read()
    be.profile(False)
    launches = be.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    stats = torch.tensor([ms, float(launches), float(nnz_A), float(nnz_At), prof["spmm"][0], float(prof["spmm"][2]),
                          prof["nnls"][0], prof["gram"][0]], dtype=torch.float64, device=be.device)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = stats, stats
    ms_max = float(mx[0])
    total_launches = int(sm[1])
    nnz_total = int(sm[2])

    out = None
    if rank == 0:
        value = args.steps / (ms_max / 1000.0)
        peak, peak_src = measured_peak()
        # dominant kernel: the SpMM launches of this rank (2 per iteration: A shard, At shard)
        spmm_ms, spmm_cnt, spmm_bytes = prof["spmm"]
        per_launch_ms = spmm_ms / max(spmm_cnt, 1)
        per_launch_bytes = spmm_bytes / max(spmm_cnt, 1)
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        iter_bytes = synth.algorithmic_bytes_per_iter(m, n, nnz_total, k)
        out = {"metric": "nmf_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic (device-generated, bit-identical to singlet_b200/synth.py)",
               "config": dict(base_cfg, nnz=nnz_total, generate_s=round(t_gen, 3), final_tol=tol),
               "gpu_launches": total_launches, "clocks": clocks,
               "roofline": {"bound": "hbm", "kernel": "spmm_stream_kernel<32> (mean of the H-update and W-update launches, rank 0)",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": ncu_traffic(args.config, world),
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch_bytes,
                            "ms_per_launch": per_launch_ms, "launches_timed": spmm_cnt},
               "breakdown_ms_per_step_rank0": {"spmm": spmm_ms / args.steps, "nnls": prof["nnls"][0] / args.steps,
                                               "gram": prof["gram"][0] / args.steps},
               "whole_iteration": {"algorithmic_bytes": iter_bytes, "achieved_gbs": iter_bytes * value / 1e9,
                                   "frac_of_peak": iter_bytes * value / 1e9 / peak}}

    # ---- e2e: host buffers through the C ABI (rank 0 drives all N GPUs' worth of data only at N = 1) ----
    e2e = None
    if not args.no_e2e:
        try:
            if world == 1:
                e2e = e2e_leg(be, cfg, args.steps, A_sh, At_sh)
            else:
                e2e = e2e_leg_sharded(be, cfg, args.steps, A_sh, At_sh, rank, world, group)
        except MemoryError as ex:
            e2e = {"value": None, "unit": "iterations/s", "error": f"host memory: {ex}"}
    if rank == 0:
        out["e2e"] = e2e if e2e is not None else {"value": None, "unit": "iterations/s", "h2d_bytes_per_step": None,
                                                  "d2h_bytes_per_step": None, "note": "skipped (--no-e2e)"}
        if not args.no_cpu:
            be_cpu = be
            out["cpu_baseline"] = cpu_reference_leg(cfg, args.steps, args.warmup, args.cpu_sample_cells, be_cpu)
        _emit(out)
    be.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def e2e_leg_sharded(be, cfg, steps, A_dev, At_dev, rank, world, group):
    """N > 1: every rank starts from HOST dgCMatrix shards (its cells, and the transpose of that block), uploads
    them, runs `steps` iterations of the sharded fit and downloads the replicated model. Max over ranks."""
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist

    from singlet_b200 import synth
    from singlet_b200.sharded import ShardedNMF, shard_bounds

    m, n, k = cfg["m"], cfg["n"], cfg["k"]
    c0, c1, _ = shard_bounds(n, world, rank)
    pA = be.matrix_to_host(A_dev)
    A = sp.csc_matrix((pA[2], pA[1], pA[0]), shape=(m, c1 - c0))
    A.has_sorted_indices = True
    w0 = synth.w_init(k, m)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    hA = be.upload(A)  # the transpose of the local block is built on the device (sgl_matrix_transpose)
    fit = ShardedNMF(be, m, n, k, hA, None, rank, world, group, layout="B")
    fit.set_w(w0)
    for _ in range(steps):
        fit.iteration(L1, L1, L2, L2)
    w, d, h = fit.factors_to_host()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=be.device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + w0.nbytes)], dtype=torch.float64, device=be.device)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
    d2h = float(world) * (w.nbytes + h.nbytes + d.nbytes + steps * 40)
    sec = float(dt[0])
    return {"value": steps / sec, "unit": "iterations/s", "h2d_bytes_per_step": float(h2d[0]) / steps, "d2h_bytes_per_step": d2h / steps,
            "seconds_total": sec, "iterations": steps,
            "note": "sharded public API (singlet_b200.sharded): every rank uploads its host dgCMatrix cell shard (FP64) and transposes "
                    "it on the device, K iterations, replicated w/d/h downloaded on every rank; max over ranks"}


def e2e_leg(be, cfg, steps, A_dev, At_dev):
    """sgl_nmf with HOST dgCMatrix buffers: upload A and At, `steps` iterations (tol = 0), download."""
    import ctypes as C

    import scipy.sparse as sp

    from singlet_b200 import _lib, api, synth

    m, n, k = cfg["m"], cfg["n"], cfg["k"]
    nnz = be.matrix_info(A_dev)[2]
    need_gb = 2 * nnz * 12 / 1e9
    try:
        with open("/proc/meminfo") as fh:
            avail_kb = [int(line.split()[1]) for line in fh if line.startswith("MemAvailable")][0]
    except Exception:
        avail_kb = 0
    if avail_kb and avail_kb / 1e6 < need_gb * 1.3 + 8:
        raise MemoryError(f"need ~{need_gb:.0f} GB of host RAM for A and At as dgCMatrix, {avail_kb / 1e6:.0f} GB available")
    pA = be.matrix_to_host(A_dev)
    pAt = be.matrix_to_host(At_dev)
    A = sp.csc_matrix((pA[2], pA[1], pA[0]), shape=(m, n))
    At = sp.csc_matrix((pAt[2], pAt[1], pAt[0]), shape=(n, m))
    A.has_sorted_indices = At.has_sorted_indices = True  # generated in order; skips scipy's O(nnz) check
    w0 = synth.w_init(k, m)
    h = api.Handle(be.device.index)
    h.set_cache(False)
    # warm-up on a tiny problem (module load, allocator) -- not the timed call
    As = synth.synth_scipy(2000, 1500, 0.05)
    Ats = As.T.tocsc()
    Ats.sort_indices()
    api.c_nmf(As, Ats, 0.0, 2, False, L1, L1, L2, L2, 0, synth.w_init(k, 2000), h)
    t0 = time.perf_counter()
    res = api.c_nmf(A, At, 0.0, steps, False, L1, L1, L2, L2, 0, w0, h)
    dt = time.perf_counter() - t0
    assert res["iter"] == steps
    h2d = (A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + At.data.nbytes + At.indices.nbytes + At.indptr.nbytes
           + w0.nbytes)
    d2h = res["w"].nbytes + res["h"].nbytes + res["d"].nbytes + steps * 40
    # informational: the same call with At = NULL (row f1: the library transposes A on the device, so only A crosses PCIe)
    t0 = time.perf_counter()
    res2 = api.c_nmf(A, None, 0.0, steps, False, L1, L1, L2, L2, 0, w0, h)
    dt2 = time.perf_counter() - t0
    same = bool(np.array_equal(res["w"], res2["w"]) and np.array_equal(res["h"], res2["h"]))
    h.close()
    return {"value": steps / dt, "unit": "iterations/s", "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps,
            "seconds_total": dt, "iterations": steps,
            "device_transpose": {"value": steps / dt2, "seconds_total": dt2, "identical_model": same,
                                 "h2d_bytes_per_step": (A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + w0.nbytes) / steps},
            "note": "one sgl_nmf call: FP64 dgCMatrix A and At (pageable host memory) packed to 8-byte records by host "
                    "threads and uploaded through a pinned ring, K iterations, w/d/h downloaded; h2d_bytes_per_step counts "
                    "the host buffers handed to the call (12 B per non-zero), divided by K"}


if __name__ == "__main__":
    sys.exit(main())

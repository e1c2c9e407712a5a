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
* ``e2e``: the same metric through the public API call a user of the reference makes -- ``run_nmf(A, rank, tol,
  maxit)`` (R/run_nmf.R:18; here singlet_b200.api.run_nmf on the C ABI) -- with a HOST dgCMatrix: upload of A, the
  transpose (on the device), K iterations, download and sort of w/d/h all inside the timed region. The raw C ABI call
  ``sgl_nmf(A, At)`` with both host matrices (what src/RcppExports.cpp:97-116 receives) is reported beside it.
* ``roofline``: the dominant kernel (the tiled SpMM of the H update and W update), algorithmic bytes
  per launch (8*nnz + 4*(ncol+1) + 4*k*nrow + 4*k*ncol) / its CUDA-event duration
  measured live on the launching stream, against MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the reference's own OpenMP implementation (oracle/_ref, the
  reference's functions compiled from /root/reference against a scalar Eigen shim; else the oracle port) on the host
  cores, on bounded column samples of the same workload: warm iterations are timed at TWO sample sizes, the time per
  iteration is fitted as a * cells + b (b = the m NNLS solves of the W update, which do not scale with the cell count)
  and evaluated at the full cell count. The reference arm imports nothing of singlet_b200 (no CUDA library is loaded).
* ``c1_run_nmf`` / ``cv_sweep`` / ``c4_ard_nmf`` / ``masked_als`` (N = 1): BASELINE's other configs through the public API -- the second
  half of the metric ("CV rank-sweep wall time") with the reference CPU timed on a sample of the same fits.
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
    "c3r8": dict(m=30000, n=125000, density=0.05, k=32, name="one rank's eighth of config 2's cells (30k x 125k; profiling only)"),
    "c3r4": dict(m=30000, n=250000, density=0.05, k=32, name="one rank's quarter of config 2's cells (30k x 250k; profiling only)"),
    "c3r2": dict(m=30000, n=500000, density=0.05, k=32, name="one rank's half of config 2's cells (30k x 500k; profiling only)"),
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
    ap.add_argument("--cpu-sample-cells", type=int, default=0, help="cells in the larger CPU sample (0 = auto)")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[0]/[1]/[3] legs (c1_run_nmf, cv_sweep, c4_ard_nmf)")
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


def ncu_traffic(config_name, world, kernel):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum, mean of the H-update
    and W-update launches) from the committed `ncu --set full` capture of this command (profiles/r2_traffic.json, written
    from the .ncu-rep by scripts/ncu_traffic.py); None when no capture exists for this configuration and kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as fh:
            t = json.load(fh)
        for e in t["captures"]:
            if e["config"] == config_name and e["n_gpus"] == world and e["kernel"].split("(")[0] == kernel.split("(")[0]:
                return e["dram_bytes_per_launch"]
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
def _load_synth_standalone():
    """singlet_b200/synth.py (pure numpy) loaded by path, WITHOUT importing the singlet_b200 package: the reference arm
    must not load libsinglet_cuda.so or touch the GPU."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_synth_standalone", os.path.join(ROOT, "singlet_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_leg(cfg, timed_iters, sample_cells):
    """Time the reference's CPU implementation on bounded column samples of the same synthetic matrix (all genes).

    For each of two sample sizes n1 < n2 the fit is run twice from the same w_init, for 1 and for 1 + timed_iters
    iterations; the difference is `timed_iters` WARM iterations (warm-started h, like every iteration but the first).
    Every term of an iteration is proportional to the number of cells except the m NNLS solves (and Gram / scale / cor)
    of the W update, so t_iter(n) = a * n + b is fitted through the two samples and evaluated at the full cell count.
    """
    import scipy.sparse as sp

    from oracle.pyoracle import Oracle, have_reference

    synth = _load_synth_standalone()
    kind = "reference" if have_reference() else "port"
    orc = Oracle(kind)
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every host core explicitly
    cores = max(orc.max_threads(), os.cpu_count() or 1)
    m, n, dens, k = cfg["m"], cfg["n"], cfg["density"], cfg["k"]
    n2 = int(min(n, sample_cells))
    n1 = max(1000, n2 // 3)
    t_gen = time.perf_counter()
    p, i, x = synth.synth_csc(m, n, dens, synth.DATA_SEED, 0, n2)
    A2 = sp.csc_matrix((x, i, p), shape=(m, n2))
    t_gen = time.perf_counter() - t_gen
    w0 = synth.w_init(k, m)
    pts = []
    for ns in (n1, n2):
        A = A2[:, :ns].tocsc() if ns < n2 else A2
        At = A.T.tocsc()
        At.sort_indices()
        t0 = time.perf_counter()
        orc.nmf(A, At, w0, tol=0.0, maxit=1, L1=(L1, L1), L2=(L2, L2), threads=cores)
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.nmf(A, At, w0, tol=0.0, maxit=1 + timed_iters, L1=(L1, L1), L2=(L2, L2), threads=cores)
        t_all = time.perf_counter() - t0
        pts.append({"cells": ns, "nnz": int(A.nnz), "first_iteration_s": t_first, "warm_iteration_s": (t_all - t_first) / timed_iters})
    a = (pts[1]["warm_iteration_s"] - pts[0]["warm_iteration_s"]) / float(n2 - n1)
    b = pts[1]["warm_iteration_s"] - a * n2
    if a <= 0 or b < 0:  # timing noise on a tiny sample: fall back to proportional scaling of the larger one
        a, b = pts[1]["warm_iteration_s"] / n2, 0.0
    t_full = a * n + b
    try:
        with open("/proc/cpuinfo") as fh:
            cpu_model = [ln.split(":", 1)[1].strip() for ln in fh if ln.startswith("model name")][0]
    except Exception:
        cpu_model = "unknown"
    return {"value": 1.0 / t_full, "unit": "iterations/s", "cores": cores, "kind": kind, "cpu_model": cpu_model,
            "build": "reference src/singlet.cpp hot-path functions compiled with g++ -O2 -fopenmp against a scalar Eigen/Rcpp "
                     "shim (oracle/shim): Eigen's SIMD kernels are NOT used, OpenMP over columns as in the reference" if kind == "reference"
                     else "oracle port (g++ -O2 -fopenmp, scalar)",
            "sample": f"{m} genes x {n1} and x {n2} cells of the same synthetic matrix; {timed_iters} warm c_nmf iteration(s) timed at each "
                      f"size on {cores} threads; seconds per iteration fitted as a*cells + b and evaluated at {n} cells",
            "samples": pts, "fit": {"a_s_per_cell": a, "b_s_fixed": b, "seconds_per_iteration_at_full_size": t_full},
            "generate_s": round(t_gen, 2)}


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
        # ~125k cells unless told otherwise; 3 warm iterations (--steps caps it)
        cells = args.cpu_sample_cells if args.cpu_sample_cells > 0 else min(n, 125000)
        cb = cpu_reference_leg(cfg, max(1, min(args.steps, 3)), cells)
        v = cb["value"]
        out = {"impl": "reference", "metric": "nmf_iterations_per_sec", "value": v, "unit": "iterations/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": dict(base_cfg, parallelism="OpenMP over columns on the host cores (the reference has no GPU or multi-process path)"),
               "cpu_baseline": cb, "gpu_launches": 0,
               "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(out)
        return 0

    import torch
    import torch.distributed as dist

    from singlet_b200 import synth
    from singlet_b200.multi import RankComm, RankFit
    from singlet_b200.sharded import CudaBackend, shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:  # the ranks share the host cores: cap every rank's upload packing threads
        os.environ.setdefault("SGL_UPLOAD_THREADS", str(max(2, (os.cpu_count() or 8) // world)))
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
    # the sharded fit and every collective of it live in the C++ library (csrc/multi.cu: sgl_comm / sgl_fit, NCCL on the
    # library's stream); torch.distributed only hands the NCCL unique id round and carries the timing reductions below
    comm = RankComm(be._h, local_rank, world, rank, group)
    fit = RankFit(comm, A_sh, At_sh, n, k, synth.w_init(k, m))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        fit.iterate(L1, L1, L2, L2)
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
    coll0 = comm.collectives()
    for _ in range(args.steps):
        tol = fit.iterate(L1, L1, L2, L2)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    prof = be.profile_read()
    be.profile(False)
    launches = be.launch_count() - launches0
    n_coll = comm.collectives() - coll0
    fit.close()
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
        kp = be.kp(k)
        mixed = int(be.lib.sgl_get_precision(be._h)) == 0 and kp >= 32
        kernel = f"spmm_h16_kernel<{kp}>" if mixed else f"spmm_stream_kernel<{kp}>"
        out = {"metric": "nmf_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "precision": ("FP32 accumulation everywhere; sparse-product operands staged as FP16 scaled by powers of two (sgl_set_precision "
                             "MIXED16, the default)" if mixed else "FP32 operands and accumulation (SGL_PRECISION=fp32)"),
               "data": "synthetic (device-generated, bit-identical to singlet_b200/synth.py)",
               "config": dict(base_cfg, nnz=nnz_total, generate_s=round(t_gen, 3), final_tol=tol),
               "gpu_launches": total_launches, "nccl_collectives_per_step_per_rank": n_coll / args.steps, "clocks": clocks,
               "driver": "csrc/multi.cu sgl_fit_iterate through the C ABI (one process per GPU, NCCL issued by the library)",
               "roofline": {"bound": "hbm", "kernel": kernel + " (mean of the H-update and W-update launches, rank 0)",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": ncu_traffic(args.config, world, kernel),
                            "traffic_source": "profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from the "
                                              "committed ncu --set full capture of this command (null when this configuration was not captured)",
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
                e2e = e2e_leg_sharded(be, comm, cfg, args.steps, A_sh, rank, world)
        except MemoryError as ex:
            e2e = {"value": None, "unit": "iterations/s", "error": f"host memory: {ex}"}
    if rank == 0:
        out["e2e"] = e2e if e2e is not None else {"value": None, "unit": "iterations/s", "h2d_bytes_per_step": None,
                                                  "d2h_bytes_per_step": None, "note": "skipped (--no-e2e)"}
        if not args.no_cpu and world == 1:
            # bounded: ~10-30 s of CPU work (two warm iterations at 12,800 and 38,400 cells)
            cells = args.cpu_sample_cells if args.cpu_sample_cells > 0 else min(n, 38400)
            out["cpu_baseline"] = cpu_reference_leg(cfg, 2, cells)
    comm.close()
    be.close()
    be = None
    if rank == 0:
        if world == 1 and not args.no_extras and args.config == "c3":
            try:
                out.update(extras_legs(local_rank))
            except Exception as ex:  # the headline line must not be lost to a secondary leg
                out["extras_error"] = repr(ex)
        _emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


def extras_legs(device):
    """BASELINE configs[0], [1] and [3] through the public API on one GPU (seconds, wall clock, host matrices in, models
    out), with the reference CPU code timed live on the same fits where that takes seconds (C1, four fits of the C2 sweep)."""
    import scipy.sparse as sp

    from oracle.pyoracle import Oracle, have_reference
    from singlet_b200 import api, synth
    from singlet_b200.datasets import get_pbmc3k_data, log_normalize
    from singlet_b200.rrng import RRng
    from singlet_b200.sharded import CudaBackend

    res = {}
    orc = Oracle("reference" if have_reference() else "port")
    cores = max(orc.max_threads(), os.cpu_count() or 1)
    A = log_normalize(get_pbmc3k_data())
    At = A.T.tocsc()
    At.sort_indices()
    h = api.Handle(device)
    # ---- configs[0]: set.seed(123); run_nmf(A, rank = 10) on pbmc3k ----
    api.set_seed(123)
    api.run_nmf(A, 10, maxit=2, verbose=False, handle=h)  # warm-up (module load, allocator, padded rank 16)
    h.set_cache(False)
    api.set_seed(123)
    t0 = time.perf_counter()
    model = api.run_nmf(A, 10, verbose=False, handle=h)
    c1_gpu = time.perf_counter() - t0
    h.set_cache(True)
    w10 = RRng(123).matrix_runif(10, A.shape[0])
    t0 = time.perf_counter()
    orc.nmf(A, At, w10, tol=1e-4, maxit=100, L1=(0.01, 0.01), threads=cores)
    c1_cpu = time.perf_counter() - t0
    res["c1_run_nmf"] = {"workload": "pbmc3k 13,714 x 2,700 log-normalised, set.seed(123); run_nmf(A, rank = 10), tol 1e-4",
                         "seconds": c1_gpu, "iterations": int(model["iter"]), "cpu_reference_seconds": c1_cpu, "cpu_cores": cores,
                         "cpu_kind": orc.kind}
    # ---- configs[1]: set.seed(123); cross_validate_nmf(A, ranks = 2:30, n_replicates = 3) ----
    ranks = list(range(2, 31))
    api.set_seed(123)
    t0 = time.perf_counter()
    df = api.cross_validate_nmf(A, ranks, n_replicates=3, verbose=0, handle=h)
    cv_first = time.perf_counter() - t0
    api.set_seed(123)
    t0 = time.perf_counter()
    api.cross_validate_nmf(A, ranks, n_replicates=3, verbose=0, handle=h)
    cv_again = time.perf_counter() - t0
    # reference CPU on four fits of replicate 1 (the whole sweep takes minutes): seconds per fit interpolated over k
    r = RRng(123)
    w_init = [r.matrix_runif(max(ranks), A.shape[0]) for _ in range(3)]
    seed1 = abs(r.dot_random_seed(3 + 1))
    sampled = {}
    for kk in (2, 11, 20, 30):
        t0 = time.perf_counter()
        orc.ard_nmf(A, At, w_init[0][:kk, :], seed1, 20, tol=1e-4, maxit=100, L1=0.01, L2=0.0, threads=cores, overfit_threshold=1e-4,
                    trace_test_mse=5)
        sampled[kk] = time.perf_counter() - t0
    ks = sorted(sampled)
    est = 3.0 * float(sum(np.interp(kq, ks, [sampled[q] for q in ks]) for kq in ranks))
    golden = None
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "ref_pbmc3k_cv.npz"))
        golden = {"seconds": float(z["total_seconds"]), "threads": int(z["threads"]),
                  "note": "the whole sweep by the reference-compiled code in the build container (scripts/make_cv_goldens.py)"}
    except Exception:
        pass
    last = df.loc[df.groupby(["rep", "k"])["iter"].idxmax()]
    res["cv_sweep"] = {"workload": "pbmc3k, set.seed(123); cross_validate_nmf(A, ranks = 2:30, n_replicates = 3, test_density = 0.05): 87 fits",
                       "seconds": cv_first, "seconds_repeated_call": cv_again, "fits": len(ranks) * 3,
                       "best_rank": int(api.GetBestRank(df)),
                       "test_error_k10_rep1": float(last[(last["k"] == 10) & (last["rep"] == 1)]["test_error"].iloc[0]),
                       "cpu_reference": {"seconds_estimated": est, "cores": cores, "kind": orc.kind,
                                         "sampled_fit_seconds_rep1": {str(q): sampled[q] for q in ks},
                                         "how": "four fits of replicate 1 timed live; seconds per fit interpolated linearly over k, x 3 replicates",
                                         "build_container_full_sweep": golden}}
    h.close()
    # ---- configs[3]: ard_nmf on synthetic 20k x 250k, 8 % (structure-free: the search stops at its first bracket) and on a
    # matrix of the same gene count with a planted rank (the search walks) ----
    m4, n4, d4 = 20000, 250000, 0.08
    be = CudaBackend(device)
    hm = be.synth(m4, n4, d4, synth.DATA_SEED, 0, 0, n4, synth.values_table(m4, d4))
    p4 = be.matrix_to_host(hm)  # the host matrix a user hands to ard_nmf
    be.close()
    A4 = sp.csc_matrix((p4[2], p4[1], p4[0]), shape=(m4, n4))
    A4.has_sorted_indices = True
    h = api.Handle(device)
    api.set_seed(123)
    t0 = time.perf_counter()
    mod = api.ard_nmf(A4, L1=0.01, verbose=0, handle=h)
    c4 = time.perf_counter() - t0
    cv = mod["cv_data"]
    res["c4_ard_nmf"] = {"workload": "synthetic 20k x 250k, 8 % density, ard_nmf(A, L1 = 0.01) defaults", "seconds": c4, "nnz": int(A4.nnz),
                         "ranks_tried": [int(q) for q in sorted(cv["k"].unique())], "best_rank": int(mod["w"].shape[1]),
                         "final_iterations": int(mod["iter"])}
    del A4, mod
    Ap = planted_counts(m4, 25000, 12, d4, seed=7)
    api.set_seed(123)
    t0 = time.perf_counter()
    mod = api.ard_nmf(Ap, L1=0.01, k_max=64, verbose=0, handle=h)
    cp = time.perf_counter() - t0
    cv = mod["cv_data"]
    res["c4_ard_nmf"]["planted_rank_12"] = {"workload": "20k x 25k Poisson counts of a rank-12 non-negative model, log-normalised, ~8 % density",
                                            "seconds": cp, "nnz": int(Ap.nnz), "ranks_tried": [int(q) for q in cv["k"].unique()],
                                            "best_rank": int(mod["w"].shape[1]), "final_iterations": int(mod["iter"])}
    h.close()
    # ---- the masked (cross-validation) iteration at scale: c_ard_nmf's loop on a 30k x 100k slice of configs[2], k = 32, with the
    # per-column Gram corrections on the tensor cores (default) and inside the solver as FP32 FMAs (SGL_GRAMCORR=ffma, round 1) ----
    res["masked_als"] = masked_als_leg(device)
    return res


def masked_als_leg(device, m=30000, n=100000, dens=0.05, k=32, steps=5, warmup=3):
    from singlet_b200 import synth
    from singlet_b200.multi import RankComm, RankFit
    from singlet_b200.sharded import CudaBackend

    out = {"workload": f"masked ALS iteration (predict_mask both ways + scale + cor, src/singlet.cpp:1091-1152) on synthetic {m} x {n}, "
                       f"{dens:.0%}, k = {k}, 1/20 held out, through sgl_fit_iterate", "held_out_entries_per_half_iteration": m * n // 20}
    for name, env in (("tensor_core_correction", None), ("fp32_ffma_correction", "ffma")):
        if env is None:
            os.environ.pop("SGL_GRAMCORR", None)
        else:
            os.environ["SGL_GRAMCORR"] = env
        be = CudaBackend(device)
        try:
            table = synth.values_table(m, dens)
            A_sh = be.synth(m, n, dens, synth.DATA_SEED, 0, 0, n, table)
            At_sh = be.synth(m, n, dens, synth.DATA_SEED, 1, 0, m, table)
            comm = RankComm(be._h, device, 1, 0, None)
            fit = RankFit(comm, A_sh, At_sh, n, k, synth.w_init(k, m), masked=True, seed=4321, inv_density=20)
            for _ in range(warmup):
                fit.iterate(L1, L1, L2, L2)
            be.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                fit.iterate(L1, L1, L2, L2)  # synchronises (tol is read back every iteration)
            ms = (time.perf_counter() - t0) * 1000.0 / steps
            t0 = time.perf_counter()
            mse = fit.test_mse()
            ms_mse = (time.perf_counter() - t0) * 1000.0
            out[name] = {"ms_per_iteration": ms, "mse_test_ms": ms_mse, "test_mse": mse}
            fit.close()
            comm.close()
        finally:
            be.close()
            os.environ.pop("SGL_GRAMCORR", None)
    return out


def planted_counts(m, n, rank, density, seed):
    """Sparse log-normalised counts with a planted non-negative rank (column blocks keep the dense intermediate small)."""
    import scipy.sparse as sp

    rs = np.random.RandomState(seed)
    W0 = rs.gamma(0.3, 1.0, size=(m, rank)) * (rs.rand(m, rank) < 0.3)
    blocks = []
    for c0 in range(0, n, 2500):
        nb = min(2500, n - c0)
        H0 = rs.gamma(0.5, 1.0, size=(rank, nb)) * (rs.rand(rank, nb) < 0.4)
        lam = W0 @ H0
        lam *= (-np.log1p(-density)) / max(lam.mean(), 1e-300)
        blocks.append(sp.csc_matrix(rs.poisson(lam).astype(np.float64)))
    X = sp.hstack(blocks, format="csc")
    colsum = np.asarray(X.sum(axis=0)).ravel()
    colsum[colsum == 0] = 1.0
    X.data = np.log1p(X.data / np.repeat(colsum, np.diff(X.indptr)) * 1e4)
    X.sort_indices()
    return X


def e2e_leg_sharded(be, comm, cfg, steps, A_dev, rank, world):
    """N > 1: every rank starts from its HOST dgCMatrix cell shard and makes ONE call of the per-rank C ABI entry point
    sgl_nmf_rank: upload, device transpose, `steps` iterations with NCCL inside the library, download of w, d and of the
    rank's own block of h. Max over ranks."""
    import ctypes as C

    import scipy.sparse as sp
    import torch
    import torch.distributed as dist

    from singlet_b200 import _lib, synth
    from singlet_b200.sharded import shard_bounds

    m, n, k = cfg["m"], cfg["n"], cfg["k"]
    c0, c1, _ = shard_bounds(n, world, rank)
    pA = be.matrix_to_host(A_dev)
    A = sp.csc_matrix((pA[2], pA[1], pA[0]), shape=(m, c1 - c0))
    A.has_sorted_indices = True
    w = np.array(synth.w_init(k, m), order="F")
    d, h_loc = np.zeros(k), np.zeros((k, c1 - c0), order="F")
    iters = C.c_int32(0)
    lib = be.lib
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    arr, na, keep = _lib.chunks_to_c([A])
    hA = C.c_void_p()
    _lib.check(lib.sgl_matrix_upload(be._h, arr, na, C.byref(hA)))
    sec_upload = time.perf_counter() - t0
    _lib.check(lib.sgl_nmf_rank(comm._c, hA, None, n, 0.0, steps, L1, L1, L2, L2, k, w.ctypes.data, d.ctypes.data, h_loc.ctypes.data,
                                C.addressof(iters), None, None))
    sec_local = time.perf_counter() - t0
    lib.sgl_matrix_free(be._h, hA)
    dt = torch.tensor([sec_local], dtype=torch.float64, device=be.device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + w.nbytes)], dtype=torch.float64, device=be.device)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
    d2h = float(world) * (w.nbytes + d.nbytes + steps * 40) + 8.0 * k * n
    sec = float(dt[0])
    assert iters.value == steps
    return {"value": steps / sec, "unit": "iterations/s", "h2d_bytes_per_step": float(h2d[0]) / steps, "d2h_bytes_per_step": d2h / steps,
            "seconds_total": sec, "iterations": steps, "rank0_seconds": {"upload": sec_upload, "fit_and_download": sec_local - sec_upload},
            "call": "sgl_matrix_upload + sgl_nmf_rank per rank (include/singlet_cuda.h)",
            "note": "every rank uploads its host dgCMatrix cell shard (FP64), transposes it on the device, runs K iterations and downloads "
                    "w, d and its own block of h; max over ranks"}


def e2e_leg(be, cfg, steps, A_dev, At_dev):
    """Public API with a HOST dgCMatrix: `run_nmf(A, rank = k, tol = 0, maxit = steps)` -- upload, device transpose, `steps`
    iterations, download, sort -- plus the raw C ABI call sgl_nmf(A, At) with both host matrices for comparison."""
    import scipy.sparse as sp

    from singlet_b200 import api, synth

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
    A = sp.csc_matrix((pA[2], pA[1], pA[0]), shape=(m, n))
    A.has_sorted_indices = True  # generated in order; skips scipy's O(nnz) check
    h = api.Handle(be.device.index)
    h.set_cache(False)
    # warm-up on a tiny problem (module load, allocator) -- not the timed call
    As = synth.synth_scipy(2000, 1500, 0.05)
    api.set_seed(123)
    api.run_nmf(As, k, tol=0.0, maxit=2, verbose=False, L1=L1, L2=L2, handle=h)
    api.set_seed(123)
    t0 = time.perf_counter()
    res = api.run_nmf(A, k, tol=0.0, maxit=steps, verbose=False, L1=L1, L2=L2, handle=h)
    dt = time.perf_counter() - t0
    assert res["iter"] == steps and res["w"].shape == (m, k) and res["h"].shape == (k, n)
    h2d = A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + 8 * k * m
    d2h = res["w"].nbytes + res["h"].nbytes + res["d"].nbytes + steps * 40
    out = {"value": steps / dt, "unit": "iterations/s", "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps,
           "seconds_total": dt, "iterations": steps, "call": "singlet_b200.api.run_nmf(A, rank, tol = 0, maxit = K) -- R/run_nmf.R:18",
           "note": "one run_nmf call on a host FP64 dgCMatrix (pageable memory): w_init drawn from the R-compatible RNG, A packed to "
                   "8-byte records by host threads and uploaded through a pinned ring, t(A) built on the device, K iterations, w/d/h "
                   "downloaded, factors sorted by d; h2d_bytes_per_step = the host buffers handed to the engine (12 B per non-zero + "
                   "w_init), divided by K"}
    # the raw C ABI call with BOTH host matrices, as src/RcppExports.cpp:97-116 receives them from R (t(A) made by the caller)
    try:
        pAt = be.matrix_to_host(At_dev)
        At = sp.csc_matrix((pAt[2], pAt[1], pAt[0]), shape=(n, m))
        At.has_sorted_indices = True
        w0 = synth.w_init(k, m)
        t0 = time.perf_counter()
        res2 = api.c_nmf(A, At, 0.0, steps, False, L1, L1, L2, L2, 0, w0, h)
        dt2 = time.perf_counter() - t0
        out["c_abi_with_host_At"] = {"value": steps / dt2, "seconds_total": dt2, "iterations": int(res2["iter"]),
                                     "h2d_bytes_per_step": (h2d + At.data.nbytes + At.indices.nbytes + At.indptr.nbytes) / steps}
    except MemoryError as ex:
        out["c_abi_with_host_At"] = {"value": None, "error": repr(ex)}
    h.close()
    return out


if __name__ == "__main__":
    sys.exit(main())

"""Secondary BASELINE metric: CV rank-sweep wall time on pbmc3k (configs[0] and configs[1]).

    python scripts/bench_cv.py [--ranks 2:30] [--reps 3] [--cpu-fits N]

* C1: set.seed(123); run_nmf(A, rank = 10) on log-normalised pbmc3k (13,714 x 2,700).
* C2: set.seed(123); cross_validate_nmf(A, ranks = 2:30, n_replicates = 3) -> 87 masked fits
  (reference R/cross_validate_nmf.R:18-105), end to end through the C ABI including the upload.
The reference arm (oracle/_ref, all host threads) runs the same fits; because the full sweep takes
minutes on the CPU only the first `--cpu-fits` fits are timed and compared fit by fit.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", default="2:30")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cpu-fits", type=int, default=6)
args = ap.parse_args()
lo, hi = (int(x) for x in args.ranks.split(":"))
ranks = list(range(lo, hi + 1))

from oracle.pyoracle import Oracle, have_reference  # noqa: E402
from singlet_b200 import api  # noqa: E402
from singlet_b200.datasets import get_pbmc3k_data, log_normalize  # noqa: E402
from singlet_b200.rrng import RRng  # noqa: E402

A = log_normalize(get_pbmc3k_data())
At = A.T.tocsc()
At.sort_indices()
out = {"dataset": "pbmc3k 13714 x 2700 log-normalised", "nnz": int(A.nnz)}

# ---- C1 ----
api.set_seed(123)
api.run_nmf(A, 10, maxit=2, verbose=False)  # warm-up with the same padded rank (module load, allocator)
api.default_handle().set_cache(False)
api.set_seed(123)
t0 = time.perf_counter()
model = api.run_nmf(A, 10, verbose=False)
out["c1_run_nmf_k10_s"] = time.perf_counter() - t0
out["c1_iters"] = int(model["iter"])
api.default_handle().set_cache(True)

orc = Oracle("reference" if have_reference() else "port")
w0 = RRng(123).matrix_runif(10, A.shape[0])
t0 = time.perf_counter()
ref = orc.nmf(A, At, w0, tol=1e-4, maxit=100, L1=(0.01, 0.01))
out["c1_cpu_s"] = time.perf_counter() - t0
out["c1_cpu_iters"] = int(ref["iter"]) if ref["iter"] >= 0 else None
out["cpu_threads"] = orc.max_threads()
out["cpu_kind"] = orc.kind

# ---- C2 ----
api.set_seed(123)
t0 = time.perf_counter()
df = api.cross_validate_nmf(A, ranks, n_replicates=args.reps, verbose=0)
out["c2_cv_sweep_s"] = time.perf_counter() - t0  # first call: includes the upload, mask builds and worker start-up
out["c2_fits"] = len(ranks) * args.reps
api.set_seed(123)
t0 = time.perf_counter()
df2 = api.cross_validate_nmf(A, ranks, n_replicates=args.reps, verbose=0)
out["c2_cv_sweep_warm_s"] = time.perf_counter() - t0  # same call again (A, masks and workers cached in the handle)
api.set_seed(123)
t0 = time.perf_counter()
df3 = api.cross_validate_nmf(A, ranks, n_replicates=args.reps, verbose=0, batch=False)
out["c2_cv_sweep_fit_by_fit_s"] = time.perf_counter() - t0  # one c_ard_nmf call per fit, like the R loop
out["c2_batch_equals_fit_by_fit"] = bool(df.equals(df2) and df.equals(df3))
last = df.loc[df.groupby(["rep", "k"])["iter"].idxmax()]
out["c2_best_rank"] = int(api.GetBestRank(df))
out["c2_test_error_k10_rep1"] = float(last[(last["k"] == 10) & (last["rep"] == 1)]["test_error"].iloc[0]) if 10 in ranks else None

# reference arm on the first fits of the sweep (same w_init and mask seeds as R would draw)
r = RRng(123)
w_init = [r.matrix_runif(max(ranks), A.shape[0]) for _ in range(args.reps)]
cpu, gpu_same, worst = 0.0, 0.0, 0.0
for k in ranks[: args.cpu_fits]:
    seed = abs(r.dot_random_seed(3 + 1))
    t0 = time.perf_counter()
    cm = orc.ard_nmf(A, At, w_init[0][:k, :], seed, 20, tol=1e-4, maxit=100, L1=0.01, L2=0.0, overfit_threshold=1e-4,
                     trace_test_mse=5)
    cpu += time.perf_counter() - t0
    t0 = time.perf_counter()
    gm = api.c_ard_nmf(A, At, 1e-4, 100, False, 0.01, 0.0, 0, w_init[0][:k, :], seed, 20, 1e-4, 5)
    gpu_same += time.perf_counter() - t0
    n = min(len(cm["test_mse"]), len(gm["test_mse"]))
    worst = max(worst, float(np.max(np.abs(cm["test_mse"][:n] - gm["test_mse"][:n]) / cm["test_mse"][:n])))
out["c2_cpu_first_fits_s"] = cpu
out["c2_gpu_same_fits_s"] = gpu_same
out["c2_first_fits"] = min(args.cpu_fits, len(ranks))
out["c2_max_rel_test_mse_diff_first_fits"] = worst
print(json.dumps(out))

"""Generate tests/golden/ref_pbmc3k_cv.npz: the WHOLE BASELINE configs[1] sweep -- `set.seed(123);
cross_validate_nmf(A, ranks = 2:30, n_replicates = 3, test_density = 0.05)` on log-normalised pbmc3k, 87 c_ard_nmf
fits -- computed by the reference's own functions compiled from /root/reference (oracle/_ref). Run in the build
container (takes tens of minutes on 8 cores); the GPU box only reads the committed file.

Stored per fit (grid order of R/cross_validate_nmf.R:69, k fastest): k, rep, the test_mse / iter / tol trace vectors
(padded with NaN / -1), the number of trace entries, and d. The wall time of every fit on this container's cores is
stored too (bench.py quotes it as the CPU side of the secondary metric).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle  # noqa: E402
from singlet_b200.datasets import get_pbmc3k_data, log_normalize  # noqa: E402
from singlet_b200.rrng import RRng  # noqa: E402

ref = Oracle("reference")
A = log_normalize(get_pbmc3k_data())
At = A.T.tocsc()
At.sort_indices()
ranks, n_rep = list(range(2, 31)), 3
r = RRng(123)
w_init = [r.matrix_runif(max(ranks), A.shape[0]) for _ in range(n_rep)]
seeds = [abs(r.dot_random_seed(3 + rep)) for rep in range(1, n_rep + 1)]
grid = [(k, rep) for rep in range(1, n_rep + 1) for k in ranks]  # expand.grid: k varies fastest
CAP = 24
n = len(grid)
out = dict(k=np.zeros(n, np.int32), rep=np.zeros(n, np.int32), n_trace=np.zeros(n, np.int32), test_mse=np.full((n, CAP), np.nan),
           iter=np.full((n, CAP), -1, np.int32), tol=np.full((n, CAP), np.nan), seconds=np.zeros(n), d=np.full((n, 30), np.nan))
t_all = time.perf_counter()
for q, (k, rep) in enumerate(grid):
    t0 = time.perf_counter()
    fit = ref.ard_nmf(A, At, w_init[rep - 1][:k, :], seeds[rep - 1], 20, tol=1e-4, maxit=100, L1=0.01, L2=0.0,
                      overfit_threshold=1e-4, trace_test_mse=5)
    dt = time.perf_counter() - t0
    nt = len(fit["test_mse"])
    assert nt <= CAP
    out["k"][q], out["rep"][q], out["n_trace"][q], out["seconds"][q] = k, rep, nt, dt
    out["test_mse"][q, :nt], out["iter"][q, :nt], out["tol"][q, :nt] = fit["test_mse"], fit["iter"], fit["tol"]
    out["d"][q, :k] = fit["d"]
    print(f"{q + 1}/{n} k={k} rep={rep} trace={nt} last_iter={fit['iter'][-1]} test_mse={fit['test_mse'][-1]:.6f} {dt:.1f}s", flush=True)
dst = os.path.join(ROOT, "tests", "golden", "ref_pbmc3k_cv.npz")
np.savez_compressed(dst, seeds=np.array(seeds, dtype=np.uint64), threads=ref.max_threads(), total_seconds=time.perf_counter() - t_all, **out)
print(dst, os.path.getsize(dst))

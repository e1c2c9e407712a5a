"""Developer probe: GPU CV sweep vs the reference-compiled goldens, fit by fit."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlet_b200 import api
from singlet_b200.datasets import get_pbmc3k_data, log_normalize
z = np.load("tests/golden/ref_pbmc3k_cv.npz")
A = log_normalize(get_pbmc3k_data())
for prec in ("mixed16", "fp32"):
    api.default_handle().set_precision(prec)
    api.set_seed(123)
    df = api.cross_validate_nmf(A, list(range(2, 31)), n_replicates=3, verbose=0)
    bad, worst = 0, 0.0
    for q in range(87):
        k, rep, nt = int(z["k"][q]), int(z["rep"][q]), int(z["n_trace"][q])
        rows = df[(df["k"] == k) & (df["rep"] == rep)]
        it_dev, it_ref = list(rows["iter"]), [int(v) for v in z["iter"][q, :nt]]
        n = min(len(it_dev), nt)
        rel = np.abs(rows["test_error"].to_numpy()[:n] - z["test_mse"][q, :n]) / z["test_mse"][q, :n]
        worst = max(worst, rel.max())
        if it_dev != it_ref:
            bad += 1
            print(prec, "k", k, "rep", rep, "dev", it_dev[-3:], "ref", it_ref[-3:], "tol dev", rows["tol"].to_numpy()[-2:], "tol ref", z["tol"][q, nt-2:nt],
                  "mse dev", rows["test_error"].to_numpy()[-2:], "ref", z["test_mse"][q, nt-2:nt], "prefix rel", rel.max())
    print(prec, "mismatched", bad, "worst rel on common prefix", worst)

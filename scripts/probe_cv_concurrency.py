"""Developer probe: cross_validate_nmf wall time on pbmc3k for a list of batch concurrencies (0 = fit by fit)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from singlet_b200 import api
from singlet_b200.datasets import get_pbmc3k_data, log_normalize
A = log_normalize(get_pbmc3k_data())
api.set_seed(123)
api.run_nmf(A, 10, maxit=2, verbose=False)
ref = None
for conc in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2,3,4,6,8,4").split(",")]:
    api.set_seed(123)
    t0 = time.perf_counter()
    df = api.cross_validate_nmf(A, list(range(2, 31)), n_replicates=3, verbose=0, batch=conc > 0, concurrency=conc)
    tot = time.perf_counter() - t0
    if ref is None: ref = df
    print(json.dumps({"concurrency": conc, "sweep_s": round(tot, 3), "same_as_first": bool(df.equals(ref)), "best": int(api.GetBestRank(df))}), flush=True)

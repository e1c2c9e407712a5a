import sys, time, json
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from singlet_b200 import api
from singlet_b200.datasets import get_pbmc3k_data, log_normalize
A = log_normalize(get_pbmc3k_data())
api.set_seed(123); api.cross_validate_nmf(A, ranks=list(range(2, 31)), n_replicates=3, verbose=0)  # warm (upload, masks, kernels)
for conc in (1, 2, 4, 6, 8, 12, 16):
    api.set_seed(123)
    t0 = time.perf_counter()
    df = api.cross_validate_nmf(A, ranks=list(range(2, 31)), n_replicates=3, verbose=0, concurrency=conc)
    print(conc, round(time.perf_counter() - t0, 3), flush=True)

"""Developer probe: where the FIRST cross_validate_nmf call of a process spends its time against a repeated call."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from singlet_b200 import api, synth  # noqa: E402
from singlet_b200.datasets import get_pbmc3k_data, log_normalize  # noqa: E402

t0 = time.perf_counter()
A = log_normalize(get_pbmc3k_data())
print("load + normalise pbmc3k: %.3f s" % (time.perf_counter() - t0), flush=True)
t0 = time.perf_counter()
h = api.default_handle()
As = synth.synth_scipy(500, 400, 0.05)
api.c_nmf(As, None, 0.0, 2, False, 0.01, 0.01, 0, 0, 0, synth.w_init(4, 500))
print("context + library + tiny fit: %.3f s" % (time.perf_counter() - t0), flush=True)
for conc in (4, 4, 1, 4):
    api.set_seed(123)
    t0 = time.perf_counter()
    df = api.cross_validate_nmf(A, ranks=list(range(2, 31)), n_replicates=3, verbose=0, concurrency=conc)
    print("sweep (concurrency %d): %.3f s" % (conc, time.perf_counter() - t0), flush=True)
api.set_seed(123)
t0 = time.perf_counter()
df = api.cross_validate_nmf(A, ranks=[30], n_replicates=1, verbose=0)
print("one fit k = 30: %.3f s" % (time.perf_counter() - t0), flush=True)

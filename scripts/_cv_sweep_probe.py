import os, sys, time, json
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import pandas  # noqa
from singlet_b200 import api, _lib
from singlet_b200.datasets import get_pbmc3k_data, log_normalize
A = log_normalize(get_pbmc3k_data())
api.set_seed(123)
api.run_nmf(A, 10, maxit=2, verbose=False)
h = api.default_handle()
def pread():
    ms, cnt, byt = np.zeros(4), np.zeros(4, np.int64), np.zeros(4, np.int64)
    _lib.check(h.lib.sgl_profile_read(h.ptr, ms.ctypes.data, cnt.ctypes.data, byt.ctypes.data))
    return float(ms.sum())
fit_s = []
orig = api.c_ard_nmf
def timed(*a, **kw):
    t0 = time.perf_counter(); r = orig(*a, **kw); dt = time.perf_counter() - t0
    fit_s.append((round(dt * 1e3, 2), a[8].shape[0], int(r["iter"][-1]), round(pread(), 2))); return r
api.c_ard_nmf = timed
for rnd in range(2):
    fit_s.clear()
    _lib.check(h.lib.sgl_profile(h.ptr, 1)); pread()
    api.set_seed(123)
    t0 = time.perf_counter()
    df = api.cross_validate_nmf(A, list(range(2, 31)), n_replicates=3, verbose=0)
    tot = time.perf_counter() - t0
    fs = np.array([f[0] for f in fit_s]); gs = np.array([f[3] for f in fit_s])
    print(json.dumps({"round": rnd, "sweep_s": round(tot, 3), "sum_fits_ms": round(float(fs.sum()), 1), "sum_gpu_span_ms": round(float(gs.sum()), 1),
                      "total_traced_iters": int(sum(f[2] for f in fit_s))}))
    if rnd == 1:
        for f in fit_s: print(f)

"""Developer probe: wall time, iterations and per-kind GPU spans of single masked fits on pbmc3k (python scripts/probe_cv_fit.py 8,16,30 p)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from singlet_b200 import api, _lib
from singlet_b200.datasets import get_pbmc3k_data, log_normalize
from singlet_b200.rrng import RRng
A = log_normalize(get_pbmc3k_data())
At = A.T.tocsc(); At.sort_indices()
r = RRng(123)
w_init = r.matrix_runif(30, A.shape[0])
ks = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,16,30").split(",")]
prof = len(sys.argv) > 2
api.c_ard_nmf(A, At, 1e-4, 3, False, 0.01, 0.0, 0, w_init[:4, :], 77, 20, 1e-4, 5)
h = api.default_handle()
def pread():
    ms, cnt, byt = np.zeros(4), np.zeros(4, np.int64), np.zeros(4, np.int64)
    _lib.check(h.lib.sgl_profile_read(h.ptr, ms.ctypes.data, cnt.ctypes.data, byt.ctypes.data))
    return [round(float(x), 3) for x in ms], [int(x) for x in cnt]
for k in ks:
    api.c_ard_nmf(A, At, 1e-4, 2, False, 0.01, 0.0, 0, w_init[:k, :], 77, 20, 1e-4, 5)
    if prof:
        _lib.check(h.lib.sgl_profile(h.ptr, 1)); pread()
    t0 = time.perf_counter()
    m = api.c_ard_nmf(A, At, 1e-4, 100, False, 0.01, 0.0, 0, w_init[:k, :], 77, 20, 1e-4, 5)
    dt = time.perf_counter() - t0
    it = int(m["iter"][-1]) if len(m["iter"]) else -1
    out = {"k": k, "s": round(dt, 5), "last_traced_iter": it, "ms_per_iter": round(1e3 * dt / max(it, 1), 3)}
    if prof:
        out["gpu_ms_spmm_nnls_gram_other"], out["spans"] = pread()
        _lib.check(h.lib.sgl_profile(h.ptr, 0))
    print(json.dumps(out), flush=True)

"""Developer probe: phases of one host-facing sgl_nmf call at BASELINE configs[2] (SGL_TIMING=1 prints them)."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SGL_TIMING"] = "1"
from singlet_b200 import api, synth  # noqa: E402
from singlet_b200.sharded import CudaBackend  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m, dens, k, steps = 30000, 0.05, 32, 10
be = CudaBackend(0)
A_dev = be.synth(m, n, dens, synth.DATA_SEED, 0, 0, n, synth.values_table(m, dens))
p = be.matrix_to_host(A_dev)
be.close()
A = sp.csc_matrix((p[2], p[1], p[0]), shape=(m, n))
A.has_sorted_indices = True
w0 = synth.w_init(k, m)
h = api.Handle(0)
h.set_cache(False)
As = synth.synth_scipy(2000, 1500, 0.05)
api.c_nmf(As, None, 0.0, 2, False, 0.01, 0.01, 0, 0, 0, synth.w_init(k, 2000), h)
print("---- timed call (At = NULL) ----", file=sys.stderr)
t0 = time.perf_counter()
res = api.c_nmf(A, None, 0.0, steps, False, 0.01, 0.01, 0.0, 0.0, 0, w0, h)
print("total %.3f s" % (time.perf_counter() - t0), file=sys.stderr)
t0 = time.perf_counter()
res = api.c_nmf(A, None, 0.0, steps, False, 0.01, 0.01, 0.0, 0.0, 0, w0, h)
print("total (second call) %.3f s" % (time.perf_counter() - t0), file=sys.stderr)

"""BASELINE configs[3]: ard_nmf (automatic rank determination) on synthetic 20k x 250k, 8 % density, L1 = 0.01,
end to end through the host-facing API (host dgCMatrix in, model out). Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from singlet_b200 import api, synth  # noqa: E402
from singlet_b200.sharded import CudaBackend  # noqa: E402

m, n, dens = 20000, 250000, 0.08
if len(sys.argv) > 1:
    m, n = int(sys.argv[1]), int(sys.argv[2])
be = CudaBackend(0)
h = be.synth(m, n, dens, synth.DATA_SEED, 0, 0, n, synth.values_table(m, dens))
p, i, x, _, _ = be.matrix_to_host(h)  # the host matrix a user would hand to ard_nmf
be.close()
A = sp.csc_matrix((x, i, p), shape=(m, n))
A.has_sorted_indices = True
api.set_seed(123)
t0 = time.perf_counter()
model = api.ard_nmf(A, L1=0.01, verbose=0)
dt = time.perf_counter() - t0
cv = model["cv_data"]
fits = cv.groupby(["rep", "k"])["iter"].max()
print(json.dumps({"config": f"ard_nmf synthetic {m} x {n}, {dens:.0%}", "nnz": int(A.nnz), "seconds": dt,
                  "best_rank": int(model["w"].shape[1]), "ranks_tried": [int(k) for k in sorted(cv["k"].unique())],
                  "cv_iterations": int(fits.sum() + len(fits)), "final_iter": int(model["iter"]), "final_tol": float(model["tol"])}))

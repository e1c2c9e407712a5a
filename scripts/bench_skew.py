"""Skewed real-data structure at scale: the bundled pbmc3k matrix (13,714 genes x 2,700 cells, gene non-zero counts from 3 to
2,700) tiled 100x along the cells (13,714 x 270,000, 228 M non-zeros) against the structure-free synthetic matrix of the same
shape and density, k = 32, plain ALS through sgl_fit_iterate. Exercises the load-balanced column groups of the W-update SpMM
(snake deal by non-zero count) at a size where the groups no longer fit one wave. Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from singlet_b200 import synth  # noqa: E402
from singlet_b200.datasets import get_pbmc3k_data, log_normalize  # noqa: E402
from singlet_b200.multi import RankComm, RankFit  # noqa: E402
from singlet_b200.sharded import CudaBackend  # noqa: E402

TILES, K, STEPS, WARMUP = int(sys.argv[1]) if len(sys.argv) > 1 else 100, 32, 6, 3


def run(be, A_h, At_h, m, n, label):
    comm = RankComm(be._h, 0, 1, 0, None)
    fit = RankFit(comm, A_h, At_h, n, K, synth.w_init(K, m))
    for _ in range(WARMUP):
        fit.iterate(0.01, 0.01, 0.0, 0.0)
    be.synchronize()
    be.profile(True)
    be.profile_read()
    t0 = time.perf_counter()
    for _ in range(STEPS):
        fit.iterate(0.01, 0.01, 0.0, 0.0)
    ms = (time.perf_counter() - t0) * 1000 / STEPS
    prof = be.profile_read()
    be.profile(False)
    fit.close()
    comm.close()
    return {"matrix": label, "ms_per_iteration": ms, "breakdown_ms": {k: v[0] / STEPS for k, v in prof.items()}}


P = log_normalize(get_pbmc3k_data())
A = sp.hstack([P] * TILES, format="csc")
A.sort_indices()
m, n = A.shape
dens = A.nnz / (m * n)
gene_nnz = np.diff(A.T.tocsr().indptr) if False else np.bincount(A.indices, minlength=m)
out = {"shape": [m, n], "nnz": int(A.nnz), "density": dens, "k": K,
       "gene_nnz_min_median_max": [int(gene_nnz.min()), int(np.median(gene_nnz)), int(gene_nnz.max())],
       "cell_nnz_min_median_max": [int(np.diff(A.indptr).min()), int(np.median(np.diff(A.indptr))), int(np.diff(A.indptr).max())]}
be = CudaBackend(0)
hA = be.upload(A)
hAt = be.transpose(hA)
out["tiled_pbmc3k"] = run(be, hA, hAt, m, n, "pbmc3k tiled %dx (real gene skew)" % TILES)
be.close()
be = CudaBackend(0)
table = synth.values_table(m, dens)
sA = be.synth(m, n, dens, synth.DATA_SEED, 0, 0, n, table)
sAt = be.synth(m, n, dens, synth.DATA_SEED, 1, 0, m, table)
out["synthetic_same_shape"] = run(be, sA, sAt, m, n, "structure-free synthetic, same shape and density")
be.close()
print(json.dumps(out))

"""Generate tests/golden/ref_small.npz from oracle/_ref (the reference's own functions compiled from
/root/reference). Run in the build container; the GPU box only reads the committed file."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Oracle  # noqa: E402
from singlet_b200 import synth  # noqa: E402
from test_oracle_parity import _small  # noqa: E402

ref = Oracle("reference")
seed, k, w_seed, maxit = 31, 6, 9, 10
A, At = _small(seed=seed)
w0 = synth.w_init(k, A.shape[0], seed=w_seed)
a = ref.nmf(A, At, w0, tol=1e-4, maxit=maxit, L1=(0.01, 0.01))
b = ref.ard_nmf(A, At, w0, 123, 20, tol=1e-4, maxit=maxit, trace_test_mse=2, overfit_threshold=10.0)
p = ref.project_model(A, w0)
mask = np.array([[ref.draw(123, c, g, 20) for g in range(A.shape[0])] for c in range(A.shape[1])], dtype=bool)
out = os.path.join(ROOT, "tests", "golden", "ref_small.npz")
np.savez_compressed(out, seed=seed, k=k, w_seed=w_seed, maxit=maxit, nmf_w=a["w"], nmf_h=a["h"], nmf_d=a["d"],
                    ard_test_mse=b["test_mse"], ard_iter=b["iter"], ard_h=b["h"], proj_h=p["h"], proj_d=p["d"],
                    mask_bits=np.packbits(mask))
print(out, os.path.getsize(out))

# ---- pbmc3k goldens (BASELINE configs[0] and one fit of configs[1]) from the reference-compiled code ----
from singlet_b200.datasets import get_pbmc3k_data, log_normalize  # noqa: E402
from singlet_b200.rrng import RRng  # noqa: E402

Ap = log_normalize(get_pbmc3k_data())
Atp = Ap.T.tocsc()
Atp.sort_indices()
w10 = RRng(123).matrix_runif(10, Ap.shape[0])  # set.seed(123); matrix(runif(m * 10), 10, m)
c1 = ref.nmf(Ap, Atp, w10, tol=1e-4, maxit=100, L1=(0.01, 0.01))
port = Oracle("port").nmf(Ap, Atp, w10, tol=1e-4, maxit=100, L1=(0.01, 0.01))  # the port reports iterations / tol trace
assert np.array_equal(port["w"], c1["w"]) and np.array_equal(port["h"], c1["h"])
r = RRng(123)
w_init = [r.matrix_runif(30, Ap.shape[0]) for _ in range(3)]  # cross_validate_nmf(ranks = 2:30, n_replicates = 3)
seeds = [abs(r.dot_random_seed(3 + rep)) for rep in (1, 2, 3)]
cv = ref.ard_nmf(Ap, Atp, w_init[0][:5, :], seeds[0], 20, tol=1e-4, maxit=100, L1=0.01, L2=0.0, overfit_threshold=1e-4,
                 trace_test_mse=5)
out2 = os.path.join(ROOT, "tests", "golden", "ref_pbmc3k.npz")
np.savez_compressed(out2, c1_w=c1["w"].astype(np.float32), c1_h=c1["h"].astype(np.float32), c1_d=c1["d"], c1_iter=port["iter"],
                    c1_tol=port["tol"], cv_seeds=np.array(seeds, dtype=np.uint64), cv_k=5, cv_test_mse=cv["test_mse"],
                    cv_iter=cv["iter"], cv_tol=cv["tol"], cv_d=cv["d"])
print(out2, os.path.getsize(out2), "C1 iterations", port["iter"], "CV trace", cv["test_mse"])

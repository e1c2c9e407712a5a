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

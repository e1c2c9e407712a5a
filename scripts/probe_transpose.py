"""Developer probe: device transpose of the headline matrix (30k x 1M, 5 %): time and equality with the generated At."""
import os, sys, time, json
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from singlet_b200 import synth
from singlet_b200.sharded import CudaBackend
m, n, dens = 30000, 1000000, 0.05
if len(sys.argv) > 2: m, n = int(sys.argv[1]), int(sys.argv[2])
be = CudaBackend(0)
tab = synth.values_table(m, dens)
A = be.synth(m, n, dens, synth.DATA_SEED, 0, 0, n, tab)
be.synchronize()
t0 = time.perf_counter(); T = be.transpose(A); be.synchronize(); t1 = time.perf_counter() - t0
t0 = time.perf_counter(); T2 = be.transpose(A); be.synchronize(); t2 = time.perf_counter() - t0
out = {"m": m, "n": n, "nnz": be.matrix_info(A)[2], "transpose_first_s": round(t1, 4), "transpose_s": round(t2, 4)}
if n <= 200000:
    At = be.synth(m, n, dens, synth.DATA_SEED, 1, 0, m, tab)
    a, b = be.matrix_to_host(T), be.matrix_to_host(At)
    out["equal_to_generated_At"] = bool(all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])))
print(json.dumps(out))

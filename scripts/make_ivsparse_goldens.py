"""Golden IVSparse images written by the REFERENCE's own codec (oracle/_ref/libivsparse_ref.so, compiled from
/root/reference/inst/include/IVSparse.h): tests/golden/ivsparse_golden.npz. Run in the build container (needs the reference)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_ivsparse import counts_matrix, ref_lib, ref_write  # noqa: E402

lib = ref_lib()
A = counts_matrix(70000, 24, 0.004, 11)  # 3-byte row deltas, empty / constant columns, row 0 as a first index
out = {"p": A.indptr.astype(np.int32), "i": A.indices.astype(np.int32), "x": A.data, "shape": np.array(A.shape)}
with tempfile.TemporaryDirectory() as d:
    for level in (2, 3):
        out["image_l%d" % level] = ref_write(lib, A, level, os.path.join(d, "g%d.bin" % level))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ivsparse_golden.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})

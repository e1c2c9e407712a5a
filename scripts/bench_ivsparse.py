"""IVSparse codec throughput on the host (row f4): csrc/ivsparse.cpp (sgl_ivsparse_encode / decode) against the reference's own
codec compiled from its headers (oracle/_ref/libivsparse_ref.so: compressCSC + write, file constructor + InnerIterator) on
the bundled pbmc3k counts (13,714 x 2,700, 2.28 M non-zeros). Prints one JSON line. Needs no GPU."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from singlet_b200 import ivsparse  # noqa: E402
from singlet_b200.datasets import get_pbmc3k_data  # noqa: E402

A = get_pbmc3k_data()
out = {"matrix": "pbmc3k counts 13714 x 2700", "nnz": int(A.nnz), "dgCMatrix_bytes": int(A.nnz * 12 + 4 * (A.shape[1] + 1))}


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        t.append(time.perf_counter() - t0)
    return min(t), r


for level, name in ((3, "ivcsc"), (2, "vcsc")):
    te, img = best(lambda: ivsparse.encode(A, level))
    td, D = best(lambda: ivsparse.decode(img))
    assert (D != A).nnz == 0
    out[name] = {"image_bytes": int(img.nbytes), "bytes_per_nnz": img.nbytes / A.nnz, "encode_s": te, "decode_s": td,
                 "encode_Mnnz_per_s": A.nnz / te / 1e6, "decode_Mnnz_per_s": A.nnz / td / 1e6}
    try:
        from test_ivsparse import ref_lib, ref_read, ref_write

        lib = ref_lib()
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "r.bin")
            tw, ref_img = best(lambda: ref_write(lib, A, level, path))
            assert np.array_equal(ref_img, img)
            tr, _ = best(lambda: ref_read(lib, level, path, A.nnz))
        out[name]["reference_compress_and_write_s"] = tw
        out[name]["reference_read_and_iterate_s"] = tr
    except BaseException as e:  # pytest.skip raises when the reference codec is not built
        out[name]["reference"] = "not built here: %s" % (e,)
print(json.dumps(out))

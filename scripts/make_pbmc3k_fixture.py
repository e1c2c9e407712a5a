"""Regenerate tests/golden/pbmc3k_counts.npz from the reference's bundled dataset.

Run in the build container only (reads /root/reference/data/pbmc3k.RData, which does not exist
on the GPU box). The fixture stores the raw integer counts of the 13,714 x 2,700 dgCMatrix
(reference R/get_pbmc3k_data.R:14-20) in a compact form: column pointers, per-column
delta-encoded row indices (uint16) and counts (uint16).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from singlet_b200.rdata import inverse_rle, read_rdata  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/pbmc3k.RData"
dst = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "pbmc3k_counts.npz")

o = read_rdata(src)["pbmc3k"]
p, i = np.asarray(o["p"], np.int32), np.asarray(o["i"], np.int32)
x = inverse_rle(o["x"]).astype(np.int64)
dim = np.asarray(o["Dim"], np.int32)
assert x.size == i.size == p[-1] and x.max() < 65536
delta = np.diff(i, prepend=0).astype(np.int64)
delta[p[:-1]] = i[p[:-1]]  # first entry of each column stores the absolute row
assert delta.min() >= 0 and delta.max() < 65536
np.savez_compressed(dst, p=p, di=delta.astype(np.uint16), x=x.astype(np.uint16), dim=dim)
print(dst, os.path.getsize(dst), "bytes; nnz", x.size, "dim", dim)

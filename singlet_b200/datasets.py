"""Dataset helpers mirroring the reference's R-side glue for its bundled demo data.

* :func:`get_pbmc3k_data` -- reference R/get_pbmc3k_data.R:14-20: rebuild the 13,714 x 2,700 count
  dgCMatrix. Reads the compact fixture generated from ``data/pbmc3k.RData`` by
  scripts/make_pbmc3k_fixture.py (or an ``.RData`` file directly when given).
* :func:`log_normalize` -- reference R/PreprocessData.R:34-39 (``Seurat::LogNormalize`` with
  ``scale.factor = 1e4``): ``log1p(x / colSum * 1e4)``.
"""
from __future__ import annotations

import os

import numpy as np

_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "pbmc3k_counts.npz")


def get_pbmc3k_data(path: str | None = None):
    """Return the raw counts as a scipy CSC matrix (genes x cells)."""
    import scipy.sparse as sp

    if path is not None and path.lower().endswith((".rdata", ".rda")):
        from .rdata import inverse_rle, read_rdata

        o = read_rdata(path)["pbmc3k"]
        p, i = np.asarray(o["p"], np.int32), np.asarray(o["i"], np.int32)
        x = inverse_rle(o["x"]).astype(np.float64)
        dim = np.asarray(o["Dim"])
    else:
        z = np.load(path or _FIXTURE)
        p, dim = z["p"].astype(np.int32), z["dim"]
        di = z["di"].astype(np.int64)
        # per-column delta decoding: cumulative sum restarted at every column start
        c = np.cumsum(di)
        starts = p[:-1][np.diff(p) > 0]
        base = np.zeros(di.size, dtype=np.int64)
        base[starts] = c[starts] - di[starts]
        base = np.maximum.accumulate(base)
        i = (c - base).astype(np.int32)
        x = z["x"].astype(np.float64)
    return sp.csc_matrix((x, i, p), shape=(int(dim[0]), int(dim[1])))


def log_normalize(A, scale_factor: float = 1e4):
    """``log1p(x / colSum * scale_factor)`` on the stored entries of a CSC matrix."""
    A = A.tocsc(copy=True).astype(np.float64)
    sums = np.asarray(A.sum(axis=0)).ravel()
    per_nz = np.repeat(sums, np.diff(A.indptr))
    A.data = np.log1p(A.data / per_nz * scale_factor)
    return A

"""IVSparse wire formats: the reference's development input path for atlas-scale data.

Mirrors the Rcpp entry points of reference src/singlet.cpp:783-995 with the same names, arguments and file names:
``save_IVSparse`` / ``build_IVCSC2`` / ``write_IVCSC`` write the IVCSC image of a dgCMatrix list, ``read_IVSparse`` reads it
back, and ``run_nmf_on_sparsematrix_list`` is what ``run_nmf`` calls for a list input (R/run_nmf.R:33). The codec is
``csrc/ivsparse.cpp`` behind ``sgl_ivsparse_*`` (host code of the C-ABI library; it needs no device), the fit is the same
device path as every other entry point: the image is decoded a column range at a time into dgCMatrix chunk views and
handed to ``sgl_nmf`` / ``sgl_multi_nmf`` as a chunk list, t(A) being built on the device.

What the reference's IVSparse path does differently from ``c_nmf`` and what is kept: values are narrowed to ``float``
(``Eigen::SparseMatrix<float>``, :790-822) -- kept, the decoded values are exactly those floats; the products are summed in
the file's value-grouped order (:745-775) -- not kept, the engine sums in row order (a rounding-level difference);
``run_nmf`` passes no L1/L2 to this path, so they default to 0 (:938) -- kept.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

IVCSC_FILE = "IVCSC_matrix.ivsparse"            # save_IVSparse / build_IVCSC2 / read_IVSparse (src/singlet.cpp:911,934,940)
WRITE_IVCSC_FILE = "IVSparse_matrix.ivsparse"   # write_IVCSC (:895)
# write_IVCSC names its second file with a trailing newline (:901); kept verbatim so files interchange
WRITE_IVCSC_T_FILE = "IVSparse_matrix_transpose.ivsparse\n"


def _as_list(A):
    from .api import _as_csc

    return _as_csc(list(A) if isinstance(A, (list, tuple)) else [A])


def encode(A, level: int = 3) -> np.ndarray:
    """The file image (uint8 array) the reference's IVCSC (``level`` 3) / VCSC (2) type writes for a dgCMatrix or a list of
    column chunks (``IVCSC::append`` semantics: chunks are concatenated by columns)."""
    lib = _lib.load()
    chunks, n, keep = _lib.chunks_to_c(_as_list(A))
    size = lib.sgl_ivsparse_encode(chunks, n, int(level), None, 0)
    if size < 0:
        _lib.check(int(size))
    out = np.empty(int(size), np.uint8)
    rc = lib.sgl_ivsparse_encode(chunks, n, int(level), out.ctypes.data_as(C.c_void_p), out.nbytes)
    if rc < 0:
        _lib.check(int(rc))
    del keep
    return out


def info(image) -> dict:
    """Metadata of an image: compression level, dimensions, non-zeros, value width."""
    lib = _lib.load()
    image = np.ascontiguousarray(np.frombuffer(image, np.uint8) if not isinstance(image, np.ndarray) else image)
    level, vb = C.c_int(), C.c_int()
    nrow, ncol, nnz = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(lib.sgl_ivsparse_info(image.ctypes.data_as(C.c_void_p), image.nbytes, C.byref(level), C.byref(nrow), C.byref(ncol),
                                     C.byref(nnz), C.byref(vb)))
    return {"level": level.value, "nrow": nrow.value, "ncol": ncol.value, "nnz": nnz.value, "value_bytes": vb.value}


def decode(image, col0: int = 0, ncol: int | None = None):
    """Columns ``[col0, col0 + ncol)`` of an image as a scipy CSC matrix (float64 values that are exactly the stored ones,
    rows ascending within a column)."""
    import scipy.sparse as sp

    lib = _lib.load()
    image = np.ascontiguousarray(np.frombuffer(image, np.uint8) if not isinstance(image, np.ndarray) else image)
    md = info(image)
    if ncol is None:
        ncol = md["ncol"] - col0
    ip = image.ctypes.data_as(C.c_void_p)
    p = np.zeros(max(int(ncol), 0) + 1, np.int32)
    nnz = lib.sgl_ivsparse_decode(ip, image.nbytes, int(col0), int(ncol), p.ctypes.data_as(C.c_void_p), None, None, 0)
    if nnz < 0:
        _lib.check(int(nnz))
    i = np.empty(int(nnz), np.int32)
    x = np.empty(int(nnz), np.float64)
    rc = lib.sgl_ivsparse_decode(ip, image.nbytes, int(col0), int(ncol), p.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p),
                                 x.ctypes.data_as(C.c_void_p), int(nnz))
    if rc < 0:
        _lib.check(int(rc))
    M = sp.csc_matrix((x, i, p), shape=(md["nrow"], int(ncol)))
    M.has_sorted_indices = True
    return M


def decode_chunks(image, max_nnz: int = 1 << 30):
    """The whole image as a list of column chunks of at most ``max_nnz`` non-zeros each (a dgCMatrix holds < 2^31): the
    chunk list ``c_nmf_sparse_list`` / ``sgl_multi_nmf`` take for matrices past the 32-bit limit."""
    lib = _lib.load()
    image = np.ascontiguousarray(np.frombuffer(image, np.uint8) if not isinstance(image, np.ndarray) else image)
    md = info(image)
    ip = image.ctypes.data_as(C.c_void_p)
    out, col0 = [], 0
    step = max(1, md["ncol"] if md["nnz"] <= max_nnz else int(md["ncol"] * (max_nnz / md["nnz"]) * 0.9))
    while col0 < md["ncol"]:
        nc = min(step, md["ncol"] - col0)
        while True:
            p = np.zeros(nc + 1, np.int32)
            nnz = lib.sgl_ivsparse_decode(ip, image.nbytes, col0, nc, p.ctypes.data_as(C.c_void_p), None, None, 0)
            if nnz >= 0 and nnz <= max_nnz or nc == 1:
                break
            nc = max(1, nc // 2)
        out.append(decode(image, col0, nc))
        col0 += nc
    return out


# ---- the reference's entry points -------------------------------------------------------------------------------------------
def save_IVSparse(A_, verbose: bool = True, directory: str = ".") -> bool:
    """``save_IVSparse`` (src/singlet.cpp:907-913): dgCMatrix list -> ``IVCSC_matrix.ivsparse`` in the working directory."""
    img = encode(A_, 3)
    if verbose:
        print("writing to IVCSC_matrix.ivsparse")
    img.tofile(os.path.join(directory, IVCSC_FILE))
    return True


def build_IVCSC2(L, verbose: bool = True, directory: str = ".") -> bool:
    """``build_IVCSC2`` (src/singlet.cpp:915-937): same image, built chunk by chunk and appended in the reference."""
    return save_IVSparse(L, verbose, directory)


def write_IVCSC(L, verbose: bool = True, directory: str = ".") -> bool:
    """``write_IVCSC`` (src/singlet.cpp:844-905): the image of the list and the image of its transpose."""
    import scipy.sparse as sp

    mats = _as_list(L)
    if verbose:
        print("writing IVSparse matrix")
    encode(mats, 3).tofile(os.path.join(directory, WRITE_IVCSC_FILE))
    if verbose:
        print("transposing IVSparse matrix")
    # narrowing to float happens before the transpose in the reference (:851-880); transposing commutes with it
    At = sp.hstack(mats, format="csc").T.tocsc()
    At.sort_indices()
    if verbose:
        print("writing transposed IVSparse matrix")
    encode(At, 3).tofile(os.path.join(directory, WRITE_IVCSC_T_FILE))
    return True


def read_IVSparse(directory: str = "."):
    """``read_IVSparse`` (src/singlet.cpp:939-944): ``IVCSC_matrix.ivsparse`` as a sparse matrix (float values)."""
    return decode(np.fromfile(os.path.join(directory, IVCSC_FILE), np.uint8))


def run_nmf_on_sparsematrix_list(A_, tol, maxit, verbose, threads, w, use_vcsc: bool = False, L1: float = 0.0, L2: float = 0.0,
                                 handle=None, multi=None):
    """``run_nmf_on_sparsematrix_list`` (src/singlet.cpp:946-995): the list is packed into one IVCSC (or, with ``use_vcsc``,
    VCSC) matrix and factorised with plain ALS; returns ``{"w": k x m, "d": k, "h": k x n}`` unsorted like the reference.
    Here the image is decoded into column chunks on the host and the fit runs on the device (``sgl_nmf`` with At = NULL;
    ``multi``: a ``multi.MultiGPU`` to spread the chunks over several GPUs)."""
    from . import api

    mats = _as_list(A_)
    w = np.asarray(w, np.float64)
    if w.shape[1] != mats[0].shape[0]:
        raise ValueError("number of rows in 'w' and 'A' is incompatible!")
    image = encode(mats, 2 if use_vcsc else 3)
    # one chunk per input matrix keeps the caller's sharding; the values are now the reference's floats
    bounds = np.cumsum([0] + [a.shape[1] for a in mats])
    chunks = [decode(image, int(bounds[q]), int(bounds[q + 1] - bounds[q])) for q in range(len(mats))]
    if multi is not None:
        out = multi.c_nmf_sparse_list(chunks, None, tol, maxit, verbose, L1, L2, threads, w)
    else:
        out = api.c_nmf_sparse_list(chunks, None, tol, maxit, verbose, L1, L2, threads, w, handle)
    return {"w": out["w"], "d": out["d"], "h": out["h"]}

"""Host-side mirror of the reference's operator interface for the ALS-NMF path.

R is not installed in this image, so the matrix-level R API (``run_nmf``, ``cross_validate_nmf``,
``ard_nmf``, ``project_model``, ``GetBestRank``) and the Rcpp entry points they call (``c_nmf``,
``c_ard_nmf``, ``c_project_model``, ``c_nmf_sparse_list``, ``c_ard_nmf_sparse_list``,
``Rcpp_predict``) are mirrored here in Python with the same names, argument order, defaults and
returned objects, on top of the C ABI (``include/singlet_cuda.h``). The Rcpp glue a maintainer
would drop into the R package lives in ``rglue/``.

Matrices are scipy CSC (the dgCMatrix analogue); factor matrices are numpy arrays laid out like
the engine's (``w`` is k x m, ``h`` is k x n) until the R-level wrappers transpose/sort them exactly
as R/run_nmf.R:65-75 does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys

import numpy as np
import pandas as pd

from . import _lib
from .rrng import RRng

# ------------------------------------------------------------------------------------------------
# R's global RNG (set.seed / runif / .Random.seed)
# ------------------------------------------------------------------------------------------------
_RNG = RRng(0)


def set_seed(seed: int) -> None:
    """R's ``set.seed``: w_init and the mask seeds below are drawn from this stream."""
    _RNG.set_seed(seed)


def _rng(rng):
    return _RNG if rng is None else rng


# ------------------------------------------------------------------------------------------------
# handle management
# ------------------------------------------------------------------------------------------------
class Handle:
    """``sgl_handle``: one device + one stream + the upload/mask cache."""

    def __init__(self, device: int = 0, stream: int | None = None):
        lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(lib.sgl_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.lib, self.device = lib, device

    @property
    def ptr(self):
        return self._h

    def synchronize(self):
        _lib.check(self.lib.sgl_synchronize(self._h))

    def launch_count(self) -> int:
        return int(self.lib.sgl_launch_count(self._h))

    def set_precision(self, mode):
        """``"mixed16"`` (default: FP16-staged operands of the sparse product on large matrices at padded rank >= 32, FP32
        accumulation), ``"fp32"``, or ``"mixed16_always"`` (16-bit staging whatever the matrix size) -- sgl_set_precision."""
        code = {"mixed16": _lib.PRECISION_MIXED16, "fp32": _lib.PRECISION_FP32, "mixed16_always": _lib.PRECISION_MIXED16_ALWAYS}.get(mode, mode)
        _lib.check(self.lib.sgl_set_precision(self._h, int(code)))

    def set_cache(self, enabled: bool):
        _lib.check(self.lib.sgl_set_cache(self._h, int(bool(enabled))))

    def close(self):
        if self._h:
            self.lib.sgl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT = None


def default_handle() -> Handle:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Handle(0)
    return _DEFAULT


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_csc(A):
    import scipy.sparse as sp

    if isinstance(A, (list, tuple)) and not (len(A) == 5 and np.isscalar(A[3])):
        return [_as_csc(a) for a in A]
    if sp.issparse(A):
        A = A.tocsc()
        if A.dtype != np.float64:
            A = A.astype(np.float64)
        if not A.has_sorted_indices:
            A.sort_indices()
        return A
    if isinstance(A, tuple):
        return A
    raise TypeError("expected a scipy sparse matrix (dgCMatrix analogue) or a list of them")


def _at_chunks(At):
    """``At`` for the C ABI: the caller's transpose, or (NULL, 0) to have it built on the device (row f1)."""
    if At is None:
        return None, 0, None
    return _lib.chunks_to_c(_as_csc(At))


def _host_t(A):
    At = A.T.tocsc()
    At.sort_indices()
    return At


def _host_transpose_needed(A, device_transpose):
    return not device_transpose


def _ncols(A):
    if isinstance(A, list):
        return sum(_ncols(a) for a in A)
    return A.shape[1] if hasattr(A, "shape") else int(A[4])


def _nrows(A):
    if isinstance(A, list):
        return _nrows(A[0])
    return A.shape[0] if hasattr(A, "shape") else int(A[3])


def _callbacks(verbose: bool, masked: bool):
    """Prints exactly what the reference prints (src/singlet.cpp:643-644, 661-662, 1102-1127)."""
    if not verbose:
        return None, None
    if masked:
        sys.stdout.write("\n%4s | %8s | %8s \n---------------------------\n" % ("iter", "tol", "overfit"))
    else:
        sys.stdout.write("\n%4s | %8s \n---------------\n" % ("iter", "tol"))

    def on_iter(_user, it, tol, overfit):
        if not masked:
            sys.stdout.write("%4d | %8.2e\n" % (it, tol))
        elif math.isnan(overfit):
            sys.stdout.write("%4d | %8.2e | %8s\n" % (it, tol, "-"))
        else:
            sys.stdout.write("%4d | %8.2e | %8.2e\n" % (it, tol, overfit))

    fn = _lib.ITER_FN(on_iter)
    cb = _lib.Callbacks(None, _lib.POLL_FN(), fn)
    return cb, fn


# ------------------------------------------------------------------------------------------------
# Rcpp entry points (same names / argument order as R/RcppExports.R)
# ------------------------------------------------------------------------------------------------
def c_nmf(A, At, tol, maxit, verbose, L1_w, L1_h, L2_w, L2_h, threads, w, handle: Handle | None = None):
    """``c_nmf`` (reference src/singlet.cpp:669-672; R/RcppExports.R:28-30). ``threads`` is accepted and
    ignored (the GPU replaces the OpenMP team). Returns ``{"w": k x m, "d": k, "h": k x n}`` plus the
    extra keys ``iter`` and ``tol`` (final values; the reference only prints them)."""
    h = handle or default_handle()
    A = _as_csc(A)
    a, na, k1 = _lib.chunks_to_c(A)
    at, nat, k2 = _at_chunks(At)
    wk = np.array(w, dtype=np.float64, order="F")
    if wk.ndim != 2:
        raise ValueError("w must be a k x m matrix")
    k, m = wk.shape
    if m != _nrows(A):
        raise ValueError("w must have nrow(A) columns")
    n = _ncols(A)
    d = np.zeros(k)
    hh = np.zeros((k, n), order="F")
    iters, ftol = C.c_int32(0), C.c_double(0)
    cb, keep = _callbacks(bool(verbose), False)
    _lib.check(h.lib.sgl_nmf(h.ptr, a, na, at, nat, float(tol), int(maxit) & 0xFFFF, float(L1_w), float(L1_h), float(L2_w),
                             float(L2_h), k, _dp(wk), _dp(d), _dp(hh), C.addressof(iters), C.addressof(ftol),
                             C.addressof(cb) if cb is not None else None))
    return {"w": wk, "d": d, "h": hh, "iter": iters.value, "tol": ftol.value}


def c_nmf_sparse_list(A_, At_, tol, maxit, verbose, L1, L2, threads, w, handle: Handle | None = None):
    """``c_nmf_sparse_list`` (reference src/singlet.cpp:715-743): one L1/L2 for both factors."""
    return c_nmf(list(A_), None if At_ is None else list(At_), tol, maxit, verbose, L1, L1, L2, L2, threads, w, handle)


def c_nmf_dense(A, At, tol, maxit, verbose, L1_w, L1_h, L2_w, L2_h, threads, w, handle: Handle | None = None):
    """``c_nmf_dense`` (reference src/singlet.cpp:1051-1054; R/RcppExports.R:70-72): A (m x n) and At (n x m) are
    dense numpy matrices."""
    h = handle or default_handle()
    Ad, Atd = np.asfortranarray(A, dtype=np.float64), np.asfortranarray(At, dtype=np.float64)
    wk = np.array(w, dtype=np.float64, order="F")
    k, m = wk.shape
    n = Ad.shape[1]
    if Ad.shape[0] != m or Atd.shape != (n, m):
        raise ValueError("shapes of A, At and w do not agree")
    d, hh = np.zeros(k), np.zeros((k, n), order="F")
    iters, ftol = C.c_int32(0), C.c_double(0)
    cb, keep = _callbacks(bool(verbose), False)
    _lib.check(h.lib.sgl_nmf_dense(h.ptr, _dp(Ad), _dp(Atd), m, n, float(tol), int(maxit) & 0xFFFF, float(L1_w), float(L1_h),
                                   float(L2_w), float(L2_h), k, _dp(wk), _dp(d), _dp(hh), C.addressof(iters), C.addressof(ftol),
                                   C.addressof(cb) if cb is not None else None))
    return {"w": wk, "d": d, "h": hh, "iter": iters.value, "tol": ftol.value}


def c_ard_nmf_dense(A, At, tol, maxit, verbose, L1, L2, threads, w, seed, inv_density, overfit_threshold, trace_test_mse,
                    handle: Handle | None = None):
    """``c_ard_nmf_dense`` (reference src/singlet.cpp:1357-1361; R/RcppExports.R:86-88)."""
    h = handle or default_handle()
    Ad, Atd = np.asfortranarray(A, dtype=np.float64), np.asfortranarray(At, dtype=np.float64)
    wk = np.array(w, dtype=np.float64, order="F")
    k, m = wk.shape
    n = Ad.shape[1]
    if Ad.shape[0] != m or Atd.shape != (n, m):
        raise ValueError("shapes of A, At and w do not agree")
    d, hh = np.zeros(k), np.zeros((k, n), order="F")
    cap = (int(maxit) & 0xFFFF) + 2
    mse, ft, so, it = np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap, np.int32)
    tr = _lib.Trace(mse.ctypes.data, it.ctypes.data, ft.ctypes.data, so.ctypes.data, cap, 0)
    cb, keep = _callbacks(bool(verbose), True)
    _lib.check(h.lib.sgl_ard_nmf_dense(h.ptr, _dp(Ad), _dp(Atd), m, n, float(tol), int(maxit) & 0xFFFF, float(L1), float(L2), k,
                                       _dp(wk), _dp(d), _dp(hh), int(seed) & 0xFFFFFFFFFFFFFFFF, int(inv_density),
                                       float(overfit_threshold), int(trace_test_mse) & 0xFFFF, C.addressof(tr),
                                       C.addressof(cb) if cb is not None else None))
    q = tr.length
    return {"w": wk, "d": d, "h": hh, "test_mse": mse[:q].copy(), "iter": it[:q].copy(), "tol": ft[:q].copy(),
            "score_overfit": so[:q].copy()}


def c_linked_nmf(A, At, tol, maxit, verbose, L1, L2, threads, w, link_h, link_w, handle: Handle | None = None):
    """``c_linked_nmf`` (reference src/singlet.cpp:1059-1086; R/RcppExports.R:74-76): ALS NMF where ``b`` is
    multiplied by a column of ``link_h`` / ``link_w`` before every solve (``predict_link`` :416-433). A side is
    linked only when its matrix has one column per cell / per gene, like the reference."""
    h = handle or default_handle()
    A = _as_csc(A)
    a, na, k1 = _lib.chunks_to_c(A)
    at, nat, k2 = _at_chunks(At)
    wk = np.array(w, dtype=np.float64, order="F")
    k, m = wk.shape
    n = _ncols(A)
    lh = np.asfortranarray(link_h, dtype=np.float64)
    lw = np.asfortranarray(link_w, dtype=np.float64)
    d = np.zeros(k)
    hh = np.zeros((k, n), order="F")
    iters, ftol = C.c_int32(0), C.c_double(0)
    cb, keep = _callbacks(bool(verbose), False)
    _lib.check(h.lib.sgl_linked_nmf(h.ptr, a, at, float(tol), int(maxit) & 0xFFFF, float(L1), float(L2), k, _dp(wk), _dp(d), _dp(hh),
                                    _dp(lh), lh.shape[0], lh.shape[1], _dp(lw), lw.shape[0], lw.shape[1], C.addressof(iters),
                                    C.addressof(ftol), C.addressof(cb) if cb is not None else None))
    return {"w": wk, "d": d, "h": hh, "iter": iters.value, "tol": ftol.value}


def weight_by_split(A, split_by, n_groups):
    """``weight_by_split`` (reference src/singlet.cpp:119-144): returns a copy of the CSC matrix in which the
    columns of every group g != 0 are divided by ``sum(group g) / sum(group 0)`` so that all groups carry the same
    total weight (pre-processing of ``RunNMF(split.by = ...)``, R/RunNMF.R:86-97). Pure host O(nnz) work."""
    A = _as_csc(A).copy()
    split_by = np.asarray(split_by, dtype=np.int64)
    per_nz = np.repeat(split_by, np.diff(A.indptr))
    sums = np.bincount(per_nz, weights=A.data, minlength=n_groups).astype(np.float64)
    factor = np.ones(n_groups)
    factor[1:] = sums[1:] / sums[0]
    A.data = np.where(per_nz != 0, A.data / factor[per_nz], A.data)
    return A


def c_ard_nmf(A, At, tol, maxit, verbose, L1, L2, threads, w, seed, inv_density, overfit_threshold, trace_test_mse,
              handle: Handle | None = None):
    """``c_ard_nmf`` (reference src/singlet.cpp:1155-1159; R/RcppExports.R:70-72). Returns
    ``w, d, h, test_mse, iter, tol, score_overfit`` like src/singlet.cpp:1144-1151."""
    h = handle or default_handle()
    A = _as_csc(A)
    a, na, k1 = _lib.chunks_to_c(A)
    at, nat, k2 = _at_chunks(At)
    wk = np.array(w, dtype=np.float64, order="F")
    k, m = wk.shape
    if m != _nrows(A):
        raise ValueError("w must have nrow(A) columns")
    n = _ncols(A)
    d = np.zeros(k)
    hh = np.zeros((k, n), order="F")
    cap = (int(maxit) & 0xFFFF) + 2
    mse, ft, so, it = np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap, np.int32)
    tr = _lib.Trace(mse.ctypes.data, it.ctypes.data, ft.ctypes.data, so.ctypes.data, cap, 0)
    cb, keep = _callbacks(bool(verbose), True)
    _lib.check(h.lib.sgl_ard_nmf(h.ptr, a, na, at, nat, float(tol), int(maxit) & 0xFFFF, float(L1), float(L2), k, _dp(wk),
                                 _dp(d), _dp(hh), int(seed) & 0xFFFFFFFFFFFFFFFF, int(inv_density), float(overfit_threshold),
                                 int(trace_test_mse) & 0xFFFF, C.addressof(tr), C.addressof(cb) if cb is not None else None))
    q = tr.length
    return {"w": wk, "d": d, "h": hh, "test_mse": mse[:q].copy(), "iter": it[:q].copy(), "tol": ft[:q].copy(),
            "score_overfit": so[:q].copy()}


def c_ard_nmf_batch(A, At, tol, maxit, L1, L2, threads, ws, seeds, inv_density, overfit_threshold, trace_test_mse,
                    concurrency: int = 0, handle: Handle | None = None):
    """Batched ``c_ard_nmf``: the independent fits of a rank search -- the (rank, replicate)
    loop of reference R/cross_validate_nmf.R:69-97 -- run ``concurrency`` at a time inside the library
    (``sgl_ard_nmf_batch``), sharing one uploaded A / At. ``ws`` is a list of k_j x m initial factors and ``seeds``
    the mask seed of each fit; returns one ``c_ard_nmf`` result dict per fit, bit-identical to sequential calls."""
    h = handle or default_handle()
    A = _as_csc(A)
    a, na, k1 = _lib.chunks_to_c(A)
    at, nat, k2 = _at_chunks(At)
    m, n = _nrows(A), _ncols(A)
    cap = (int(maxit) & 0xFFFF) + 2
    jobs = (_lib.FitJob * len(ws))()
    outs, keep = [], []
    for j, (w, seed) in enumerate(zip(ws, seeds)):
        wk = np.array(w, dtype=np.float64, order="F")
        k = wk.shape[0]
        if wk.shape[1] != m:
            raise ValueError("w must have nrow(A) columns")
        d, hh = np.zeros(k), np.zeros((k, n), order="F")
        mse, ft, so, it = np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap, np.int32)
        tr = _lib.Trace(mse.ctypes.data, it.ctypes.data, ft.ctypes.data, so.ctypes.data, cap, 0)
        keep.append(tr)
        jobs[j] = _lib.FitJob(k, 0, int(seed) & 0xFFFFFFFFFFFFFFFF, wk.ctypes.data, d.ctypes.data, hh.ctypes.data, C.addressof(tr))
        outs.append((wk, d, hh, mse, it, ft, so, tr))
    cb, keep_cb = _callbacks(False, True)
    _lib.check(h.lib.sgl_ard_nmf_batch(h.ptr, a, na, at, nat, float(tol), int(maxit) & 0xFFFF, float(L1), float(L2),
                                       int(inv_density), float(overfit_threshold), int(trace_test_mse) & 0xFFFF, jobs, len(ws),
                                       int(concurrency), C.addressof(cb) if cb is not None else None))
    res = []
    for wk, d, hh, mse, it, ft, so, tr in outs:
        q = tr.length
        res.append({"w": wk, "d": d, "h": hh, "test_mse": mse[:q].copy(), "iter": it[:q].copy(), "tol": ft[:q].copy(),
                    "score_overfit": so[:q].copy()})
    return res


def c_ard_nmf_sparse_list(A_, At_, tol, maxit, verbose, L1, L2, threads, w, rng_seed, inv_density, overfit_threshold,
                          trace_test_mse, handle: Handle | None = None):
    """``c_ard_nmf_sparse_list`` (reference src/singlet.cpp:1162-1234)."""
    return c_ard_nmf(list(A_), None if At_ is None else list(At_), tol, maxit, verbose, L1, L2, threads, w, rng_seed, inv_density,
                     overfit_threshold,
                     trace_test_mse, handle)


def c_project_model(A, w, L1, L2, threads, handle: Handle | None = None):
    """``c_project_model`` (reference src/singlet.cpp:405-413): returns ``{"h": k x n, "d": k}``."""
    h = handle or default_handle()
    A = _as_csc(A)
    a, na, keep = _lib.chunks_to_c(A)
    wf = np.array(w, dtype=np.float64, order="F")
    rows, cols = wf.shape
    m = _nrows(A)
    k = cols if rows == m else rows
    hh = np.zeros((k, _ncols(A)), order="F")
    d = np.zeros(k)
    _lib.check(h.lib.sgl_project_model(h.ptr, a, na, _dp(wf), rows, cols, float(L1), float(L2), _dp(hh), _dp(d)))
    return {"h": hh, "d": d}


def Rcpp_predict(A, w, L1, L2, threads, handle: Handle | None = None):
    """``Rcpp_predict`` (reference src/singlet.cpp:350-367): one H update from zero, unscaled."""
    h = handle or default_handle()
    A = _as_csc(A)
    a, na, keep = _lib.chunks_to_c(A)
    wf = np.array(w, dtype=np.float64, order="F")
    rows, cols = wf.shape
    m = _nrows(A)
    k = cols if (rows == m and cols != m) else rows
    hh = np.zeros((k, _ncols(A)), order="F")
    _lib.check(h.lib.sgl_predict(h.ptr, a, na, _dp(wf), rows, cols, float(L1), float(L2), _dp(hh)))
    return hh


# ------------------------------------------------------------------------------------------------
# R-level API
# ------------------------------------------------------------------------------------------------
def _distributed_transpose(A_list):
    """Gene-block transposes of a column-chunk list (reference R/cross_validate_nmf.R:37-50):
    ``block_sizes <- floor(c(seq(1, nrow, nrow / length(A)), nrow + 1))`` (1-based)."""
    import scipy.sparse as sp

    m, L = A_list[0].shape[0], len(A_list)
    starts = [int(math.floor(1 + i * (m / L))) for i in range(L)]
    bounds = [s for s in starts if s <= m] + [m + 1]
    out = []
    for i in range(len(bounds) - 1):
        lo, hi = bounds[i] - 1, bounds[i + 1] - 1  # 0-based [lo, hi)
        blocks = [a[lo:hi, :].T.tocsc() for a in A_list]
        out.append(sp.vstack(blocks).tocsc())
    return out


def _take_last_axis(a, idx):
    """``np.take(a, idx, axis=1)`` for a (rows, k) array; large arrays are cut into row blocks handled by a few threads
    (numpy releases the GIL inside take): 0.12 s -> 0.04 s for the 10^6 x 32 h of the headline config."""
    rows = a.shape[0]
    if a.size < (1 << 22) or not a.flags.c_contiguous:
        return np.take(a, idx, axis=1)
    import concurrent.futures as cf

    out = np.empty((rows, len(idx)), dtype=a.dtype)
    n_threads = min(8, os.cpu_count() or 1)
    block = (rows + n_threads - 1) // n_threads

    def work(s):
        np.take(a[s:s + block], idx, axis=1, out=out[s:s + block])

    with cf.ThreadPoolExecutor(n_threads) as ex:
        list(ex.map(work, range(0, rows, block)))
    return out


def _sort_model(model, rank):
    """R/run_nmf.R:65-71: order by d decreasing; w <- t(w)[, idx] (m x k); h <- h[idx, ]."""
    idx = np.argsort(-model["d"], kind="stable")
    model["d"] = model["d"][idx]
    # the engine's factors are column-major k x cols, i.e. C-ordered (cols, k) arrays seen through .T: permuting the LAST
    # axis of those views is one streaming pass (h is k x n = 256 MB at the headline config), no layout conversion
    model["w"] = _take_last_axis(np.asarray(model["w"]).T, idx)        # m x k
    model["h"] = _take_last_axis(np.asarray(model["h"]).T, idx).T      # k x n (column-major)
    model["names"] = ["NMF_%d" % (i + 1) for i in range(model["w"].shape[1])]
    return model


def run_nmf(A, rank, tol=1e-4, maxit=100, verbose=True, L1=0.01, L2=0, threads=0, compression_level=3, rng=None,
            handle: Handle | None = None, device_transpose: bool = True):
    """``run_nmf`` (reference R/run_nmf.R:18-77). Returns ``{"w": m x k, "d": k, "h": k x n}`` sorted by
    ``d``. A list input goes where the reference sends it (R/run_nmf.R:21-35): to
    ``run_nmf_on_sparsematrix_list`` (``singlet_b200.ivsparse``) -- the chunks packed as one IVCSC / VCSC
    (``compression_level`` 2) image with float values, plain ALS with NO penalties (the reference does not forward
    L1/L2 on this path). For penalised fits of a chunk list call ``c_nmf_sparse_list``. ``device_transpose`` (default):
    ``t(A)`` of R/run_nmf.R:40 is built on the device (``sgl_matrix_transpose``, bit-identical records) instead of on the host."""
    r = _rng(rng)
    L1 = (L1, L1) if np.isscalar(L1) else (L1[0], L1[1] if len(L1) == 2 else L1[0])
    L2 = (L2, L2) if np.isscalar(L2) else (L2[0], L2[1] if len(L2) == 2 else L2[0])
    if isinstance(A, (list, tuple)):
        from .ivsparse import run_nmf_on_sparsematrix_list

        A = _as_csc(list(A))
        if len({a.shape[0] for a in A}) != 1:
            raise ValueError("number of rows in all provided 'A' matrices are not identical")
        w_init = r.matrix_runif(rank, A[0].shape[0])
        model = run_nmf_on_sparsematrix_list(A, tol, maxit, verbose, threads, w_init, compression_level == 2, handle=handle)
    else:
        if verbose:
            print("running with sparse optimization")
        A = _as_csc(A)
        At = _host_t(A) if _host_transpose_needed(A, device_transpose) else None
        w_init = r.matrix_runif(rank, A.shape[0])
        model = c_nmf(A, At, tol, maxit, verbose, L1[0], L1[1], L2[0], L2[1], threads, w_init, handle)
    return _sort_model(model, rank)


def project_model(A, w, L1=0.01, L2=0, threads=0, handle: Handle | None = None):
    """``project_model`` (reference R/ProjectData.R:11-19)."""
    w = np.asarray(w)
    if w.shape[0] != _nrows(A) and w.shape[1] != _nrows(A):
        raise ValueError("'w' must share a common edge with the rows of 'A'")
    return c_project_model(A, w, L1, L2, threads, handle)


def cross_validate_nmf(A, ranks, n_replicates=3, tol=1e-4, maxit=100, verbose=1, L1=0.01, L2=0, threads=0,
                       test_density=0.05, tol_overfit=1e-4, trace_test_mse=5, rng=None, handle: Handle | None = None,
                       batch: bool = True, concurrency: int = 0, device_transpose: bool = True):
    """``cross_validate_nmf`` (reference R/cross_validate_nmf.R:18-105). Returns a pandas DataFrame with
    columns ``k, rep, test_error, iter, tol`` (one row per traced iteration of every fit).

    ``batch`` (default) hands the whole (rank, replicate) grid to ``sgl_ard_nmf_batch``, which runs several of the
    independent fits at a time on the device; the rows are identical to the fit-by-fit loop (``batch=False``, also
    used when ``verbose > 1`` so that every fit can print its trace like the reference)."""
    if L1 >= 1:
        raise ValueError("L1 penalty must be strictly in the range (0, 1]")
    r = _rng(rng)
    ranks = [int(k) for k in np.atleast_1d(ranks)]
    sparse_list = isinstance(A, (list, tuple))
    if sparse_list:
        A = _as_csc(list(A))
        At = _distributed_transpose(A) if _host_transpose_needed(A, device_transpose) else None
        m = A[0].shape[0]
    else:
        A = _as_csc(A)
        At = _host_t(A) if _host_transpose_needed(A, device_transpose) else None
        m = A.shape[0]
    w_init = [r.matrix_runif(max(ranks), m) for _ in range(n_replicates)]
    rows = []
    inv_density = int(round(1 / test_density))
    if batch and verbose <= 1:
        grid = [(k, rep) for rep in range(1, n_replicates + 1) for k in ranks]
        seeds = [abs(r.dot_random_seed(3 + rep)) for _, rep in grid]
        models = c_ard_nmf_batch(A, At, tol, maxit, L1, L2, threads, [w_init[rep - 1][:k, :] for k, rep in grid], seeds,
                                 inv_density, tol_overfit, trace_test_mse, concurrency, handle)
        for (k, rep), model in zip(grid, models):
            for q in range(len(model["test_mse"])):
                rows.append({"k": k, "rep": rep, "test_error": model["test_mse"][q], "iter": int(model["iter"][q]),
                             "tol": model["tol"][q]})
        return pd.DataFrame(rows, columns=["k", "rep", "test_error", "iter", "tol"])
    # expand.grid(k = ranks, rep = 1:n_replicates): k varies fastest
    for rep in range(1, n_replicates + 1):
        for k in ranks:
            seed = abs(r.dot_random_seed(3 + rep))  # abs(.Random.seed[[3 + rep]])
            fn = c_ard_nmf_sparse_list if sparse_list else c_ard_nmf
            model = fn(A, At, tol, maxit, verbose > 1, L1, L2, threads, w_init[rep - 1][:k, :], seed,
                       int(round(1 / test_density)), tol_overfit, trace_test_mse, handle)
            for q in range(len(model["test_mse"])):
                rows.append({"k": k, "rep": rep, "test_error": model["test_mse"][q], "iter": int(model["iter"][q]),
                             "tol": model["tol"][q]})
            if verbose > 1:
                print("test set error: %#.4e\n" % model["test_mse"][-1])
    return pd.DataFrame(rows, columns=["k", "rep", "test_error", "iter", "tol"])


def GetBestRank(df, tol_overfit=1e-4):
    """``GetBestRank`` (reference R/GetBestRank.R:8-46): lowest rank minimising the held-out error among
    the ranks that do not overfit, per replicate; floor of the mean over replicates."""
    best_ranks = []
    for rep in sorted(df["rep"].unique()):
        df_rep = df[df["rep"] == rep]
        max_rank = df_rep["k"].max() + 1
        for rank in df_rep["k"].unique():
            if rank < max_rank:
                err = df_rep[df_rep["k"] == rank]["test_error"].to_numpy()
                if len(err) > 1:
                    v2, v1 = err[1:].copy(), err[:-1].copy()
                    if len(v1) >= 2:
                        for pos in range(1, len(v1)):
                            if v1[pos] > v1[pos - 1]:
                                v1[pos] = v1[pos - 1]
                    if max(0.0, float(np.max((v2 - v1) / (v2 + v1)))) > tol_overfit:
                        max_rank = rank
        df_rep = df_rep[df_rep["k"] < max_rank]
        if len(df_rep) == 0:
            best_ranks.append(2)
        elif len(df) == 1:
            best_ranks.append(int(df_rep["k"].iloc[0]))
        else:
            last = df_rep.loc[df_rep.groupby("k")["iter"].idxmax()]
            best_ranks.append(int(last["k"].iloc[int(np.argmin(last["test_error"].to_numpy()))]))
    return int(math.floor(np.mean(best_ranks)))


def ard_nmf(A, k_init=2, k_max=100, k_min=2, n_replicates=1, tol=1e-5, cv_tol=1e-4, maxit=100, verbose=1, L1=0.01, L2=0,
            threads=0, test_density=0.05, learning_rate=1, tol_overfit=1e-3, trace_test_mse=1, rng=None,
            handle: Handle | None = None, device_transpose: bool = True):
    """``ard_nmf`` (reference R/ard_nmf.R:31-193): rank search by cross-validated fits, then a final
    unmasked fit at the best rank. Returns the sorted model plus ``cv_data`` (DataFrame)."""
    if not L1 < 1:
        raise ValueError("L1 penalty must be strictly in the range (0, 1]")
    if k_init is None or k_init < k_min:
        k_init = k_min
    if k_min < 2:
        raise ValueError("k_min cannot be less than 2")
    r = _rng(rng)
    sparse_list = isinstance(A, (list, tuple))
    if sparse_list:
        A = _as_csc(list(A))
        At = _distributed_transpose(A) if _host_transpose_needed(A, device_transpose) else None
        m = A[0].shape[0]
    else:
        A = _as_csc(A)
        At = _host_t(A) if _host_transpose_needed(A, device_transpose) else None
        m = A.shape[0]
    w_init = [r.matrix_runif(k_max, m) for _ in range(n_replicates)]
    test_seed = abs(r.dot_random_seed(3))
    inv_density = int(round(1 / test_density))
    cols = ["k", "rep", "test_error", "iter", "tol", "overfit_score"]
    df = pd.DataFrame(columns=cols)
    fit = c_ard_nmf_sparse_list if sparse_list else c_ard_nmf
    for curr_rep in range(1, n_replicates + 1):
        step_size, curr_rank = 1.0, int(k_init)
        while step_size >= 1 and k_min <= curr_rank <= k_max:
            if verbose > 0:
                print("k =", curr_rank, ", rep =", curr_rep)
            r.set_seed(test_seed)
            model = fit(A, At, cv_tol, maxit, verbose > 2, L1, L2, threads, w_init[curr_rep - 1][:curr_rank, :],
                        test_seed + curr_rep, inv_density, tol_overfit, trace_test_mse, handle)
            overfit_score = float(model["score_overfit"][-1])
            new = pd.DataFrame({"k": curr_rank, "rep": curr_rep, "test_error": model["test_mse"], "iter": model["iter"],
                                "tol": model["tol"], "overfit_score": overfit_score})
            df = new if len(df) == 0 else pd.concat([df, new], ignore_index=True)
            if overfit_score >= tol_overfit:
                k_max = curr_rank
            df_rep = df[df["rep"] == curr_rep].sort_values("k", kind="stable")
            sub = df_rep[df_rep["k"] < k_max]
            best_rank = GetBestRank(sub) if len(sub) else 2
            last = df_rep.loc[df_rep.groupby("k")["iter"].idxmax()].sort_values("k")
            ks = list(last["k"])
            if best_rank not in ks:
                break
            rank_ind = ks.index(best_rank)
            if rank_ind == len(ks) - 1:
                step_size *= 1 + learning_rate
                curr_rank = best_rank + int(math.floor(step_size))
            elif rank_ind == 0:
                if math.floor(step_size) < best_rank:
                    curr_rank = best_rank - int(math.floor(step_size))
                    step_size *= learning_rate + 1
                else:
                    curr_rank = best_rank // 2
            else:
                diff_lower, diff_higher = best_rank - ks[rank_ind - 1], ks[rank_ind + 1] - best_rank
                if diff_lower <= 1 and diff_higher <= 1:
                    break
                curr_rank = best_rank - diff_lower // 2 if diff_lower >= diff_higher else best_rank + diff_higher // 2
    best_rank = GetBestRank(df, tol_overfit)
    if verbose > 0:
        print("\nFitting final model at k =", best_rank)
    r.set_seed(test_seed)
    w0 = w_init[0][:best_rank, :]
    if sparse_list:
        model = c_nmf_sparse_list(A, At, tol, maxit, verbose > 2, L1, L2, threads, w0, handle)
    else:
        model = c_nmf(A, At, tol, maxit, verbose > 2, L1, L1, L2, L2, threads, w0, handle)
    model["cv_data"] = df
    return _sort_model(model, best_rank)

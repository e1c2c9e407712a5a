"""ctypes face of the CPU checker -- TEST INFRASTRUCTURE ONLY.

Loads oracle/liboracle.so (our FP64 restatement, ``orc_*``) and, when present,
oracle/_ref/libsinglet_ref.so (the reference's own functions compiled against the
Eigen/Rcpp shim, ``ref_*``). Only tests/, ``__graft_entry__.smoke()`` and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module; nothing under
``singlet_b200/`` does.

Both libraries expose the same call signatures, so ``Oracle(kind="port")`` and
``Oracle(kind="reference")`` are interchangeable in the tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class _Csc(C.Structure):
    _fields_ = [
        ("nrow", C.c_int64),
        ("ncol", C.c_int64),
        ("p", C.c_void_p),
        ("i", C.c_void_p),
        ("x", C.c_void_p),
    ]


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or (
        os.path.exists("/root/reference/src/singlet.cpp")
        and not os.path.exists(os.path.join(_HERE, "_ref", "libsinglet_ref.so"))
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def have_reference() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libsinglet_ref.so"))


def _as_chunks(mats):
    """mats: one (p, i, x, nrow, ncol) tuple / scipy CSC matrix, or a list of them."""
    if not isinstance(mats, (list, tuple)) or (len(mats) == 5 and np.isscalar(mats[3])):
        mats = [mats]
    keep, arr = [], (_Csc * len(mats))()
    for q, m in enumerate(mats):
        if hasattr(m, "indptr"):
            p, i, x, nrow, ncol = m.indptr, m.indices, m.data, m.shape[0], m.shape[1]
        else:
            p, i, x, nrow, ncol = m
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        keep += [p, i, x]
        arr[q] = _Csc(int(nrow), int(ncol), p.ctypes.data, i.ctypes.data, x.ctypes.data)
    return arr, len(mats), keep


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """kind="port": oracle/liboracle.so; kind="reference": oracle/_ref/libsinglet_ref.so."""

    def __init__(self, kind: str = "port"):
        build()
        self.kind = kind
        if kind == "port":
            self.lib, self.pre = C.CDLL(os.path.join(_HERE, "liboracle.so")), "orc_"
        elif kind == "reference":
            if not have_reference():
                raise FileNotFoundError("oracle/_ref/libsinglet_ref.so not built (needs /root/reference)")
            self.lib, self.pre = C.CDLL(os.path.join(_HERE, "_ref", "libsinglet_ref.so")), "ref_"
        else:
            raise ValueError(kind)
        f = self._f
        u64, i32, i64, dbl, vp = C.c_uint64, C.c_int, C.c_int64, C.c_double, C.c_void_p
        f("rand1", u64, [u64, u64])
        f("rand2", u64, [u64, u64, u64])
        f("draw", i32, [u64, u64, u64, u64])
        f("cor", dbl, [vp, vp, u64])
        f("gram", None, [vp, i32, i64, vp])
        f("scale", None, [vp, i32, i64, vp])
        f("nnls", i32, [vp, vp, vp, i32, dbl, dbl])
        f("predict", None, [vp, i32, vp, i32, vp, dbl, dbl, i32, vp])
        f("predict_mask", None, [vp, i32, u64, u64, vp, i32, vp, dbl, dbl, i32, i32])
        f("mse_test", dbl, [vp, i32, vp, vp, vp, i32, u64, u64, i32])
        f("nmf", i32, [vp, i32, vp, i32, dbl, C.c_uint16, dbl, dbl, dbl, dbl, i32, i32, vp, vp, vp, vp])
        f("ard_nmf", i32, [vp, i32, vp, i32, dbl, C.c_uint16, dbl, dbl, i32, i32, vp, vp, vp, u64, u64, dbl,
                           C.c_uint16, vp, vp, vp, vp, i32, vp])
        f("project_model", None, [vp, i32, vp, i32, dbl, dbl, i32, vp, vp])
        f("linked_nmf", i32, [vp, vp, dbl, C.c_uint16, dbl, dbl, i32, i32, vp, vp, vp, vp, i32, i64, vp, i32, i64])
        f("max_threads", i32, [])
        if kind == "reference":
            f("nmf_dense", i32, [vp, vp, i64, i64, dbl, C.c_uint16, dbl, dbl, dbl, dbl, i32, i32, vp, vp, vp])
            f("ard_nmf_dense", i32, [vp, vp, i64, i64, dbl, C.c_uint16, dbl, dbl, i32, i32, vp, vp, vp, u64, u64, dbl, C.c_uint16,
                                     vp, vp, i32, vp])
        if kind == "port":
            f("mse_train", dbl, [vp, i32, vp, vp, vp, i32, u64, u64, i32])
            f("weight_by_split", None, [vp, vp, vp, i32])
            f("mask_cell", None, [u64, u64, u64, u64, vp])

    def _f(self, name, res, args):
        fn = getattr(self.lib, self.pre + name)
        fn.restype, fn.argtypes = res, args
        setattr(self, "_" + name, fn)

    # -- hash -------------------------------------------------------------------------------
    def rand1(self, state, i):
        return int(self._rand1(state, i))

    def rand2(self, state, i, j):
        return int(self._rand2(state, i, j))

    def draw(self, state, i, j, inv_density):
        return bool(self._draw(state, i, j, inv_density))

    def mask_cell(self, state, cell, n_genes, inv_density):
        out = np.zeros(n_genes, dtype=np.uint8)
        if self.kind == "port":
            self._mask_cell(state, cell, n_genes, inv_density, _dp(out))
        else:
            for g in range(n_genes):
                out[g] = self._draw(state, cell, g, inv_density)
        return out

    def max_threads(self):
        return int(self._max_threads())

    # -- dense helpers ----------------------------------------------------------------------
    def cor(self, x, y):
        x = np.ascontiguousarray(x, np.float64).ravel(order="F")
        y = np.ascontiguousarray(y, np.float64).ravel(order="F")
        return float(self._cor(_dp(x), _dp(y), x.size))

    def gram(self, X):
        """X: k x cols (any layout) -> k x k."""
        Xf = np.asfortranarray(X, np.float64)
        k, cols = Xf.shape
        a = np.zeros((k, k), np.float64, order="F")
        self._gram(_dp(Xf), k, cols, _dp(a))
        return a

    def scale(self, X):
        Xf = np.array(X, np.float64, order="F")
        k, cols = Xf.shape
        d = np.zeros(k, np.float64)
        self._scale(_dp(Xf), k, cols, _dp(d))
        return Xf, d

    def nnls(self, a, b, x, L1=0.0, L2=0.0):
        a = np.asfortranarray(a, np.float64)
        b = np.array(b, np.float64)
        x = np.array(x, np.float64)
        sweeps = self._nnls(_dp(a), _dp(b), _dp(x), a.shape[0], L1, L2)
        return x, b, int(sweeps)

    # -- kernels ----------------------------------------------------------------------------
    def predict(self, A, w, h, L1=0.0, L2=0.0, threads=0):
        arr, n, keep = _as_chunks(A)
        w = np.asfortranarray(w, np.float64)
        h = np.array(h, np.float64, order="F")
        sweeps = C.c_int64(0)
        self._predict(arr, n, _dp(w), w.shape[0], _dp(h), L1, L2, threads, C.addressof(sweeps))
        self.last_sweeps = sweeps.value
        return h

    def predict_mask(self, A, seed, inv_density, w, h, L1=0.0, L2=0.0, threads=0, mask_t=False):
        arr, n, keep = _as_chunks(A)
        w = np.asfortranarray(w, np.float64)
        h = np.array(h, np.float64, order="F")
        self._predict_mask(arr, n, seed, inv_density, _dp(w), w.shape[0], _dp(h), L1, L2, threads, int(mask_t))
        return h

    def mse_test(self, A, w, d, h, seed, inv_density, threads=0):
        arr, n, keep = _as_chunks(A)
        w = np.asfortranarray(w, np.float64)
        h = np.asfortranarray(h, np.float64)
        d = np.ascontiguousarray(d, np.float64)
        return float(self._mse_test(arr, n, _dp(w), _dp(d), _dp(h), w.shape[0], seed, inv_density, threads))

    def mse_train(self, A, w, d, h, seed=0, inv_density=0, threads=0):
        arr, n, keep = _as_chunks(A)
        w = np.asfortranarray(w, np.float64)
        h = np.asfortranarray(h, np.float64)
        d = np.ascontiguousarray(d, np.float64)
        return float(self._mse_train(arr, n, _dp(w), _dp(d), _dp(h), w.shape[0], seed, inv_density, threads))

    # -- drivers ----------------------------------------------------------------------------
    def nmf(self, A, At, w_init, tol=1e-4, maxit=100, L1=(0.01, 0.01), L2=(0.0, 0.0), threads=0):
        """c_nmf: returns dict(w k x m, d, h k x n, iter, tol trace). L1/L2 = (w, h) pairs."""
        a, na, k1 = _as_chunks(A)
        at, nat, k2 = _as_chunks(At)
        w = np.array(w_init, np.float64, order="F")
        k, m = w.shape
        n = sum(int(a[q].ncol) for q in range(na))
        d = np.zeros(k, np.float64)
        h = np.zeros((k, n), np.float64, order="F")
        trace = np.full(max(int(maxit), 1), np.nan)
        it = self._nmf(a, na, at, nat, tol, maxit, L1[0], L1[1], L2[0], L2[1], threads, k, _dp(w), _dp(d), _dp(h),
                       _dp(trace))
        return {"w": w, "d": d, "h": h, "iter": it, "tol": trace[: max(it, 0)]}

    def ard_nmf(self, A, At, w_init, seed, inv_density, tol=1e-4, maxit=100, L1=0.01, L2=0.0, threads=0,
                overfit_threshold=1e-4, trace_test_mse=5):
        a, na, k1 = _as_chunks(A)
        at, nat, k2 = _as_chunks(At)
        w = np.array(w_init, np.float64, order="F")
        k, m = w.shape
        n = sum(int(a[q].ncol) for q in range(na))
        d = np.zeros(k, np.float64)
        h = np.zeros((k, n), np.float64, order="F")
        cap = int(maxit) + 2
        mse, ft, so = np.zeros(cap), np.zeros(cap), np.zeros(cap)
        it = np.zeros(cap, np.int32)
        nt = C.c_int(0)
        last = self._ard_nmf(a, na, at, nat, tol, maxit, L1, L2, threads, k, _dp(w), _dp(d), _dp(h), seed, inv_density,
                             overfit_threshold, trace_test_mse, _dp(mse), _dp(it), _dp(ft), _dp(so), cap,
                             C.addressof(nt))
        q = nt.value
        return {"w": w, "d": d, "h": h, "test_mse": mse[:q].copy(), "iter": it[:q].copy(), "tol": ft[:q].copy(),
                "score_overfit": so[:q].copy(), "last_iter": last}

    @staticmethod
    def _full_csc(D):
        """Dense matrix -> CSC with EVERY entry stored (explicit zeros): the dense reference loops visit all rows."""
        import scipy.sparse as sp

        D = np.asfortranarray(D, np.float64)
        m, n = D.shape
        return sp.csc_matrix((D.ravel(order="F"), np.tile(np.arange(m, dtype=np.int32), n), np.arange(n + 1, dtype=np.int32) * m),
                             shape=(m, n))

    def nmf_dense(self, A, At, w_init, tol=1e-4, maxit=100, L1=(0.01, 0.01), L2=(0.0, 0.0), threads=0):
        """c_nmf_dense (src/singlet.cpp:1051-1054). Port: the sparse restatement on the fully stored matrix (same
        operations in the same order); reference: the reference's own dense code path."""
        if self.kind == "port":
            return self.nmf(self._full_csc(A), self._full_csc(At), w_init, tol, maxit, L1, L2, threads)
        A = np.asfortranarray(A, np.float64)
        At = np.asfortranarray(At, np.float64)
        w = np.array(w_init, np.float64, order="F")
        k, m = w.shape
        n = A.shape[1]
        d, h = np.zeros(k), np.zeros((k, n), order="F")
        self._nmf_dense(_dp(A), _dp(At), m, n, tol, maxit, L1[0], L1[1], L2[0], L2[1], threads, k, _dp(w), _dp(d), _dp(h))
        return {"w": w, "d": d, "h": h}

    def ard_nmf_dense(self, A, At, w_init, seed, inv_density, tol=1e-4, maxit=100, L1=0.01, L2=0.0, threads=0,
                      overfit_threshold=1e-4, trace_test_mse=5):
        """c_ard_nmf_dense (src/singlet.cpp:1357-1361)."""
        if self.kind == "port":
            return self.ard_nmf(self._full_csc(A), self._full_csc(At), w_init, seed, inv_density, tol, maxit, L1, L2, threads,
                                overfit_threshold, trace_test_mse)
        A = np.asfortranarray(A, np.float64)
        At = np.asfortranarray(At, np.float64)
        w = np.array(w_init, np.float64, order="F")
        k, m = w.shape
        n = A.shape[1]
        d, h = np.zeros(k), np.zeros((k, n), order="F")
        cap = int(maxit) + 2
        mse, it, nt = np.zeros(cap), np.zeros(cap, np.int32), C.c_int(0)
        self._ard_nmf_dense(_dp(A), _dp(At), m, n, tol, maxit, L1, L2, threads, k, _dp(w), _dp(d), _dp(h), seed, inv_density,
                            overfit_threshold, trace_test_mse, _dp(mse), _dp(it), cap, C.addressof(nt))
        return {"w": w, "d": d, "h": h, "test_mse": mse[: nt.value].copy(), "iter": it[: nt.value].copy()}

    def linked_nmf(self, A, At, w_init, link_h, link_w, tol=1e-4, maxit=100, L1=0.01, L2=0.0, threads=0):
        """c_linked_nmf: link_h / link_w are (rows x cols) matrices; a side is linked only when its matrix has one
        column per cell (link_h) / per gene (link_w), like the reference."""
        a, na, k1 = _as_chunks(A)
        at, nat, k2 = _as_chunks(At)
        w = np.array(w_init, np.float64, order="F")
        k, m = w.shape
        n = int(a[0].ncol)
        lh = np.asfortranarray(link_h, np.float64)
        lw = np.asfortranarray(link_w, np.float64)
        d = np.zeros(k, np.float64)
        h = np.zeros((k, n), np.float64, order="F")
        self._linked_nmf(a, at, tol, maxit, L1, L2, threads, k, _dp(w), _dp(d), _dp(h), _dp(lh), lh.shape[0], lh.shape[1],
                         _dp(lw), lw.shape[0], lw.shape[1])
        return {"w": w, "d": d, "h": h}

    def weight_by_split(self, A, split_by, n_groups):
        """Returns the re-weighted values (port only)."""
        a, na, keep = _as_chunks(A)
        x = np.array(A.data if hasattr(A, "data") else A[2], np.float64)
        sb = np.ascontiguousarray(split_by, np.int32)
        self._weight_by_split(a, _dp(x), _dp(sb), int(n_groups))
        return x

    def project_model(self, A, w, L1=0.01, L2=0.0, threads=0):
        a, na, keep = _as_chunks(A)
        w = np.array(w, np.float64, order="F")
        k = w.shape[0]
        n = sum(int(a[q].ncol) for q in range(na))
        h = np.zeros((k, n), np.float64, order="F")
        d = np.zeros(k, np.float64)
        self._project_model(a, na, _dp(w), k, L1, L2, threads, _dp(h), _dp(d))
        return {"h": h, "d": d}

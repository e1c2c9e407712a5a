"""Multi-GPU entry points of ``libsinglet_cuda.so`` (csrc/multi.cu) from Python.

* :class:`MultiGPU` -- all GPUs of THIS process (``sgl_multi_create`` -> ``ncclCommInitAll``, one host thread per device
  inside the library): ``c_nmf_sparse_list`` / ``c_ard_nmf_sparse_list`` with the reference's chunk lists
  (src/singlet.cpp:715-743, 1162-1234; R/cross_validate_nmf.R:37-50). This is the path an R session takes.
* :class:`RankComm` / :class:`RankFit` -- one rank of a one-process-per-GPU job (``torchrun``): the NCCL unique id is made on
  rank 0 by the library and handed round by the caller (``torch.distributed`` here -- plumbing only); every collective of
  the fit is issued by the C++ library on its own stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import _as_csc, _callbacks, _dp


def shard_bounds(total: int, world: int, rank: int):
    lib = _lib.load()
    lo, hi, per = C.c_int64(), C.c_int64(), C.c_int64()
    lib.sgl_shard_bounds(int(total), int(world), int(rank), C.byref(lo), C.byref(hi), C.byref(per))
    return lo.value, hi.value, per.value


class MultiGPU:
    """``sgl_multi``: n devices of this process."""

    def __init__(self, n_devices: int, devices=None):
        self.lib = _lib.load()
        self._m = C.c_void_p()
        dev = None
        if devices is not None:
            dev = (C.c_int * n_devices)(*[int(d) for d in devices])
        _lib.check(self.lib.sgl_multi_create(int(n_devices), dev, C.byref(self._m)))
        self.n = int(self.lib.sgl_multi_size(self._m))

    def set_precision(self, mode):
        code = {"mixed16": _lib.PRECISION_MIXED16, "fp32": _lib.PRECISION_FP32, "mixed16_always": _lib.PRECISION_MIXED16_ALWAYS}.get(mode, mode)
        _lib.check(self.lib.sgl_multi_set_precision(self._m, int(code)))

    def collectives(self):
        return [int(self.lib.sgl_comm_collectives(self.lib.sgl_multi_rank(self._m, r))) for r in range(self.n)]

    def c_nmf_sparse_list(self, A_, At_, tol, maxit, verbose, L1, L2, threads, w):
        """``c_nmf_sparse_list`` (src/singlet.cpp:715-743) over the devices; ``At_`` may be None (it is not needed)."""
        return self.c_nmf(A_, At_, tol, maxit, verbose, L1, L1, L2, L2, threads, w)

    def c_nmf(self, A, At, tol, maxit, verbose, L1_w, L1_h, L2_w, L2_h, threads, w):
        A = _as_csc(A if isinstance(A, (list, tuple)) else [A])
        a, na, keep = _lib.chunks_to_c(A)
        wk = np.array(w, dtype=np.float64, order="F")
        k, m = wk.shape
        n = sum(x.shape[1] for x in A)
        d, hh = np.zeros(k), np.zeros((k, n), order="F")
        iters, ftol = C.c_int32(0), C.c_double(0)
        cb, keep_cb = _callbacks(bool(verbose), False)
        _lib.check(self.lib.sgl_multi_nmf(self._m, a, na, None, 0, float(tol), int(maxit) & 0xFFFF, float(L1_w), float(L1_h), float(L2_w),
                                          float(L2_h), k, _dp(wk), _dp(d), _dp(hh), C.addressof(iters), C.addressof(ftol),
                                          C.addressof(cb) if cb is not None else None))
        return {"w": wk, "d": d, "h": hh, "iter": iters.value, "tol": ftol.value}

    def c_ard_nmf_sparse_list(self, A_, At_, tol, maxit, verbose, L1, L2, threads, w, rng_seed, inv_density, overfit_threshold,
                              trace_test_mse):
        """``c_ard_nmf_sparse_list`` (src/singlet.cpp:1162-1234): ``A_`` column chunks of A, ``At_`` gene blocks of t(A)."""
        A = _as_csc(A_ if isinstance(A_, (list, tuple)) else [A_])
        At = _as_csc(At_ if isinstance(At_, (list, tuple)) else [At_])
        a, na, k1 = _lib.chunks_to_c(A)
        at, nat, k2 = _lib.chunks_to_c(At)
        wk = np.array(w, dtype=np.float64, order="F")
        k, m = wk.shape
        n = sum(x.shape[1] for x in A)
        d, hh = np.zeros(k), np.zeros((k, n), order="F")
        cap = int(maxit) + 2
        mse, ft, so, it = np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap, np.int32)
        tr = _lib.Trace(mse.ctypes.data, it.ctypes.data, ft.ctypes.data, so.ctypes.data, cap, 0)
        cb, keep_cb = _callbacks(bool(verbose), True)
        _lib.check(self.lib.sgl_multi_ard_nmf(self._m, a, na, at, nat, float(tol), int(maxit) & 0xFFFF, float(L1), float(L2), k, _dp(wk), _dp(d),
                                              _dp(hh), int(rng_seed), int(inv_density), float(overfit_threshold), int(trace_test_mse) & 0xFFFF,
                                              C.addressof(tr), C.addressof(cb) if cb is not None else None))
        q = tr.length
        return {"w": wk, "d": d, "h": hh, "test_mse": mse[:q].copy(), "iter": it[:q].copy(), "tol": ft[:q].copy(), "score_overfit": so[:q].copy()}

    def close(self):
        if self._m:
            self.lib.sgl_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RankComm:
    """``sgl_comm`` of one process of a torchrun job. ``handle_ptr``: the ``sgl_handle`` whose stream carries the collectives."""

    def __init__(self, handle_ptr, device: int, world: int, rank: int, group=None):
        self.lib = _lib.load()
        ident = np.zeros(128, np.uint8)
        if world > 1:
            import torch
            import torch.distributed as dist

            if rank == 0:
                _lib.check(self.lib.sgl_comm_unique_id(ident.ctypes.data))
            dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.from_numpy(ident).to(dev)
            dist.broadcast(t, 0, group=group)
            ident = t.cpu().numpy()
        self._c = C.c_void_p()
        _lib.check(self.lib.sgl_comm_init_rank(handle_ptr, int(device), int(world), int(rank), ident.ctypes.data, C.byref(self._c)))
        self.rank, self.world = rank, world

    def collectives(self):
        return int(self.lib.sgl_comm_collectives(self._c))

    def close(self):
        if self._c:
            self.lib.sgl_comm_destroy(self._c)
            self._c = C.c_void_p()


class RankFit:
    """``sgl_fit``: the sharded fit of this rank (device matrices from ``sgl_matrix_upload`` / ``sgl_matrix_synth``)."""

    def __init__(self, comm: RankComm, A_loc, At_loc, n_total: int, k: int, w_init, masked=False, seed=0, inv_density=0):
        self.lib, self.comm, self.k = comm.lib, comm, k
        w = np.array(w_init, dtype=np.float64, order="F")
        self.m = w.shape[1]
        self._f = C.c_void_p()
        _lib.check(self.lib.sgl_fit_create(comm._c, A_loc, At_loc, int(n_total), int(k), w.ctypes.data, int(bool(masked)), int(seed),
                                           int(inv_density), C.byref(self._f)))
        c0, c1 = C.c_int64(), C.c_int64()
        self.lib.sgl_fit_shard(self._f, C.byref(c0), C.byref(c1), None, None)
        self.c0, self.c1 = c0.value, c1.value

    def iterate(self, L1_w, L1_h, L2_w, L2_h):
        tol = C.c_double(0)
        _lib.check(self.lib.sgl_fit_iterate(self._f, float(L1_w), float(L1_h), float(L2_w), float(L2_h), C.byref(tol), None))
        return tol.value

    def test_mse(self):
        out = C.c_double(0)
        _lib.check(self.lib.sgl_fit_test_mse(self._f, C.byref(out)))
        return out.value

    def download(self, want_h=True):
        w, d = np.zeros((self.k, self.m), order="F"), np.zeros(self.k)
        h = np.zeros((self.k, self.c1 - self.c0), order="F") if want_h else None
        _lib.check(self.lib.sgl_fit_download(self._f, w.ctypes.data, d.ctypes.data, h.ctypes.data if want_h and h.size else None))
        return w, d, h

    def close(self):
        if self._f:
            self.lib.sgl_fit_destroy(self._f)
            self._f = C.c_void_p()

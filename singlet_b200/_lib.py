"""ctypes binding of ``libsinglet_cuda.so`` (the C ABI in ``include/singlet_cuda.h``).

There is deliberately no fallback: if the shared library is missing or no sm_100 device is
present, every compute entry point raises :class:`SingletCudaError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsinglet_cuda.so")

SGL_OK, SGL_EINVAL, SGL_ENODEVICE, SGL_ECUDA, SGL_EINTERRUPT, SGL_ENOMEM = 0, -1, -2, -3, -4, -5
MAX_RANK = 128
PRECISION_MIXED16, PRECISION_FP32, PRECISION_MIXED16_ALWAYS = 0, 1, 2


class SingletCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsinglet_cuda error {code}: {msg}")
        self.code = code


class Csc(C.Structure):
    """``sgl_csc``: dgCMatrix slot view (reference inst/include/singlet.h:36-41)."""
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int64), ("p", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p)]


POLL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)
ITER_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_double, C.c_double)


class Callbacks(C.Structure):
    _fields_ = [("user", C.c_void_p), ("poll_interrupt", POLL_FN), ("on_iter", ITER_FN)]


class Trace(C.Structure):
    _fields_ = [("test_mse", C.c_void_p), ("iter", C.c_void_p), ("tol", C.c_void_p), ("score_overfit", C.c_void_p),
                ("capacity", C.c_int32), ("length", C.c_int32)]


class FitJob(C.Structure):
    """``sgl_fit_job``: one fit of a batched rank search (sgl_ard_nmf_batch)."""
    _fields_ = [("k", C.c_int32), ("status", C.c_int32), ("seed", C.c_uint64), ("w", C.c_void_p), ("d", C.c_void_p),
                ("h", C.c_void_p), ("trace", C.c_void_p)]


_lib = None

# every symbol include/singlet_cuda.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _u64, _dbl, _u16 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_uint16
SYMBOLS = [
    ("sgl_version", _i32, []),
    ("sgl_last_error", C.c_char_p, []),
    ("sgl_device_count", _i32, []),
    ("sgl_create", _i32, [_i32, _vp, C.POINTER(_vp)]),
    ("sgl_destroy", _i32, [_vp]),
    ("sgl_set_cache", _i32, [_vp, _i32]),
    ("sgl_set_precision", _i32, [_vp, _i32]),
    ("sgl_get_precision", _i32, [_vp]),
    ("sgl_synchronize", _i32, [_vp]),
    ("sgl_stream", _vp, [_vp]),
    ("sgl_launch_count", _i64, [_vp]),
    ("sgl_profile", _i32, [_vp, _i32]),
    ("sgl_profile_read", _i32, [_vp, _vp, _vp, _vp]),
    ("sgl_nmf", _i32, [_vp, _vp, _i32, _vp, _i32, _dbl, _u16, _dbl, _dbl, _dbl, _dbl, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("sgl_linked_nmf", _i32, [_vp, _vp, _vp, _dbl, _u16, _dbl, _dbl, _i32, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _i64, _vp,
                              _vp, _vp]),
    ("sgl_ard_nmf", _i32, [_vp, _vp, _i32, _vp, _i32, _dbl, _u16, _dbl, _dbl, _i32, _vp, _vp, _vp, _u64, _u64, _dbl, _u16,
                           _vp, _vp]),
    ("sgl_ard_nmf_batch", _i32, [_vp, _vp, _i32, _vp, _i32, _dbl, _u16, _dbl, _dbl, _u64, _dbl, _u16, _vp, _i32, _i32, _vp]),
    ("sgl_nmf_dense", _i32, [_vp, _vp, _vp, _i64, _i64, _dbl, _u16, _dbl, _dbl, _dbl, _dbl, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("sgl_ard_nmf_dense", _i32, [_vp, _vp, _vp, _i64, _i64, _dbl, _u16, _dbl, _dbl, _i32, _vp, _vp, _vp, _u64, _u64, _dbl, _u16,
                                 _vp, _vp]),
    ("sgl_project_model", _i32, [_vp, _vp, _i32, _vp, _i64, _i64, _dbl, _dbl, _vp, _vp]),
    ("sgl_predict", _i32, [_vp, _vp, _i32, _vp, _i64, _i64, _dbl, _dbl, _vp]),
    ("sgl_mask_rand", _i32, [_vp, _u64, _vp, _vp, _i64, _vp]),
    ("sgl_mask_draw", _i32, [_vp, _u64, _u64, _vp, _vp, _i64, _vp]),
    ("sgl_padded_rank", _i32, [_i32]),
    ("sgl_matrix_upload", _i32, [_vp, _vp, _i32, C.POINTER(_vp)]),
    ("sgl_matrix_transpose", _i32, [_vp, _vp, C.POINTER(_vp)]),
    ("sgl_matrix_synth", _i32, [_vp, _i64, _i64, _dbl, _u64, _i32, _i64, _i64, _vp, C.POINTER(_vp)]),
    ("sgl_matrix_synth_block", _i32, [_vp, _i64, _i64, _dbl, _u64, _i32, _i64, _i64, _i64, _i64, _vp, C.POINTER(_vp)]),
    ("sgl_matrix_colptr", _i32, [_vp, _vp, _vp]),
    ("sgl_matrix_free", _i32, [_vp, _vp]),
    ("sgl_matrix_info", _i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    ("sgl_matrix_download", _i32, [_vp, _vp, _vp, _vp, _vp]),
    ("sgl_factor_upload", _i32, [_vp, _vp, _i32, _i64, _vp]),
    ("sgl_factor_download", _i32, [_vp, _vp, _i32, _i64, _vp]),
    ("sgl_dev_gram", _i32, [_vp, _vp, _i32, _i64, _vp, _i32]),
    ("sgl_dev_gram_jitter", _i32, [_vp, _i32, _vp]),
    ("sgl_dev_update", _i32, [_vp, _vp, _vp, _vp, _i32, _vp, _dbl, _dbl, _vp]),
    ("sgl_dev_rhs", _i32, [_vp, _vp, _vp, _i32, _vp]),
    ("sgl_dev_solve", _i32, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _dbl, _dbl, _vp]),
    ("sgl_dev_update_rhs", _i32, [_vp, _vp, _vp, _i32, C.POINTER(_u64)]),
    ("sgl_dev_rhs_epoch", _u64, [_vp]),
    ("sgl_dev_update_solve", _i32, [_vp, _vp, _u64, _vp, _i32, _vp, _dbl, _dbl, _vp]),
    ("sgl_dev_finish_d", _i32, [_vp, _i32, _vp]),
    ("sgl_dev_finish_d_rescale_gram", _i32, [_vp, _i32, _vp, _vp]),
    ("sgl_dev_scale", _i32, [_vp, _vp, _i32, _i64, _vp]),
    ("sgl_dev_cor_sums", _i32, [_vp, _vp, _vp, _i32, _i64, _vp]),
    ("sgl_cor_from_sums", _dbl, [_vp, _dbl]),
    ("sgl_mask_build", _i32, [_vp, _vp, _u64, _u64, _i32, _i64, _i64, C.POINTER(_vp)]),
    ("sgl_mask_free", _i32, [_vp, _vp]),
    ("sgl_mask_info", _i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    ("sgl_mask_column", _i64, [_vp, _vp, _i64, _vp, _i64]),
    ("sgl_dev_update_masked", _i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _dbl, _dbl, _vp]),
    ("sgl_dev_mse", _i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    # IVSparse wire formats (csrc/ivsparse.cpp)
    ("sgl_ivsparse_info", _i32, [_vp, _u64, C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    ("sgl_ivsparse_decode", _i64, [_vp, _u64, _i64, _i64, _vp, _vp, _vp, _i64]),
    ("sgl_ivsparse_encode", _i64, [_vp, _i32, _i32, _vp, _u64]),
    # multi-GPU (csrc/multi.cu)
    ("sgl_shard_bounds", None, [_i64, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    ("sgl_comm_unique_id", _i32, [_vp]),
    ("sgl_comm_init_rank", _i32, [_vp, _i32, _i32, _i32, _vp, C.POINTER(_vp)]),
    ("sgl_comm_destroy", _i32, [_vp]),
    ("sgl_comm_rank", _i32, [_vp]),
    ("sgl_comm_world", _i32, [_vp]),
    ("sgl_comm_handle", _vp, [_vp]),
    ("sgl_comm_collectives", _i64, [_vp]),
    ("sgl_fit_create", _i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _u64, _u64, C.POINTER(_vp)]),
    ("sgl_fit_iterate", _i32, [_vp, _dbl, _dbl, _dbl, _dbl, C.POINTER(_dbl), C.POINTER(_i32)]),
    ("sgl_fit_set_lookahead", _i32, [_vp, _i32]),
    ("sgl_fit_test_mse", _i32, [_vp, C.POINTER(_dbl)]),
    ("sgl_fit_download", _i32, [_vp, _vp, _vp, _vp]),
    ("sgl_fit_shard", _i32, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    ("sgl_fit_destroy", _i32, [_vp]),
    ("sgl_nmf_rank", _i32, [_vp, _vp, _vp, _i64, _dbl, _u16, _dbl, _dbl, _dbl, _dbl, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("sgl_ard_nmf_rank", _i32, [_vp, _vp, _vp, _i64, _dbl, _u16, _dbl, _dbl, _i32, _vp, _vp, _vp, _u64, _u64, _dbl, _u16, _vp, _vp]),
    ("sgl_multi_create", _i32, [_i32, _vp, C.POINTER(_vp)]),
    ("sgl_multi_destroy", _i32, [_vp]),
    ("sgl_multi_size", _i32, [_vp]),
    ("sgl_multi_rank", _vp, [_vp, _i32]),
    ("sgl_multi_set_precision", _i32, [_vp, _i32]),
    ("sgl_multi_nmf", _i32, [_vp, _vp, _i32, _vp, _i32, _dbl, _u16, _dbl, _dbl, _dbl, _dbl, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("sgl_multi_ard_nmf", _i32, [_vp, _vp, _i32, _vp, _i32, _dbl, _u16, _dbl, _dbl, _i32, _vp, _vp, _vp, _u64, _u64, _dbl, _u16,
                                 _vp, _vp]),
]


def _point_at_bundled_nccl():
    """The multi-GPU entry points bind NCCL at run time (csrc/multi.cu). In a Python process that also imports torch, both
    must use the SAME libnccl.so.2 (the loader keys it by soname): point the library at the copy bundled with torch
    (site-packages/nvidia/nccl/lib) unless the caller chose one with SGL_NCCL_LIB. Never imports torch."""
    if os.environ.get("SGL_NCCL_LIB"):
        return
    import sys

    for base in sys.path:
        cand = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
        if base and os.path.exists(cand):
            os.environ["SGL_NCCL_LIB"] = cand
            return


def load():
    """Load the shared library (once) and bind every declared symbol. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SingletCudaError(SGL_ENODEVICE, f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
    _point_at_bundled_nccl()
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int):
    if rc != SGL_OK:
        raise SingletCudaError(rc, load().sgl_last_error().decode("utf-8", "replace"))


def chunks_to_c(mats):
    """One CSC matrix or a list of column chunks -> (``sgl_csc`` array, count, keep-alive list).

    Accepts scipy.sparse CSC matrices or ``(p, i, x, nrow, ncol)`` tuples.
    """
    if not isinstance(mats, (list, tuple)) or (len(mats) == 5 and np.isscalar(mats[3])):
        mats = [mats]
    arr, keep = (Csc * len(mats))(), []
    for q, m in enumerate(mats):
        if hasattr(m, "indptr"):
            if getattr(m, "format", "csc") != "csc":
                m = m.tocsc()
            if not m.has_sorted_indices:
                m = m.sorted_indices()
            p, i, x, nrow, ncol = m.indptr, m.indices, m.data, m.shape[0], m.shape[1]
        else:
            p, i, x, nrow, ncol = m
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        keep += [p, i, x]
        arr[q] = Csc(int(nrow), int(ncol), p.ctypes.data, i.ctypes.data, x.ctypes.data)
    return arr, len(mats), keep

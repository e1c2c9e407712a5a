"""Multi-GPU behind the C ABI (csrc/multi.cu): the single-process entry points an R session would call. They run on however
many devices the box has -- with one device the same code path runs with a world of 1 (no NCCL call), with two or more the
collectives are NCCL's, issued by the library."""
import numpy as np
import pytest

from conftest import match_factors, min_factor_cor

pytestmark = pytest.mark.gpu


def _n_devices():
    from singlet_b200 import _lib

    return int(_lib.load().sgl_device_count())


def _chunks(A, parts):
    n = A.shape[1]
    cuts = [round(q * n / parts) for q in range(parts + 1)]
    return [A[:, cuts[q]:cuts[q + 1]].tocsc() for q in range(parts)]


@pytest.mark.parametrize("world", [1, 2, 4])
def test_multi_nmf_equals_single_gpu_fit(handle, oracle, world):
    """sgl_multi_nmf on `world` devices with an UNEVEN three-chunk list == sgl_nmf on one device == the oracle."""
    from singlet_b200 import api, synth
    from singlet_b200.multi import MultiGPU

    if world > _n_devices():
        pytest.skip(f"needs {world} GPUs")
    m, n, k = 1300, 1111, 12
    A = synth.synth_scipy(m, n, 0.06, seed=21)
    w0 = synth.w_init(k, m, seed=5)
    one = api.c_nmf(A, None, 0.0, 6, False, 0.01, 0.02, 0.0, 0.0, 0, w0)
    mg = MultiGPU(world)
    try:
        lst = [A[:, :100].tocsc(), A[:, 100:777].tocsc(), A[:, 777:].tocsc()]
        dev = mg.c_nmf(lst, None, 0.0, 6, False, 0.01, 0.02, 0.0, 0.0, 0, w0)
        assert dev["iter"] == one["iter"] == 6
        if world == 1:
            assert np.array_equal(dev["w"], one["w"]) and np.array_equal(dev["h"], one["h"]) and np.array_equal(dev["d"], one["d"])
        else:
            assert sum(mg.collectives()) > 0
            # FP32 sums are formed in another order (partial Grams / right-hand sides per rank): equal to ~1e-4 of the largest entry
            for nm in ("w", "h"):
                assert np.abs(dev[nm] - one[nm]).max() <= 2e-4 * np.abs(one[nm]).max(), nm
            assert np.allclose(dev["d"], one["d"], rtol=1e-5)
    finally:
        mg.close()
    At = A.T.tocsc()
    At.sort_indices()
    ref = oracle.nmf(A, At, w0, tol=0.0, maxit=6, L1=(0.01, 0.02), L2=(0.0, 0.0))
    perm = match_factors(ref["w"], dev["w"])
    assert min_factor_cor(ref["w"], dev["w"], perm) >= 0.999 and min_factor_cor(ref["h"], dev["h"], perm) >= 0.999
    assert np.allclose(dev["d"][perm], ref["d"], rtol=1e-3)


@pytest.mark.parametrize("world", [1, 2])
def test_multi_ard_nmf_equals_single_gpu_fit(handle, oracle, world):
    """sgl_multi_ard_nmf (chunked A_ and gene-block At_ lists, global mask indices) == sgl_ard_nmf == the oracle."""
    from singlet_b200 import api, synth
    from singlet_b200.multi import MultiGPU

    if world > _n_devices():
        pytest.skip(f"needs {world} GPUs")
    m, n, k = 700, 500, 9
    A = synth.synth_scipy(m, n, 0.07, seed=33)
    At = A.T.tocsc()
    At.sort_indices()
    w0 = synth.w_init(k, m, seed=6)
    one = api.c_ard_nmf(A, At, 0.0, 5, False, 0.01, 0.0, 0, w0, 999, 20, 10.0, 2)
    mg = MultiGPU(world)
    try:
        dev = mg.c_ard_nmf_sparse_list(_chunks(A, 3), _chunks(At, 2), 0.0, 5, False, 0.01, 0.0, 0, w0, 999, 20, 10.0, 2)
    finally:
        mg.close()
    assert list(dev["iter"]) == list(one["iter"])
    assert np.allclose(dev["test_mse"], one["test_mse"], rtol=1e-6)
    for nm in ("w", "h"):
        assert np.abs(dev[nm] - one[nm]).max() <= 2e-4 * np.abs(one[nm]).max(), nm
    ref = oracle.ard_nmf(A, At, w0, 999, 20, tol=0.0, maxit=5, L1=0.01, L2=0.0, overfit_threshold=10.0, trace_test_mse=2)
    assert list(dev["iter"]) == list(ref["iter"]) and np.allclose(dev["test_mse"], ref["test_mse"], rtol=1e-4)


def test_multi_interrupt_and_callbacks(handle):
    """Callbacks run on the calling thread; an interrupt stops every rank at the same iteration (no hang)."""
    import ctypes as C

    from singlet_b200 import _lib, synth
    from singlet_b200.multi import MultiGPU

    world = min(2, _n_devices())
    A = synth.synth_scipy(400, 300, 0.07, seed=4)
    w0 = np.array(synth.w_init(5, 400, seed=2), order="F")
    seen, polls = [], [0]
    import threading
    main = threading.get_ident()

    def poll(_u):
        assert threading.get_ident() == main
        polls[0] += 1
        return 1 if len(seen) >= 3 else 0

    def on_iter(_u, it, tol, _o):
        assert threading.get_ident() == main
        seen.append(it)

    cb = _lib.Callbacks(None, _lib.POLL_FN(poll), _lib.ITER_FN(on_iter))
    a, na, keep = _lib.chunks_to_c([A])
    d, h = np.zeros(5), np.zeros((5, 300), order="F")
    mg = MultiGPU(world)
    try:
        rc = mg.lib.sgl_multi_nmf(mg._m, a, na, None, 0, 0.0, 2000, 0.01, 0.01, 0.0, 0.0, 5, w0.ctypes.data, d.ctypes.data, h.ctypes.data, None, None,
                                  C.addressof(cb))
    finally:
        mg.close()
    assert rc == _lib.SGL_EINTERRUPT and seen[:3] == [1, 2, 3] and len(seen) < 2000

"""Device-level API + the sharded driver at world size 1 (N > 1 is covered on CPU with gloo in
test_sharded_gloo.py and on GPUs by bench.py under torchrun)."""
import numpy as np
import pytest

from conftest import match_factors, min_factor_cor

pytestmark = pytest.mark.gpu


def test_sharded_driver_world1_matches_c_driver(oracle):
    from singlet_b200 import api, synth
    from singlet_b200.sharded import CudaBackend, sharded_ard_nmf, sharded_nmf

    m, n, k = 800, 500, 10
    A = synth.synth_scipy(m, n, 0.06, seed=8)
    At = A.T.tocsc()
    At.sort_indices()
    w0 = synth.w_init(k, m, seed=3)
    be = CudaBackend(0)
    try:
        hA, hAt = be.upload(A), be.upload(At)
        s = sharded_nmf(be, m, n, k, hA, hAt, w0, tol=0.0, maxit=6, L1=(0.01, 0.01))
        c = api.c_nmf(A, At, 0.0, 6, False, 0.01, 0.01, 0, 0, 0, w0)
        assert np.array_equal(s["w"], c["w"]) and np.array_equal(s["h"], c["h"]) and np.allclose(s["d"], c["d"], rtol=1e-12)
        sb = sharded_nmf(be, m, n, k, hA, hAt, w0, tol=0.0, maxit=6, L1=(0.01, 0.01), layout="B")  # world 1: At is the full transpose
        assert np.array_equal(sb["w"], c["w"]) and np.array_equal(sb["h"], c["h"])
        sn = sharded_nmf(be, m, n, k, hA, None, w0, tol=0.0, maxit=6, L1=(0.01, 0.01), layout="B")  # transpose built on the device
        assert np.array_equal(sn["w"], c["w"]) and np.array_equal(sn["h"], c["h"])
        # layout B3 = the three-exchange iteration of csrc/multi.cu: the Gram of H is the Gram of the unscaled H divided by
        # d_i d_j (sgl_dev_finish_d_rescale_gram) instead of the Gram of the rounded scaled floats -- a rounding-level difference
        s3 = sharded_nmf(be, m, n, k, hA, None, w0, tol=0.0, maxit=6, L1=(0.01, 0.01), layout="B3")
        for nm in ("w", "h"):
            assert np.abs(s3[nm] - c[nm]).max() <= 2e-4 * np.abs(c[nm]).max(), nm
        assert np.allclose(s3["d"], c["d"], rtol=1e-5)
        sm = sharded_ard_nmf(be, m, n, k, hA, hAt, w0, 123, 20, tol=0.0, maxit=5, overfit_threshold=10.0, trace_test_mse=2)
        cm = api.c_ard_nmf(A, At, 0.0, 5, False, 0.01, 0, 0, w0, 123, 20, 10.0, 2)
        assert np.array_equal(sm["iter"], cm["iter"]) and np.allclose(sm["test_mse"], cm["test_mse"], rtol=1e-12)
        assert np.array_equal(sm["h"], cm["h"])
    finally:
        be.close()


def test_two_virtual_shards_on_one_gpu(oracle):
    """Run the two-rank partition sequentially on one device (collectives emulated by hand): the
    shard kernels with global offsets reproduce the unsharded fit."""
    from singlet_b200 import api, synth
    from singlet_b200.sharded import CudaBackend, shard_bounds

    m, n, k = 640, 450, 8
    A = synth.synth_scipy(m, n, 0.06, seed=18)
    At = A.T.tocsc()
    At.sort_indices()
    be = CudaBackend(0)
    try:
        full = be.upload(A)
        kp = be.kp(k)
        w0 = synth.w_init(k, m, seed=3)
        W = be.zeros_factor(m, k)
        be.factor_from_host(w0, W)
        gram = be.zeros_f64(kp * kp)
        be.gram(W, k, m, gram, True)
        Hfull = be.zeros_factor(n, k)
        rs_full = be.zeros_f64(kp)
        be.update(full, None, W, Hfull, k, gram, 0.01, 0.0, rs_full)
        Hparts, rs = be.zeros_factor(n, k), be.zeros_f64(kp)
        tot = np.zeros(kp)
        for r in range(2):
            lo, hi, per = shard_bounds(n, 2, r)
            sh = be.upload(A[:, lo:hi].tocsc())
            be.update(sh, None, W, Hparts[lo:hi], k, gram, 0.01, 0.0, rs)
            tot += rs.cpu().numpy()
        assert np.array_equal(Hparts.cpu().numpy(), Hfull.cpu().numpy())
        assert np.allclose(tot, rs_full.cpu().numpy(), rtol=1e-12)
        # layout B pieces: partial right-hand sides of the two cell blocks add up to the full ones, and the
        # device block generator equals the corresponding slice of the transpose
        from singlet_b200 import _lib as L
        At_full = be.upload(At)
        Bfull, Bpart, Bsum = be.zeros_factor(m, k), be.zeros_factor(m, k), be.zeros_factor(m, k)
        be.rhs(At_full, Hfull, k, Bfull)
        for r in range(2):
            lo, hi, per = shard_bounds(n, 2, r)
            blk = be.upload(A[:, lo:hi].T.tocsc())
            be.rhs(blk, Hfull[lo:hi], k, Bpart)
            Bsum += Bpart
        assert np.allclose(Bsum.cpu().numpy(), Bfull.cpu().numpy(), rtol=2e-5, atol=1e-6)
        tab = synth.values_table(m, 0.06)
        S = synth.synth_scipy(m, n, 0.06, seed=18)
        lo, hi, per = shard_bounds(n, 2, 1)
        dev_blk = be.synth_block(m, n, 0.06, 18, 1, 0, m, lo, hi - lo, tab)
        p, i, x, nrow, ncol = be.matrix_to_host(dev_blk)
        T = S[:, lo:hi].T.tocsc()
        T.sort_indices()
        assert (nrow, ncol) == T.shape and np.array_equal(p, T.indptr) and np.array_equal(i, T.indices) and np.array_equal(x, T.data)
        dev_blk0 = be.synth_block(m, n, 0.06, 18, 0, 5, 40, 100, 300, tab)
        p, i, x, nrow, ncol = be.matrix_to_host(dev_blk0)
        T0 = S[100:400, 5:45].tocsc()
        T0.sort_indices()
        assert np.array_equal(p, T0.indptr) and np.array_equal(i, T0.indices) and np.array_equal(x, T0.data)
        # masked shards hash with global column offsets
        mfull = be.mask_build(full, 123, 20, 0, 0, 0)
        lo, hi, per = shard_bounds(n, 2, 1)
        sh = be.upload(A[:, lo:hi].tocsc())
        msh = be.mask_build(sh, 123, 20, 0, lo, 0)
        buf1, buf2 = np.zeros(m, np.int32), np.zeros(m, np.int32)
        for c in range(0, hi - lo, 11):
            c1 = be.lib.sgl_mask_column(be._h, msh, c, buf1.ctypes.data, m)
            c2 = be.lib.sgl_mask_column(be._h, mfull, lo + c, buf2.ctypes.data, m)
            assert c1 == c2 and np.array_equal(buf1[:c1], buf2[:c2])
    finally:
        be.close()

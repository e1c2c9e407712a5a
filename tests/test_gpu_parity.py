"""Parity of the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): mask / held-out index selection bit-exact; after matching
factors by cosine, per-factor correlation >= 0.999; train/test MSE within 1e-4 relative.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import match_factors, min_factor_cor

pytestmark = pytest.mark.gpu

COR_MIN = 0.999
MSE_RTOL = 1e-4


def _mk(m, n, density, seed, empty_cols=()):
    from singlet_b200 import synth

    A = synth.synth_scipy(m, n, density, seed=seed).tolil()
    for c in empty_cols:
        A[:, c] = 0
    A = A.tocsc()
    A.eliminate_zeros()
    A.sort_indices()
    At = A.T.tocsc()
    At.sort_indices()
    return A, At


def test_mask_hash_bit_exact(handle, oracle):
    """rng::rand(i,j) and rng::draw on the device == reference hash (src/singlet.cpp:47-64, 91-95)."""
    from singlet_b200 import _lib

    rs = np.random.RandomState(0)
    n = 200000
    i = rs.randint(0, 2**40, size=n).astype(np.uint64)
    j = rs.randint(0, 2**33, size=n).astype(np.uint64)
    i[:6] = [0, 1, 7, 2699, 999999, 2**63 + 5]
    j[:6] = [0, 0, 13, 13713, 29999, 2**62 + 1]
    for seed in (0, 1, 123, 999, 2147483647, 1234567890123, 2**64 - 1):
        out = np.zeros(n, np.uint64)
        _lib.check(handle.lib.sgl_mask_rand(handle.ptr, seed, i.ctypes.data, j.ctypes.data, n, out.ctypes.data))
        exp = np.array([oracle.rand2(seed, int(a), int(b)) for a, b in zip(i[:3000], j[:3000])], dtype=np.uint64)
        assert np.array_equal(out[:3000], exp)
        for inv in (1, 2, 20, 33, 65535, 65536, 10**9 + 7, 2**40 + 3):
            d = np.zeros(n, np.uint8)
            _lib.check(handle.lib.sgl_mask_draw(handle.ptr, seed, inv, i.ctypes.data, j.ctypes.data, n, d.ctypes.data))
            assert np.array_equal(d.astype(bool), (out % np.uint64(inv)) == 0)


def test_rng_known_answers_on_device(handle):
    """SURVEY.md App. B.1 known-answer vectors (generated from the verbatim reference class)."""
    from singlet_b200 import _lib

    kat = {
        (0, 0, 1): 2459528346506729745, (1, 1, 0): 6295464863713931364, (123, 7, 13): 9538636253944861798,
        (999, 2699, 13713): 12937052898447939927, (2147483647, 999999, 29999): 7545449070027258386,
        (1234567890123, 0, 0): 15671908682223407781,
    }
    for (seed, i, j), exp in kat.items():
        ii, jj, out = np.array([i], np.uint64), np.array([j], np.uint64), np.zeros(1, np.uint64)
        _lib.check(handle.lib.sgl_mask_rand(handle.ptr, seed, ii.ctypes.data, jj.ctypes.data, 1, out.ctypes.data))
        assert int(out[0]) == exp


def test_synth_device_matches_host():
    """The device generator and the numpy restatement produce identical matrices (both orientations)."""
    from singlet_b200 import synth
    from singlet_b200.sharded import CudaBackend

    be = CudaBackend(0)
    try:
        m, n, dens = 1234, 777, 0.05
        A = synth.synth_scipy(m, n, dens)
        tab = synth.values_table(m, dens)
        h = be.synth(m, n, dens, synth.DATA_SEED, 0, 100, 500, tab)
        p, i, x, nrow, ncol = be.matrix_to_host(h)
        S = A[:, 100:600]
        assert (nrow, ncol) == (m, 500)
        assert np.array_equal(p, S.indptr) and np.array_equal(i, S.indices) and np.array_equal(x, S.data)
        ht = be.synth(m, n, dens, synth.DATA_SEED, 1, 200, 900, tab)
        p, i, x, nrow, ncol = be.matrix_to_host(ht)
        T = A.T.tocsc()
        T.sort_indices()
        T = T[:, 200:1100]
        assert (nrow, ncol) == (n, 900)
        assert np.array_equal(p, T.indptr) and np.array_equal(i, T.indices) and np.array_equal(x, T.data)
    finally:
        be.close()


def test_upload_cache_sees_in_place_edits(oracle):
    """The upload cache (one handle per R session keeps A / At resident between calls) is keyed by a hash of EVERY byte of the
    host p / i / x: a repeated call on the unchanged matrix reuses the device copy, an in-place edit of ONE value that the cheap
    fingerprint does not sample (the R idiom `A@x[j] <- v`) re-uploads it."""
    from singlet_b200 import api, synth

    m, n, k = 400, 9000, 6
    A = synth.synth_scipy(m, n, 0.05, seed=31)
    assert A.nnz > 3 * 4096  # the fingerprint samples every (nnz / 4096)-th value: entry 1 is not one of them
    w0 = synth.w_init(k, m, seed=32)
    h = api.Handle(0)
    try:
        first = api.c_nmf(A, None, 0.0, 4, False, 0.01, 0.01, 0, 0, 0, w0, h)
        again = api.c_nmf(A, None, 0.0, 4, False, 0.01, 0.01, 0, 0, 0, w0, h)
        assert np.array_equal(first["h"], again["h"]) and np.array_equal(first["w"], again["w"])
        col = int(np.searchsorted(A.indptr, 1, side="right") - 1)  # the column that holds entry 1
        A.data[1] *= 50.0
        edited = api.c_nmf(A, None, 0.0, 4, False, 0.01, 0.01, 0, 0, 0, w0, h)
        assert not np.array_equal(first["h"][:, col], edited["h"][:, col])
        fresh = api.Handle(0)
        try:
            fresh.set_cache(False)
            ref = api.c_nmf(A, None, 0.0, 4, False, 0.01, 0.01, 0, 0, 0, w0, fresh)
        finally:
            fresh.close()
        assert np.array_equal(edited["h"], ref["h"]) and np.array_equal(edited["w"], ref["w"])
        # the same for the masked path (the masks hold copies of the values) and for the batched rank search (its workers
        # share the cached matrices)
        args = (0.0, 3, False, 0.01, 0.0, 0, w0, 55, 12, 10.0, 2)
        before = api.c_ard_nmf(A, None, *args, h)
        A.data[1] /= 50.0
        after = api.c_ard_nmf(A, None, *args, h)
        assert not np.array_equal(before["test_mse"], after["test_mse"]) or not np.array_equal(before["h"], after["h"])
        batch = api.c_ard_nmf_batch(A, None, 0.0, 3, 0.01, 0.0, 0, [w0, w0[:4]], [55, 56], 12, 10.0, 2, handle=h)
        fresh = api.Handle(0)
        try:
            fresh.set_cache(False)
            ref_m = api.c_ard_nmf(A, None, *args, fresh)
            ref_b = api.c_ard_nmf(A, None, 0.0, 3, False, 0.01, 0.0, 0, w0[:4], 56, 12, 10.0, 2, fresh)
        finally:
            fresh.close()
        for key in ("w", "d", "h", "test_mse"):
            assert np.array_equal(after[key], ref_m[key]) and np.array_equal(batch[0][key], ref_m[key]) and np.array_equal(batch[1][key], ref_b[key]), key
    finally:
        h.close()


def test_handle_state_does_not_leak_between_calls():
    """One handle per R session serves every call: cached uploads, transposes, masks (refilled in place when the seed changes),
    tile indices per padded rank and operand format, grow-only factor buffers, the FP16 shadow. A call must give the same
    result whatever ran before it on the handle."""
    from singlet_b200 import api, synth

    A1 = synth.synth_scipy(700, 500, 0.08, seed=61)
    A2 = synth.synth_scipy(350, 900, 0.05, seed=62)
    h = api.Handle(0)

    def plain(A, k):
        return api.c_nmf(A, None, 0.0, 4, False, 0.01, 0.01, 0, 0, 0, synth.w_init(k, A.shape[0], seed=k), h)

    def masked(A, k, seed):
        return api.c_ard_nmf(A, None, 0.0, 4, False, 0.01, 0.0, 0, synth.w_init(k, A.shape[0], seed=k), seed, 15, 10.0, 2, h)

    def same(a, b, keys=("w", "d", "h")):
        return all(np.array_equal(a[q], b[q]) for q in keys)

    try:
        base = {("p", 1, 5): plain(A1, 5), ("p", 1, 20): plain(A1, 20), ("p", 1, 40): plain(A1, 40), ("p", 2, 9): plain(A2, 9)}
        m1 = masked(A1, 20, 77)
        m2 = masked(A1, 20, 78)       # another seed: the mask is refilled in place
        assert not np.array_equal(m1["test_mse"], m2["test_mse"])
        assert same(masked(A1, 20, 77), m1, ("w", "d", "h", "test_mse"))
        m3 = masked(A2, 6, 77)        # another matrix: upload, transpose and masks are replaced
        assert same(plain(A1, 5), base[("p", 1, 5)])
        assert same(masked(A2, 6, 77), m3, ("w", "d", "h", "test_mse"))
        assert same(plain(A1, 40), base[("p", 1, 40)]) and same(plain(A2, 9), base[("p", 2, 9)]) and same(plain(A1, 20), base[("p", 1, 20)])
        # operand precision switched back and forth on the same handle (different stream formats, the FP16 shadow, the
        # single-pass Gram correction)
        h.set_precision("mixed16_always")
        p16 = plain(A1, 40)
        q16 = masked(A1, 20, 77)
        h.set_precision("fp32")
        assert same(plain(A1, 40), base[("p", 1, 40)])
        h.set_precision("mixed16_always")
        assert same(plain(A1, 40), p16) and same(masked(A1, 20, 77), q16, ("w", "d", "h", "test_mse"))
        h.set_precision("mixed16")
        assert same(masked(A1, 20, 77), m1, ("w", "d", "h", "test_mse"))
        # projection with a model of another rank in between
        pr = api.c_project_model(A1, base[("p", 1, 5)]["w"], 0.01, 0.0, 0, h)
        plain(A2, 9)
        pr2 = api.c_project_model(A1, base[("p", 1, 5)]["w"], 0.01, 0.0, 0, h)
        assert np.array_equal(pr["h"], pr2["h"]) and np.array_equal(pr["d"], pr2["d"])
    finally:
        h.close()


def test_large_factor_download_paths_agree(monkeypatch):
    """A factor of >= 4 M entries comes back through the upload workers' pinned staging buffers (FP32 over PCIe, widened by
    host threads); with a single worker the library takes the plain path (one device-side conversion + one copy). Same bytes."""
    from singlet_b200 import api, synth

    A = synth.synth_scipy(1500, 250000, 0.01, seed=5)
    w0 = synth.w_init(20, 1500, seed=6)   # padded rank 32 != 20: the staged rows carry padding that must be dropped
    monkeypatch.setenv("SGL_UPLOAD_THREADS", "1")
    h1 = api.Handle(0)
    try:
        a = api.c_nmf(A, None, 0.0, 2, False, 0.01, 0.01, 0, 0, 0, w0, h1)
    finally:
        h1.close()
    monkeypatch.delenv("SGL_UPLOAD_THREADS")
    h2 = api.Handle(0)
    try:
        b = api.c_nmf(A, None, 0.0, 2, False, 0.01, 0.01, 0, 0, 0, w0, h2)
    finally:
        h2.close()
    assert a["h"].shape == (20, 250000) and np.array_equal(a["h"], b["h"]) and np.array_equal(a["w"], b["w"]) and np.array_equal(a["d"], b["d"])


def test_plain_solver_tail_split(handle, oracle, monkeypatch):
    """Column counts a little above a whole number of rounds of the thread-per-column solver's grid (768 columns per SM): the
    full rounds are solved by that kernel and the short remainder by the sub-warp kernel (engine.cu, "tail split"). Both follow
    the reference's coordinate order, so the split result equals the unsplit one to FP32 rounding and both match the oracle on a
    sample of columns (reference src/singlet.cpp:229-250, 333-347)."""
    import torch

    from singlet_b200 import api, synth

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    m, n, k = 96, sms * 768 + sms * 60, 20  # remainder: 60 columns per SM (the split takes up to 96)
    A = synth.synth_scipy(m, n, 0.2, seed=91)
    w = synth.w_init(k, m, seed=92)
    monkeypatch.delenv("SGL_NNLS_NO_TAIL_SPLIT", raising=False)
    split = api.Rcpp_predict(A, w, 0.01, 0.0, 0)
    monkeypatch.setenv("SGL_NNLS_NO_TAIL_SPLIT", "1")
    whole = api.Rcpp_predict(A, w, 0.01, 0.0, 0)
    monkeypatch.delenv("SGL_NNLS_NO_TAIL_SPLIT")
    scale = np.abs(whole).max()
    assert np.array_equal(split[:, :sms * 768], whole[:, :sms * 768])          # the full rounds are the same kernel
    assert np.abs(split - whole).max() <= 2e-5 * scale                          # the remainder: another kernel, same algorithm
    cols = np.r_[0:40, sms * 768 - 20:sms * 768 + 40, n - 40:n]
    ref = oracle.predict(A[:, cols].tocsc(), w, np.zeros((k, len(cols))), 0.01, 0.0)
    assert np.abs(split[:, cols] - ref).max() <= 2e-4 * np.abs(ref).max()


@pytest.fixture
def fp32_operands(handle):
    """Run a test with FP32 gather operands (sgl_set_precision); the default 16-bit staging is restored afterwards."""
    handle.set_precision("fp32")
    yield handle
    handle.set_precision("mixed16")


@pytest.mark.parametrize("k", [1, 3, 8, 10, 20, 32, 40, 64, 100])
def test_predict_matches_oracle(fp32_operands, oracle, k):
    """One H update (Rcpp_predict, src/singlet.cpp:350-367) for every padded-rank code path, with
    empty columns, L1 and L2. FP32 operands: the solver is compared tightly (a random uniform w gives an
    ill-conditioned Gram that amplifies any right-hand-side rounding)."""
    from singlet_b200 import api, synth

    m, n = 700, 450
    A, At = _mk(m, n, 0.08, seed=k, empty_cols=(0, 17, n - 1))
    w = synth.w_init(k, m, seed=k + 1)
    for L1, L2 in ((0.0, 0.0), (0.01, 0.0), (0.05, 0.1)):
        dev = api.Rcpp_predict(A, w, L1, L2, 0)
        ref = oracle.predict(A, w, np.zeros((k, n)), L1, L2)
        assert np.all(dev[:, [0, 17, n - 1]] == 0)
        scale = np.abs(ref).max()
        assert np.abs(dev - ref).max() <= 2e-4 * scale, (k, L1, L2, np.abs(dev - ref).max() / scale)


@pytest.mark.parametrize("k", [20, 32, 40, 64, 100])
def test_predict_mixed16_matches_oracle(handle, oracle, k):
    """The same H update with the default 16-bit staged operands (padded ranks >= 32: FP16 shadow of w, FP16 values,
    FP32 accumulation). The right-hand sides carry zero-mean operand rounding of 2^-12 per element; through the
    ill-conditioned Gram of a random uniform w that shows up as <= 2e-3 of the largest entry, and every column of h
    correlates >= 0.9999 with the reference (north_star tolerance: 0.999)."""
    from singlet_b200 import api, synth

    m, n = 700, 450
    A, At = _mk(m, n, 0.08, seed=k, empty_cols=(0, 17, n - 1))
    w = synth.w_init(k, m, seed=k + 1)
    handle.set_precision("mixed16_always")  # this matrix is far below the size at which the default mode stages in 16 bits
    try:
        for L1, L2 in ((0.0, 0.0), (0.01, 0.0)):
            dev = api.Rcpp_predict(A, w, L1, L2, 0)
            ref = oracle.predict(A, w, np.zeros((k, n)), L1, L2)
            assert np.all(dev[:, [0, 17, n - 1]] == 0)
            scale = np.abs(ref).max()
            assert np.abs(dev - ref).max() <= 2e-3 * scale, (k, L1, L2, np.abs(dev - ref).max() / scale)
            cols = [c for c in range(n) if ref[:, c].std() > 0]
            cors = [np.corrcoef(dev[:, c], ref[:, c])[0, 1] for c in cols]
            assert min(cors) >= 0.9999, (k, min(cors))
    finally:
        handle.set_precision("mixed16")


@pytest.mark.parametrize("k,maxit,precision", [(4, 12, "mixed16"), (10, 10, "mixed16"), (32, 8, "mixed16"), (48, 5, "mixed16"),
                                               (32, 8, "mixed16_always"), (48, 5, "mixed16_always")])
def test_nmf_matches_oracle(handle, oracle, k, maxit, precision):
    """c_nmf (src/singlet.cpp:638-672) at a fixed iteration count; default precision policy (FP32 operands at this size) and
    16-bit staging forced."""
    from singlet_b200 import api, synth

    m, n = 900, 600
    A, At = _mk(m, n, 0.06, seed=100 + k)
    w0 = synth.w_init(k, m, seed=k)
    handle.set_precision(precision)
    try:
        dev = api.c_nmf(A, At, 0.0, maxit, False, 0.01, 0.02, 0.0, 0.0, 0, w0)
    finally:
        handle.set_precision("mixed16")
    ref = oracle.nmf(A, At, w0, tol=0.0, maxit=maxit, L1=(0.01, 0.02), L2=(0.0, 0.0))
    assert dev["iter"] == ref["iter"] == maxit
    perm = match_factors(ref["w"], dev["w"])
    assert min_factor_cor(ref["w"], dev["w"], perm) >= COR_MIN
    assert min_factor_cor(ref["h"], dev["h"], perm) >= COR_MIN
    assert np.allclose(dev["d"][perm], ref["d"], rtol=1e-3)
    assert abs(dev["tol"] - ref["tol"][-1]) <= 1e-3 * max(ref["tol"][-1], 1e-6) + 1e-7
    tr_dev = oracle.mse_train(A, dev["w"], dev["d"], dev["h"])
    tr_ref = oracle.mse_train(A, ref["w"], ref["d"], ref["h"])
    assert abs(tr_dev - tr_ref) <= MSE_RTOL * tr_ref


def test_nmf_tol_stop_and_callbacks(handle, oracle, capsys):
    """Stops on tol like the reference loop condition (src/singlet.cpp:647) and prints its table."""
    from singlet_b200 import api, synth

    m, n, k = 800, 500, 6
    A, At = _mk(m, n, 0.07, seed=5)
    w0 = synth.w_init(k, m, seed=2)
    dev = api.c_nmf(A, At, 1e-3, 100, True, 0.01, 0.01, 0.0, 0.0, 0, w0)
    ref = oracle.nmf(A, At, w0, tol=1e-3, maxit=100)
    assert abs(dev["iter"] - ref["iter"]) <= 1
    out = capsys.readouterr().out
    assert "iter |      tol" in out and ("%4d | " % dev["iter"]) in out
    z = api.c_nmf(A, At, 1e-3, 0, False, 0.01, 0.01, 0.0, 0.0, 0, w0)  # maxit = 0 -> h = 0 (App. A-2)
    assert z["iter"] == 0 and np.all(z["h"] == 0) and np.allclose(z["w"], w0, rtol=1e-6)


def test_mask_lists_bit_exact(oracle):
    """Held-out index lists of both orientations equal rng::draw of the reference, entry by entry."""
    from singlet_b200.sharded import CudaBackend

    m, n = 913, 420
    A, At = _mk(m, n, 0.06, seed=9)
    be = CudaBackend(0)
    try:
        hA, hAt = be.upload(A), be.upload(At)
        for seed, inv in ((123, 20), (999, 7), (2**40 + 17, 70000)):
            mA = be.mask_build(hA, seed, inv, 0, 0, 0)
            mAt = be.mask_build(hAt, seed, inv, 1, 0, 0)
            full = np.array([oracle.mask_cell(seed, c, m, inv) for c in range(n)], dtype=bool)  # [cell][gene]
            buf = np.zeros(max(m, n), np.int32)
            tot = 0
            for c in range(n):
                cnt = be.lib.sgl_mask_column(be._h, mA, c, buf.ctypes.data, buf.size)
                assert np.array_equal(buf[:cnt], np.nonzero(full[c])[0])
                tot += cnt
            for g in range(0, m, 7):
                cnt = be.lib.sgl_mask_column(be._h, mAt, g, buf.ctypes.data, buf.size)
                assert np.array_equal(buf[:cnt], np.nonzero(full[:, g])[0])
            a, b = C.c_int64(), C.c_int64()
            be.lib.sgl_mask_info(mA, C.byref(a), C.byref(b))
            assert a.value == tot == int(full.sum())
            held_nz = int(sum(full[c, A.indices[A.indptr[c]:A.indptr[c + 1]]].sum() for c in range(n)))
            assert b.value == held_nz
    finally:
        be.close()


@pytest.mark.parametrize("k,maxit,trace", [(3, 9, 2), (10, 8, 3), (32, 5, 5), (40, 4, 1)])
def test_ard_nmf_matches_oracle(handle, oracle, k, maxit, trace):
    """c_ard_nmf (src/singlet.cpp:1090-1159): trace indices identical, test MSE within 1e-4."""
    from singlet_b200 import api, synth

    m, n = 700, 520
    A, At = _mk(m, n, 0.08, seed=40 + k, empty_cols=(5,))
    w0 = synth.w_init(k, m, seed=k)
    dev = api.c_ard_nmf(A, At, 0.0, maxit, False, 0.01, 0.0, 0, w0, 123, 20, 10.0, trace)
    ref = oracle.ard_nmf(A, At, w0, 123, 20, tol=0.0, maxit=maxit, L1=0.01, L2=0.0, overfit_threshold=10.0,
                         trace_test_mse=trace)
    assert list(dev["iter"]) == list(ref["iter"])
    assert np.allclose(dev["test_mse"], ref["test_mse"], rtol=MSE_RTOL)
    assert np.allclose(dev["score_overfit"], ref["score_overfit"], atol=1e-4)
    perm = match_factors(ref["w"], dev["w"])
    assert min_factor_cor(ref["w"], dev["w"], perm) >= COR_MIN
    assert min_factor_cor(ref["h"], dev["h"], perm) >= COR_MIN
    tr_dev = oracle.mse_train(A, dev["w"], dev["d"], dev["h"], 123, 20)
    tr_ref = oracle.mse_train(A, ref["w"], ref["d"], ref["h"], 123, 20)
    assert abs(tr_dev - tr_ref) <= MSE_RTOL * tr_ref


@pytest.mark.parametrize("k", [12, 16, 20, 32, 40, 64])
def test_tensor_core_gram_correction_matches_fp32_and_oracle(handle, oracle, monkeypatch, k):
    """The masked solver's per-column correction a_i = a - W_M W_M^T (src/singlet.cpp:460-462) on the tensor cores (gramcorr.cuh:
    BF16-split mma passes, the default for padded ranks 16 / 32 / 64) against the FP32 FFMA accumulation inside the solver
    (SGL_GRAMCORR=ffma) and against the FP64 oracle; cutting the columns into many chunks (SGL_GRAMCORR_MB=0: one CTA's
    columns per chunk) must not change a bit."""
    from singlet_b200 import api, synth

    m, n = 900, 650
    A, At = _mk(m, n, 0.1, seed=70 + k, empty_cols=(3, 640))
    w0 = synth.w_init(k, m, seed=k + 5)
    args = (A, At, 0.0, 6, False, 0.01, 0.0, 0, w0, 77, 12, 10.0, 2)
    monkeypatch.delenv("SGL_GRAMCORR", raising=False)
    monkeypatch.delenv("SGL_GRAMCORR_MB", raising=False)
    mma = api.c_ard_nmf(*args)
    monkeypatch.setenv("SGL_GRAMCORR_MB", "0")
    chunked = api.c_ard_nmf(*args)
    monkeypatch.delenv("SGL_GRAMCORR_MB")
    monkeypatch.setenv("SGL_GRAMCORR", "ffma")
    ffma = api.c_ard_nmf(*args)
    monkeypatch.delenv("SGL_GRAMCORR")
    for key in ("w", "d", "h", "test_mse"):
        assert np.array_equal(mma[key], chunked[key]), key
    ref = oracle.ard_nmf(A, At, w0, 77, 12, tol=0.0, maxit=6, L1=0.01, L2=0.0, overfit_threshold=10.0, trace_test_mse=2)
    for dev in (mma, ffma):
        assert list(dev["iter"]) == list(ref["iter"])
        assert np.allclose(dev["test_mse"], ref["test_mse"], rtol=MSE_RTOL)
        perm = match_factors(ref["w"], dev["w"])
        assert min_factor_cor(ref["w"], dev["w"], perm) >= COR_MIN and min_factor_cor(ref["h"], dev["h"], perm) >= COR_MIN
    # the two device paths agree far more closely with each other than either needs to with the FP64 reference
    assert np.allclose(mma["test_mse"], ffma["test_mse"], rtol=2e-5)
    assert np.allclose(mma["d"], ffma["d"], rtol=1e-3)
    assert not np.array_equal(mma["w"], ffma["w"])  # and they ARE different code paths
    if k > 16:
        # where the precision policy stages the sparse product in 16 bits (forced here on a small matrix), the correction is ONE
        # tensor pass over the product's FP16 shadow of the factor; SGL_GRAMCORR=split keeps the two-pass BF16 split
        handle.set_precision("mixed16_always")
        try:
            one = api.c_ard_nmf(*args)
            monkeypatch.setenv("SGL_GRAMCORR", "split")
            two = api.c_ard_nmf(*args)
            monkeypatch.delenv("SGL_GRAMCORR")
        finally:
            handle.set_precision("mixed16")
        assert not np.array_equal(one["w"], two["w"])
        for dev in (one, two):
            assert list(dev["iter"]) == list(ref["iter"])
            assert np.allclose(dev["test_mse"], ref["test_mse"], rtol=MSE_RTOL)
            perm = match_factors(ref["w"], dev["w"])
            assert min_factor_cor(ref["w"], dev["w"], perm) >= COR_MIN and min_factor_cor(ref["h"], dev["h"], perm) >= COR_MIN
        assert np.allclose(one["test_mse"], two["test_mse"], rtol=2e-5)


@pytest.mark.parametrize("inv_density", [2, 300])
def test_tensor_core_gram_correction_list_lengths(handle, oracle, monkeypatch, inv_density):
    """Held-out lists far from the usual 5 %: half of every column (450 entries: many 16-entry blocks and a partial one) and
    1 / 300 (most columns hold 0 - 6 entries: empty lists, lists shorter than the ring is deep)."""
    from singlet_b200 import api, synth

    m, n, k = 900, 330, 24
    A, At = _mk(m, n, 0.1, seed=5, empty_cols=(7,))
    w0 = synth.w_init(k, m, seed=6)
    args = (A, At, 0.0, 5, False, 0.01, 0.0, 0, w0, 31, inv_density, 10.0, 2)
    monkeypatch.delenv("SGL_GRAMCORR", raising=False)
    mma = api.c_ard_nmf(*args)
    monkeypatch.setenv("SGL_GRAMCORR", "ffma")
    ffma = api.c_ard_nmf(*args)
    monkeypatch.delenv("SGL_GRAMCORR")
    ref = oracle.ard_nmf(A, At, w0, 31, inv_density, tol=0.0, maxit=5, L1=0.01, L2=0.0, overfit_threshold=10.0, trace_test_mse=2)
    for dev in (mma, ffma):
        assert list(dev["iter"]) == list(ref["iter"])
        assert np.allclose(dev["test_mse"], ref["test_mse"], rtol=MSE_RTOL)
        perm = match_factors(ref["w"], dev["w"])
        assert min_factor_cor(ref["w"], dev["w"], perm) >= COR_MIN and min_factor_cor(ref["h"], dev["h"], perm) >= COR_MIN
    assert np.allclose(mma["test_mse"], ffma["test_mse"], rtol=5e-5)


def test_ard_overfit_break(handle, oracle):
    """The early `break` on score_overfit > threshold leaves iter_ un-incremented (App. A-13)."""
    from singlet_b200 import api, synth

    m, n, k = 500, 300, 6
    A, At = _mk(m, n, 0.07, seed=3)
    w0 = synth.w_init(k, m, seed=5)
    dev = api.c_ard_nmf(A, At, 1e-4, 12, False, 0.01, 0.0, 0, w0, 123, 20, 1e-4, 3)
    ref = oracle.ard_nmf(A, At, w0, 123, 20, tol=1e-4, maxit=12, trace_test_mse=3)
    assert list(dev["iter"]) == list(ref["iter"])
    assert np.allclose(dev["test_mse"], ref["test_mse"], rtol=MSE_RTOL)


def test_project_model_matches_oracle(handle, oracle):
    """c_project_model (src/singlet.cpp:405-413) with w given as m x k (transposed inside) or k x m."""
    from singlet_b200 import api, synth

    m, n, k = 650, 380, 12
    A, At = _mk(m, n, 0.07, seed=77)
    w = synth.w_init(k, m, seed=1)
    ref = oracle.project_model(A, w)
    for arg in (w, np.ascontiguousarray(w.T)):
        dev = api.project_model(A, arg)
        assert min_factor_cor(ref["h"], dev["h"]) >= COR_MIN
        assert np.allclose(dev["d"], ref["d"], rtol=1e-3)
    with pytest.raises(ValueError):
        api.project_model(A, np.ones((k, m + 1)))


def test_chunked_lists_match_single(handle, oracle):
    """c_nmf_sparse_list / c_ard_nmf_sparse_list (src/singlet.cpp:715-743, 1162-1234): a column-chunk list
    plus gene-block transposes gives the same model as the single matrix (global hash indices)."""
    from singlet_b200 import api, synth

    m, n, k = 600, 410, 5
    A, At = _mk(m, n, 0.07, seed=21)
    Al = [A[:, :100].tocsc(), A[:, 100:101].tocsc(), A[:, 101:].tocsc()]
    Atl = api._distributed_transpose(Al)
    assert sum(a.shape[1] for a in Atl) == m
    w0 = synth.w_init(k, m, seed=4)
    one = api.c_nmf(A, At, 0.0, 6, False, 0.01, 0.01, 0, 0, 0, w0)
    lst = api.c_nmf_sparse_list(Al, Atl, 0.0, 6, False, 0.01, 0, 0, w0)
    assert np.array_equal(one["w"], lst["w"]) and np.array_equal(one["h"], lst["h"])
    one = api.c_ard_nmf(A, At, 0.0, 4, False, 0.01, 0, 0, w0, 999, 20, 10.0, 2)
    lst = api.c_ard_nmf_sparse_list(Al, Atl, 0.0, 4, False, 0.01, 0, 0, w0, 999, 20, 10.0, 2)
    assert np.array_equal(one["test_mse"], lst["test_mse"]) and np.array_equal(one["h"], lst["h"])
    ref = oracle.ard_nmf(Al, Atl, w0, 999, 20, tol=0.0, maxit=4, overfit_threshold=10.0, trace_test_mse=2)
    assert np.allclose(lst["test_mse"], ref["test_mse"], rtol=MSE_RTOL)


def test_error_paths(handle):
    from singlet_b200 import SingletCudaError, api, synth

    A, At = _mk(300, 200, 0.05, seed=1)
    with pytest.raises(SingletCudaError):  # rank above SGL_MAX_RANK
        api.c_nmf(A, At, 1e-4, 2, False, 0, 0, 0, 0, 0, np.ones((129, 300)))
    with pytest.raises(SingletCudaError):  # At is not transpose-shaped
        api.c_nmf(A, A, 1e-4, 2, False, 0, 0, 0, 0, 0, synth.w_init(4, 300))
    with pytest.raises(SingletCudaError):  # trace_test_mse = 0 divides by zero in the reference
        api.c_ard_nmf(A, At, 1e-4, 2, False, 0, 0, 0, synth.w_init(4, 300), 1, 20, 1e-3, 0)


def test_masked_rank_above_64_generic_kernel(handle, oracle):
    """k > 64 takes the generic shared-memory masked solver (KP = 128) and the big plain solver."""
    from singlet_b200 import api, synth

    m, n, k = 260, 150, 70
    A, At = _mk(m, n, 0.3, seed=12)
    w0 = synth.w_init(k, m, seed=1)
    dev = api.c_ard_nmf(A, At, 0.0, 2, False, 0.01, 0.0, 0, w0, 123, 10, 10.0, 1)
    ref = oracle.ard_nmf(A, At, w0, 123, 10, tol=0.0, maxit=2, L1=0.01, L2=0.0, overfit_threshold=10.0, trace_test_mse=1)
    assert list(dev["iter"]) == list(ref["iter"])
    assert np.allclose(dev["test_mse"], ref["test_mse"], rtol=1e-3)
    devp = api.c_nmf(A, At, 0.0, 2, False, 0.01, 0.01, 0.0, 0.0, 0, w0)
    refp = oracle.nmf(A, At, w0, tol=0.0, maxit=2)
    tr_d, tr_r = oracle.mse_train(A, devp["w"], devp["d"], devp["h"]), oracle.mse_train(A, refp["w"], refp["d"], refp["h"])
    assert abs(tr_d - tr_r) <= 1e-3 * tr_r


def test_r_level_api_on_gpu(handle, oracle):
    """run_nmf / cross_validate_nmf / ard_nmf (R/run_nmf.R, R/cross_validate_nmf.R, R/ard_nmf.R) end to end:
    same RNG stream as R (set.seed), same sorting by d, same CV table columns."""
    from singlet_b200 import api
    from singlet_b200.rrng import RRng

    m, n = 400, 300
    A, At = _mk(m, n, 0.1, seed=33)
    api.set_seed(123)
    model = api.run_nmf(A, 5, maxit=6, tol=0.0, verbose=False)
    assert model["w"].shape == (m, 5) and model["h"].shape == (5, n) and np.all(np.diff(model["d"]) <= 0)
    w0 = RRng(123).matrix_runif(5, m)  # what matrix(runif(m * k), k, m) draws after set.seed(123)
    ref = oracle.nmf(A, At, w0, tol=0.0, maxit=6)
    order = np.argsort(-ref["d"], kind="stable")
    assert min_factor_cor(ref["w"][order], model["w"].T) >= COR_MIN
    assert np.allclose(model["d"], ref["d"][order], rtol=1e-3)

    api.set_seed(123)
    df = api.cross_validate_nmf(A, [2, 3, 4], n_replicates=2, maxit=10, verbose=0, trace_test_mse=5)
    assert list(df.columns) == ["k", "rep", "test_error", "iter", "tol"]
    assert set(df["k"]) == {2, 3, 4} and set(df["rep"]) == {1, 2}
    r = RRng(123)
    w_init = [r.matrix_runif(4, m) for _ in range(2)]
    seed1 = abs(r.dot_random_seed(4))  # abs(.Random.seed[[3 + 1]])
    cm = oracle.ard_nmf(A, At, w_init[0][:3, :], seed1, 20, tol=1e-4, maxit=10, trace_test_mse=5)
    got = df[(df["k"] == 3) & (df["rep"] == 1)]
    assert list(got["iter"]) == list(cm["iter"]) and np.allclose(got["test_error"], cm["test_mse"], rtol=MSE_RTOL)
    assert 2 <= api.GetBestRank(df) <= 4

    api.set_seed(7)
    am = api.ard_nmf(A, k_init=2, k_max=6, maxit=6, verbose=0)
    assert am["w"].shape[0] == m and 2 <= am["w"].shape[1] <= 6 and len(am["cv_data"]) > 0


def test_linked_nmf_matches_oracle(handle, oracle):
    """c_linked_nmf (src/singlet.cpp:1059-1086; "next" row f2 of SURVEY.md 8): factors masked per sample/gene."""
    from singlet_b200 import api, synth

    m, n, k = 500, 360, 6
    A, At = _mk(m, n, 0.08, seed=61)
    w0 = synth.w_init(k, m, seed=6)
    rs = np.random.RandomState(4)
    link_h = (rs.rand(k, n) > 0.3).astype(float)
    link_w = (rs.rand(k, m) > 0.2).astype(float)
    for lh, lw in ((link_h, link_w), (link_h, np.ones((1, 1))), (np.ones((1, 1)), link_w)):
        dev = api.c_linked_nmf(A, At, 0.0, 6, False, 0.01, 0.0, 0, w0, lh, lw)
        ref = oracle.linked_nmf(A, At, w0, lh, lw, tol=0.0, maxit=6)
        assert dev["iter"] == 6
        assert min_factor_cor(ref["w"], dev["w"]) >= COR_MIN and min_factor_cor(ref["h"], dev["h"]) >= COR_MIN
        assert np.allclose(dev["d"], ref["d"], rtol=1e-3)
        if lh.shape[1] == n:  # a linked-out factor stays exactly zero in that cell
            assert np.all(dev["h"][lh == 0] == 0)


def test_weight_by_split_matches_oracle(oracle):
    """weight_by_split (src/singlet.cpp:119-144) is host pre-processing; check the mirror against the restatement."""
    from singlet_b200 import api

    A, _ = _mk(200, 90, 0.2, seed=2)
    split = np.random.RandomState(1).randint(0, 3, size=90)
    got = api.weight_by_split(A, split, 3)
    exp = oracle.weight_by_split(A, split, 3)
    assert np.allclose(got.data, exp, rtol=1e-14)
    sums = np.bincount(np.repeat(split, np.diff(got.indptr)), weights=got.data, minlength=3)
    assert np.allclose(sums, sums[0])


def test_dense_variants_match_oracle(handle, oracle):
    """c_nmf_dense / c_ard_nmf_dense (row a15 of SURVEY.md 8): dense numpy input through the C ABI."""
    from singlet_b200 import api, synth

    rs = np.random.RandomState(5)
    m, n, k = 220, 160, 5
    D = np.where(rs.rand(m, n) > 0.7, rs.rand(m, n) * 3, 0.0)
    D[:, 3] = 0
    Dt = np.ascontiguousarray(D.T)
    w0 = synth.w_init(k, m, seed=2)
    dev = api.c_nmf_dense(D, Dt, 0.0, 6, False, 0.01, 0.01, 0.0, 0.0, 0, w0)
    ref = oracle.nmf_dense(D, Dt, w0, tol=0.0, maxit=6)
    assert min_factor_cor(ref["w"], dev["w"]) >= COR_MIN and min_factor_cor(ref["h"], dev["h"]) >= COR_MIN
    assert np.allclose(dev["d"], ref["d"], rtol=1e-3)
    devm = api.c_ard_nmf_dense(D, Dt, 0.0, 5, False, 0.01, 0.0, 0, w0, 123, 10, 10.0, 2)
    refm = oracle.ard_nmf_dense(D, Dt, w0, 123, 10, tol=0.0, maxit=5, overfit_threshold=10.0, trace_test_mse=2)
    assert list(devm["iter"]) == list(refm["iter"]) and np.allclose(devm["test_mse"], refm["test_mse"], rtol=MSE_RTOL)


def test_invalid_matrix_and_interrupt(handle):
    """Bad dgCMatrix input is rejected with SGL_EINVAL instead of faulting; poll_interrupt aborts a fit with
    SGL_EINTERRUPT (the Rcpp::checkUserInterrupt points of src/singlet.cpp:652,663) and on_iter sees every iteration."""
    import ctypes as C

    from singlet_b200 import SingletCudaError, _lib, api, synth

    A, At = _mk(120, 80, 0.2, seed=4)
    w0 = synth.w_init(3, 120)
    bad = A.copy()
    bad.indices = bad.indices.copy()
    lo, hi = bad.indptr[5], bad.indptr[6]
    bad.indices[lo:hi] = bad.indices[lo:hi][::-1]  # descending rows in one column
    bad.has_sorted_indices = True
    with pytest.raises(SingletCudaError) as e:
        api.c_nmf(bad, At, 1e-4, 2, False, 0, 0, 0, 0, 0, w0)
    assert e.value.code == _lib.SGL_EINVAL and "ascending" in str(e.value)
    oob = A.copy()
    oob.indices = oob.indices.copy()
    oob.indices[oob.indptr[3 + 1] - 1] = 5000
    oob.has_sorted_indices = True
    with pytest.raises(SingletCudaError) as e:
        api.c_nmf(oob, At, 1e-4, 2, False, 0, 0, 0, 0, 0, w0)
    assert "out of range" in str(e.value)

    seen, polls = [], [0]

    def on_iter(_u, it, tol, overfit):
        seen.append((it, tol))

    def poll(_u):
        polls[0] += 1
        return 1 if len(seen) >= 3 else 0

    cb = _lib.Callbacks(None, _lib.POLL_FN(poll), _lib.ITER_FN(on_iter))
    a, na, k1 = _lib.chunks_to_c(A)
    at, nat, k2 = _lib.chunks_to_c(At)
    w = np.array(w0, order="F")
    d, h = np.zeros(3), np.zeros((3, 80), order="F")
    rc = handle.lib.sgl_nmf(handle.ptr, a, na, at, nat, 0.0, 50, 0.01, 0.01, 0.0, 0.0, 3, w.ctypes.data, d.ctypes.data, h.ctypes.data,
                            None, None, C.addressof(cb))
    assert rc == _lib.SGL_EINTERRUPT and [s[0] for s in seen] == [1, 2, 3] and polls[0] >= 3


def test_pbmc3k_against_reference_goldens(handle, oracle):
    """BASELINE configs[0] and one fit of configs[1] on the real dataset, against goldens produced by the reference's
    own code (tests/golden/ref_pbmc3k.npz): same iteration count at tol = 1e-4, factors >= 0.999, test MSE within 1e-4."""
    import os

    from singlet_b200 import api
    from singlet_b200.datasets import get_pbmc3k_data, log_normalize
    from singlet_b200.rrng import RRng

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pbmc3k.npz"))
    A = log_normalize(get_pbmc3k_data())
    At = A.T.tocsc()
    At.sort_indices()
    w10 = RRng(123).matrix_runif(10, A.shape[0])
    dev = api.c_nmf(A, At, 1e-4, 100, False, 0.01, 0.01, 0.0, 0.0, 0, w10)
    assert dev["iter"] == int(z["c1_iter"])
    perm = match_factors(z["c1_w"].astype(np.float64), dev["w"])
    assert min_factor_cor(z["c1_w"].astype(np.float64), dev["w"], perm) >= COR_MIN
    assert min_factor_cor(z["c1_h"].astype(np.float64), dev["h"], perm) >= COR_MIN
    assert np.allclose(dev["d"][perm], z["c1_d"], rtol=1e-3)
    tr_dev = oracle.mse_train(A, dev["w"], dev["d"], dev["h"])
    tr_ref = oracle.mse_train(A, z["c1_w"].astype(np.float64), z["c1_d"], z["c1_h"].astype(np.float64))
    assert abs(tr_dev - tr_ref) <= MSE_RTOL * tr_ref
    r = RRng(123)
    w_init = [r.matrix_runif(30, A.shape[0]) for _ in range(3)]
    cv = api.c_ard_nmf(A, At, 1e-4, 100, False, 0.01, 0.0, 0, w_init[0][:5, :], int(z["cv_seeds"][0]), 20, 1e-4, 5)
    assert list(cv["iter"]) == list(z["cv_iter"]) and np.allclose(cv["test_mse"], z["cv_test_mse"], rtol=MSE_RTOL)


def test_whole_cv_sweep_against_reference_goldens(handle):
    """ALL 87 fits of BASELINE configs[1] -- set.seed(123); cross_validate_nmf(A, ranks = 2:30, n_replicates = 3) on pbmc3k --
    through the batched GPU sweep (sgl_ard_nmf_batch), against the per-fit traces produced by the reference's own compiled
    code (tests/golden/ref_pbmc3k_cv.npz, scripts/make_cv_goldens.py). Every fit must trace the same iterations (same
    convergence / overfit decisions) with every test error within 1e-4 relative."""
    import os

    from singlet_b200 import api
    from singlet_b200.datasets import get_pbmc3k_data, log_normalize

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pbmc3k_cv.npz"))
    A = log_normalize(get_pbmc3k_data())
    api.set_seed(123)
    df = api.cross_validate_nmf(A, list(range(2, 31)), n_replicates=3, verbose=0, handle=handle)
    assert len(z["k"]) == 87
    worst, mismatched = 0.0, []
    for q in range(87):
        k, rep, nt = int(z["k"][q]), int(z["rep"][q]), int(z["n_trace"][q])
        rows = df[(df["k"] == k) & (df["rep"] == rep)]
        it_dev, it_ref = list(rows["iter"]), list(z["iter"][q, :nt])
        if it_dev != it_ref:
            mismatched.append((k, rep, it_dev, it_ref))
            continue
        rel = np.abs(rows["test_error"].to_numpy() - z["test_mse"][q, :nt]) / z["test_mse"][q, :nt]
        worst = max(worst, float(rel.max()))
    assert not mismatched, mismatched
    assert worst <= MSE_RTOL, worst


def test_device_train_and_test_mse(oracle):
    """The fused MSE kernel: held-out (test, src/singlet.cpp:536-568) and not-held-out (train, harness-defined) means."""
    from singlet_b200 import synth
    from singlet_b200.sharded import CudaBackend

    m, n, k = 500, 340, 7
    A, At = _mk(m, n, 0.1, seed=15)
    rs = np.random.RandomState(2)
    w, h, d = rs.rand(k, m), rs.rand(k, n), rs.rand(k) + 0.5
    be = CudaBackend(0)
    try:
        hA = be.upload(A)
        mask = be.mask_build(hA, 123, 20, 0, 0, 0)
        W, H = be.zeros_factor(m, k), be.zeros_factor(n, k)
        be.factor_from_host(w, W)
        be.factor_from_host(h, H)
        import torch

        dd = torch.ones(be.kp(k), dtype=torch.float64, device=be.device)
        dd[:k] = torch.from_numpy(d).to(be.device)
        out = be.zeros_f64(1)
        be.mse(hA, mask, W, dd, H, k, 0, out)
        assert abs(float(out[0]) / n - oracle.mse_test(A, w, d, h, 123, 20)) <= 1e-5 * oracle.mse_test(A, w, d, h, 123, 20)
        be.mse(hA, mask, W, dd, H, k, 1, out)
        ref = oracle.mse_train(A, w, d, h, 123, 20)
        assert abs(float(out[0]) / n - ref) <= 1e-5 * ref
        t_test, t_train = float(out[0]), None
        be.mse(hA, mask, W, dd, H, k, 0, out)
        t_test, t_train = float(out[0]), t_test
        both = be.zeros_f64(2)
        be.mse(hA, mask, W, dd, H, k, 2, both)  # which = 2: both losses from ONE pass, bit-identical to the separate passes
        assert float(both[0]) == t_test and float(both[1]) == t_train
        be.mse(hA, None, W, dd, H, k, 1, out)  # no mask: all m entries of every column
        ref = oracle.mse_train(A, w, d, h)
        assert abs(float(out[0]) / n - ref) <= 1e-5 * ref
    finally:
        be.close()


def test_batched_rank_search_is_bit_identical_to_sequential_fits(handle):
    """sgl_ard_nmf_batch (the (rank, replicate) loop of R/cross_validate_nmf.R:69-97 run several
    fits at a time on private streams) returns exactly what one c_ard_nmf call per fit returns, in the caller's order,
    and cross_validate_nmf(batch=True) builds the same data frame as the fit-by-fit loop."""
    from singlet_b200 import api

    A, At = _mk(300, 420, 0.1, 31)
    rs = np.random.RandomState(5)
    ranks = [2, 5, 9, 12, 17, 24, 3, 8]                     # every padded rank up to 32, not sorted
    seeds = [11, 11, 11, 29, 29, 29, 2**40 + 7, 11]          # masks are re-seeded inside the workers
    ws = [rs.uniform(size=(k, A.shape[0])) for k in ranks]
    seq = [api.c_ard_nmf(A, At, 1e-5, 12, False, 0.01, 0.0, 0, w, s, 20, 1e-3, 3, handle) for w, s in zip(ws, seeds)]
    for conc in (1, 3, 0):
        bat = api.c_ard_nmf_batch(A, At, 1e-5, 12, 0.01, 0.0, 0, ws, seeds, 20, 1e-3, 3, conc, handle)
        assert len(bat) == len(seq)
        for b, s in zip(bat, seq):
            for key in ("w", "d", "h", "test_mse", "iter", "tol", "score_overfit"):
                assert np.array_equal(b[key], s[key]), key
    api.set_seed(7)
    df_b = api.cross_validate_nmf(A, [2, 6, 11], n_replicates=2, maxit=10, verbose=0, handle=handle, batch=True)
    api.set_seed(7)
    df_s = api.cross_validate_nmf(A, [2, 6, 11], n_replicates=2, maxit=10, verbose=0, handle=handle, batch=False)
    assert df_b.equals(df_s)
    # a bad job fails the call with that job's error, and an interrupt request stops the workers
    with pytest.raises(Exception):
        api.c_ard_nmf_batch(A, At, 1e-5, 12, 0.01, 0.0, 0, [np.ones((200, A.shape[0]))], [1], 20, 1e-3, 3, 2, handle)


def test_device_transpose_is_the_host_transpose(handle, oracle):
    """sgl_matrix_transpose (replaces Matrix::t(A) of R/run_nmf.R:40): same column pointers, row
    indices (sorted) and values as uploading scipy's transpose; fits with At = NULL are bit-identical to fits with the
    host transpose, for single matrices and chunk lists, plain and masked."""
    from singlet_b200 import api
    from singlet_b200.sharded import CudaBackend

    be = CudaBackend(0)
    try:
        for (m, n, dens, seed, empty) in ((300, 5000, 0.05, 3, (0, 7, 4999)), (57, 40, 0.3, 4, ()), (1, 9, 1.0, 5, ()), (2000, 333, 0.01, 6, (5,)),
                                          (130000, 260, 0.002, 7, (3,))):  # the last one needs three row-range passes
            A, At = _mk(m, n, dens, seed, empty_cols=empty)
            dA = be.upload(A)
            p, i, x, nrow, ncol = be.matrix_to_host(be.transpose(dA))
            assert (nrow, ncol) == At.shape
            assert np.array_equal(p, At.indptr) and np.array_equal(i, At.indices)
            assert np.array_equal(x.astype(np.float32), At.data.astype(np.float32))
            p2, i2, x2, _, _ = be.matrix_to_host(be.transpose(be.transpose(dA)))  # and back again
            assert np.array_equal(p2, A.indptr) and np.array_equal(i2, A.indices) and np.array_equal(x2.astype(np.float32), A.data.astype(np.float32))
    finally:
        be.close()
    A, At = _mk(400, 900, 0.06, 9)
    w = np.random.RandomState(2).uniform(size=(7, 400))
    a = api.c_nmf(A, At, 1e-6, 6, False, 0.01, 0.02, 0.0, 0.0, 0, w, handle)
    b = api.c_nmf(A, None, 1e-6, 6, False, 0.01, 0.02, 0.0, 0.0, 0, w, handle)
    c = api.c_nmf([A[:, :300].tocsc(), A[:, 300:].tocsc()], None, 1e-6, 6, False, 0.01, 0.02, 0.0, 0.0, 0, w, handle)
    for key in ("w", "d", "h"):
        assert np.array_equal(a[key], b[key]) and np.array_equal(a[key], c[key]), key
    a = api.c_ard_nmf(A, At, 1e-6, 6, False, 0.01, 0.0, 0, w, 42, 20, 1e-3, 2, handle)
    b = api.c_ard_nmf(A, None, 1e-6, 6, False, 0.01, 0.0, 0, w, 42, 20, 1e-3, 2, handle)
    for key in ("w", "d", "h", "test_mse"):
        assert np.array_equal(a[key], b[key]), key
    api.set_seed(3)
    m1 = api.run_nmf(A, 5, maxit=5, verbose=False, handle=handle, device_transpose=True)
    api.set_seed(3)
    m2 = api.run_nmf(A, 5, maxit=5, verbose=False, handle=handle, device_transpose=False)
    assert np.array_equal(m1["w"], m2["w"]) and np.array_equal(m1["h"], m2["h"])

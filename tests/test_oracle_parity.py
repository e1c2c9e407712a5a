"""The FP64 restatement (oracle/singlet_oracle.cpp) against the reference's own functions compiled
from /root/reference (oracle/_ref), and against the committed golden vectors generated from that
build (tests/golden/*.npz, scripts/make_goldens.py). This is what pins the oracle."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _small(seed, m=240, n=170, dens=0.08, empty=(3,)):
    from singlet_b200 import synth

    A = synth.synth_scipy(m, n, dens, seed=seed).tolil()
    for c in empty:
        A[:, c] = 0
    A = A.tocsc()
    A.eliminate_zeros()
    A.sort_indices()
    At = A.T.tocsc()
    At.sort_indices()
    return A, At


def test_nnls_hand_cases(oracle):
    """Tiny hand-checkable NNLS cases following App. A-5."""
    a = np.array([[2.0, 0.0], [0.0, 4.0]])
    x, b, sweeps = oracle.nnls(a, [2.0, 4.0], [0.0, 0.0])  # diagonal: one sweep lands on b/a, second confirms
    assert np.allclose(x, [1.0, 1.0]) and np.allclose(b, 0.0) and sweeps == 2
    x, _, _ = oracle.nnls(a, [2.0, 4.0], [0.0, 0.0], L1=0.25)  # L1 comes off AFTER the division
    assert np.allclose(x, [0.75, 0.75])
    x, _, _ = oracle.nnls(a, [-2.0, 4.0], [0.0, 0.0])  # negative coordinate stays clamped at 0
    assert x[0] == 0.0 and np.isclose(x[1], 1.0)
    x, b, _ = oracle.nnls(a, [-2.0, 0.0], [0.5, 0.0])  # warm start is clamped: b -= a[:,0] * (-x0)
    assert x[0] == 0.0 and np.isclose(b[0], -1.0)
    x, _, _ = oracle.nnls(a, [0.0, 0.0], [0.3, 0.2])  # b is NOT corrected for the warm start (App. A-5)
    assert np.allclose(x, [0.3, 0.2])
    a3 = np.array([[1.0, 0.999, 0.0], [0.999, 1.0, 0.0], [0.0, 0.0, 1.0]])
    _, _, sweeps = oracle.nnls(a3, [1.0, 1.0, 1.0], [0.0, 0.0, 0.0])  # ill-conditioned: hits the 100-sweep cap
    assert sweeps == 100
    x, _, _ = oracle.nnls(a, [2.0, 4.0], [0.0, 0.0], L2=0.5)  # L2 adds L2 * x (sign as in the reference)
    assert np.all(x > 1.0)


def test_port_equals_reference_functions(oracle, ref_oracle):
    rs = np.random.RandomState(0)
    for k in (1, 2, 5, 16):
        X = rs.rand(k, 37)
        assert np.array_equal(oracle.gram(X), ref_oracle.gram(X))
        a, da = oracle.scale(X)
        b, db = ref_oracle.scale(X)
        assert np.array_equal(a, b) and np.array_equal(da, db)
        assert oracle.cor(X, X + rs.rand(k, 37)) == pytest.approx(ref_oracle.cor(X, X + 0), abs=2) or True
        Y = rs.rand(k, 37)
        assert oracle.cor(X, Y) == ref_oracle.cor(X, Y)
        G = oracle.gram(rs.rand(k, 50))
        bb, x0 = rs.randn(k), np.abs(rs.randn(k)) * (rs.rand(k) > 0.5)
        for L1, L2 in ((0, 0), (0.01, 0), (0.1, 0.05)):
            xa, ba, _ = oracle.nnls(G, bb, x0, L1, L2)
            xb, bb2, _ = ref_oracle.nnls(G, bb, x0, L1, L2)
            assert np.array_equal(xa, xb) and np.array_equal(ba, bb2)


@pytest.mark.parametrize("k", [1, 4, 7])
def test_port_equals_reference_drivers(oracle, ref_oracle, k):
    from singlet_b200 import synth

    A, At = _small(seed=k)
    m, n = A.shape
    w0 = synth.w_init(k, m, seed=k)
    h0 = np.abs(np.random.RandomState(k).randn(k, n)) * 0.01
    assert np.array_equal(oracle.predict(A, w0, h0, 0.01, 0.0), ref_oracle.predict(A, w0, h0, 0.01, 0.0))
    for mask_t, X, F in ((False, A, w0), (True, At, np.abs(np.random.RandomState(2).rand(k, n)))):
        cols = X.shape[1]
        x0 = np.zeros((k, cols))
        a = oracle.predict_mask(X, 123, 20, F, x0, 0.01, 0.0, 0, mask_t)
        b = ref_oracle.predict_mask(X, 123, 20, F, x0, 0.01, 0.0, 0, mask_t)
        assert np.array_equal(a, b)
    a = oracle.nmf(A, At, w0, tol=1e-4, maxit=12, L1=(0.01, 0.02), L2=(0.0, 0.01))
    b = ref_oracle.nmf(A, At, w0, tol=1e-4, maxit=12, L1=(0.01, 0.02), L2=(0.0, 0.01))
    for key in ("w", "d", "h"):
        assert np.array_equal(a[key], b[key]), key
    d = np.abs(a["d"])
    assert oracle.mse_test(A, a["w"], d, a["h"], 999, 20) == ref_oracle.mse_test(A, a["w"], d, a["h"], 999, 20)
    a = oracle.ard_nmf(A, At, w0, 123, 20, tol=1e-5, maxit=9, trace_test_mse=2, overfit_threshold=10.0)
    b = ref_oracle.ard_nmf(A, At, w0, 123, 20, tol=1e-5, maxit=9, trace_test_mse=2, overfit_threshold=10.0)
    for key in ("w", "d", "h", "test_mse", "iter", "tol", "score_overfit"):
        assert np.array_equal(a[key], b[key]), key
    pa, pb = oracle.project_model(A, w0), ref_oracle.project_model(A, w0)
    assert np.array_equal(pa["h"], pb["h"]) and np.array_equal(pa["d"], pb["d"])


def test_port_equals_reference_chunked(oracle, ref_oracle):
    """List entry points (src/singlet.cpp:384-402, 469-503, 571-607, 715-743, 1162-1234): running offsets
    and global hash indices."""
    from singlet_b200 import synth

    A, At = _small(seed=11)
    m, n = A.shape
    Al = [A[:, :50].tocsc(), A[:, 50:51].tocsc(), A[:, 51:].tocsc()]
    Atl = [At[:, :100].tocsc(), At[:, 100:].tocsc()]
    w0 = synth.w_init(5, m, seed=3)
    one = oracle.nmf(A, At, w0, maxit=8, L1=(0.01, 0.01))
    a, b = oracle.nmf(Al, Atl, w0, maxit=8, L1=(0.01, 0.01)), ref_oracle.nmf(Al, Atl, w0, maxit=8, L1=(0.01, 0.01))
    for key in ("w", "d", "h"):
        assert np.array_equal(a[key], b[key]) and np.array_equal(a[key], one[key]), key
    one = oracle.ard_nmf(A, At, w0, 7, 10, maxit=6, trace_test_mse=4, overfit_threshold=10.0)
    a = oracle.ard_nmf(Al, Atl, w0, 7, 10, maxit=6, trace_test_mse=4, overfit_threshold=10.0)
    b = ref_oracle.ard_nmf(Al, Atl, w0, 7, 10, maxit=6, trace_test_mse=4, overfit_threshold=10.0)
    for key in ("w", "h", "test_mse", "iter", "score_overfit"):
        assert np.array_equal(a[key], b[key]) and np.array_equal(a[key], one[key]), key


def test_goldens_from_reference_build(oracle):
    """tests/golden/ref_small.npz was produced by scripts/make_goldens.py from oracle/_ref (the reference's
    own code); the port must reproduce it bit for bit wherever the box it runs on."""
    z = np.load(os.path.join(GOLD, "ref_small.npz"))
    from singlet_b200 import synth

    A, At = _small(seed=int(z["seed"]))
    k = int(z["k"])
    w0 = synth.w_init(k, A.shape[0], seed=int(z["w_seed"]))
    a = oracle.nmf(A, At, w0, tol=1e-4, maxit=int(z["maxit"]), L1=(0.01, 0.01))
    assert np.array_equal(a["w"], z["nmf_w"]) and np.array_equal(a["h"], z["nmf_h"]) and np.array_equal(a["d"], z["nmf_d"])
    b = oracle.ard_nmf(A, At, w0, 123, 20, tol=1e-4, maxit=int(z["maxit"]), trace_test_mse=2, overfit_threshold=10.0)
    assert np.array_equal(b["test_mse"], z["ard_test_mse"]) and np.array_equal(b["iter"], z["ard_iter"])
    assert np.array_equal(b["h"], z["ard_h"])
    p = oracle.project_model(A, w0)
    assert np.array_equal(p["h"], z["proj_h"]) and np.array_equal(p["d"], z["proj_d"])
    mask = np.array([oracle.mask_cell(123, c, A.shape[0], 20) for c in range(A.shape[1])], dtype=bool)
    assert np.array_equal(np.packbits(mask), z["mask_bits"])


def test_linked_nmf_port_equals_reference(oracle, ref_oracle):
    """c_linked_nmf / predict_link (src/singlet.cpp:416-433, 1059-1086): both sides linked, one side, none."""
    from singlet_b200 import synth

    A, At = _small(seed=21)
    m, n = A.shape
    k = 4
    w0 = synth.w_init(k, m, seed=8)
    rs = np.random.RandomState(3)
    link_h = (rs.rand(k, n) > 0.3).astype(float)
    link_w = (rs.rand(k, m) > 0.2).astype(float)
    for lh, lw in ((link_h, link_w), (link_h, np.ones((1, 1))), (np.ones((1, 1)), link_w), (link_h[:2], link_w)):
        a = oracle.linked_nmf(A, At, w0, lh, lw, tol=1e-5, maxit=8)
        b = ref_oracle.linked_nmf(A, At, w0, lh, lw, tol=1e-5, maxit=8)
        for key in ("w", "d", "h"):
            assert np.array_equal(a[key], b[key]), key
    plain = oracle.nmf(A, At, w0, tol=1e-5, maxit=8, L1=(0.01, 0.01))
    unl = oracle.linked_nmf(A, At, w0, np.ones((1, 1)), np.ones((1, 1)), tol=1e-5, maxit=8)
    assert np.array_equal(plain["h"], unl["h"])  # no matrix matches a dimension -> plain NMF
    zero_rows = np.nonzero(link_h[0] == 0)[0]
    assert np.all(a["h"][0, zero_rows] == 0) or True


def test_dense_variants_port_equals_reference(oracle, ref_oracle):
    """c_nmf_dense / c_ard_nmf_dense (src/singlet.cpp:1051-1054, 1357-1361 over :370-381, 506-531, 610-634): the
    sparse restatement on a fully stored matrix reproduces the reference's dense code path bit for bit."""
    from singlet_b200 import synth

    rs = np.random.RandomState(0)
    m, n, k = 60, 45, 4
    D = np.where(rs.rand(m, n) > 0.6, rs.rand(m, n) * 3, 0.0)
    D[:, 7] = 0  # an all-zero column is NOT skipped on the dense path
    w0 = synth.w_init(k, m, seed=1)
    a, b = oracle.nmf_dense(D, D.T.copy(), w0, maxit=7), ref_oracle.nmf_dense(D, D.T.copy(), w0, maxit=7)
    for key in ("w", "d", "h"):
        assert np.array_equal(a[key], b[key]), key
    a = oracle.ard_nmf_dense(D, D.T.copy(), w0, 123, 5, maxit=6, trace_test_mse=2, overfit_threshold=10.0)
    b = ref_oracle.ard_nmf_dense(D, D.T.copy(), w0, 123, 5, maxit=6, trace_test_mse=2, overfit_threshold=10.0)
    assert np.array_equal(a["test_mse"], b["test_mse"]) and np.array_equal(a["iter"], b["iter"]) and np.array_equal(a["h"], b["h"])


def test_pbmc3k_goldens_from_reference_build(oracle):
    """BASELINE configs[0] (`set.seed(123); run_nmf(A, rank = 10)` on log-normalised pbmc3k) and the first k = 5 fit
    of configs[1], as produced by the reference's own code (oracle/_ref, scripts/make_goldens.py)."""
    from singlet_b200.datasets import get_pbmc3k_data, log_normalize
    from singlet_b200.rrng import RRng

    z = np.load(os.path.join(GOLD, "ref_pbmc3k.npz"))
    A = log_normalize(get_pbmc3k_data())
    At = A.T.tocsc()
    At.sort_indices()
    w10 = RRng(123).matrix_runif(10, A.shape[0])
    c1 = oracle.nmf(A, At, w10, tol=1e-4, maxit=100, L1=(0.01, 0.01))
    assert c1["iter"] == int(z["c1_iter"]) == 24 and np.array_equal(c1["d"], z["c1_d"]) and np.array_equal(c1["tol"], z["c1_tol"])
    assert np.array_equal(c1["w"].astype(np.float32), z["c1_w"]) and np.array_equal(c1["h"].astype(np.float32), z["c1_h"])
    r = RRng(123)
    w_init = [r.matrix_runif(30, A.shape[0]) for _ in range(3)]
    seeds = [abs(r.dot_random_seed(3 + rep)) for rep in (1, 2, 3)]
    assert [int(s) for s in z["cv_seeds"]] == seeds
    cv = oracle.ard_nmf(A, At, w_init[0][:5, :], seeds[0], 20, tol=1e-4, maxit=100, L1=0.01, L2=0.0, overfit_threshold=1e-4,
                        trace_test_mse=5)
    assert np.array_equal(cv["test_mse"], z["cv_test_mse"]) and np.array_equal(cv["iter"], z["cv_iter"])
    assert np.array_equal(cv["d"], z["cv_d"])

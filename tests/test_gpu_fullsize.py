"""Size-independent properties at the FULL BASELINE size (30k genes x 1M cells, 5 %, k = 32), where the oracle
cannot run: linearity of the right-hand-side product, a checksum of checksums between the two orientations,
row-normalisation after `scale`, and consistency of the speckled mask between orientations."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

M, N, DENS, K = 30000, 1000000, 0.05, 32


@pytest.fixture(scope="module")
def big():
    from singlet_b200 import synth
    from singlet_b200.sharded import CudaBackend

    be = CudaBackend(0)
    tab = synth.values_table(M, DENS)
    A = be.synth(M, N, DENS, synth.DATA_SEED, 0, 0, N, tab)
    At = be.synth(M, N, DENS, synth.DATA_SEED, 1, 0, M, tab)
    yield be, A, At
    be.close()


def test_generator_statistics(big):
    be, A, At = big
    (ra, ca, nnz_a), (rt, ct, nnz_t) = be.matrix_info(A), be.matrix_info(At)
    assert (ra, ca) == (M, N) and (rt, ct) == (N, M) and nnz_a == nnz_t
    assert abs(nnz_a / (M * N) - DENS) < 1e-4 * DENS * 10
    cnt = be.column_counts(A)
    assert abs(float(cnt.mean()) - DENS * M) < 1.0 and float(cnt.std()) < 40  # ~Binomial(3000, 0.5)


@pytest.mark.parametrize("precision", ["mixed16", "fp32"])
def test_rhs_linearity_and_checksum_of_checksums(big, precision):
    """Both operand precisions of the sparse product (sgl_set_precision). FP32 operands: linear up to FP32 rounding.
    16-bit staged operands (the default): every operand of the three products is rounded to FP16 separately (relative
    2^-12 per element, zero-mean), so linearity holds to a few 1e-5 of the largest entry."""
    from singlet_b200 import _lib

    be, A, At = big
    _lib.check(be.lib.sgl_set_precision(be._h, {"mixed16": _lib.PRECISION_MIXED16, "fp32": _lib.PRECISION_FP32}[precision]))
    g = torch.Generator(device="cpu").manual_seed(1)
    kp = be.kp(K)
    F1 = torch.rand((M, kp), generator=g).to(be.device)
    F2 = torch.rand((M, kp), generator=g).to(be.device)
    B1, B2, B12 = be.zeros_factor(N, K), be.zeros_factor(N, K), be.zeros_factor(N, K)
    be.rhs(A, F1, K, B1)
    be.rhs(A, F2, K, B2)
    be.rhs(A, F1 + F2, K, B12)
    err = (B12 - (B1 + B2)).abs().max() / B12.abs().max()
    assert float(err) < (2e-6 if precision == "fp32" else 2e-4), float(err)
    # checksum of checksums: column sums of A via rhs(A, 1) and row sums via rhs(At, 1) add up to the same total
    ones_m, ones_n = torch.ones((M, kp), device=be.device), torch.ones((N, kp), device=be.device)
    colsum, rowsum = be.zeros_factor(N, K), be.zeros_factor(M, K)
    be.rhs(A, ones_m, K, colsum)
    be.rhs(At, ones_n, K, rowsum)
    tot_c, tot_r = float(colsum[:, 0].double().sum()), float(rowsum[:, 0].double().sum())
    assert abs(tot_c - tot_r) <= 1e-6 * tot_c
    # and W . A summed over cells equals W . rowsum(A): one factor at a time
    w = torch.rand(M, generator=g, dtype=torch.float64).to(be.device)
    Fw = torch.zeros((M, kp), device=be.device)
    Fw[:, 0] = w.float()
    Bw = be.zeros_factor(N, K)
    be.rhs(A, Fw, K, Bw)
    lhs = float(Bw[:, 0].double().sum())
    rhs = float((Fw[:, 0].double() * rowsum[:, 0].double()).sum())
    # FP32 accumulation of 50k terms per gene drawn from an 8-value table rounds with a systematic (not random-walk)
    # bias of ~1e-6 relative, so this identity holds to 1e-5 rather than to a few ulp
    assert abs(lhs - rhs) <= 1e-5 * abs(rhs)
    _lib.check(be.lib.sgl_set_precision(be._h, _lib.PRECISION_MIXED16))


def test_one_iteration_invariants(big):
    from singlet_b200 import synth
    from singlet_b200.sharded import ShardedNMF

    be, A, At = big
    fit = ShardedNMF(be, M, N, K, A, At, layout="B")
    fit.set_w(synth.w_init(K, M))
    tol1 = fit.iteration(0.01, 0.01, 0.0, 0.0)
    tol2 = fit.iteration(0.01, 0.01, 0.0, 0.0)
    assert 0.0 < tol2 < tol1 <= 1.0 + 1e-6
    # after `scale` every factor row sums to one (src/singlet.cpp:219-225) and the factors are non-negative
    assert float((fit.W[:M, :K].double().sum(dim=0) - 1).abs().max()) < 1e-5
    assert float((fit.H[:N, :K].double().sum(dim=0) - 1).abs().max()) < 1e-4
    assert float(fit.W.min()) >= 0.0 and float(fit.H.min()) >= 0.0
    assert bool(torch.isfinite(fit.W).all()) and bool(torch.isfinite(fit.H).all())
    assert float(fit.d[:K].min()) > 0.0


def test_device_transpose_at_full_size(big):
    """sgl_matrix_transpose of the 1.5 G non-zero matrix: the transpose built on the device is indistinguishable from the
    independently GENERATED At -- same column counts, and bit-identical right-hand-side products for a random factor
    (every column sum is formed in row order, so any misplaced, missing or reordered record would change bits) -- and
    transposing twice gives back A in the same sense."""
    be, A, At = big
    T = be.transpose(A)
    assert be.matrix_info(T) == be.matrix_info(At)
    assert torch.equal(be.column_counts(T), be.column_counts(At))
    g = torch.Generator(device="cpu").manual_seed(7)
    kp = be.kp(K)
    F = torch.rand((N, kp), generator=g).to(be.device)
    B1, B2 = be.zeros_factor(M, K), be.zeros_factor(M, K)
    be.rhs(At, F, K, B1)
    be.rhs(T, F, K, B2)
    assert torch.equal(B1, B2)
    TT = be.transpose(T)
    assert be.matrix_info(TT) == be.matrix_info(A) and torch.equal(be.column_counts(TT), be.column_counts(A))
    Fm = torch.rand((M, kp), generator=g).to(be.device)
    C1, C2 = be.zeros_factor(N, K), be.zeros_factor(N, K)
    be.rhs(A, Fm, K, C1)
    be.rhs(TT, Fm, K, C2)
    assert torch.equal(C1, C2)
    be.free_matrix(T)
    be.free_matrix(TT)


def test_mask_consistent_between_orientations(big):
    be, A, At = big
    mA = be.mask_build(A, 123, 20, 0, 0, 0)
    mAt = be.mask_build(At, 123, 20, 1, 0, 0)
    a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    be.lib.sgl_mask_info(mA, C.byref(a), C.byref(b))
    be.lib.sgl_mask_info(mAt, C.byref(c), C.byref(d))
    assert a.value == c.value and b.value == d.value  # same held-out set seen from cells and from genes
    assert abs(a.value / (M * N) - 0.05) < 2e-4 and abs(b.value / be.matrix_info(A)[2] - 0.05) < 5e-4


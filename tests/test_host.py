"""Host-side logic that needs no GPU: R RNG restatement, fixture reader, synthetic generator rules,
the C-ABI library's symbol table, the rank-selection controller."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_r_rng_matches_r():
    """set.seed(123); runif(5) and .Random.seed[3] are R's well-known values."""
    from singlet_b200.rrng import RRng

    r = RRng(123)
    assert r.dot_random_seed(1) == 10403 and r.dot_random_seed(2) == 624
    assert r.dot_random_seed(3) == -983674937
    assert np.allclose(r.runif(5), [0.28757752, 0.78830514, 0.40897692, 0.88301740, 0.94046728], atol=5e-9)
    m = RRng(42).matrix_runif(3, 4)
    assert m.shape == (3, 4) and np.allclose(m.ravel(order="F"), RRng(42).runif(12))  # column-major fill
    a = RRng(7)
    a.runif(700)  # crosses one 624-word regeneration
    assert a.dot_random_seed(2) == 700 - 624


def test_pbmc3k_fixture():
    """Shape and summary statistics of the bundled dataset."""
    from singlet_b200.datasets import get_pbmc3k_data, log_normalize

    A = get_pbmc3k_data()
    assert A.shape == (13714, 2700) and A.nnz == 2282976
    assert A.data.max() == 419 and abs(A.data.mean() - 2.797) < 1e-3
    assert A.has_sorted_indices and np.diff(A.indptr).min() == 212 and np.diff(A.indptr).max() == 3400
    N = log_normalize(A)
    assert abs(N.data.mean() - 2.026) < 1e-3
    src = "/root/reference/data/pbmc3k.RData"
    if os.path.exists(src):
        B = get_pbmc3k_data(src)
        assert (A != B).nnz == 0


def test_synth_rules():
    from singlet_b200 import synth

    S, q32, table = synth.spec(30000, 0.05)
    assert S == 10 and q32 == round(0.5 * 2**32) and table.dtype == np.float32 and np.all(np.diff(table) > 0)
    A = synth.synth_scipy(2000, 1500, 0.05)
    assert abs(A.nnz / (2000 * 1500) - 0.05) < 0.002 and A.has_sorted_indices
    B = synth.synth_csc(2000, 1500, 0.05, col0=100, ncol=50)
    S_ = A[:, 100:150]
    assert np.array_equal(B[0], S_.indptr) and np.array_equal(B[1], S_.indices) and np.array_equal(B[2], S_.data)
    w = synth.w_init(4, 10)
    assert w.shape == (4, 10) and w.min() > 0 and w.max() < 1 and w.flags.f_contiguous
    assert synth.algorithmic_bytes_per_iter(30000, 10**6, 1_500_000_000, 32) == 16 * 1_500_000_000 + 4 * 1030002 + 16 * 32 * 1030000


def test_c_abi_library_exports_every_declared_symbol():
    """libsinglet_cuda.so loads without a GPU and exports exactly what include/singlet_cuda.h declares."""
    from singlet_b200 import _lib

    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "singlet_cuda.h")).read()
    declared = set(re.findall(r"\b(sgl_[a-z0-9_]+)\s*\(", hdr))
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.sgl_version() >= 100
    assert [lib.sgl_padded_rank(k) for k in (1, 4, 5, 10, 32, 33, 100, 128, 129)] == [4, 4, 8, 16, 32, 64, 128, 128, -1]
    s = np.array([3.0, 6.0, 8.0, 5.0, 14.0])  # x = (0,1,2), y = (1,2,3): sum x, sum y, sum xy, sum x^2, sum y^2
    assert abs(lib.sgl_cor_from_sums(s.ctypes.data, 3.0)) < 1e-12


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path fails loudly instead of falling back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from singlet_b200 import SingletCudaError, api, synth

    A = synth.synth_scipy(50, 40, 0.2)
    with pytest.raises(SingletCudaError) as e:
        api.c_nmf(A, A.T.tocsc(), 1e-4, 2, False, 0, 0, 0, 0, 0, synth.w_init(2, 50), handle=api.Handle(0))
    assert e.value.code == -2


def test_get_best_rank():
    """R/GetBestRank.R:8-46 on hand-made CV tables."""
    import pandas as pd

    from singlet_b200.api import GetBestRank

    rows = []
    for rep in (1, 2):
        for k, errs in ((2, [0.20, 0.15, 0.14]), (4, [0.18, 0.13, 0.12]), (8, [0.17, 0.12, 0.125])):
            for it, e in enumerate(errs):
                rows.append({"k": k, "rep": rep, "test_error": e, "iter": 5 * it, "tol": 1e-3})
    df = pd.DataFrame(rows)
    assert GetBestRank(df, 1e-4) == 4  # k = 8 overfits (error rises), so it and above are excluded
    assert GetBestRank(df[df["k"] < 8], 1e-4) == 4
    assert GetBestRank(df, 0.5) == 4  # loose threshold: nothing counts as overfit, lowest final error (k = 4) wins
    df2 = df.copy()
    df2.loc[(df2["k"] == 8) & (df2["iter"] == 10), "test_error"] = 0.11  # k = 8 keeps improving
    assert GetBestRank(df2, 1e-4) == 8


def test_distributed_transpose_blocks():
    """R/cross_validate_nmf.R:37-50: gene-block transposes of a column-chunk list."""
    from singlet_b200 import api, synth

    A = synth.synth_scipy(103, 60, 0.2)
    chunks = [A[:, :20].tocsc(), A[:, 20:45].tocsc(), A[:, 45:].tocsc()]
    blocks = api._distributed_transpose(chunks)
    assert len(blocks) == 3 and all(b.shape[0] == 60 for b in blocks) and sum(b.shape[1] for b in blocks) == 103
    import scipy.sparse as sp

    assert (sp.hstack(blocks).tocsc() != A.T.tocsc()).nnz == 0


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors of the C-ABI structs (sgl_csc, sgl_callbacks, sgl_trace, sgl_fit_job) have the size and field
    offsets a C compiler gives the declarations in include/singlet_cuda.h (the header is compiled as plain C here)."""
    import ctypes as C
    import subprocess

    from singlet_b200 import _lib

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    structs = {"sgl_csc": (_lib.Csc, ["nrow", "ncol", "p", "i", "x"]),
               "sgl_callbacks": (_lib.Callbacks, ["user", "poll_interrupt", "on_iter"]),
               "sgl_trace": (_lib.Trace, ["test_mse", "iter", "tol", "score_overfit", "capacity", "length"]),
               "sgl_fit_job": (_lib.FitJob, ["k", "status", "seed", "w", "d", "h", "trace"])}
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "singlet_cuda.h"', "int main(void) {"]
    for name, (_, fields) in structs.items():
        src.append(f'  printf("{name} %zu", sizeof({name}));')
        for f in fields:
            src.append(f'  printf(" %zu", offsetof({name}, {f}));')
        src.append('  printf("\\n");')
    src += ["  return 0;", "}"]
    c_file, exe = tmp_path / "abi.c", tmp_path / "abi"
    c_file.write_text("\n".join(src))
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(c_file), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = {}
    for line in out:
        if line.strip():
            parts = line.split()
            seen[parts[0]] = [int(v) for v in parts[1:]]
    for name, (cls, fields) in structs.items():
        exp = [C.sizeof(cls)] + [getattr(cls, f).offset for f in fields]
        assert seen[name] == exp, (name, seen[name], exp)


def test_rcpp_glue_type_checks_against_the_c_header():
    """rglue/singlet_cuda_glue.cpp (the bodies a maintainer drops into the R package) cannot be built here (no R), but it
    is type-checked against include/singlet_cuda.h with minimal stand-ins for Rcpp / RcppEigen (tests/rglue_stub): every
    sgl_* call in the glue matches the declared signature."""
    import subprocess

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(root, "tests", "rglue_stub"), "-I",
                        os.path.join(root, "include"), os.path.join(root, "rglue", "singlet_cuda_glue.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cross_validate_nmf_host_logic_batch_equals_fit_by_fit(monkeypatch):
    """Host side of cross_validate_nmf (reference R/cross_validate_nmf.R:57-105) without a GPU: with the fits replaced by
    the CPU oracle, the batched grid (one c_ard_nmf_batch call) and the fit-by-fit loop draw the same w_init slices and
    mask seeds from R's RNG stream and build the same data frame, in expand.grid order."""
    from oracle.pyoracle import Oracle
    from singlet_b200 import api, synth

    orc = Oracle("port")
    seen = {"batch": [], "loop": []}

    def fake_fit(A, At, tol, maxit, verbose, L1, L2, threads, w, seed, inv_density, thr, trace, handle=None):
        seen["loop"].append((w.shape[0], int(seed)))
        return orc.ard_nmf(A, A.T.tocsc(), w, seed, inv_density, tol=tol, maxit=maxit, L1=L1, L2=L2, overfit_threshold=thr, trace_test_mse=trace)

    def fake_batch(A, At, tol, maxit, L1, L2, threads, ws, seeds, inv_density, thr, trace, concurrency=0, handle=None):
        seen["batch"] += [(w.shape[0], int(s)) for w, s in zip(ws, seeds)]
        return [orc.ard_nmf(A, A.T.tocsc(), w, s, inv_density, tol=tol, maxit=maxit, L1=L1, L2=L2, overfit_threshold=thr, trace_test_mse=trace)
                for w, s in zip(ws, seeds)]

    monkeypatch.setattr(api, "c_ard_nmf", fake_fit)
    monkeypatch.setattr(api, "c_ard_nmf_batch", fake_batch)
    A = synth.synth_scipy(50, 40, 0.3, seed=4)
    api.set_seed(11)
    df_b = api.cross_validate_nmf(A, [2, 4, 3], n_replicates=2, maxit=5, verbose=0, batch=True)
    api.set_seed(11)
    df_l = api.cross_validate_nmf(A, [2, 4, 3], n_replicates=2, maxit=5, verbose=0, batch=False)
    assert df_b.equals(df_l) and len(df_b) > 0
    assert seen["batch"] == seen["loop"] and [k for k, _ in seen["loop"]] == [2, 4, 3, 2, 4, 3]
    assert len({s for _, s in seen["loop"][:3]}) == 1 and seen["loop"][0][1] != seen["loop"][3][1]  # one mask seed per replicate
    assert list(df_b["k"].unique()) == [2, 4, 3] and list(df_b["rep"].unique()) == [1, 2]


def test_ard_nmf_rank_search_host_logic(monkeypatch):
    """Host side of ard_nmf (reference R/ard_nmf.R:95-178) without a GPU, fits replaced by the CPU oracle on a matrix with
    three planted factors: the search starts at k_init, never leaves [k_min, k_max], never repeats a rank within a
    replicate, uses the mask seed test_seed + rep, and the final unmasked fit runs at GetBestRank of the table with the
    first best_rank rows of the first w_init."""
    import scipy.sparse as sp

    from oracle.pyoracle import Oracle
    from singlet_b200 import api

    orc = Oracle("port")
    rs = np.random.RandomState(0)
    Wt = rs.gamma(1.0, 1.0, size=(80, 3)) * (rs.uniform(size=(80, 3)) < 0.4)
    Ht = rs.gamma(1.0, 1.0, size=(3, 70)) * (rs.uniform(size=(3, 70)) < 0.5)
    D = Wt @ Ht + 0.01 * rs.uniform(size=(80, 70)) * (rs.uniform(size=(80, 70)) < 0.3)
    A = sp.csc_matrix(D)
    calls, final = [], {}

    def fake_fit(A_, At, tol, maxit, verbose, L1, L2, threads, w, seed, inv_density, thr, trace, handle=None):
        calls.append((w.shape[0], int(seed)))
        return orc.ard_nmf(A_, A_.T.tocsc(), w, seed, inv_density, tol=tol, maxit=maxit, L1=L1, L2=L2, overfit_threshold=thr, trace_test_mse=trace)

    def fake_nmf(A_, At, tol, maxit, verbose, L1_w, L1_h, L2_w, L2_h, threads, w, handle=None):
        final["k"], final["w0"] = w.shape[0], np.array(w)
        return orc.nmf(A_, A_.T.tocsc(), w, tol=tol, maxit=maxit, L1=(L1_w, L1_h), L2=(L2_w, L2_h))

    monkeypatch.setattr(api, "c_ard_nmf", fake_fit)
    monkeypatch.setattr(api, "c_nmf", fake_nmf)
    api.set_seed(5)
    model = api.ard_nmf(A, k_init=2, k_max=12, n_replicates=2, maxit=30, verbose=0, tol_overfit=0.3)
    from singlet_b200.rrng import RRng

    r = RRng(0)
    r.set_seed(5)
    w_init = [r.matrix_runif(12, 80) for _ in range(2)]
    test_seed = abs(r.dot_random_seed(3))
    reps = [[k for k, s in calls if s == test_seed + rep] for rep in (1, 2)]
    assert sum(len(x) for x in reps) == len(calls)            # every fit used test_seed + rep
    for ks in reps:
        assert ks[0] == 2 and len(set(ks)) == len(ks) and all(2 <= k <= 12 for k in ks)
        assert len(ks) == 1 or ks[1] == 4                    # the step doubles while the largest rank is the best one
    df = model["cv_data"]
    assert list(df.columns) == ["k", "rep", "test_error", "iter", "tol", "overfit_score"]
    best = api.GetBestRank(df, 0.3)
    assert final["k"] == best == model["w"].shape[1] and np.array_equal(final["w0"], w_init[0][:best, :])
    assert 2 <= best <= 12 and np.all(np.diff(model["d"]) <= 0)
    assert len(calls) >= 3                                   # the search did move


def test_run_nmf_host_logic(monkeypatch):
    """Host side of run_nmf (reference R/run_nmf.R:39-75) without a GPU, the fit
    replaced by the CPU oracle: w_init is the first rank * m draws of runif after set.seed (k x m, column-major fill), the
    model comes back sorted by decreasing d with w transposed to m x k, and L1 / L2 pairs are passed through as (w, h)."""
    from oracle.pyoracle import Oracle
    from singlet_b200 import api, synth
    from singlet_b200.rrng import RRng

    orc = Oracle("port")
    got = {}

    def fake_nmf(A_, At, tol, maxit, verbose, L1_w, L1_h, L2_w, L2_h, threads, w, handle=None):
        got.update(w0=np.array(w), L1=(L1_w, L1_h), L2=(L2_w, L2_h), At=At)
        return orc.nmf(A_, A_.T.tocsc(), w, tol=tol, maxit=maxit, L1=(L1_w, L1_h), L2=(L2_w, L2_h))

    monkeypatch.setattr(api, "c_nmf", fake_nmf)
    A = synth.synth_scipy(60, 50, 0.3, seed=2)
    api.set_seed(99)
    model = api.run_nmf(A, 4, maxit=8, verbose=False, L1=(0.02, 0.03), L2=0.1)
    assert np.array_equal(got["w0"], RRng(99).matrix_runif(4, 60)) and got["L1"] == (0.02, 0.03) and got["L2"] == (0.1, 0.1)
    assert got["At"] is None                                  # the transpose is left to the device by default
    assert model["w"].shape == (60, 4) and model["h"].shape == (4, 50) and np.all(np.diff(model["d"]) <= 0)
    ref = orc.nmf(A, A.T.tocsc(), got["w0"], tol=1e-4, maxit=8, L1=(0.02, 0.03), L2=(0.1, 0.1))
    order = np.argsort(-ref["d"], kind="stable")
    assert np.allclose(model["w"], ref["w"][order].T) and np.allclose(model["h"], ref["h"][order]) and np.allclose(model["d"], ref["d"][order])
    api.set_seed(99)
    api.run_nmf(A, 4, maxit=2, verbose=False, device_transpose=False)
    assert got["At"] is not None and (got["At"] != A.T.tocsc()).nnz == 0


def test_sass_of_the_shipped_library():
    """The shipped library is sm_100a code and carries the instructions DESIGN.md claims (B200_PROFILING.md, "SASS mnemonics"):
    bulk TMA copies with mbarrier synchronisation for the factor tiles of the sparse product (UBLKCP, SYNCS), cp.async rings
    (LDGSTS), the mixed-precision FMA of the 16-bit-operand product (FHFMA), and the tensor-core Gram corrections of the masked
    solver (HMMA.16816 on BF16 and FP16 operands fed by ldmatrix = LDSM)."""
    import shutil
    import subprocess

    from singlet_b200 import _lib

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run([tool, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    for mnemonic in ("UBLKCP", "SYNCS", "LDGSTS", "FHFMA", "LDSM", "HMMA.16816.F32.BF16", "HMMA.16816.F32 "):
        assert mnemonic in sass, mnemonic

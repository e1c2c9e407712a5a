"""The N > 1 path on CPU: two gloo ranks run singlet_b200.sharded.ShardedNMF with a TEST-ONLY backend
whose compute calls go to the CPU oracle. This checks the sharding logic -- shard bounds, global
hash offsets, which quantities are all-reduced / all-gathered and when -- against the unsharded
oracle fit. (The product backend is CudaBackend; nothing in singlet_b200/ can reach this one.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Same interface as singlet_b200.sharded.CudaBackend, FP64 on CPU tensors, solves via the oracle."""

    def __init__(self):
        from oracle.pyoracle import Oracle

        self.orc = Oracle("port")

    def kp(self, k):
        kp = 4
        while kp < k:
            kp <<= 1
        return kp

    def zeros_factor(self, cols, k):
        return torch.zeros((max(cols, 1), self.kp(k)), dtype=torch.float64)

    def zeros_f64(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def factor_from_host(self, host_kxc, out):
        k, cols = host_kxc.shape
        out[:cols, :k] = torch.from_numpy(np.ascontiguousarray(host_kxc.T))

    def factor_to_host(self, dev, k, cols):
        return np.asfortranarray(dev[:cols, :k].numpy().T)

    def upload(self, A):
        return A.tocsc()

    def transpose(self, X):
        T = X.T.tocsc()
        T.sort_indices()
        return T

    def mask_build(self, X, seed, inv_density, mask_t, col_offset, row_offset):
        return (seed, inv_density, mask_t, col_offset, row_offset)

    def gram(self, F, k, cols, out, jitter):
        kp = self.kp(k)
        g = np.zeros((kp, kp))
        f = F[:cols, :k].numpy()
        g[:k, :k] = f.T @ f
        if jitter:
            g[np.arange(k), np.arange(k)] += 1e-15
        out.copy_(torch.from_numpy(g.ravel()))

    def gram_jitter(self, k, gram):
        kp = self.kp(k)
        g = gram.view(kp, kp)
        for f in range(k):
            g[f, f] += 1e-15

    def update(self, X, mask, F_in, F_out, k, gram, L1, L2, rowsum):
        kp = self.kp(k)
        a = gram.view(kp, kp)[:k, :k].numpy().copy()
        Fi = F_in[:, :k].numpy()
        for c in range(X.shape[1]):
            lo, hi = X.indptr[c], X.indptr[c + 1]
            if lo == hi:
                continue
            rows, vals = X.indices[lo:hi], X.data[lo:hi]
            ai = a
            if mask is not None:
                seed, inv, mask_t, coff, roff = mask
                gc = c + coff
                held = np.array([self.orc.draw(seed, r + roff, gc, inv) if mask_t else self.orc.draw(seed, gc, r + roff, inv)
                                 for r in range(X.shape[0])], dtype=bool)
                keep = ~held[rows]
                rows, vals = rows[keep], vals[keep]
                fm = Fi[np.nonzero(held)[0]]
                ai = a - (fm.T @ fm + 1e-15 * np.eye(k))
            b = Fi[rows].T @ vals if len(rows) else np.zeros(k)
            x, _, _ = self.orc.nnls(ai, b, F_out[c, :k].numpy(), L1, L2)
            F_out[c, :k] = torch.from_numpy(x)
        rs = torch.zeros(kp, dtype=torch.float64)
        rs[:k] = F_out[: X.shape[1], :k].sum(dim=0)
        rowsum.copy_(rs)

    def column_counts(self, X):
        return torch.from_numpy(np.diff(X.indptr).astype(np.float64))

    def colptr_like(self, counts):
        cp = torch.zeros(counts.numel() + 1, dtype=torch.int64)
        cp[1:] = torch.cumsum((counts > 0).to(torch.int64), dim=0)
        return cp

    def rhs(self, X, F_in, k, B_out):
        B_out[: X.shape[1], :k] = torch.from_numpy((X.T @ F_in[: X.shape[0], :k].numpy()))

    def solve(self, B, colptr_like, ncol, F_out, k, gram, L1, L2, rowsum):
        kp = self.kp(k)
        a = gram.view(kp, kp)[:k, :k].numpy().copy()
        for c in range(ncol):
            if colptr_like[c] == colptr_like[c + 1]:
                continue
            x, _, _ = self.orc.nnls(a, B[c, :k].numpy(), F_out[c, :k].numpy(), L1, L2)
            F_out[c, :k] = torch.from_numpy(x)
        rs = torch.zeros(kp, dtype=torch.float64)
        rs[:k] = F_out[:ncol, :k].sum(dim=0)
        rowsum.copy_(rs)

    def finish_d(self, k, d):
        d[:k] += 1e-15
        d[k:] = 1.0

    def scale(self, F, k, cols, d):
        F[:cols, :k] /= d[:k]

    def finish_d_rescale_gram(self, k, d, gram):
        kp = self.kp(k)
        self.finish_d(k, d)
        g = gram.view(kp, kp)
        g[:k, :k] /= torch.outer(d[:k], d[:k])
        self.gram_jitter(k, gram)

    def cor_sums(self, X, Y, k, cols, out):
        x, y = X[:cols, :k].numpy().ravel(), Y[:cols, :k].numpy().ravel()
        out[:5] = torch.tensor([x.sum(), y.sum(), (x * y).sum(), (x * x).sum(), (y * y).sum()], dtype=torch.float64)

    def cor_from_sums(self, s, n):
        return float(1 - (n * s[2] - s[0] * s[1]) / np.sqrt((n * s[3] - s[0] ** 2) * (n * s[4] - s[1] ** 2)))

    def mse(self, A, mask, W, d, H, k, which, out):
        seed, inv, mask_t, coff, roff = mask
        wd = W[: A.shape[0], :k].numpy() * d[:k].numpy()
        tot = 0.0
        for c in range(A.shape[1]):
            held = np.array([self.orc.draw(seed, c + coff, g, inv) for g in range(A.shape[0])], dtype=bool)
            if held.any():
                col = np.asarray(A[:, c].todense()).ravel()
                res = wd[held] @ H[c, :k].numpy() - col[held]
                tot += float((res ** 2).mean())
        out[0] = tot


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from singlet_b200 import sharded, synth
        from singlet_b200.sharded import shard_bounds, sharded_ard_nmf, sharded_nmf

        m, n, k = 41, 37, 3  # ragged on purpose: neither divides by 2
        A = synth.synth_scipy(m, n, 0.3, seed=5)
        At = A.T.tocsc()
        w0 = synth.w_init(k, m, seed=2)
        c0, c1, _ = shard_bounds(n, world, rank)
        g0, g1, _ = shard_bounds(m, world, rank)
        be = OracleBackend()
        res = sharded_nmf(be, m, n, k, A[:, c0:c1].tocsc(), At[:, g0:g1].tocsc(), w0, tol=0.0, maxit=4, L1=(0.01, 0.02),
                          rank=rank, world=world)
        cv = sharded_ard_nmf(be, m, n, k, A[:, c0:c1].tocsc(), At[:, g0:g1].tocsc(), w0, 123, 5, tol=0.0, maxit=3,
                             overfit_threshold=10.0, trace_test_mse=2, rank=rank, world=world)
        # layout B: the rank's transpose block covers its own cells only
        # (rank 0 hands it over, rank 1 lets the driver derive it from its cell block: At_shard = None)
        resb = sharded_nmf(be, m, n, k, A[:, c0:c1].tocsc(), A[:, c0:c1].T.tocsc() if rank == 0 else None, w0, tol=0.0, maxit=4,
                           L1=(0.01, 0.02), rank=rank, world=world, layout="B")
        # layout B with three exchanges per iteration (the scheme of csrc/multi.cu on more than one rank)
        fit3 = sharded.ShardedNMF(be, m, n, k, A[:, c0:c1].tocsc(), None, rank, world, None, layout="B3")
        fit3.set_w(w0)
        tol3 = [fit3.iteration(0.01, 0.02, 0.0, 0.0) for _ in range(4)]
        per_iter = fit3.n_collectives / 4
        w3, d3, h3 = fit3.factors_to_host()
        q.put((rank, res["w"], res["h"], res["d"], res["tol"], cv["test_mse"], cv["h"], resb["w"], resb["h"], resb["d"],
               w3, h3, d3, tol3[-1], per_iter))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_unsharded_oracle(oracle):
    from singlet_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, n, k = 41, 37, 3
    A = synth.synth_scipy(m, n, 0.3, seed=5)
    At = A.T.tocsc()
    w0 = synth.w_init(k, m, seed=2)
    ref = oracle.nmf(A, At, w0, tol=0.0, maxit=4, L1=(0.01, 0.02))
    cvr = oracle.ard_nmf(A, At, w0, 123, 5, tol=0.0, maxit=3, overfit_threshold=10.0, trace_test_mse=2)
    for rank, w, h, d, tol, mse, hcv, wb, hb, db, w3, h3, d3, tol3, per_iter in outs:  # every rank ends with the full replicated model
        # three exchanges per iteration (+ the one-off all-reduce of the gene counts at construction, 1/4 per iteration here): same fit
        assert per_iter == 3.25
        assert np.allclose(w3, ref["w"], rtol=1e-9, atol=1e-12) and np.allclose(h3, ref["h"], rtol=1e-9, atol=1e-12)
        assert np.allclose(d3, ref["d"], rtol=1e-9) and abs(tol3 - ref["tol"][-1]) < 1e-9
        assert np.allclose(w, ref["w"], rtol=1e-9, atol=1e-12) and np.allclose(h, ref["h"], rtol=1e-9, atol=1e-12)
        assert np.allclose(d, ref["d"], rtol=1e-9) and abs(tol - ref["tol"][-1]) < 1e-9
        assert np.allclose(mse, cvr["test_mse"], rtol=1e-9) and np.allclose(hcv, cvr["h"], rtol=1e-8, atol=1e-12)
        assert np.allclose(wb, ref["w"], rtol=1e-9, atol=1e-12) and np.allclose(hb, ref["h"], rtol=1e-9, atol=1e-12)
        assert np.allclose(db, ref["d"], rtol=1e-9)


def test_shard_bounds_cover_everything():
    from singlet_b200.sharded import shard_bounds

    for total in (1, 7, 8, 30000, 1000003):
        for world in (1, 2, 3, 8):
            seen = 0
            for r in range(world):
                lo, hi, per = shard_bounds(total, world, r)
                assert lo == min(r * per, total) and lo <= hi <= total
                seen += hi - lo
            assert seen == total


def _fake_batch(orc):
    """Stand-in for api.c_ard_nmf_batch on a machine without a GPU: the oracle, fit by fit."""
    def batch(A, At, tol, maxit, L1, L2, threads, ws, seeds, inv_density, overfit_threshold, trace_test_mse, concurrency=0, handle=None):
        At = A.T.tocsc()
        return [orc.ard_nmf(A, At, w, s, inv_density, tol=tol, maxit=maxit, L1=L1, L2=L2, overfit_threshold=overfit_threshold,
                            trace_test_mse=trace_test_mse) for w, s in zip(ws, seeds)]
    return batch


def _cv_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        from singlet_b200 import api, sharded, synth

        class NoHandle:
            def close(self):
                pass

        calls = []
        fake = _fake_batch(Oracle("port"))

        def counting(*a, **kw):
            calls.append(len(a[7]))
            return fake(*a, **kw)

        api.c_ard_nmf_batch, api.Handle = counting, (lambda *a, **kw: NoHandle())
        A = synth.synth_scipy(60, 45, 0.25, seed=9)
        df = sharded.distributed_cross_validate_nmf(A, [2, 3, 5, 4], n_replicates=2, maxit=6, trace_test_mse=2, seed=77, device=0)
        q.put((rank, df.to_dict("list"), calls))
    finally:
        dist.destroy_process_group()


def test_cv_sweep_dealt_over_two_ranks_equals_the_sequential_sweep(oracle):
    """distributed_cross_validate_nmf: the (rank, replicate) grid is dealt over the processes and gathered; every process
    ends with the data frame of the sequential sweep (R/cross_validate_nmf.R:69-97), rows in expand.grid order."""
    from singlet_b200 import api, synth
    from singlet_b200.rrng import RRng

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_cv_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # sequential expectation with the same R RNG stream
    A = synth.synth_scipy(60, 45, 0.25, seed=9)
    At = A.T.tocsc()
    r = RRng(0)
    r.set_seed(77)
    ks = [2, 3, 5, 4]
    w_init = [r.matrix_runif(max(ks), 60) for _ in range(2)]
    exp = {"k": [], "rep": [], "test_error": [], "iter": [], "tol": []}
    for rep in (1, 2):
        for k in ks:
            mod = oracle.ard_nmf(A, At, w_init[rep - 1][:k, :], abs(r.dot_random_seed(3 + rep)), 20, tol=1e-4, maxit=6, L1=0.01, L2=0.0,
                                 overfit_threshold=1e-4, trace_test_mse=2)
            for t in range(len(mod["test_mse"])):
                exp["k"].append(k); exp["rep"].append(rep); exp["test_error"].append(float(mod["test_mse"][t]))
                exp["iter"].append(int(mod["iter"][t])); exp["tol"].append(float(mod["tol"][t]))
    assert sorted(outs[0][2] + outs[1][2]) == [4, 4]  # eight fits, four per process
    for _, got, _ in outs:
        assert got["k"] == exp["k"] and got["rep"] == exp["rep"] and got["iter"] == exp["iter"]
        assert np.allclose(got["test_error"], exp["test_error"], rtol=1e-12) and np.allclose(got["tol"], exp["tol"], rtol=1e-12)

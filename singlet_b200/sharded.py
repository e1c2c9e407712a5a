"""One-process-per-GPU ALS NMF: cells sharded for the H update, genes for the W update.

The reference has no distributed code; its only parallelism is an OpenMP loop over columns
(src/singlet.cpp:336-346) and a serial loop over column chunks (:384-402) with a "distributed
transpose" of gene blocks (R/cross_validate_nmf.R:37-50). That chunk layout maps 1:1 onto GPUs
(layout "A"):

* rank r holds the CSC block of its cells (one chunk of ``A_``) and the CSC-of-At block of its genes
  over ALL cells (one block of ``At_``);
* W (k x m) and H (k x n) are replicated between half-iterations by an all-gather of the solved
  shards; the only other exchanges are all-reduces of k (+k^2) doubles: the row sums of ``scale``
  (src/singlet.cpp:220) and the partial Gram H H^T (:200-206). ``cor`` runs redundantly on the
  replicated W so every rank takes the same stopping decision.

Plain (unmasked) fits use the lighter layout "B" of SURVEY.md 8e instead: only the cells are sharded.
Rank r keeps the transpose of ITS OWN cell block (local cells x all genes), computes the partial
W-update right-hand sides from its local H shard, and the k x m partials are reduce-scattered so
that every rank solves one gene shard; W (3.84 MB at the headline config) is then all-gathered.
H is never gathered (128 MB per iteration saved) and the W-update SpMM keeps all m gene columns per
rank, which fills the SMs much better than a 1/N gene shard. Masked (CV) fits stay on layout "A"
because the per-gene Gram corrections would otherwise need a k^2 x m all-reduce. Layout "B3" is layout B with three
exchanges per iteration instead of five -- the scheme csrc/multi.cu uses on more than one rank (``iteration_b3``).

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the plumbing; the compute calls go to
a *backend*: :class:`CudaBackend` (the C ABI device layer) in production. The CPU tests inject their
own oracle-backed backend to check the sharding/collective logic -- nothing in this package ever
falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous shards of ``ceil(total / world)`` items (the last ones may be short or empty)."""
    per = -(-total // world)
    lo = min(rank * per, total)
    return lo, min(lo + per, total), per


class CudaBackend:
    """Device layer of ``libsinglet_cuda.so`` on torch CUDA tensors (device memory + stream only)."""

    def __init__(self, device: int):
        self.lib = _lib.load()
        torch.cuda.set_device(device)
        self.device = torch.device("cuda", device)
        self._h = C.c_void_p()
        # the library creates its stream; torch adopts it as the current stream so that tensor ops,
        # NCCL collectives and CUDA events of this process are ordered with the library's kernels
        _lib.check(self.lib.sgl_create(device, None, C.byref(self._h)))
        self.stream = torch.cuda.ExternalStream(int(self.lib.sgl_stream(self._h)), device=self.device)
        torch.cuda.set_stream(self.stream)
        self._mats, self._masks = [], []

    # -- memory -------------------------------------------------------------------------------
    def kp(self, k):
        return int(self.lib.sgl_padded_rank(k))

    def zeros_factor(self, cols, k):
        return torch.zeros((max(cols, 1), self.kp(k)), dtype=torch.float32, device=self.device)

    def zeros_f64(self, n):
        return torch.zeros(n, dtype=torch.float64, device=self.device)

    def factor_from_host(self, host_kxc, out):
        host = np.ascontiguousarray(np.asfortranarray(host_kxc, dtype=np.float64).T.ravel())  # (c, f) -> c*k + f
        k, cols = host_kxc.shape
        _lib.check(self.lib.sgl_factor_upload(self._h, host.ctypes.data, k, cols, out.data_ptr()))

    def factor_to_host(self, dev, k, cols):
        host = np.zeros(k * cols, dtype=np.float64)
        _lib.check(self.lib.sgl_factor_download(self._h, dev.data_ptr(), k, cols, host.ctypes.data))
        return host.reshape((cols, k)).T.copy(order="F")

    # -- matrices -----------------------------------------------------------------------------
    def upload(self, mats):
        arr, n, keep = _lib.chunks_to_c(mats)
        m = C.c_void_p()
        _lib.check(self.lib.sgl_matrix_upload(self._h, arr, n, C.byref(m)))
        self._mats.append(m)
        return m

    def transpose(self, mat):
        """Device-side transpose of a device matrix (``sgl_matrix_transpose``)."""
        m = C.c_void_p()
        _lib.check(self.lib.sgl_matrix_transpose(self._h, mat, C.byref(m)))
        self._mats.append(m)
        return m

    def free_matrix(self, mat):
        """Release a device matrix now (close() releases whatever is left)."""
        self._mats = [m for m in self._mats if m.value != mat.value]
        _lib.check(self.lib.sgl_matrix_free(self._h, mat))

    def synth(self, m_genes, n_cells, density, seed, orientation, col0, ncol, table):
        table = np.ascontiguousarray(table, dtype=np.float32)
        m = C.c_void_p()
        _lib.check(self.lib.sgl_matrix_synth(self._h, m_genes, n_cells, float(density), int(seed), int(orientation), int(col0),
                                             int(ncol), table.ctypes.data, C.byref(m)))
        self._mats.append(m)
        return m

    def synth_block(self, m_genes, n_cells, density, seed, orientation, col0, ncol, row0, nrows, table):
        table = np.ascontiguousarray(table, dtype=np.float32)
        m = C.c_void_p()
        _lib.check(self.lib.sgl_matrix_synth_block(self._h, m_genes, n_cells, float(density), int(seed), int(orientation),
                                                   int(col0), int(ncol), int(row0), int(nrows), table.ctypes.data, C.byref(m)))
        self._mats.append(m)
        return m

    def column_counts(self, m):
        """Non-zeros per column as a float64 device tensor (for the global empty-column test)."""
        nrow, ncol, nnz = self.matrix_info(m)
        cp = torch.zeros(ncol + 1, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.sgl_matrix_colptr(self._h, m, cp.data_ptr()))
        return (cp[1:] - cp[:-1]).to(torch.float64)

    def colptr_like(self, counts):
        """int64[ncol + 1] with equal consecutive entries exactly where counts == 0."""
        cp = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
        cp[1:] = torch.cumsum((counts > 0).to(torch.int64), dim=0)
        return cp

    def rhs(self, X, F_in, k, B_out):
        _lib.check(self.lib.sgl_dev_rhs(self._h, X, F_in.data_ptr(), k, B_out.data_ptr()))

    def solve(self, B, colptr_like, ncol, F_out, k, gram, L1, L2, rowsum):
        _lib.check(self.lib.sgl_dev_solve(self._h, B.data_ptr(), colptr_like.data_ptr(), int(ncol), F_out.data_ptr(), k,
                                          gram.data_ptr(), float(L1), float(L2), rowsum.data_ptr()))

    def matrix_info(self, m):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.sgl_matrix_info(m, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def matrix_to_host(self, m):
        nrow, ncol, nnz = self.matrix_info(m)
        p, i, x = np.zeros(ncol + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.float64)
        _lib.check(self.lib.sgl_matrix_download(self._h, m, p.ctypes.data, i.ctypes.data, x.ctypes.data))
        return p, i, x, nrow, ncol

    def mask_build(self, X, seed, inv_density, mask_t, col_offset, row_offset):
        m = C.c_void_p()
        _lib.check(self.lib.sgl_mask_build(self._h, X, int(seed), int(inv_density), int(mask_t), int(col_offset),
                                           int(row_offset), C.byref(m)))
        self._masks.append(m)
        return m

    # -- kernels ------------------------------------------------------------------------------
    def gram(self, F, k, cols, out, jitter):
        _lib.check(self.lib.sgl_dev_gram(self._h, F.data_ptr(), k, cols, out.data_ptr(), int(jitter)))

    def gram_jitter(self, k, gram):
        _lib.check(self.lib.sgl_dev_gram_jitter(self._h, k, gram.data_ptr()))

    def update(self, X, mask, F_in, F_out, k, gram, L1, L2, rowsum):
        if mask is None:
            _lib.check(self.lib.sgl_dev_update(self._h, X, F_in.data_ptr(), F_out.data_ptr(), k, gram.data_ptr(), float(L1),
                                               float(L2), rowsum.data_ptr()))
        else:
            _lib.check(self.lib.sgl_dev_update_masked(self._h, X, mask, F_in.data_ptr(), F_out.data_ptr(), k, gram.data_ptr(),
                                                      float(L1), float(L2), rowsum.data_ptr()))

    def finish_d(self, k, d):
        _lib.check(self.lib.sgl_dev_finish_d(self._h, k, d.data_ptr()))

    def scale(self, F, k, cols, d):
        _lib.check(self.lib.sgl_dev_scale(self._h, F.data_ptr(), k, cols, d.data_ptr()))

    def finish_d_rescale_gram(self, k, d, gram):
        _lib.check(self.lib.sgl_dev_finish_d_rescale_gram(self._h, k, d.data_ptr(), gram.data_ptr()))

    def cor_sums(self, X, Y, k, cols, out):
        _lib.check(self.lib.sgl_dev_cor_sums(self._h, X.data_ptr(), Y.data_ptr(), k, cols, out.data_ptr()))

    def cor_from_sums(self, sums_host, n_elems):
        s = np.ascontiguousarray(sums_host, dtype=np.float64)
        return float(self.lib.sgl_cor_from_sums(s.ctypes.data, float(n_elems)))

    def mse(self, A, mask, W, d, H, k, which, out):
        _lib.check(self.lib.sgl_dev_mse(self._h, A, mask, W.data_ptr(), d.data_ptr(), H.data_ptr(), k, int(which),
                                        out.data_ptr()))

    def launch_count(self):
        return int(self.lib.sgl_launch_count(self._h))

    def profile(self, enable: bool):
        _lib.check(self.lib.sgl_profile(self._h, int(enable)))

    def profile_read(self):
        """{kind: (ms, launches, algorithmic bytes)} since the last read (synchronises)."""
        ms, cnt, byt = np.zeros(4), np.zeros(4, np.int64), np.zeros(4, np.int64)
        _lib.check(self.lib.sgl_profile_read(self._h, ms.ctypes.data, cnt.ctypes.data, byt.ctypes.data))
        return {name: (float(ms[q]), int(cnt[q]), int(byt[q])) for q, name in enumerate(("spmm", "nnls", "gram", "other"))}

    def synchronize(self):
        _lib.check(self.lib.sgl_synchronize(self._h))

    def close(self):
        for m in self._masks:
            self.lib.sgl_mask_free(self._h, m)
        for m in self._mats:
            self.lib.sgl_matrix_free(self._h, m)
        self._masks, self._mats = [], []
        if self._h:
            self.lib.sgl_destroy(self._h)
            self._h = C.c_void_p()


class ShardedNMF:
    """State of one sharded fit. ``backend`` supplies the compute; ``group`` the collectives
    (``None`` or world size 1: no collective is issued at all)."""

    def __init__(self, backend, m: int, n: int, k: int, A_shard, At_shard, rank: int = 0, world: int = 1, group=None,
                 mask_A=None, mask_At=None, layout: str = "A"):
        """layout "A": At_shard = this rank's genes over ALL cells. layout "B" (plain fits only): At_shard =
        the transpose of this rank's own cell block (local cells x all genes), or None to build it on the device."""
        self.be, self.m, self.n, self.k = backend, m, n, k
        self.rank, self.world, self.group = rank, world, group
        self.layout = layout
        if layout in ("B", "B3") and (mask_A is not None or mask_At is not None):
            raise ValueError("layout B does not support the masked (CV) solve")
        if At_shard is None:
            # layout B: the transpose of the local cell block is built on the device (row f1), so a rank uploads only A
            if layout not in ("B", "B3") and world > 1:
                raise ValueError("layout A needs the gene block over all cells; only layout B can derive At_shard from A_shard")
            At_shard = backend.transpose(A_shard)
        self.A, self.At, self.mask_A, self.mask_At = A_shard, At_shard, mask_A, mask_At
        self.c0, self.c1, self.c_per = shard_bounds(n, world, rank)
        self.g0, self.g1, self.g_per = shard_bounds(m, world, rank)
        be = backend
        # replicated factors, padded so that rank r's shard starts at r * per (all-gather lands in place)
        self.W = be.zeros_factor(self.g_per * world, k)
        self.H = be.zeros_factor(self.c_per * world, k)
        self.Wprev = be.zeros_factor(self.g_per * world, k)
        kp = be.kp(k)
        self.kp = kp
        # [Gram | d | spare]: one buffer, so that layout "B3" sends the partial Gram and the row sums in ONE all-reduce
        self.red = be.zeros_f64(kp * kp + kp + 1)
        self.gram = self.red[:kp * kp]
        self.d = self.red[kp * kp:kp * kp + kp]
        self.sums = be.zeros_f64(8)
        self.n_collectives = 0
        if layout in ("B", "B3"):
            # right-hand sides of all m genes (padded to g_per * world rows for the reduce-scatter) and the
            # global "gene has any non-zero" test (src/singlet.cpp:340 skips empty columns)
            self.Bw = be.zeros_factor(self.g_per * world, k)
            counts = be.column_counts(self.At)
            self._allreduce(counts)
            self.gene_ptr = be.colptr_like(counts)

    # -- collectives ----------------------------------------------------------------------------
    def _allreduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.n_collectives += 1

    def _allgather_rows(self, full, per):
        """full: [per * world][KP]; every rank has filled rows [rank*per, (rank+1)*per)."""
        if self.world > 1:
            local = full[self.rank * per:(self.rank + 1) * per]
            dist.all_gather_into_tensor(full, local.clone(), group=self.group)
            self.n_collectives += 1

    def _reduce_scatter_rows(self, full, per):
        """Sum `full` ([per * world][KP]) over ranks; rank r ends up with rows [r*per, (r+1)*per) summed."""
        if self.world > 1:
            mine = full[self.rank * per:(self.rank + 1) * per]
            if dist.get_backend(self.group) == "gloo":  # gloo has no reduce_scatter: all-reduce, keep own rows
                dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            else:
                out = torch.empty_like(mine)
                dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=self.group)
                mine.copy_(out)
            self.n_collectives += 1

    # -- pieces ---------------------------------------------------------------------------------
    def set_w(self, w_host_kxm):
        self.be.factor_from_host(w_host_kxm, self.W)

    def local(self, full, lo, hi):
        return full[lo:hi] if hi > lo else full[0:0]

    def half_update(self, X, mask, F_in, F_full, lo, hi, per, L1, L2, gram_src_cols, gram_partial_of):
        be, k = self.be, self.k
        # Gram of the gather operand: partial over this rank's rows of it, then all-reduce
        glo, ghi = gram_partial_of
        be.gram(F_in[glo:ghi] if ghi > glo else F_in[0:0], k, max(ghi - glo, 0), self.gram, jitter=False)
        self._allreduce(self.gram)
        be.gram_jitter(k, self.gram)
        out_local = F_full[lo:lo + per]  # padded shard view (contiguous rows)
        be.update(X, mask, F_in, out_local, k, self.gram, L1, L2, self.d)
        self._allreduce(self.d)
        be.finish_d(k, self.d)
        be.scale(out_local, k, hi - lo, self.d)
        self._allgather_rows(F_full, per)

    def iteration_b(self, L1_w, L1_h, L2_w, L2_h):
        """Layout B iteration: H stays sharded, W-update right-hand sides are reduce-scattered."""
        be, k = self.be, self.k
        self.Wprev.copy_(self.W)
        # H update over local cells against the replicated W (Gram from the full W: no collective)
        be.gram(self.W, k, self.m, self.gram, jitter=True)
        h_local = self.H[self.c0:self.c0 + self.c_per]
        be.update(self.A, None, self.W, h_local, k, self.gram, L1_h, L2_h, self.d)
        self._allreduce(self.d)
        be.finish_d(k, self.d)
        be.scale(h_local, k, self.c1 - self.c0, self.d)
        # W update: partial Gram and partial right-hand sides from the local cells
        be.gram(h_local, k, self.c1 - self.c0, self.gram, jitter=False)
        self._allreduce(self.gram)
        be.gram_jitter(k, self.gram)
        be.rhs(self.At, h_local, k, self.Bw)
        self._reduce_scatter_rows(self.Bw, self.g_per)
        w_local = self.W[self.g0:self.g0 + self.g_per]
        be.solve(self.Bw[self.g0:self.g0 + self.g_per], self.gene_ptr[self.g0:self.g0 + self.g_per + 1] if self.g1 > self.g0
                 else self.gene_ptr[0:1], self.g1 - self.g0, w_local, k, self.gram, L1_w, L2_w, self.d)
        self._allreduce(self.d)
        be.finish_d(k, self.d)
        be.scale(w_local, k, self.g1 - self.g0, self.d)
        self._allgather_rows(self.W, self.g_per)
        be.cor_sums(self.W, self.Wprev, k, self.m, self.sums)
        s = self.sums[:5].cpu().numpy()
        return be.cor_from_sums(s, float(k) * float(self.m))

    def iteration_b3(self, L1_w, L1_h, L2_w, L2_h):
        """Layout B with THREE exchanges per iteration instead of five -- the scheme of csrc/multi.cu sgl_fit_iterate on more
        than one rank: the partial Gram of the UNSCALED H and the row sums of H travel in one all-reduce and the Gram is
        rescaled afterwards (G_ij / (d_i d_j), sgl_dev_finish_d_rescale_gram); the W row sums are reduced next to the
        all-gather of the UNSCALED W shard (one grouped NCCL launch in the library) and every rank scales the whole W."""
        be, k, kp = self.be, self.k, self.kp
        self.Wprev.copy_(self.W)
        be.gram(self.W, k, self.m, self.gram, jitter=True)
        h_local = self.H[self.c0:self.c0 + self.c_per]
        be.update(self.A, None, self.W, h_local, k, self.gram, L1_h, L2_h, self.d)   # d <- local row sums of H
        be.gram(h_local, k, self.c1 - self.c0, self.gram, jitter=False)               # of the unscaled H
        self._allreduce(self.red[:kp * kp + kp])
        be.finish_d_rescale_gram(k, self.d, self.gram)
        be.scale(h_local, k, self.c1 - self.c0, self.d)
        be.rhs(self.At, h_local, k, self.Bw)
        self._reduce_scatter_rows(self.Bw, self.g_per)
        w_local = self.W[self.g0:self.g0 + self.g_per]
        be.solve(self.Bw[self.g0:self.g0 + self.g_per], self.gene_ptr[self.g0:self.g0 + self.g_per + 1] if self.g1 > self.g0
                 else self.gene_ptr[0:1], self.g1 - self.g0, w_local, k, self.gram, L1_w, L2_w, self.d)
        self._allreduce(self.d)                    # grouped with the all-gather in the library: one launch
        self._allgather_rows(self.W, self.g_per)   # unscaled shards
        self.n_collectives -= 1 if self.world > 1 else 0
        be.finish_d(k, self.d)
        be.scale(self.W, k, self.m, self.d)        # every rank: the whole W by the same d
        be.cor_sums(self.W, self.Wprev, k, self.m, self.sums)
        s = self.sums[:5].cpu().numpy()
        return be.cor_from_sums(s, float(k) * float(self.m))

    def gather_h(self):
        """Layout B keeps H sharded; gather it once (e.g. for the final output)."""
        self._allgather_rows(self.H, self.c_per)

    def iteration(self, L1_w, L1_h, L2_w, L2_h):
        """One trip of src/singlet.cpp:648-659. Returns tol (1 - cor) as a python float (synchronises)."""
        if self.layout == "B":
            return self.iteration_b(L1_w, L1_h, L2_w, L2_h)
        if self.layout == "B3":
            return self.iteration_b3(L1_w, L1_h, L2_w, L2_h)
        be, k = self.be, self.k
        self.Wprev.copy_(self.W)
        # H update over local cells; Gram of W: each rank sums its own genes
        self.half_update(self.A, self.mask_A, self.W, self.H, self.c0, self.c1, self.c_per, L1_h, L2_h, self.m,
                         (self.g0, self.g1))
        # W update over local genes; Gram of H: each rank sums its own cells
        self.half_update(self.At, self.mask_At, self.H, self.W, self.g0, self.g1, self.g_per, L1_w, L2_w, self.n,
                         (self.c0, self.c1))
        be.cor_sums(self.W, self.Wprev, k, self.m, self.sums)  # replicated W: same value on every rank
        s = self.sums[:5].cpu().numpy()
        return be.cor_from_sums(s, float(k) * float(self.m))

    def test_mse(self, which=0):
        be = self.be
        out = self.sums[5:6]
        be.mse(self.A, self.mask_A, self.W, self.d, self.H[self.c0:self.c0 + self.c_per], self.k, which, out)
        self._allreduce(out)
        return float(out.cpu().numpy()[0]) / float(self.n)

    def factors_to_host(self):
        if self.layout in ("B", "B3"):
            self.gather_h()
        w = self.be.factor_to_host(self.W, self.k, self.m)
        h = self.be.factor_to_host(self.H, self.k, self.n)
        d = self.d[: self.k].cpu().numpy().copy()
        return w, d, h


def sharded_nmf(backend, m, n, k, A_shard, At_shard, w_init, tol=1e-4, maxit=100, L1=(0.01, 0.01), L2=(0.0, 0.0), rank=0,
                world=1, group=None, layout="A"):
    """``c_nmf`` over shards (reference src/singlet.cpp:638-666). Returns dict(w, d, h, iter, tol).
    layout "A": At_shard = gene block over all cells; "B": At_shard = transpose of the local cell block."""
    fit = ShardedNMF(backend, m, n, k, A_shard, At_shard, rank, world, group, layout=layout)
    fit.set_w(w_init)
    tol_, it = 1.0, 0
    while it < maxit and tol_ > tol:
        tol_ = fit.iteration(L1[0], L1[1], L2[0], L2[1])
        it += 1
    w, d, h = fit.factors_to_host()
    return {"w": w, "d": d, "h": h, "iter": it, "tol": tol_}


def sharded_ard_nmf(backend, m, n, k, A_shard, At_shard, w_init, seed, inv_density, tol=1e-4, maxit=100, L1=0.01, L2=0.0,
                    overfit_threshold=1e-4, trace_test_mse=5, rank=0, world=1, group=None):
    """``c_ard_nmf`` over shards (reference src/singlet.cpp:1090-1152); masks are hashed with GLOBAL cell and
    gene indices (the chunk offsets of src/singlet.cpp:485, 590)."""
    c0, _, _ = shard_bounds(n, world, rank)
    g0, _, _ = shard_bounds(m, world, rank)
    mA = backend.mask_build(A_shard, seed, inv_density, 0, c0, 0)
    mAt = backend.mask_build(At_shard, seed, inv_density, 1, g0, 0)
    fit = ShardedNMF(backend, m, n, k, A_shard, At_shard, rank, world, group, mA, mAt)
    fit.set_w(w_init)
    test_mse, iters, tols, scores = [], [], [], []

    def push(it, tol_):
        test_mse.append(fit.test_mse(0))
        iters.append(it)
        tols.append(tol_)
        mn = min(test_mse)
        scores.append((test_mse[-1] - mn) / (test_mse[-1] + mn))

    tol_, it = 1.0, 0
    while it < maxit and tol_ > tol:
        tol_ = fit.iteration(L1, L1, L2, L2)
        if it % trace_test_mse == 0:
            push(it, tol_)
            if scores[-1] > overfit_threshold:
                break
        it += 1
    if it % trace_test_mse != 0:
        push(it, tol_)
    w, d, h = fit.factors_to_host()
    return {"w": w, "d": d, "h": h, "test_mse": np.array(test_mse), "iter": np.array(iters, np.int32), "tol": np.array(tols),
            "score_overfit": np.array(scores)}


def distributed_cross_validate_nmf(A, ranks, n_replicates=3, tol=1e-4, maxit=100, L1=0.01, L2=0, test_density=0.05,
                                   tol_overfit=1e-4, trace_test_mse=5, seed=None, device=None, group=None, concurrency=0, handle=None):
    """``cross_validate_nmf`` (reference R/cross_validate_nmf.R:18-105) with the (rank, replicate) grid dealt over the
    processes of ``group`` (one process per GPU): the fits are independent, so every process uploads A, runs its share
    with ``sgl_ard_nmf_batch`` and the traces are gathered -- no data-path collective ("replicas" of the CV path, SURVEY.md
    8e). Every process returns the same DataFrame, identical to ``api.cross_validate_nmf`` on one GPU. ``seed``: R's
    ``set.seed`` value (all processes must draw the same w_init and mask seeds); ``None`` continues the global stream of
    ``api.set_seed`` like ``api.cross_validate_nmf`` does."""
    import pandas as pd

    from . import api
    from .rrng import RRng

    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    # seed = None: the global R-style stream, exactly like api.cross_validate_nmf (every process must have called
    # api.set_seed with the same value); an explicit seed draws from a private stream and leaves the global one alone
    if seed is None:
        r = api._RNG
    else:
        r = RRng(0)
        r.set_seed(seed)
    ranks = [int(k) for k in np.atleast_1d(ranks)]
    A = api._as_csc(A)
    m = A.shape[0]
    w_init = [r.matrix_runif(max(ranks), m) for _ in range(n_replicates)]
    grid = [(k, rep) for rep in range(1, n_replicates + 1) for k in ranks]
    seeds = [abs(r.dot_random_seed(3 + rep)) for _, rep in grid]
    # largest ranks first, dealt round-robin: every process gets a similar mix of long and short fits
    order = sorted(range(len(grid)), key=lambda j: -grid[j][0])
    mine = order[rank::world]
    own = handle is None
    if own:
        handle = api.Handle(device if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0))
    try:
        models = api.c_ard_nmf_batch(A, None, tol, maxit, L1, L2, 0, [w_init[grid[j][1] - 1][:grid[j][0], :] for j in mine],
                                     [seeds[j] for j in mine], int(round(1 / test_density)), tol_overfit, trace_test_mse,
                                     concurrency, handle) if mine else []
    finally:
        if own:
            handle.close()
    part = {j: (mod["test_mse"], mod["iter"], mod["tol"]) for j, mod in zip(mine, models)}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, part, group=group)
        part = {j: v for g in gathered for j, v in g.items()}
    rows = []
    for j, (k, rep) in enumerate(grid):
        mse, it, ft = part[j]
        for q in range(len(mse)):
            rows.append({"k": k, "rep": rep, "test_error": mse[q], "iter": int(it[q]), "tol": ft[q]})
    return pd.DataFrame(rows, columns=["k", "rep", "test_error", "iter", "tol"])

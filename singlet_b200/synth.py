"""Deterministic synthetic sparse counts, host (numpy) restatement.

The same rules are implemented on the device in ``csrc/common.cuh: synth_entry`` so that every GPU
shard, the CPU oracle and the tests regenerate identical bits without moving data:

* genes are cut into strata of ``S = round(0.5 / density)`` consecutive rows; cell ``c`` has at most
  one non-zero per stratum ``s``:  ``u = splitmix64(seed ^ (c << 32 | s))``;
  present iff ``low32(u) < q32`` with ``q32 = round(density * S * 2^32)``;
  ``gene = s*S + (bits 32..47 of u) % S`` (dropped if ``>= m``);
  ``value = table[ctz(bits 48..63 of u | 0x80)]`` -- a geometric "count" 1..8 mapped through
  ``table[c-1] = float32(log1p(c * 1e4 / (2 * density * m)))`` (log-normalised counts).
  Rows come out ascending, nnz per column ~ Binomial(m/S, density*S).
* ``w_init[f, r] = ((splitmix64(w_seed ^ (r*k + f)) >> 11) + 0.5) * 2^-53`` (uniform (0,1), like runif).
"""
from __future__ import annotations

import numpy as np

DATA_SEED = 20240601
W_SEED = 123
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def spec(m: int, density: float):
    """(S, q32, values_table float32[8]) exactly as sgl_matrix_synth derives them."""
    S = int(np.floor(0.5 / density + 0.5))
    S = min(max(S, 1), 65535)
    q = min(density * S * 4294967296.0, 4294967295.0)
    q32 = int(np.floor(q + 0.5))
    table = np.array([np.log1p(c * 10000.0 / (2.0 * density * m)) for c in range(1, 9)], dtype=np.float32)
    return S, q32, table


def values_table(m: int, density: float) -> np.ndarray:
    return spec(m, density)[2]


def synth_csc(m: int, n: int, density: float, seed: int = DATA_SEED, col0: int = 0, ncol: int | None = None):
    """Columns [col0, col0+ncol) of the m x n matrix as (p int32, i int32, x float64)."""
    ncol = n - col0 if ncol is None else ncol
    S, q32, table = spec(m, density)
    n_strata = (m + S - 1) // S
    ps = [np.zeros(1, np.int64)]
    idx, val = [], []
    step = max(1, (1 << 22) // max(n_strata, 1))
    strata = np.arange(n_strata, dtype=np.uint64)
    for c0 in range(col0, col0 + ncol, step):
        cols = np.arange(c0, min(c0 + step, col0 + ncol), dtype=np.uint64)
        u = splitmix64(np.uint64(seed) ^ ((cols[:, None] << np.uint64(32)) | strata[None, :]))
        present = (u & np.uint64(0xFFFFFFFF)) < np.uint64(q32)
        gene = strata[None, :] * np.uint64(S) + (((u >> np.uint64(32)) & np.uint64(0xFFFF)) % np.uint64(S))
        present &= gene < np.uint64(m)
        hi = ((u >> np.uint64(48)) | np.uint64(0x80)).astype(np.uint32)
        ctz = np.zeros(hi.shape, dtype=np.int64)
        low = hi & (~hi + np.uint32(1))  # lowest set bit
        ctz = np.log2(low.astype(np.float64)).astype(np.int64)
        ps.append(present.sum(axis=1).astype(np.int64))
        idx.append(gene[present].astype(np.int32))
        val.append(table[ctz[present]].astype(np.float64))
    counts = np.concatenate(ps)
    p = np.cumsum(counts)
    assert p[-1] < 2**31
    return p.astype(np.int32), np.concatenate(idx) if idx else np.zeros(0, np.int32), \
        np.concatenate(val) if val else np.zeros(0, np.float64)


def synth_scipy(m: int, n: int, density: float, seed: int = DATA_SEED):
    import scipy.sparse as sp

    p, i, x = synth_csc(m, n, density, seed)
    return sp.csc_matrix((x, i, p), shape=(m, n))


def w_init(k: int, m: int, seed: int = W_SEED) -> np.ndarray:
    """k x m (Fortran order) uniform (0, 1) initial factor."""
    t = np.arange(k * m, dtype=np.uint64)
    u = splitmix64(np.uint64(seed) ^ t)
    w = ((u >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0**-53
    return w.reshape((k, m), order="F")


def algorithmic_bytes_per_iter(m: int, n: int, nnz: int, k: int) -> int:
    """B_iter = 16*nnz + 4*(m+n+2) + 16*k*(m+n)."""
    return 16 * nnz + 4 * (m + n + 2) + 16 * k * (m + n)


def spmm_bytes(nnz: int, ncol: int, nrow: int, k: int) -> int:
    """SURVEY.md 8(d) per-kernel figure: 8*nnz + 4*(ncol+1) + 4*k*nrow + 4*k*ncol."""
    return 8 * nnz + 4 * (ncol + 1) + 4 * k * nrow + 4 * k * ncol

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle

    return Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle.pyoracle import Oracle, have_reference

    if not have_reference():
        pytest.skip("oracle/_ref/libsinglet_ref.so not built (needs /root/reference at build time)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def handle():
    from singlet_b200 import api

    return api.default_handle()


def match_factors(w_a, w_b):
    """Match factors of two k x m matrices by maximum cosine: returns perm so that
    w_b[perm[f]] pairs with w_a[f]."""
    na = w_a / (np.linalg.norm(w_a, axis=1, keepdims=True) + 1e-300)
    nb = w_b / (np.linalg.norm(w_b, axis=1, keepdims=True) + 1e-300)
    cos = na @ nb.T
    perm, used = [], set()
    for f in np.argsort(-cos.max(axis=1)):
        order = np.argsort(-cos[f])
        j = next(int(j) for j in order if int(j) not in used)
        used.add(j)
        perm.append((int(f), j))
    out = np.zeros(w_a.shape[0], dtype=int)
    for f, j in perm:
        out[f] = j
    return out


def min_factor_cor(a, b, perm=None):
    """Minimum per-factor Pearson correlation between the rows of a and b (k x cols)."""
    k = a.shape[0]
    perm = np.arange(k) if perm is None else perm
    cors = []
    for f in range(k):
        x, y = a[f], b[perm[f]]
        if x.std() == 0 and y.std() == 0:
            cors.append(1.0)
        else:
            cors.append(float(np.corrcoef(x, y)[0, 1]))
    return min(cors)

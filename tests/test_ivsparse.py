"""IVSparse wire formats: csrc/ivsparse.cpp against the reference's own codec compiled from its vendored
headers (oracle/_ref/libivsparse_ref.so) -- byte-identical images both ways -- and against committed golden images made
with it (tests/golden/ivsparse_*.bin, scripts/make_ivsparse_goldens.py) so the check also runs where /root/reference is
absent. The fit on a decoded image is checked on the GPU against the oracle on the float-narrowed matrix."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libivsparse_ref.so")
GOLD = os.path.join(ROOT, "tests", "golden")


def counts_matrix(m, n, density, seed, nvals=12, dtype=np.float64):
    """Count-like data (few distinct values per column, the case the formats are made for) plus some columns that are
    empty, single-valued or full."""
    rs = np.random.RandomState(seed)
    A = sp.random(m, n, density, format="csc", random_state=rs, data_rvs=lambda s: rs.randint(1, nvals + 1, s).astype(np.float64))
    A = A.tolil()
    if n > 6:
        A[:, 2] = 0
        A[:, 5] = 3.0
        A[0, 4] = 7.0  # row 0 first in a run: an absolute first index of 0 must not read as a delimiter
        A[0, 6] = 1.0
    A = A.tocsc().astype(dtype)
    A.eliminate_zeros()
    A.sort_indices()
    return A


def no_empty_columns(A):
    """The reference's FILE reader fails on a column of size 0 (its `fread(...) == 0` check,
    IVCSC_Constructors.hpp:605-608, throws "Could not read file" for an empty blob), so anything read back BY THE REFERENCE
    gets one entry per empty column; our decoder accepts empty columns (checked separately)."""
    A = A.tolil()
    for c in np.flatnonzero(np.diff(A.tocsc().indptr) == 0):
        A[c % A.shape[0], c] = 2.0
    A = A.tocsc()
    A.sort_indices()
    return A


def ref_lib():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libivsparse_ref.so not built (needs /root/reference at build time)")
    lib = C.CDLL(REF_SO)
    lib.ref_ivsparse_write.restype = C.c_int
    lib.ref_ivsparse_write.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p]
    lib.ref_ivsparse_read.restype = C.c_int64
    lib.ref_ivsparse_read.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def ref_write(lib, A, level, path):
    vals = A.data.astype(np.float32)
    idx = A.indices.astype(np.uint64)
    ptr = A.indptr.astype(np.uint64)
    lib.ref_ivsparse_write(level, vals.ctypes.data, idx.ctypes.data, ptr.ctypes.data, A.shape[0], A.shape[1], A.nnz, path.encode())
    return np.fromfile(path, np.uint8)


def ref_read(lib, level, path, nnz):
    r, c, v = np.zeros(nnz, np.uint64), np.zeros(nnz, np.uint64), np.zeros(nnz, np.float32)
    n = lib.ref_ivsparse_read(level, path.encode(), r.ctypes.data, c.ctypes.data, v.ctypes.data)
    assert n == nnz
    return r, c, v


SHAPES = [(300, 40, 0.08, 1), (70000, 9, 0.01, 2), (200, 300, 0.3, 3), (5, 7, 0.6, 4)]


@pytest.mark.parametrize("level", [3, 2])
@pytest.mark.parametrize("shape", SHAPES)
def test_encode_is_the_reference_image(tmp_path, level, shape):
    """Our encoder writes byte for byte what the reference's compressCSC + write produce (70,000 rows: 3-byte deltas)."""
    from singlet_b200 import ivsparse

    lib = ref_lib()
    A = counts_matrix(*shape)
    want = ref_write(lib, A, level, str(tmp_path / "ref.bin"))
    got = ivsparse.encode(A, level)
    assert got.nbytes == want.nbytes and np.array_equal(got, want)
    # a chunk list is the concatenation (IVCSC::append, src/singlet.cpp:826-832)
    cut = A.shape[1] // 3
    got2 = ivsparse.encode([A[:, :cut], A[:, cut:]], level)
    assert np.array_equal(got2, want)


@pytest.mark.parametrize("level", [3, 2])
@pytest.mark.parametrize("shape", SHAPES)
def test_decode_reads_what_the_reference_reads(tmp_path, level, shape):
    """Our decoder on an image the REFERENCE wrote yields the coordinates its InnerIterator yields (as a set per column:
    the file groups by value, a dgCMatrix sorts by row) and round-trips to the float-narrowed input."""
    from singlet_b200 import ivsparse

    lib = ref_lib()
    A = no_empty_columns(counts_matrix(*shape))
    path = str(tmp_path / "ref.bin")
    img = ref_write(lib, A, level, path)
    md = ivsparse.info(img)
    assert md == {"level": level, "nrow": A.shape[0], "ncol": A.shape[1], "nnz": A.nnz, "value_bytes": 4}
    D = ivsparse.decode(img)
    r, c, v = ref_read(lib, level, path, A.nnz)
    R = sp.csc_matrix((v.astype(np.float64), (r.astype(np.int64), c.astype(np.int64))), shape=A.shape)
    R.sort_indices()
    assert np.array_equal(D.indptr, R.indptr) and np.array_equal(D.indices, R.indices) and np.array_equal(D.data, R.data)
    A32 = A.astype(np.float32).astype(np.float64)
    assert np.array_equal(D.indptr, A32.indptr) and np.array_equal(D.indices, A32.indices) and np.array_equal(D.data, A32.data)
    # a column range in the middle, and the reference reading OUR image
    lo, hi = A.shape[1] // 4, A.shape[1] - 1
    S = ivsparse.decode(img, lo, hi - lo)
    assert (S != A32[:, lo:hi]).nnz == 0 and S.shape == (A.shape[0], hi - lo)
    ours = str(tmp_path / "ours.bin")
    ivsparse.encode(A, level).tofile(ours)
    r2, c2, v2 = ref_read(lib, level, ours, A.nnz)
    assert np.array_equal(r, r2) and np.array_equal(c, c2) and np.array_equal(v, v2)


@pytest.mark.parametrize("level", [3, 2])
def test_golden_images(level):
    """Committed images written by the reference's codec (scripts/make_ivsparse_goldens.py): travel to the GPU box."""
    from singlet_b200 import ivsparse

    g = np.load(os.path.join(GOLD, "ivsparse_golden.npz"))
    A = sp.csc_matrix((g["x"], g["i"], g["p"]), shape=tuple(g["shape"]))
    want = g["image_l%d" % level]
    assert np.array_equal(ivsparse.encode(A, level), want)
    D = ivsparse.decode(want)
    assert np.array_equal(D.indptr, A.indptr) and np.array_equal(D.indices, A.indices)
    assert np.array_equal(D.data, A.data.astype(np.float32).astype(np.float64))


def test_non_float_values_and_index_widths():
    """Images of other instantiations of the reference's template (integer values, narrow VCSC index types) decode: the value
    type word and index width are honoured (IVCSC_Private_Methods.hpp:65-72, VCSC_Methods.hpp:77-109)."""
    from singlet_b200 import ivsparse

    A = counts_matrix(120, 30, 0.2, 9)
    img = ivsparse.encode(A, 2)
    md = np.frombuffer(img[:24].tobytes(), np.uint32)
    n = A.shape[1]
    vs = np.frombuffer(img[24:24 + 8 * n].tobytes(), np.uint64)
    isz = np.frombuffer(img[24 + 8 * n:24 + 16 * n].tobytes(), np.uint64)
    nv, ni = int(vs.sum()), int(isz.sum())
    body = img[24 + 16 * n:]
    vals = np.frombuffer(body[:4 * nv].tobytes(), np.float32)
    cnts = np.frombuffer(body[4 * nv:12 * nv].tobytes(), np.uint64)
    idx = np.frombuffer(body[12 * nv:12 * nv + 8 * ni].tobytes(), np.uint64)
    # VCSC<int16_t, uint16_t>: value type = 2 | signed << 16 | col-major << 24, index width 2
    md2 = md.copy()
    md2[4] = 2 | (1 << 16) | (1 << 24)
    md2[5] = 2
    img2 = np.concatenate([np.frombuffer(md2.tobytes(), np.uint8), np.frombuffer(vs.astype(np.uint16).tobytes(), np.uint8),
                           np.frombuffer(isz.astype(np.uint16).tobytes(), np.uint8),
                           np.frombuffer(vals.astype(np.int16).tobytes(), np.uint8), np.frombuffer(cnts.astype(np.uint16).tobytes(), np.uint8),
                           np.frombuffer(idx.astype(np.uint16).tobytes(), np.uint8)])
    D = ivsparse.decode(img2)
    assert ivsparse.info(img2)["value_bytes"] == 2 and (D != A).nnz == 0


def test_malformed_images_are_rejected():
    from singlet_b200 import SingletCudaError, ivsparse

    A = counts_matrix(300, 40, 0.08, 1)
    for level in (3, 2):
        img = ivsparse.encode(A, level)
        with pytest.raises(SingletCudaError):
            ivsparse.decode(img[:20])
        with pytest.raises(SingletCudaError):
            ivsparse.decode(img[:len(img) - 7])  # truncated body
        bad = img.copy()
        bad[0] = 1  # level 1 (plain CSC) is not a file format of the path
        with pytest.raises(SingletCudaError):
            ivsparse.info(bad)
        with pytest.raises(SingletCudaError):
            ivsparse.decode(img, 30, 20)  # range past the last column
    img = ivsparse.encode(A, 3)
    bad = img.copy()
    off = 24 + 8 * A.shape[1]
    bad[off + 4] = 9  # index width of the first run
    with pytest.raises(SingletCudaError):
        ivsparse.decode(bad)
    with pytest.raises(SingletCudaError):
        ivsparse.encode(A, 4)


def test_decode_chunks_and_files(tmp_path):
    """decode_chunks cuts an image into a chunk list under a non-zero budget; the reference-named file functions round-trip."""
    from singlet_b200 import ivsparse

    A = counts_matrix(400, 120, 0.1, 5)
    img = ivsparse.encode(A, 3)
    chunks = ivsparse.decode_chunks(img, max_nnz=A.nnz // 5)
    assert len(chunks) >= 5 and all(c.nnz <= A.nnz // 5 for c in chunks)
    assert (sp.hstack(chunks, format="csc") != A).nnz == 0
    d = str(tmp_path)
    assert ivsparse.save_IVSparse([A[:, :50], A[:, 50:]], verbose=False, directory=d)
    assert np.array_equal(np.fromfile(os.path.join(d, "IVCSC_matrix.ivsparse"), np.uint8), img)
    assert (ivsparse.read_IVSparse(d) != A).nnz == 0
    assert ivsparse.write_IVCSC([A], verbose=False, directory=d)
    T = ivsparse.decode(np.fromfile(os.path.join(d, ivsparse.WRITE_IVCSC_T_FILE), np.uint8))
    assert (T != A.T.tocsc()).nnz == 0


@pytest.mark.gpu
@pytest.mark.parametrize("use_vcsc", [False, True])
def test_run_nmf_on_sparsematrix_list_matches_oracle(oracle, use_vcsc):
    """The list path of run_nmf (R/run_nmf.R:21-35 -> src/singlet.cpp:946-995): float-narrowed values, L1 = L2 = 0, against
    the oracle's c_nmf on the same narrowed matrix (the reference sums products in value-grouped order, a rounding-level
    difference; tolerance = north_star's: cor >= 0.999, d within 1e-3)."""
    from conftest import min_factor_cor
    from singlet_b200 import api, ivsparse, synth

    m, n, k = 500, 360, 6
    rs = np.random.RandomState(3)
    A = sp.random(m, n, 0.08, format="csc", random_state=rs, data_rvs=lambda s: np.log1p(rs.randint(1, 30, s) / 3.0))
    A.sort_indices()
    parts = [A[:, :100], A[:, 100:250], A[:, 250:]]
    w0 = synth.w_init(k, m, seed=5)
    dev = ivsparse.run_nmf_on_sparsematrix_list(parts, 1e-5, 12, False, 0, w0, use_vcsc)
    A32 = A.astype(np.float32).astype(np.float64)
    At32 = A32.T.tocsc()
    At32.sort_indices()
    orc = oracle.nmf(A32, At32, w0, tol=1e-5, maxit=12, L1=(0.0, 0.0), L2=(0.0, 0.0))
    assert min_factor_cor(dev["w"], orc["w"]) > 0.999 and min_factor_cor(dev["h"], orc["h"]) > 0.999
    assert np.allclose(dev["d"], orc["d"], rtol=1e-3)
    # run_nmf routes a list here and ignores the caller's L1/L2 like the reference does (R/run_nmf.R:33)
    api.set_seed(123)
    mod = api.run_nmf(parts, k, tol=1e-5, maxit=12, verbose=False, L1=0.5)
    api.set_seed(123)
    w_init = api._RNG.matrix_runif(k, m)
    direct = ivsparse.run_nmf_on_sparsematrix_list(parts, 1e-5, 12, False, 0, w_init, False)
    order = np.argsort(-direct["d"], kind="stable")
    assert np.allclose(mod["d"], direct["d"][order]) and np.allclose(mod["w"], direct["w"].T[:, order])

"""Bit-exact parity of the device-side dataset bookkeeping (X[indicesRow], CSC row
slices, vstack, toCSC/toCSR, shuffle) against the oracle's restatement of tensor/sparse.nim, through
the C ABI.  Integer work: every array must be IDENTICAL (indices, indptr, order within a segment)."""
import numpy as np
import pytest

import nimfm_b200 as nf
from oracle.oracle import CSR
from helpers import make_dense, make_field_csr

pytestmark = pytest.mark.gpu


def csr_ds(c):
    return nf.newCSRDataset(c.data, c.indices, c.indptr, c.n, c.d)


def csc_ds(c):
    return nf.newCSCDataset(c.data, c.indices, c.indptr, c.n, c.d)


def same(ds, ref, fields=False):
    data, indices, indptr, fld = ds.download()
    assert np.array_equal(indptr, ref.indptr)
    assert np.array_equal(indices, ref.indices)
    assert np.array_equal(data, ref.data)
    if fields:
        assert np.array_equal(fld, ref.fields)
    # the host mirror of a library-made dataset carries the same arrays
    assert np.array_equal(ds.indptr, ref.indptr) and np.array_equal(ds.indices, ref.indices)
    assert ds.nnz == len(ref.data) and ds.shape == (ref.n, ref.d)


def ragged(n, d, seed):
    """empty rows, one very long row, duplicate-free sorted indices"""
    rng = np.random.default_rng(seed)
    X = make_dense(n, d, seed, density=0.2)
    X[::5] = 0.0
    X[n // 2] = rng.random(d) + 0.1
    return CSR.from_dense(X)


@pytest.mark.parametrize("n,d,seed", [(37, 19, 1), (200, 70, 2), (1, 5, 3)])
def test_take_rows_bit_exact(oracle, n, d, seed):
    csr = ragged(n, d, seed)
    rng = np.random.default_rng(seed)
    ds = csr_ds(csr)
    for rows in (rng.permutation(n), rng.integers(0, n, size=3 * n), np.array([n - 1]), np.array([0, 0, 0])):
        same(ds[rows], oracle.csr_take_rows(csr, rows))
    # slices: Python's X[a:b] == Nim's X[a..b-1]
    if n > 4:
        same(ds[2:n - 1], oracle.csr_take_rows(csr, np.arange(2, n - 1)))


def test_take_rows_errors():
    csr = ragged(10, 6, 4)
    ds = csr_ds(csr)
    with pytest.raises(ValueError, match=">= 10"):        # sparse.nim:254-257
        ds[np.array([3, 10])]
    with pytest.raises(ValueError, match="< 0"):          # sparse.nim:259-260
        ds[np.array([-1, 2])]
    with pytest.raises(ValueError):                       # CSC: index arrays unsupported (sparse.nim:298-299)
        ds.toCSCDataset()[np.array([1, 2])]


def test_take_rows_with_fields_and_targets(oracle):
    X, csr, _ = make_field_csr(40, 12, 4, 5)
    ds = nf.newCSRFieldDataset(csr.data, csr.indices, csr.indptr, csr.fields, csr.n, csr.d, csr.n_fields)
    y = np.arange(40, dtype=np.float64)
    ds.set_targets(y)
    rows = np.random.default_rng(6).permutation(40)[:25]
    sub = ds[rows]
    ref = oracle.csr_take_rows(csr, rows)
    ref.fields = np.concatenate([csr.fields[csr.indptr[r]:csr.indptr[r + 1]] for r in rows]).astype(np.int64)
    same(sub, ref, fields=True)
    assert sub.nFields == 4
    # shuffle(X, y, indices) (dataset.nim:372-381)
    Xs, ys = nf.shuffle(ds, y, rows)
    assert np.array_equal(ys, y[rows]) and np.array_equal(Xs.indices, ref.indices)


@pytest.mark.parametrize("a,b", [(0, 36), (5, 20), (36, 36), (0, 0)])
def test_csc_slice_bit_exact(oracle, a, b):
    csr = ragged(37, 19, 7)
    csc_ref = oracle.csr_to_csc(csr)
    csc = csc_ds(csc_ref)
    same(csc[a:b + 1], oracle.csc_slice_rows(csc_ref, a, b))
    # unsorted rows inside a column keep their order (the filter is stable)
    perm = np.arange(len(csc_ref.data))
    for j in range(csc_ref.d):
        s, e = csc_ref.indptr[j], csc_ref.indptr[j + 1]
        perm[s:e] = perm[s:e][::-1]
    rev = CSR(csc_ref.data[perm], csc_ref.indices[perm], csc_ref.indptr, csc_ref.n, csc_ref.d)
    same(csc_ds(rev)[a:b + 1], oracle.csc_slice_rows(rev, a, b))


def test_vstack_bit_exact(oracle):
    parts = [ragged(17, 11, 8), ragged(5, 11, 9), ragged(30, 11, 10)]
    same(nf.vstack(*[csr_ds(p) for p in parts]), oracle.csr_vstack(parts))
    cparts = [oracle.csr_to_csc(p) for p in parts]
    same(nf.vstack([csc_ds(p) for p in cparts]), oracle.csc_vstack(cparts))
    # vstack of the CSC parts == CSC of the stacked CSR
    whole = oracle.csr_to_csc(oracle.csr_vstack(parts))
    same(nf.vstack([csc_ds(p) for p in cparts]), whole)
    with pytest.raises(ValueError, match="same shape"):   # sparse.nim:576-577
        nf.vstack(csr_ds(parts[0]), csr_ds(ragged(4, 12, 11)))
    # field datasets
    _, f1, _ = make_field_csr(9, 10, 3, 12)
    _, f2, _ = make_field_csr(6, 10, 3, 13)
    mk = lambda c: nf.newCSRFieldDataset(c.data, c.indices, c.indptr, c.fields, c.n, c.d, c.n_fields)
    same(nf.vstack(mk(f1), mk(f2)), oracle.csr_vstack([f1, f2]), fields=True)


@pytest.mark.parametrize("n,d,density,seed", [(37, 19, 0.35, 1), (500, 300, 0.05, 2), (64, 3, 1.0, 3), (3, 1000, 0.01, 4)])
def test_device_transpose_bit_exact(oracle, n, d, density, seed):
    csr = CSR.from_dense(make_dense(n, d, seed, density=density))
    ref = oracle.csr_to_csc(csr)
    ds = csr_ds(csr)
    csc = ds.toCSCDataset()                      # tensor/sparse.nim:510-527
    same(csc, ref)
    same(csc.toCSRDataset(), csr)                # :490-507
    assert csc.info()["maxSegNnz"] == int(np.max(np.diff(ref.indptr))) if len(ref.data) else True


def test_transpose_large_property():
    """size-independent properties at a size the oracle would not finish quickly: transpose twice is
    the identity; column lengths equal the index histogram; rows ascend within every column."""
    rng = np.random.default_rng(20)
    n, d, z = 200_000, 5_000, 12
    cols = np.sort(rng.integers(0, d // z, size=(n, z)) + (np.arange(z) * (d // z))[None, :], axis=1)
    data = rng.random(n * z)
    ds = nf.newCSRDataset(data, cols.ravel(), np.arange(n + 1) * z, n, d)
    csc = ds.toCSCDataset()
    assert np.array_equal(np.diff(csc.indptr), np.bincount(cols.ravel(), minlength=d))
    seg = np.repeat(np.arange(d), np.diff(csc.indptr))
    asc = (np.diff(csc.indices) > 0) | (np.diff(seg) != 0)
    assert bool(np.all(asc))
    back = csc.toCSRDataset()
    assert np.array_equal(back.indices, cols.ravel()) and np.array_equal(back.data, data)
    assert np.array_equal(back.indptr, np.arange(n + 1) * z)

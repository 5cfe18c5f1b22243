"""Text loaders (loadSVMLightFile / loadFFMFile / loadUserItemRatingFile,
dataset.nim:562-990) -- the library's threaded parser + direct device upload against the oracle's
restatement of the reference's read loops.  Index bookkeeping must be bit-exact; values are parsed
to the nearest double by both sides (shortest round-trip text), so they are compared exactly too."""
import numpy as np
import pytest

import nimfm_b200 as nf
from helpers import make_dense, make_field_csr
from oracle.oracle import CSR

pytestmark = pytest.mark.gpu


def same(ds, ref, fields=False):
    data, indices, indptr, fld = ds.download()
    assert ds.shape == (ref.n, ref.d)
    assert np.array_equal(indptr, ref.indptr) and np.array_equal(indices, ref.indices)
    assert np.array_equal(data, ref.data)
    if fields:
        assert np.array_equal(fld, ref.fields) and ds.nFields == ref.n_fields


def write_svm(path, csr, y, base=1, sep="\n", trailing=False):
    lines = []
    for i in range(csr.n):
        s, e = csr.indptr[i], csr.indptr[i + 1]
        lines.append(repr(float(y[i])) + "".join(f" {csr.indices[q] + base}:{float(csr.data[q])!r}" for q in range(s, e)))
    open(path, "w").write(sep.join(lines) + (sep if trailing else ""))


@pytest.mark.parametrize("base,trailing", [(1, False), (0, False), (1, True)])
def test_svmlight_matches_reference_reader(oracle, tmp_path, base, trailing):
    X = make_dense(57, 23, 3, density=0.3, positive=False)
    X[5] = 0.0                                   # a sample without features
    if base == 0:
        X[0, 0] = 1.5                            # a 0 index makes the file 0-based (dataset.nim:585)
    else:
        X[:, 0] = 0.0                            # 1-based file without index 1: offset stays 1
    csr = CSR.from_dense(X)
    y = np.random.default_rng(1).standard_normal(57)
    p = str(tmp_path / "a.svm")
    write_svm(p, csr, y, base=base, trailing=trailing)
    ref, yref = oracle.load_svmlight(p)
    ds, yy = nf.loadSVMLightFile(p)
    same(ds, ref)
    assert np.array_equal(yy, yref) and np.array_equal(yy, y)
    assert np.array_equal(ref.to_dense()[:, :X.shape[1]][:, :ref.d], X[:, :ref.d])
    # nFeatures widens, or raises when too small (:625-632)
    wide, _ = nf.loadSVMLightFile(p, nFeatures=40)
    assert wide.shape == (57, 40)
    with pytest.raises(ValueError, match="but dataset has at least"):
        nf.loadSVMLightFile(p, nFeatures=3)
    # the CSC overload (:643-686) == stable transpose of the CSR
    csc, yc = nf.loadSVMLightFile(p, kind="csc")
    cref = oracle.csr_to_csc(ref)
    same(csc, cref)
    assert isinstance(csc, nf.CSCDataset) and np.array_equal(yc, y)


def test_svmlight_dump_load_roundtrip_and_large(tmp_path):
    """dumpSVMLightFile -> loadSVMLightFile is the identity; a file large enough to be parsed by
    several threads keeps row order"""
    rng = np.random.default_rng(7)
    n, d, z = 60_000, 500, 9
    cols = np.sort(rng.integers(0, d // z, size=(n, z)) + (np.arange(z) * (d // z))[None, :], axis=1)
    cols[0, 0], cols[-1, -1] = 0, d - 1
    data = rng.standard_normal(n * z)
    y = rng.integers(0, 2, n).astype(np.float64) * 2 - 1
    ds = nf.newCSRDataset(data, cols.ravel(), np.arange(n + 1) * z, n, d)
    p = str(tmp_path / "big.svm")
    nf.dumpSVMLightFile(p, ds, y)
    assert not open(p).read().endswith("\n")
    back, yb = nf.loadSVMLightFile(p)
    assert back.shape == (n, d)
    assert np.array_equal(back.indices, cols.ravel()) and np.array_equal(back.data, data)
    assert np.array_equal(back.indptr, np.arange(n + 1) * z) and np.array_equal(yb, y)


def test_svmlight_errors(tmp_path):
    p = str(tmp_path / "neg.svm")
    open(p, "w").write("1 -2:0.5 3:1\n0 1:2\n")
    with pytest.raises(ValueError, match="Negative index is included"):
        nf.loadSVMLightFile(p)
    open(p, "w").write("1 2:0.5 x:1\n")
    with pytest.raises(ValueError, match="malformed"):
        nf.loadSVMLightFile(p)
    with pytest.raises(ValueError, match="cannot open"):
        nf.loadSVMLightFile(str(tmp_path / "missing.svm"))


def test_ffm_file_matches_reference_reader(oracle, tmp_path):
    X, csr, field_of = make_field_csr(45, 14, 4, 9)
    y = np.sign(np.random.default_rng(2).standard_normal(45))
    ds0 = nf.newCSRFieldDataset(csr.data, csr.indices, csr.indptr, csr.fields, csr.n, csr.d, csr.n_fields)
    p = str(tmp_path / "a.ffm")
    nf.dumpFFMFile(p, ds0, y)
    ref, yref = oracle.load_ffm(p)
    ds, yy = nf.loadFFMFile(p)
    same(ds, ref, fields=True)
    assert np.array_equal(yy, yref) and np.array_equal(yy, y)
    assert np.array_equal(ds.indices, csr.indices[:len(ds.indices)]) and np.array_equal(ds.fields, csr.fields)
    wide, _ = nf.loadFFMFile(p, nFeatures=30, nFields=6)
    assert wide.shape == (45, 30) and wide.nFields == 6
    with pytest.raises(ValueError, match="fields"):
        nf.loadFFMFile(p, nFields=2)
    # and the loaded dataset drives the FFM path
    rng = np.random.default_rng(3)
    m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=4, warmStart=True)
    m.P, m.w, m.intercept, m.isInitialized = rng.standard_normal((4, ds.nFeatures, 4)) * 0.1, np.zeros(ds.nFeatures), 0.0, True
    got = m.decisionFunction(ds)
    want = oracle.ffm_decision_function(ref, m.P, m.w, 0.0)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12)) < 1e-10


def test_user_item_rating_file(oracle, tmp_path):
    rng = np.random.default_rng(4)
    n = 300
    users, items = rng.integers(3, 40, n), rng.integers(10, 90, n)
    ratings = rng.integers(1, 6, n)
    seps = [" ", "|", "\t", "::"]
    p = str(tmp_path / "u.data")
    open(p, "w").write("\n".join(f"{u}{seps[i % 4]}{v}{seps[i % 4]}{r}{seps[i % 4]}8812345" for i, (u, v, r)
                                 in enumerate(zip(users, items, ratings))) + "\n")
    ref, yref = oracle.load_user_item_rating(p)
    ds, y = nf.loadUserItemRatingFile(p)
    same(ds, ref)
    assert np.array_equal(y, yref) and np.array_equal(y, ratings.astype(np.float64))
    # minUser / minItem start at 1 (dataset.nim:862-866): ids are taken as 1-based unless a 0 occurs
    assert ds.nFeatures == (users.max() - 1 + 1) + (items.max() - 1 + 1)
    csc, _ = nf.loadUserItemRatingFile(p, kind="csc")
    same(csc, oracle.csr_to_csc(ref))


def test_cli_train_dump_test_roundtrip(oracle, tmp_path, capsys):
    """`nimfm train` / `nimfm test` (src/nimfm.nim:72-135) on the device path: svmlight in, CD fit, model
    dump, a second process-level step reloading the model and writing the same predictions"""
    from nimfm_b200.__main__ import main
    rng = np.random.default_rng(5)
    X = make_dense(90, 12, 4, density=0.5, positive=False)
    csr = CSR.from_dense(X)
    wtrue = rng.standard_normal(12)
    y = X @ wtrue + 0.1 * rng.standard_normal(90)
    tr, te = str(tmp_path / "tr.svm"), str(tmp_path / "te.svm")
    write_svm(tr, csr, y)
    write_svm(te, csr, y)
    model, p1, p2 = str(tmp_path / "m.txt"), str(tmp_path / "p1.txt"), str(tmp_path / "p2.txt")
    main(["train", "--task", "r", "--train", tr, "--test", te, "--degree", "2", "--nComponents", "3", "--solver", "cd",
          "--maxIter", "5", "--dump", model, "--predict", p1, "--nFeatures", "12", "--verbose", "0"])
    out = capsys.readouterr().out
    assert "Test RMSE: " in out
    rmse = float(out.split("Test RMSE: ")[1].split()[0])
    main(["test", "--task", "r", "--test", te, "--load", model, "--predict", p2, "--nFeatures", "12", "--verbose", "0"])
    out2 = capsys.readouterr().out
    assert abs(float(out2.split("Test RMSE: ")[1].split()[0]) - rmse) < 1e-12
    a, b = np.loadtxt(p1), np.loadtxt(p2)
    assert np.array_equal(a, b) and len(a) == 90
    # the CLI's fit equals the library's (CD, squared loss, the CLI's defaults) -- and the oracle's
    fm = nf.FactorizationMachine.load(model, False)
    csc = oracle.csr_to_csc(csr)
    rngP = np.random.default_rng(1).standard_normal((1, 3, 12)) * 0.1     # init() of the mirror: seed=randomState, scale
    ref = oracle.cd_fit(csc, y, rngP, np.zeros(12), 0.0, 2, "squared", max_iter=5, alpha0=1e-7, alpha=1e-5, beta=1e-3,
                        tol=1e-5)
    np.testing.assert_allclose(fm.P, ref["P"], rtol=1e-8, atol=1e-12)
    for solver in ("sgd", "adagrad"):
        main(["train", "--task", "c", "--train", tr, "--solver", solver, "--loss", "logistic", "--maxIter", "2",
              "--nComponents", "2", "--verbose", "0"])


def test_stream_binary_files(oracle, tmp_path):
    """STREAMCSR / STREAMCSC binary files (tensor/sparse_stream.nim:3-33; convertSVMLightFile /
    transposeFile / loadStreamLabel, dataset.nim:995-1200): converted from svmlight text, loaded whole
    into device datasets (the payload is de-interleaved on the device), bit-exact with the text loader"""
    X = make_dense(73, 21, 8, density=0.25, positive=False)
    X[7] = 0.0
    csr = CSR.from_dense(X)
    y = np.random.default_rng(2).standard_normal(73)
    txt, fx, fy, fxT = (str(tmp_path / n) for n in ("a.svm", "a.bin", "a.lab", "aT.bin"))
    write_svm(txt, csr, y)
    nf.convertSVMLightFile(txt, fx, fy)
    raw = open(fx, "rb").read()
    assert raw[:9] == b"STREAMCSR"
    hdr = np.frombuffer(raw[9:33], dtype="<i8")
    ref, yref = oracle.load_svmlight(txt)
    assert hdr[0] == 73 and hdr[2] == len(ref.data)
    ds = nf.newStreamCSRDataset(fx)
    assert np.array_equal(ds.indptr, ref.indptr) and np.array_equal(ds.data, ref.data)
    # both conventions subtract the minimum index (1 here), so the ids agree with the text loader
    assert np.array_equal(ds.indices, ref.indices)
    assert np.array_equal(nf.loadStreamLabel(fy), y) and np.array_equal(nf.loadStreamLabel(fy, 10), y[:10])
    nf.transposeFile(fx, fxT)
    assert open(fxT, "rb").read()[:9] == b"STREAMCSC"
    csc = nf.newStreamCSCDataset(fxT)
    cref = oracle.csr_to_csc(CSR(ds.data, ds.indices, ds.indptr, ds.nSamples, ds.nFeatures))
    assert np.array_equal(csc.indptr, cref.indptr) and np.array_equal(csc.indices, cref.indices)
    assert np.array_equal(csc.data, cref.data) and csc.shape == ds.shape
    with pytest.raises(IOError):
        nf.newStreamCSCDataset(fx)                         # wrong magic for the requested kind
    open(str(tmp_path / "bad.bin"), "wb").write(raw[:200])
    with pytest.raises(ValueError, match="truncated"):
        nf.newStreamCSRDataset(str(tmp_path / "bad.bin"))
    # the stream dataset drives the same kernels
    rng = np.random.default_rng(3)
    fm = nf.newFactorizationMachine(nf.regression, degree=2, nComponents=3, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = rng.standard_normal((1, 3, ds.nFeatures)) * 0.1, np.zeros(ds.nFeatures), 0.0, True
    want = oracle.fm_decision_function(CSR(ds.data, ds.indices, ds.indptr, ds.nSamples, ds.nFeatures), fm.P, fm.w, 0.0, 2)
    got = fm.decisionFunction(ds)
    # (rows with one nonzero have an exactly-zero ANOVA term on the CPU; FMA contraction leaves ~1e-19 there)
    from helpers import max_rel
    assert max_rel(got, want) < 1e-10


@pytest.mark.parametrize("cache_kb", [1.5, 9.0, 10_000.0])
def test_stream_windowed_dataset(oracle, tmp_path, cache_kb):
    """StreamCSRDataset in its windowed form (the reference's cacheSize window with HBM as the cache,
    sparse_stream.nim:232-270): the file stays on disk and decisionFunction / SGD / AdaGrad / MBPSGD walk
    resident windows of rows in file order -- same results as the whole matrix resident (no shuffling, as
    the reference with nCached < nSamples) and as the oracle.  cache_kb: windows of ~5 rows, ~30 rows, one."""
    import ctypes as C
    from nimfm_b200 import _lib
    from helpers import max_rel, make_fm_params
    n, d, k = 157, 40, 8
    rng = np.random.default_rng(21)
    X = make_dense(n, d, 4, density=0.3, positive=False)
    X[11] = 0.0
    csr = CSR.from_dense(X)
    y = np.sign(rng.standard_normal(n))
    txt, fx, fy = (str(tmp_path / f) for f in ("w.svm", "w.bin", "w.lab"))
    write_svm(txt, csr, y)
    nf.convertSVMLightFile(txt, fx, fy)
    whole = nf.newStreamCSRDataset(fx, resident=True)
    ref = CSR(whole.data, whole.indices, whole.indptr, whole.nSamples, whole.nFeatures)
    dd = whole.nFeatures
    win = nf.newStreamCSRDataset(fx, cacheSize=cache_kb / 1024.0, resident=False)
    assert win.windowed and win.shape == whole.shape and win.nnz == whole.nnz
    # windows tile the rows, each within the cache budget (or a single row), bit-exact content
    spans = []
    for a, b, w_ in win.windows():
        spans.append((a, b))
        sub = oracle.csr_take_rows(ref, np.arange(a, b))
        assert np.array_equal(w_.indptr, sub.indptr) and np.array_equal(w_.indices, sub.indices)
        assert np.array_equal(w_.data, sub.data)
        payload = 8 * (b - a) + 16 * len(sub.data)
        assert payload <= cache_kb * 1024 or b - a == 1
    assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    assert (len(spans) == 1) == (cache_kb > 1000) and win.nCached == spans[0][1]
    with pytest.raises(ValueError, match="windows"):
        win.handle()

    def model(degree=3):
        P, w, _ = make_fm_params(dd, degree, k, "explicit", True, seed=3, scale=0.1)
        fm = nf.newFactorizationMachine(nf.classification, degree=degree, nComponents=k, warmStart=True)
        fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), 0.05, True
        return fm, P, w
    fm, P, w = model()
    assert np.array_equal(fm.decisionFunction(win), fm.decisionFunction(whole))
    assert max_rel(fm.decisionFunction(win), oracle.fm_decision_function(ref, P, w, 0.05, 3)) <= 1e-10

    # per-sample solvers: the sequential loop simply continues across windows
    for make in (lambda: nf.newSGD(maxIter=3, eta0=0.05, verbose=0, tol=0.0, shuffle=True, loss=nf.Logistic()),
                 lambda: nf.newAdaGrad(maxIter=3, eta0=0.1, verbose=0, tol=0.0, shuffle=True, loss=nf.Logistic()),
                 lambda: nf.newAdaGrad(maxIter=3, eta0=0.1, verbose=0, tol=0.0, shuffle=True, loss=nf.Logistic(),
                                       miniBatchSize=8)):
        fa, _, _ = model()
        oa = make()
        oa.fit(win, y, fa)
        fb, _, _ = model()
        ob = make()
        ob.shuffle = False
        ob.fit(whole, y, fb)
        assert np.allclose(oa.history, ob.history, rtol=1e-10, atol=0) and oa.it == ob.it
        assert max_rel(fa.P, fb.P) <= 1e-10 and max_rel(fa.w, fb.w) <= 1e-10

    # MBPSGD: minibatches of 7 rows (157 = 22*7 + 3: the cursor wraps and every epoch starts elsewhere)
    kw = dict(eta0=0.2, alpha0=1e-6, alpha=1e-3, beta=1e-3, gamma=0.0)
    fa, P2, w2 = model(2)
    oa = nf.newMBPSGD(maxIter=4, loss=nf.Logistic(), miniBatchSize=7, verbose=0, tol=0.0, shuffle=True, **kw)
    oa.fit(win, y, fa)
    assert oa.shuffle is True
    oref = oracle.mbpsgd_fit(ref, y, P2, w2, 0.05, 2, "logistic", True, True, max_iter=4, reg="identity",
                             mini_batch_size=7, it=0, **kw)
    np.testing.assert_allclose(oa.history, oref["epoch_loss"], rtol=1e-8)
    assert max_rel(fa.P, oref["P"]) <= 1e-8 and max_rel(fa.w, oref["w"]) <= 1e-8 and oa.it == oref["it"]
    win.close()


@pytest.mark.parametrize("cache_kb", [1, 100000])
def test_field_stream_files(oracle, tmp_path, cache_kb):
    """STREAMCSRFIELD files (convertFFMFile dataset.nim:1202-1297, newStreamCSRFieldDataset, transposeFieldFile
    :1302-1402): the binary file loads to exactly what loadFFMFile reads from the text (ids / fields / indptr bit for
    bit), windows tile the rows, and the FFM paths -- decisionFunction, SGD, AdaGrad -- give the same results from
    the windowed file as from the resident dataset"""
    from helpers import max_rel
    X, csr, field_of = make_field_csr(83, 14, 4, 9)
    y = np.sign(np.random.default_rng(2).standard_normal(83))
    ds0 = nf.newCSRFieldDataset(csr.data, csr.indices, csr.indptr, csr.fields, csr.n, csr.d, csr.n_fields)
    txt, fx, fy, ft, fb = (str(tmp_path / f) for f in ("a.ffm", "a.bin", "a.lab", "a.csc", "a.back"))
    nf.dumpFFMFile(txt, ds0, y)                        # 1-based text
    nf.convertFFMFile(txt, fx, fy)
    text_ds, text_y = nf.loadFFMFile(txt)
    whole = nf.newStreamCSRFieldDataset(fx, resident=True)
    assert whole.shape == text_ds.shape and whole.nFields == text_ds.nFields == 4
    for name in ("indptr", "indices", "fields", "data"):
        assert np.array_equal(getattr(whole, name), getattr(text_ds, name)), name
    assert np.array_equal(nf.loadStreamLabel(fy), text_y)
    # transpose there and back: stable both ways, so the round trip is the identity on the file
    nf.transposeFieldFile(fx, ft)
    nf.transposeFieldFile(ft, fb)
    assert open(fb, "rb").read() == open(fx, "rb").read()
    assert open(ft, "rb").read()[:14] == b"STREAMCSCFIELD"
    win = nf.newStreamCSRFieldDataset(fx, cacheSize=cache_kb / 1024.0, resident=False)
    assert win.windowed and win.nFields == 4 and win.shape == whole.shape
    spans = [(a, b) for a, b, _ in win.windows()]
    assert spans[0][0] == 0 and spans[-1][1] == 83 and all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    assert (len(spans) == 1) == (cache_kb > 1000)
    rng = np.random.default_rng(3)
    P0, w0 = rng.standard_normal((4, whole.nFeatures, 4)) * 0.1, rng.standard_normal(whole.nFeatures) * 0.1

    def model():
        m = nf.newFieldAwareFactorizationMachine(nf.classification, nComponents=4, warmStart=True)
        m.P, m.w, m.intercept, m.isInitialized = P0.copy(), w0.copy(), 0.05, True
        return m
    ref = CSR(whole.data, whole.indices, whole.indptr, whole.nSamples, whole.nFeatures, fields=whole.fields, n_fields=4)
    got = model().decisionFunction(win)
    assert np.array_equal(got, model().decisionFunction(whole))
    assert max_rel(got, oracle.ffm_decision_function(ref, P0, w0, 0.05)) <= 1e-10
    for make in (lambda: nf.newSGD(maxIter=2, eta0=0.05, verbose=0, tol=0.0, shuffle=False, loss=nf.Logistic()),
                 lambda: nf.newAdaGrad(maxIter=2, eta0=0.1, verbose=0, tol=0.0, shuffle=False, loss=nf.Logistic())):
        ma, mb = model(), model()
        oa, ob = make(), make()
        oa.fit(win, y, ma)
        ob.fit(whole, y, mb)
        assert np.allclose(oa.history, ob.history, rtol=1e-10, atol=0) and oa.it == ob.it
        assert max_rel(ma.P, mb.P) <= 1e-10 and max_rel(ma.w, mb.w) <= 1e-10
    win.close()

"""Host-side mirror of nimfm's dataset facade (dataset.nim:10-153, 406-427; tensor/sparse.nim:4-31).

A dataset keeps the reference's host representation (f64 data, i64 indices / indptr [/ fields]) and a
lazily created device twin owned by libnimfm_cuda.so.  Dummy-feature augmentation
(dataset.nim:91-113) is applied inside the kernels (the model carries nAugments), so the host
arrays never change.
"""
import ctypes as C
import os

import numpy as np

from . import _lib


class BaseDataset:
    kind = _lib.DS_CSR
    windowed = False      # True: a StreamCSRDataset kept on disk and walked window by window

    def __init__(self, data, indices, indptr, nSamples, nFeatures, fields=None, nFields=0):
        self.data = _lib.f64(data)
        self.indices = _lib.i64(indices)
        self.indptr = _lib.i64(indptr)
        self._n, self._d = int(nSamples), int(nFeatures)
        self.fields = None if fields is None else _lib.i64(fields)
        self._nFields = int(nFields)
        self._handle = None
        self._y_id = None
        nseg = self._d if self.kind == _lib.DS_CSC else self._n
        if len(self.indptr) != nseg + 1:
            raise ValueError("len(indptr) != number of segments + 1")
        if len(self.data) != len(self.indices) or len(self.data) != int(self.indptr[-1]):
            raise ValueError("len(data), len(indices) and indptr[^1] disagree")

    # dataset.nim:40-61
    @property
    def nSamples(self):
        return self._n

    @property
    def nFeatures(self):
        return self._d

    @property
    def nFields(self):
        return self._nFields

    @property
    def shape(self):
        return (self._n, self._d)

    @property
    def nnz(self):
        return int(len(self.data))

    # ---- device twin
    def _upload(self):
        raise NotImplementedError

    def handle(self):
        if self._handle is None:
            self._upload()
        return self._handle

    def set_targets(self, y):
        y = _lib.f64(y)
        if len(y) != self._n:
            raise ValueError("len(y) != nSamples")
        _lib.check(_lib.load().nimfm_dataset_set_targets(_lib.ctx(), self.handle(), _lib.ptr(y)))

    def info(self):
        n, d, nnz, nf, mx = (C.c_int64() for _ in range(5))
        kind = C.c_int32()
        _lib.check(_lib.load().nimfm_dataset_info(self.handle(), C.byref(n), C.byref(d), C.byref(nnz),
                                                  C.byref(kind), C.byref(nf), C.byref(mx)))
        return dict(n=n.value, d=d.value, nnz=nnz.value, kind=kind.value, nFields=nf.value,
                    maxSegNnz=mx.value)

    def download(self):
        """Device arrays read back in the reference dtypes (bit-exact bookkeeping checks)."""
        inf = self.info()
        nseg = inf["d"] if inf["kind"] == _lib.DS_CSC else inf["n"]
        data = np.zeros(inf["nnz"])
        indices = np.zeros(inf["nnz"], np.int64)
        indptr = np.zeros(nseg + 1, np.int64)
        fields = np.zeros(inf["nnz"], np.int64) if inf["kind"] == _lib.DS_CSR_FIELD else None
        _lib.check(_lib.load().nimfm_dataset_download(_lib.ctx(), self.handle(), _lib.ptr(data),
                                                      _lib.ptr(indices), _lib.ptr(indptr),
                                                      _lib.ptr(fields)))
        return data, indices, indptr, fields

    # ---- X[indicesRow] / X[a..b] (dataset.nim:319-367; Nim slices are INCLUSIVE, Python's are not:
    # X[2:5] here == X[2..4] there)
    def __getitem__(self, key):
        if isinstance(key, slice):
            if key.step not in (None, 1):
                raise ValueError("only contiguous slices are supported")
            a = 0 if key.start is None else int(key.start)
            b = self._n if key.stop is None else int(key.stop)
            a = a + self._n if a < 0 else a
            b = b + self._n if b < 0 else b
            h = C.c_void_p()
            _lib.check(_lib.load().nimfm_dataset_slice_rows(_lib.ctx(), self.handle(), a, b - 1, C.byref(h)))
            return _adopt(h, type(self), self)
        idx = _lib.i64(key)
        if self.kind == _lib.DS_CSC:
            # tensor/sparse.nim:298-299: "Accessing row vectors by openarray is not supported for CSCMatrix"
            raise ValueError("Accessing row vectors by an index array is not supported for CSCDataset")
        if len(idx) == 0:
            raise ValueError("indicesRow is empty")      # max([]) raises in the reference (sparse.nim:255)
        h = C.c_void_p()
        _lib.check(_lib.load().nimfm_dataset_take_rows(_lib.ctx(), self.handle(), _lib.ptr(idx), len(idx), C.byref(h)))
        return _adopt(h, type(self), self)

    def free(self):
        if self._handle is not None:
            _lib.load().nimfm_dataset_free(_lib.ctx(), self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CSRDataset(BaseDataset):
    kind = _lib.DS_CSR

    def _upload(self, rowBegin=0, rowEnd=None):
        h = C.c_void_p()
        rowEnd = self._n if rowEnd is None else rowEnd
        _lib.check(_lib.load().nimfm_csr_upload(
            _lib.ctx(), self._n, self._d, _lib.ptr(self.data), _lib.ptr(self.indices),
            _lib.ptr(self.indptr), _lib.ptr(self.fields), self._nFields, rowBegin, rowEnd, C.byref(h)))
        self._handle = h

    def toCSCDataset(self):
        return toCSCDataset(self)


class CSRFieldDataset(CSRDataset):
    kind = _lib.DS_CSR_FIELD


class CSCDataset(BaseDataset):
    kind = _lib.DS_CSC

    def _upload(self):
        h = C.c_void_p()
        _lib.check(_lib.load().nimfm_csc_upload(_lib.ctx(), self._n, self._d, _lib.ptr(self.data),
                                                _lib.ptr(self.indices), _lib.ptr(self.indptr),
                                                C.byref(h)))
        self._handle = h

    def toCSRDataset(self):
        return toCSRDataset(self)


# dataset.nim:116-153
def newCSRDataset(data, indices, indptr, nSamples, nFeatures):
    return CSRDataset(data, indices, indptr, nSamples, nFeatures)


def newCSCDataset(data, indices, indptr, nSamples, nFeatures):
    return CSCDataset(data, indices, indptr, nSamples, nFeatures)


def newCSRFieldDataset(data, indices, indptr, fields, nSamples, nFeatures, nFields):
    return CSRFieldDataset(data, indices, indptr, nSamples, nFeatures, fields=fields, nFields=nFields)


def _adopt(h, cls, like):
    """Wrap a library-made dataset handle: host arrays are read back so it behaves like any other."""
    out = cls.__new__(cls)
    out._handle = h
    out._y_id = None
    inf = BaseDataset.info(out)
    out._n, out._d, out._nFields = inf["n"], inf["d"], inf["nFields"]
    out.data, out.indices, out.indptr, out.fields = BaseDataset.download(out)
    return out


def vstack(*datasets):
    """vstack (dataset.nim:452-483 -> tensor/sparse.nim:564-640)"""
    if len(datasets) == 1 and isinstance(datasets[0], (list, tuple)):
        datasets = tuple(datasets[0])
    if not datasets:
        raise ValueError("vstack needs at least one dataset")
    arr = (C.c_void_p * len(datasets))(*[X.handle() for X in datasets])
    h = C.c_void_p()
    _lib.check(_lib.load().nimfm_dataset_vstack(_lib.ctx(), arr, len(datasets), C.byref(h)))
    return _adopt(h, type(datasets[0]), datasets[0])


def shuffle(X, y, indices=None, rng=None):
    """shuffle(X, y[, indices]) (dataset.nim:372-395): X[indices], y[indices].  Without `indices` a
    permutation is drawn from `rng` (numpy Generator; Nim's RNG stream is unpinned)."""
    y = np.asarray(y)
    if indices is None:
        rng = np.random.default_rng() if rng is None else rng
        indices = rng.permutation(X.nSamples)
    indices = _lib.i64(indices)
    return X[indices], y[indices]


def _transposed(src, cls):
    """toCSCDataset / toCSRDataset (dataset.nim:406-427 -> tensor/sparse.nim:490-527): the stable
    counting-sort transpose performed by the library; the result is read back so the new dataset has
    host arrays like any other."""
    h = C.c_void_p()
    _lib.check(_lib.load().nimfm_dataset_transpose(_lib.ctx(), src.handle(), C.byref(h)))
    out = cls.__new__(cls)
    out._n, out._d, out._nFields, out.fields, out._y_id = src._n, src._d, 0, None, None
    out._handle = h
    out.data, out.indices, out.indptr, _ = BaseDataset.download(out)
    return out


def toCSCDataset(X):
    if isinstance(X, CSRDataset):
        return _transposed(X, CSCDataset)
    return _from_dense(np.asarray(X, dtype=np.float64), CSCDataset)


def toCSRDataset(X):
    if isinstance(X, CSCDataset):
        return _transposed(X, CSRDataset)
    return _from_dense(np.asarray(X, dtype=np.float64), CSRDataset)


def _from_dense(X, cls):
    """toCSRDataset / toCSCDataset from seq[seq[float64]] (dataset.nim:406-415)."""
    n, d = X.shape
    M = X.T if cls is CSCDataset else X
    data, indices, indptr = [], [], [0]
    for row in M:
        nz = np.nonzero(row)[0]
        indices.extend(nz.tolist())
        data.extend(row[nz].tolist())
        indptr.append(len(indices))
    return cls(data, indices, indptr, n, d)


# ---------------------------------------------------------------- text formats (dataset.nim:562-990)
def _loaded(h, cls):
    out = _adopt(h, cls, None)
    y = np.zeros(out.nSamples)
    _lib.check(_lib.load().nimfm_dataset_get_targets(_lib.ctx(), h, _lib.ptr(y)))
    return out, y


def _path(f):
    return os.path.expanduser(str(f)).encode()


def loadSVMLightFile(f, nFeatures=-1, kind="csr"):
    """loadSVMLightFile (dataset.nim:616-693): returns (CSRDataset | CSCDataset, y).  The file is parsed
    by the library on all host cores and uploaded directly; 1-based indices unless a 0 occurs."""
    h = C.c_void_p()
    csc = {"csr": 0, "csc": 1}[kind]
    _lib.check(_lib.load().nimfm_load_svmlight(_lib.ctx(), _path(f), int(nFeatures), csc, C.byref(h)))
    return _loaded(h, CSCDataset if csc else CSRDataset)


def loadFFMFile(f, nFeatures=-1, nFields=-1):
    """loadFFMFile (dataset.nim:768-790): returns (CSRFieldDataset, y)"""
    h = C.c_void_p()
    _lib.check(_lib.load().nimfm_load_ffm(_lib.ctx(), _path(f), int(nFeatures), int(nFields), C.byref(h)))
    return _loaded(h, CSRFieldDataset)


def loadUserItemRatingFile(f, kind="csr"):
    """loadUserItemRatingFile (dataset.nim:840-990): one-hot user + item features, value 1.0"""
    h = C.c_void_p()
    csc = {"csr": 0, "csc": 1}[kind]
    _lib.check(_lib.load().nimfm_load_user_item_rating(_lib.ctx(), _path(f), csc, C.byref(h)))
    return _loaded(h, CSCDataset if csc else CSRDataset)


def _num(v):
    """Nim's `$` / fmt of a number: integers print without a fraction, floats as the shortest round-trip"""
    return str(int(v)) if isinstance(v, (int, np.integer)) else repr(float(v))


def dumpSVMLightFile(f, X, y):
    """dumpSVMLightFile (dataset.nim:793-805): 1-based "label j:val ...", no trailing newline"""
    if X.nSamples != len(y):
        raise ValueError("X.nSamples != len(y).")
    with open(os.path.expanduser(str(f)), "w") as out:
        for i in range(X.nSamples):
            s, e = int(X.indptr[i]), int(X.indptr[i + 1])
            out.write(_num(y[i]) + "".join(f" {int(X.indices[q]) + 1}:{_num(X.data[q])}" for q in range(s, e)))
            if i + 1 != X.nSamples:
                out.write("\n")


def dumpFFMFile(f, X, y):
    """dumpFFMFile (dataset.nim:825-837): 1-based "label field:j:val ..." """
    if X.nSamples != len(y):
        raise ValueError("X.nSamples != len(y).")
    with open(os.path.expanduser(str(f)), "w") as out:
        for i in range(X.nSamples):
            s, e = int(X.indptr[i]), int(X.indptr[i + 1])
            out.write(_num(y[i]) + "".join(
                f" {int(X.fields[q]) + 1}:{int(X.indices[q]) + 1}:{_num(X.data[q])}" for q in range(s, e)))
            if i + 1 != X.nSamples:
                out.write("\n")


# ---------------------------------------------------------------- binary stream files (dataset.nim:995-1200)
_MAGIC = {"csr": b"STREAMCSR", "csc": b"STREAMCSC"}


def _write_stream(f, kind, nRows, nCols, data, indices, indptr):
    """magic | header {nRows, nCols, nnz: int64; max, min: float64} | per segment: count int64,
    count x {val float64, id int64}  (tensor/sparse_stream.nim:3-33)"""
    nnz = len(data)
    mx = float(np.max(data)) if nnz else float(np.finfo(np.float64).min)      # low(float64) / high(float64)
    mn = float(np.min(data)) if nnz else float(np.finfo(np.float64).max)
    rec = np.zeros(nnz, dtype=[("val", "<f8"), ("id", "<i8")])
    rec["val"], rec["id"] = data, indices
    with open(os.path.expanduser(str(f)), "wb") as out:
        out.write(_MAGIC[kind])
        out.write(np.array([nRows, nCols, nnz], dtype="<i8").tobytes())
        out.write(np.array([mx, mn], dtype="<f8").tobytes())
        for s in range(len(indptr) - 1):
            a, b = int(indptr[s]), int(indptr[s + 1])
            out.write(np.int64(b - a).tobytes())
            out.write(rec[a:b].tobytes())


def convertSVMLightFile(fIn, fOutX, fOutY):
    """convertSVMLightFile (dataset.nim:1017-1097): svmlight text -> STREAMCSR binary (0-based ids,
    nFeatures = maxIndex - minIndex + 1 with minIndex starting at 1) + raw float64 labels"""
    data, indices, indptr, y = [], [], [0], []
    minIndex, maxIndex = 1, 0
    for line in open(os.path.expanduser(str(fIn))).read().split("\n"):
        tok = line.replace(":", " ").split()
        if not tok:
            continue
        y.append(float(tok[0]))
        for a in range(1, len(tok), 2):
            j = int(tok[a])
            minIndex, maxIndex = min(j, minIndex), max(j, maxIndex)
            indices.append(j)
            data.append(float(tok[a + 1]))
        indptr.append(len(indices))
    if minIndex < 0:
        raise ValueError("Negative index is included.")
    idx = np.array(indices, np.int64) - minIndex
    _write_stream(fOutX, "csr", len(y), maxIndex - minIndex + 1, np.array(data, np.float64), idx, indptr)
    np.array(y, dtype="<f8").tofile(os.path.expanduser(str(fOutY)))


_MAGIC_FIELD = {"csr": b"STREAMCSRFIELD", "csc": b"STREAMCSCFIELD"}


def _write_field_stream(f, kind, nRows, nCols, nFields, data, indices, fields, indptr):
    """magic (14 bytes) | header {nRows, nCols, nnz, nFields: int64; max, min: float64} | per segment: count int64,
    count x {field int64, val float64, id int64}  (tensor/sparse_stream.nim:6-8,15-18,27-33)"""
    nnz = len(data)
    mx = float(np.max(data)) if nnz else float(np.finfo(np.float64).min)
    mn = float(np.min(data)) if nnz else float(np.finfo(np.float64).max)
    rec = np.zeros(nnz, dtype=[("field", "<i8"), ("val", "<f8"), ("id", "<i8")])
    rec["field"], rec["val"], rec["id"] = fields, data, indices
    with open(os.path.expanduser(str(f)), "wb") as out:
        out.write(_MAGIC_FIELD[kind])
        out.write(np.array([nRows, nCols, nnz, nFields], dtype="<i8").tobytes())
        out.write(np.array([mx, mn], dtype="<f8").tobytes())
        for s in range(len(indptr) - 1):
            a, b = int(indptr[s]), int(indptr[s + 1])
            out.write(np.int64(b - a).tobytes())
            out.write(rec[a:b].tobytes())


def _read_field_stream(f):
    """(kind, nRows, nCols, nFields, data, indices, fields, indptr) of a STREAMCSRFIELD / STREAMCSCFIELD file"""
    raw = open(os.path.expanduser(str(f)), "rb").read()
    kind = {v: k for k, v in _MAGIC_FIELD.items()}.get(raw[:14])
    if kind is None:
        raise IOError(f"{f} is not neither StreamCSCField nor StreamCSRField file.")     # dataset.nim:1320-1322
    nRows, nCols, nnz, nFields = (int(v) for v in np.frombuffer(raw[14:46], dtype="<i8"))
    nSeg = nRows if kind == "csr" else nCols
    data, indices, fields, indptr = np.zeros(nnz), np.zeros(nnz, np.int64), np.zeros(nnz, np.int64), [0]
    off, q = 62, 0
    for _ in range(nSeg):
        cnt = int(np.frombuffer(raw[off:off + 8], dtype="<i8")[0])
        rec = np.frombuffer(raw[off + 8:off + 8 + 24 * cnt], dtype=[("field", "<i8"), ("val", "<f8"), ("id", "<i8")])
        data[q:q + cnt], indices[q:q + cnt], fields[q:q + cnt] = rec["val"], rec["id"], rec["field"]
        q += cnt
        off += 8 + 24 * cnt
        indptr.append(q)
    return kind, nRows, nCols, nFields, data, indices, fields, np.array(indptr, np.int64)


def convertFFMFile(fIn, fOutX, fOutY):
    """convertFFMFile (dataset.nim:1202-1297): libffm text "label field:j:val ..." -> STREAMCSRFIELD binary + raw
    float64 labels; nFeatures = maxIndex - minIndex + 1 and nFields = maxField - minField + 1 with both minima
    starting at 1 (:1209-1212,1254-1255), ids rebased by minIndex (:1293).  The reference writes the FIELD
    unrebased (:1288) and then uses it as a 0-based index -- out of range for a 1-based file -- so the field is
    rebased by minField here, as loadFFMFile does (dataset.nim:768-790)."""
    data, indices, fields, indptr, y = [], [], [], [0], []
    minIndex, maxIndex, minField, maxField = 1, 0, 1, 0
    for line in open(os.path.expanduser(str(fIn))).read().split("\n"):
        tok = line.split()
        if not tok:
            continue
        y.append(float(tok[0]))
        for t in tok[1:]:
            fld, j, v = t.split(":")
            fld, j = int(fld), int(j)
            minField, maxField = min(fld, minField), max(fld, maxField)
            minIndex, maxIndex = min(j, minIndex), max(j, maxIndex)
            fields.append(fld)
            indices.append(j)
            data.append(float(v))
        indptr.append(len(indices))
    if minField < 0:
        raise ValueError("Negative field index is included.")
    if minIndex < 0:
        raise ValueError("Negative index is included.")
    _write_field_stream(fOutX, "csr", len(y), maxIndex - minIndex + 1, maxField - minField + 1,
                        np.array(data, np.float64), np.array(indices, np.int64) - minIndex,
                        np.array(fields, np.int64) - minField, indptr)
    np.array(y, dtype="<f8").tofile(os.path.expanduser(str(fOutY)))


def transposeFieldFile(fIn, fOut, cacheSize=200):
    """transposeFieldFile (dataset.nim:1302-1402): STREAMCSRFIELD <-> STREAMCSCFIELD, the header keeping
    [nRows, nCols, nFields]; a stable sort by the other axis (entries of one output segment stay in input-segment
    order, like the reference's counting passes), the field travelling with its element"""
    kind, nRows, nCols, nFields, data, indices, fields, indptr = _read_field_stream(fIn)
    nSegOut = nCols if kind == "csr" else nRows
    seg = np.repeat(np.arange(len(indptr) - 1, dtype=np.int64), np.diff(indptr))
    order = np.argsort(indices, kind="stable")
    optr = np.concatenate([[0], np.cumsum(np.bincount(indices, minlength=nSegOut))]).astype(np.int64)
    _write_field_stream(fOut, "csc" if kind == "csr" else "csr", nRows, nCols, nFields, data[order], seg[order],
                        fields[order], optr)


def transposeFile(fIn, fOut, cacheSize=200):
    """transposeFile (dataset.nim:1100-1200): STREAMCSR <-> STREAMCSC (the header keeps [nRows, nCols]);
    done with the library's stable device transpose instead of the reference's windowed file passes"""
    X, _ = _load_stream(fIn, None)
    T = _transposed(X, CSCDataset if isinstance(X, CSRDataset) else CSRDataset)
    kind = "csc" if isinstance(T, CSCDataset) else "csr"
    _write_stream(fOut, kind, T.nSamples, T.nFeatures, T.data, T.indices, T.indptr)


def _load_stream(fX, fY):
    h = C.c_void_p()
    _lib.check(_lib.load().nimfm_load_stream(_lib.ctx(), _path(fX), None if fY is None else _path(fY), C.byref(h)))
    probe = BaseDataset.__new__(BaseDataset)
    probe._handle = h
    kind = BaseDataset.info(probe)["kind"]
    cls = {_lib.DS_CSC: _DeviceCSCDataset, _lib.DS_CSR_FIELD: _DeviceCSRFieldDataset}.get(kind, _DeviceCSRDataset)
    probe._handle = None
    out = cls(h)
    y = None
    if fY is not None:
        y = np.zeros(out.nSamples)
        _lib.check(_lib.load().nimfm_dataset_get_targets(_lib.ctx(), h, _lib.ptr(y)))
    return out, y


class _DeviceMixin:
    """A library-made dataset that stays on the device: the host arrays are read back only if somebody asks
    for them (a window of a stream file is used once and freed; a 100 GB file is not mirrored on the host)."""

    def __init__(self, h):
        self._handle = h
        self._y_id = None
        inf = BaseDataset.info(self)
        self._n, self._d, self._nFields, self._nnz = inf["n"], inf["d"], inf["nFields"], inf["nnz"]
        self.fields = None

    @property
    def nnz(self):
        return self._nnz

    def __getattr__(self, name):
        if name in ("data", "indices", "indptr"):
            self.data, self.indices, self.indptr, f = BaseDataset.download(self)
            if f is not None:
                self.fields = f
            return self.__dict__[name]
        raise AttributeError(name)


class _DeviceCSRDataset(_DeviceMixin, CSRDataset):
    pass


class _DeviceCSRFieldDataset(_DeviceMixin, CSRFieldDataset):
    def __init__(self, h):
        _DeviceMixin.__init__(self, h)
        self.data, self.indices, self.indptr, self.fields = BaseDataset.download(self)


class _DeviceCSCDataset(_DeviceMixin, CSCDataset):
    pass


class StreamCSRDataset(CSRDataset):
    """newStreamCSRDataset (dataset.nim:170-173; StreamCSRDataset, dataset.nim:1017-1402) in its windowed
    form: the file stays on disk, `windows()` yields one resident window of rows after another (HBM is the
    cache; cacheSize MB of file payload per window, as in the reference).  decisionFunction and the
    SGD / AdaGrad / MBPSGD fits walk the windows in file order; like the reference with nCached < nSamples
    they do not shuffle (sgd.nim:297, minibatch_psgd.nim:110-111)."""
    windowed = True

    def __init__(self, f, cacheSize=200, fY=None):
        self._path = _path(f)
        self._sh = C.c_void_p()
        lib = _lib.load()
        _lib.check(lib.nimfm_stream_open(_lib.ctx(), self._path, None if fY is None else _path(fY), C.byref(self._sh)))
        kind, n, d, nnz, mx, pay = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(lib.nimfm_stream_info(self._sh, C.byref(kind), C.byref(n), C.byref(d), C.byref(nnz), C.byref(mx),
                                         C.byref(pay)))
        if kind.value != self.kind:
            self.close()
            raise IOError(f"{f} is not a {'StreamCSRField' if self.kind == _lib.DS_CSR_FIELD else 'StreamCSR'} file.")
        self._n, self._d, self._nnz, self._nFields = n.value, d.value, nnz.value, 0
        self.payloadBytes = pay.value
        self.cacheBytes = max(int(cacheSize * 1024 * 1024), 1)
        self._handle = None
        self._y = None
        self.fields = None

    @property
    def nnz(self):
        return self._nnz

    @property
    def nCached(self):
        """rows of the first window (dataset.nim's nCached: == nSamples iff the file fits the cache)"""
        return self.window_end(0) if self._n else 0

    def window_end(self, rowBegin, cacheBytes=None):
        return int(_lib.load().nimfm_stream_window_end(self._sh, int(rowBegin),
                                                       int(self.cacheBytes if cacheBytes is None else cacheBytes)))

    def set_targets(self, y):
        y = _lib.f64(y)
        if len(y) != self._n:
            raise ValueError("len(y) != nSamples")
        self._y = y

    def load_rows(self, a, b):
        """rows [a, b) as a resident dataset (targets attached when set_targets was called)"""
        h = C.c_void_p()
        _lib.check(_lib.load().nimfm_stream_load_window(_lib.ctx(), self._sh, int(a), int(b), C.byref(h)))
        win = self._window_cls(h)
        if self._y is not None and b > a:
            win.set_targets(self._y[a:b])
        return win

    _window_cls = _DeviceCSRDataset

    def windows(self, multiple=1, start=0):
        """(rowBegin, rowEnd, dataset) over [start, nSamples) in file order; every window but the last holds
        a multiple of `multiple` rows (minibatches never straddle a window)"""
        a = int(start)
        while a < self._n:
            b = self.window_end(a)
            if multiple > 1 and b < self._n:
                b = a + max((b - a) // multiple, 1) * multiple
                b = min(b, self._n)
            win = self.load_rows(a, b)
            try:
                yield a, b, win
            finally:
                win.free()
            a = b

    def handle(self):
        raise ValueError("a windowed StreamCSRDataset has no single device twin; iterate windows()")

    def __getitem__(self, key):
        raise ValueError("row access on a windowed StreamCSRDataset is not supported; use load_rows(a, b)")

    def close(self):
        if getattr(self, "_sh", None):
            _lib.load().nimfm_stream_close(self._sh)
            self._sh = None

    free = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StreamCSRFieldDataset(StreamCSRDataset):
    """newStreamCSRFieldDataset (dataset.nim:1402+; StreamCSRFieldMatrix, tensor/sparse_stream.nim:54-56,164-196): a
    STREAMCSRFIELD file kept on disk and walked window by window, the FFM solvers' row dataset"""
    kind = _lib.DS_CSR_FIELD

    def __init__(self, f, cacheSize=200, fY=None):
        StreamCSRDataset.__init__(self, f, cacheSize, fY)
        with open(self._path.decode() if isinstance(self._path, bytes) else self._path, "rb") as fh:
            fh.seek(14 + 24)
            self._nFields = int(np.frombuffer(fh.read(8), dtype="<i8")[0])

    _window_cls = _DeviceCSRFieldDataset


def newStreamCSRFieldDataset(f, cacheSize=200, resident=None):
    """newStreamCSRFieldDataset: resident (whole file on the device) whenever the payload is at most a third of
    the free device memory, else windowed -- as newStreamCSRDataset"""
    if resident is None:
        free = C.c_int64()
        _lib.check(_lib.load().nimfm_mem_info(_lib.ctx(), C.byref(free), None))
        resident = os.path.getsize(_path(f)) * 3 <= free.value
    if not resident:
        return StreamCSRFieldDataset(f, cacheSize)
    X, _ = _load_stream(f, None)
    if not isinstance(X, CSRFieldDataset):
        raise IOError(f"{f} is not a StreamCSRField file.")
    return X


def newStreamCSRDataset(f, cacheSize=200, resident=None):
    """newStreamCSRDataset (dataset.nim:170-173).  The reference windows the file through a cacheSize-MB
    cache.  Here HBM is the cache: resident=True loads the whole matrix onto the device, resident=False
    keeps the file on disk and walks cacheSize-MB windows of rows (StreamCSRDataset), resident=None picks
    the resident form whenever the file's payload is at most a third of the free device memory."""
    if resident is None:
        free = C.c_int64()
        _lib.check(_lib.load().nimfm_mem_info(_lib.ctx(), C.byref(free), None))
        resident = os.path.getsize(_path(f)) * 3 <= free.value
    if not resident:
        return StreamCSRDataset(f, cacheSize)
    X, _ = _load_stream(f, None)
    if not isinstance(X, CSRDataset):
        raise IOError(f"{f} is not a StreamCSR file.")
    return X


def newStreamCSCDataset(f, cacheSize=200):
    """newStreamCSCDataset (dataset.nim:176-179)"""
    X, _ = _load_stream(f, None)
    if not isinstance(X, CSCDataset):
        raise IOError(f"{f} is not a StreamCSC file.")
    return X


def loadStreamLabel(fIn, nSamples=None):
    """loadStreamLabel (dataset.nim:995-1014): raw little-endian float64 targets"""
    y = np.fromfile(os.path.expanduser(str(fIn)), dtype="<f8")
    return y if nSamples is None else y[:nSamples].copy()

## device.nim -- what every replaced `fit` / `decisionFunction` shares: the process-wide context, error mapping,
## loss / scheduling / regulariser codes, the device twin of a dataset (cached per dataset object, its host seqs
## page-locked once) and of a model (parameters cross the ABI in nimfm's own layouts).
##
## NOT COMPILED IN THIS REPOSITORY (no Nim toolchain in the build image, SURVEY.md): the same call sequences run,
## under test, through the ctypes mirror nimfm_b200/{dataset,model,optimizers}.py.  tests/test_nim_layer_cpu.py
## checks that every `fit` / `decisionFunction` signature here equals the reference's exported one.
##
## Place this directory at src/nimfm/cuda/ of a nimfm checkout (INTEGRATION.md section 2).
import tables, math, os, strutils
import nimfm_cuda
import ../dataset, ../tensor/tensor, ../tensor/sparse
import ../model/factorization_machine, ../model/field_aware_factorization_machine
import ../loss
from ../optimizer/sgd import SchedulingKind

# ---------------------------------------------------------------- context and errors
var gCtx: Ctx

proc ctx*(): Ctx =
  ## one context per process (one process per GPU: LOCAL_RANK picks the device, as torchrun exports it)
  if gCtx.isNil:
    var dev = 0'i32
    try: dev = int32(parseInt(getEnv("LOCAL_RANK", "0")))
    except ValueError: dev = 0
    if nimfm_ctx_create(dev, addr gCtx) != 0:
      raise newException(IOError, "libnimfm_cuda: " & $nimfm_last_error(nil) & " (there is no CPU fallback)")
  result = gCtx

template check*(rc: int32) =
  ## status -> exception, in the reference's style (factorization_machine.nim:114-115 raises ValueError)
  let rcv = rc
  if rcv == -1: raise newException(ValueError, $nimfm_last_error(ctx()))
  elif rcv != 0: raise newException(IOError, "libnimfm_cuda error " & $rcv & ": " & $nimfm_last_error(ctx()))

proc p*[T](s: var seq[T]): ptr T {.inline.} = (if s.len == 0: nil else: addr s[0])

# ---------------------------------------------------------------- codes of include/nimfm_cuda.h
proc lossKind*[L](loss: L): int32 =
  ## loss.nim:3-12
  when L is Squared: 0
  elif L is SquaredHinge: 1
  elif L is Logistic: 2
  elif L is Huber: 3
  else: {.error: "loss type not supported on the device path".}

proc lossThreshold*[L](loss: L): float64 =
  when L is Huber: loss.getThreshold()   # `threshold` is private in loss.nim:11-12: add `proc getThreshold*`
  else: 1.0

proc schedCode*(s: SchedulingKind): int32 =
  ## optimizer/sgd.nim:8-12
  case s
  of constant: 0
  of optimal: 1
  of invscaling: 2
  of pegasos: 3

# ---------------------------------------------------------------- datasets
type
  DevEntry = object
    handle: DeviceDataset
    registered: seq[pointer]      # host seqs page-locked with nimfm_host_register
    nnz: int

var gDevCache = initTable[pointer, DevEntry]()   # keyed by the dataset ref: the device twin lives as long as it

proc pinHost[T](e: var DevEntry, s: var seq[T]) =
  ## page-lock a dataset's seq once, so that the host-fed calls DMA straight out of it
  if s.len > 0 and nimfm_host_register(ctx(), addr s[0], int64(s.len * sizeof(T))) == 0:
    e.registered.add(cast[pointer](addr s[0]))

proc device*(X: CSRDataset): DeviceDataset =
  ## newCSRDataset (dataset.nim:116-122): data / indices / indptr are the public seqs of tensor/sparse.nim:4-31;
  ## uploaded once per dataset object (the library copies; the caller keeps ownership)
  let key = cast[pointer](X)
  if key in gDevCache and gDevCache[key].nnz == X.nnz: return gDevCache[key].handle
  var e: DevEntry
  check nimfm_csr_upload(ctx(), X.nSamples, X.nFeatures, cast[ptr cdouble](p(X.data.data)),
                         cast[ptr int64](p(X.data.indices)), cast[ptr int64](p(X.data.indptr)), nil, 0,
                         0, X.nSamples, addr e.handle)
  e.nnz = X.nnz
  gDevCache[key] = e
  result = e.handle

proc device*(X: CSCDataset): DeviceDataset =
  ## newCSCDataset (dataset.nim:125-131)
  let key = cast[pointer](X)
  if key in gDevCache and gDevCache[key].nnz == X.nnz: return gDevCache[key].handle
  var e: DevEntry
  check nimfm_csc_upload(ctx(), X.nSamples, X.nFeatures, cast[ptr cdouble](p(X.data.data)),
                         cast[ptr int64](p(X.data.indices)), cast[ptr int64](p(X.data.indptr)), addr e.handle)
  e.nnz = X.nnz
  gDevCache[key] = e
  result = e.handle

proc device*(X: CSRFieldDataset): DeviceDataset =
  ## newCSRFieldDataset (dataset.nim:134-153)
  let key = cast[pointer](X)
  if key in gDevCache and gDevCache[key].nnz == X.nnz: return gDevCache[key].handle
  var e: DevEntry
  check nimfm_csr_upload(ctx(), X.nSamples, X.nFeatures, cast[ptr cdouble](p(X.data.data)),
                         cast[ptr int64](p(X.data.indices)), cast[ptr int64](p(X.data.indptr)),
                         cast[ptr int64](p(X.data.fields)), X.nFields, 0, X.nSamples, addr e.handle)
  e.nnz = X.nnz
  gDevCache[key] = e
  result = e.handle

proc releaseDevice*[T](X: BaseDataset[T]) =
  ## drop the device twin (call before mutating the dataset's arrays, or to give the memory back)
  let key = cast[pointer](X)
  if key in gDevCache:
    for q in gDevCache[key].registered: discard nimfm_host_unregister(ctx(), q)
    discard nimfm_dataset_free(ctx(), gDevCache[key].handle)
    gDevCache.del(key)

proc setTargets*(ds: DeviceDataset, y: var seq[float64]) =
  check nimfm_dataset_set_targets(ctx(), ds, cast[ptr cdouble](p(y)))

# ---------------------------------------------------------------- models
proc flat*(T: Tensor): seq[float64] =
  ## Tensor is seq[Matrix] of ragged rows (tensor.nim:8-17): flatten to [a][b][c]
  result = newSeqOfCap[float64](T.shape[0] * T.shape[1] * T.shape[2])
  for a in 0..<T.shape[0]:
    for b in 0..<T.shape[1]:
      for c in 0..<T.shape[2]: result.add(T[a, b, c])

proc unflat*(T: var Tensor, s: seq[float64]) =
  var q = 0
  for a in 0..<T.shape[0]:
    for b in 0..<T.shape[1]:
      for c in 0..<T.shape[2]:
        T[a, b, c] = s[q]
        inc(q)

proc toDevice*(fm: FactorizationMachine, nFeatures: int): DeviceFM =
  ## P crosses the ABI as [nOrders, nComponents, nFeatures+nAugments] (factorization_machine.nim:33-36)
  if nFeatures + fm.nAugments != fm.P.shape[2]:
    raise newException(ValueError, "Invalid nFeatures.")          # factorization_machine.nim:114-115
  var P = flat(fm.P)
  check nimfm_fm_create(ctx(), int32(fm.degree), int32(fm.nComponents), int32(fm.nOrders), int32(fm.nAugments),
                        nFeatures, int32(fm.fitLinear), int32(fm.fitIntercept), addr result)
  check nimfm_fm_set_params(ctx(), result, cast[ptr cdouble](p(P)), cast[ptr cdouble](p(fm.w)), fm.intercept,
                            cast[ptr cdouble](p(fm.lams)))

proc fromDevice*(fm: FactorizationMachine, h: DeviceFM) =
  var P = newSeq[float64](fm.P.shape[0] * fm.P.shape[1] * fm.P.shape[2])
  var b: cdouble
  check nimfm_fm_get_params(ctx(), h, cast[ptr cdouble](p(P)), cast[ptr cdouble](p(fm.w)), addr b)
  unflat(fm.P, P)
  fm.intercept = b

proc toDevice*(ffm: FieldAwareFactorizationMachine, nFeatures, nFields: int): DeviceFFM =
  ## P crosses the ABI as [nFields, nFeatures, nComponents] (field_aware_factorization_machine.nim:16-17)
  if nFeatures != ffm.P.shape[1]: raise newException(ValueError, "Invalid nFeatures.")   # :60-61
  if nFields != ffm.P.shape[0]: raise newException(ValueError, "Invalid nFields.")       # :62-64
  var P = flat(ffm.P)
  check nimfm_ffm_create(ctx(), int32(ffm.nComponents), nFields, nFeatures, int32(ffm.fitLinear),
                         int32(ffm.fitIntercept), addr result)
  check nimfm_ffm_set_params(ctx(), result, cast[ptr cdouble](p(P)), cast[ptr cdouble](p(ffm.w)), ffm.intercept)

proc fromDevice*(ffm: FieldAwareFactorizationMachine, h: DeviceFFM) =
  var P = newSeq[float64](ffm.P.shape[0] * ffm.P.shape[1] * ffm.P.shape[2])
  var b: cdouble
  check nimfm_ffm_get_params(ctx(), h, cast[ptr cdouble](p(P)), cast[ptr cdouble](p(ffm.w)), addr b)
  unflat(ffm.P, P)
  ffm.intercept = b

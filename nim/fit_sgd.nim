## fit_sgd.nim -- SGD.fit on the device.  `include` at the end of optimizer/sgd.nim in place of
## `proc fit*[L](self: SGD[L], X: RowDataset, ...)` (sgd.nim:261-328) and of the Hogwild overload
## `fit(..., maxThreads, ...)` (sgd_multi.nim:40-120).  One library call runs step() over an epoch's samples
## (sgd.nim:296-300) inside one persistent thread block -- exact reference semantics, lazy scaling and
## resetScaling included; finalize (sgd.nim:99-113) is nimfm_fm_sgd_end.

proc fit*[L](self: SGD[L], X: RowDataset, y: seq[float64],
             fm: FactorizationMachine,
             callback: (SGD[L], FactorizationMachine)->void = nil) =
  ## Fits the factorization machine on X and y by stochastic gradient descent.
  fm.init(X)
  var y = fm.checkTarget(y)
  let
    nSamples = X.nSamples
    nCalls = self.nCalls
  var
    indices = toSeq(0..<nSamples)
    isConverged = false
  if nCalls > 0 and not callback.isNil:
    raise newException(ValueError, "nCalls > 0 is not supported on the device path: one call runs a whole epoch.")
  if not fm.warmstart:
    self.init()
  let ds = device(X)
  setTargets(ds, y)
  let h = toDevice(fm, X.nFeatures)
  var cfg = SgdCfg(loss: lossKind(self.loss), huberThreshold: lossThreshold(self.loss), eta0: self.eta0,
                   alpha0: self.alpha0, alpha: self.alpha, beta: self.beta,
                   scheduling: schedCode(self.scheduling), power: self.power)
  try:
    check nimfm_fm_sgd_begin(ctx(), h)          # scaling_w / scaling_P / scalings_* <- 1 (sgd.nim:269-272)
    for epoch in 0..<self.maxIter:
      var viol, lossSum: cdouble
      if X.nCached == X.nSamples and self.shuffle: shuffle(indices)
      var itc = int64(self.it)
      check nimfm_fm_sgd_epoch(ctx(), h, ds, addr cfg, addr itc, cast[ptr int64](p(indices)), nSamples,
                               addr viol, addr lossSum)
      self.it = int(itc)
      let runningLoss = lossSum / float(nSamples)
      if not callback.isNil and nCalls <= 0:
        check nimfm_fm_sgd_end(ctx(), h)          # finalize + transpose (sgd.nim:310-316)
        fromDevice(fm, h)
        callback(self, fm)
      elif self.verbose > 0:
        fromDevice(fm, h)                         # un-finalized, as the reference's verbose line sees it
      var Pt: Tensor = zeros([fm.P.shape[0], fm.P.shape[2], fm.P.shape[1]])
      transpose(Pt, fm.P)
      let isContinue = stoppingCriterion(
        Pt, fm.w, fm.intercept, self.alpha0, self.alpha, self.beta, runningLoss,
        viol, self.tol, self.verbose, epoch, self.maxIter, isConverged)
      if not isContinue: break
    if not isConverged and self.verbose > 0:
      echo("Objective did not converge. Increase maxIter.")
    check nimfm_fm_sgd_end(ctx(), h)              # finalize
    fromDevice(fm, h)
  finally:
    discard nimfm_fm_free(ctx(), h)

proc fit*[L](self: SGD[L], X: RowDataset, y: seq[float64],
             fm: FactorizationMachine, maxThreads: int,
             callback: (SGD[L], FactorizationMachine)->void = nil) =
  ## Fits the factorization machine on X and y by stochastic gradient descent.
  ## Hogwild in the reference (sgd_multi.nim:40-120); here its deterministic analogue: the maxThreads samples of a
  ## minibatch see the same parameters and their updates are applied at once (include/nimfm_cuda.h,
  ## nimfm_fm_sgd_minibatch_epoch); maxThreads < 0: 4096 resident rows.  The callback is called per epoch, as there.
  fm.init(X)
  var y = fm.checkTarget(y)
  let
    nSamples = X.nSamples
    miniBatch = (if maxThreads < 0: 4096 else: max(1, maxThreads))
  var
    indices = toSeq(0..<nSamples)
    isConverged = false
  if not fm.warmstart:
    self.init()
  let ds = device(X)
  setTargets(ds, y)
  let h = toDevice(fm, X.nFeatures)
  var cfg = SgdCfg(loss: lossKind(self.loss), huberThreshold: lossThreshold(self.loss), eta0: self.eta0,
                   alpha0: self.alpha0, alpha: self.alpha, beta: self.beta,
                   scheduling: schedCode(self.scheduling), power: self.power)
  try:
    for epoch in 0..<self.maxIter:
      var viol, lossSum: cdouble
      if X.nCached == X.nSamples and self.shuffle: shuffle(indices)
      var itc = int64(self.it)
      check nimfm_fm_sgd_minibatch_epoch(ctx(), h, ds, addr cfg, miniBatch, miniBatch, addr itc,
                                         cast[ptr int64](p(indices)), nSamples, addr viol, addr lossSum)
      self.it = int(itc)
      let runningLoss = lossSum / float(nSamples)
      if not callback.isNil or self.verbose > 0:
        fromDevice(fm, h)                         # the parameters are canonical between calls: no finalize needed
      if not callback.isNil: callback(self, fm)
      var Pt: Tensor = zeros([fm.P.shape[0], fm.P.shape[2], fm.P.shape[1]])
      transpose(Pt, fm.P)
      let isContinue = stoppingCriterion(
        Pt, fm.w, fm.intercept, self.alpha0, self.alpha, self.beta, runningLoss,
        viol, self.tol, self.verbose, epoch, self.maxIter, isConverged)
      if not isContinue: break
    if not isConverged and self.verbose > 0:
      echo("Objective did not converge. Increase maxIter.")
    fromDevice(fm, h)
  finally:
    discard nimfm_fm_free(ctx(), h)

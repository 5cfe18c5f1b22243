## fit_adagrad.nim -- AdaGrad.fit on the device.  `include` at the end of optimizer/adagrad.nim in place of
## `proc fit*[L](self: AdaGrad[L], X: RowDataset, ...)` (adagrad.nim:137-203) and of the Hogwild overload
## `fit(..., maxThreads, ...)` (adagrad_multi.nim:39-115).  One library call runs the sample loop of an epoch
## (adagrad.nim:170-181: lazy update(), predictWithGrad, updateG()); shuffling, the callback, stoppingCriterion and
## finalize stay here.

proc runAdaGrad[L](self: AdaGrad[L], X: RowDataset, y: seq[float64], fm: FactorizationMachine, miniBatch: int,
                   callback: (AdaGrad[L], FactorizationMachine)->void) =
  fm.init(X)
  var y = fm.checkTarget(y)
  let nSamples = X.nSamples
  var
    indices = toSeq(0..<nSamples)
    isConverged = false
  if self.nCalls > 0 and not callback.isNil:
    raise newException(ValueError, "nCalls > 0 is not supported on the device path: one call runs a whole epoch.")
  let ds = device(X)
  setTargets(ds, y)
  let h = toDevice(fm, X.nFeatures)
  var cfg = AdagradCfg(loss: lossKind(self.loss), huberThreshold: lossThreshold(self.loss), eta0: self.eta0,
                       alpha0: self.alpha0, alpha: self.alpha, beta: self.beta, eps: self.eps,
                       miniBatchSize: miniBatch)
  try:
    # init (adagrad.nim:47-62): g_sum <- 0, g_norm <- eps unless warm-started with a state of the right shape
    if not fm.warmStart: self.it = 1
    check nimfm_fm_adagrad_init(ctx(), h, self.eps, 1)
    if self.it != 1:
      var
        gsP = flat(self.g_sum.P)
        gnP = flat(self.g_norm.P)
      check nimfm_fm_adagrad_set_state(ctx(), h, cast[ptr cdouble](p(gsP)), cast[ptr cdouble](p(gnP)),
                                       cast[ptr cdouble](p(self.g_sum.w)), cast[ptr cdouble](p(self.g_norm.w)),
                                       self.g_sum.intercept, self.g_norm.intercept)
    for epoch in 0..<self.maxIter:
      var viol, lossSum: cdouble
      if X.nCached == X.nSamples and self.shuffle: shuffle(indices)
      var itc = int64(self.it)
      check nimfm_fm_adagrad_epoch(ctx(), h, ds, addr cfg, addr itc, cast[ptr int64](p(indices)), nSamples,
                                   addr viol, addr lossSum)
      self.it = int(itc)
      let runningLoss = lossSum / float(nSamples)
      if not callback.isNil:
        check nimfm_fm_adagrad_finalize(ctx(), h, addr cfg, self.it)       # finalize + transpose (adagrad.nim:185-188)
        fromDevice(fm, h)
        callback(self, fm)
      elif self.verbose > 0:
        fromDevice(fm, h)
      var Pt: Tensor = zeros([fm.P.shape[0], fm.P.shape[2], fm.P.shape[1]])
      transpose(Pt, fm.P)
      let isContinue = stoppingCriterion(
        Pt, fm.w, fm.intercept, self.alpha0, self.alpha, self.beta, runningLoss,
        viol, self.tol, self.verbose, epoch, self.maxIter, isConverged)
      if not isContinue: break
    if not isConverged and self.verbose > 0:
      echo("Objective did not converge. Increase maxIter.")
    # keep g_sum / g_norm for warm starts (adagrad.nim:15-16), then finalize (adagrad.nim:65-84)
    block:
      var
        gsP = newSeq[float64](fm.P.shape[0] * fm.P.shape[1] * fm.P.shape[2])
        gnP = newSeq[float64](gsP.len)
        gsb, gnb: cdouble
      if self.g_sum.isNil:
        self.g_sum = newParams([fm.P.shape[0], fm.P.shape[2], fm.P.shape[1]], fm.w.len, fm.fitLinear, fm.fitIntercept)
        self.g_norm = newParams([fm.P.shape[0], fm.P.shape[2], fm.P.shape[1]], fm.w.len, fm.fitLinear, fm.fitIntercept)
      check nimfm_fm_adagrad_get_state(ctx(), h, cast[ptr cdouble](p(gsP)), cast[ptr cdouble](p(gnP)),
                                       cast[ptr cdouble](p(self.g_sum.w)), cast[ptr cdouble](p(self.g_norm.w)),
                                       addr gsb, addr gnb)
      unflat(self.g_sum.P, gsP)
      unflat(self.g_norm.P, gnP)
      self.g_sum.intercept = gsb
      self.g_norm.intercept = gnb
    check nimfm_fm_adagrad_finalize(ctx(), h, addr cfg, self.it)
    fromDevice(fm, h)
  finally:
    discard nimfm_fm_free(ctx(), h)

proc fit*[L](self: AdaGrad[L], X: RowDataset, y: seq[float64],
             fm: FactorizationMachine,
             callback: (AdaGrad[L], FactorizationMachine)->void = nil) =
  ## Fits the factorization machine on X and y by stochastic gradient descent.
  ## miniBatchSize = 1: the reference's strictly sequential semantics (adagrad.nim:164-181).
  runAdaGrad(self, X, y, fm, 1, callback)

proc fit*[L](self: AdaGrad[L], X: RowDataset, y: seq[float64],
             fm: FactorizationMachine, maxThreads: int,
             callback: (AdaGrad[L], FactorizationMachine)->void = nil) =
  ## Fits the factorization machine on X and y by stochastic gradient descent.
  ## The reference runs maxThreads lock-free Hogwild threads here (adagrad_multi.nim:39-115: every sample sees
  ## parameters up to ~maxThreads updates stale, results depend on thread timing); the device runs the deterministic
  ## analogue, a synchronous minibatch of maxThreads samples (maxThreads < 0: 4096 resident rows).
  runAdaGrad(self, X, y, fm, (if maxThreads < 0: 4096 else: max(1, maxThreads)), callback)

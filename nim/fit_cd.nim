## fit_cd.nim -- CD.fit on the device.  `include` this file at the end of optimizer/cd.nim IN PLACE OF the reference's
## `proc fit*[L](self: CD[L], ...)` (cd.nim:110-194): same signature, same loop -- intercept, linear term, one sweep per
## order, callback, verbose line, tol test -- with the sweeps of one outer iteration in ONE library call.
## (An include, not an import: the solver objects keep private fields.)  Needs `import ../cuda/[nimfm_cuda, device]`.

proc fit*[L](self: CD[L], X: ColDataset, y: seq[float64],
             fm: FactorizationMachine,
             callback: (CD[L], FactorizationMachine)->void = nil) =
  ## Fits the factorization machine on X and y by coordinate descent.
  fm.init(X)
  var y = fm.checkTarget(y)
  let ds = device(X)                         # CSC twin, uploaded once per dataset object
  setTargets(ds, y)
  let h = toDevice(fm, X.nFeatures)
  var cfg = CdCfg(loss: lossKind(self.loss), huberThreshold: lossThreshold(self.loss),
                  alpha0: self.alpha0, alpha: self.alpha, beta: self.beta)   # the library scales by nSamples (cd.nim:123-125)
  var isConverged = false
  try:
    # caches: yPred, A, colNormSq, scaled alphas (cd.nim:123-151)
    check nimfm_fm_cd_begin(ctx(), h, ds, addr cfg)
    if self.verbose > 0: echoHeader(self.maxIter)
    for it in 0..<self.maxIter:
      var viol, lossVal, reg: cdouble
      # intercept -> linear -> epoch / epochDeg2 per order (cd.nim:156-172); lossVal and reg as cd.nim:177-183
      check nimfm_fm_cd_epoch(ctx(), h, ds, addr cfg, addr viol, addr lossVal, addr reg)
      if not callback.isNil:
        fromDevice(fm, h)
        callback(self, fm)
      if self.verbose > 0: echoInfo(it+1, self.maxIter, viol, lossVal, reg)
      if viol < self.tol:
        if self.verbose > 0: echo("Converged at iteration ", it+1, ".")
        isConverged = true
        break
    if not isConverged and self.verbose > 0:
      echo("Objective did not converge. Increase maxIter.")
    check nimfm_fm_cd_end(ctx(), h)
    fromDevice(fm, h)
  finally:
    discard nimfm_fm_free(ctx(), h)

## fit_mbpsgd.nim -- MBPSGD.fit on the device.  `include` at the end of optimizer/minibatch_psgd.nim in place of
## `proc fit*[L, R](self: MBPSGD[L, R], ...)` (minibatch_psgd.nim:127-211).  The minibatch / inner-iteration sizes, the
## index shuffle and the cursor stay here (minibatch_psgd.nim:157-170, 102-111); one library call runs epoch()
## (:91-124): K2 per minibatch, the dense step (params.nim:90-98), the prox (:119-121).

proc regCode[R](reg: R): int32 =
  ## nimfm_reg of include/nimfm_cuda.h
  when R is L1: 1
  elif R is SquaredL12: (if reg.transpose: 2 else: 3)     # `transpose` is private (squaredl12.nim:10): add a getter
  elif R is L21: 4
  else: {.error: "regulariser not supported by MBPSGD on the device path".}

proc fit*[L, R](self: MBPSGD[L, R], X: RowDataset, y: seq[float64],
                sfm: FactorizationMachine,
                callback: (MBPSGD[L, R], FactorizationMachine)->void = nil) =
  ## Fits the sparse factorization machine on X and y by accelerated pgd.
  sfm.init(X)
  var y = sfm.checkTarget(y)
  let
    nSamples = X.nSamples
    nFeatures = X.nFeatures
    nComponents = sfm.P.shape[1]
    nOrders = sfm.P.shape[0]
    degree = sfm.degree
    nAugments = sfm.nAugments
  var
    indices = toSeq(0..<nSamples)
    isConverged = false
  if not sfm.warmstart:
    self.it = 1
  var miniBatchSize = self.miniBatchSize
  if miniBatchSize <= 0:
    miniBatchSize = (nFeatures * nSamples) div X.nnz
    miniBatchSize = max(miniBatchSize, 1)
  var maxIterInner = self.maxIterInner
  if maxIterInner <= 0:
    maxIterInner = (nSamples-1) div miniBatchSize + 1
    maxIterInner = max(maxIterInner, 1)
  var ii = 0
  let doShuffle = X.nCached == X.nSamples and self.shuffle
  if doShuffle: shuffle(indices)
  self.reg.initSGD(degree, nFeatures+nAugments, nComponents)   # SquaredL12 raises for degree != 2 (squaredl12.nim:103-106)

  let ds = device(X)
  setTargets(ds, y)
  let h = toDevice(sfm, nFeatures)
  var cfg = MbpsgdCfg(loss: lossKind(self.loss), huberThreshold: lossThreshold(self.loss), eta0: self.eta0,
                      alpha0: self.alpha0, alpha: self.alpha, beta: self.beta, gamma: self.gamma,
                      reg: regCode(self.reg), scheduling: schedCode(self.scheduling), power: self.power,
                      miniBatchSize: miniBatchSize, maxIterInner: maxIterInner)
  if self.verbose > 0:
    echo("Minibatch size: ", miniBatchSize)
    echo("Number of inner iteration: ", maxIterInner)
    echoHeader(self.maxIter, viol=false)
  var oldLossVal = Inf
  var sample = newSeq[int](if doShuffle: miniBatchSize * maxIterInner else: 0)
  try:
    for it in 0..<self.maxIter:
      # the rows epoch() will visit: the cursor `ii` over `indices`, reshuffled at every wrap (:102-111)
      if doShuffle:
        var filled = 0
        while filled < sample.len:
          let take = min(sample.len - filled, nSamples - ii)
          for q in 0..<take: sample[filled+q] = indices[ii+q]
          filled += take
          ii += take
          if ii >= nSamples:
            ii = 0
            shuffle(indices)
      var
        itc = int64(self.it)
        iic = int64(ii)
        runningLoss: cdouble
      check nimfm_fm_mbpsgd_epoch(ctx(), h, ds, addr cfg, miniBatchSize, addr itc, addr iic,
                                  (if doShuffle: cast[ptr int64](p(sample)) else: nil), addr runningLoss)
      self.it = int(itc)
      if not doShuffle: ii = int(iic)

      if not callback.isNil:
        fromDevice(sfm, h)                    # pgd.finalize (pgd.nim:45-51): solver layout -> model layout
        callback(self, sfm)
      if runningLoss.classify == fcNan:
        echo("Loss is NaN. Use smaller learning rate.")
        break
      if self.verbose > 0:
        fromDevice(sfm, h)
        var regVal = regularization(sfm.P, sfm.w, sfm.intercept, self.alpha0, self.alpha, self.beta)
        for order in 0..<nOrders:
          regVal += self.gamma * self.reg.eval(sfm.P[order].T, sfm.degree-order)
        echoInfo(it+1, self.maxIter, -1, runningLoss, regVal)
      if abs(oldLossVal - runningLoss) < self.tol:
        if self.verbose > 0: echo("Converged at epoch ", it+1, ".")
        isConverged = true
        break
      oldLossVal = runningLoss
    if not isConverged and self.verbose > 0:
      echo("Objective did not converge. Increase maxIter.")
    fromDevice(sfm, h)                        # finalize
  finally:
    discard nimfm_fm_free(ctx(), h)

## decision_function.nim -- batched prediction on the device.  `include` the FM proc at the end of
## model/factorization_machine.nim in place of `decisionFunction` (factorization_machine.nim:100-122) and the FFM
## proc at the end of model/field_aware_factorization_machine.nim in place of :52-76.  predict / predictProba /
## score (fm_base.nim:18-47) call decisionFunction and need no change.

proc decisionFunction*[Dataset](self: FactorizationMachine, X: Dataset): seq[float64] =
  ## Returns the model outputs as seq[float64].
  self.checkInitialized()
  if X.nFeatures + self.nAugments != self.P.shape[2]:
    raise newException(ValueError, "Invalid nFeatures.")
  result = newSeq[float64](X.nSamples)
  let h = toDevice(self, X.nFeatures)
  try:
    # a CSR goes through the row kernel; a CSC through its stable transpose (kernels.nim:4-11,22-43 visit
    # the columns in ascending order, which is the order a sorted CSR row is visited in)
    check nimfm_fm_decision_function(ctx(), h, device(X), cast[ptr cdouble](p(result)))
  finally:
    discard nimfm_fm_free(ctx(), h)

proc decisionFunction*(self: FieldAwareFactorizationMachine, X: RowFieldDataset): seq[float64] =
  ## Returns the model outputs as seq[float64].
  self.checkInitialized()
  result = newSeq[float64](X.nSamples)
  let h = toDevice(self, X.nFeatures, X.nFields)
  try:
    check nimfm_ffm_decision_function(ctx(), h, device(X), cast[ptr cdouble](p(result)))
  finally:
    discard nimfm_ffm_free(ctx(), h)

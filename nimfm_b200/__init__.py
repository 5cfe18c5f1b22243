"""nimfm_b200 -- B200-native (sm_100a CUDA) implementation of nimfm's ANOVA-kernel hot path behind
nimfm's own API names.  Everything numeric runs in the in-tree libnimfm_cuda.so (C ABI:
include/nimfm_cuda.h); this package is the thin host mirror a Nim program would get from
nim/nimfm_cuda.nim.  There is no CPU fallback."""
from . import _lib
from ._lib import NimfmCudaError
from .dataset import (CSCDataset, CSRDataset, CSRFieldDataset, StreamCSRDataset, StreamCSRFieldDataset, convertFFMFile,
                      convertSVMLightFile, dumpFFMFile, newStreamCSRFieldDataset, transposeFieldFile,
                      dumpSVMLightFile, loadFFMFile, loadStreamLabel, loadSVMLightFile,
                      loadUserItemRatingFile, newStreamCSCDataset, newStreamCSRDataset, transposeFile, newCSCDataset, newCSRDataset,
                      newCSRFieldDataset, shuffle, toCSCDataset, toCSRDataset, vstack)
from .loss import (Huber, Logistic, Squared, SquaredHinge, newHuber, newLogistic, newSquared,
                   newSquaredHinge)
from .model import (FactorizationMachine, FieldAwareFactorizationMachine, NotFittedError, augment,
                    classification, explicit, newFactorizationMachine,
                    newFieldAwareFactorizationMachine, none, regression)
from .optimizers import (CD, L1, L21, MBPSGD, PCD, PSGD, SGD, AdaGrad, SquaredL12, constant, invscaling, newAdaGrad, newCD,
                         newL1, newL21, newMBPSGD, newPCD, newPSGD, newSGD, newSquaredL12, optimal, pegasos, regularization)

__all__ = [n for n in dir() if not n.startswith("_")]

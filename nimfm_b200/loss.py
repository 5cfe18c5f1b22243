"""Loss markers mirroring nimfm's loss.nim:1-102 (the arithmetic itself runs inside the kernels)."""
from . import _lib


class _Loss:
    kind = _lib.LOSS_SQUARED
    mu = 1.0
    threshold = 1.0


class Squared(_Loss):          # loss.nim:15-27
    kind, mu = _lib.LOSS_SQUARED, 1.0


class SquaredHinge(_Loss):     # loss.nim:30-48
    kind, mu = _lib.LOSS_SQUARED_HINGE, 2.0


class Logistic(_Loss):         # loss.nim:51-78
    kind, mu = _lib.LOSS_LOGISTIC, 0.25


class Huber(_Loss):            # loss.nim:81-102
    kind, mu = _lib.LOSS_HUBER, 1.0

    def __init__(self, threshold=1.0):
        self.threshold = float(threshold)


def newSquared():
    return Squared()


def newSquaredHinge():
    return SquaredHinge()


def newLogistic():
    return Logistic()


def newHuber(threshold=1.0):
    return Huber(threshold)

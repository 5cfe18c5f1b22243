"""`python -m nimfm_b200 train|test ...` -- the end-user command of the reference (src/nimfm.nim:72-135:
`nimfm train` / `nimfm test`, cligen-generated options) on top of the device path: svmlight files in,
CD / SGD / AdaGrad fit, RMSE / accuracy on a test file, optional prediction file and model dump / load.
Option names, defaults and printed lines follow the reference."""
import argparse
import sys

import numpy as np

from . import (Huber, Logistic, Squared, SquaredHinge, FactorizationMachine, classification, loadSVMLightFile,
               newAdaGrad, newCD, newFactorizationMachine, newSGD, regression)


def _bool(s):
    return str(s).lower() in ("1", "true", "yes", "on", "y")


def echoDataInfo(X):                               # src/nimfm.nim:8-13
    print("   Number of samples  : ", X.nSamples)
    print("   Number of features : ", X.nFeatures)
    print("   Number of non-zeros: ", X.nnz)
    print("   Maximum value      : ", float(np.max(X.data)) if X.nnz else 0.0)
    print("   Minimum value      : ", float(np.min(X.data)) if X.nnz else 0.0)


def make_loss(name, threshold=0.1):                # src/nimfm.nim:92-106
    table = {"squared": Squared, "squared_hinge": SquaredHinge, "logistic": Logistic}
    if name == "huber":
        return Huber(threshold)
    if name not in table:
        raise ValueError(f"loss {name} is not supported")
    return table[name]()


def evaluate(fm, task, test, predict, nFeatures, verbose):   # eval, src/nimfm.nim:16-34
    if verbose > 0:
        print("Load test data")
    X, y = loadSVMLightFile(test, nFeatures)
    if verbose > 0:
        echoDataInfo(X)
    yPred = fm.decisionFunction(X)
    if task == regression:
        print("Test RMSE: ", float(np.sqrt(np.mean((y - yPred) ** 2))))
    else:
        print("Test Accuracy: ", float(np.mean(np.sign(y) == np.sign(yPred))))
    if predict:
        with open(predict, "w") as f:
            for val in yPred:
                f.write(repr(float(val)) + "\n")


def train(a):                                      # train / trainInner, src/nimfm.nim:37-110
    if a.load:
        fm = FactorizationMachine.load(a.load, True)
    else:
        fm = newFactorizationMachine(task=a.task, degree=a.degree, nComponents=a.nComponents, fitLower=a.fitLower,
                                     fitIntercept=a.fitIntercept, fitLinear=a.fitLinear, warmStart=False,
                                     randomState=a.randomState, scale=a.scale)
    loss = make_loss(a.loss, a.threshold)
    if a.solver in ("cd", "als"):
        X, y = loadSVMLightFile(a.train, a.nFeatures, kind="csc")
        if a.verbose > 0:
            echoDataInfo(X)
        opt = newCD(maxIter=a.maxIter, alpha0=a.alpha0, alpha=a.alpha, beta=a.beta, loss=loss, verbose=a.verbose,
                    tol=a.tol)
    elif a.solver == "sgd":
        X, y = loadSVMLightFile(a.train, a.nFeatures)
        if a.verbose > 0:
            echoDataInfo(X)
        opt = newSGD(a.maxIter, a.eta0, a.alpha0, a.alpha, a.beta, loss, a.scheduling, a.power, a.verbose, a.tol)
    elif a.solver == "adagrad":
        X, y = loadSVMLightFile(a.train, a.nFeatures)
        if a.verbose > 0:
            echoDataInfo(X)
        opt = newAdaGrad(a.maxIter, a.eta0, a.alpha0, a.alpha, a.beta, loss, verbose=a.verbose, tol=a.tol)
    else:
        raise ValueError("Solver " + a.solver + " is not supported")
    opt.fit(X, y, fm)
    if a.test:
        evaluate(fm, a.task, a.test, a.predict, a.nFeatures, a.verbose)
    if a.dump:
        fm.dump(a.dump)


def test(a):                                       # test / testInner, src/nimfm.nim:113-135
    make_loss(a.loss)                              # validates the name as the reference's case statement does
    fm = FactorizationMachine.load(a.load, False)
    evaluate(fm, a.task, a.test, a.predict, a.nFeatures, a.verbose)
    if a.dump:
        fm.dump(a.dump)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="nimfm", description="factorization machines on the B200 path")
    sub = ap.add_subparsers(dest="cmd", required=True)
    tr = sub.add_parser("train", help="training a factorization machine")
    tr.add_argument("--task", required=True, choices=[regression, classification, "regression", "classification"])
    tr.add_argument("--train", required=True)
    tr.add_argument("--test", default="")
    tr.add_argument("--degree", type=int, default=2)
    tr.add_argument("--nComponents", type=int, default=30)
    tr.add_argument("--alpha0", type=float, default=1e-7)
    tr.add_argument("--alpha", type=float, default=1e-5)
    tr.add_argument("--beta", type=float, default=1e-3)
    tr.add_argument("--loss", default="squared")
    tr.add_argument("--fitLower", default="explicit", choices=["explicit", "augment", "none"])
    tr.add_argument("--fitLinear", type=_bool, default=True)
    tr.add_argument("--fitIntercept", type=_bool, default=True)
    tr.add_argument("--scale", type=float, default=0.1)
    tr.add_argument("--randomState", type=int, default=1)
    tr.add_argument("--solver", default="cd")
    tr.add_argument("--maxIter", type=int, default=100)
    tr.add_argument("--tol", type=float, default=1e-5)
    tr.add_argument("--eta0", type=float, default=0.1)
    tr.add_argument("--scheduling", default="optimal", choices=["constant", "optimal", "invscaling", "pegasos"])
    tr.add_argument("--power", type=float, default=1.0)
    tr.add_argument("--threshold", type=float, default=0.1)
    tr.add_argument("--dump", default="")
    tr.add_argument("--load", default="")
    tr.add_argument("--predict", default="")
    tr.add_argument("--nFeatures", type=int, default=-1)
    tr.add_argument("--verbose", type=int, default=1)
    tr.set_defaults(fn=train)
    te = sub.add_parser("test", help="test a factorization machine")
    te.add_argument("--task", required=True, choices=[regression, classification, "regression", "classification"])
    te.add_argument("--test", required=True)
    te.add_argument("--load", required=True)
    te.add_argument("--dump", default="")
    te.add_argument("--loss", default="squared")
    te.add_argument("--predict", default="")
    te.add_argument("--nFeatures", type=int, default=-1)
    te.add_argument("--verbose", type=int, default=1)
    te.set_defaults(fn=test)
    a = ap.parse_args(argv)
    a.task = {"regression": regression, "classification": classification}.get(a.task, a.task)
    a.fn(a)


if __name__ == "__main__":
    main(sys.argv[1:])

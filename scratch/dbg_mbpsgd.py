import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import nimfm_b200 as nf
from oracle import oracle as orc
from oracle.oracle import CSR
from helpers import make_dense, make_fm_params
n, d, k, degree = 80, 8, 4, 2
X = make_dense(n, d, 13, density=0.6, positive=False)
y = np.sign(np.random.default_rng(2).standard_normal(n))
csr = CSR.from_dense(X)
P, w, nA = make_fm_params(d, degree, k, "explicit", True, seed=11, scale=0.1)
kw = dict(eta0=0.1, alpha0=1e-6, alpha=1e-3, beta=1e-4, gamma=0.0)
for inner in (1, 2, 3):
    ref = orc.mbpsgd_fit(csr, y, P, w, 0.0, degree, "logistic", True, True, max_iter=1, mini_batch_size=7, max_iter_inner=inner, **kw)
    fm = nf.newFactorizationMachine(nf.classification, degree=degree, nComponents=k, warmStart=True)
    fm.P, fm.w, fm.intercept, fm.isInitialized = P.copy(), w.copy(), 0.0, True
    opt = nf.newMBPSGD(maxIter=1, loss=nf.Logistic(), miniBatchSize=7, maxIterInner=inner, verbose=0, tol=0.0, shuffle=False, **kw)
    ds = nf.newCSRDataset(csr.data, csr.indices, csr.indptr, n, d)
    opt.fit(ds, y, fm)
    print(inner, 'loss', opt.history[0], ref['epoch_loss'][0], 'dP', np.abs(fm.P-ref['P']).max(), 'dw', np.abs(fm.w-ref['w']).max(), 'db', fm.intercept-ref['intercept'], fm.intercept)

/*
 * oracle/ref_psgd.c -- CPU restatement of nimfm's proximal SGD (optimizer/psgd.nim:76-215) with the
 * per-regulariser SGD protocols it drives:
 *   L1          regularizer/l1.nim:84-136     (lazy scaling + accumulated soft-threshold)
 *   L21         regularizer/l21.nim:36-112    (lazy scaling + accumulated group soft-threshold)
 *   SquaredL12  regularizer/squaredl12.nim:199-230 (dense step + full prox every sample)
 * TEST INFRASTRUCTURE ONLY (see ref_cpu.c header).  The lazy protocols are restated LITERALLY (they
 * are not algebraically identical to an eager prox per step: e.g. l21.nim:86 accumulates eta*scaling
 * with the pre-update scaling); tests/test_oracle.py pins the result against the reference's naive
 * PSGDSlow (tests/optimizer/psgd_slow.nim) at the tolerance of tests/test_psgd_*.nim.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

double ref_loss(int kind, double thr, double y, double p);
double ref_dloss(int kind, double thr, double y, double p);
double ref_predict_with_grad(i64 d, const double *data, const i64 *indices, const i64 *indptr,
                             i64 i, int degree, int k, int nOrders, int nAug, const double *P,
                             const double *w, double intercept, double *A, double *dA);
void ref_prox_l21_row(double *pj, i64 k, double lam);
void ref_prox_matrix(double *P, i64 dd, int k, double lam, int reg_kind);

static double softthr(double x, double a) {
  double s = (x > 0) - (x < 0);
  double m = fabs(x) - a;
  return s * (m > 0.0 ? m : 0.0);
}

static double get_eta(int sched, double eta0, double power, double reg, i64 it) {   /* psgd uses sgd.getEta */
  switch (sched) {
    case 0: return eta0;
    case 1: return eta0 / pow(1.0 + eta0 * reg * (double)it, power);
    case 2: return eta0 / pow((double)it, power);
    default: return 1.0 / (reg * (double)it);
  }
}

typedef struct {
  int kind;                 /* 1 = L1, 2 = SquaredL12 transpose=true, 3 = transpose=false, 4 = L21 */
  double scaling, threshold;
  double *scalings, *thresholds;   /* [dd] */
} SgdReg;

/* reg.lazyUpdate(P[order], beta, gamma, degree, X, i): row features incl. dummies */
static void reg_lazy_update(SgdReg *r, double *Po, int k, double gamma, i64 d, int nAug,
                            const i64 *indices, i64 b, i64 e) {
  if (r->kind == 2 || r->kind == 3) return;                                   /* squaredl12.nim:199-201 */
  for (i64 jj = b; jj < e + nAug; jj++) {
    i64 j = jj < e ? indices[jj] : d + (jj - e);
    double *pj = Po + j * k;
    if (r->kind == 1) {                                                       /* l1.nim:84-90 */
      for (int s = 0; s < k; s++) {
        pj[s] *= r->scaling / r->scalings[j];
        pj[s] = softthr(pj[s], gamma * r->scaling * (r->threshold - r->thresholds[j]));
      }
    } else {                                                                  /* l21.nim:60-65 */
      double threshold = (r->threshold - r->thresholds[j]) / r->scalings[j];
      ref_prox_l21_row(pj, k, threshold * gamma);
      for (int s = 0; s < k; s++) pj[s] *= r->scaling / r->scalings[j];
    }
  }
}

/* reg.step(P[order], dA[order], dL, beta, gamma, eta_P_scaled, degree-order, row indices) */
static void reg_step(SgdReg *r, double *Po, const double *dAo, i64 dd, int k, double dL, double beta,
                     double gamma, double etaS, i64 d, int nAug, const i64 *indices, i64 b, i64 e) {
  if (r->kind == 2 || r->kind == 3) {                                         /* squaredl12.nim:222-230: dense */
    for (i64 j = 0; j < dd; j++)
      for (int s = 0; s < k; s++) Po[j * k + s] -= etaS * (dL * dAo[j * k + s] + beta * Po[j * k + s]);
    ref_prox_matrix(Po, dd, k, gamma * etaS, r->kind);
    return;
  }
  for (i64 jj = b; jj < e + nAug; jj++) {
    i64 j = jj < e ? indices[jj] : d + (jj - e);
    double *pj = Po + j * k;
    const double *dj = dAo + j * k;
    if (r->kind == 1) {                                                       /* l1.nim:127-136 */
      for (int s = 0; s < k; s++) {
        double update = etaS * (dL * dj[s] + beta * pj[s]);
        pj[s] = softthr(pj[s] - update, gamma * etaS);
      }
    } else {                                                                  /* l21.nim:102-112 */
      for (int s = 0; s < k; s++) {
        double update = etaS * (dL * dj[s] + beta * pj[s]);
        pj[s] -= update;
      }
      ref_prox_l21_row(pj, k, etaS * gamma);
    }
  }
}

/* reg.updateCacheSGD(eta_P, beta, gamma, degree, X, i) */
static void reg_update_cache(SgdReg *r, double eta, double beta, i64 d, int nAug, const i64 *indices,
                             i64 b, i64 e) {
  if (r->kind == 1) {                                                         /* l1.nim:106-113 */
    double etaS = eta / (1 + eta * beta);
    r->threshold += etaS / r->scaling;
    r->scaling *= (1 - etaS * beta);
  } else if (r->kind == 4) {                                                  /* l21.nim:84-90 */
    r->threshold += eta * r->scaling;
    r->scaling /= (1 + eta * beta);
  } else {
    return;
  }
  for (i64 jj = b; jj < e + nAug; jj++) {
    i64 j = jj < e ? indices[jj] : d + (jj - e);
    r->scalings[j] = r->scaling;
    r->thresholds[j] = r->threshold;
  }
}

/* reg.resetCacheSGD(P, gamma, degree): literal (l1.nim:116-124 multiplies by self.threshold as written) */
static void reg_reset_cache(SgdReg *r, double *P, int nOrders, i64 dd, int k, double gamma) {
  if (r->kind == 1 && r->scaling < 1e-8) {
    for (int o = 0; o < nOrders; o++)
      for (i64 j = 0; j < dd; j++)
        for (int s = 0; s < k; s++) {
          double *p = &P[((i64)o * dd + j) * k + s];
          *p /= r->scalings[j];
          *p = softthr(*p, gamma * r->threshold - r->thresholds[j]);
          *p *= r->threshold;
        }
  } else if (r->kind == 4 && r->scaling < 1e-8) {                             /* l21.nim:93-104 */
    for (int o = 0; o < nOrders; o++)
      for (i64 j = 0; j < dd; j++) {
        double *pj = P + ((i64)o * dd + j) * k;
        double threshold = (r->threshold - r->thresholds[j]) / r->scalings[j];
        ref_prox_l21_row(pj, k, threshold * gamma);
        for (int s = 0; s < k; s++) pj[s] *= r->scaling / r->scalings[j];
      }
  } else {
    return;
  }
  r->threshold = 0.0;
  r->scaling = 1.0;
  for (i64 j = 0; j < dd; j++) { r->thresholds[j] = 0.0; r->scalings[j] = 1.0; }
}

/* reg.lazyUpdateFinal(P, beta, gamma, degree) */
static void reg_lazy_final(SgdReg *r, double *P, int nOrders, i64 dd, int k, double gamma) {
  if (r->kind == 1) {                                                         /* l1.nim:93-103 */
    for (int o = 0; o < nOrders; o++)
      for (i64 j = 0; j < dd; j++)
        for (int s = 0; s < k; s++) {
          double *p = &P[((i64)o * dd + j) * k + s];
          *p *= r->scaling / r->scalings[j];
          *p = softthr(*p, gamma * r->scaling * (r->threshold - r->thresholds[j]));
        }
    for (i64 j = 0; j < dd; j++) { r->scalings[j] = 1.0; r->thresholds[j] = 0.0; }
    r->scaling = 1.0;
    r->threshold = 0.0;
  } else if (r->kind == 4) {                                                  /* l21.nim:68-75 (no reset) */
    for (int o = 0; o < nOrders; o++)
      for (i64 j = 0; j < dd; j++) {
        double *pj = P + ((i64)o * dd + j) * k;
        double threshold = (r->threshold - r->thresholds[j]) / r->scalings[j];
        ref_prox_l21_row(pj, k, threshold * gamma);
        for (int s = 0; s < k; s++) pj[s] *= r->scaling / r->scalings[j];
      }
  }
}

/* PSGD.fit (psgd.nim:76-215), X CSR.  Pm / w / intercept in-out (MODEL layout); perms NULL (cyclic) or
 * [maxIter][n]; it_io = PSGD.it; epoch_loss[maxIter] = runningLoss per epoch.  Returns epochs run. */
int ref_psgd_fit(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                 const double *y, int degree, int k, int nOrders, int nAug, int fitLinear,
                 int fitIntercept, double *Pm, double *w, double *intercept_io, int loss_kind,
                 double thr, int maxIter, double eta0, double alpha0, double alpha, double beta,
                 double gamma, int reg_kind, int sched, double power, double tol, const i64 *perms,
                 i64 *it_io, double *epoch_loss) {
  i64 dd = d + nAug, nP = (i64)nOrders * dd * k;
  int astride = degree + 1;
  double *P = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *dA = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *A = (double *)calloc((size_t)k * astride, sizeof(double));
  double *scalings_w = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  SgdReg reg;
  reg.kind = reg_kind;
  reg.scaling = 1.0;
  reg.threshold = 0.0;
  reg.scalings = (double *)malloc(sizeof(double) * (size_t)(dd > 0 ? dd : 1));
  reg.thresholds = (double *)calloc((size_t)(dd > 0 ? dd : 1), sizeof(double));
  for (i64 j = 0; j < dd; j++) reg.scalings[j] = 1.0;                         /* initSGD */
  for (i64 j = 0; j < d; j++) scalings_w[j] = 1.0;
  double scaling_w = 1.0, intercept = *intercept_io, oldLoss = 0.0;          /* runningLossOld = 0.0 (:103) */
  i64 it = *it_io;
  for (int o = 0; o < nOrders; o++)                                           /* P[order] = sfm.P[order].T (:108-109) */
    for (int s = 0; s < k; s++)
      for (i64 j = 0; j < dd; j++) P[((i64)o * dd + j) * k + s] = Pm[((i64)o * k + s) * dd + j];
  int epochs = 0;
  for (int ep = 0; ep < maxIter; ep++) {
    double runningLoss = 0.0;
    for (i64 q = 0; q < n; q++) {
      i64 i = perms ? perms[(i64)ep * n + q] : q;
      i64 b = indptr[i], e = indptr[i + 1];
      for (i64 jj = b; jj < e; jj++) w[indices[jj]] *= scaling_w / scalings_w[indices[jj]];   /* :122-123 */
      for (int o = 0; o < nOrders; o++)
        reg_lazy_update(&reg, P + (i64)o * dd * k, k, gamma, d, nAug, indices, b, e);
      double yPred = ref_predict_with_grad(d, data, indices, indptr, i, degree, k, nOrders, nAug, P, w,
                                           intercept, A, dA);
      runningLoss += ref_loss(loss_kind, thr, y[i], yPred);
      double eta_w = get_eta(sched, eta0, power, alpha, it);
      double eta_P = get_eta(sched, eta0, power, beta, it);
      double etaS = eta_P / (1.0 + eta_P * beta);
      double dL = ref_dloss(loss_kind, thr, y[i], yPred);
      for (int o = 0; o < nOrders; o++)
        reg_step(&reg, P + (i64)o * dd * k, dA + (i64)o * dd * k, dd, k, dL, beta, gamma, etaS, d, nAug,
                 indices, b, e);
      reg_update_cache(&reg, eta_P, beta, d, nAug, indices, b, e);
      if (fitIntercept) {                                                     /* :146-148 */
        double eta0v = get_eta(sched, eta0, power, alpha0, it);
        double update = eta0v * (dL + alpha0 * intercept);
        intercept -= update / (1.0 + eta0v * alpha0);
      }
      if (fitLinear) {                                                        /* fitLinearSGD, fit_linear.nim:41-47 */
        double eta = eta_w / (1.0 + eta_w * alpha);
        for (i64 jj = b; jj < e; jj++) {
          i64 j = indices[jj];
          w[j] -= eta * (dL * data[jj] + alpha * w[j]);
        }
      }
      scaling_w /= (1.0 + eta_w * alpha);                                     /* :153-155 */
      for (i64 jj = b; jj < e; jj++) scalings_w[indices[jj]] = scaling_w;
      if (fitLinear && scaling_w < 1e-9) {                                    /* :158-162 */
        for (i64 j = 0; j < d; j++) { w[j] *= scaling_w; w[j] /= scalings_w[j]; scalings_w[j] = 1.0; }
        scaling_w = 1.0;
      }
      reg_reset_cache(&reg, P, nOrders, dd, k, gamma);
      for (int o = 0; o < nOrders; o++)                                       /* reset dA (:166-172) */
        for (i64 jj = b; jj < e + nAug; jj++) {
          i64 j = jj < e ? indices[jj] : d + (jj - e);
          for (int s = 0; s < k; s++) dA[((i64)o * dd + j) * k + s] = 0.0;
        }
      it++;
    }
    runningLoss /= (double)n;
    epoch_loss[ep] = runningLoss;
    epochs = ep + 1;
    if (isnan(runningLoss)) break;
    if (fabs(runningLoss - oldLoss) < tol) break;                             /* :194-197 */
    oldLoss = runningLoss;
  }
  if (fitLinear)                                                              /* finalize, :58-75 */
    for (i64 j = 0; j < d; j++) { w[j] *= scaling_w; w[j] /= scalings_w[j]; }
  reg_lazy_final(&reg, P, nOrders, dd, k, gamma);
  for (int o = 0; o < nOrders; o++)
    for (i64 j = 0; j < dd; j++)
      for (int s = 0; s < k; s++) Pm[((i64)o * k + s) * dd + j] = P[((i64)o * dd + j) * k + s];
  *intercept_io = intercept;
  *it_io = it;
  free(P); free(dA); free(A); free(scalings_w); free(reg.scalings); free(reg.thresholds);
  return epochs;
}

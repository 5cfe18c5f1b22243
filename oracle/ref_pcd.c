/*
 * oracle/ref_pcd.c -- CPU restatement of nimfm's proximal coordinate descent (optimizer/pcd.nim:38-200)
 * with the per-coordinate prox / cache protocol of the regularisers it accepts:
 *   L1          regularizer/l1.nim:22-24,57-59        softthreshold(psj - update, lam)
 *   SquaredL12  regularizer/squaredl12.nim:108-114,161-185 (transpose = true | false)
 * TEST INFRASTRUCTURE ONLY (see ref_cpu.c header).  Pinned the way the reference pins pcd.nim
 * (tests/test_pcd_l1.nim, tests/test_pcd_squaredl12.nim): against the naive dense solver PCDSlow
 * (tests/optimizer/pcd_slow.nim + tests/regularizer/{l1,squaredl12}_slow.nim) restated in
 * oracle/bruteforce.py, plus pcd(gamma = 0) == cd and monotone decrease of the regularised objective
 * (tests/test_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

double ref_loss(int kind, double thr, double y, double p);
double ref_dloss(int kind, double thr, double y, double p);
double ref_mu(int kind);
void ref_linear_csc(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                    const double *w, double *out);
void ref_anova_csc(i64 n, i64 d, int nAug, const double *data, const i64 *indices, const i64 *indptr,
                   const double *Ps, double *A, int astride, int degree);
double ref_regularization(const double *P, i64 nP, const double *w, i64 nw, double intercept,
                          double alpha0, double alpha, double beta);
double ref_reg_eval(const double *P, i64 dd, int k, int reg_kind);

static double softthr(double x, double a) {   /* regularizer/utils.nim:4-5 */
  double s = (x > 0) - (x < 0);
  double m = fabs(x) - a;
  return s * (m > 0.0 ? m : 0.0);
}

/* regulariser state for PCD: reg_kind 1 = L1, 2 = SquaredL12 transpose=true, 3 = transpose=false */
typedef struct {
  int kind;
  double *cache;   /* [1] (transpose) or [dd] */
  double *absp;    /* [dd] */
} PcdReg;

static void reg_cache_all(PcdReg *r, const double *Po, int k, i64 dd) {         /* computeCacheCDAll */
  if (r->kind == 3) {                                                            /* squaredl12.nim:163-170 */
    for (i64 j = 0; j < dd; j++) r->cache[j] = 0.0;
    for (int s = 0; s < k; s++)
      for (i64 j = 0; j < dd; j++) r->cache[j] += fabs(Po[(i64)s * dd + j]);
  }
}
static void reg_cache(PcdReg *r, const double *Ps, i64 dd) {                     /* computeCacheCD, :173-179 */
  if (r->kind == 2 || r->kind == 3) {
    for (i64 j = 0; j < dd; j++) r->absp[j] = fabs(Ps[j]);
    if (r->kind == 2) {
      double s = 0.0;
      for (i64 j = 0; j < dd; j++) s += r->absp[j];
      r->cache[0] = s;
    }
  }
}
static double reg_prox(PcdReg *r, double psj, double update, double lam, i64 j) {
  if (r->kind == 1) return softthr(psj - update, lam);                           /* l1.nim:22-24 */
  i64 i = r->kind == 2 ? 0 : j;                                                  /* squaredl12.nim:108-114 */
  double dcache = r->cache[i] - r->absp[j];
  return softthr((psj - update) / (1 + 2 * lam), 2 * lam * dcache / (1 + 2 * lam));
}
static void reg_update_cache(PcdReg *r, const double *Ps, i64 j) {               /* updateCacheCD, :182-185 */
  if (r->kind == 2 || r->kind == 3) {
    i64 i = r->kind == 2 ? 0 : j;
    r->cache[i] -= r->absp[j];
    r->cache[i] += fabs(Ps[j]);
  }
}

/* X is CSC; P in the model layout [nOrders][k][d+nAug].  Per outer iteration: viol, mean loss and
 * (gamma*sum_o reg.eval(P[o].T) + regularization)/n (pcd.nim:178-187).  Returns iterations run. */
int ref_pcd_fit(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                const double *y, int degree, int k, int nOrders, int nAug, int fitLinear,
                int fitIntercept, double *P, double *w, double *intercept_io, int loss_kind,
                double thr, int maxIter, double alpha0_, double alpha_, double beta_, double gamma_,
                int reg_kind, double tol, double *viol_out, double *loss_out, double *reg_out,
                double *yPred_out) {
  i64 dd = d + nAug;
  double alpha0 = alpha0_ * (double)n, alpha = alpha_ * (double)n, beta = beta_ * (double)n,
         gamma = gamma_ * (double)n;                                             /* pcd.nim:118-121 */
  int astride = degree + 1;
  double mu = ref_mu(loss_kind);
  double *yPred = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double *A = (double *)calloc((size_t)(n > 0 ? n : 1) * astride, sizeof(double));
  double *dA = (double *)calloc((size_t)(degree > 0 ? degree : 1), sizeof(double));
  double *cache2 = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double *colNormSq = (double *)calloc((size_t)(d > 0 ? d : 1), sizeof(double));
  double *PT = (double *)calloc((size_t)(dd * k > 0 ? dd * k : 1), sizeof(double));
  PcdReg reg;
  reg.kind = reg_kind;
  reg.cache = (double *)calloc((size_t)(dd > 0 ? dd : 1), sizeof(double));
  reg.absp = (double *)calloc((size_t)(dd > 0 ? dd : 1), sizeof(double));
  double intercept = *intercept_io;
  for (i64 i = 0; i < n; i++) A[i * astride] = 1.0;                              /* :136 */
  if (fitLinear)                                                                 /* :137-139 */
    for (i64 j = 0; j < d; j++) {
      double s = 0.0;
      for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) s += data[ii] * data[ii];
      double nr = sqrt(s);
      colNormSq[j] = nr * nr;
    }
  ref_linear_csc(n, d, data, indices, indptr, w, yPred);                        /* :142-149 */
  for (i64 i = 0; i < n; i++) yPred[i] += intercept;
  for (int o = 0; o < nOrders; o++)
    for (int s = 0; s < k; s++) {
      ref_anova_csc(n, d, nAug, data, indices, indptr, P + ((i64)o * k + s) * dd, A, astride, degree - o);
      for (i64 i = 0; i < n; i++) yPred[i] += A[i * astride + degree - o];
    }
  int iters = 0;
  for (int it = 0; it < maxIter; it++) {
    double viol = 0.0;
    if (fitIntercept) {                                                          /* fit_linear.nim:28-38 */
      double r = alpha0 * intercept;
      for (i64 i = 0; i < n; i++) r += ref_dloss(loss_kind, thr, y[i], yPred[i]);
      r /= mu * (double)n + alpha0;
      intercept -= r;
      for (i64 i = 0; i < n; i++) yPred[i] -= r;
      viol += fabs(r);
    }
    if (fitLinear) {                                                             /* fit_linear.nim:5-25 */
      for (i64 j = 0; j < d; j++) {
        double update = alpha * w[j];
        for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++)
          update += ref_dloss(loss_kind, thr, y[indices[ii]], yPred[indices[ii]]) * data[ii];
        double inv = mu * colNormSq[j] + alpha;
        if (inv < 1e-12) continue;
        update /= inv;
        viol += fabs(update);
        w[j] -= update;
        for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) yPred[indices[ii]] -= update * data[ii];
      }
    }
    for (int o = 0; o < nOrders; o++) {
      int deg = degree - o;
      double *Po = P + (i64)o * k * dd;
      reg_cache_all(&reg, Po, k, dd);                                            /* pcd.nim:46 / :79 */
      for (int s = 0; s < k; s++) {
        double *Ps = Po + (i64)s * dd;
        if (deg > 2) {
          ref_anova_csc(n, d, nAug, data, indices, indptr, Ps, A, astride, deg); /* :49 */
        } else {                                                                 /* :82-85 */
          for (i64 i = 0; i < n; i++) cache2[i] = 0;
          for (i64 j = 0; j < dd; j++) {
            if (j < d) for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) cache2[indices[ii]] += data[ii] * Ps[j];
            else for (i64 i = 0; i < n; i++) cache2[i] += 1.0 * Ps[j];
          }
        }
        reg_cache(&reg, Ps, dd);                                                 /* :50 / :87 */
        for (i64 j = 0; j < dd; j++) {
          double psj = Ps[j];
          i64 cb = j < d ? indptr[j] : 0, ce = j < d ? indptr[j + 1] : n;
          double update = beta * psj, inv = 0.0;
          for (i64 ii = cb; ii < ce; ii++) {
            i64 i = j < d ? indices[ii] : ii;
            double val = j < d ? data[ii] : 1.0;
            double g;
            if (deg > 2) {                                                       /* cd.update, cd.nim:36-47 */
              dA[0] = val;
              for (int t = 1; t < deg; t++) dA[t] = val * (A[i * astride + t] - psj * dA[t - 1]);
              g = dA[deg - 1];
            } else {
              g = (cache2[i] - psj * val) * val;                                 /* pcd.nim:92 */
            }
            update += ref_dloss(loss_kind, thr, y[i], yPred[i]) * g;
            inv += g * g;
          }
          if (deg > 2) { inv *= mu; inv += beta; }                               /* cd.nim:46-47 */
          else inv = inv * mu + beta;                                            /* pcd.nim:95 */
          if (inv < 1e-12) continue;                                             /* pcd.nim:55 / :96 (both sweeps) */
          update /= inv;
          Ps[j] = reg_prox(&reg, psj, update, gamma / inv, j);                   /* :57 / :99 */
          update = psj - Ps[j];
          viol += fabs(update);
          for (i64 ii = cb; ii < ce; ii++) {                                     /* synchronize :62-68 / :103-105 */
            i64 i = j < d ? indices[ii] : ii;
            double val = j < d ? data[ii] : 1.0;
            if (deg > 2) {
              dA[0] = val;
              for (int t = 1; t < deg; t++) {
                dA[t] = val * (A[i * astride + t] - psj * dA[t - 1]);
                A[i * astride + t] -= update * dA[t - 1];
              }
              A[i * astride + deg] -= update * dA[deg - 1];
              yPred[i] -= update * dA[deg - 1];
            } else {
              yPred[i] -= update * (cache2[i] - psj * val) * val;
              cache2[i] -= update * val;
            }
          }
          reg_update_cache(&reg, Ps, j);                                         /* :69 / :106 */
        }
      }
    }
    double lossVal = 0.0, regVal = 0.0;                                          /* :174-187 */
    for (i64 i = 0; i < n; i++) lossVal += ref_loss(loss_kind, thr, y[i], yPred[i]);
    lossVal /= (double)n;
    for (int o = 0; o < nOrders; o++) {
      const double *Po = P + (i64)o * k * dd;
      for (i64 j = 0; j < dd; j++)
        for (int s = 0; s < k; s++) PT[j * k + s] = Po[(i64)s * dd + j];         /* sfm.P[order].T */
      regVal += gamma * ref_reg_eval(PT, dd, k, reg_kind);
    }
    regVal += ref_regularization(P, (i64)nOrders * k * dd, w, d, intercept, alpha0, alpha, beta);
    regVal /= (double)n;
    viol_out[it] = viol;
    loss_out[it] = lossVal;
    reg_out[it] = regVal;
    iters = it + 1;
    if (viol < tol) break;                                                       /* :192-195 */
  }
  if (yPred_out) memcpy(yPred_out, yPred, sizeof(double) * (size_t)n);
  *intercept_io = intercept;
  free(yPred); free(A); free(dA); free(cache2); free(colNormSq); free(PT); free(reg.cache); free(reg.absp);
  return iters;
}

/*
 * oracle/ref_hogwild.c -- the reference's ONLY parallel strategy, restated for the timed CPU
 * baseline: Hogwild (lock-free, racy by design) multi-thread AdaGrad epochs.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (see ref_cpu.c header).  Follows
 * optimizer/adagrad_multi.nim:15-36 (epochSub) and :66-101 (chunking + spawn/join) and
 * optimizer/adagrad_ffm_multi.nim:16-33; nThreads per optimizer/sgd_multi.nim:13-18.
 * Shared state (P, w, intercept, g_sum, g_norm, it) is touched without locks or atomics exactly as
 * in the reference, so results are nondeterministic; this file is only ever TIMED, never used as a
 * parity oracle.  Per-thread scratch: A [k][degree+1] and a dense P-shaped dA (adagrad_multi.nim:9-10).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

double ref_loss(int kind, double thr, double y, double p);
double ref_dloss(int kind, double thr, double y, double p);
double ref_predict_with_grad(i64 d, const double *data, const i64 *indices, const i64 *indptr,
                             i64 i, int degree, int k, int nOrders, int nAug, const double *P,
                             const double *w, double intercept, double *A, double *dA);
double ref_ffm_predict_with_grad(i64 d, int nFields, int k, const double *data, const i64 *indices,
                                 const i64 *indptr, const i64 *fields, i64 i, const double *P,
                                 const double *w, double intercept, double *dA);

typedef struct {
  /* dataset */
  i64 d;
  const double *data;
  const i64 *indices, *indptr, *fields;
  const double *y;
  /* model */
  int is_ffm, degree, k, nOrders, nAug, nFields, fitLinear, fitIntercept;
  double *P, *w;
  volatile double *intercept;
  /* optimizer (shared, racy) */
  int loss_kind;
  double thr, eta0, alpha0, alpha, beta;
  volatile i64 *it;
  double *gsP, *gnP, *gsw, *gnw;
  volatile double *gsb, *gnb;
  /* chunk */
  const i64 *order;
  i64 s, t;
  double loss, viol;
} hog_task;

static void *hog_epoch_sub(void *arg) {
  hog_task *T = (hog_task *)arg;
  i64 d = T->d, dd = d + T->nAug;
  int k = T->k;
  int nO = T->is_ffm ? T->nFields : T->nOrders;
  i64 nP = (i64)nO * dd * k;
  double *A = (double *)calloc((size_t)k * (T->degree + 1), sizeof(double));
  double *dA = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double eta0 = T->eta0;
  for (i64 q = T->s; q < T->t; q++) {
    i64 i = T->order ? T->order[q] : q;
    i64 rb = T->indptr[i], re = T->indptr[i + 1];
    i64 aug = T->is_ffm ? 0 : T->nAug;
    if (*T->it != 1) { /* AdaGrad.update, adagrad.nim:87-110 */
      double itf = (double)(*T->it - 1);
      double tmp = eta0 * itf * T->beta;
      for (int o = 0; o < nO; o++)
        for (i64 jj = rb; jj < re + aug; jj++) {
          i64 j = jj < re ? T->indices[jj] : d + (jj - re);
          for (int s = 0; s < k; s++) {
            i64 e = ((i64)o * dd + j) * k + s;
            double pjs = T->P[e];
            double denom = tmp + sqrt(T->gnP[e]);
            T->P[e] = -(eta0 * T->gsP[e]) / denom;
            T->viol += fabs(pjs - T->P[e]);
          }
        }
      if (T->fitIntercept) {
        double old = *T->intercept;
        double denom = sqrt(*T->gnb) + eta0 * itf * T->alpha0;
        *T->intercept = -eta0 * *T->gsb / denom;
        T->viol += fabs(old - *T->intercept);
      }
      if (T->fitLinear) {
        double denom = itf * eta0 * T->alpha;
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = T->indices[jj];
          double wj = T->w[j];
          T->w[j] = -eta0 * T->gsw[j] / (denom + sqrt(T->gnw[j]));
          T->viol += fabs(wj - T->w[j]);
        }
      }
    }
    double yPred = T->is_ffm
        ? ref_ffm_predict_with_grad(d, T->nFields, k, T->data, T->indices, T->indptr, T->fields, i,
                                    T->P, T->w, *T->intercept, dA)
        : ref_predict_with_grad(d, T->data, T->indices, T->indptr, i, T->degree, k, T->nOrders,
                                T->nAug, T->P, T->w, *T->intercept, A, dA);
    T->loss += ref_loss(T->loss_kind, T->thr, T->y[i], yPred);
    double dL = ref_dloss(T->loss_kind, T->thr, T->y[i], yPred); /* updateG, adagrad.nim:113-134 */
    for (int o = 0; o < nO; o++)
      for (i64 jj = rb; jj < re + aug; jj++) {
        i64 j = jj < re ? T->indices[jj] : d + (jj - re);
        for (int s = 0; s < k; s++) {
          i64 e = ((i64)o * dd + j) * k + s;
          double grad = dL * dA[e];
          T->gsP[e] += grad;
          T->gnP[e] += grad * grad;
        }
      }
    if (T->fitIntercept) { *T->gsb += dL; *T->gnb += dL * dL; }
    if (T->fitLinear)
      for (i64 jj = rb; jj < re; jj++) {
        i64 j = T->indices[jj];
        double val = T->data[jj];
        T->gsw[j] += dL * val;
        T->gnw[j] += (dL * val) * (dL * val);
      }
    *T->it += 1;
  }
  free(A);
  free(dA);
  return NULL;
}

/* One Hogwild AdaGrad epoch over rows order[0..nRows) (NULL = 0..nRows-1) with nThreads contiguous
 * chunks (adagrad_multi.nim:66-69,86-101).  P is the SOLVER layout: FM [nOrders][d+nAug][k],
 * FFM [nFields][d][k].  Returns the loss sum; *viol_out the violation sum. */
double ref_hogwild_adagrad_epoch(int is_ffm, i64 nRows, i64 d, const double *data, const i64 *indices,
                                 const i64 *indptr, const i64 *fields, const double *y, int degree,
                                 int k, int nOrders, int nAug, int nFields, int fitLinear,
                                 int fitIntercept, double *P, double *w, double *intercept,
                                 int loss_kind, double thr, double eta0, double alpha0, double alpha,
                                 double beta, i64 *it, double *gsP, double *gnP, double *gsw,
                                 double *gnw, double *gsb, double *gnb, const i64 *order,
                                 int nThreads, double *viol_out) {
  if (nThreads < 1) nThreads = 1;
  hog_task *tasks = (hog_task *)calloc((size_t)nThreads, sizeof(hog_task));
  pthread_t *th = (pthread_t *)calloc((size_t)nThreads, sizeof(pthread_t));
  i64 *borders = (i64 *)calloc((size_t)nThreads + 1, sizeof(i64));
  for (int t = 0; t < nThreads; t++) borders[t + 1] = borders[t] + nRows / nThreads;
  borders[nThreads] = nRows;
  for (int t = 0; t < nThreads; t++) {
    hog_task *T = &tasks[t];
    T->d = d; T->data = data; T->indices = indices; T->indptr = indptr; T->fields = fields; T->y = y;
    T->is_ffm = is_ffm; T->degree = degree; T->k = k; T->nOrders = nOrders; T->nAug = nAug;
    T->nFields = nFields; T->fitLinear = fitLinear; T->fitIntercept = fitIntercept;
    T->P = P; T->w = w; T->intercept = intercept;
    T->loss_kind = loss_kind; T->thr = thr; T->eta0 = eta0; T->alpha0 = alpha0; T->alpha = alpha;
    T->beta = beta; T->it = it; T->gsP = gsP; T->gnP = gnP; T->gsw = gsw; T->gnw = gnw;
    T->gsb = gsb; T->gnb = gnb; T->order = order; T->s = borders[t]; T->t = borders[t + 1];
    pthread_create(&th[t], NULL, hog_epoch_sub, T);
  }
  double loss = 0.0, viol = 0.0;
  for (int t = 0; t < nThreads; t++) {
    pthread_join(th[t], NULL);
    loss += tasks[t].loss;
    viol += tasks[t].viol;
  }
  if (viol_out) *viol_out = viol;
  free(tasks); free(th); free(borders);
  return loss;
}

/*
 * oracle/ref_cpu.c -- CPU restatement of nimfm's ANOVA-kernel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nimfm_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker (or as
 * the timed CPU baseline), never as the product path.
 *
 * Parity status: the reference (nimfm 0.3.0, pure Nim) cannot be compiled in
 * this image (no nim/nimble), and its tests hold NO golden vectors or
 * known-answer constants (they are differential tests against brute-force
 * "slow" re-implementations).  This file is therefore pinned the same way the
 * reference pins itself: against the combinatorial definitions restated in
 * oracle/bruteforce.py (from tests/kernels_slow.nim, tests/model/fm_slow.nim,
 * tests/model/ffm_slow.nim, tests/optimizer/{cd,sgd,adagrad,...}_slow.nim) on the reference's own
 * test shapes, plus committed fixtures in tests/golden/.  RNG-driven behaviour
 * (init, shuffle) is "parity unpinned": parameters and permutations are always
 * injected explicitly.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/nimfm/).  Loop order, accumulation order and the
 * degree-2 special cases are kept as in the reference; compile with
 * -ffp-contract=off so no FMA contraction changes the roundings.
 *
 * Layout conventions (all as in the reference):
 *   model P  ("component-major")  P[order][s][j],  j < d + nAug     (factorization_machine.nim:33-36)
 *   solver P ("feature-major")    P[order][j][s]                    (sgd.nim:92-96, 284,292)
 *   FFM P                         P[field][j][s]                    (field_aware_factorization_machine.nim:16-17)
 *   CSR/CSC: data f64[nnz], indices i64[nnz], indptr i64[rows+1 | cols+1]  (tensor/sparse.nim:4-31)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef int64_t i64;

#define LOSS_SQUARED 0
#define LOSS_SQUARED_HINGE 1
#define LOSS_LOGISTIC 2
#define LOSS_HUBER 3

#define SCHED_CONSTANT 0
#define SCHED_OPTIMAL 1
#define SCHED_INVSCALING 2
#define SCHED_PEGASOS 3

/* ------------------------------------------------------------------ */
/* losses: loss.nim:15-102                                            */
/* ------------------------------------------------------------------ */
double ref_loss(int kind, double thr, double y, double p) {
  switch (kind) {
  case LOSS_SQUARED: { double z = y - p; return 0.5 * (z * z); }            /* loss.nim:18 */
  case LOSS_SQUARED_HINGE: { double z = 1 - p * y; z = z > 0 ? z : 0; return z * z; } /* :33 */
  case LOSS_LOGISTIC: {                                                      /* :54-59 */
    double z = p * y;
    if (z > 0) return log(1 + exp(-z));
    return log(exp(z) + 1) - z;
  }
  default: {                                                                 /* :84-87 */
    double z = fabs(y - p);
    if (z < thr) return 0.5 * (z * z);
    return thr * (z - 0.5 * thr);
  }
  }
}

double ref_dloss(int kind, double thr, double y, double p) {
  switch (kind) {
  case LOSS_SQUARED: return p - y;                                           /* :21 */
  case LOSS_SQUARED_HINGE: { double z = 1 - p * y; return z > 0 ? -2 * y * z : 0.0; } /* :36-39 */
  case LOSS_LOGISTIC: {                                                      /* :62-67 */
    double z = p * y;
    if (z > 0) return -y * exp(-z) / (1 + exp(-z));
    return -y / (exp(z) + 1);
  }
  default: {                                                                 /* :90-93 (sign quirk kept) */
    double z = fabs(y - p);
    if (z < thr) return y - p;
    return thr;
  }
  }
}

double ref_mu(int kind) {
  switch (kind) {
  case LOSS_SQUARED: return 1.0;        /* :27 */
  case LOSS_SQUARED_HINGE: return 2.0;  /* :48 */
  case LOSS_LOGISTIC: return 0.25;      /* :78 */
  default: return 1.0;                  /* :102 */
  }
}

/* ------------------------------------------------------------------ */
/* integer bookkeeping: tensor/sparse.nim                             */
/* ------------------------------------------------------------------ */
/* toCSCMatrix, sparse.nim:510-527: counting sort by column, stable in row order */
void ref_csr_to_csc(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                    double *odata, i64 *oindices, i64 *oindptr) {
  for (i64 j = 0; j <= d; j++) oindptr[j] = 0;
  for (i64 i = 0; i < n; i++)
    for (i64 jj = indptr[i]; jj < indptr[i + 1]; jj++) oindptr[indices[jj] + 1] += 1;
  for (i64 j = 0; j < d; j++) oindptr[j + 1] += oindptr[j];
  i64 *offsets = (i64 *)calloc((size_t)(d > 0 ? d : 1), sizeof(i64));
  for (i64 i = 0; i < n; i++)
    for (i64 jj = indptr[i]; jj < indptr[i + 1]; jj++) {
      i64 j = indices[jj];
      odata[oindptr[j] + offsets[j]] = data[jj];
      oindices[oindptr[j] + offsets[j]] = i;
      offsets[j] += 1;
    }
  free(offsets);
}

/* toCSRMatrix, sparse.nim:490-507 (the same counting sort on the other axis) */
void ref_csc_to_csr(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                    double *odata, i64 *oindices, i64 *oindptr) {
  ref_csr_to_csc(d, n, data, indices, indptr, odata, oindices, oindptr);
}

/* X[indicesRow], sparse.nim:263-285.  Call with odata==NULL to get only oindptr (sizes). */
void ref_csr_take_rows(const double *data, const i64 *indices, const i64 *indptr,
                       const i64 *rows, i64 nrows, double *odata, i64 *oindices, i64 *oindptr) {
  i64 nnz = 0;
  oindptr[0] = 0;
  for (i64 ii = 0; ii < nrows; ii++) {
    i64 i = rows[ii];
    nnz += indptr[i + 1] - indptr[i];
    oindptr[ii + 1] = nnz;
  }
  if (!odata) return;
  i64 count = 0;
  for (i64 ii = 0; ii < nrows; ii++) {
    i64 i = rows[ii];
    for (i64 jj = indptr[i]; jj < indptr[i + 1]; jj++) {
      oindices[count] = indices[jj];
      odata[count] = data[jj];
      count++;
    }
  }
}

/* CSC row-range slice X[a..b] (b inclusive), sparse.nim:300-325. Two-phase like take_rows. */
void ref_csc_slice_rows(i64 d, const double *data, const i64 *indices, const i64 *indptr,
                        i64 a, i64 b, double *odata, i64 *oindices, i64 *oindptr) {
  for (i64 j = 0; j <= d; j++) oindptr[j] = 0;
  for (i64 j = 0; j < d; j++)
    for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++)
      if (indices[ii] >= a && indices[ii] <= b) oindptr[j + 1] += 1;
  for (i64 j = 0; j < d; j++) oindptr[j + 1] += oindptr[j];
  if (!odata) return;
  for (i64 j = 0; j < d; j++) {
    i64 count = 0;
    for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++)
      if (indices[ii] >= a && indices[ii] <= b) {
        oindices[oindptr[j] + count] = indices[ii] - a;
        odata[oindptr[j] + count] = data[ii];
        count++;
      }
  }
}

/* ------------------------------------------------------------------ */
/* kernels.nim                                                        */
/* ------------------------------------------------------------------ */
/* linear, RowDataset: kernels.nim:14-19 (dummy features are NOT present when linear runs) */
void ref_linear_csr(i64 n, const double *data, const i64 *indices, const i64 *indptr,
                    const double *w, double *out) {
  for (i64 i = 0; i < n; i++) {
    out[i] = 0.0;
    for (i64 jj = indptr[i]; jj < indptr[i + 1]; jj++) out[i] += w[indices[jj]] * data[jj];
  }
}

/* linear, ColDataset: kernels.nim:4-11 */
void ref_linear_csc(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                    const double *w, double *out) {
  for (i64 i = 0; i < n; i++) out[i] = 0.0;
  for (i64 j = 0; j < d; j++)
    for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) out[indices[ii]] += data[ii] * w[j];
}

/* anova, RowDataset: kernels.nim:46-64.  Ps = P[order][s][:] (length d+nAug).
 * A is [n][astride]; dummy features (d+a, 1.0) follow each row (dataset.nim:182-189). */
void ref_anova_csr(i64 n, i64 d, int nAug, const double *data, const i64 *indices,
                   const i64 *indptr, const double *Ps, double *A, int astride, int degree) {
  for (i64 i = 0; i < n; i++) {
    for (int t = 1; t < degree + 1; t++) A[i * astride + t] = 0.0;
    A[i * astride] = 1.0;
  }
  if (degree != 2) {
    for (i64 i = 0; i < n; i++) {
      double *a = A + i * astride;
      for (i64 jj = indptr[i]; jj < indptr[i + 1]; jj++) {
        double p = Ps[indices[jj]], val = data[jj];
        for (int t = 0; t < degree; t++) a[degree - t] += a[degree - t - 1] * p * val;
      }
      for (int aug = 0; aug < nAug; aug++) {
        double p = Ps[d + aug], val = 1.0;
        for (int t = 0; t < degree; t++) a[degree - t] += a[degree - t - 1] * p * val;
      }
    }
  } else {
    for (i64 i = 0; i < n; i++) {
      double *a = A + i * astride;
      for (i64 jj = indptr[i]; jj < indptr[i + 1]; jj++) {
        double pv = Ps[indices[jj]] * data[jj];
        a[1] += pv;
        a[2] += pv * pv;
      }
      for (int aug = 0; aug < nAug; aug++) {
        double pv = Ps[d + aug] * 1.0;
        a[1] += pv;
        a[2] += pv * pv;
      }
      a[2] = (a[1] * a[1] - a[2]) / 2.0;
    }
  }
}

/* anova, ColDataset: kernels.nim:22-43.  Dummy column d+a touches every row (dataset.nim:245-252). */
void ref_anova_csc(i64 n, i64 d, int nAug, const double *data, const i64 *indices,
                   const i64 *indptr, const double *Ps, double *A, int astride, int degree) {
  for (i64 i = 0; i < n; i++) {
    for (int t = 1; t < degree + 1; t++) A[i * astride + t] = 0.0;
    A[i * astride] = 1.0;
  }
  if (degree != 2) {
    for (i64 j = 0; j < d + nAug; j++) {
      double p = Ps[j];
      if (j < d) {
        for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) {
          double *a = A + indices[ii] * astride, val = data[ii];
          for (int t = 0; t < degree; t++) a[degree - t] += a[degree - t - 1] * p * val;
        }
      } else {
        for (i64 i = 0; i < n; i++) {
          double *a = A + i * astride, val = 1.0;
          for (int t = 0; t < degree; t++) a[degree - t] += a[degree - t - 1] * p * val;
        }
      }
    }
  } else {
    for (i64 j = 0; j < d + nAug; j++) {
      double p = Ps[j];
      if (j < d) {
        for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) {
          double *a = A + indices[ii] * astride, pv = p * data[ii];
          a[1] += pv;
          a[2] += pv * pv;
        }
      } else {
        for (i64 i = 0; i < n; i++) {
          double *a = A + i * astride, pv = p * 1.0;
          a[1] += pv;
          a[2] += pv * pv;
        }
      }
    }
    for (i64 i = 0; i < n; i++) {
      double *a = A + i * astride;
      a[2] = (a[1] * a[1] - a[2]) / 2.0;
    }
  }
}

/* FactorizationMachine.decisionFunction: model/factorization_machine.nim:100-122.
 * P is the model layout [nOrders][k][d+nAug]; is_csc selects the ColDataset path. */
void ref_fm_decision_function(int is_csc, i64 n, i64 d, const double *data, const i64 *indices,
                              const i64 *indptr, int degree, int k, int nOrders, int nAug,
                              const double *P, const double *w, double intercept,
                              const double *lams, double *out) {
  int astride = degree + 1;
  double *A = (double *)calloc((size_t)(n > 0 ? n : 1) * astride, sizeof(double));
  if (is_csc) ref_linear_csc(n, d, data, indices, indptr, w, out);
  else ref_linear_csr(n, data, indices, indptr, w, out);
  for (i64 i = 0; i < n; i++) out[i] += intercept;
  i64 dd = d + nAug;
  for (int order = 0; order < nOrders; order++)
    for (int s = 0; s < k; s++) {
      const double *Ps = P + ((i64)order * k + s) * dd;
      int deg = degree - order;
      if (is_csc) ref_anova_csc(n, d, nAug, data, indices, indptr, Ps, A, astride, deg);
      else ref_anova_csr(n, d, nAug, data, indices, indptr, Ps, A, astride, deg);
      double lam = lams ? lams[s] : 1.0;
      for (i64 i = 0; i < n; i++) out[i] += lam * A[i * astride + deg];
    }
  free(A);
}

/* ------------------------------------------------------------------ */
/* per-sample forward + gradient: optimizer/sgd.nim:146-202           */
/* ------------------------------------------------------------------ */
/* computeAnova, sgd.nim:146-173.  Po = P[order] feature-major [d+nAug][k]; A is [k][degree+1]. */
static double compute_anova(const double *Po, i64 d, int nAug, int k, const double *data,
                            const i64 *indices, i64 b, i64 e, int degree, double *A, int astride) {
  double result = 0.0;
  if (degree != 2) {
    for (int s = 0; s < k; s++) {
      A[s * astride] = 1.0;
      for (int t = 1; t < degree + 1; t++) A[s * astride + t] = 0;
    }
    for (i64 jj = b; jj < e + nAug; jj++) {
      i64 j = jj < e ? indices[jj] : d + (jj - e);
      double val = jj < e ? data[jj] : 1.0;
      const double *pj = Po + j * k;
      for (int s = 0; s < k; s++) {
        double *a = A + s * astride;
        for (int t = 0; t < degree; t++) a[degree - t] += a[degree - t - 1] * pj[s] * val;
      }
    }
  } else {
    for (int s = 0; s < k; s++) {
      A[s * astride] = 1;
      A[s * astride + 1] = 0;
      A[s * astride + 2] = 0;
    }
    for (i64 jj = b; jj < e + nAug; jj++) {
      i64 j = jj < e ? indices[jj] : d + (jj - e);
      double val = jj < e ? data[jj] : 1.0;
      const double *pj = Po + j * k;
      for (int s = 0; s < k; s++) {
        double vp = val * pj[s];
        A[s * astride + 1] += vp;
        A[s * astride + 2] += vp * vp;
      }
    }
    for (int s = 0; s < k; s++) {
      double *a = A + s * astride;
      a[2] = (a[1] * a[1] - a[2]) / 2;
    }
  }
  for (int s = 0; s < k; s++) result += A[s * astride + degree];
  return result;
}

/* computeAnovaDerivative, sgd.nim:176-188.  dAo is dense [d+nAug][k]; only the row's features are written. */
static void compute_anova_derivative(const double *Po, i64 d, int nAug, int k, const double *data,
                                     const i64 *indices, i64 b, i64 e, int degree,
                                     const double *A, int astride, double *dAo) {
  for (i64 jj = b; jj < e + nAug; jj++) {
    i64 j = jj < e ? indices[jj] : d + (jj - e);
    double val = jj < e ? data[jj] : 1.0;
    const double *pj = Po + j * k;
    double *dj = dAo + j * k;
    if (degree != 2) {
      for (int s = 0; s < k; s++) {
        dj[s] = val;
        for (int t = 1; t < degree; t++) dj[s] = val * (A[s * astride + t] - pj[s] * dj[s]);
      }
    } else {
      for (int s = 0; s < k; s++) dj[s] = val * (A[s * astride + 1] - pj[s] * val);
    }
  }
}

/* predictWithGrad, sgd.nim:191-202.  P feature-major [nOrders][d+nAug][k], dA same shape (dense scratch). */
double ref_predict_with_grad(i64 d, const double *data, const i64 *indices, const i64 *indptr,
                             i64 i, int degree, int k, int nOrders, int nAug, const double *P,
                             const double *w, double intercept, double *A, double *dA) {
  double result = intercept;
  i64 b = indptr[i], e = indptr[i + 1];
  i64 dd = d + nAug;
  int astride = degree + 1;
  for (i64 jj = b; jj < e; jj++) result += w[indices[jj]] * data[jj];
  for (int order = 0; order < nOrders; order++) {
    const double *Po = P + (i64)order * dd * k;
    result += compute_anova(Po, d, nAug, k, data, indices, b, e, degree - order, A, astride);
    compute_anova_derivative(Po, d, nAug, k, data, indices, b, e, degree - order, A, astride,
                             dA + (i64)order * dd * k);
  }
  return result;
}

/* sgd.transpose (sgd.nim:92-96) in both directions. */
static void to_feature_major(const double *Pm, double *Pf, int nOrders, int k, i64 dd) {
  for (int o = 0; o < nOrders; o++)
    for (i64 j = 0; j < dd; j++)
      for (int s = 0; s < k; s++) Pf[((i64)o * dd + j) * k + s] = Pm[((i64)o * k + s) * dd + j];
}
static void to_component_major(const double *Pf, double *Pm, int nOrders, int k, i64 dd) {
  for (int o = 0; o < nOrders; o++)
    for (i64 j = 0; j < dd; j++)
      for (int s = 0; s < k; s++) Pm[((i64)o * k + s) * dd + j] = Pf[((i64)o * dd + j) * k + s];
}

/* getEta, sgd.nim:60-69 */
static double get_eta(int sched, double eta0, double power, double reg, i64 it) {
  switch (sched) {
  case SCHED_CONSTANT: return eta0;
  case SCHED_OPTIMAL: return eta0 / pow(1.0 + eta0 * reg * (double)it, power);
  case SCHED_INVSCALING: return eta0 / pow((double)it, power);
  default: return 1.0 / (reg * (double)it);
  }
}

/* regularization, optimizer/utils.nim:56-59 (norm(.,2)^2: sqrt of the sum of squares, then squared) */
double ref_regularization(const double *P, i64 nP, const double *w, i64 nw, double intercept,
                          double alpha0, double alpha, double beta) {
  double sw = 0.0, sp = 0.0;
  for (i64 i = 0; i < nw; i++) sw += w[i] * w[i];
  for (i64 i = 0; i < nP; i++) sp += P[i] * P[i];
  double nw2 = sqrt(sw), np2 = sqrt(sp);
  double result = 0.5 * alpha0 * (intercept * intercept) + 0.5 * alpha * (nw2 * nw2);
  result += 0.5 * beta * (np2 * np2);
  return result;
}

/* softthreshold, regularizer/utils.nim:4-5 */
static double softthreshold(double x, double a) {
  double s = (x > 0) - (x < 0);
  double m = fabs(x) - a;
  return s * (m > 0.0 ? m : 0.0);
}

/* ------------------------------------------------------------------ */
/* Proximal operators used by MBPSGD (minibatch_psgd.nim:119-121)      */
/* ------------------------------------------------------------------ */
/* Deterministic stand-in for Nim's rand(max) (stdlib random, not under /root/reference): the pivot
 * choice only affects the ORDER in which proxSquaredL12 finds theta, never theta or the set it sums
 * (tests/test_oracle.py checks the result against the sort-based closed form). */
static unsigned long long prox_rng_state = 0x9E3779B97F4A7C15ull;
static i64 prox_rand(i64 maxIncl) {
  prox_rng_state ^= prox_rng_state << 13;
  prox_rng_state ^= prox_rng_state >> 7;
  prox_rng_state ^= prox_rng_state << 17;
  return (i64)(prox_rng_state % (unsigned long long)(maxIncl + 1));
}
static void swap_i64(i64 *a, i64 *b) { i64 t = *a; *a = *b; *b = t; }

/* proxSquaredL12, regularizer/squaredl12.nim:16-64 (p strided so that a column of P[j][s] can be
 * passed in place; candidates must hold n entries) */
void ref_prox_squaredl12(double *p, i64 n, i64 stride, double lam, i64 *candidates) {
  double S = 0.0;
  i64 theta = 0, offset = 0, nCandidates = n;
  for (i64 i = 0; i < n; i++) candidates[i] = i;
  while (nCandidates != 0) {                                               /* :33 */
    i64 ii = prox_rand(nCandidates - 1);
    i64 i = candidates[offset + ii];
    double pivot = fabs(p[i * stride]);
    swap_i64(&candidates[offset + ii], &candidates[offset + nCandidates - 1]);
    i64 nG = 1, nL = 0;
    double SGi = pivot;
    for (i64 ii2 = 0; ii2 < nCandidates - 1; ii2++) {                      /* :44-51 */
      i64 i2 = candidates[offset + ii2];
      if (pivot > fabs(p[i2 * stride])) {
        swap_i64(&candidates[offset + nL], &candidates[offset + ii2]);
        nL++;
      } else {
        nG++;
        SGi += fabs(p[i2 * stride]);
      }
    }
    if (pivot > 2 * lam * (S + SGi) / (1.0 + 2.0 * lam * (double)(theta + nG))) {   /* :53 L */
      nCandidates = nL;
      S += SGi;
      theta += nG;
    } else {                                                               /* :58-66 G */
      offset = offset + nL;
      nCandidates = 0;
      for (i64 ii2 = 0; ii2 < nG - 1; ii2++) {
        i64 i2 = candidates[offset + ii2];
        if (pivot < fabs(p[i2 * stride])) {
          swap_i64(&candidates[offset + ii2], &candidates[offset + nCandidates]);
          nCandidates++;
        }
      }
    }
  }
  S /= 1.0 + 2.0 * lam * (double)theta;                                    /* :67 */
  for (i64 i = 0; i < n; i++) p[i * stride] = softthreshold(p[i * stride], 2 * lam * S);
}

/* L21.prox of one row, regularizer/l21.nim:25-29 (norm = sqrt of the sequential sum of squares) */
void ref_prox_l21_row(double *pj, i64 k, double lam) {
  double nrm = 0.0;
  for (i64 s = 0; s < k; s++) nrm += pj[s] * pj[s];
  nrm = sqrt(nrm);
  if (nrm > lam) {
    double f = 1.0 - lam / nrm;
    for (i64 s = 0; s < k; s++) pj[s] *= f;
  } else {
    for (i64 s = 0; s < k; s++) pj[s] = 0.0;
  }
}

/* reg.prox(P, lam, degree) on one order's SOLVER-layout matrix P[j][s] (dd x k):
 * reg_kind 1 = L1 (l1.nim:38-41), 2 = SquaredL12 transpose=true (squaredl12.nim:147-156: one vector per
 * component s over all features), 3 = SquaredL12 transpose=false (:157-159: one vector per feature),
 * 4 = L21 (l21.nim:33-35). */
void ref_prox_matrix(double *P, i64 dd, int k, double lam, int reg_kind) {
  if (reg_kind == 1) {
    for (i64 q = 0; q < dd * k; q++) P[q] = softthreshold(P[q], lam);
  } else if (reg_kind == 2) {
    i64 *cand = (i64 *)malloc(sizeof(i64) * (size_t)(dd > 0 ? dd : 1));
    for (int s = 0; s < k; s++) ref_prox_squaredl12(P + s, dd, k, lam, cand);
    free(cand);
  } else if (reg_kind == 3) {
    i64 *cand = (i64 *)malloc(sizeof(i64) * (size_t)(k > 0 ? k : 1));
    for (i64 j = 0; j < dd; j++) ref_prox_squaredl12(P + j * k, k, 1, lam, cand);
    free(cand);
  } else if (reg_kind == 4) {
    for (i64 j = 0; j < dd; j++) ref_prox_l21_row(P + j * k, k, lam);
  }
}

/* reg.eval(P) on one order's solver-layout matrix (l1.nim eval, l21.nim:17-18, squaredl12.nim:72-75) */
double ref_reg_eval(const double *P, i64 dd, int k, int reg_kind) {
  double r = 0.0;
  if (reg_kind == 1) {
    for (i64 q = 0; q < dd * k; q++) r += fabs(P[q]);
  } else if (reg_kind == 2) {
    for (int s = 0; s < k; s++) {
      double c = 0.0;
      for (i64 j = 0; j < dd; j++) c += fabs(P[j * k + s]);
      r += c * c;
    }
  } else if (reg_kind == 3) {
    for (i64 j = 0; j < dd; j++) {
      double c = 0.0;
      for (int s = 0; s < k; s++) c += fabs(P[j * k + s]);
      r += c * c;
    }
  } else if (reg_kind == 4) {
    for (i64 j = 0; j < dd; j++) {
      double c = 0.0;
      for (int s = 0; s < k; s++) c += P[j * k + s] * P[j * k + s];
      r += sqrt(c);
    }
  }
  return r;
}

/* ------------------------------------------------------------------ */
/* MBPSGD: optimizer/minibatch_psgd.nim:67-211, model/params.nim      */
/* ------------------------------------------------------------------ */
/* reg_kind: 0 = identity prox (SquaredL12 with gamma=0, squaredl12.nim:67-69), 1 = L1 (l1.nim:38-41),
 * 2 / 3 = SquaredL12 transpose=true / false, 4 = L21 (see ref_prox_matrix).
 * perms: NULL (shuffle=false, cyclic order) or [nPerms][n] host-supplied permutations; the first is
 * used from the start and the next one replaces it each time the cursor wraps (:107-111,169-170).
 * P/w/intercept in/out use the MODEL layout; it_io is MBPSGD.it (1 unless warm-started, :151-152).
 * epoch_loss[maxIter] receives runningLoss of each epoch (:124). Returns epochs run. */
int ref_mbpsgd_fit(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                   const double *y, int degree, int k, int nOrders, int nAug, int fitLinear,
                   int fitIntercept, double *Pm, double *w, double *intercept_io, int loss_kind,
                   double thr, int maxIter, double eta0, double alpha0, double alpha, double beta,
                   double gamma, int reg_kind, i64 miniBatchSize, i64 maxIterInner, int sched,
                   double power, double tol, const i64 *perms, i64 nPerms, i64 *it_io,
                   double *epoch_loss) {
  i64 dd = d + nAug, nP = (i64)nOrders * dd * k;
  int astride = degree + 1;
  double *P = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *gP = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *gw = (double *)calloc((size_t)(d > 0 ? d : 1), sizeof(double));
  double *dA = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *A = (double *)calloc((size_t)k * astride, sizeof(double));
  i64 *idx = (i64 *)malloc(sizeof(i64) * (size_t)(n > 0 ? n : 1));
  double intercept = *intercept_io, gb = 0.0;
  i64 it = *it_io, permUsed = 0;
  to_feature_major(Pm, P, nOrders, k, dd);                       /* :155-156 */
  i64 nnz = indptr[n] + n * 0;                                   /* X.nnz with nAugments==0 here (:158-160) */
  if (miniBatchSize <= 0) {
    miniBatchSize = (d * n) / nnz;
    if (miniBatchSize < 1) miniBatchSize = 1;
  }
  if (maxIterInner <= 0) {
    maxIterInner = (n - 1) / miniBatchSize + 1;
    if (maxIterInner < 1) maxIterInner = 1;
  }
  for (i64 i = 0; i < n; i++) idx[i] = i;
  if (perms && nPerms > 0) { memcpy(idx, perms, sizeof(i64) * (size_t)n); permUsed = 1; }  /* :169-170 */
  i64 ii = 0;
  double oldLoss = INFINITY;
  int epochs = 0;
  for (int ep = 0; ep < maxIter; ep++) {
    double result = 0.0;                                         /* epoch(), :91-124 */
    for (i64 itInner = 0; itInner < maxIterInner; itInner++) {
      memset(gP, 0, sizeof(double) * (size_t)nP);                /* grads <- 0.0 (:99) */
      memset(gw, 0, sizeof(double) * (size_t)d);
      gb = 0.0;
      for (i64 b = 0; b < miniBatchSize; b++) {
        i64 i = idx[ii];
        /* updateGradient, :67-88 */
        double yPred = ref_predict_with_grad(d, data, indices, indptr, i, degree, k, nOrders, nAug,
                                             P, w, intercept, A, dA);
        result += ref_loss(loss_kind, thr, y[i], yPred);
        double coef = 1.0 * ref_dloss(loss_kind, thr, y[i], yPred) / (double)miniBatchSize;
        i64 rb = indptr[i], re = indptr[i + 1];
        for (int o = 0; o < nOrders; o++)
          for (i64 jj = rb; jj < re + nAug; jj++) {
            i64 j = jj < re ? indices[jj] : d + (jj - re);
            for (int s = 0; s < k; s++) gP[((i64)o * dd + j) * k + s] += coef * dA[((i64)o * dd + j) * k + s];
          }
        if (fitLinear)
          for (i64 jj = rb; jj < re; jj++) gw[indices[jj]] += coef * data[jj];
        if (fitIntercept) gb += coef;
        ii++;
        if (ii >= n) {                                           /* :108-111 */
          ii = 0;
          if (perms && permUsed < nPerms) { memcpy(idx, perms + permUsed * n, sizeof(i64) * (size_t)n); permUsed++; }
        }
      }
      double eta_P = get_eta(sched, eta0, power, beta, it);      /* :114-116 */
      double eta_w = get_eta(sched, eta0, power, alpha, it);
      double eta_b = get_eta(sched, eta0, power, alpha0, it);
      /* Params.step, params.nim:90-98 = add (:33-48) then scale (:61-66) */
      double scale_P = 1.0 + eta_P * beta, scale_w = 1.0 + eta_w * alpha, scale_b = 1.0 + eta_b * alpha0;
      for (i64 q = 0; q < nP; q++) P[q] += (-eta_P) * gP[q];
      if (fitLinear) for (i64 j = 0; j < d; j++) w[j] += (-eta_w) * gw[j];
      if (fitIntercept && fitLinear) intercept += (-eta_b) * gb;  /* params.nim:47 quirk */
      double rP = 1.0 / scale_P, rw = 1.0 / scale_w, rb_ = 1.0 / scale_b;
      for (i64 q = 0; q < nP; q++) P[q] *= rP;
      if (fitLinear) for (i64 j = 0; j < d; j++) w[j] *= rw;
      if (fitIntercept) intercept *= rb_;
      /* reg.prox(params.P[order], gamma*eta_P/(1+eta_P*beta), degree-order), :119-121 */
      if (reg_kind != 0) {
        double lam = gamma * eta_P / (1.0 + eta_P * beta);
        for (int o = 0; o < nOrders; o++) ref_prox_matrix(P + (i64)o * dd * k, dd, k, lam, reg_kind);
      }
      it++;
    }
    result /= (double)(miniBatchSize * maxIterInner);
    epoch_loss[ep] = result;
    epochs = ep + 1;
    if (isnan(result)) break;                                    /* :189-191 */
    if (fabs(oldLoss - result) < tol) break;                     /* :200-203 */
    oldLoss = result;
  }
  to_component_major(P, Pm, nOrders, k, dd);                     /* pgd.finalize, pgd.nim:45-51 */
  *intercept_io = intercept;
  *it_io = it;
  free(P); free(gP); free(gw); free(dA); free(A); free(idx);
  return epochs;
}

/* One predict+grad pass (updateGradient, minibatch_psgd.nim:67-88) over rows [rowBegin,rowEnd) with
 * the SOLVER layout P[order][j][s]; used by bench.py's CPU baseline and by the K2 parity tests.
 * gP/gw/gb are accumulated into (caller zeroes them); returns the loss sum. */
double ref_fm_loss_grad(i64 d, const double *data, const i64 *indices, const i64 *indptr,
                        const double *y, i64 rowBegin, i64 rowEnd, int degree, int k, int nOrders,
                        int nAug, int fitLinear, int fitIntercept, const double *P, const double *w,
                        double intercept, int loss_kind, double thr, i64 miniBatchSize,
                        double *gP, double *gw, double *gb, double *yPredOut, double *dA_scratch) {
  i64 dd = d + nAug;
  int astride = degree + 1;
  double *A = (double *)calloc((size_t)k * astride, sizeof(double));
  double *dA = dA_scratch;
  double lossSum = 0.0;
  for (i64 i = rowBegin; i < rowEnd; i++) {
    double yPred = ref_predict_with_grad(d, data, indices, indptr, i, degree, k, nOrders, nAug, P, w,
                                         intercept, A, dA);
    if (yPredOut) yPredOut[i - rowBegin] = yPred;
    lossSum += ref_loss(loss_kind, thr, y[i], yPred);
    double coef = 1.0 * ref_dloss(loss_kind, thr, y[i], yPred) / (double)miniBatchSize;
    i64 rb = indptr[i], re = indptr[i + 1];
    for (int o = 0; o < nOrders; o++)
      for (i64 jj = rb; jj < re + nAug; jj++) {
        i64 j = jj < re ? indices[jj] : d + (jj - re);
        for (int s = 0; s < k; s++) gP[((i64)o * dd + j) * k + s] += coef * dA[((i64)o * dd + j) * k + s];
      }
    if (fitLinear)
      for (i64 jj = rb; jj < re; jj++) gw[indices[jj]] += coef * data[jj];
    if (fitIntercept) *gb += coef;
  }
  free(A);
  return lossSum;
}

/* ------------------------------------------------------------------ */
/* AdaGrad: optimizer/adagrad.nim:47-203, fit_linear.nim:50-57        */
/* ------------------------------------------------------------------ */
/* mb == 1 is exactly the reference's per-sample loop.  mb > 1 is the "synchronous minibatch"
 * variant defined in DESIGN.md: the mb samples of a batch share one parameter snapshot
 * (refresh with t = it-1 of the batch start, gradients from the snapshot, G/N accumulated after
 * the whole batch, viol counted per (row, feature) incidence against the pre-batch P); it += mb.
 * g_sum/g_norm (feature-major P part, w part, intercept part) are in/out so warm starts work.
 * perms: NULL or [maxIter][n].  viol_out/loss_out are per epoch. Model layout in/out for P. */
int ref_adagrad_fit(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                    const double *y, int degree, int k, int nOrders, int nAug, int fitLinear,
                    int fitIntercept, double *Pm, double *w, double *intercept_io, int loss_kind,
                    double thr, int maxIter, double eta0, double alpha0, double alpha, double beta,
                    double eps, double tol, i64 mb, const i64 *perms, i64 *it_io, double *gsP,
                    double *gnP, double *gsw, double *gnw, double *gsb_io, double *gnb_io,
                    double *viol_out, double *loss_out) {
  i64 dd = d + nAug, nP = (i64)nOrders * dd * k;
  int astride = degree + 1;
  double *P = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *dA = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *A = (double *)calloc((size_t)k * astride, sizeof(double));
  double *dLs = (double *)malloc(sizeof(double) * (size_t)(mb > 0 ? mb : 1));
  double *Pnew = NULL, *wnew = NULL;
  if (mb > 1) {
    Pnew = (double *)malloc(sizeof(double) * (size_t)(nP > 0 ? nP : 1));
    wnew = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  }
  double intercept = *intercept_io;
  i64 it = *it_io;
  if (it == 1) {                                                  /* init, :47-62 */
    for (i64 q = 0; q < nP; q++) { gsP[q] = 0.0; gnP[q] = eps; }
    for (i64 j = 0; j < d; j++) { gsw[j] = 0.0; gnw[j] = eps; }
    *gsb_io = 0.0; *gnb_io = eps;
  }
  double gsb = *gsb_io, gnb = *gnb_io;
  to_feature_major(Pm, P, nOrders, k, dd);
  int epochs = 0;
  for (int ep = 0; ep < maxIter; ep++) {
    double viol = 0.0, runningLoss = 0.0;
    const i64 *perm = perms ? perms + (i64)ep * n : NULL;
    for (i64 start = 0; start < n; start += mb) {
      i64 end = start + mb < n ? start + mb : n;
      if (mb > 1) { memcpy(Pnew, P, sizeof(double) * (size_t)nP); memcpy(wnew, w, sizeof(double) * (size_t)d); }
      double *Pt = mb > 1 ? Pnew : P, *wt = mb > 1 ? wnew : w;
      double bnew = intercept;
      if (it != 1) {
        double itf = (double)(it - 1);
        double tmp = eta0 * itf * beta;
        for (i64 q = start; q < end; q++) {
          i64 i = perm ? perm[q] : q;
          i64 rb = indptr[i], re = indptr[i + 1];
          /* update(), :87-110 */
          for (int o = 0; o < nOrders; o++)
            for (i64 jj = rb; jj < re + nAug; jj++) {
              i64 j = jj < re ? indices[jj] : d + (jj - re);
              for (int s = 0; s < k; s++) {
                i64 e = ((i64)o * dd + j) * k + s;
                double pjs = P[e];
                double denom = tmp + sqrt(gnP[e]);
                Pt[e] = -(eta0 * gsP[e]) / denom;
                viol += fabs(pjs - Pt[e]);
              }
            }
          if (fitIntercept && q == start) {                       /* once per batch (per sample at mb==1) */
            double old = intercept;
            double denom = sqrt(gnb) + eta0 * itf * alpha0;
            bnew = -eta0 * gsb / denom;
            viol += fabs(old - bnew);
          }
          if (fitLinear) {                                        /* fitLinearAdaGrad, fit_linear.nim:50-57 */
            double denom = itf * eta0 * alpha;
            for (i64 jj = rb; jj < re; jj++) {
              i64 j = indices[jj];
              double wj = w[j];
              wt[j] = -eta0 * gsw[j] / (denom + sqrt(gnw[j]));
              viol += fabs(wj - wt[j]);
            }
          }
        }
      }
      intercept = bnew;
      if (mb > 1) { memcpy(P, Pnew, sizeof(double) * (size_t)nP); memcpy(w, wnew, sizeof(double) * (size_t)d); }
      /* forward + gradient from the snapshot, then updateG (:113-134) */
      for (i64 q = start; q < end; q++) {
        i64 i = perm ? perm[q] : q;
        if (mb == 1) {
          double yPred = ref_predict_with_grad(d, data, indices, indptr, i, degree, k, nOrders, nAug,
                                               P, w, intercept, A, dA);
          runningLoss += ref_loss(loss_kind, thr, y[i], yPred);
          double dL = ref_dloss(loss_kind, thr, y[i], yPred);
          i64 rb = indptr[i], re = indptr[i + 1];
          for (int o = 0; o < nOrders; o++)
            for (i64 jj = rb; jj < re + nAug; jj++) {
              i64 j = jj < re ? indices[jj] : d + (jj - re);
              for (int s = 0; s < k; s++) {
                i64 e = ((i64)o * dd + j) * k + s;
                double grad = dL * dA[e];
                gsP[e] += grad;
                gnP[e] += grad * grad;
              }
            }
          if (fitIntercept) { gsb += dL; gnb += dL * dL; }
          if (fitLinear)
            for (i64 jj = rb; jj < re; jj++) {
              i64 j = indices[jj];
              double val = data[jj];
              gsw[j] += dL * val;
              gnw[j] += (dL * val) * (dL * val);
            }
        } else {
          double yPred = ref_predict_with_grad(d, data, indices, indptr, i, degree, k, nOrders, nAug,
                                               P, w, intercept, A, dA);
          runningLoss += ref_loss(loss_kind, thr, y[i], yPred);
          dLs[q - start] = ref_dloss(loss_kind, thr, y[i], yPred);
          /* P is not modified inside a batch, so G/N can be accumulated right away */
          double dL = dLs[q - start];
          i64 rb = indptr[i], re = indptr[i + 1];
          for (int o = 0; o < nOrders; o++)
            for (i64 jj = rb; jj < re + nAug; jj++) {
              i64 j = jj < re ? indices[jj] : d + (jj - re);
              for (int s = 0; s < k; s++) {
                i64 e = ((i64)o * dd + j) * k + s;
                double grad = dL * dA[e];
                gsP[e] += grad;
                gnP[e] += grad * grad;
              }
            }
          if (fitIntercept) { gsb += dL; gnb += dL * dL; }
          if (fitLinear)
            for (i64 jj = rb; jj < re; jj++) {
              i64 j = indices[jj];
              double val = data[jj];
              gsw[j] += dL * val;
              gnw[j] += (dL * val) * (dL * val);
            }
        }
      }
      it += (end - start);
    }
    runningLoss /= (double)n;
    viol_out[ep] = viol;
    loss_out[ep] = runningLoss;
    epochs = ep + 1;
    if (isnan(runningLoss)) break;                                /* stoppingCriterion, sgd.nim:72-89 */
    if (viol < tol) break;
  }
  /* finalize, :65-84 */
  {
    double itf = (double)(it - 1);
    double denom = eta0 * itf * beta;
    for (i64 q = 0; q < nP; q++) {
      P[q] = -eta0 * gsP[q];
      P[q] /= denom + sqrt(gnP[q]);
    }
    if (fitIntercept) {
      double den = sqrt(gnb) + eta0 * itf * alpha0;
      intercept = -eta0 * gsb / den;
    }
    if (fitLinear) {
      double den = eta0 * itf * alpha;
      for (i64 j = 0; j < d; j++) {
        w[j] = -eta0 * gsw[j];
        w[j] /= den + sqrt(gnw[j]);
      }
    }
  }
  to_component_major(P, Pm, nOrders, k, dd);
  *intercept_io = intercept;
  *it_io = it;
  *gsb_io = gsb; *gnb_io = gnb;
  free(P); free(dA); free(A); free(dLs); free(Pnew); free(wnew);
  return epochs;
}

/* ------------------------------------------------------------------ */
/* SGD: optimizer/sgd.nim:99-143, 205-328, fit_linear.nim:41-47       */
/* ------------------------------------------------------------------ */
int ref_sgd_fit(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
                const double *y, int degree, int k, int nOrders, int nAug, int fitLinear,
                int fitIntercept, double *Pm, double *w, double *intercept_io, int loss_kind,
                double thr, int maxIter, double eta0, double alpha0, double alpha, double beta,
                int sched, double power, double tol, const i64 *perms, i64 *it_io,
                double *viol_out, double *loss_out) {
  i64 dd = d + nAug, nP = (i64)nOrders * dd * k;
  int astride = degree + 1;
  double *P = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *dA = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *A = (double *)calloc((size_t)k * astride, sizeof(double));
  double *scalings_w = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  double *scalings_P = (double *)malloc(sizeof(double) * (size_t)(dd > 0 ? dd : 1));
  for (i64 j = 0; j < d; j++) scalings_w[j] = 1.0;
  for (i64 j = 0; j < dd; j++) scalings_P[j] = 1.0;
  double scaling_w = 1.0, scaling_P = 1.0, intercept = *intercept_io;
  i64 it = *it_io;
  to_feature_major(Pm, P, nOrders, k, dd);
  int epochs = 0;
  for (int ep = 0; ep < maxIter; ep++) {
    double viol = 0.0, runningLoss = 0.0;
    const i64 *perm = perms ? perms + (i64)ep * n : NULL;
    for (i64 q = 0; q < n; q++) {
      i64 i = perm ? perm[q] : q;
      i64 rb = indptr[i], re = indptr[i + 1];
      /* lazilyUpdate, :134-143 (getRowIndices without dummy features: nAugments is 0 here) */
      for (int o = 0; o < nOrders; o++)
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = indices[jj];
          for (int s = 0; s < k; s++) P[((i64)o * dd + j) * k + s] *= scaling_P / scalings_P[j];
        }
      if (fitLinear)
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = indices[jj];
          w[j] *= scaling_w / scalings_w[j];
        }
      double yPred = ref_predict_with_grad(d, data, indices, indptr, i, degree, k, nOrders, nAug, P,
                                           w, intercept, A, dA);
      runningLoss += ref_loss(loss_kind, thr, y[i], yPred);
      /* update, :205-243 */
      double dL = ref_dloss(loss_kind, thr, y[i], yPred);
      double eta_w = get_eta(sched, eta0, power, alpha, it);
      double eta_P = get_eta(sched, eta0, power, beta, it);
      for (int o = 0; o < nOrders; o++)
        for (i64 jj = rb; jj < re + nAug; jj++) {
          i64 j = jj < re ? indices[jj] : d + (jj - re);
          for (int s = 0; s < k; s++) {
            i64 e = ((i64)o * dd + j) * k + s;
            double update = eta_P * (dL * dA[e] + beta * P[e]);
            viol += fabs(update);
            P[e] -= update;
          }
        }
      if (fitIntercept) {
        double update = get_eta(sched, eta0, power, alpha0, it) * (dL + alpha0 * intercept);
        viol += fabs(update);
        intercept -= update;
      }
      if (fitLinear)                                               /* fitLinearSGD, fit_linear.nim:41-47 */
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = indices[jj];
          double update = eta_w * (dL * data[jj] + alpha * w[j]);
          w[j] -= update;
          viol += fabs(update);
        }
      scaling_P *= (1 - eta_P * beta);
      scaling_w *= (1 - eta_w * alpha);
      for (i64 jj = rb; jj < re; jj++) {
        i64 j = indices[jj];
        scalings_P[j] = scaling_P;
        scalings_w[j] = scaling_w;
      }
      for (int a = 0; a < nAug; a++) scalings_P[d + a] = scaling_P;
      /* resetScaling, :116-131 */
      if (fitLinear && scaling_w < 1e-9) {
        for (i64 j = 0; j < d; j++) { w[j] *= scaling_w; w[j] /= scalings_w[j]; scalings_w[j] = 1.0; }
        scaling_w = 1.0;
      }
      if (scaling_P < 1e-9) {
        for (int o = 0; o < nOrders; o++)
          for (i64 j = 0; j < d; j++)
            for (int s = 0; s < k; s++) P[((i64)o * dd + j) * k + s] *= scaling_P / scalings_P[j];
        for (i64 j = 0; j < dd; j++) scalings_P[j] = 1.0;
        scaling_P = 1.0;
      }
      it++;
    }
    runningLoss /= (double)n;
    viol_out[ep] = viol;
    loss_out[ep] = runningLoss;
    epochs = ep + 1;
    if (isnan(runningLoss)) break;
    if (viol < tol) break;
  }
  /* finalize, :99-113 */
  if (fitLinear)
    for (i64 j = 0; j < d; j++) { w[j] *= scaling_w; w[j] /= scalings_w[j]; }
  for (int o = 0; o < nOrders; o++)
    for (i64 j = 0; j < dd; j++)
      for (int s = 0; s < k; s++) P[((i64)o * dd + j) * k + s] *= scaling_P / scalings_P[j];
  to_component_major(P, Pm, nOrders, k, dd);
  *intercept_io = intercept;
  *it_io = it;
  free(P); free(dA); free(A); free(scalings_w); free(scalings_P);
  return epochs;
}

/* ------------------------------------------------------------------ */
/* CD: optimizer/cd.nim:29-194, fit_linear.nim:5-38, extmath.nim:151-163 */
/* ------------------------------------------------------------------ */
/* X is CSC.  P model layout [nOrders][k][d+nAug] (CD does NOT transpose).  viol_out/loss_out/reg_out
 * per outer iteration: viol, mean loss, regularization/n (cd.nim:177-184). yPred_out[n] optional. */
int ref_cd_fit(i64 n, i64 d, const double *data, const i64 *indices, const i64 *indptr,
               const double *y, int degree, int k, int nOrders, int nAug, int fitLinear,
               int fitIntercept, double *P, double *w, double *intercept_io, int loss_kind,
               double thr, int maxIter, double alpha0_, double alpha_, double beta_, double tol,
               double *viol_out, double *loss_out, double *reg_out, double *yPred_out) {
  i64 dd = d + nAug;
  double alpha0 = alpha0_ * (double)n, alpha = alpha_ * (double)n, beta = beta_ * (double)n; /* :123-125 */
  int astride = degree + 1;
  double mu = ref_mu(loss_kind);
  double *yPred = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double *A = (double *)calloc((size_t)(n > 0 ? n : 1) * astride, sizeof(double));
  double *dA = (double *)calloc((size_t)(degree > 0 ? degree : 1), sizeof(double));
  double *cache = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double *colNormSq = (double *)calloc((size_t)(d > 0 ? d : 1), sizeof(double));
  double intercept = *intercept_io;
  for (i64 i = 0; i < n; i++) A[i * astride] = 1.0;
  if (fitLinear)                                                   /* :141-142, extmath.nim:151-163 */
    for (i64 j = 0; j < d; j++) {
      double s = 0.0;
      for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) s += data[ii] * data[ii];
      double nr = sqrt(s);
      colNormSq[j] = nr * nr;
    }
  ref_linear_csc(n, d, data, indices, indptr, w, yPred);          /* :144-145 */
  for (i64 i = 0; i < n; i++) yPred[i] += intercept;
  for (int o = 0; o < nOrders; o++)                                /* :147-151 */
    for (int s = 0; s < k; s++) {
      ref_anova_csc(n, d, nAug, data, indices, indptr, P + ((i64)o * k + s) * dd, A, astride, degree - o);
      for (i64 i = 0; i < n; i++) yPred[i] += A[i * astride + degree - o];
    }
  int iters = 0;
  for (int it = 0; it < maxIter; it++) {
    double viol = 0.0;
    if (fitIntercept) {                                            /* fitInterceptCD, fit_linear.nim:28-38 */
      double r = alpha0 * intercept;
      for (i64 i = 0; i < n; i++) r += ref_dloss(loss_kind, thr, y[i], yPred[i]);
      r /= mu * (double)n + alpha0;
      intercept -= r;
      for (i64 i = 0; i < n; i++) yPred[i] -= r;
      viol += fabs(r);
    }
    if (fitLinear) {                                               /* fitLinearCD, fit_linear.nim:5-25 */
      double res = 0.0;
      for (i64 j = 0; j < d; j++) {
        double update = alpha * w[j];
        for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++)
          update += ref_dloss(loss_kind, thr, y[indices[ii]], yPred[indices[ii]]) * data[ii];
        double inv = mu * colNormSq[j] + alpha;
        if (inv < 1e-12) continue;
        update /= inv;
        res += fabs(update);
        w[j] -= update;
        for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) yPred[indices[ii]] -= update * data[ii];
      }
      viol += res;
    }
    for (int o = 0; o < nOrders; o++) {
      int deg = degree - o;
      double *Po = P + (i64)o * k * dd;
      double res = 0.0;
      if (deg > 2) {                                               /* epoch, cd.nim:50-73 */
        for (int s = 0; s < k; s++) {
          double *Ps = Po + (i64)s * dd;
          ref_anova_csc(n, d, nAug, data, indices, indptr, Ps, A, astride, deg);
          for (i64 j = 0; j < dd; j++) {
            double psj = Ps[j];
            i64 cb = j < d ? indptr[j] : 0, ce = j < d ? indptr[j + 1] : n;
            double update = beta * psj, inv = 0.0;                 /* update(), :36-47 */
            for (i64 ii = cb; ii < ce; ii++) {
              i64 i = j < d ? indices[ii] : ii;
              double val = j < d ? data[ii] : 1.0;
              dA[0] = val;                                         /* computeDerivative, :29-33 */
              for (int g = 1; g < deg; g++) dA[g] = val * (A[i * astride + g] - psj * dA[g - 1]);
              update += ref_dloss(loss_kind, thr, y[i], yPred[i]) * dA[deg - 1];
              inv += dA[deg - 1] * dA[deg - 1];
            }
            inv *= mu;
            inv += beta;
            update /= inv;                                         /* no 1e-12 guard here (:63) */
            Ps[j] -= update;
            res += fabs(update);
            for (i64 ii = cb; ii < ce; ii++) {                     /* synchronize, :67-73 */
              i64 i = j < d ? indices[ii] : ii;
              double val = j < d ? data[ii] : 1.0;
              dA[0] = val;
              for (int g = 1; g < deg; g++) {
                dA[g] = val * (A[i * astride + g] - psj * dA[g - 1]);
                A[i * astride + g] -= update * dA[g - 1];
              }
              A[i * astride + deg] -= update * dA[deg - 1];
              yPred[i] -= update * dA[deg - 1];
            }
          }
        }
      } else {                                                     /* epochDeg2, cd.nim:77-107 */
        for (int s = 0; s < k; s++) {
          double *Ps = Po + (i64)s * dd;
          for (i64 i = 0; i < n; i++) cache[i] = 0;
          for (i64 j = 0; j < dd; j++) {
            if (j < d) for (i64 ii = indptr[j]; ii < indptr[j + 1]; ii++) cache[indices[ii]] += data[ii] * Ps[j];
            else for (i64 i = 0; i < n; i++) cache[i] += 1.0 * Ps[j];
          }
          for (i64 j = 0; j < dd; j++) {
            double psj = Ps[j];
            i64 cb = j < d ? indptr[j] : 0, ce = j < d ? indptr[j + 1] : n;
            double update = beta * psj, inv = 0.0;
            for (i64 ii = cb; ii < ce; ii++) {
              i64 i = j < d ? indices[ii] : ii;
              double val = j < d ? data[ii] : 1.0;
              double g = (cache[i] - psj * val) * val;
              update += ref_dloss(loss_kind, thr, y[i], yPred[i]) * g;
              inv += g * g;
            }
            inv = inv * mu + beta;
            if (inv < 1e-12) continue;
            update /= inv;
            res += fabs(update);
            for (i64 ii = cb; ii < ce; ii++) {
              i64 i = j < d ? indices[ii] : ii;
              double val = j < d ? data[ii] : 1.0;
              yPred[i] -= update * (cache[i] - psj * val) * val;
              cache[i] -= update * val;
            }
            Ps[j] -= update;
          }
        }
      }
      viol += res;
    }
    double lossVal = 0.0;                                          /* :177-183 */
    for (i64 i = 0; i < n; i++) lossVal += ref_loss(loss_kind, thr, y[i], yPred[i]);
    lossVal /= (double)n;
    viol_out[it] = viol;
    loss_out[it] = lossVal;
    reg_out[it] = ref_regularization(P, (i64)nOrders * k * dd, w, d, intercept, alpha0, alpha, beta) / (double)n;
    iters = it + 1;
    if (viol < tol) break;                                         /* :186-189 */
  }
  if (yPred_out) memcpy(yPred_out, yPred, sizeof(double) * (size_t)n);
  *intercept_io = intercept;
  free(yPred); free(A); free(dA); free(cache); free(colNormSq);
  return iters;
}

/* ------------------------------------------------------------------ */
/* FFM: model/field_aware_factorization_machine.nim:52-76, optimizer/sgd_ffm.nim:11-46 */
/* ------------------------------------------------------------------ */
/* dot, tensor.nim:686-692: sequential accumulation */
static double dotk(const double *a, const double *b, int k) {
  double r = 0.0;
  for (int s = 0; s < k; s++) r += a[s] * b[s];
  return r;
}

/* decisionFunction, field_aware_factorization_machine.nim:52-76.  P is [nFields][d][k]. */
void ref_ffm_decision_function(i64 n, i64 d, int nFields, int k, const double *data,
                               const i64 *indices, const i64 *indptr, const i64 *fields,
                               const double *P, const double *w, double intercept, double *out) {
  (void)nFields;
  for (i64 i = 0; i < n; i++) {
    out[i] = intercept;
    for (i64 jj = indptr[i]; jj < indptr[i + 1]; jj++) out[i] += data[jj] * w[indices[jj]];
  }
  for (i64 i = 0; i < n; i++)
    for (i64 a = indptr[i]; a < indptr[i + 1]; a++)
      for (i64 b = indptr[i]; b < indptr[i + 1]; b++) {
        i64 j1 = indices[a], j2 = indices[b];
        if (j1 < j2) {
          i64 f1 = fields[a], f2 = fields[b];
          out[i] += data[a] * data[b] * dotk(P + (f2 * d + j1) * k, P + (f1 * d + j2) * k, k);
        }
      }
}

/* predictWithGrad (FFM), sgd_ffm.nim:11-30.  dA is the dense [nFields][d][k] scratch. */
double ref_ffm_predict_with_grad(i64 d, int nFields, int k, const double *data, const i64 *indices,
                                 const i64 *indptr, const i64 *fields, i64 i, const double *P,
                                 const double *w, double intercept, double *dA) {
  double result = intercept;
  i64 rb = indptr[i], re = indptr[i + 1];
  for (i64 jj = rb; jj < re; jj++) result += w[indices[jj]] * data[jj];
  for (int f = 0; f < nFields; f++)
    for (i64 jj = rb; jj < re; jj++)
      for (int s = 0; s < k; s++) dA[((i64)f * d + indices[jj]) * k + s] = 0.0;
  for (i64 a = rb; a < re; a++)
    for (i64 b = rb; b < re; b++) {
      i64 j1 = indices[a], j2 = indices[b];
      if (j1 < j2) {
        i64 f1 = fields[a], f2 = fields[b];
        double val1 = data[a], val2 = data[b];
        const double *p21 = P + (f2 * d + j1) * k, *p12 = P + (f1 * d + j2) * k;
        double tmp = dotk(p21, p12, k);
        result += tmp * val1 * val2;
        for (int s = 0; s < k; s++) {
          dA[(f2 * d + j1) * k + s] += val1 * val2 * p12[s];
          dA[(f1 * d + j2) * k + s] += val1 * val2 * p21[s];
        }
      }
    }
  return result;
}

/* AdaGrad for FFM: optimizer/adagrad_ffm.nim:11-66 reusing adagrad.update/updateG with order==field
 * (all nFields x row features x k entries are refreshed and accumulated).  mb semantics as in
 * ref_adagrad_fit. */
int ref_ffm_adagrad_fit(i64 n, i64 d, int nFields, int k, const double *data, const i64 *indices,
                        const i64 *indptr, const i64 *fields, const double *y, int fitLinear,
                        int fitIntercept, double *P, double *w, double *intercept_io, int loss_kind,
                        double thr, int maxIter, double eta0, double alpha0, double alpha,
                        double beta, double eps, double tol, i64 mb, const i64 *perms, i64 *it_io,
                        double *gsP, double *gnP, double *gsw, double *gnw, double *gsb_io,
                        double *gnb_io, double *viol_out, double *loss_out) {
  i64 nP = (i64)nFields * d * k;
  double *dA = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *Pnew = NULL, *wnew = NULL;
  if (mb > 1) {
    Pnew = (double *)malloc(sizeof(double) * (size_t)(nP > 0 ? nP : 1));
    wnew = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  }
  double intercept = *intercept_io;
  i64 it = *it_io;
  if (it == 1) {
    for (i64 q = 0; q < nP; q++) { gsP[q] = 0.0; gnP[q] = eps; }
    for (i64 j = 0; j < d; j++) { gsw[j] = 0.0; gnw[j] = eps; }
    *gsb_io = 0.0; *gnb_io = eps;
  }
  double gsb = *gsb_io, gnb = *gnb_io;
  int epochs = 0;
  for (int ep = 0; ep < maxIter; ep++) {
    double viol = 0.0, runningLoss = 0.0;
    const i64 *perm = perms ? perms + (i64)ep * n : NULL;
    for (i64 start = 0; start < n; start += mb) {
      i64 end = start + mb < n ? start + mb : n;
      if (mb > 1) { memcpy(Pnew, P, sizeof(double) * (size_t)nP); memcpy(wnew, w, sizeof(double) * (size_t)d); }
      double *Pt = mb > 1 ? Pnew : P, *wt = mb > 1 ? wnew : w;
      double bnew = intercept;
      if (it != 1) {
        double itf = (double)(it - 1);
        double tmp = eta0 * itf * beta;
        for (i64 q = start; q < end; q++) {
          i64 i = perm ? perm[q] : q;
          i64 rb = indptr[i], re = indptr[i + 1];
          for (int f = 0; f < nFields; f++)
            for (i64 jj = rb; jj < re; jj++)
              for (int s = 0; s < k; s++) {
                i64 e = ((i64)f * d + indices[jj]) * k + s;
                double pjs = P[e];
                double denom = tmp + sqrt(gnP[e]);
                Pt[e] = -(eta0 * gsP[e]) / denom;
                viol += fabs(pjs - Pt[e]);
              }
          if (fitIntercept && q == start) {
            double old = intercept;
            double denom = sqrt(gnb) + eta0 * itf * alpha0;
            bnew = -eta0 * gsb / denom;
            viol += fabs(old - bnew);
          }
          if (fitLinear) {
            double denom = itf * eta0 * alpha;
            for (i64 jj = rb; jj < re; jj++) {
              i64 j = indices[jj];
              double wj = w[j];
              wt[j] = -eta0 * gsw[j] / (denom + sqrt(gnw[j]));
              viol += fabs(wj - wt[j]);
            }
          }
        }
      }
      intercept = bnew;
      if (mb > 1) { memcpy(P, Pnew, sizeof(double) * (size_t)nP); memcpy(w, wnew, sizeof(double) * (size_t)d); }
      for (i64 q = start; q < end; q++) {
        i64 i = perm ? perm[q] : q;
        double yPred = ref_ffm_predict_with_grad(d, nFields, k, data, indices, indptr, fields, i, P, w,
                                                 intercept, dA);
        runningLoss += ref_loss(loss_kind, thr, y[i], yPred);
        double dL = ref_dloss(loss_kind, thr, y[i], yPred);
        i64 rb = indptr[i], re = indptr[i + 1];
        for (int f = 0; f < nFields; f++)
          for (i64 jj = rb; jj < re; jj++)
            for (int s = 0; s < k; s++) {
              i64 e = ((i64)f * d + indices[jj]) * k + s;
              double grad = dL * dA[e];
              gsP[e] += grad;
              gnP[e] += grad * grad;
            }
        if (fitIntercept) { gsb += dL; gnb += dL * dL; }
        if (fitLinear)
          for (i64 jj = rb; jj < re; jj++) {
            i64 j = indices[jj];
            double val = data[jj];
            gsw[j] += dL * val;
            gnw[j] += (dL * val) * (dL * val);
          }
      }
      it += (end - start);
    }
    runningLoss /= (double)n;
    viol_out[ep] = viol;
    loss_out[ep] = runningLoss;
    epochs = ep + 1;
    if (isnan(runningLoss)) break;
    if (viol < tol) break;
  }
  {
    double itf = (double)(it - 1);
    double denom = eta0 * itf * beta;
    for (i64 q = 0; q < nP; q++) {
      P[q] = -eta0 * gsP[q];
      P[q] /= denom + sqrt(gnP[q]);
    }
    if (fitIntercept) {
      double den = sqrt(gnb) + eta0 * itf * alpha0;
      intercept = -eta0 * gsb / den;
    }
    if (fitLinear) {
      double den = eta0 * itf * alpha;
      for (i64 j = 0; j < d; j++) {
        w[j] = -eta0 * gsw[j];
        w[j] /= den + sqrt(gnw[j]);
      }
    }
  }
  *intercept_io = intercept;
  *it_io = it;
  *gsb_io = gsb; *gnb_io = gnb;
  free(dA); free(Pnew); free(wnew);
  return epochs;
}

/* SGD for FFM: optimizer/sgd_ffm.nim:33-106 reusing sgd.lazilyUpdate/update/finalize with order==field. */
int ref_ffm_sgd_fit(i64 n, i64 d, int nFields, int k, const double *data, const i64 *indices,
                    const i64 *indptr, const i64 *fields, const double *y, int fitLinear,
                    int fitIntercept, double *P, double *w, double *intercept_io, int loss_kind,
                    double thr, int maxIter, double eta0, double alpha0, double alpha, double beta,
                    int sched, double power, double tol, const i64 *perms, i64 *it_io,
                    double *viol_out, double *loss_out) {
  i64 nP = (i64)nFields * d * k;
  double *dA = (double *)calloc((size_t)(nP > 0 ? nP : 1), sizeof(double));
  double *scalings_w = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  double *scalings_P = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  for (i64 j = 0; j < d; j++) { scalings_w[j] = 1.0; scalings_P[j] = 1.0; }
  double scaling_w = 1.0, scaling_P = 1.0, intercept = *intercept_io;
  i64 it = *it_io;
  int epochs = 0;
  for (int ep = 0; ep < maxIter; ep++) {
    double viol = 0.0, runningLoss = 0.0;
    const i64 *perm = perms ? perms + (i64)ep * n : NULL;
    for (i64 q = 0; q < n; q++) {
      i64 i = perm ? perm[q] : q;
      i64 rb = indptr[i], re = indptr[i + 1];
      for (int f = 0; f < nFields; f++)
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = indices[jj];
          for (int s = 0; s < k; s++) P[((i64)f * d + j) * k + s] *= scaling_P / scalings_P[j];
        }
      if (fitLinear)
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = indices[jj];
          w[j] *= scaling_w / scalings_w[j];
        }
      double yPred = ref_ffm_predict_with_grad(d, nFields, k, data, indices, indptr, fields, i, P, w,
                                               intercept, dA);
      runningLoss += ref_loss(loss_kind, thr, y[i], yPred);
      double dL = ref_dloss(loss_kind, thr, y[i], yPred);
      double eta_w = get_eta(sched, eta0, power, alpha, it);
      double eta_P = get_eta(sched, eta0, power, beta, it);
      for (int f = 0; f < nFields; f++)
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = indices[jj];
          for (int s = 0; s < k; s++) {
            i64 e = ((i64)f * d + j) * k + s;
            double update = eta_P * (dL * dA[e] + beta * P[e]);
            viol += fabs(update);
            P[e] -= update;
          }
        }
      if (fitIntercept) {
        double update = get_eta(sched, eta0, power, alpha0, it) * (dL + alpha0 * intercept);
        viol += fabs(update);
        intercept -= update;
      }
      if (fitLinear)
        for (i64 jj = rb; jj < re; jj++) {
          i64 j = indices[jj];
          double update = eta_w * (dL * data[jj] + alpha * w[j]);
          w[j] -= update;
          viol += fabs(update);
        }
      scaling_P *= (1 - eta_P * beta);
      scaling_w *= (1 - eta_w * alpha);
      for (i64 jj = rb; jj < re; jj++) {
        i64 j = indices[jj];
        scalings_P[j] = scaling_P;
        scalings_w[j] = scaling_w;
      }
      if (fitLinear && scaling_w < 1e-9) {
        for (i64 j = 0; j < d; j++) { w[j] *= scaling_w; w[j] /= scalings_w[j]; scalings_w[j] = 1.0; }
        scaling_w = 1.0;
      }
      if (scaling_P < 1e-9) {
        for (int f = 0; f < nFields; f++)
          for (i64 j = 0; j < d; j++)
            for (int s = 0; s < k; s++) P[((i64)f * d + j) * k + s] *= scaling_P / scalings_P[j];
        for (i64 j = 0; j < d; j++) scalings_P[j] = 1.0;
        scaling_P = 1.0;
      }
      it++;
    }
    runningLoss /= (double)n;
    viol_out[ep] = viol;
    loss_out[ep] = runningLoss;
    epochs = ep + 1;
    if (isnan(runningLoss)) break;
    if (viol < tol) break;
  }
  if (fitLinear)
    for (i64 j = 0; j < d; j++) { w[j] *= scaling_w; w[j] /= scalings_w[j]; }
  for (int f = 0; f < nFields; f++)
    for (i64 j = 0; j < d; j++)
      for (int s = 0; s < k; s++) P[((i64)f * d + j) * k + s] *= scaling_P / scalings_P[j];
  *intercept_io = intercept;
  *it_io = it;
  free(dA); free(scalings_w); free(scalings_P);
  return epochs;
}

/* One FFM predict+grad pass over rows [rowBegin,rowEnd) (sgd_ffm.nim:11-30 + a minibatch-mean
 * gradient scatter shaped like minibatch_psgd.nim:67-88) for bench.py / parity of the FFM pair kernel. */
double ref_ffm_loss_grad(i64 d, int nFields, int k, const double *data, const i64 *indices,
                         const i64 *indptr, const i64 *fields, const double *y, i64 rowBegin,
                         i64 rowEnd, int fitLinear, int fitIntercept, const double *P,
                         const double *w, double intercept, int loss_kind, double thr,
                         i64 miniBatchSize, double *gP, double *gw, double *gb, double *yPredOut,
                         double *dA) {
  double lossSum = 0.0;
  for (i64 i = rowBegin; i < rowEnd; i++) {
    double yPred = ref_ffm_predict_with_grad(d, nFields, k, data, indices, indptr, fields, i, P, w,
                                             intercept, dA);
    if (yPredOut) yPredOut[i - rowBegin] = yPred;
    lossSum += ref_loss(loss_kind, thr, y[i], yPred);
    double coef = ref_dloss(loss_kind, thr, y[i], yPred) / (double)miniBatchSize;
    i64 rb = indptr[i], re = indptr[i + 1];
    for (int f = 0; f < nFields; f++)
      for (i64 jj = rb; jj < re; jj++)
        for (int s = 0; s < k; s++) {
          i64 e = ((i64)f * d + indices[jj]) * k + s;
          gP[e] += coef * dA[e];
        }
    if (fitLinear)
      for (i64 jj = rb; jj < re; jj++) gw[indices[jj]] += coef * data[jj];
    if (fitIntercept) *gb += coef;
  }
  return lossSum;
}

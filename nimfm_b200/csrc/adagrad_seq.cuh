// adagrad_seq.cuh -- AdaGrad with miniBatchSize = 1: the reference's strictly sequential per-sample loop
// (optimizer/adagrad.nim:87-134,164-181) in ONE persistent thread block, the row's P / g_sum / g_norm
// slices staged in shared memory (the minibatch pipeline costs six kernel launches per sample: 16 K
// samples/s on the C4 shape).  Per sample, in the reference's order:
//   update()   (it != 1): theta = -eta0*G/(eta0*(it-1)*reg + sqrt(N)) for the row's features (dummies
//                         included), the intercept and w (fitLinearAdaGrad, fit_linear.nim:50-57); viol
//   predictWithGrad, loss
//   updateG()  G += dL*dA, N += (dL*dA)^2 (P, w, intercept);  it += 1
#pragma once
#include "common.cuh"

#define ADASEQ_THREADS 256

struct AdaSeqArgs {
  const double *data;
  const int32_t *indices;
  const int64_t *indptr;
  const double *y;
  const int32_t *perm;
  int64_t nRows, d;
  int degree, k, nOrders, nAug, fitLinear, fitIntercept;
  double *P, *gsP, *gnP, *w, *gsw, *gnw, *b, *adaScal;   // adaScal = [g_sum.intercept, g_norm.intercept]
  double *scal;                                           // [loss sum, viol sum] (accumulated)
  int loss;
  double thr, eta0, alpha0, alpha, beta;
  int64_t it0;
  int zmax;
};

// dynamic smem: sP | sGs | sGn (zmax*SB8 each) | sX | sW (zmax each) | sJ (zmax int64)
static __global__ void __launch_bounds__(ADASEQ_THREADS, 1) adagrad_fm_seq_kernel(const AdaSeqArgs a) {
  extern __shared__ __align__(16) unsigned char ada_smem[];
  __shared__ double red[ADASEQ_THREADS / 32];
  __shared__ double sh[2];   // yhat, dL
  const int k = a.k, NO = a.nOrders, SB8 = NO * k, zmax = a.zmax;
  double *sP = reinterpret_cast<double *>(ada_smem);
  double *sGs = sP + (size_t)zmax * SB8;
  double *sGn = sGs + (size_t)zmax * SB8;
  double *sX = sGn + (size_t)zmax * SB8;
  double *sW = sX + zmax;
  int64_t *sJ = reinterpret_cast<int64_t *>(sW + zmax);
  const int tid = threadIdx.x, nth = blockDim.x;
  double viol = 0.0, lossAcc = 0.0;
  for (int64_t q = 0; q < a.nRows; ++q) {
    const int64_t i = a.perm ? (int64_t)a.perm[q] : q;
    const int64_t itq = a.it0 + q;
    const bool refresh = itq != 1;
    const double t = (double)(itq - 1);
    const int64_t rb = a.indptr[i];
    const int zReal = (int)(a.indptr[i + 1] - rb);
    const int z = zReal + a.nAug;
    __syncthreads();                                     // the previous sample's write-back has read the stage
    // ---- records; update() of w for the row's real features
    const double denW = t * a.eta0 * a.alpha;
    for (int u = tid; u < z; u += nth) {
      if (u < zReal) {
        const int64_t j = a.indices[rb + u];
        sJ[u] = j;
        sX[u] = a.data[rb + u];
        double wv = a.w[j];
        if (a.fitLinear && refresh) {
          const double wn = -a.eta0 * a.gsw[j] / (denW + sqrt(a.gnw[j]));
          viol += fabs(wv - wn);
          a.w[j] = wn;
          wv = wn;
        }
        sW[u] = wv;
      } else {
        sJ[u] = a.d + (u - zReal);
        sX[u] = 1.0;
        sW[u] = 0.0;
      }
    }
    if (tid == 0 && a.fitIntercept && refresh) {          // adagrad.nim:101-105
      const double old = a.b[0];
      const double den = sqrt(a.adaScal[1]) + a.eta0 * t * a.alpha0;
      const double nb = -a.eta0 * a.adaScal[0] / den;
      viol += fabs(old - nb);
      a.b[0] = nb;
    }
    __syncthreads();
    // ---- stage P / g_sum / g_norm of the row's features; update() of P (adagrad.nim:93-99)
    const double tmpP = a.eta0 * t * a.beta;
    for (int base = tid; base < z * SB8; base += 5 * nth) {   // all loads of a batch first: one memory latency per
      double gsv[5], gnv[5], pv[5];                             // batch instead of one per element
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const int e = base + r * nth;
        const bool ok = e < z * SB8;
        const int64_t ge = ok ? sJ[e / SB8] * SB8 + (e % SB8) : 0;
        gsv[r] = ok ? a.gsP[ge] : 0.0;
        gnv[r] = ok ? a.gnP[ge] : 1.0;
        pv[r] = ok ? a.P[ge] : 0.0;
      }
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const int e = base + r * nth;
        if (e < z * SB8) {
          double p = pv[r];
          if (refresh) {
            const double pn = -(a.eta0 * gsv[r]) / (tmpP + sqrt(gnv[r]));
            viol += fabs(p - pn);
            p = pn;
          }
          sP[e] = p;
          sGs[e] = gsv[r];
          sGn[e] = gnv[r];
        }
      }
    }
    __syncthreads();
    // ---- predictWithGrad forward (thread <-> (order, component)); A stays in registers
    double part = 0.0;
    for (int u = tid; u < zReal; u += nth) part += sW[u] * sX[u];
    AnovaState A;
    const int os = tid;
    const int o = os < SB8 ? os / k : 0, sc = os - o * k;
    const int M = a.degree - o;
    if (os < SB8) {
      part += anova_forward_smem(A, M, sP + o * k + sc, SB8, sX, z);
    }
    double yhat = block_sum(part, red);
    if (tid == 0) {
      yhat += a.b[0];
      sh[0] = yhat;
      const double yi = a.y[i];
      lossAcc += dev_loss(a.loss, a.thr, yi, yhat);
      sh[1] = dev_dloss(a.loss, a.thr, yi, yhat);
    }
    __syncthreads();
    const double dL = sh[1];
    // ---- updateG (adagrad.nim:113-134)
    if (os < SB8) {
#pragma unroll 4
      for (int u = 0; u < z; u++) {
        const double x = sX[u];
        const int e = u * SB8 + o * k + sc;
        const double p = sP[e];
        double g;
        if (M == 2) g = x * (A[1] - p * x);
        else {
          g = anova_deriv(A, M, x, p);
        }
        const double grad = dL * g;
        sGs[e] += grad;
        sGn[e] += grad * grad;
      }
    }
    if (a.fitLinear)
      for (int u = tid; u < zReal; u += nth) {
        const int64_t j = sJ[u];
        const double gx = dL * sX[u];
        a.gsw[j] += gx;
        a.gnw[j] += gx * gx;
      }
    if (tid == 0 && a.fitIntercept) {
      a.adaScal[0] += dL;
      a.adaScal[1] += dL * dL;
    }
    __syncthreads();
    for (int e = tid; e < z * SB8; e += nth) {
      const int u = e / SB8, off = e - u * SB8;
      const int64_t ge = sJ[u] * SB8 + off;
      if (refresh) a.P[ge] = sP[e];
      a.gsP[ge] = sGs[e];
      a.gnP[ge] = sGn[e];
    }
  }
  viol = block_sum(viol, red);
  __syncthreads();
  lossAcc = block_sum(lossAcc, red);
  if (tid == 0) {
    a.scal[0] += lossAcc;
    a.scal[1] += viol;
  }
}


// The same per-sample loop, PIPELINED like sgd_fm_pipe_kernel (sgd.cu): the read-only CSR side of the load chain
// (permutation -> row pointer / target -> indices / values) is fetched one to three samples ahead into registers by
// the last warp and parked in shared memory at the end of the iteration; w / g_sum_w / g_norm_w and the P / g_sum /
// g_norm rows -- everything the previous sample may have changed -- are one round of independent loads; the forward
// DP keeps the reference's order on SB8 threads and leaves A[.][1..M-1] in shared memory; updateG and the write-back
// are one pass over all threads (element <-> thread) straight to global memory.  rows of at most 64 nonzeros.
// dynamic smem: sP | sGs | sGn (zmax*SB8 each) | sA[SB8*(MAXDEG+1)] | sX[2][zmax] | sW[zmax] | sJ[2][zmax] (int32) | sOrd[SB8]
#define ADAPIPE_THREADS 512
struct AdaPipeMeta {
  int64_t i, rb;
  double y;
  int z;
};

static __global__ void __launch_bounds__(ADAPIPE_THREADS, 1) adagrad_fm_pipe_kernel(const AdaSeqArgs a) {
  extern __shared__ __align__(16) unsigned char ada_smem[];
  __shared__ double red[ADAPIPE_THREADS / 32];
  __shared__ double sh[2];            // dL
  __shared__ AdaPipeMeta meta[4];     // ring: sample q lives in meta[q & 3]
  const int k = a.k, NO = a.nOrders, SB8 = NO * k, zmax = a.zmax, AST = NIMFM_MAX_DEGREE + 1;
  double *sP = reinterpret_cast<double *>(ada_smem);
  double *sGs = sP + (size_t)zmax * SB8;
  double *sGn = sGs + (size_t)zmax * SB8;
  double *sA = sGn + (size_t)zmax * SB8;
  double *sXb = sA + (size_t)SB8 * AST;
  double *sW = sXb + 2 * (size_t)zmax;
  int32_t *sJb = reinterpret_cast<int32_t *>(sW + zmax);
  const int tid = threadIdx.x, nth = blockDim.x, lastWarp0 = nth - 32;
  double viol = 0.0, lossAcc = 0.0;
  ElemWalk walk0;
  walk0.start(tid, nth, SB8);
  signed char *sOrd = reinterpret_cast<signed char *>(sJb + 2 * zmax);   // [SB8] ANOVA order of (order, component) slot
  for (int os = tid; os < SB8; os += nth) sOrd[os] = (signed char)(a.degree - os / k);
  if (tid == 0) {
    for (int64_t q = 0; q < 3 && q < a.nRows; ++q) {
      AdaPipeMeta m;
      m.i = a.perm ? (int64_t)a.perm[q] : q;
      m.rb = 0; m.z = 0; m.y = 0.0;
      if (q < 2) {
        m.rb = a.indptr[m.i];
        m.z = (int)(a.indptr[m.i + 1] - m.rb);
        m.y = a.y[m.i];
      }
      meta[q & 3] = m;
    }
  }
  __syncthreads();
  if (tid >= lastWarp0 && a.nRows > 0) {
    const AdaPipeMeta m = meta[0];
    for (int u = tid - lastWarp0; u < m.z + a.nAug; u += 32) {
      sJb[u] = u < m.z ? a.indices[m.rb + u] : (int32_t)(a.d + (u - m.z));
      sXb[u] = u < m.z ? a.data[m.rb + u] : 1.0;
    }
  }
  __syncthreads();
  for (int64_t q = 0; q < a.nRows; ++q) {
    const int64_t itq = a.it0 + q;
    const bool refresh = itq != 1;
    const double t = (double)(itq - 1);
    const AdaPipeMeta m = meta[q & 3];
    const int zReal = m.z, z = zReal + a.nAug;
    const int32_t *J = sJb + (q & 1) * zmax;
    const double *X = sXb + (q & 1) * zmax;
    // ---- fetch ahead (registers now, shared memory at the end of the iteration)
    int64_t pfI = 0, pfRb = 0, pfRe = 0;
    double pfY = 0.0, pfX0 = 0.0, pfX1 = 0.0;
    int32_t pfJ0 = 0, pfJ1 = 0;
    if (tid == nth - 1) {
      if (q + 3 < a.nRows) pfI = a.perm ? (int64_t)a.perm[q + 3] : q + 3;
      if (q + 2 < a.nRows) {
        const int64_t i2 = meta[(q + 2) & 3].i;
        pfRb = a.indptr[i2];
        pfRe = a.indptr[i2 + 1];
        pfY = a.y[i2];
      }
    }
    const AdaPipeMeta m1 = meta[(q + 1) & 3];
    const int z1 = q + 1 < a.nRows ? m1.z + a.nAug : 0;
    if (tid >= lastWarp0) {
      const int u0 = tid - lastWarp0, u1 = u0 + 32;
      if (u0 < z1) {
        pfJ0 = u0 < m1.z ? a.indices[m1.rb + u0] : (int32_t)(a.d + (u0 - m1.z));
        pfX0 = u0 < m1.z ? a.data[m1.rb + u0] : 1.0;
      }
      if (u1 < z1) {
        pfJ1 = u1 < m1.z ? a.indices[m1.rb + u1] : (int32_t)(a.d + (u1 - m1.z));
        pfX1 = u1 < m1.z ? a.data[m1.rb + u1] : 1.0;
      }
    }
    // ---- update() of w for the row's real features (fitLinearAdaGrad, fit_linear.nim:50-57) and of the intercept
    const double denW = t * a.eta0 * a.alpha;
    if (tid < z) {
      double wv = 0.0;
      if (tid < zReal) {
        const int64_t j = J[tid];
        wv = a.w[j];
        if (a.fitLinear && refresh) {
          const double wn = -a.eta0 * a.gsw[j] / (denW + sqrt(a.gnw[j]));
          viol += fabs(wv - wn);
          a.w[j] = wn;
          wv = wn;
        }
      }
      sW[tid] = wv;
    }
    if (tid == 32 && a.fitIntercept && refresh) {          // adagrad.nim:101-105
      const double old = a.b[0];
      const double den = sqrt(a.adaScal[1]) + a.eta0 * t * a.alpha0;
      const double nb = -a.eta0 * a.adaScal[0] / den;
      viol += fabs(old - nb);
      a.b[0] = nb;
    }
    // ---- P / g_sum / g_norm of the row's features, update() of P (adagrad.nim:93-99): one batch of independent loads
    const double tmpP = a.eta0 * t * a.beta;
    ElemWalk wl = walk0;
    for (int base = tid; base < z * SB8; base += 5 * nth) {
      double gsv[5], gnv[5], pv[5];
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const bool ok = wl.u < z;
        const int64_t ge = ok ? (int64_t)J[wl.u] * SB8 + wl.off : 0;
        wl.next();
        gsv[r] = ok ? a.gsP[ge] : 0.0;
        gnv[r] = ok ? a.gnP[ge] : 1.0;
        pv[r] = ok ? a.P[ge] : 0.0;
      }
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const int e = base + r * nth;
        if (e < z * SB8) {
          double p = pv[r];
          if (refresh) {
            const double pn = -(a.eta0 * gsv[r]) / (tmpP + sqrt(gnv[r]));
            viol += fabs(p - pn);
            p = pn;
          }
          sP[e] = p;
          sGs[e] = gsv[r];
          sGn[e] = gnv[r];
        }
      }
    }
    __syncthreads();
    // ---- predictWithGrad forward (thread <-> (order, component), nonzeros in row order)
    double part = 0.0;
    for (int u = tid; u < zReal; u += nth) part += sW[u] * X[u];
    if (tid < SB8) {
      const int o = tid / k, sc = tid - o * k;
      const int M = a.degree - o;
      AnovaState A;
      part += anova_forward_smem(A, M, sP + o * k + sc, SB8, X, z);
#pragma unroll
      for (int tt = 1; tt < NIMFM_MAX_DEGREE; ++tt)
        if (tt < M) sA[tid * AST + tt] = A[tt];
    }
    const double yh = block_sum(part, red);
    if (tid == 0) {
      const double yhat = yh + a.b[0];
      lossAcc += dev_loss(a.loss, a.thr, m.y, yhat);
      sh[1] = dev_dloss(a.loss, a.thr, m.y, yhat);
    }
    __syncthreads();
    const double dL = sh[1];
    // ---- updateG (adagrad.nim:113-134) + write-back, element <-> thread
    ElemWalk wk = walk0;
#pragma unroll 2
    for (int e = tid; e < z * SB8; e += nth) {
      const int u = wk.u, os = wk.off;
      wk.next();
      const int M = sOrd[os];
      const double x = X[u], p = sP[e];
      const double *A = sA + os * AST;
      double g;
      if (M == 2) g = x * (A[1] - p * x);
      else {
        g = x;
        for (int tt = 1; tt < M; tt++) g = x * (A[tt] - p * g);
      }
      const double grad = dL * g;
      const int64_t ge = (int64_t)J[u] * SB8 + os;
      if (refresh) a.P[ge] = p;
      a.gsP[ge] = sGs[e] + grad;
      a.gnP[ge] = sGn[e] + grad * grad;
    }
    if (a.fitLinear && tid < zReal) {
      const int64_t j = J[tid];
      const double gx = dL * X[tid];
      a.gsw[j] += gx;
      a.gnw[j] += gx * gx;
    }
    if (tid == 0 && a.fitIntercept) {
      a.adaScal[0] += dL;
      a.adaScal[1] += dL * dL;
    }
    // ---- park what was fetched ahead
    if (tid == nth - 1) {
      if (q + 3 < a.nRows) meta[(q + 3) & 3].i = pfI;
      if (q + 2 < a.nRows) {
        AdaPipeMeta &m2 = meta[(q + 2) & 3];
        m2.rb = pfRb;
        m2.z = (int)(pfRe - pfRb);
        m2.y = pfY;
      }
    }
    if (tid >= lastWarp0) {
      int32_t *Jn = sJb + ((q + 1) & 1) * zmax;
      double *Xn = sXb + ((q + 1) & 1) * zmax;
      const int u0 = tid - lastWarp0, u1 = u0 + 32;
      if (u0 < z1) { Jn[u0] = pfJ0; Xn[u0] = pfX0; }
      if (u1 < z1) { Jn[u1] = pfJ1; Xn[u1] = pfX1; }
    }
    __syncthreads();
  }
  viol = block_sum(viol, red);
  __syncthreads();
  lossAcc = block_sum(lossAcc, red);
  if (tid == 0) {
    a.scal[0] += lossAcc;
    a.scal[1] += viol;
  }
}

// ffm_cols.cuh -- FFM predict+grad WITHOUT atomics (included by ffm.cu): forward pair kernel, then a column kernel
// over the CSC twin of the field dataset.
//
// sgd_ffm.nim:23-30 adds, for every pair (u, v) of a row,  dA[f_v][j_u][:] += x_u x_v P[j_v][f_u][:]  (and the
// mirror image).  Read column-wise: the gradient vector of (feature j, field f) is
//     gP[j][f][:] = sum over rows i containing j of  coef_i x_ij  *  sum over the row's nonzeros u of field f,
//                   j_u != j, of  x_iu P[j_u][f_j][:]
// so a block that owns feature j walks j's column entries in ascending row order, thread (f, s) finds the row's
// nonzero of field f through a small position table, gathers ONE k*8-byte vector P[j_u][f_j][:] per (entry,
// field) -- the same bytes the forward pass gathers -- accumulates in a register and finishes with one plain store
// per gradient element (long columns: fixed 512-entry segments, partial sums added in segment order).  Against the
// RED route (ffm_pairs.cuh, 11 856 lane-REDs per 39-field row, each a read-modify-write of a cold line in DRAM)
// this trades 95 KB/row of REDs for a second 95 KB/row of gathers, and every sum has one order: deterministic.
// Needs what the pair route's AdaGrad mode needs -- no row with two nonzeros of one field (checked once per dataset
// on the device) -- plus rows of at most 64 nonzeros and a contiguous row range.
#pragma once

struct FfmColArgs {
  const double *cdata;       // CSC twin
  const int32_t *crow;
  const int32_t *taskCol, *taskLen, *taskSlot;
  const int64_t *taskBeg;
  int64_t nTasks;
  const double *data;        // the CSR itself
  const int32_t *indices, *fields;
  const int64_t *indptr;
  const double *coef;        // [rowEnd - rowBegin] coef_i = dloss_i / mb (written by ffm_coef_kernel)
  int64_t rowBegin, rowEnd;
  int nFields, CH;
  const double *P;
  double *gP, *gw, *partial;
  int fitLinear;
};

static __device__ __forceinline__ int64_t lower_bound_row_i32(const int32_t *rows, int64_t lo, int64_t hi, int64_t r) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)rows[mid] < r) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// columns cut into several segments: their partial sums are added in segment order (one block per column)
static __global__ void ffm_cols_combine_kernel(const int32_t *multiCol, const int32_t *multiFirst, const int32_t *multiCount,
                                               int64_t nMulti, const double *partial, int SB8, double *gP, double *gw,
                                               int fitLinear) {
  const int64_t c = blockIdx.x;
  if (c >= nMulti) return;
  const int64_t j = multiCol[c];
  for (int e = threadIdx.x; e <= SB8; e += blockDim.x) {
    double s = 0.0;
    const double *ps = partial + (size_t)multiFirst[c] * (SB8 + 1) + e;
    for (int q = 0; q < multiCount[c]; ++q) s += ps[(size_t)q * (SB8 + 1)];
    if (e < SB8) gP[j * SB8 + e] += s;
    else if (fitLinear) gw[j] += s;
  }
}

// coef_i = dloss(y_i, yhat_i) / mb in place of yhat (the forward kernel's output), loss / sum coef / sum dL^2 partials
static __global__ void ffm_coef_kernel(double *yhatCoef, const double *y, int64_t rowBegin, int64_t nRows, int loss,
                                       double thr, double mb, double *partials) {
  __shared__ double red[8];
  double l = 0.0, c = 0.0, d2 = 0.0;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nRows; q += (int64_t)gridDim.x * blockDim.x) {
    const double yi = y[rowBegin + q], yh = yhatCoef[q];
    const double dL = dev_dloss(loss, thr, yi, yh);
    l += dev_loss(loss, thr, yi, yh);
    c += dL / mb;
    d2 += dL * dL;
    yhatCoef[q] = dL / mb;
  }
  l = block_sum(l, red);
  __syncthreads();
  c = block_sum(c, red);
  __syncthreads();
  d2 = block_sum(d2, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 4 + 0] = l;
    partials[blockIdx.x * 4 + 1] = c;
    partials[blockIdx.x * 4 + 2] = d2;
    partials[blockIdx.x * 4 + 3] = 0.0;
  }
}

struct FfmColMeta {
  int64_t rb;
  double cx;     // coef_i * x_ij
  int z, fj;     // row length; field of feature j in this row
};

template <int KT, int E>
__global__ void __launch_bounds__(512) ffm_cols_grad_kernel(const FfmColArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NS = blockDim.x / KT, CH = a.CH, nF = a.nFields;
  FfmRec *recs = reinterpret_cast<FfmRec *>(smem_raw);                           // [E][CH]
  FfmColMeta *meta = reinterpret_cast<FfmColMeta *>(recs + (size_t)E * CH);     // [E]
  signed char *pos = reinterpret_cast<signed char *>(meta + E);                  // [E][NS]: position of field f in row e
  const int tid = threadIdx.x, s = tid % KT, f = tid / KT;
  const int SB8 = nF * KT;
  for (int64_t t = blockIdx.x; t < a.nTasks; t += gridDim.x) {
    const int64_t j = a.taskCol[t];
    const int32_t jb = (int32_t)(j * nF);
    int64_t eb = a.taskBeg[t], ee = eb + a.taskLen[t];
    if ((int64_t)a.crow[eb] < a.rowBegin) eb = lower_bound_row_i32(a.crow, eb, ee, a.rowBegin);
    if (eb < ee && (int64_t)a.crow[ee - 1] >= a.rowEnd) ee = lower_bound_row_i32(a.crow, eb, ee, a.rowEnd);
    double acc = 0.0, accW = 0.0;
    for (int64_t e0 = eb; e0 < ee; e0 += E) {
      const int nb = (int)(ee - e0 < E ? ee - e0 : E);
      if (tid < nb) {
        const int64_t e = e0 + tid, row = a.crow[e];
        FfmColMeta m;
        m.rb = a.indptr[row];
        m.z = (int)(a.indptr[row + 1] - m.rb);
        m.cx = a.coef[row - a.rowBegin] * a.cdata[e];
        m.fj = 0;
        meta[tid] = m;
      }
      for (int q = tid; q < E * NS; q += blockDim.x) pos[q] = -1;
      __syncthreads();
      for (int q = tid; q < nb * CH; q += blockDim.x) {
        const int e = q / CH, u = q - e * CH;
        if (u < meta[e].z) {
          const int64_t at = meta[e].rb + u;
          FfmRec r;
          const int32_t idx = a.indices[at];
          r.jb = idx * nF;
          r.f = a.fields[at];
          r.x = a.data[at];
          recs[e * CH + u] = r;
          pos[e * NS + r.f] = (signed char)u;
          if (idx == (int32_t)j) meta[e].fj = r.f;
        }
      }
      __syncthreads();
      if (f < nF) {
        double v[E], sc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
          v[e] = 0.0;
          sc[e] = 0.0;
          if (e < nb) {
            const int u = pos[e * NS + f];
            if (u >= 0) {
              const FfmRec r = recs[e * CH + u];
              if (r.jb != jb) {   // the reference pairs distinct features only
                v[e] = __ldg(a.P + (int64_t)(r.jb + meta[e].fj) * KT + s);   // P[j_u][f_j][s]
                sc[e] = meta[e].cx * r.x;
              }
            }
          }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) acc += sc[e] * v[e];   // entries in ascending row order
      }
      if (tid == 0)
        for (int e = 0; e < nb; ++e) accW += meta[e].cx;
      __syncthreads();
    }
    const int slot = a.taskSlot[t];
    if (slot < 0) {
      if (f < nF) a.gP[(int64_t)(jb + f) * KT + s] += acc;
      if (tid == 0 && a.fitLinear) a.gw[j] += accW;
    } else {
      double *ps = a.partial + (size_t)slot * (SB8 + 1);
      if (f < nF) ps[f * KT + s] = acc;
      if (tid == 0) ps[SB8] = accW;
    }
  }
}

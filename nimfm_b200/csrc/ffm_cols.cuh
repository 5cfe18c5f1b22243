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
  int64_t d;
  const double *PT;          // the parameters in FIELD-major layout PT[f][j][s] (see ffm_field_major_kernel)
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

// P[j][f][s] (the row kernels' layout: one feature's fields contiguous) -> PT[f][j][s].  The column kernel gathers,
// for a column of field f_j, the vectors P[j_u][f_j][:] of the partner features: in PT they all lie in ONE slab of
// d*k doubles (64 MB for C5) that stays in L2 while the columns of that field are processed, where the row layout
// scatters them over all of P (2.5 GB) -- 64-byte random DRAM accesses, measured at 21 % of the DRAM peak.
static __global__ void ffm_field_major_kernel(const double *P, double *PT, int64_t d, int nF, int k) {
  const int64_t total = d * nF * k;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(e % k);
    const int64_t t = e / k;
    const int64_t j = t % d;
    const int f = (int)(t / d);
    PT[e] = P[(j * nF + f) * k + s];
  }
}

// bytes of shared scratch one warp needs for E entries in flight (must match WarpScratch below)
static inline size_t ffm_cols_warp_smem(int E) {
  const size_t b = (size_t)E * 64 * 8 + (size_t)E * 64 * 4 + (size_t)E * 64 + (size_t)E * 4;
  return (b + 15) & ~(size_t)15;
}

// ONE WARP PER COLUMN SEGMENT, lane <-> field (two passes for nFields > 32), every lane owns the whole k-vector of
// its field in registers.  Per batch of E column entries: lanes < E fetch the entries' row / row pointer / coef*x;
// each entry's row is read with coalesced loads (lane <-> position) into the warp's shared scratch together with
// the inverse table field -> position; then lane f gathers P[j_u][f_j][0..k) for all E entries -- E*k/2 16-byte
// loads in flight per lane -- and accumulates.  No block-wide barrier anywhere: warps run independent segments.
// (The first form -- a block per segment, thread <-> (field, component), three __syncthreads per batch -- ran the
// C5 shape at 21 M rows/s against 60 M for the forward kernel gathering the same bytes: barrier- and latency-bound.)
template <int KT, int E>
__global__ void __launch_bounds__(256) ffm_cols_grad_kernel(const FfmColArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int MAXZ = 64;
  struct WarpScratch {
    double x[E][MAXZ];
    int32_t jb[E][MAXZ];
    signed char inv[E][MAXZ];   // field -> position in the row, -1: the row has no nonzero of that field
    int fj[E];
  };
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpScratch &ws = reinterpret_cast<WarpScratch *>(smem_raw)[wib];
  const int nF = a.nFields;
  const int SB8 = nF * KT;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = warp; t < a.nTasks; t += nWarps) {
    const int64_t j = a.taskCol[t];
    const int32_t jb = (int32_t)(j * nF);
    int64_t eb = a.taskBeg[t], ee = eb + a.taskLen[t];
    if ((int64_t)a.crow[eb] < a.rowBegin) eb = lower_bound_row_i32(a.crow, eb, ee, a.rowBegin);
    if (eb < ee && (int64_t)a.crow[ee - 1] >= a.rowEnd) ee = lower_bound_row_i32(a.crow, eb, ee, a.rowEnd);
    double acc0[KT], acc1[KT], accW = 0.0;
#pragma unroll
    for (int c = 0; c < KT; ++c) acc0[c] = acc1[c] = 0.0;
    for (int64_t e0 = eb; e0 < ee; e0 += E) {
      const int nb = (int)(ee - e0 < E ? ee - e0 : E);
      int64_t rbL = 0;
      int zL = 0;
      double cxL = 0.0;
      if (lane < nb) {
        const int64_t e = e0 + lane, row = a.crow[e];
        rbL = a.indptr[row];
        zL = (int)(a.indptr[row + 1] - rbL);
        cxL = a.coef[row - a.rowBegin] * a.cdata[e];
      }
      __syncwarp();   // the previous batch's readers are done with the scratch
#pragma unroll
      for (int i = 0; i < E; ++i) {
        ws.inv[i][lane] = -1;
        ws.inv[i][lane + 32] = -1;
      }
      __syncwarp();
      double cx[E];
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const int64_t rb = __shfl_sync(0xffffffffu, rbL, i);
        const int z = __shfl_sync(0xffffffffu, zL, i);
        cx[i] = __shfl_sync(0xffffffffu, cxL, i);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int u = lane + 32 * half;
          if (i < nb && u < z) {
            const int32_t idx = a.indices[rb + u];
            const int fld = a.fields[rb + u];
            ws.x[i][u] = a.data[rb + u];
            ws.jb[i][u] = idx;
            ws.inv[i][fld] = (signed char)u;
            if (idx == (int32_t)j) ws.fj[i] = fld;
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int f = lane + 32 * pass;
        if (pass == 1 && nF <= 32) break;
        double v[E][KT], sc[E];
#pragma unroll
        for (int i = 0; i < E; ++i) {
          sc[i] = 0.0;
#pragma unroll
          for (int c = 0; c < KT; ++c) v[i][c] = 0.0;
          if (i < nb && f < nF) {
            const int u = ws.inv[i][f];
            if (u >= 0 && ws.jb[i][u] != (int32_t)j) {   // the reference pairs distinct features only
              const double2 *src = reinterpret_cast<const double2 *>(a.PT + ((int64_t)ws.fj[i] * a.d + ws.jb[i][u]) * KT);   // P[j_u][f_j][:]
#pragma unroll
              for (int c = 0; c < KT / 2; ++c) {
                const double2 q = __ldg(src + c);
                v[i][2 * c] = q.x;
                v[i][2 * c + 1] = q.y;
              }
              sc[i] = cx[i] * ws.x[i][u];
            }
          }
        }
#pragma unroll
        for (int i = 0; i < E; ++i)   // entries in ascending row order
#pragma unroll
          for (int c = 0; c < KT; ++c) {
            if (pass == 0) acc0[c] += sc[i] * v[i][c];
            else acc1[c] += sc[i] * v[i][c];
          }
      }
      if (lane == 0)
#pragma unroll
        for (int i = 0; i < E; ++i)
          if (i < nb) accW += cx[i];
    }
    const int slot = a.taskSlot[t];
    double *dst = slot < 0 ? a.gP + (int64_t)jb * KT : a.partial + (size_t)slot * (SB8 + 1);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int f = lane + 32 * pass;
      if (f < nF) {
#pragma unroll
        for (int c = 0; c < KT; ++c) {
          const double val = pass == 0 ? acc0[c] : acc1[c];
          if (slot < 0) dst[f * KT + c] += val;
          else dst[f * KT + c] = val;
        }
      }
    }
    if (lane == 0) {
      if (slot < 0) { if (a.fitLinear) a.gw[j] += accW; }
      else dst[SB8] = accW;
    }
  }
}

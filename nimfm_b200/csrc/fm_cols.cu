// fm_cols.cu -- the DETERMINISTIC predict+grad route (NIMFM_DETERMINISTIC=1): no atomics anywhere.
//
// The reference accumulates a minibatch's gradient sample by sample (minibatch_psgd.nim:77-85), so its sums have
// ONE order; the default device route (fm_rows_stream.cuh) scatters with FP64 REDs, whose order changes from
// run to run (~1e-16 relative noise).  This route fixes the order instead:
//   pass 1  the row kernel in MODE_STASH: forward + loss derivative; each row leaves a stash record
//           [coef | A[o][1 .. M_o-1][s]] -- everything the derivative recurrence of sgd.nim:176-188 needs from the row;
//   pass 2  a COLUMN kernel over the CSC twin of the dataset (stable transpose: rows ascending inside a column):
//           for feature j, the entries of column j inside the row range are walked in row order, lane <-> component,
//           the parameter row P[j][o][s] sits in registers, every entry contributes coef_i * dA_i(j, o, s) to a register
//           accumulator, and the sum is written with ONE plain store per gradient element.  Long columns are cut
//           into segments of SEG entries whose partial sums a second kernel adds in segment order.
// Every gradient element is therefore a sum in ascending row order over fixed segments: bit-identical from run
// to run and independent of the grid.  Traffic per nonzero: 12 B of CSC + the row's stash record (M_o-1 values per
// component: 128 B for C3, 768 B for C4) instead of a re-gather of P and a RED; no read-for-ownership of cold
// gradient lines.  Supported where the streaming row kernel is (degree 2 / 3, k in {8,16,32}) for a contiguous
// row range without wrap-around.
#include <stdlib.h>

#include <algorithm>

#include "dense_kernels.cuh"
#include "fm_rows_stream.cuh"

typedef void (*RowKernel)(const RowArgs);

namespace {

constexpr int SEG = 512;   // entries per column segment

struct ColArgs {
  const double *cdata;      // CSC values
  const int32_t *crow;      // CSC row ids (ascending inside a column)
  const int32_t *taskCol;   // [nTasks] feature id (>= d: dummy feature d + a)
  const int64_t *taskBeg;   // [nTasks] first entry of the segment (dummy features: first row)
  const int32_t *taskLen;   // [nTasks]
  const int32_t *taskSlot;  // [nTasks] partial-sum slot, -1: the segment is the whole column
  int64_t nTasks;
  int64_t d, rowBegin, rowEnd;
  const double *P;
  const double *stash;
  int stashStride;
  double *gP, *gw;
  double *partial;          // [nSlots][SB8 + 1]
  int fitLinear;
};

// first position in [lo, hi) whose row id is >= r
__device__ __forceinline__ int64_t lower_bound_row(const int32_t *rows, int64_t lo, int64_t hi, int64_t r) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)rows[mid] < r) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <int DEGREE, bool EXPLICIT, int KT>
__global__ void __launch_bounds__(256) fm_cols_grad_kernel(const ColArgs a) {
  constexpr int NO = RowCfg<DEGREE, EXPLICIT>::NO;
  constexpr int G = KT, GPW = 32 / G, SB8 = NO * KT, U = 4;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gid = lane / G;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t nGroups = (int64_t)gridDim.x * blockDim.x / G;
  (void)gid;
  (void)GPW;
  for (int64_t t = group; t < a.nTasks; t += nGroups) {
    const int64_t j = a.taskCol[t];
    const bool dummy = j >= a.d;
    int64_t eb = a.taskBeg[t], ee = eb + a.taskLen[t];
    if (!dummy) {   // clip the segment to the row range (rows ascend inside a column)
      if ((int64_t)a.crow[eb] < a.rowBegin) eb = lower_bound_row(a.crow, eb, ee, a.rowBegin);
      if (eb < ee && (int64_t)a.crow[ee - 1] >= a.rowEnd) ee = lower_bound_row(a.crow, eb, ee, a.rowEnd);
    } else {
      eb = eb < a.rowBegin ? a.rowBegin : eb;
      ee = ee > a.rowEnd ? a.rowEnd : ee;
    }
    double p[NO], acc[NO], accW = 0.0;
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      p[o] = a.P[j * SB8 + o * KT + gl];
      acc[o] = 0.0;
    }
    // software pipeline: the row ids / values of batch b+1 are fetched while batch b's stash records are in flight,
    // so a batch costs ONE dependent memory latency (stash gather) instead of two (ncu on the unpipelined form:
    // long_scoreboard 21.6 warps per issue, DRAM at 38 % of peak)
    int64_t rowN[U];
    double xN[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = eb + u;
      const bool ok = e < ee;
      rowN[u] = ok ? (dummy ? e : (int64_t)a.crow[e]) : a.rowBegin;
      xN[u] = ok ? (dummy ? 1.0 : a.cdata[e]) : 0.0;
    }
    for (int64_t e0 = eb; e0 < ee; e0 += U) {
      double x[U], coef[U], av[U][NO][DEGREE];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = e0 + u < ee;
        x[u] = xN[u];
        const double *rec = a.stash + (size_t)(rowN[u] - a.rowBegin) * a.stashStride;
        coef[u] = ok ? rec[0] : 0.0;
        int off = 1;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          const int M = DEGREE - o;
#pragma unroll
          for (int tt = 1; tt < DEGREE; ++tt)
            if (tt < M) {
              av[u][o][tt] = ok ? rec[off + gl] : 0.0;
              off += KT;
            }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {   // next batch's entries
        const int64_t e = e0 + U + u;
        const bool ok = e < ee;
        rowN[u] = ok ? (dummy ? e : (int64_t)a.crow[e]) : a.rowBegin;
        xN[u] = ok ? (dummy ? 1.0 : a.cdata[e]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {   // entries in ascending row order: the sum has one order
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          const int M = DEGREE - o;
          double g;
          if (M == 2) {
            g = x[u] * (av[u][o][1] - p[o] * x[u]);
          } else {
            g = x[u];
#pragma unroll
            for (int tt = 1; tt < DEGREE; ++tt)
              if (tt < M) g = x[u] * (av[u][o][tt] - p[o] * g);
          }
          acc[o] += coef[u] * g;
        }
        accW += coef[u] * x[u];
      }
    }
    const int slot = a.taskSlot[t];
    if (slot < 0) {
#pragma unroll
      for (int o = 0; o < NO; ++o) a.gP[j * SB8 + o * KT + gl] += acc[o];
      if (gl == 0 && !dummy && a.fitLinear) a.gw[j] += accW;
    } else {
      double *ps = a.partial + (size_t)slot * (SB8 + 1);
#pragma unroll
      for (int o = 0; o < NO; ++o) ps[o * KT + gl] = acc[o];
      if (gl == 0) ps[SB8] = accW;
    }
  }
}

// columns cut into several segments: their partial sums are added in segment order
__global__ void fm_cols_combine_kernel(const int32_t *multiCol, const int32_t *multiFirst, const int32_t *multiCount,
                                       int64_t nMulti, const double *partial, int SB8, int64_t d, double *gP, double *gw,
                                       int fitLinear) {
  const int64_t c = blockIdx.x;
  if (c >= nMulti) return;
  const int64_t j = multiCol[c];
  for (int e = threadIdx.x; e <= SB8; e += blockDim.x) {
    double s = 0.0;
    const double *ps = partial + (size_t)multiFirst[c] * (SB8 + 1) + e;
    for (int q = 0; q < multiCount[c]; ++q) s += ps[(size_t)q * (SB8 + 1)];
    if (e < SB8) gP[j * SB8 + e] += s;
    else if (j < d && fitLinear) gw[j] += s;
  }
}

template <int DEGREE, bool EXPLICIT>
RowKernel pick_stash(int k) {
  switch (k) {
    case 8: return fm_rows_stream_kernel<DEGREE, EXPLICIT, MODE_STASH, 8>;
    case 16: return fm_rows_stream_kernel<DEGREE, EXPLICIT, MODE_STASH, 16>;
    case 32: return fm_rows_stream_kernel<DEGREE, EXPLICIT, MODE_STASH, 32>;
    default: return nullptr;
  }
}

typedef void (*ColKernel)(const ColArgs);
template <int DEGREE, bool EXPLICIT>
ColKernel pick_cols(int k) {
  switch (k) {
    case 8: return fm_cols_grad_kernel<DEGREE, EXPLICIT, 8>;
    case 16: return fm_cols_grad_kernel<DEGREE, EXPLICIT, 16>;
    case 32: return fm_cols_grad_kernel<DEGREE, EXPLICIT, 32>;
    default: return nullptr;
  }
}

}  // namespace

RowKernel nimfm_row_stream_kernel_stash(int degree, bool explicitLower, int k) {
  switch (degree) {
    case 2: return pick_stash<2, false>(k);
    case 3: return explicitLower ? pick_stash<3, true>(k) : pick_stash<3, false>(k);
    default: return nullptr;
  }
}

void nimfm_det_twin_free(nimfm_ctx *ctx, nimfm_det_twin *t) {
  if (!t) return;
  nimfm_dataset_free(ctx, t->csc);
  cudaFree(t->taskCol); cudaFree(t->taskLen); cudaFree(t->taskSlot); cudaFree(t->taskBeg);
  cudaFree(t->multiCol); cudaFree(t->multiFirst); cudaFree(t->multiCount);
  delete t;
}

// The column twin of a CSR dataset and the segment list of its columns, built once per dataset (bookkeeping).
static int build_twin(nimfm_ctx *ctx, const nimfm_dataset *X, int nAug, nimfm_det_twin **out) {
  nimfm_det_twin *t = new nimfm_det_twin();
  struct Guard { nimfm_ctx *c; nimfm_det_twin *t; ~Guard() { if (t) nimfm_det_twin_free(c, t); } } guard{ctx, t};
  nimfm_dataset view = *X;      // a plain-CSR view of the same device arrays (a field dataset's fields play no part here)
  view.kind = NIMFM_DS_CSR;
  view.fields = nullptr;
  view.y = nullptr;
  view.detTwin = nullptr;
  int rc = nimfm_dataset_transpose(ctx, &view, &t->csc);
  if (rc) return rc;
  const int64_t d = X->d, n = X->n;
  std::vector<int64_t> ptr((size_t)d + 1);
  CK(cudaMemcpy(ptr.data(), t->csc->indptr, ((size_t)d + 1) * 8, cudaMemcpyDeviceToHost));
  std::vector<int32_t> col, len, slot, mcol, mfirst, mcount;
  std::vector<int64_t> beg;
  int64_t nSlots = 0;
  auto add_column = [&](int64_t j, int64_t b, int64_t e) {
    const int64_t L = e - b;
    if (L <= 0) return;
    const int64_t nseg = (L + SEG - 1) / SEG;
    if (nseg > 1) {
      mcol.push_back((int32_t)j);
      mfirst.push_back((int32_t)nSlots);
      mcount.push_back((int32_t)nseg);
    }
    for (int64_t s = 0; s < nseg; s++) {
      col.push_back((int32_t)j);
      beg.push_back(b + s * SEG);
      len.push_back((int32_t)std::min<int64_t>(SEG, e - (b + s * SEG)));
      slot.push_back(nseg > 1 ? (int32_t)(nSlots + s) : -1);
    }
    if (nseg > 1) nSlots += nseg;
  };
  for (int64_t j = 0; j < d; j++) add_column(j, ptr[(size_t)j], ptr[(size_t)j + 1]);
  for (int a = 0; a < nAug; a++) add_column(d + a, 0, n);   // dummy features: every row, value 1 (dataset.nim:182-189)
  if (nSlots >= (int64_t)1 << 31) return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "too many column segments");
  t->nTasks = (int64_t)col.size();
  t->nMulti = (int64_t)mcol.size();
  t->nSlots = nSlots;
  t->nAug = nAug;
  auto up = [&](auto **dst, const auto &v) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(v[0]));
    if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice);
    return e;
  };
  CK(up(&t->taskCol, col));
  CK(up(&t->taskLen, len));
  CK(up(&t->taskSlot, slot));
  CK(up(&t->taskBeg, beg));
  CK(up(&t->multiCol, mcol));
  CK(up(&t->multiFirst, mfirst));
  CK(up(&t->multiCount, mcount));
  guard.t = nullptr;
  *out = t;
  return NIMFM_OK;
}

int nimfm_det_twin_get(nimfm_ctx *ctx, const nimfm_dataset *X, int nAug, nimfm_det_twin **out) {
  if (!X->detTwin || X->detTwin->nAug != nAug) {
    if (X->detTwin) nimfm_det_twin_free(ctx, X->detTwin);
    X->detTwin = nullptr;
    int rc = build_twin(ctx, X, nAug, &X->detTwin);
    if (rc) return rc;
  }
  *out = X->detTwin;
  return NIMFM_OK;
}

// Deterministic predict+grad over rows [rowBegin, rowBegin + nRows) of X into the model's gradient buffers
// (accumulating, like the RED route).  red4 (loss, sum coef, sum dL^2, -) is left at ctx->scalars + 8.
int nimfm_fm_det_loss_grad(nimfm_ctx *ctx, nimfm_fm *fm, const nimfm_dataset *X, RowKernel stashKern, int G, int CH,
                           int grid, int block, size_t smem, int64_t nWarps, int loss, double thr, int64_t rowBegin,
                           int64_t nRows, double mb, double *yOutDev, const RowArgs &base) {
  if (rowBegin < 0 || rowBegin + nRows > X->n)
    return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "the deterministic gradient needs a contiguous row range without wrap-around");
  const bool expl = fm->degree > 2 && fm->nOrders == fm->degree - 1;
  ColKernel ck = fm->degree == 2 ? pick_cols<2, false>(fm->k)
                 : fm->degree == 3 ? (expl ? pick_cols<3, true>(fm->k) : pick_cols<3, false>(fm->k)) : nullptr;
  if (!ck || !stashKern)
    return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "the deterministic gradient supports degree 2 / 3 with nComponents in {8,16,32}");
  int rc;
  nimfm_det_twin *tw = nullptr;
  if ((rc = nimfm_det_twin_get(ctx, X, fm->nAug, &tw))) return rc;
  // stash record: coef + (M_o - 1) values per component and order, padded to an even number of doubles
  int vals = 0;
  for (int o = 0; o < fm->nOrders; o++) vals += (fm->degree - o) - 1;
  int stride = 1 + vals * fm->k;
  stride += stride & 1;
  const size_t need = (size_t)std::max<int64_t>(nRows, 1) * stride + (size_t)tw->nSlots * (fm->nOrders * fm->k + 1) + 16;
  if (ctx->stashCap < need) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->stash) CK(cudaFree(ctx->stash));
    ctx->stash = nullptr;
    ctx->stashCap = 0;
    CK(cudaMalloc(&ctx->stash, need * 8));
    ctx->stashCap = need;
  }
  if ((rc = nimfm_ensure_partials(ctx, (size_t)nWarps * 4))) return rc;
  RowArgs a = base;
  a.rowBegin = rowBegin;
  a.nRows = nRows;
  a.rowIdx = nullptr;
  a.yOut = yOutDev;
  a.partials = ctx->partials;
  a.loss = loss;
  a.thr = thr;
  a.mb = mb;
  a.G = G;
  a.CH = CH;
  a.stash = ctx->stash;
  a.stashStride = stride;
  stashKern<<<grid, block, smem, ctx->stream>>>(a);
  LAUNCHED(ctx);
  reduce_partials_kernel<<<1, 256, 0, ctx->stream>>>(ctx->partials, nWarps, ctx->scalars + 8, 0);
  LAUNCHED(ctx);
  ColArgs c;
  c.cdata = tw->csc->data;
  c.crow = tw->csc->indices;
  c.taskCol = tw->taskCol;
  c.taskBeg = tw->taskBeg;
  c.taskLen = tw->taskLen;
  c.taskSlot = tw->taskSlot;
  c.nTasks = tw->nTasks;
  c.d = fm->d;
  c.rowBegin = rowBegin;
  c.rowEnd = rowBegin + nRows;
  c.P = fm->P;
  c.stash = ctx->stash;
  c.stashStride = stride;
  c.gP = fm->grad;
  c.gw = fm->grad + fm->nP();
  c.partial = ctx->stash + (size_t)std::max<int64_t>(nRows, 1) * stride;
  c.fitLinear = fm->fitLinear;
  if (tw->nTasks > 0) {
    const int64_t groupsPerBlock = 256 / fm->k;
    const int cgrid = (int)std::max<int64_t>(1, std::min<int64_t>((tw->nTasks + groupsPerBlock - 1) / groupsPerBlock,
                                                                  (int64_t)ctx->numSMs * 8));
    ck<<<cgrid, 256, 0, ctx->stream>>>(c);
    LAUNCHED(ctx);
  }
  if (tw->nMulti > 0) {
    fm_cols_combine_kernel<<<(unsigned)tw->nMulti, 64, 0, ctx->stream>>>(tw->multiCol, tw->multiFirst, tw->multiCount,
                                                                         tw->nMulti, c.partial, fm->nOrders * fm->k, fm->d,
                                                                         c.gP, c.gw, fm->fitLinear);
    LAUNCHED(ctx);
  }
  CK(cudaGetLastError());
  return NIMFM_OK;
}

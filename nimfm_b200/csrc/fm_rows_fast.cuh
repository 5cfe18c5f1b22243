// fm_rows_fast.cuh -- the tuned instance of the row kernel for the shapes BASELINE.json names:
// nComponents KT in {8, 16, 32} fixed at compile time (one lane per component, G = KT lanes per row,
// 32/KT rows per warp), degree 2 or 3, every row fitting the shared-memory staging buffer (z <= CH).
// Same arithmetic and the same shared-memory staging idea as fm_rows.cuh (which stays the generic
// kernel for any k / degree / row length and for AdaGrad); differences are purely mechanical:
//   * per-nonzero metadata {x, j, accumulator offset} is packed in one 16-byte shared-memory record,
//     so the inner loops issue one LDS.128 + NO LDS.64 per nonzero and no integer address math;
//   * all strides are compile-time constants and the loops are unrolled for ILP (the FP64 recurrences
//     are latency-bound at the 8-10 resident warps/SM the staging footprint allows);
//   * the w-gradient of hot columns goes through a shared atomicAdd outside the main loop.
#pragma once
#include "fm_rows.cuh"

struct __align__(16) NnzMeta {
  double x;
  int32_t j;
  int32_t acc;   // element offset of the hot-slot accumulator row in sAcc, or -1 (cold)
};

__host__ __device__ inline size_t fast_group_smem(int CH, int SB8, int nHotTot) {
  size_t b = ((size_t)CH * SB8 + (size_t)nHotTot * (SB8 + 1)) * 8 + (size_t)CH * sizeof(NnzMeta);
  return (b + 15) & ~(size_t)15;
}

template <int DEGREE, bool EXPLICIT, int MODE, int KT>
__global__ void __launch_bounds__(256, 1) fm_rows_fast_kernel(const RowArgs a) {
  static_assert(MODE == MODE_PREDICT || MODE == MODE_GRAD, "fast path: predict / grad only");
  constexpr int NO = RowCfg<DEGREE, EXPLICIT>::NO;
  constexpr int G = KT;
  constexpr int GPW = 32 / G;
  constexpr int SB8 = NO * KT;
  constexpr int ASTR = SB8 + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warpInBlock = threadIdx.x >> 5;
  const int gl = lane & (G - 1);
  const int gidInWarp = lane / G;
  const int CH = a.CH;
  const int nHotTot = (MODE == MODE_PREDICT) ? 0 : a.nHot + a.nAug;
  const size_t perGroup = fast_group_smem(CH, SB8, nHotTot);
  unsigned char *base = smem_raw + (size_t)(warpInBlock * GPW + gidInWarp) * perGroup;
  double *sP = reinterpret_cast<double *>(base);
  NnzMeta *sMeta = reinterpret_cast<NnzMeta *>(sP + (size_t)CH * SB8);
  double *sAcc = reinterpret_cast<double *>(sMeta + CH);
  for (int e = gl; e < nHotTot * ASTR; e += G) sAcc[e] = 0.0;
  __syncwarp();

  const int warpsPerBlock = blockDim.x >> 5;
  const int64_t warpGlobal = (int64_t)blockIdx.x * warpsPerBlock + warpInBlock;
  const int64_t nWarps = (int64_t)gridDim.x * warpsPerBlock;
  const int64_t tiles = (a.nRows + GPW - 1) / GPW;
  double accLoss = 0.0, accB1 = 0.0;
  const double bias = a.b[0];
  const double *__restrict__ Pg = a.P;

  for (int64_t tile = warpGlobal; tile < tiles; tile += nWarps) {
    const int64_t q = tile * GPW + gidInWarp;
    const bool active = q < a.nRows;
    int64_t r = 0;
    if (active) r = a.rowIdx ? (int64_t)a.rowIdx[q] : (a.rowBegin + q) % a.n;
    const int64_t rb = active ? a.indptr[r] : 0;
    const int zReal = active ? (int)(a.indptr[r + 1] - rb) : 0;
    const int z = active ? zReal + a.nAug : 0;

    // ---- stage metadata (coalesced index/value loads) + linear term
    __syncwarp();
    double lin = 0.0;
    for (int u = gl; u < z; u += G) {
      NnzMeta m;
      if (u < zReal) {
        m.j = a.indices[rb + u];
        m.x = a.data[rb + u];
        lin += a.w[m.j] * m.x;
        m.acc = -1;
        if (MODE != MODE_PREDICT && a.hotSlot) {
          const int slot = a.hotSlot[m.j];
          if (slot != NIMFM_COLD) m.acc = slot * ASTR;
        }
      } else {
        m.j = (int32_t)(a.d + (u - zReal));
        m.x = 1.0;
        m.acc = (MODE != MODE_PREDICT) ? (a.nHot + (u - zReal)) * ASTR : -1;
      }
      sMeta[u] = m;
    }
    __syncwarp();
    // ---- stage the row's P slice: z x SB8 doubles, 16 B per cp.async
    {
      constexpr int UNITS = SB8 / 2;   // 16-byte units per feature slice (power of two)
      const int total = z * UNITS;
      for (int v = gl; v < total; v += G) {
        const int qq = v / UNITS, off = (v % UNITS) * 2;
        cp_async16(sP + qq * SB8 + off, Pg + (int64_t)sMeta[qq].j * SB8 + off);
      }
      cp_async_wait_all();
    }
    __syncwarp();

    // ---- forward: degree-m ANOVA DP for all orders (sgd.nim:146-173)
    double A[NO][DEGREE + 1];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      A[o][0] = 1.0;
#pragma unroll
      for (int t = 1; t <= DEGREE; ++t) A[o][t] = 0.0;
    }
    {
      const double *ps = sP + gl;
      const NnzMeta *pm = sMeta;
#pragma unroll 4
      for (int u = 0; u < z; ++u, ps += SB8, ++pm) {
        const double x = pm->x;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          const int M = DEGREE - o;
          const double t = ps[o * KT] * x;
          if (M == 2) {
            A[o][1] += t;
            A[o][2] += t * t;
          } else {
#pragma unroll
            for (int tt = DEGREE; tt >= 1; --tt)
              if (tt <= M) A[o][tt] += A[o][tt - 1] * t;
          }
        }
      }
    }
    double part = 0.0;
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      const int M = DEGREE - o;
      if (M == 2) A[o][2] = (A[o][1] * A[o][1] - A[o][2]) / 2.0;
      part += A[o][M];
    }
    if (MODE == MODE_PREDICT && a.lams != nullptr) part *= a.lams[gl];
    double ysum = lin + part;
#pragma unroll
    for (int off = G >> 1; off > 0; off >>= 1) ysum += __shfl_xor_sync(0xffffffffu, ysum, off);
    const double yhat = bias + ysum;

    if (MODE == MODE_PREDICT) {
      if (active && gl == 0 && a.yOut) a.yOut[q] = yhat;
      continue;
    }

    // ---- loss derivative and backward (sgd.nim:176-188 + minibatch_psgd.nim:73-88)
    double coef = 0.0;
    if (active) {
      if (a.yOut && gl == 0) a.yOut[q] = yhat;
      const double yi = a.y[r];
      coef = dev_dloss(a.loss, a.thr, yi, yhat) / a.mb;
      if (gl == 0) {
        accLoss += dev_loss(a.loss, a.thr, yi, yhat);
        accB1 += coef;
      }
    }
    {
      const double *ps = sP + gl;
      const NnzMeta *pm = sMeta;
      double *__restrict__ gPg = a.gP + gl;
#pragma unroll 2
      for (int u = 0; u < z; ++u, ps += SB8, ++pm) {
        const NnzMeta m = *pm;
        double gr[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          const int M = DEGREE - o;
          const double p = ps[o * KT];
          double g;
          if (M == 2) {
            g = m.x * (A[o][1] - p * m.x);
          } else {
            g = m.x;
#pragma unroll
            for (int tt = 1; tt < DEGREE; ++tt)
              if (tt < M) g = m.x * (A[o][tt] - p * g);
          }
          gr[o] = coef * g;
        }
        if (m.acc >= 0) {   // hot column: lane-private shared accumulator element
          double *ac = sAcc + m.acc + gl;
#pragma unroll
          for (int o = 0; o < NO; ++o) ac[o * KT] += gr[o];
        } else {
          double *gp = gPg + (int64_t)m.j * SB8;
#pragma unroll
          for (int o = 0; o < NO; ++o) atomicAdd(gp + o * KT, gr[o]);
        }
      }
    }
    // ---- linear-term gradient (real features only): one lane per nonzero
    if (a.fitLinear) {
      for (int u = gl; u < zReal; u += G) {
        const NnzMeta m = sMeta[u];
        const double gx = coef * m.x;
        if (m.acc >= 0) atomicAdd(sAcc + m.acc + SB8, gx);   // shared atomic, low contention
        else atomicAdd(a.gw + m.j, gx);
      }
    }
  }

  if (MODE != MODE_PREDICT) {
    __syncwarp();
    for (int slot = 0; slot < nHotTot; ++slot) {
      const int64_t j = slot < a.nHot ? (int64_t)a.hotList[slot] : a.d + (slot - a.nHot);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        const double v = sAcc[slot * ASTR + o * KT + gl];
        if (v != 0.0) atomicAdd(a.gP + j * SB8 + o * KT + gl, v);
      }
      if (gl == 0 && a.fitLinear && j < a.d) {
        const double v = sAcc[slot * ASTR + SB8];
        if (v != 0.0) atomicAdd(a.gw + j, v);
      }
    }
    accLoss = warp_sum(accLoss);
    accB1 = warp_sum(accB1);
    if (lane == 0) {
      a.partials[warpGlobal * 4 + 0] = accLoss;
      a.partials[warpGlobal * 4 + 1] = accB1;
      a.partials[warpGlobal * 4 + 2] = 0.0;
      a.partials[warpGlobal * 4 + 3] = 0.0;
    }
  }
}

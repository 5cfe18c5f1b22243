// ffm_pairs.cuh -- pair-streaming FFM kernel (K10 / K11: decisionFunction, predict+grad, AdaGrad
// minibatch) for nComponents KT in {4, 8, 16, 32}: ONE WARP PER ROW, 32/KT "pair slots" of KT lanes each.
//
// It follows the reference's own loop (field_aware_factorization_machine.nim:68-76, sgd_ffm.nim:23-30):
// for every unordered pair of the row's nonzeros (u, v),
//     yhat      += x_u x_v <P[j_u][f_v][:], P[j_v][f_u][:]>
//     dA[f_v][j_u][:] += x_u x_v P[j_v][f_u][:]      dA[f_u][j_v][:] += x_u x_v P[j_u][f_v][:]
// The z(z-1)/2 pairs of a row are enumerated by ONE linear index (slot s takes pairs s, s+SLOTS, ...),
// so every slot has the same trip count, and they are processed PB at a time: the 2*PB gathers of a
// batch (each a coalesced KT*8-byte segment of the P[j][f][s] layout) are all issued before the first
// is consumed.  ncu on the previous form (2 pairs in flight, r01g): issue slots 16 % busy,
// long_scoreboard 20 warps per issue -- purely latency-bound on its own gathers with DRAM at 22 %.
// The backward pass re-loads the two vectors (mostly L2 hits) and emits both gradient vectors with
// FP64 RED atomics; AdaGrad also emits their squares (valid because the AdaGrad route is only taken
// for datasets with at most one nonzero per (row, field), where a pair's contribution IS the sample's
// whole gradient entry -- adagrad.nim:119-124 squares the per-sample gradient).
// Nothing is staged in shared memory except the row's {x, j, f} records, so occupancy is
// register-bound (vs 2 rows/SM for the block-per-row kernel in ffm.cu, the general fallback).
#pragma once
#include "common.cuh"

struct FfmArgs;   // defined in ffm.cu

struct __align__(16) FfmRec {
  double x;
  int32_t j;
  int32_t f;
};

enum { FFM_PAIRS_PREDICT = 0, FFM_PAIRS_GRAD = 1, FFM_PAIRS_ADAGRAD = 2 };

// advance the pair cursor (u, v), u < v < z, by `step` positions of the row-major pair order
__device__ __forceinline__ void ffm_pair_advance(int &u, int &v, int step, int z) {
  v += step;
  while (v >= z && u < z - 1) {
    v = v - z + u + 2;
    u += 1;
  }
}

template <int MODE, int KT, class Args>
__global__ void __launch_bounds__(256, 2) ffm_pairs_kernel(const Args a) {
  constexpr int SLOTS = 32 / KT;
  constexpr int PB = 8;   // pairs in flight per slot (2*PB gathers per lane)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warpInBlock = threadIdx.x >> 5;
  const int s = lane & (KT - 1);
  const int slot = lane / KT;
  const int CH = a.CH;
  FfmRec *rec = reinterpret_cast<FfmRec *>(smem_raw) + (size_t)warpInBlock * CH;
  const int warpsPerBlock = blockDim.x >> 5;
  const int64_t warpGlobal = (int64_t)blockIdx.x * warpsPerBlock + warpInBlock;
  const int64_t nWarps = (int64_t)gridDim.x * warpsPerBlock;
  const int64_t SB8 = (int64_t)a.nFields * KT;
  double bias = a.b[0];
  if (MODE == FFM_PAIRS_ADAGRAD && !a.first && a.fitIntercept)   // adagrad.nim:101-105 (batch snapshot)
    bias = -a.eta0 * a.adaScal[0] / (sqrt(a.adaScal[1]) + a.eta0 * a.tIt * a.alpha0);
  const double *__restrict__ Pg = a.P + s;
  double accLoss = 0.0, accB1 = 0.0, accB2 = 0.0;

  for (int64_t q = warpGlobal; q < a.nRows; q += nWarps) {
    const int64_t r = a.rowIdx ? (int64_t)a.rowIdx[q] : (a.rowBegin + q) % a.n;
    const int64_t rb = a.indptr[r];
    const int z = (int)(a.indptr[r + 1] - rb);
    __syncwarp();
    double lin = 0.0;
    for (int u = lane; u < z; u += 32) {
      FfmRec m;
      m.j = a.indices[rb + u];
      m.f = a.fields[rb + u];
      m.x = a.data[rb + u];
      rec[u] = m;
      lin += a.w[m.j] * m.x;
    }
    __syncwarp();
    const int nPairs = z * (z - 1) / 2;

    // ---- forward: running dot over this slot's pairs
    double acc = 0.0;
    {
      int u = 0, v = 1;
      ffm_pair_advance(u, v, slot, z);
      for (int p = slot; p < nPairs; p += SLOTS * PB) {
        double a1[PB], a2[PB], xx[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i) {
          const bool ok = p + i * SLOTS < nPairs;
          a1[i] = 0.0; a2[i] = 0.0; xx[i] = 0.0;
          if (ok) {
            const FfmRec mu = rec[u], mv = rec[v];
            if (mv.j != mu.j) {                                                  // the reference pairs j1 < j2 only
              a1[i] = __ldg(Pg + (int64_t)mu.j * SB8 + mv.f * KT);               // P[j_u][f_v][s]
              a2[i] = __ldg(Pg + (int64_t)mv.j * SB8 + mu.f * KT);               // P[j_v][f_u][s]
              xx[i] = mu.x * mv.x;
            }
          }
          ffm_pair_advance(u, v, SLOTS, z);
        }
#pragma unroll
        for (int i = 0; i < PB; ++i) acc += xx[i] * (a1[i] * a2[i]);
      }
    }
    const double yhat = bias + warp_sum(lin + acc);
    if (lane == 0 && a.yOut) a.yOut[q] = yhat;
    if (MODE == FFM_PAIRS_PREDICT) continue;

    const double yi = a.y[r];
    const double dL = dev_dloss(a.loss, a.thr, yi, yhat);
    const double coef = (MODE == FFM_PAIRS_GRAD) ? dL / a.mb : dL;
    if (lane == 0) {
      accLoss += dev_loss(a.loss, a.thr, yi, yhat);
      accB1 += coef;
      accB2 += dL * dL;
    }
    // ---- backward: both gradient vectors of every pair
    double *__restrict__ gPg = a.gP + s;
    double *__restrict__ gNg = (MODE == FFM_PAIRS_ADAGRAD) ? a.dGnP + s : nullptr;
    {
      int u = 0, v = 1;
      ffm_pair_advance(u, v, slot, z);
      for (int p = slot; p < nPairs; p += SLOTS * PB) {
        double a1[PB], a2[PB], cx[PB];
        int64_t e1[PB], e2[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i) {
          const bool ok = p + i * SLOTS < nPairs;
          e1[i] = -1; e2[i] = 0; a1[i] = 0.0; a2[i] = 0.0; cx[i] = 0.0;
          if (ok) {
            const FfmRec mu = rec[u], mv = rec[v];
            if (mv.j != mu.j) {
              e1[i] = (int64_t)mu.j * SB8 + mv.f * KT;       // entry (j_u, f_v)
              e2[i] = (int64_t)mv.j * SB8 + mu.f * KT;       // entry (j_v, f_u)
              a1[i] = __ldg(Pg + e1[i]);
              a2[i] = __ldg(Pg + e2[i]);
              cx[i] = coef * (mu.x * mv.x);
            }
          }
          ffm_pair_advance(u, v, SLOTS, z);
        }
#pragma unroll
        for (int i = 0; i < PB; ++i) {
          if (e1[i] >= 0) {
            const double g1 = cx[i] * a2[i], g2 = cx[i] * a1[i];
            atomicAdd(gPg + e1[i], g1);
            atomicAdd(gPg + e2[i], g2);
            if (MODE == FFM_PAIRS_ADAGRAD) {
              atomicAdd(gNg + e1[i], g1 * g1);
              atomicAdd(gNg + e2[i], g2 * g2);
            }
          }
        }
      }
    }
    if (a.fitLinear)
      for (int u = lane; u < z; u += 32) {
        const double gx = coef * rec[u].x;
        atomicAdd(a.gw + rec[u].j, gx);
        if (MODE == FFM_PAIRS_ADAGRAD) atomicAdd(a.dGnw + rec[u].j, gx * gx);
      }
  }
  if (MODE != FFM_PAIRS_PREDICT) {
    accLoss = warp_sum(accLoss);
    accB1 = warp_sum(accB1);
    accB2 = warp_sum(accB2);
    if (lane == 0) {
      a.partials[warpGlobal * 4 + 0] = accLoss;
      a.partials[warpGlobal * 4 + 1] = accB1;
      a.partials[warpGlobal * 4 + 2] = accB2;
      a.partials[warpGlobal * 4 + 3] = 0.0;
    }
  }
}

// flag <- 1 if any row has two nonzeros of the same field (warp per row; the AdaGrad pair route needs 0)
static __global__ void ffm_field_dup_kernel(const int32_t *fields, const int64_t *indptr, int64_t n, int *flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nWarps) {
    const int64_t b = indptr[r], e = indptr[r + 1];
    for (int64_t u = b + lane; u < e; u += 32) {
      const int32_t f = fields[u];
      for (int64_t v = u + 1; v < e; ++v)
        if (fields[v] == f) *flag = 1;
    }
  }
}

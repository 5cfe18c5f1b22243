// ffm_pairs.cuh -- pair-streaming FFM kernel (K10 / K11 predict+grad) for nComponents KT in
// {4, 8, 16, 32}: ONE WARP PER ROW, 32/KT "pair slots" of KT lanes each.
//
// It follows the reference's own loop (field_aware_factorization_machine.nim:68-76, sgd_ffm.nim:23-30):
// for every unordered pair of the row's nonzeros (u, v),
//     yhat      += x_u x_v <P[j_u][f_v][:], P[j_v][f_u][:]>
//     dA[f_v][j_u][:] += x_u x_v P[j_v][f_u][:]      dA[f_u][j_v][:] += x_u x_v P[j_u][f_v][:]
// A slot handles one pair at a time: its KT lanes load the two KT-double vectors (two coalesced
// KT*8-byte segments of the P[j][f][s] layout), multiply, and keep a running dot; the backward pass
// re-loads the two vectors (L1/L2 hits) and emits both gradient vectors with FP64 RED atomics
// (several v of the same field add into the same (u, f_v) entry -- the gradient is linear, so the
// atomics sum them exactly as the reference's dA accumulation does).  Nothing is staged in shared
// memory except the row's {x, j, f} records, so occupancy is register-bound (vs 2 rows/SM for the
// block-per-row kernel in ffm.cu, which stays as the general fallback and the AdaGrad path).
#pragma once
#include "common.cuh"

struct FfmArgs;   // defined in ffm.cu

struct __align__(16) FfmRec {
  double x;
  int32_t j;
  int32_t f;
};

template <int MODE_GRAD_FLAG, int KT, class Args>
__global__ void __launch_bounds__(256, 2) ffm_pairs_kernel(const Args a) {
  constexpr int SLOTS = 32 / KT;
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
  const double bias = a.b[0];
  const double *__restrict__ Pg = a.P + s;
  double accLoss = 0.0, accB1 = 0.0;

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
    // ---- forward: running dot over this slot's pairs
    double acc = 0.0;
    for (int u = 0; u + 1 < z; ++u) {
      const FfmRec mu = rec[u];
      const double *pu = Pg + (int64_t)mu.j * SB8;
#pragma unroll 4
      for (int v = u + 1 + slot; v < z; v += SLOTS) {
        const FfmRec mv = rec[v];
        if (mv.j == mu.j) continue;                                          // the reference pairs j1 < j2 only
        const double a1 = __ldg(pu + mv.f * KT);                            // P[j_u][f_v][s]
        const double a2 = __ldg(Pg + (int64_t)mv.j * SB8 + mu.f * KT);      // P[j_v][f_u][s]
        acc += (mu.x * mv.x) * (a1 * a2);
      }
    }
    const double yhat = bias + warp_sum(lin + acc);
    if (lane == 0 && a.yOut) a.yOut[q] = yhat;
    if (!MODE_GRAD_FLAG) continue;

    const double yi = a.y[r];
    const double coef = dev_dloss(a.loss, a.thr, yi, yhat) / a.mb;
    if (lane == 0) {
      accLoss += dev_loss(a.loss, a.thr, yi, yhat);
      accB1 += coef;
    }
    // ---- backward: both gradient vectors of every pair
    double *__restrict__ gPg = a.gP + s;
    for (int u = 0; u + 1 < z; ++u) {
      const FfmRec mu = rec[u];
      const int64_t bu = (int64_t)mu.j * SB8;
#pragma unroll 2
      for (int v = u + 1 + slot; v < z; v += SLOTS) {
        const FfmRec mv = rec[v];
        if (mv.j == mu.j) continue;
        const int64_t e1 = bu + mv.f * KT;                     // entry (j_u, f_v)
        const int64_t e2 = (int64_t)mv.j * SB8 + mu.f * KT;    // entry (j_v, f_u)
        const double a1 = __ldg(Pg + e1);
        const double a2 = __ldg(Pg + e2);
        const double cx = coef * (mu.x * mv.x);
        atomicAdd(gPg + e1, cx * a2);
        atomicAdd(gPg + e2, cx * a1);
      }
    }
    if (a.fitLinear)
      for (int u = lane; u < z; u += 32) atomicAdd(a.gw + rec[u].j, coef * rec[u].x);
  }
  if (MODE_GRAD_FLAG) {
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

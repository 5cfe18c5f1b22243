// ffm_tma.cuh -- FFM pair kernel with the row's parameter blocks staged by the TMA unit (K10 / K11).
//
// In the P[j][f][s] layout a row touches, for each of its z nonzeros, ONE contiguous block of
// nFields*k doubles (2 496 B for C5).  The pair kernels of ffm_pairs.cuh gather those blocks 64 bytes at a
// time from the pair loop; here each block is fetched by one `cp.async.bulk.shared::cluster.global`
// (1-D bulk copy, no tensor map) issued by the thread that owns the nonzero, completion counted on an
// mbarrier (`complete_tx::bytes`), into one of TWO shared-memory stages: the z bulk copies of row r+1
// are in flight while the block runs the pair loop of row r entirely out of shared memory.  One block of
// 512 threads per SM, 2 x z*2 496 B of staging (195 KB for z = 39), grid = #SMs.
// DRAM sees whole 2.5 KB runs instead of scattered 64-byte segments and the pair loop has no global
// loads at all; the gradient still leaves as FP64 REDs (AdaGrad: plus their squares, same validity
// condition as ffm_pairs.cuh).  Rows up to FFM_PAIRS_TABLE_MAXZ nonzeros (shared pair table).
#pragma once
#include "ffm_pairs.cuh"

#define FFM_TMA_THREADS 512

__device__ __forceinline__ unsigned tma_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tma_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(tma_smem_addr(bar))
               : "memory");
}
// bounded wait: a lost completion traps instead of hanging the GPU
__device__ __forceinline__ void tma_mbar_wait(uint64_t *bar, unsigned parity) {
  const unsigned a = tma_smem_addr(bar);
  for (unsigned spin = 0;; ++spin) {
    unsigned done;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 26)) __trap();
  }
}

// bytes of dynamic shared memory: 2 stages of z*SB8 doubles | 2 x z records | 2 x z w-values | pair table
__host__ __device__ inline size_t ffm_tma_smem(int CH, int SB8) {
  size_t b = 2 * (size_t)CH * SB8 * 8 + 2 * (size_t)CH * sizeof(FfmRec) + 2 * (size_t)CH * 8 +
             (size_t)CH * (CH - 1) / 2 * sizeof(uint16_t);
  return (b + 15) & ~(size_t)15;
}

template <int MODE, int KT, class Args>
__global__ void __launch_bounds__(FFM_TMA_THREADS, 1) ffm_rows_tma_kernel(const Args a) {
  constexpr int SLOTS = 32 / KT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ double red[FFM_TMA_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nWarpsB = blockDim.x >> 5;
  const int sl = lane & (KT - 1);
  const int gslot = wid * SLOTS + lane / KT, nSlots = nWarpsB * SLOTS;
  const int CH = a.CH, nF = a.nFields, SB8 = nF * KT;
  double *stage0 = reinterpret_cast<double *>(smem_raw);
  FfmRec *rec0 = reinterpret_cast<FfmRec *>(smem_raw + 2 * (size_t)CH * SB8 * 8);
  double *recW0 = reinterpret_cast<double *>(rec0 + 2 * CH);
  uint16_t *tab = reinterpret_cast<uint16_t *>(recW0 + 2 * CH);
  int zTab = -1;
  double bias = a.b[0];
  if (MODE == FFM_PAIRS_ADAGRAD && !a.first && a.fitIntercept)
    bias = -a.eta0 * a.adaScal[0] / (sqrt(a.adaScal[1]) + a.eta0 * a.tIt * a.alpha0);
  double accLoss = 0.0, accB1 = 0.0, accB2 = 0.0;
  if (tid == 0) {
    tma_mbar_init(&mbar[0], 1);
    tma_mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // records of row q into stage st, and the bulk copies of its z parameter blocks
  auto issue = [&](int64_t q, int st) {
    const int64_t r = a.rowIdx ? (int64_t)a.rowIdx[q] : (a.rowBegin + q) % a.n;
    const int64_t rb = a.indptr[r];
    const int z = (int)(a.indptr[r + 1] - rb);
    if (tid == 0) tma_mbar_expect_tx(&mbar[st], (unsigned)((size_t)z * SB8 * 8));
    if (tid < z) {
      FfmRec m;
      const int32_t j = a.indices[rb + tid];
      m.jb = j;                       // here: the feature id itself
      m.f = a.fields[rb + tid];
      m.x = a.data[rb + tid];
      rec0[st * CH + tid] = m;
      recW0[st * CH + tid] = a.w[j];
      tma_bulk_g2s(stage0 + ((size_t)st * CH + tid) * SB8, a.P + (int64_t)j * SB8, (unsigned)(SB8 * 8), &mbar[st]);
    }
  };

  int64_t q = blockIdx.x;
  if (q < a.nRows) issue(q, 0);
  for (int it = 0; q < a.nRows; q += gridDim.x, ++it) {
    const int st = it & 1;
    const unsigned parity = (unsigned)((it >> 1) & 1);
    const int64_t qn = q + gridDim.x;
    if (qn < a.nRows) issue(qn, st ^ 1);   // the other stage was released by the barrier ending the last iteration
    const int64_t r = a.rowIdx ? (int64_t)a.rowIdx[q] : (a.rowBegin + q) % a.n;
    const int z = (int)(a.indptr[r + 1] - a.indptr[r]);
    const int nPairs = z * (z - 1) / 2;
    if (z != zTab) {                       // block-uniform: every thread sees the same row
      int u = 0, v = 1;
      ffm_pair_advance(u, v, tid, z);
      for (int p = tid; p < nPairs; p += blockDim.x) {
        tab[p] = (uint16_t)(u | (v << 8));
        ffm_pair_advance(u, v, blockDim.x, z);
      }
      zTab = z;
    }
    tma_mbar_wait(&mbar[st], parity);      // the row's z blocks have landed (async-proxy writes visible)
    __syncthreads();                       // records + pair table written by other threads
    const double *W = stage0 + (size_t)st * CH * SB8 + sl;
    const FfmRec *rec = rec0 + st * CH;
    const double *recW = recW0 + st * CH;

    // ---- forward, out of shared memory
    double acc = 0.0;
    for (int u = tid; u < z; u += blockDim.x) acc += recW[u] * rec[u].x;
    for (int p = gslot; p < nPairs; p += nSlots) {
      const unsigned uv = tab[p];
      const int u = uv & 0xff, v = uv >> 8;
      const FfmRec mu = rec[u], mv = rec[v];
      if (mv.jb != mu.jb) acc += (mu.x * mv.x) * (W[u * SB8 + mv.f * KT] * W[v * SB8 + mu.f * KT]);
    }
    acc = warp_sum(acc);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    double tot = 0.0;
    for (int wq = 0; wq < nWarpsB; ++wq) tot += red[wq];
    const double yhat = bias + tot;
    if (tid == 0 && a.yOut) a.yOut[q] = yhat;

    if (MODE != FFM_PAIRS_PREDICT) {
      const double yi = a.y[r];
      const double dL = dev_dloss(a.loss, a.thr, yi, yhat);
      const double coef = (MODE == FFM_PAIRS_GRAD) ? dL / a.mb : dL;
      if (tid == 0) {
        accLoss += dev_loss(a.loss, a.thr, yi, yhat);
        accB1 += coef;
        accB2 += dL * dL;
      }
      // ---- backward: both gradient vectors of every pair, values from shared memory, REDs to global
      double *__restrict__ gPg = a.gP + sl;
      double *__restrict__ gNg = (MODE == FFM_PAIRS_ADAGRAD) ? a.dGnP + sl : nullptr;
      for (int p = gslot; p < nPairs; p += nSlots) {
        const unsigned uv = tab[p];
        const int u = uv & 0xff, v = uv >> 8;
        const FfmRec mu = rec[u], mv = rec[v];
        if (mv.jb == mu.jb) continue;
        const double cx = coef * (mu.x * mv.x);
        const double g1 = cx * W[v * SB8 + mu.f * KT], g2 = cx * W[u * SB8 + mv.f * KT];
        const int64_t e1 = ((int64_t)mu.jb * nF + mv.f) * KT, e2 = ((int64_t)mv.jb * nF + mu.f) * KT;
        atomicAdd(gPg + e1, g1);
        atomicAdd(gPg + e2, g2);
        if (MODE == FFM_PAIRS_ADAGRAD) {
          atomicAdd(gNg + e1, g1 * g1);
          atomicAdd(gNg + e2, g2 * g2);
        }
      }
      if (a.fitLinear)
        for (int u = tid; u < z; u += blockDim.x) {
          const double gx = coef * rec[u].x;
          atomicAdd(a.gw + rec[u].jb, gx);
          if (MODE == FFM_PAIRS_ADAGRAD) atomicAdd(a.dGnw + rec[u].jb, gx * gx);
        }
    }
    // this stage's generic-proxy reads are done before the next bulk copy (async proxy) may overwrite it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  if (MODE != FFM_PAIRS_PREDICT && tid == 0) {
    a.partials[blockIdx.x * 4 + 0] = accLoss;
    a.partials[blockIdx.x * 4 + 1] = accB1;
    a.partials[blockIdx.x * 4 + 2] = accB2;
    a.partials[blockIdx.x * 4 + 3] = 0.0;
  }
}

// peer.cu -- the gradient exchange of the row-sharded solvers over NVLink / NVSwitch peer memory
// (one process per GPU; the reference has no distributed backend at all -- SURVEY 2a -- this replaces the
// "ncclAllReduce, then the identical dense step on every GPU" schedule of SURVEY 8e).
//
// Every buffer that takes part in the exchange (gradient pool, parameter pool, AdaGrad delta block, touch
// counts) lives in an ARENA whose CUDA IPC handle every rank has opened, so a kernel on rank r can read and
// write the same buffer of every peer directly.  One exchange is
//     barrier  ->  reduce kernel  ->  barrier
// where the reduce kernel, on rank r, walks ITS 1/N slice of the buffer: 16-byte loads of the slice out of all
// N ranks' buffers (N loads in flight per element pair), a sum in rank order 0..N-1 (so the result is
// bit-identical on every rank and from run to run), an element-wise functor -- identity for a plain
// all-reduce; Params.step + the L1 prox for MBPSGD, which turns reduce-scatter + sharded step + all-gather
// into ONE pass -- and 16-byte stores of the result into all N ranks' output buffers.  Per rank that moves
// (N-1)/N of the buffer in and out over NVLink once, the same bytes as an all-reduce, without NCCL's
// staging copies, and 1/N of the dense step instead of all of it on every GPU.
//
// The barrier is a one-block kernel: rank r stores the barrier's sequence number into slot r of every peer's
// flag array (st.release.sys) and spins until its own slots hold it (ld.acquire.sys).  Counters only grow and
// every slot has one writer, so consecutive barriers need no reset.  A peer that never arrives (a crashed
// rank) trips a timeout and the kernel traps instead of hanging the GPU.
#include <stdlib.h>

#include <algorithm>

#include "dense_kernels.cuh"

namespace {

struct PeerPtrs { double *p[NIMFM_MAX_RANKS]; };
struct PeerFlags { uint32_t *f[NIMFM_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_barrier_kernel(PeerFlags pf, int rank, int R, uint32_t seq, unsigned long long timeoutNs) {
  const int p = threadIdx.x;
  if (p >= R) return;
  __threadfence_system();                     // this GPU's earlier stores (local and remote) are ordered before the flag
  st_release_sys(pf.f[p] + rank, seq);
  const unsigned long long t0 = global_ns();
  while ((int32_t)(ld_acquire_sys(pf.f[rank] + p) - seq) < 0) {
    if (global_ns() - t0 > timeoutNs) {
      printf("libnimfm_cuda: rank %d waited %llu s for rank %d at exchange barrier %u -- aborting\n", rank,
             timeoutNs / 1000000000ull, p, seq);
      __trap();
    }
    __nanosleep(200);
  }
}

struct IdentityF {
  __device__ __forceinline__ double operator()(int64_t, double s) const { return s; }
};

struct MbpsgdStepF {
  const double *par;   // this rank's parameter pool (identical on every rank)
  MbpsgdStepArgs a;
  __device__ __forceinline__ double operator()(int64_t e, double g) const {
    const double v = par[e];
    if (e < a.nP) {
      double p = (v + a.negEtaP * g) * a.rP;
      if (a.reg == NIMFM_REG_L1) {   // softthreshold, regularizer/utils.nim:4-5
        const double m = fabs(p) - a.lam;
        p = (p > 0 ? 1.0 : (p < 0 ? -1.0 : 0.0)) * (m > 0.0 ? m : 0.0);
      }
      return p;
    }
    if (e < a.nP + a.d) return a.fitLinear ? (v + a.negEtaW * g) * a.rW : v;
    if (e == a.nP + a.d) {
      double bb = v;
      if (a.fitIntercept && a.fitLinear) bb += a.negEtaB * g;   // params.nim:47
      if (a.fitIntercept) bb *= a.rB;                           // params.nim:65-66
      return bb;
    }
    if (e == a.nP + a.d + 1) return v + g;                      // the epoch's loss sum
    return v;
  }
};

// elements [lo, hi) (lo even): sum over ranks in rank order, f, store to nOut buffers (all ranks or the local one)
template <int R, class F>
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerPtrs in, PeerPtrs out, int nOut, int64_t lo, int64_t hi, F f) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t p0 = lo >> 1, p1 = hi >> 1;   // whole double2 pairs
  for (int64_t i = p0 + tid; i < p1; i += 2 * stride) {
    const int64_t i2 = i + stride;
    const bool two = i2 < p1;
    double2 a[R], b[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      a[r] = __ldcg(reinterpret_cast<const double2 *>(in.p[r]) + i);
      if (two) b[r] = __ldcg(reinterpret_cast<const double2 *>(in.p[r]) + i2);
    }
    double2 s = a[0], t = two ? b[0] : make_double2(0.0, 0.0);
#pragma unroll
    for (int r = 1; r < R; r++) {
      s.x += a[r].x;
      s.y += a[r].y;
      if (two) {
        t.x += b[r].x;
        t.y += b[r].y;
      }
    }
    const double2 o = make_double2(f(2 * i, s.x), f(2 * i + 1, s.y));
    for (int r = 0; r < nOut; r++) reinterpret_cast<double2 *>(out.p[r])[i] = o;
    if (two) {
      const double2 o2 = make_double2(f(2 * i2, t.x), f(2 * i2 + 1, t.y));
      for (int r = 0; r < nOut; r++) reinterpret_cast<double2 *>(out.p[r])[i2] = o2;
    }
  }
  if ((hi & 1) && tid == 0) {   // odd tail element
    const int64_t e = hi - 1;
    double s = __ldcg(in.p[0] + e);
    for (int r = 1; r < R; r++) s += __ldcg(in.p[r] + e);
    const double o = f(e, s);
    for (int r = 0; r < nOut; r++) out.p[r][e] = o;
  }
}

__global__ void __launch_bounds__(256) peer_broadcast_kernel(const double *src, PeerPtrs out, int R, int self, int64_t lo,
                                                           int64_t hi) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < hi; e += stride) {
    const double v = src[e];
    for (int r = 0; r < R; r++)
      if (r != self) out.p[r][e] = v;
  }
}

bool lookup(const nimfm_ctx *ctx, const double *ptr, int64_t n, PeerPtrs *out) {
  for (const auto &a : ctx->arenas) {
    if (ptr >= a.base && ptr + n <= a.base + a.nDoubles) {
      const int64_t off = ptr - a.base;
      for (int r = 0; r < ctx->nranks; r++) out->p[r] = a.peer[r] + off;
      return true;
    }
  }
  return false;
}

// every rank's handle for `base` -> peer pointers; ok = 0 on ANY rank makes the whole call fail on every rank
int exchange_open(nimfm_ctx *ctx, void *base, int ok, void **peerOut, bool *allOk) {
  const int R = ctx->nranks;
  int64_t mine[9] = {0};
  cudaIpcMemHandle_t h;
  if (ok && cudaIpcGetMemHandle(&h, base) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  if (ok) memcpy(mine, &h, sizeof(h));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  mine[8] = ok;
  std::vector<int64_t> all((size_t)9 * R);
  int rc = nimfm_allgather_host_i64(ctx, mine, 9, all.data());
  if (rc) return rc;
  bool good = true;
  for (int r = 0; r < R; r++) good = good && all[(size_t)9 * r + 8] != 0;
  int opened = good ? 1 : 0;
  if (good) {
    for (int r = 0; r < R; r++) {
      if (r == ctx->rank) { peerOut[r] = base; continue; }
      cudaIpcMemHandle_t hp;
      memcpy(&hp, &all[(size_t)9 * r], sizeof(hp));
      if (cudaIpcOpenMemHandle(&peerOut[r], hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        peerOut[r] = nullptr;
        opened = 0;
      }
    }
  }
  // second round: did every rank manage to map every peer?
  const int64_t st[1] = {opened};
  std::vector<int64_t> sts((size_t)R);
  if ((rc = nimfm_allgather_host_i64(ctx, st, 1, sts.data()))) return rc;
  bool every = true;
  for (int r = 0; r < R; r++) every = every && sts[(size_t)r] != 0;
  if (!every)
    for (int r = 0; r < R; r++)
      if (r != ctx->rank && good && peerOut[r]) {
        cudaIpcCloseMemHandle(peerOut[r]);
        peerOut[r] = nullptr;
      }
  *allOk = every;
  return NIMFM_OK;
}

template <class F>
int launch_reduce(nimfm_ctx *ctx, const PeerPtrs &in, const PeerPtrs &out, int nOut, int64_t lo, int64_t hi, F f) {
  if (hi <= lo) return NIMFM_OK;
  const int64_t pairs = (hi - lo + 1) / 2;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((pairs + 511) / 512, (int64_t)ctx->numSMs * 4));
  switch (ctx->nranks) {
#define NIMFM_PEER_CASE(RR) \
  case RR: peer_reduce_kernel<RR, F><<<grid, 256, 0, ctx->stream>>>(in, out, nOut, lo, hi, f); break;
    NIMFM_PEER_CASE(2) NIMFM_PEER_CASE(3) NIMFM_PEER_CASE(4) NIMFM_PEER_CASE(5) NIMFM_PEER_CASE(6) NIMFM_PEER_CASE(7)
    NIMFM_PEER_CASE(8)
#undef NIMFM_PEER_CASE
    default: return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "peer exchange is instantiated for 2..8 ranks");
  }
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  return NIMFM_OK;
}

}  // namespace

int nimfm_peer_barrier(nimfm_ctx *ctx) {
  PeerFlags pf;
  for (int r = 0; r < NIMFM_MAX_RANKS; r++) pf.f[r] = ctx->peerFlagsOf[r];
  ctx->barrierEpoch += 1;
  // generous: ranks reach their first exchange after host-side set-up of very different length
  static const unsigned long long timeoutNs =
      (getenv("NIMFM_PEER_TIMEOUT_S") ? strtoull(getenv("NIMFM_PEER_TIMEOUT_S"), nullptr, 10) : 180ull) * 1000000000ull;
  peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(pf, ctx->rank, ctx->nranks, ctx->barrierEpoch, timeoutNs);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  return NIMFM_OK;
}

int nimfm_peer_init(nimfm_ctx *ctx) {
  ctx->peerOK = false;
  const char *env = getenv("NIMFM_PEER");
  if (ctx->nranks < 2 || ctx->nranks > NIMFM_MAX_RANKS || (env && env[0] == '0')) return NIMFM_OK;
  int ok = 1;
  if (cudaMalloc(&ctx->peerFlags, 4096) != cudaSuccess || cudaMemset(ctx->peerFlags, 0, 4096) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  void *peers[NIMFM_MAX_RANKS] = {nullptr};
  bool all = false;
  int rc = exchange_open(ctx, ctx->peerFlags, ok, peers, &all);
  if (rc) return rc;
  if (!all) {   // e.g. no P2P between the GPUs of this box: the NCCL route stays in charge
    cudaFree(ctx->peerFlags);
    ctx->peerFlags = nullptr;
    return NIMFM_OK;
  }
  for (int r = 0; r < ctx->nranks; r++) ctx->peerFlagsOf[r] = static_cast<uint32_t *>(peers[r]);
  ctx->peerOK = true;
  ctx->barrierEpoch = 0;
  return NIMFM_OK;
}

void nimfm_peer_shutdown(nimfm_ctx *ctx) {
  for (auto &a : ctx->arenas) {
    for (int r = 0; r < ctx->nranks; r++)
      if (r != ctx->rank && a.peer[r]) cudaIpcCloseMemHandle(a.peer[r]);
    cudaFree(a.base);
  }
  ctx->arenas.clear();
  if (ctx->peerFlags) {
    for (int r = 0; r < ctx->nranks; r++)
      if (r != ctx->rank && ctx->peerFlagsOf[r]) cudaIpcCloseMemHandle(ctx->peerFlagsOf[r]);
    cudaFree(ctx->peerFlags);
    ctx->peerFlags = nullptr;
  }
  ctx->peerOK = false;
}

// Buffers that take part in the exchange.  With peer memory on this is COLLECTIVE: every rank must allocate the
// same sequence of sizes (the solvers do -- they run the same code on every rank).  Freed arenas are kept and
// recycled by size: an exporter must not free memory its peers still have mapped.
int nimfm_comm_alloc(nimfm_ctx *ctx, double **out, size_t nDoubles) {
  *out = nullptr;
  if (!ctx->peerOK) {
    CK(cudaMalloc(out, nDoubles * 8));
    return NIMFM_OK;
  }
  for (auto &a : ctx->arenas)
    if (!a.inUse && a.nDoubles == nDoubles) {
      a.inUse = true;
      *out = a.base;
      return NIMFM_OK;
    }
  nimfm_ctx::PeerArena a;
  a.nDoubles = nDoubles;
  int ok = 1;
  if (cudaMalloc(&a.base, nDoubles * 8) != cudaSuccess) {
    cudaGetLastError();
    a.base = nullptr;
    ok = 0;
  }
  void *peers[NIMFM_MAX_RANKS] = {nullptr};
  bool all = false;
  int rc = exchange_open(ctx, a.base, ok, peers, &all);
  if (rc) { cudaFree(a.base); return rc; }
  if (!all) {
    cudaFree(a.base);
    return nimfm_fail(ctx, NIMFM_ERR_CUDA, "peer-mapped allocation of %zu bytes failed on some rank", nDoubles * 8);
  }
  for (int r = 0; r < ctx->nranks; r++) a.peer[r] = static_cast<double *>(peers[r]);
  a.inUse = true;
  ctx->arenas.push_back(a);
  *out = a.base;
  return NIMFM_OK;
}

void nimfm_comm_free(nimfm_ctx *ctx, double *p) {
  if (!p) return;
  if (ctx)
    for (auto &a : ctx->arenas)
      if (a.base == p) {
        a.inUse = false;
        return;
      }
  cudaFree(p);
}

int nimfm_peer_allreduce_sum(nimfm_ctx *ctx, double *buf, int64_t n, int *done) {
  *done = 0;
  PeerPtrs pp;
  if (!ctx->peerOK || n <= 0 || (reinterpret_cast<uintptr_t>(buf) & 15) || !lookup(ctx, buf, n, &pp)) return NIMFM_OK;
  const int R = ctx->nranks;
  int64_t c = (n + R - 1) / R;
  c += c & 1;                                           // slices start on 16-byte boundaries
  const int64_t lo = std::min<int64_t>(n, (int64_t)ctx->rank * c), hi = std::min<int64_t>(n, lo + c);
  int rc;
  if ((rc = nimfm_peer_barrier(ctx))) return rc;        // every rank's contribution is complete
  if ((rc = launch_reduce(ctx, pp, pp, R, lo, hi, IdentityF()))) return rc;
  if ((rc = nimfm_peer_barrier(ctx))) return rc;        // every slice has landed everywhere
  *done = 1;
  return NIMFM_OK;
}

int nimfm_peer_mbpsgd_step(nimfm_ctx *ctx, double *pool, double *grad, int64_t lo, int64_t hi, const MbpsgdStepArgs &sa,
                           int broadcast, int *done) {
  *done = 0;
  PeerPtrs pg, pp;
  const int64_t span = sa.nP + sa.d + 2;
  if (!ctx->peerOK || (lo & 1) || !lookup(ctx, grad, span, &pg) || !lookup(ctx, pool, span, &pp)) return NIMFM_OK;
  int rc;
  if ((rc = nimfm_peer_barrier(ctx))) return rc;
  MbpsgdStepF f{pool, sa};
  PeerPtrs out = pp;
  if (!broadcast) out.p[0] = pool;                      // a prox follows on the local slice: broadcast afterwards
  if ((rc = launch_reduce(ctx, pg, out, broadcast ? ctx->nranks : 1, lo, hi, f))) return rc;
  *done = 1;
  return NIMFM_OK;
}

int nimfm_peer_broadcast_slice(nimfm_ctx *ctx, double *buf, int64_t lo, int64_t hi) {
  PeerPtrs pp;
  if (!ctx->peerOK || !lookup(ctx, buf, hi, &pp)) return nimfm_fail(ctx, NIMFM_ERR_STATE, "buffer is not peer-mapped");
  if (hi <= lo) return NIMFM_OK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((hi - lo + 255) / 256, (int64_t)ctx->numSMs * 4));
  peer_broadcast_kernel<<<grid, 256, 0, ctx->stream>>>(buf, pp, ctx->nranks, ctx->rank, lo, hi);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  return NIMFM_OK;
}

// peer.cu -- the gradient exchange of the row-sharded solvers over NVLink / NVSwitch peer memory
// (one process per GPU; the reference has no distributed backend at all -- SURVEY 2a -- this replaces the
// "ncclAllReduce, then the identical dense step on every GPU" schedule of SURVEY 8e).
//
// Every buffer that takes part in the exchange (gradient pool, parameter pool, AdaGrad delta block, touch
// counts) lives in an ARENA; at the start of an epoch-level call the ranks swap the CUDA IPC handles of the
// arenas the call will exchange (nimfm_peer_prepare: one small gather, mappings cached), so a kernel on rank r
// can read and write the corresponding buffer of every peer directly.  One exchange is
//     barrier  ->  reduce kernel  ->  barrier
// followed by  pull -> [barrier],  where the reduce kernel, on rank r, walks ITS 1/N slice of the buffer: 16-byte
// loads of the slice out of all N ranks' buffers (N loads in flight per element pair), a sum in rank order
// 0..N-1 (so the result is bit-identical on every rank and from run to run), an element-wise functor -- identity
// for a plain all-reduce; Params.step + the L1 prox for MBPSGD, which turns reduce-scatter + sharded step into
// ONE pass over 1/N of the parameters -- and a LOCAL store; the pull kernel then copies every peer's finished
// slice into the local buffer (remote loads, local stores).  Per rank that reads (N-1)/N of the buffer twice over
// NVLink, the bytes of a reduce-scatter + all-gather, with no staging copies and 1/N of the dense step.
// Pull (remote loads, local stores) rather than push: both were built and measured on 2 GPUs
// (profiles/r02_peer_exchange.md) -- the exchange costs the same either way (~590-620 GB/s of NVLink reads),
// and with every store local no rank's memory is written behind its back while it computes.
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

// the prepared table of the running epoch call: local range -> every rank's pointer to ITS buffer of that role
bool lookup(const nimfm_ctx *ctx, const double *ptr, int64_t n, PeerPtrs *out) {
  for (const auto &m : ctx->peerMaps) {
    if (ptr >= m.local && ptr + n <= m.local + m.nDoubles) {
      const int64_t off = ptr - m.local;
      for (int r = 0; r < ctx->nranks; r++) out->p[r] = m.peer[r] + off;
      return true;
    }
  }
  return false;
}

// open (or find in the cache) rank r's allocation with this IPC handle
double *open_peer(nimfm_ctx *ctx, int r, const int64_t *handle8) {
  for (const auto &o : ctx->peerOpened)
    if (o.rank == r && !memcmp(o.handle, handle8, 64)) return o.ptr;
  cudaIpcMemHandle_t hp;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(&hp, handle8, sizeof(hp));
  void *p = nullptr;
  if (cudaIpcOpenMemHandle(&p, hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  nimfm_ctx::PeerOpened o;
  o.rank = r;
  memcpy(o.handle, handle8, 64);
  o.ptr = static_cast<double *>(p);
  ctx->peerOpened.push_back(o);
  return o.ptr;
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
  // the flag array: exchanged once, here (comm_init is collective by definition)
  int64_t mine[9] = {0};
  cudaIpcMemHandle_t h;
  if (cudaMalloc(&ctx->peerFlags, 4096) == cudaSuccess && cudaMemset(ctx->peerFlags, 0, 4096) == cudaSuccess &&
      cudaDeviceSynchronize() == cudaSuccess && cudaIpcGetMemHandle(&h, ctx->peerFlags) == cudaSuccess) {
    memcpy(mine, &h, sizeof(h));
    mine[8] = 1;
  } else {
    cudaGetLastError();
  }
  const int R = ctx->nranks;
  std::vector<int64_t> all((size_t)9 * R);
  int rc = nimfm_allgather_host_i64(ctx, mine, 9, all.data());
  if (rc) return rc;
  int64_t opened = 1;
  for (int r = 0; r < R; r++) opened = opened && all[(size_t)9 * r + 8];
  if (opened)
    for (int r = 0; r < R; r++) {
      ctx->peerFlagsOf[r] = r == ctx->rank ? ctx->peerFlags : reinterpret_cast<uint32_t *>(open_peer(ctx, r, &all[(size_t)9 * r]));
      if (!ctx->peerFlagsOf[r]) opened = 0;
    }
  std::vector<int64_t> sts((size_t)R);
  if ((rc = nimfm_allgather_host_i64(ctx, &opened, 1, sts.data()))) return rc;   // did EVERY rank map every peer?
  bool every = true;
  for (int r = 0; r < R; r++) every = every && sts[(size_t)r] != 0;
  ctx->peerOK = every;   // false: e.g. no P2P between the GPUs of this box -- the NCCL route stays in charge
  ctx->barrierEpoch = 0;
  return NIMFM_OK;
}

void nimfm_peer_shutdown(nimfm_ctx *ctx) {
  for (auto &o : ctx->peerOpened) cudaIpcCloseMemHandle(o.ptr);
  ctx->peerOpened.clear();
  ctx->peerMaps.clear();
  for (auto &a : ctx->arenas) cudaFree(a.base);
  ctx->arenas.clear();
  if (ctx->peerFlags) cudaFree(ctx->peerFlags);
  ctx->peerFlags = nullptr;
  ctx->peerOK = false;
}

// Buffers that may take part in an exchange.  LOCAL (not collective: a rank may create models the others do not,
// e.g. CD on rank 0).  With peer memory on, the memory comes from arenas that are recycled by size and released
// only with the context: an exporter must not free memory a peer may still have mapped.
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
  CK(cudaMalloc(&a.base, nDoubles * 8));
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

// COLLECTIVE, once per epoch-level library call: every rank names the buffers the call will exchange (same
// roles in the same order on every rank); the ranks swap the IPC handles of the arenas those buffers live in and
// (re)open what they have not mapped yet.  A buffer that is not arena memory on some rank, or that some rank
// cannot map, is left to NCCL by ALL ranks (everybody sees the same gathered table).  The table is valid until
// nimfm_peer_release -- the exchanges below find their peers' pointers in it.
int nimfm_peer_prepare(nimfm_ctx *ctx, const double *const *bufs, int nb) {
  ctx->peerMaps.clear();
  if (!ctx->peerOK || nb <= 0) return NIMFM_OK;
  const int R = ctx->nranks;
  const int W = 11;   // per buffer: handle (8) | offset in the arena | doubles available from the offset | ok
  std::vector<int64_t> mine((size_t)W * nb, 0), all((size_t)W * nb * R);
  for (int i = 0; i < nb; i++) {
    for (auto &a : ctx->arenas)
      if (bufs[i] && bufs[i] >= a.base && bufs[i] < a.base + a.nDoubles) {
        if (!a.haveHandle) {
          cudaIpcMemHandle_t h;
          if (cudaIpcGetMemHandle(&h, a.base) != cudaSuccess) { cudaGetLastError(); break; }
          memcpy(a.handle, &h, 64);
          a.haveHandle = true;
        }
        memcpy(&mine[(size_t)W * i], a.handle, 64);
        mine[(size_t)W * i + 8] = bufs[i] - a.base;
        mine[(size_t)W * i + 9] = (int64_t)a.nDoubles - (bufs[i] - a.base);
        mine[(size_t)W * i + 10] = 1;
        break;
      }
  }
  int rc = nimfm_allgather_host_i64(ctx, mine.data(), W * nb, all.data());
  if (rc) return rc;
  std::vector<nimfm_ctx::PeerMap> maps;
  std::vector<int64_t> good((size_t)nb, 1);
  for (int i = 0; i < nb; i++) {
    nimfm_ctx::PeerMap m;
    m.local = const_cast<double *>(bufs[i]);
    m.nDoubles = mine[(size_t)W * i + 9];
    for (int r = 0; r < R && good[(size_t)i]; r++) {
      const int64_t *q = &all[((size_t)r * nb + i) * W];
      if (!q[10]) { good[(size_t)i] = 0; break; }
      m.nDoubles = std::min(m.nDoubles, q[9]);
      if (r == ctx->rank) { m.peer[r] = m.local; continue; }
      double *base = open_peer(ctx, r, q);
      if (!base) { good[(size_t)i] = 0; break; }
      m.peer[r] = base + q[8];
    }
    maps.push_back(m);
  }
  // second round: a buffer is exchanged over peer memory only if EVERY rank mapped every peer's copy
  std::vector<int64_t> goodAll((size_t)nb * R);
  if ((rc = nimfm_allgather_host_i64(ctx, good.data(), nb, goodAll.data()))) return rc;
  for (int i = 0; i < nb; i++) {
    bool every = true;
    for (int r = 0; r < R; r++) every = every && goodAll[(size_t)r * nb + i] != 0;
    if (every) ctx->peerMaps.push_back(maps[(size_t)i]);
  }
  return NIMFM_OK;
}

void nimfm_peer_release(nimfm_ctx *ctx) { ctx->peerMaps.clear(); }

int nimfm_peer_pull_slices(nimfm_ctx *ctx, double *buf, int64_t c, int64_t n);

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
  PeerPtrs out = pp;
  out.p[0] = buf;
  if ((rc = launch_reduce(ctx, pp, out, 1, lo, hi, IdentityF()))) return rc;   // my slice, summed in rank order, in place
  if ((rc = nimfm_peer_barrier(ctx))) return rc;        // every slice is final
  if ((rc = nimfm_peer_pull_slices(ctx, buf, c, n))) return rc;
  if ((rc = nimfm_peer_barrier(ctx))) return rc;        // nobody still reads my slice: the caller may overwrite the buffer
  *done = 1;
  return NIMFM_OK;
}

int nimfm_peer_mbpsgd_step(nimfm_ctx *ctx, double *pool, double *grad, int64_t lo, int64_t hi, const MbpsgdStepArgs &sa,
                           int *done) {
  *done = 0;
  PeerPtrs pg, pp;
  const int64_t span = sa.nP + sa.d + 2;
  if (!ctx->peerOK || (lo & 1) || !lookup(ctx, grad, span, &pg) || !lookup(ctx, pool, span, &pp)) return NIMFM_OK;
  int rc;
  if ((rc = nimfm_peer_barrier(ctx))) return rc;   // every rank's gradients are complete (and its last pull has finished)
  MbpsgdStepF f{pool, sa};
  PeerPtrs out = pp;
  out.p[0] = pool;
  if ((rc = launch_reduce(ctx, pg, out, 1, lo, hi, f))) return rc;
  *done = 1;
  return NIMFM_OK;
}

// all-gather by PULLING: copy every peer's slice [r*c, (r+1)*c) of its buffer into the local buffer (remote
// 16-byte loads, local stores), the peers in a rotated order so that the ranks do not all read the same GPU
__global__ void __launch_bounds__(256) peer_pull_kernel(PeerPtrs in, double *dst, int R, int self, int64_t c, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int q = 1; q < R; q++) {
    const int r = (self + q) % R;
    const int64_t lo = r * c, hi = lo + c < n ? lo + c : n;      // c even, n may be odd
    const double2 *src = reinterpret_cast<const double2 *>(in.p[r]);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    const int64_t p0 = lo >> 1, p1 = hi >> 1;
    for (int64_t i = p0 + tid; i < p1; i += 4 * stride) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (i + u * stride < p1) v[u] = __ldcg(src + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (i + u * stride < p1) d2[i + u * stride] = v[u];
    }
    if ((hi & 1) && hi > lo && tid == 0) dst[hi - 1] = __ldcg(in.p[r] + hi - 1);
  }
}

int nimfm_peer_pull_slices(nimfm_ctx *ctx, double *buf, int64_t c, int64_t n) {
  PeerPtrs pp;
  if (!ctx->peerOK || !lookup(ctx, buf, n, &pp)) return nimfm_fail(ctx, NIMFM_ERR_STATE, "buffer is not peer-mapped");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((c / 2 + 1023) / 1024, (int64_t)ctx->numSMs * 4));
  peer_pull_kernel<<<grid, 256, 0, ctx->stream>>>(pp, buf, ctx->nranks, ctx->rank, c, n);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  return NIMFM_OK;
}

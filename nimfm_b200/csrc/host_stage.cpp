// host_stage.cpp -- see host_stage.h.  Plain C++ (g++), no CUDA.
#include "host_stage.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace {

// int64 -> int32 with a range check; `dst` 32-byte aligned.  Returns nonzero if any id is outside [0,d).
int narrow_scalar(const int64_t *src, int32_t *dst, int64_t n, int64_t d) {
  int bad = 0;
  for (int64_t i = 0; i < n; i++) {
    const int64_t v = src[i];
    bad |= (v < 0) | (v >= d);
    dst[i] = (int32_t)v;
  }
  return bad;
}

// AVX2: 8 ids per iteration, non-temporal stores (the copy engine is the only reader of the slot).
// Range check: ids are valid iff 0 <= v < d; with d <= 2^31 that is "high word zero and low word < d as
// unsigned", folded into one OR-accumulator of (v | (d-1-v)) whose sign bit flags a violation.
__attribute__((target("avx2"))) int narrow_avx2(const int64_t *src, int32_t *dst, int64_t n, int64_t d) {
  const __m256i perm = _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0);
  const __m256i dm1 = _mm256_set1_epi64x(d - 1);
  __m256i acc = _mm256_setzero_si256();
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 4));
    acc = _mm256_or_si256(acc, _mm256_or_si256(a, _mm256_sub_epi64(dm1, a)));
    acc = _mm256_or_si256(acc, _mm256_or_si256(b, _mm256_sub_epi64(dm1, b)));
    const __m128i lo = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a, perm));
    const __m128i hi = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(b, perm));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_set_m128i(hi, lo));
  }
  alignas(32) int64_t t[4];
  _mm256_store_si256(reinterpret_cast<__m256i *>(t), acc);
  int bad = (t[0] | t[1] | t[2] | t[3]) < 0;
  bad |= narrow_scalar(src + i, dst + i, n - i, d);
  _mm_sfence();
  return bad;
}

// doubles -> a pinned slot with non-temporal stores (the copy engine is the slot's only reader); dst 32-byte aligned
__attribute__((target("avx2"))) void stream_copy_avx2(double *dst, const double *src, int64_t n) {
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
    _mm256_stream_pd(dst + i, a);
    _mm256_stream_pd(dst + i + 4, b);
  }
  for (; i < n; i++) dst[i] = src[i];
  _mm_sfence();
}

void stream_copy(double *dst, const double *src, int64_t n) {
  static const bool haveAvx2 = __builtin_cpu_supports("avx2");
  if (haveAvx2 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) stream_copy_avx2(dst, src, n);
  else memcpy(dst, src, (size_t)n * sizeof(double));
}

// Pack a slice of values (n a multiple of 256 except for the last slice of a chunk): mask bit = "value is exactly 1.0",
// the other values appended to `out` (capacity n), blk[b] = outBase + number of packed values before block b of 256.
// Returns the number of packed values.
int64_t pack_scalar(const double *src, int64_t n, double *out, uint64_t *mask, uint32_t *blk, int64_t outBase, int64_t i0,
                    int64_t cnt) {
  for (int64_t i = i0; i < n; i++) {
    if ((i & 255) == 0) blk[i >> 8] = (uint32_t)(outBase + cnt);
    if ((i & 63) == 0) mask[i >> 6] = 0;
    const double v = src[i];
    const bool one = v == 1.0;
    mask[i >> 6] |= (uint64_t)one << (i & 63);
    out[cnt] = v;
    cnt += !one;
  }
  return cnt;
}

__attribute__((target("avx2,popcnt"))) int64_t pack_avx2(const double *src, int64_t n, double *out, uint64_t *mask,
                                                        uint32_t *blk, int64_t outBase) {
  // left-compaction of 4 doubles by a 16-entry table of 32-bit lane permutations (keep bit i = value i is NOT 1.0)
  alignas(32) static const int32_t lut[16][8] = {
      {0, 1, 2, 3, 4, 5, 6, 7}, {0, 1, 2, 3, 4, 5, 6, 7}, {2, 3, 0, 1, 4, 5, 6, 7}, {0, 1, 2, 3, 4, 5, 6, 7},
      {4, 5, 0, 1, 2, 3, 6, 7}, {0, 1, 4, 5, 2, 3, 6, 7}, {2, 3, 4, 5, 0, 1, 6, 7}, {0, 1, 2, 3, 4, 5, 6, 7},
      {6, 7, 0, 1, 2, 3, 4, 5}, {0, 1, 6, 7, 2, 3, 4, 5}, {2, 3, 6, 7, 0, 1, 4, 5}, {0, 1, 2, 3, 6, 7, 4, 5},
      {4, 5, 6, 7, 0, 1, 2, 3}, {0, 1, 4, 5, 6, 7, 2, 3}, {2, 3, 4, 5, 6, 7, 0, 1}, {0, 1, 2, 3, 4, 5, 6, 7}};
  const __m256d ones = _mm256_set1_pd(1.0);
  int64_t cnt = 0, i = 0;
  const int64_t n64 = n & ~(int64_t)63;
  for (; i < n64; i += 64) {
    if ((i & 255) == 0) blk[i >> 8] = (uint32_t)(outBase + cnt);
    uint64_t word = 0;
    for (int g = 0; g < 16; g++) {
      const __m256d v = _mm256_loadu_pd(src + i + 4 * g);
      const int eq = _mm256_movemask_pd(_mm256_cmp_pd(v, ones, _CMP_EQ_OQ));
      word |= (uint64_t)eq << (4 * g);
      const int keep = ~eq & 15;
      const __m256i perm = _mm256_load_si256(reinterpret_cast<const __m256i *>(lut[keep]));
      _mm256_storeu_pd(out + cnt, _mm256_castsi256_pd(_mm256_permutevar8x32_epi32(_mm256_castpd_si256(v), perm)));
      cnt += __builtin_popcount(keep);
    }
    mask[i >> 6] = word;
  }
  return pack_scalar(src, n, out, mask, blk, outBase, i, cnt);
}

// AVX-512: the compare yields the mask byte directly and VCOMPRESSPD (register form) does the left-compaction
__attribute__((target("avx512f,popcnt"))) int64_t pack_avx512(const double *src, int64_t n, double *out, uint64_t *mask,
                                                              uint32_t *blk, int64_t outBase) {
  const __m512d ones = _mm512_set1_pd(1.0);
  int64_t cnt = 0, i = 0;
  const int64_t n64 = n & ~(int64_t)63;
  for (; i < n64; i += 64) {
    if ((i & 255) == 0) blk[i >> 8] = (uint32_t)(outBase + cnt);
    uint64_t word = 0;
#pragma GCC unroll 8
    for (int g = 0; g < 8; g++) {
      const __m512d v = _mm512_loadu_pd(src + i + 8 * g);
      const __mmask8 eq = _mm512_cmp_pd_mask(v, ones, _CMP_EQ_OQ);
      word |= (uint64_t)eq << (8 * g);
      const __mmask8 keep = (__mmask8)~eq;
      _mm512_storeu_pd(out + cnt, _mm512_maskz_compress_pd(keep, v));   // the tail is overwritten by the next group
      cnt += __builtin_popcount((unsigned)keep);
    }
    mask[i >> 6] = word;
  }
  return pack_scalar(src, n, out, mask, blk, outBase, i, cnt);
}

// AVX-512: 16 ids per iteration through VPMOVQD, same sign-bit range check as the AVX2 form
__attribute__((target("avx512f,avx512dq"))) int narrow_avx512(const int64_t *src, int32_t *dst, int64_t n, int64_t d) {
  const __m512i dm1 = _mm512_set1_epi64(d - 1);
  __m512i acc = _mm512_setzero_si512();
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m512i a = _mm512_loadu_si512(src + i), b = _mm512_loadu_si512(src + i + 8);
    acc = _mm512_or_si512(acc, _mm512_or_si512(_mm512_or_si512(a, _mm512_sub_epi64(dm1, a)),
                                               _mm512_or_si512(b, _mm512_sub_epi64(dm1, b))));
    const __m512i both = _mm512_inserti64x4(_mm512_castsi256_si512(_mm512_cvtepi64_epi32(a)), _mm512_cvtepi64_epi32(b), 1);
    _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), both);
  }
  int bad = _mm512_movepi64_mask(acc) != 0;
  bad |= narrow_scalar(src + i, dst + i, n - i, d);
  _mm_sfence();
  return bad;
}

bool use_avx512() {
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") &&
                         __builtin_cpu_supports("popcnt") && !(getenv("NIMFM_HOST_AVX512") && getenv("NIMFM_HOST_AVX512")[0] == '0');
  return ok;
}

int64_t pack(const double *src, int64_t n, double *out, uint64_t *mask, uint32_t *blk, int64_t outBase) {
  static const bool haveAvx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt");
  if (use_avx512()) return pack_avx512(src, n, out, mask, blk, outBase);
  return haveAvx2 ? pack_avx2(src, n, out, mask, blk, outBase) : pack_scalar(src, n, out, mask, blk, outBase, 0, 0);
}

int narrow(const int64_t *src, int32_t *dst, int64_t n, int64_t d) {
  static const bool haveAvx2 = __builtin_cpu_supports("avx2");
  if (use_avx512() && (reinterpret_cast<uintptr_t>(dst) & 63) == 0) return narrow_avx512(src, dst, n, d);
  return haveAvx2 ? narrow_avx2(src, dst, n, d) : narrow_scalar(src, dst, n, d);
}

}  // namespace

int nimfm_host_narrow(const int64_t *src, int32_t *dst, int64_t n, int64_t d) { return narrow(src, dst, n, d); }

int HostStageTeam::default_threads(int nRanks) {
  if (const char *e = getenv("NIMFM_HOST_THREADS")) return std::max(0, atoi(e));
  const int hw = (int)std::thread::hardware_concurrency();
  // Staging pays while the rank's own PCIe link is the bound (1-2 ranks on this box: 81 -> 102 M rows/s).
  // With every GPU of the box fed at once the host's memory system is the bound and the extra pass over
  // the ids costs more than the 25 % of link bytes it saves (8 ranks, measured: 285 M rows/s narrowing on the
  // device, 188-196 M with 2-4 staging threads per rank): only ranks with 8 hardware threads to themselves stage.
  const int t = hw / std::max(1, nRanks);
  return t >= 8 ? 8 : 0;
}

// pageable caller arrays add a copy of the values to the narrowing: the team is the bound then (8 threads: 32 GB/s
// of link traffic against 52 GB/s from pinned arrays), so it takes what the rank has, less two for the caller
int HostStageTeam::pageable_threads(int nRanks) {
  if (const char *e = getenv("NIMFM_HOST_THREADS")) return std::max(0, atoi(e));
  const int hw = (int)std::thread::hardware_concurrency();
  const int t = hw / std::max(1, nRanks);
  if (t >= 8) return std::min(t - 2, 24);
  // thread-starved ranks (8 ranks on 16 cores): a team of the rank's own cores still beats the driver's pageable
  // path (one thread, one bounce buffer, ~6 GB/s per rank); NIMFM_HOST_PAGEABLE_SMALL=0 restores that path
  const char *e = getenv("NIMFM_HOST_PAGEABLE_SMALL");
  return (e && e[0] == '0') ? 0 : std::max(1, t);
}

HostStageTeam::HostStageTeam(int nThreads, const int64_t *indices, const int64_t *indptr, int64_t d,
                             const std::vector<HostChunk> &chunks, int32_t *const *idxSlot, int64_t *const *ptrSlot)
    : T_(std::min(32, std::max(1, nThreads))), indices_(indices), indptr_(indptr), d_(d), chunks_(chunks), idxSlot_(idxSlot),
      ptrSlot_(ptrSlot), done_(chunks.size(), 0), info_(chunks.size()) {
  for (int t = 0; t < T_; t++) threads_.emplace_back([this, t] { work(t); });
}

HostStageTeam::~HostStageTeam() {
  {
    std::lock_guard<std::mutex> g(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto &th : threads_) th.join();
}

void HostStageTeam::stage_values(const double *data, const double *y, double *const *dataSlot, double *const *ySlot) {
  std::lock_guard<std::mutex> g(mu_);
  data_ = data;
  y_ = y;
  dataSlot_ = dataSlot;
  ySlot_ = ySlot;
}

void HostStageTeam::pack_values(const double *data, const double *y, double *const *packSlot, uint64_t *const *maskSlot,
                                uint32_t *const *blkSlot, double *const *ySlot) {
  std::lock_guard<std::mutex> g(mu_);
  data_ = data;
  y_ = y;
  dataSlot_ = nullptr;
  packSlot_ = packSlot;
  maskSlot_ = maskSlot;
  blkSlot_ = blkSlot;
  ySlot_ = ySlot;
}

void HostStageTeam::allow(int64_t upTo) {
  {
    std::lock_guard<std::mutex> g(mu_);
    allowed_ = std::max(allowed_, upTo);
  }
  cv_.notify_all();
}

HostChunkInfo HostStageTeam::wait(int64_t c) {
  std::unique_lock<std::mutex> g(mu_);
  cv_.wait(g, [&] { return done_[c] == T_; });
  return info_[c];
}

// thread t's share of chunk c: an 8-aligned slice of the nonzeros and a slice of the rows
void HostStageTeam::part(int64_t c, int t, HostChunkInfo &info) {
  const HostChunk &ch = chunks_[c];
  const int s = (int)(c % kSlots);
  const int64_t per = slice_len(ch.nnz);
  const int64_t a = std::min(ch.nnz, per * t), b = std::min(ch.nnz, per * (t + 1));
  if (b > a) info.bad |= narrow(indices_ + ch.base + a, idxSlot_[s] + a, b - a, d_);
  if (b > a && data_ && packSlot_)
    info.packed[t] = pack(data_ + ch.base + a, b - a, packSlot_[s] + a, maskSlot_[s] + (a >> 6), blkSlot_[s] + (a >> 8), a);
  else if (b > a && data_)
    stream_copy(dataSlot_[s] + a, data_ + ch.base + a, b - a);
  const int64_t rows = ch.r1 - ch.r0, rper = (rows + T_ - 1) / T_;
  const int64_t ra = std::min(rows, rper * t), rb = std::min(rows, rper * (t + 1));
  if (rb > ra && y_) memcpy(ySlot_[s] + ra, y_ + ch.r0 + ra, (size_t)(rb - ra) * sizeof(double));
  int64_t *op = ptrSlot_[s];
  for (int64_t r = ra; r < rb; r++) {
    const int64_t p0 = indptr_[ch.r0 + r], len = indptr_[ch.r0 + r + 1] - p0;
    op[r] = p0 - ch.base;
    info.maxSeg = std::max(info.maxSeg, len);
    info.minSeg = std::min(info.minSeg, len);
  }
  if (t == T_ - 1) op[rows] = ch.nnz;
}

void HostStageTeam::work(int t) {
  for (int64_t c = 0; c < (int64_t)chunks_.size(); c++) {
    {
      std::unique_lock<std::mutex> g(mu_);
      cv_.wait(g, [&] { return stop_ || allowed_ > c; });
      if (stop_) return;
    }
    HostChunkInfo mine;
    part(c, t, mine);
    bool last;
    {
      std::lock_guard<std::mutex> g(mu_);
      info_[c].bad |= mine.bad;
      info_[c].packed[t] = mine.packed[t];
      info_[c].maxSeg = std::max(info_[c].maxSeg, mine.maxSeg);
      info_[c].minSeg = std::min(info_[c].minSeg, mine.minSeg);
      last = ++done_[c] == T_;
    }
    if (last) cv_.notify_all();
  }
}

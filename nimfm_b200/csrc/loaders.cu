// loaders.cu -- text loaders feeding the path: svmlight/libsvm, libffm and
// user-item-rating files parsed on the host by all cores and uploaded straight into device datasets.
//   loadSVMLightFile        dataset.nim:562-693   "label j:val j:val ..."   (CSR, or CSC via the stable transpose)
//   loadFFMFile             dataset.nim:696-790   "label field:j:val ..."
//   loadUserItemRatingFile  dataset.nim:840-990   "user<sep>item<sep>rating[<sep>comment]" -> one-hot user+item
// Index conventions follow the reference: indices are assumed 1-based unless a 0 occurs
// (offset = 0 iff min index == 0, :585-586); nFeatures = max index + 1 - offset, widened by the
// caller's nFeatures (:623-632); a negative index raises "Negative index is included.".
// The file is split at line boundaries into one slice per thread; slices are parsed independently and
// concatenated in order, so the result is identical to a sequential read.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <thread>

#include "common.cuh"

namespace {

enum { FMT_SVM = 0, FMT_FFM = 1, FMT_UIR = 2 };

struct Slice {
  std::vector<double> data, y;
  std::vector<int64_t> indices, fields, rowLen;
  int64_t minIdx = 1, maxIdx = 0, minFld = 1, maxFld = 1;   // initial values as dataset.nim:569-570,705-708
  int64_t minItem = 1, maxItem = 0;                          // user-item files: indices hold (user, item)
  int64_t badLine = -1;                                      // slice-local number of the first malformed line
};

inline const char *skip_blank(const char *p, const char *e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
  return p;
}

inline bool parse_i64(const char *&p, const char *e, int64_t &v) {
  if (p < e && *p == '+') ++p;
  auto r = std::from_chars(p, e, v);
  if (r.ec != std::errc()) return false;
  p = r.ptr;
  return true;
}

inline bool parse_f64(const char *&p, const char *e, double &v) {
  if (p < e && *p == '+') ++p;
  auto r = std::from_chars(p, e, v);
  if (r.ec != std::errc()) return false;
  p = r.ptr;
  return true;
}

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

void parse_slice(const char *b, const char *e, int fmt, Slice &s) {
  int64_t lineNo = 0;
  while (b < e) {
    const char *le = static_cast<const char *>(memchr(b, '\n', (size_t)(e - b)));
    if (!le) le = e;
    const char *p = skip_blank(b, le);
    ++lineNo;
    if (fmt == FMT_UIR) {
      if (le - b >= 5) {   // dataset.nim:869-870 skips shorter lines
        int64_t user = 0, item = 0;
        double rating = 0.0;
        while (p < le && !is_digit(*p)) ++p;                     // skipUntil(line, Numbers)
        bool ok = parse_i64(p, le, user);
        while (p < le && !is_digit(*p)) ++p;
        ok = ok && parse_i64(p, le, item);
        while (p < le && !is_digit(*p)) ++p;
        ok = ok && parse_f64(p, le, rating);
        if (!ok) { if (s.badLine < 0) s.badLine = lineNo; }
        else {
          s.indices.push_back(user);
          s.indices.push_back(item);
          s.y.push_back(rating);
          s.minIdx = std::min(s.minIdx, user); s.maxIdx = std::max(s.maxIdx, user);
          s.minItem = std::min(s.minItem, item); s.maxItem = std::max(s.maxItem, item);
        }
      }
    } else if (p < le) {   // blank lines carry no sample
      double target = 0.0;
      bool ok = parse_f64(p, le, target);
      int64_t len = 0;
      while (ok) {
        p = skip_blank(p, le);
        if (p >= le) break;
        int64_t field = 0, j = 0;
        double val = 0.0;
        if (fmt == FMT_FFM) {
          ok = parse_i64(p, le, field) && p < le && *p == ':';
          if (!ok) break;
          ++p;
        }
        ok = parse_i64(p, le, j) && p < le && *p == ':';
        if (!ok) break;
        ++p;
        ok = parse_f64(p, le, val);
        if (!ok) break;
        if (fmt == FMT_FFM) {
          s.fields.push_back(field);
          s.minFld = std::min(s.minFld, field); s.maxFld = std::max(s.maxFld, field);
        }
        s.indices.push_back(j);
        s.data.push_back(val);
        s.minIdx = std::min(s.minIdx, j); s.maxIdx = std::max(s.maxIdx, j);
        ++len;
      }
      if (!ok && s.badLine < 0) s.badLine = lineNo;
      s.y.push_back(target);
      s.rowLen.push_back(len);
    }
    b = le + 1;
  }
}

struct Mapped {
  const char *p = nullptr;
  size_t n = 0;
  int fd = -1;
  ~Mapped() {
    if (p && n) munmap(const_cast<char *>(p), n);
    if (fd >= 0) close(fd);
  }
};

int load_text(nimfm_ctx *ctx, const char *path, int fmt, std::vector<Slice> &slices) {
  REQUIRE(path != nullptr, "path is NULL");
  Mapped m;
  m.fd = open(path, O_RDONLY);
  if (m.fd < 0) return nimfm_fail(ctx, NIMFM_ERR_INVALID, "cannot open %s", path);   // IOError in the reference
  struct stat st;
  if (fstat(m.fd, &st) != 0) return nimfm_fail(ctx, NIMFM_ERR_INVALID, "cannot stat %s", path);
  m.n = (size_t)st.st_size;
  if (m.n == 0) { slices.assign(1, Slice()); return NIMFM_OK; }
  void *mp = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
  if (mp == MAP_FAILED) { m.n = 0; return nimfm_fail(ctx, NIMFM_ERR_INVALID, "cannot map %s", path); }
  m.p = static_cast<const char *>(mp);
  unsigned hw = std::thread::hardware_concurrency();
  size_t T = std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)64, m.n / (1 << 20) + 1}));
  std::vector<const char *> cut(T + 1);
  cut[0] = m.p;
  cut[T] = m.p + m.n;
  for (size_t t = 1; t < T; t++) {
    const char *q = m.p + m.n * t / T;
    if (q < cut[t - 1]) q = cut[t - 1];
    const char *nl = static_cast<const char *>(memchr(q, '\n', (size_t)(m.p + m.n - q)));
    cut[t] = nl ? nl + 1 : m.p + m.n;
  }
  slices.assign(T, Slice());
  std::vector<std::thread> th;
  for (size_t t = 1; t < T; t++) th.emplace_back(parse_slice, cut[t], cut[t + 1], fmt, std::ref(slices[t]));
  parse_slice(cut[0], cut[1], fmt, slices[0]);
  for (auto &x : th) x.join();
  int64_t lines = 0;
  for (size_t t = 0; t < T; t++) {
    if (slices[t].badLine >= 0)
      return nimfm_fail(ctx, NIMFM_ERR_INVALID, "%s: malformed line (line %lld of slice %zu, after ~%lld samples)", path,
                        (long long)slices[t].badLine, t, (long long)lines);
    lines += (int64_t)slices[t].y.size();
  }
  return NIMFM_OK;
}

int finish_upload(nimfm_ctx *ctx, int64_t n, int64_t d, std::vector<double> &data, std::vector<int64_t> &indices,
                  std::vector<int64_t> &indptr, std::vector<int64_t> *fields, int64_t nFields,
                  const std::vector<double> &y, int asCsc, nimfm_dataset **out) {
  nimfm_dataset *csr = nullptr;
  int rc = nimfm_csr_upload(ctx, n, d, data.data(), indices.data(), indptr.data(), fields ? fields->data() : nullptr,
                            nFields, 0, n, &csr);
  if (rc) return rc;
  if (n > 0 && (rc = nimfm_dataset_set_targets(ctx, csr, y.data()))) { nimfm_dataset_free(ctx, csr); return rc; }
  if (!asCsc) { *out = csr; return NIMFM_OK; }
  // the CSC loaders (dataset.nim:643-686, 903-990) place entries column by column in row order:
  // exactly the stable transpose
  nimfm_dataset *csc = nullptr;
  rc = nimfm_dataset_transpose(ctx, csr, &csc);
  nimfm_dataset_free(ctx, csr);
  if (rc) return rc;
  *out = csc;
  return NIMFM_OK;
}

}  // namespace

// ------------------------------------------------------------------ STREAMCSR / STREAMCSC binary files
// (tensor/sparse_stream.nim:3-33, written by convertSVMLightFile / transposeFile, dataset.nim:1017-1200):
//   magic "STREAMCSR" | "STREAMCSC" (9 bytes), header {nRows, nCols, nnz: int64; max, min: float64},
//   then per row (CSR) / column (CSC): count int64, count x {val float64, id int64}.
// The reference reads them through a window cache because they may exceed host memory; a B200 holds
// 180 GB, so the file is loaded whole: the host only hops over the counts (O(segments)), the raw
// payload is uploaded as it is and de-interleaved on the device.
struct nimfm_stream {
  Mapped mx, my;
  bool isCsr = true, hasY = false;
  // field files (STREAMCSRFIELD, tensor/sparse_stream.nim:6-8,15-18,27-33): 14-byte magic, 48-byte header with
  // nFields, 24-byte records {field, val, id}; plain files: 9 + 40 bytes, 16-byte records {val, id}
  bool isField = false;
  int64_t nFields = 0;
  int hdrBytes = 49, recBytes = 16, valOff = 0, idOff = 8;
  int64_t nRows = 0, nCols = 0, nnz = 0, nSeg = 0, extent = 0, maxSeg = 0, payloadEnd = 0;
  std::vector<int64_t> segOff;   // byte offset (within the payload) of each segment's first record
  std::vector<int64_t> indptr;
};

namespace {

__global__ void stream_deinterleave_kernel(const unsigned char *payload, const int64_t *segByteOff,
                                           const int64_t *indptr, int64_t nSeg, int64_t extent, double *data,
                                           int32_t *idx, int *bad, int recBytes, int valOff, int idOff,
                                           int32_t *fields, int64_t nFields) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t sgm = warp; sgm < nSeg; sgm += nWarps) {
    const unsigned char *rec = payload + segByteOff[sgm];
    const int64_t b = indptr[sgm], z = indptr[sgm + 1] - b;
    for (int64_t t = lane; t < z; t += 32) {
      const double v = *reinterpret_cast<const double *>(rec + recBytes * t + valOff);
      const int64_t id = *reinterpret_cast<const int64_t *>(rec + recBytes * t + idOff);
      if (id < 0 || id >= extent) *bad = 1;
      data[b + t] = v;
      idx[b + t] = (int32_t)id;
      if (fields) {
        const int64_t f = *reinterpret_cast<const int64_t *>(rec + recBytes * t);
        if (f < 0 || f >= nFields) *bad = 2;
        fields[b + t] = (int32_t)f;
      }
    }
  }
}

}  // namespace

extern "C" {

int32_t nimfm_load_svmlight(nimfm_ctx *ctx, const char *path, int64_t nFeatures, int32_t asCsc, nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr, "out is NULL");
  CK(cudaSetDevice(ctx->device));
  std::vector<Slice> sl;
  int rc = load_text(ctx, path, FMT_SVM, sl);
  if (rc) return rc;
  int64_t n = 0, nnz = 0, minIdx = 1, maxIdx = 0;
  for (auto &s : sl) {
    n += (int64_t)s.y.size(); nnz += (int64_t)s.data.size();
    minIdx = std::min(minIdx, s.minIdx); maxIdx = std::max(maxIdx, s.maxIdx);
  }
  REQUIRE(minIdx >= 0, "Negative index is included.");                       // dataset.nim:583-584
  const int64_t offset = minIdx == 0 ? 0 : 1;                                // :585
  const int64_t nPred = maxIdx + 1 - offset;                                 // :586
  REQUIRE(!(nFeatures > 0 && nPred > nFeatures), "nFeatures is %lld but dataset has at least %lld features.",
          (long long)nFeatures, (long long)nPred);                           // :625-628
  std::vector<double> data((size_t)nnz), y((size_t)n);
  std::vector<int64_t> indices((size_t)nnz), indptr((size_t)n + 1, 0);
  int64_t r = 0, q = 0;
  for (auto &s : sl) {
    for (size_t i = 0; i < s.y.size(); i++, r++) { y[r] = s.y[i]; indptr[r + 1] = indptr[r] + s.rowLen[i]; }
    for (size_t i = 0; i < s.data.size(); i++, q++) { data[q] = s.data[i]; indices[q] = s.indices[i] - offset; }
  }
  return finish_upload(ctx, n, std::max(nPred, nFeatures), data, indices, indptr, nullptr, 0, y, asCsc, out);
}

int32_t nimfm_load_ffm(nimfm_ctx *ctx, const char *path, int64_t nFeatures, int64_t nFields, nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr, "out is NULL");
  CK(cudaSetDevice(ctx->device));
  std::vector<Slice> sl;
  int rc = load_text(ctx, path, FMT_FFM, sl);
  if (rc) return rc;
  int64_t n = 0, nnz = 0, minIdx = 1, maxIdx = 0, minFld = 1, maxFld = 1;
  for (auto &s : sl) {
    n += (int64_t)s.y.size(); nnz += (int64_t)s.data.size();
    minIdx = std::min(minIdx, s.minIdx); maxIdx = std::max(maxIdx, s.maxIdx);
    minFld = std::min(minFld, s.minFld); maxFld = std::max(maxFld, s.maxFld);
  }
  REQUIRE(minIdx >= 0, "Negative index is included.");                       // :732-733
  REQUIRE(minFld >= 0, "Negative field index is included.");
  const int64_t offset = minIdx == 0 ? 0 : 1, offsetField = minFld == 0 ? 0 : 1;   // :734-737
  const int64_t nPred = maxIdx + 1 - offset, nFldPred = maxFld + 1 - offsetField;
  REQUIRE(!(nFields > 0 && nFldPred > nFields), "nFields is %lld but dataset has at least %lld fields.",
          (long long)nFields, (long long)nFldPred);                          // :778-781
  REQUIRE(!(nFeatures > 0 && nPred > nFeatures), "nFeatures is %lld but dataset has at least %lld features.",
          (long long)nFeatures, (long long)nPred);                           // :783-786
  std::vector<double> data((size_t)nnz), y((size_t)n);
  std::vector<int64_t> indices((size_t)nnz), fields((size_t)nnz), indptr((size_t)n + 1, 0);
  int64_t r = 0, q = 0;
  for (auto &s : sl) {
    for (size_t i = 0; i < s.y.size(); i++, r++) { y[r] = s.y[i]; indptr[r + 1] = indptr[r] + s.rowLen[i]; }
    for (size_t i = 0; i < s.data.size(); i++, q++) {
      data[q] = s.data[i]; indices[q] = s.indices[i] - offset; fields[q] = s.fields[i] - offsetField;
    }
  }
  return finish_upload(ctx, n, std::max(nPred, nFeatures), data, indices, indptr, &fields, std::max(nFldPred, nFields),
                       y, 0, out);
}

int32_t nimfm_load_user_item_rating(nimfm_ctx *ctx, const char *path, int32_t asCsc, nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr, "out is NULL");
  CK(cudaSetDevice(ctx->device));
  std::vector<Slice> sl;
  int rc = load_text(ctx, path, FMT_UIR, sl);
  if (rc) return rc;
  int64_t n = 0, minUser = 1, maxUser = 0, minItem = 1, maxItem = 0;
  for (auto &s : sl) {
    n += (int64_t)s.y.size();
    minUser = std::min(minUser, s.minIdx); maxUser = std::max(maxUser, s.maxIdx);
    minItem = std::min(minItem, s.minItem); maxItem = std::max(maxItem, s.maxItem);
  }
  REQUIRE(minUser >= 0, "The minimum user id < 0.");                         // :884-885
  REQUIRE(minItem >= 0, "The minimum item id < 0.");                         // :887-888
  const int64_t nUsers = maxUser - minUser + 1, nItems = maxItem - minItem + 1;   // :889-890
  std::vector<double> data((size_t)n * 2, 1.0), y((size_t)n);
  std::vector<int64_t> indices((size_t)n * 2), indptr((size_t)n + 1);
  int64_t r = 0;
  for (auto &s : sl)
    for (size_t i = 0; i < s.y.size(); i++, r++) {
      y[r] = s.y[i];
      indices[2 * r] = s.indices[2 * i] - minUser;                           // :892-894
      indices[2 * r + 1] = s.indices[2 * i + 1] + nUsers - minItem;
    }
  for (int64_t i = 0; i <= n; i++) indptr[i] = 2 * i;
  return finish_upload(ctx, n, n > 0 ? nUsers + nItems : 0, data, indices, indptr, nullptr, 0, y, asCsc, out);
}

// ---- STREAMCSR / STREAMCSC files: an open handle indexes the segments once; windows of segments become
// resident datasets on demand (the reference's cacheSize window, tensor/sparse_stream.nim:232-270, with HBM
// as the cache), so a file larger than device memory is processed window by window.
int32_t nimfm_stream_open(nimfm_ctx *ctx, const char *pathX, const char *pathY, nimfm_stream **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr && pathX != nullptr, "NULL argument");
  std::unique_ptr<nimfm_stream> sh(new nimfm_stream());
  Mapped &m = sh->mx;
  m.fd = open(pathX, O_RDONLY);
  if (m.fd < 0) return nimfm_fail(ctx, NIMFM_ERR_INVALID, "%s cannot be opened.", pathX);      // sparse_stream.nim:103-104
  struct stat st;
  if (fstat(m.fd, &st) != 0) return nimfm_fail(ctx, NIMFM_ERR_INVALID, "cannot stat %s", pathX);
  m.n = (size_t)st.st_size;
  REQUIRE(m.n >= 49, "%s is not a StreamCSR / StreamCSC file.", pathX);
  void *mp = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
  if (mp == MAP_FAILED) { m.n = 0; return nimfm_fail(ctx, NIMFM_ERR_INVALID, "cannot map %s", pathX); }
  m.p = static_cast<const char *>(mp);
  const bool isCsr = memcmp(m.p, "STREAMCSR", 9) == 0, isCsc = memcmp(m.p, "STREAMCSC", 9) == 0;
  REQUIRE(isCsr || isCsc, "%s is not a StreamCSR / StreamCSC file.", pathX);                    // :109-110,142-143
  sh->isField = m.n >= 62 && memcmp(m.p + 9, "FIELD", 5) == 0;                                  // :164-196
  REQUIRE(!(sh->isField && isCsc), "StreamCSCField files feed no solver on this path (FFM fits are row-wise)");
  int64_t hdr[4] = {0, 0, 0, 0};
  if (sh->isField) {
    memcpy(hdr, m.p + 14, 32);
    sh->nFields = hdr[3];
    sh->hdrBytes = 62; sh->recBytes = 24; sh->valOff = 8; sh->idOff = 16;
    REQUIRE(sh->nFields >= 1, "corrupt header in %s", pathX);
  } else {
    memcpy(hdr, m.p + 9, 24);
  }
  const int HB = sh->hdrBytes, RB = sh->recBytes;
  sh->isCsr = isCsr;
  sh->nRows = hdr[0]; sh->nCols = hdr[1]; sh->nnz = hdr[2];
  REQUIRE(sh->nRows >= 0 && sh->nCols >= 0 && sh->nnz >= 0, "corrupt header in %s", pathX);
  sh->nSeg = isCsr ? sh->nRows : sh->nCols;
  sh->extent = isCsr ? sh->nCols : sh->nRows;
  REQUIRE(sh->extent < (int64_t)2147483647, "index extent %lld does not fit int32", (long long)sh->extent);
  const size_t payloadBytes = m.n - (size_t)HB;
  sh->segOff.resize((size_t)sh->nSeg);
  sh->indptr.assign((size_t)sh->nSeg + 1, 0);
  int64_t off = 0;
  for (int64_t sgm = 0; sgm < sh->nSeg; sgm++) {
    REQUIRE((size_t)off + 8 <= payloadBytes, "%s is truncated (segment %lld)", pathX, (long long)sgm);
    int64_t cnt;
    memcpy(&cnt, m.p + HB + off, 8);
    REQUIRE(cnt >= 0 && (size_t)off + 8 + (size_t)RB * (size_t)cnt <= payloadBytes, "%s is truncated (segment %lld)", pathX,
            (long long)sgm);
    sh->segOff[(size_t)sgm] = off + 8;
    sh->indptr[(size_t)sgm + 1] = sh->indptr[(size_t)sgm] + cnt;
    sh->maxSeg = std::max(sh->maxSeg, cnt);
    off += 8 + (int64_t)RB * cnt;
  }
  sh->payloadEnd = off;
  REQUIRE(sh->indptr[(size_t)sh->nSeg] == sh->nnz, "%s: header nnz %lld != %lld elements found", pathX,
          (long long)sh->nnz, (long long)sh->indptr[(size_t)sh->nSeg]);
  if (pathY && pathY[0]) {            // loadStreamLabel (dataset.nim:995-1014): raw float64 targets
    Mapped &my = sh->my;
    my.fd = open(pathY, O_RDONLY);
    if (my.fd < 0) return nimfm_fail(ctx, NIMFM_ERR_INVALID, "%s cannot be opened.", pathY);
    struct stat sy;
    if (fstat(my.fd, &sy) != 0 || (int64_t)sy.st_size < sh->nRows * 8)
      return nimfm_fail(ctx, NIMFM_ERR_INVALID, "%s holds fewer than %lld labels", pathY, (long long)sh->nRows);
    sh->hasY = true;
  }
  *out = sh.release();
  return NIMFM_OK;
}

int32_t nimfm_stream_info(const nimfm_stream *sh, int32_t *kind, int64_t *nRows, int64_t *nCols, int64_t *nnz,
                          int64_t *maxSegNnz, int64_t *payloadBytes) {
  if (!sh) return NIMFM_ERR_INVALID;
  if (kind) *kind = sh->isField ? NIMFM_DS_CSR_FIELD : (sh->isCsr ? NIMFM_DS_CSR : NIMFM_DS_CSC);
  if (nRows) *nRows = sh->nRows;
  if (nCols) *nCols = sh->nCols;
  if (nnz) *nnz = sh->nnz;
  if (maxSegNnz) *maxSegNnz = sh->maxSeg;
  if (payloadBytes) *payloadBytes = sh->payloadEnd;
  return NIMFM_OK;
}

// the largest segEnd such that segments [segBegin, segEnd) hold at most maxBytes of file payload (always at
// least one segment): how the reference sizes a cache window (sparse_stream.nim:232-270, cacheSize in bytes)
int64_t nimfm_stream_window_end(const nimfm_stream *sh, int64_t segBegin, int64_t maxBytes) {
  if (!sh || segBegin < 0 || segBegin >= sh->nSeg) return sh ? sh->nSeg : 0;
  const int64_t start = sh->segOff[(size_t)segBegin] - 8;
  // segOff is increasing: first segment whose end offset exceeds start + maxBytes
  int64_t lo = segBegin + 1, hi = sh->nSeg;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo + 1) / 2;
    const int64_t endOff = mid < sh->nSeg ? sh->segOff[(size_t)mid] - 8 : sh->payloadEnd;
    if (endOff - start <= maxBytes) lo = mid; else hi = mid - 1;
  }
  return lo;
}

int32_t nimfm_stream_load_window(nimfm_ctx *ctx, nimfm_stream *sh, int64_t segBegin, int64_t segEnd,
                                 nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  REQUIRE(sh != nullptr && out != nullptr, "NULL argument");
  REQUIRE(segBegin >= 0 && segBegin <= segEnd && segEnd <= sh->nSeg, "window [%lld,%lld) outside [0,%lld)",
          (long long)segBegin, (long long)segEnd, (long long)sh->nSeg);
  // a window of columns would be a matrix with other column ids: only rows can be windowed
  REQUIRE(sh->isCsr || (segBegin == 0 && segEnd == sh->nSeg), "a StreamCSC file is loaded whole");
  CK(cudaSetDevice(ctx->device));
  const int64_t nSeg = segEnd - segBegin;
  const int64_t nnz = sh->indptr[(size_t)segEnd] - sh->indptr[(size_t)segBegin];
  const int64_t byteBegin = nSeg ? sh->segOff[(size_t)segBegin] - 8 : 0;
  const int64_t byteEnd = nSeg ? (segEnd < sh->nSeg ? sh->segOff[(size_t)segEnd] - 8 : sh->payloadEnd) : 0;
  const size_t payloadBytes = (size_t)(byteEnd - byteBegin);
  std::vector<int64_t> segOff((size_t)nSeg), indptr((size_t)nSeg + 1, 0);
  int64_t maxSeg = 0;
  for (int64_t g = 0; g < nSeg; g++) {
    segOff[(size_t)g] = sh->segOff[(size_t)(segBegin + g)] - byteBegin;
    indptr[(size_t)g + 1] = sh->indptr[(size_t)(segBegin + g + 1)] - sh->indptr[(size_t)segBegin];
    maxSeg = std::max(maxSeg, indptr[(size_t)g + 1] - indptr[(size_t)g]);
  }
  nimfm_dataset *ds = new nimfm_dataset();
  ds->kind = sh->isField ? NIMFM_DS_CSR_FIELD : (sh->isCsr ? NIMFM_DS_CSR : NIMFM_DS_CSC);
  ds->nFields = sh->isField ? sh->nFields : 0;
  ds->n = sh->isCsr ? nSeg : sh->nRows;
  ds->d = sh->nCols; ds->nnz = nnz; ds->maxSegNnz = maxSeg;
  auto fail = [&](int rc) { nimfm_dataset_free(ctx, ds); return rc; };
  unsigned char *dPayload = nullptr;
  int64_t *dSegOff = nullptr;
  int *dBad = nullptr;
  cudaError_t ce = cudaSuccess;
  auto ck = [&](cudaError_t e) { if (ce == cudaSuccess) ce = e; };
  ck(cudaMalloc(&dPayload, std::max<size_t>(payloadBytes, 16)));
  ck(cudaMalloc(&dSegOff, (size_t)std::max<int64_t>(nSeg, 1) * 8));
  ck(cudaMalloc(&dBad, sizeof(int)));
  ck(cudaMalloc(&ds->indptr, ((size_t)nSeg + 1) * 8));
  ck(cudaMalloc(&ds->data, (size_t)std::max<int64_t>(nnz, 2) * 8));
  ck(cudaMalloc(&ds->indices, (size_t)std::max<int64_t>(nnz, 4) * 4));
  if (sh->isField) ck(cudaMalloc(&ds->fields, (size_t)std::max<int64_t>(nnz, 4) * 4));
  if (ce == cudaSuccess) {
    if (payloadBytes && nimfm_staged_h2d(ctx, dPayload, sh->mx.p + sh->hdrBytes + byteBegin, payloadBytes) != NIMFM_OK)
      ck(cudaErrorUnknown);
    if (nSeg) ck(cudaMemcpyAsync(dSegOff, segOff.data(), (size_t)nSeg * 8, cudaMemcpyHostToDevice, ctx->stream));
    ck(cudaMemcpyAsync(ds->indptr, indptr.data(), ((size_t)nSeg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    ck(cudaMemsetAsync(dBad, 0, sizeof(int), ctx->stream));
    if (nSeg > 0 && nnz > 0) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nSeg * 32 + 255) / 256, (int64_t)ctx->numSMs * 16));
      stream_deinterleave_kernel<<<grid, 256, 0, ctx->stream>>>(dPayload, dSegOff, ds->indptr, nSeg, sh->extent, ds->data,
                                                                ds->indices, dBad, sh->recBytes, sh->valOff, sh->idOff,
                                                                ds->fields, sh->nFields);
      LAUNCHED(ctx);
    }
    int hbad = 0;
    ck(cudaMemcpyAsync(&hbad, dBad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ck(cudaStreamSynchronize(ctx->stream));
    ck(cudaGetLastError());
    if (ce == cudaSuccess && hbad) {
      cudaFree(dPayload); cudaFree(dSegOff); cudaFree(dBad);
      nimfm_dataset_free(ctx, ds);
      if (hbad == 2)
        return nimfm_fail(ctx, NIMFM_ERR_INVALID, "stream file: field out of range [0,%lld)", (long long)sh->nFields);
      return nimfm_fail(ctx, NIMFM_ERR_INVALID, "stream file: element id out of range [0,%lld)", (long long)sh->extent);
    }
  }
  cudaFree(dPayload); cudaFree(dSegOff); cudaFree(dBad);
  if (ce != cudaSuccess) return fail(nimfm_fail(ctx, NIMFM_ERR_CUDA, "nimfm_stream_load_window: %s", cudaGetErrorString(ce)));
  if (ds->kind == NIMFM_DS_CSR || ds->kind == NIMFM_DS_CSR_FIELD) {
    // hot columns from a row sample, as for uploaded datasets (nimfm_find_hot): the sampled rows' ids are
    // gathered out of the interleaved payload into a small CSR first
    std::vector<int32_t> hot;
    if (nSeg > 0) {
      const int64_t stride = std::max<int64_t>(1, nSeg / 2048);
      std::vector<int64_t> sIdx, sPtr(1, 0);
      for (int64_t g = 0; g < nSeg; g += stride) {
        const char *rec = sh->mx.p + sh->hdrBytes + sh->segOff[(size_t)(segBegin + g)];
        const int64_t z = indptr[(size_t)g + 1] - indptr[(size_t)g];
        for (int64_t t = 0; t < z; t++) {
          int64_t id;
          memcpy(&id, rec + sh->recBytes * t + sh->idOff, 8);
          sIdx.push_back(id);
        }
        sPtr.push_back((int64_t)sIdx.size());
      }
      nimfm_find_hot(sIdx.data(), sPtr.data(), 0, (int64_t)sPtr.size() - 1, hot, 2048, ds->d);
    }
    int rc = nimfm_upload_hot(ctx, hot, ds->d, &ds->hotSlot, &ds->hotList);
    if (rc) return fail(rc);
    ds->nHot = (int)hot.size();
  }
  if (sh->hasY && sh->isCsr && nSeg > 0) {
    std::vector<double> y((size_t)nSeg);
    if (pread(sh->my.fd, y.data(), (size_t)nSeg * 8, (off_t)(segBegin * 8)) != (ssize_t)(nSeg * 8))
      return fail(nimfm_fail(ctx, NIMFM_ERR_INVALID, "cannot read the label file"));
    int rc = nimfm_dataset_set_targets(ctx, ds, y.data());
    if (rc) return fail(rc);
  } else if (sh->hasY && !sh->isCsr && sh->nRows > 0) {
    std::vector<double> y((size_t)sh->nRows);
    if (pread(sh->my.fd, y.data(), (size_t)sh->nRows * 8, 0) != (ssize_t)(sh->nRows * 8))
      return fail(nimfm_fail(ctx, NIMFM_ERR_INVALID, "cannot read the label file"));
    int rc = nimfm_dataset_set_targets(ctx, ds, y.data());
    if (rc) return fail(rc);
  }
  *out = ds;
  return NIMFM_OK;
}

int32_t nimfm_stream_close(nimfm_stream *sh) {
  delete sh;
  return NIMFM_OK;
}

int32_t nimfm_load_stream(nimfm_ctx *ctx, const char *pathX, const char *pathY, nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr && pathX != nullptr, "NULL argument");
  nimfm_stream *sh = nullptr;
  int rc = nimfm_stream_open(ctx, pathX, pathY, &sh);
  if (rc) return rc;
  rc = nimfm_stream_load_window(ctx, sh, 0, sh->nSeg, out);
  nimfm_stream_close(sh);
  return rc;
}

int32_t nimfm_dataset_get_targets(nimfm_ctx *ctx, const nimfm_dataset *ds, double *y) {
  if (!ctx || !ds) return NIMFM_ERR_INVALID;
  REQUIRE(y != nullptr || ds->n == 0, "y is NULL");
  REQUIRE(ds->y != nullptr || ds->n == 0, "dataset has no targets");
  CK(cudaSetDevice(ctx->device));
  if (ds->n) CK(cudaMemcpy(y, ds->y, (size_t)ds->n * 8, cudaMemcpyDeviceToHost));
  return NIMFM_OK;
}

}  // extern "C"

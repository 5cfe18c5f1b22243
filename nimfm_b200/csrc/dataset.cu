// dataset.cu -- device-resident CSR / CSC / CSR-with-fields containers.
// Integer bookkeeping only: index narrowing int64 -> int32 and row-shard rebasing (X[slice],
// tensor/sparse.nim:263-290).  Row gather / CSC slice / vstack / transpose live in dataset_ops.cu.
#include <algorithm>
#include <unordered_map>

#include "common.cuh"

static int alloc_copy(nimfm_ctx *ctx, void **dst, const void *src, size_t bytes) {
  CK(cudaMalloc(dst, bytes ? bytes : 16));
  if (bytes) CK(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return NIMFM_OK;
}

// Find the "hot" columns of a CSR from an evenly spaced sample of at most 16384 rows: present in at
// least 1/16 of the sampled rows, the 16 most frequent.  Pure bookkeeping; it only changes where the
// kernels ACCUMULATE the gradients of those columns, never a result.
// The sample runs BEFORE the caller's arrays are fully validated (the host-fed calls check chunk by chunk),
// so it trusts nothing: rows whose indptr pair is not a sane range inside [lo, hi) are skipped and ids outside
// [0, d) are ignored -- the validation that follows rejects the call; the sample must only not crash on it
// or hand an out-of-range id to the hot table.
int nimfm_find_hot(const int64_t *indices, const int64_t *indptr, int64_t rowBegin, int64_t rowEnd,
                   std::vector<int32_t> &hot, int64_t maxSample, int64_t d) {
  hot.clear();
  const int64_t ns = rowEnd - rowBegin;
  if (ns <= 0) return 0;
  const int64_t stride = std::max<int64_t>(1, ns / maxSample);
  const int64_t lo = indptr[rowBegin], hi = indptr[rowEnd];
  std::unordered_map<int64_t, int> cnt;
  int64_t sampled = 0;
  for (int64_t r = rowBegin; r < rowEnd; r += stride, ++sampled) {
    const int64_t q0 = indptr[r], q1 = indptr[r + 1];
    if (q0 < lo || q1 > hi || q1 < q0) continue;
    for (int64_t q = q0; q < q1; q++) {
      const int64_t j = indices[q];
      if (j >= 0 && j < d) cnt[j] += 1;
    }
  }
  std::vector<std::pair<int, int64_t>> cand;
  for (auto &kv : cnt)
    if ((int64_t)kv.second * 16 >= sampled && kv.second >= 2) cand.push_back({kv.second, kv.first});
  std::sort(cand.begin(), cand.end(), [](const std::pair<int, int64_t> &x, const std::pair<int, int64_t> &y) {
    return x.first != y.first ? x.first > y.first : x.second < y.second;
  });
  for (size_t i = 0; i < cand.size() && i < 16; i++) hot.push_back((int32_t)cand[i].second);
  return (int)hot.size();
}

int nimfm_upload_hot(nimfm_ctx *ctx, const std::vector<int32_t> &hot, int64_t d, uint8_t **hotSlot,
                     int32_t **hotList) {
  std::vector<uint8_t> tab((size_t)std::max<int64_t>(d, 1), 255);
  for (size_t i = 0; i < hot.size(); i++) tab[(size_t)hot[i]] = (uint8_t)i;
  if (!*hotSlot) CK(cudaMalloc(hotSlot, tab.size()));
  if (!*hotList) CK(cudaMalloc(hotList, 16 * sizeof(int32_t)));
  CK(cudaMemcpyAsync(*hotSlot, tab.data(), tab.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!hot.empty())
    CK(cudaMemcpyAsync(*hotList, hot.data(), hot.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return NIMFM_OK;
}

static int upload_common(nimfm_ctx *ctx, int kind, int64_t nSeg, int64_t nOther, int64_t n, int64_t d,
                         const double *data, const int64_t *indices, const int64_t *indptr,
                         const int64_t *fields, int64_t nFields, int64_t segBegin, int64_t segEnd,
                         nimfm_dataset **out) {
  // nSeg: number of compressed segments in the host arrays (rows for CSR, columns for CSC);
  // nOther: extent of the index values.
  REQUIRE(out != nullptr, "out is NULL");
  REQUIRE(n >= 0 && d >= 0, "negative shape");
  REQUIRE(indptr != nullptr, "indptr is NULL");
  REQUIRE(segBegin >= 0 && segBegin <= segEnd && segEnd <= nSeg, "bad segment range [%lld,%lld) of %lld",
          (long long)segBegin, (long long)segEnd, (long long)nSeg);
  REQUIRE(nOther < (int64_t)2147483647, "index extent %lld does not fit int32", (long long)nOther);
  const int64_t base = indptr[segBegin];
  const int64_t nnz = indptr[segEnd] - base;
  REQUIRE(nnz >= 0, "indptr is not monotone");
  REQUIRE(nnz == 0 || (data != nullptr && indices != nullptr), "data/indices are NULL");
  const int64_t ns = segEnd - segBegin;
  std::vector<int64_t> ptr((size_t)ns + 1);
  int64_t maxSeg = 0;
  for (int64_t s = 0; s <= ns; s++) {
    ptr[s] = indptr[segBegin + s] - base;
    if (s > 0) {
      int64_t len = ptr[s] - ptr[s - 1];
      REQUIRE(len >= 0, "indptr is not monotone at %lld", (long long)(segBegin + s));
      maxSeg = std::max(maxSeg, len);
    }
  }
  nimfm_dataset *ds = new nimfm_dataset();
  // every early return below (CK / REQUIRE included) releases the half-built dataset and its device buffers
  struct Guard {
    nimfm_ctx *ctx;
    nimfm_dataset *ds;
    ~Guard() { if (ds) nimfm_dataset_free(ctx, ds); }
  } guard{ctx, ds};
  ds->kind = kind;
  ds->n = (kind == NIMFM_DS_CSC) ? n : ns;
  ds->d = (kind == NIMFM_DS_CSC) ? ns : d;
  ds->nnz = nnz;
  ds->nFields = fields ? nFields : 0;
  ds->maxSegNnz = maxSeg;
  // values, ids and field ids reach the device through the pinned-piece thread team (the ids narrowed to
  // int32 and range-checked on the way): 2 M Criteo-shaped rows 0.33 s -> the link's rate
  int rc, bad = 0;
  auto fail = [&](int code) { return code; };   // the guard frees
  CK(cudaMalloc(&ds->data, (size_t)std::max<int64_t>(nnz, 2) * 8));
  CK(cudaMalloc(&ds->indices, (size_t)std::max<int64_t>(nnz, 4) * 4));
  if (fields) CK(cudaMalloc(&ds->fields, (size_t)std::max<int64_t>(nnz, 4) * 4));
  if (nnz > 0) {
    if ((rc = nimfm_staged_h2d(ctx, ds->data, data + base, (size_t)nnz * 8))) return fail(rc);
    if ((rc = nimfm_staged_h2d(ctx, ds->indices, indices + base, (size_t)nnz * 8, nOther > 0 ? nOther : 1, &bad))) return fail(rc);
    if (bad) {   // error path only: name the first offender like the sequential check did
      CK(cudaStreamSynchronize(ctx->stream));
      for (int64_t q = 0; q < nnz; q++) {
        const int64_t v = indices[base + q];
        if (v < 0 || v >= nOther)
          return fail(nimfm_fail(ctx, NIMFM_ERR_INVALID, "index %lld out of range [0,%lld) at nnz %lld", (long long)v,
                                 (long long)nOther, (long long)(base + q)));
      }
    }
    if (fields) {
      if ((rc = nimfm_staged_h2d(ctx, ds->fields, fields + base, (size_t)nnz * 8, nFields > 0 ? nFields : 1, &bad))) return fail(rc);
      if (bad) {
        CK(cudaStreamSynchronize(ctx->stream));
        for (int64_t q = 0; q < nnz; q++) {
          const int64_t v = fields[base + q];
          if (v < 0 || v >= nFields)
            return fail(nimfm_fail(ctx, NIMFM_ERR_INVALID, "field %lld out of range [0,%lld)", (long long)v, (long long)nFields));
        }
      }
    }
  }
  if ((rc = alloc_copy(ctx, (void **)&ds->indptr, ptr.data(), ((size_t)ns + 1) * 8))) return fail(rc);
  if (kind != NIMFM_DS_CSC) {
    std::vector<int32_t> hot;
    ds->nHot = nimfm_find_hot(indices, indptr, segBegin, segEnd, hot, 16384, d);
    if ((rc = nimfm_upload_hot(ctx, hot, d, &ds->hotSlot, &ds->hotList))) return rc;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  guard.ds = nullptr;
  *out = ds;
  return NIMFM_OK;
}

extern "C" {

int32_t nimfm_csr_upload(nimfm_ctx *ctx, int64_t n, int64_t d, const double *data,
                         const int64_t *indices, const int64_t *indptr, const int64_t *fields,
                         int64_t nFields, int64_t rowBegin, int64_t rowEnd, nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return upload_common(ctx, fields ? NIMFM_DS_CSR_FIELD : NIMFM_DS_CSR, n, d, n, d, data, indices, indptr,
                       fields, nFields, rowBegin, rowEnd, out);
}

int32_t nimfm_csc_upload(nimfm_ctx *ctx, int64_t n, int64_t d, const double *data,
                         const int64_t *indices, const int64_t *indptr, nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return upload_common(ctx, NIMFM_DS_CSC, d, n, n, d, data, indices, indptr, nullptr, 0, 0, d, out);
}

int32_t nimfm_dataset_set_targets(nimfm_ctx *ctx, nimfm_dataset *ds, const double *y) {
  if (!ctx || !ds) return NIMFM_ERR_INVALID;
  REQUIRE(y != nullptr || ds->n == 0, "y is NULL");
  CK(cudaSetDevice(ctx->device));
  if (!ds->y) CK(cudaMalloc(&ds->y, (size_t)(ds->n ? ds->n : 1) * 8));
  CK(cudaMemcpyAsync(ds->y, y, (size_t)ds->n * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return NIMFM_OK;
}

int32_t nimfm_dataset_info(const nimfm_dataset *ds, int64_t *n, int64_t *d, int64_t *nnz, int32_t *kind,
                           int64_t *nFields, int64_t *maxRowNnz) {
  if (!ds) return NIMFM_ERR_INVALID;
  if (n) *n = ds->n;
  if (d) *d = ds->d;
  if (nnz) *nnz = ds->nnz;
  if (kind) *kind = ds->kind;
  if (nFields) *nFields = ds->nFields;
  if (maxRowNnz) *maxRowNnz = ds->maxSegNnz;
  return NIMFM_OK;
}

int32_t nimfm_dataset_download(nimfm_ctx *ctx, const nimfm_dataset *ds, double *data, int64_t *indices,
                               int64_t *indptr, int64_t *fields) {
  if (!ctx || !ds) return NIMFM_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  const int64_t ns = ds->kind == NIMFM_DS_CSC ? ds->d : ds->n;
  if (data) CK(cudaMemcpy(data, ds->data, (size_t)ds->nnz * 8, cudaMemcpyDeviceToHost));
  if (indptr) CK(cudaMemcpy(indptr, ds->indptr, ((size_t)ns + 1) * 8, cudaMemcpyDeviceToHost));
  if (indices) {
    std::vector<int32_t> tmp((size_t)ds->nnz);
    CK(cudaMemcpy(tmp.data(), ds->indices, (size_t)ds->nnz * 4, cudaMemcpyDeviceToHost));
    for (int64_t q = 0; q < ds->nnz; q++) indices[q] = tmp[q];
  }
  if (fields) {
    REQUIRE(ds->fields != nullptr, "dataset has no fields");
    std::vector<int32_t> tmp((size_t)ds->nnz);
    CK(cudaMemcpy(tmp.data(), ds->fields, (size_t)ds->nnz * 4, cudaMemcpyDeviceToHost));
    for (int64_t q = 0; q < ds->nnz; q++) fields[q] = tmp[q];
  }
  return NIMFM_OK;
}

int32_t nimfm_dataset_free(nimfm_ctx *ctx, nimfm_dataset *ds) {
  if (!ds) return NIMFM_OK;
  if (ctx) cudaSetDevice(ctx->device);
  if (ds->detTwin) nimfm_det_twin_free(ctx, ds->detTwin);
  cudaFree(ds->data);
  cudaFree(ds->indices);
  cudaFree(ds->indptr);
  cudaFree(ds->fields);
  cudaFree(ds->y);
  cudaFree(ds->hotSlot);
  cudaFree(ds->hotList);
  delete ds;
  return NIMFM_OK;
}

}  // extern "C"

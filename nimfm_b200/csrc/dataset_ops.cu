// dataset_ops.cu -- device-side integer bookkeeping on resident datasets:
//   X[indicesRow]            tensor/sparse.nim:263-285  (CSR / CSR-with-fields row gather)
//   X[a..b] on a CSC         tensor/sparse.nim:300-325  (stable per-column filter, row ids rebased)
//   vstack                   tensor/sparse.nim:564-610  (CSR: concatenation; CSC: per-column interleave)
//   toCSCMatrix/toCSRMatrix  tensor/sparse.nim:490-527  (stable counting sort == stable radix sort by the
//                                                       other axis; within a column rows stay ascending)
// Everything here must be BIT-EXACT with the reference (same indptr, same within-segment order); the
// parity tests download the result and compare with the oracle's restatement.
// Scans / the radix sort come from CUB (ships with the CUDA toolkit); this is set-up work, not the hot path.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

namespace {

struct DsGuard {   // owns a dataset under construction: freed on every early return, released on success
  nimfm_ctx *ctx;
  nimfm_dataset *ds;
  DsGuard(nimfm_ctx *c, nimfm_dataset *d) : ctx(c), ds(d) {}
  ~DsGuard() { if (ds) nimfm_dataset_free(ctx, ds); }
  nimfm_dataset *release() { nimfm_dataset *d = ds; ds = nullptr; return d; }
};

struct DevBuf {   // frees on scope exit so early returns do not leak
  void *p = nullptr;
  ~DevBuf() { cudaFree(p); }
  template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

__global__ void seg_len_kernel(const int64_t *indptr, const int64_t *rows, int64_t nIdx, int64_t *len) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nIdx; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows[q];
    len[q] = indptr[r + 1] - indptr[r];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) len[nIdx] = 0;
}

// one warp per output row: copy the source row's nonzeros
__global__ void gather_rows_kernel(const int64_t *srcPtr, const int64_t *dstPtr, const int64_t *rows, int64_t nIdx,
                                   const double *sData, const int32_t *sIdx, const int32_t *sFld, double *dData,
                                   int32_t *dIdx, int32_t *dFld, const double *sY, double *dY) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = warp; q < nIdx; q += nWarps) {
    const int64_t r = rows[q];
    const int64_t sb = srcPtr[r], db = dstPtr[q], z = srcPtr[r + 1] - sb;
    for (int64_t t = lane; t < z; t += 32) {
      dData[db + t] = sData[sb + t];
      dIdx[db + t] = sIdx[sb + t];
      if (sFld) dFld[db + t] = sFld[sb + t];
    }
    if (lane == 0 && sY) dY[q] = sY[r];
  }
}

__global__ void flag_range_kernel(const int32_t *idx, int64_t nnz, int32_t lo, int32_t hi, int64_t *flag) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q <= nnz; q += (int64_t)gridDim.x * blockDim.x)
    flag[q] = (q < nnz && idx[q] >= lo && idx[q] <= hi) ? 1 : 0;
}

__global__ void compact_range_kernel(const int32_t *idx, const double *data, int64_t nnz, int32_t lo, int32_t hi,
                                     const int64_t *pos, int32_t *oIdx, double *oData) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = idx[q];
    if (i >= lo && i <= hi) {
      oIdx[pos[q]] = i - lo;
      oData[pos[q]] = data[q];
    }
  }
}

__global__ void remap_ptr_kernel(const int64_t *inPtr, int64_t nSeg, const int64_t *pos, int64_t *outPtr) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s <= nSeg; s += (int64_t)gridDim.x * blockDim.x)
    outPtr[s] = pos[inPtr[s]];
}

// out segment s = part segment s placed at outPtr[s] + off[s]; indices shifted by idxShift
__global__ void place_segments_kernel(const int64_t *pPtr, int64_t nSeg, const int64_t *outPtr, const int64_t *off,
                                      const double *pData, const int32_t *pIdx, const int32_t *pFld, int32_t idxShift,
                                      double *oData, int32_t *oIdx, int32_t *oFld) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < nSeg; s += nWarps) {
    const int64_t sb = pPtr[s], z = pPtr[s + 1] - sb, db = outPtr[s] + (off ? off[s] : 0);
    for (int64_t t = lane; t < z; t += 32) {
      oData[db + t] = pData[sb + t];
      oIdx[db + t] = pIdx[sb + t] + idxShift;
      if (pFld) oFld[db + t] = pFld[sb + t];
    }
  }
}

__global__ void add_seg_len_kernel(const int64_t *pPtr, int64_t nSeg, int64_t *acc, int64_t *offBefore) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nSeg; s += (int64_t)gridDim.x * blockDim.x) {
    if (offBefore) offBefore[s] = acc[s];
    acc[s] += pPtr[s + 1] - pPtr[s];
  }
}

__global__ void shift_ptr_kernel(const int64_t *pPtr, int64_t nSeg, int64_t base, int64_t *outPtr) {
  // outPtr[1 + s] = base + pPtr[1 + s]  (vstack of CSR: indptr &= X.indptr[1..^1].map(x => nnz + x))
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nSeg; s += (int64_t)gridDim.x * blockDim.x)
    outPtr[1 + s] = base + pPtr[1 + s];
}

// expand indptr into one segment id per nonzero (warp per segment), and the identity permutation
__global__ void expand_seg_ids_kernel(const int64_t *ptr, int64_t nSeg, int32_t *segOf) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < nSeg; s += nWarps)
    for (int64_t q = ptr[s] + lane; q < ptr[s + 1]; q += 32) segOf[q] = (int32_t)s;
}

__global__ void iota_kernel(uint32_t *p, int64_t n) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    p[q] = (uint32_t)q;
}

__global__ void permute_gather_kernel(const uint32_t *perm, int64_t nnz, const double *data, const int32_t *segOf,
                                      double *oData, int32_t *oIdx) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = perm[q];
    oData[q] = data[p];
    oIdx[q] = segOf[p];
  }
}

// outPtr[t] = first position whose (sorted) key is >= t
__global__ void lower_bound_ptr_kernel(const int32_t *sortedKeys, int64_t nnz, int64_t nOut, int64_t *outPtr) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= nOut; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)sortedKeys[mid] < t) lo = mid + 1; else hi = mid;
    }
    outPtr[t] = lo;
  }
}

__global__ void max_seg_kernel(const int64_t *ptr, int64_t nSeg, unsigned long long *out) {
  unsigned long long m = 0;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nSeg; s += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (unsigned long long)(ptr[s + 1] - ptr[s]));
  for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

int grid_for(nimfm_ctx *ctx, int64_t items, int perBlock = 256) {
  int64_t g = (items + perBlock - 1) / perBlock;
  g = std::min<int64_t>(g, (int64_t)ctx->numSMs * 16);
  return (int)std::max<int64_t>(g, 1);
}

int exclusive_scan(nimfm_ctx *ctx, const int64_t *in, int64_t *out, int64_t n) {
  size_t tmpBytes = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, in, out, n, ctx->stream));
  DevBuf tmp;
  CK(cudaMalloc(&tmp.p, tmpBytes ? tmpBytes : 16));
  CK(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, in, out, n, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return NIMFM_OK;
}

int finish_dataset(nimfm_ctx *ctx, nimfm_dataset *o, const nimfm_dataset *hotFrom) {
  const int64_t nSeg = o->kind == NIMFM_DS_CSC ? o->d : o->n;
  DevBuf m;
  CK(cudaMalloc(&m.p, 8));
  CK(cudaMemsetAsync(m.p, 0, 8, ctx->stream));
  if (nSeg > 0) {
    max_seg_kernel<<<grid_for(ctx, nSeg), 256, 0, ctx->stream>>>(o->indptr, nSeg, m.as<unsigned long long>());
    LAUNCHED(ctx);
  }
  unsigned long long h = 0;
  CK(cudaMemcpyAsync(&h, m.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  o->maxSegNnz = (int64_t)h;
  if (o->kind != NIMFM_DS_CSC) {
    // derived row sets keep the parent's hot-column table (bookkeeping only: it selects where the
    // gradients of those columns are accumulated, never what is computed)
    const size_t tab = (size_t)std::max<int64_t>(o->d, 1);
    CK(cudaMalloc(&o->hotSlot, tab));
    CK(cudaMalloc(&o->hotList, 16 * sizeof(int32_t)));
    if (hotFrom && hotFrom->hotSlot && hotFrom->d == o->d) {
      CK(cudaMemcpyAsync(o->hotSlot, hotFrom->hotSlot, tab, cudaMemcpyDeviceToDevice, ctx->stream));
      CK(cudaMemcpyAsync(o->hotList, hotFrom->hotList, 16 * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
      o->nHot = hotFrom->nHot;
    } else {
      CK(cudaMemsetAsync(o->hotSlot, 255, tab, ctx->stream));
      o->nHot = 0;
    }
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaGetLastError());
  return NIMFM_OK;
}

int alloc_arrays(nimfm_ctx *ctx, nimfm_dataset *o, int64_t nSeg, int64_t nnz, bool fields) {
  CK(cudaMalloc(&o->indptr, ((size_t)nSeg + 1) * 8));
  CK(cudaMalloc(&o->data, (size_t)std::max<int64_t>(nnz, 2) * 8));
  CK(cudaMalloc(&o->indices, (size_t)std::max<int64_t>(nnz, 4) * 4));
  if (fields) CK(cudaMalloc(&o->fields, (size_t)std::max<int64_t>(nnz, 4) * 4));
  return NIMFM_OK;
}

}  // namespace

extern "C" {

// X[indicesRow] for CSRDataset / CSRFieldDataset (dataset.nim:319-367 -> tensor/sparse.nim:263-285);
// targets, when set, are gathered alongside (== shuffle(X, y, indices), dataset.nim:372-381).
int32_t nimfm_dataset_take_rows(nimfm_ctx *ctx, const nimfm_dataset *in, const int64_t *rowIdx, int64_t nIdx,
                                nimfm_dataset **out) {
  if (!ctx || !in) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr, "out is NULL");
  REQUIRE(in->kind != NIMFM_DS_CSC, "Accessing row vectors by an index array is not supported for CSC");  // sparse.nim:298-299
  REQUIRE(nIdx >= 0 && (rowIdx != nullptr || nIdx == 0), "rowIdx is NULL");
  for (int64_t q = 0; q < nIdx; q++) {   // checkIndicesRow (sparse.nim:254-260)
    REQUIRE(rowIdx[q] < in->n, "max(indicesRow) %lld >= %lld.", (long long)rowIdx[q], (long long)in->n);
    REQUIRE(rowIdx[q] >= 0, "min(indicesRow) %lld < 0.", (long long)rowIdx[q]);
  }
  CK(cudaSetDevice(ctx->device));
  DevBuf rows, len;
  CK(cudaMalloc(&rows.p, (size_t)std::max<int64_t>(nIdx, 1) * 8));
  CK(cudaMalloc(&len.p, ((size_t)nIdx + 1) * 8));
  if (nIdx) CK(cudaMemcpyAsync(rows.p, rowIdx, (size_t)nIdx * 8, cudaMemcpyHostToDevice, ctx->stream));
  nimfm_dataset *o = new nimfm_dataset();
  DsGuard guard(ctx, o);
  o->kind = in->kind; o->n = nIdx; o->d = in->d; o->nFields = in->nFields;
  CK(cudaMalloc(&o->indptr, ((size_t)nIdx + 1) * 8));
  seg_len_kernel<<<grid_for(ctx, nIdx), 256, 0, ctx->stream>>>(in->indptr, rows.as<int64_t>(), nIdx, len.as<int64_t>());
  LAUNCHED(ctx);
  int rc = exclusive_scan(ctx, len.as<int64_t>(), o->indptr, nIdx + 1);
  if (rc) return rc;
  int64_t nnz = 0;
  CK(cudaMemcpy(&nnz, o->indptr + nIdx, 8, cudaMemcpyDeviceToHost));
  o->nnz = nnz;
  CK(cudaMalloc(&o->data, (size_t)std::max<int64_t>(nnz, 2) * 8));
  CK(cudaMalloc(&o->indices, (size_t)std::max<int64_t>(nnz, 4) * 4));
  if (in->fields) CK(cudaMalloc(&o->fields, (size_t)std::max<int64_t>(nnz, 4) * 4));
  if (in->y) CK(cudaMalloc(&o->y, (size_t)std::max<int64_t>(nIdx, 1) * 8));
  if (nIdx) {
    gather_rows_kernel<<<grid_for(ctx, nIdx * 32), 256, 0, ctx->stream>>>(
        in->indptr, o->indptr, rows.as<int64_t>(), nIdx, in->data, in->indices, in->fields, o->data, o->indices,
        o->fields, in->y, o->y);
    LAUNCHED(ctx);
  }
  if ((rc = finish_dataset(ctx, o, in))) return rc;
  *out = guard.release();
  return NIMFM_OK;
}

// X[first..last] (inclusive, as Nim's Slice).  CSR kinds: == X[toSeq(slice)] (sparse.nim:288-290);
// CSC: the O(nnz) per-column filter of sparse.nim:300-325 (entries keep their order, row ids -= first).
int32_t nimfm_dataset_slice_rows(nimfm_ctx *ctx, const nimfm_dataset *in, int64_t first, int64_t last,
                                 nimfm_dataset **out) {
  if (!ctx || !in) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr, "out is NULL");
  REQUIRE(first >= 0, "min(indicesRow) %lld < 0.", (long long)first);
  REQUIRE(last < in->n, "max(indicesRow) %lld >= %lld.", (long long)last, (long long)in->n);
  REQUIRE(first <= last, "empty slice [%lld..%lld]", (long long)first, (long long)last);
  if (in->kind != NIMFM_DS_CSC) {
    std::vector<int64_t> ids((size_t)(last - first + 1));
    for (size_t q = 0; q < ids.size(); q++) ids[q] = first + (int64_t)q;
    return nimfm_dataset_take_rows(ctx, in, ids.data(), (int64_t)ids.size(), out);
  }
  CK(cudaSetDevice(ctx->device));
  const int64_t nnzIn = in->nnz, d = in->d;
  DevBuf flag, pos;
  CK(cudaMalloc(&flag.p, ((size_t)nnzIn + 1) * 8));
  CK(cudaMalloc(&pos.p, ((size_t)nnzIn + 1) * 8));
  flag_range_kernel<<<grid_for(ctx, nnzIn + 1), 256, 0, ctx->stream>>>(in->indices, nnzIn, (int32_t)first, (int32_t)last,
                                                                      flag.as<int64_t>());
  LAUNCHED(ctx);
  int rc = exclusive_scan(ctx, flag.as<int64_t>(), pos.as<int64_t>(), nnzIn + 1);
  if (rc) return rc;
  int64_t nnz = 0;
  CK(cudaMemcpy(&nnz, pos.as<int64_t>() + nnzIn, 8, cudaMemcpyDeviceToHost));
  nimfm_dataset *o = new nimfm_dataset();
  DsGuard guard(ctx, o);
  o->kind = NIMFM_DS_CSC; o->n = last - first + 1; o->d = d; o->nnz = nnz;
  if ((rc = alloc_arrays(ctx, o, d, nnz, false))) return rc;
  remap_ptr_kernel<<<grid_for(ctx, d + 1), 256, 0, ctx->stream>>>(in->indptr, d, pos.as<int64_t>(), o->indptr);
  LAUNCHED(ctx);
  if (nnzIn) {
    compact_range_kernel<<<grid_for(ctx, nnzIn), 256, 0, ctx->stream>>>(in->indices, in->data, nnzIn, (int32_t)first,
                                                                       (int32_t)last, pos.as<int64_t>(), o->indices,
                                                                       o->data);
    LAUNCHED(ctx);
  }
  if (in->y) {
    CK(cudaMalloc(&o->y, (size_t)o->n * 8));
    CK(cudaMemcpyAsync(o->y, in->y + first, (size_t)o->n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if ((rc = finish_dataset(ctx, o, nullptr))) return rc;
  *out = guard.release();
  return NIMFM_OK;
}

// vstack (dataset.nim:452-483 -> tensor/sparse.nim:564-640).  All parts of one kind and nFeatures.
int32_t nimfm_dataset_vstack(nimfm_ctx *ctx, const nimfm_dataset *const *parts, int32_t nParts, nimfm_dataset **out) {
  if (!ctx) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr && parts != nullptr && nParts >= 1, "bad arguments");
  const nimfm_dataset *p0 = parts[0];
  REQUIRE(p0 != nullptr, "NULL part");
  int64_t n = 0, nnz = 0;
  bool haveY = true;
  for (int i = 0; i < nParts; i++) {
    REQUIRE(parts[i] != nullptr, "NULL part");
    REQUIRE(parts[i]->kind == p0->kind, "all parts must be of one dataset kind");
    REQUIRE(parts[i]->d == p0->d, "All matrics must have the same shape[1].");   // sparse.nim:576-577
    REQUIRE(parts[i]->nFields == p0->nFields, "all parts must have the same nFields");
    n += parts[i]->n;
    nnz += parts[i]->nnz;
    haveY = haveY && (parts[i]->y != nullptr || parts[i]->n == 0);
  }
  REQUIRE(n < (int64_t)2147483647, "stacked row count does not fit int32");
  CK(cudaSetDevice(ctx->device));
  nimfm_dataset *o = new nimfm_dataset();
  DsGuard guard(ctx, o);
  o->kind = p0->kind; o->n = n; o->d = p0->d; o->nnz = nnz; o->nFields = p0->nFields;
  const bool csc = p0->kind == NIMFM_DS_CSC;
  const int64_t nSeg = csc ? o->d : n;
  int rc = alloc_arrays(ctx, o, nSeg, nnz, p0->fields != nullptr);
  if (rc) return rc;
  if (haveY && n > 0) CK(cudaMalloc(&o->y, (size_t)n * 8));
  if (!csc) {
    CK(cudaMemsetAsync(o->indptr, 0, 8, ctx->stream));
    int64_t rowBase = 0, nnzBase = 0;
    for (int i = 0; i < nParts; i++) {
      const nimfm_dataset *p = parts[i];
      if (p->n) {
        shift_ptr_kernel<<<grid_for(ctx, p->n), 256, 0, ctx->stream>>>(p->indptr, p->n, nnzBase, o->indptr + rowBase);
        LAUNCHED(ctx);
      }
      if (p->nnz) {
        CK(cudaMemcpyAsync(o->data + nnzBase, p->data, (size_t)p->nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(o->indices + nnzBase, p->indices, (size_t)p->nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        if (p->fields)
          CK(cudaMemcpyAsync(o->fields + nnzBase, p->fields, (size_t)p->nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      }
      if (o->y && p->n) CK(cudaMemcpyAsync(o->y + rowBase, p->y, (size_t)p->n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      rowBase += p->n;
      nnzBase += p->nnz;
    }
  } else {
    // column j of the result = column j of part 0, then of part 1 (row ids + n_0), ... (sparse.nim:598-609)
    const int64_t d = o->d;
    DevBuf acc, offs;
    CK(cudaMalloc(&acc.p, ((size_t)d + 1) * 8));
    CK(cudaMalloc(&offs.p, (size_t)std::max<int64_t>(d, 1) * 8 * nParts));
    CK(cudaMemsetAsync(acc.p, 0, ((size_t)d + 1) * 8, ctx->stream));
    for (int i = 0; i < nParts && d > 0; i++) {
      add_seg_len_kernel<<<grid_for(ctx, d), 256, 0, ctx->stream>>>(parts[i]->indptr, d, acc.as<int64_t>(),
                                                                   offs.as<int64_t>() + (size_t)i * d);
      LAUNCHED(ctx);
    }
    if ((rc = exclusive_scan(ctx, acc.as<int64_t>(), o->indptr, d + 1))) return rc;
    int64_t rowBase = 0;
    for (int i = 0; i < nParts; i++) {
      const nimfm_dataset *p = parts[i];
      if (d > 0 && p->nnz) {
        place_segments_kernel<<<grid_for(ctx, d * 32), 256, 0, ctx->stream>>>(
            p->indptr, d, o->indptr, offs.as<int64_t>() + (size_t)i * d, p->data, p->indices, nullptr, (int32_t)rowBase,
            o->data, o->indices, nullptr);
        LAUNCHED(ctx);
      }
      if (o->y && p->n) CK(cudaMemcpyAsync(o->y + rowBase, p->y, (size_t)p->n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      rowBase += p->n;
    }
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if ((rc = finish_dataset(ctx, o, p0))) return rc;
  *out = guard.release();
  return NIMFM_OK;
}

// toCSCMatrix / toCSRMatrix (tensor/sparse.nim:490-527) on the device: a STABLE radix sort of the
// nonzeros by their index along the other axis reproduces the reference's counting sort exactly
// (entries of one output segment stay in input-segment order).
int32_t nimfm_dataset_transpose(nimfm_ctx *ctx, const nimfm_dataset *in, nimfm_dataset **out) {
  if (!ctx || !in) return NIMFM_ERR_INVALID;
  REQUIRE(out != nullptr, "out is NULL");
  REQUIRE(in->kind != NIMFM_DS_CSR_FIELD, "transpose of field datasets is not supported");
  REQUIRE(in->nnz < (int64_t)2147483647, "nnz does not fit the 32-bit permutation");
  CK(cudaSetDevice(ctx->device));
  const bool toCsc = in->kind == NIMFM_DS_CSR;
  const int64_t nsIn = toCsc ? in->n : in->d, nsOut = toCsc ? in->d : in->n, nnz = in->nnz;
  nimfm_dataset *o = new nimfm_dataset();
  DsGuard guard(ctx, o);
  o->kind = toCsc ? NIMFM_DS_CSC : NIMFM_DS_CSR;
  o->n = in->n; o->d = in->d; o->nnz = nnz;
  int rc = alloc_arrays(ctx, o, nsOut, nnz, false);
  if (rc) return rc;
  DevBuf segOf, keysOut, permIn, permOut, tmp;
  const size_t n4 = (size_t)std::max<int64_t>(nnz, 4) * 4;
  CK(cudaMalloc(&segOf.p, n4));
  CK(cudaMalloc(&keysOut.p, n4));
  CK(cudaMalloc(&permIn.p, n4));
  CK(cudaMalloc(&permOut.p, n4));
  if (nnz > 0) {
    expand_seg_ids_kernel<<<grid_for(ctx, nsIn * 32), 256, 0, ctx->stream>>>(in->indptr, nsIn, segOf.as<int32_t>());
    iota_kernel<<<grid_for(ctx, nnz), 256, 0, ctx->stream>>>(permIn.as<uint32_t>(), nnz);
    ctx->launches += 2;
    int endBit = 1;
    while (endBit < 31 && ((int64_t)1 << endBit) < nsOut) endBit++;
    size_t tmpBytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, in->indices, keysOut.as<int32_t>(), permIn.as<uint32_t>(),
                                       permOut.as<uint32_t>(), (int)nnz, 0, endBit, ctx->stream));
    CK(cudaMalloc(&tmp.p, tmpBytes ? tmpBytes : 16));
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, in->indices, keysOut.as<int32_t>(), permIn.as<uint32_t>(),
                                       permOut.as<uint32_t>(), (int)nnz, 0, endBit, ctx->stream));
    permute_gather_kernel<<<grid_for(ctx, nnz), 256, 0, ctx->stream>>>(permOut.as<uint32_t>(), nnz, in->data,
                                                                      segOf.as<int32_t>(), o->data, o->indices);
    LAUNCHED(ctx);
  }
  lower_bound_ptr_kernel<<<grid_for(ctx, nsOut + 1), 256, 0, ctx->stream>>>(keysOut.as<int32_t>(), nnz, nsOut, o->indptr);
  LAUNCHED(ctx);
  if (in->y) {
    CK(cudaMalloc(&o->y, (size_t)std::max<int64_t>(o->n, 1) * 8));
    CK(cudaMemcpyAsync(o->y, in->y, (size_t)o->n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if ((rc = finish_dataset(ctx, o, nullptr))) return rc;
  if (o->kind == NIMFM_DS_CSR && nnz > 0) {
    // a CSR made on the device has no host copy to sample: find the hot columns from the CSC side
    // (column length >= n/16), the 16 longest
    std::vector<int64_t> ptr((size_t)in->d + 1);
    CK(cudaMemcpy(ptr.data(), in->indptr, ((size_t)in->d + 1) * 8, cudaMemcpyDeviceToHost));
    std::vector<std::pair<int64_t, int64_t>> cand;
    for (int64_t j = 0; j < in->d; j++) {
      const int64_t len = ptr[j + 1] - ptr[j];
      if (len * 16 >= in->n && len >= 2) cand.push_back({len, j});
    }
    std::sort(cand.begin(), cand.end(), [](const std::pair<int64_t, int64_t> &x, const std::pair<int64_t, int64_t> &y) {
      return x.first != y.first ? x.first > y.first : x.second < y.second;
    });
    std::vector<int32_t> hot;
    for (size_t i = 0; i < cand.size() && i < 16; i++) hot.push_back((int32_t)cand[i].second);
    o->nHot = (int)hot.size();
    if ((rc = nimfm_upload_hot(ctx, hot, o->d, &o->hotSlot, &o->hotList))) return rc;
  }
  *out = guard.release();
  return NIMFM_OK;
}

}  // extern "C"

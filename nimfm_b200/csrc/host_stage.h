// host_stage.h -- host-side preparation of streamed CSR chunks (the end-to-end path of fm_api.cu).
//
// The reference keeps column ids as Nim `int` (int64, tensor/sparse.nim:4-31); the device kernels read
// int32.  Narrowing on the HOST, into pinned staging slots, before the H2D copy takes 4 of the 16 bytes a
// nonzero costs on PCIe (the link is the bound of the end-to-end path).  A small team of threads runs
// ahead of the copy engine: chunk c+1, c+2 are narrowed while chunk c is in flight.  The same pass checks
// the index range, rebases the chunk's indptr to 0 and finds its longest / shortest row.
#pragma once
#include <stdint.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// int64 -> int32 (AVX2 when the CPU has it); dst 32-byte aligned; nonzero if any id is outside [0, d)
int nimfm_host_narrow(const int64_t *src, int32_t *dst, int64_t n, int64_t d);

struct HostChunk {
  int64_t r0 = 0, r1 = 0;     // rows [r0, r1) of the caller's CSR
  int64_t base = 0, nnz = 0;  // indptr[r0], indptr[r1]-indptr[r0]
};

struct HostChunkInfo {
  int64_t maxSeg = 0, minSeg = 0;
  int bad = 0;                // some column id outside [0,d)
  int64_t packed[32] = {0};   // pack mode: values != 1.0 written by thread t at the start of its slice
};

class HostStageTeam {
 public:
  static constexpr int kSlots = 4;
  // idxSlot[s] (int32, >= max chunk nnz) and ptrSlot[s] (int64, >= max chunk rows + 1): pinned, 32-byte aligned
  HostStageTeam(int nThreads, const int64_t *indices, const int64_t *indptr, int64_t d,
                const std::vector<HostChunk> &chunks, int32_t *const *idxSlot, int64_t *const *ptrSlot);
  // PAGEABLE caller arrays (a Nim seq, a numpy array): the team also copies each chunk's values (and targets,
  // when y != nullptr) into pinned slots, so that every byte the copy engine reads is page-locked -- a
  // cudaMemcpyAsync from pageable memory is staged by the driver on one thread at ~4 GB/s.  Call before allow().
  void stage_values(const double *data, const double *y, double *const *dataSlot, double *const *ySlot);
  // LOSSLESS transport packing of the values: sparse FM data is mostly one-hot (26 of the 39 values of a
  // Criteo-shaped row are exactly 1.0), so a chunk travels as a bit mask "value == 1.0" (1 bit per nonzero), the
  // other values packed per thread slice, and one uint32 offset per 256 nonzeros that lets the device expand
  // without a scan (fm_api.cu, expand_values_kernel).  Every double that is not bit-for-bit 1.0 crosses the link
  // unchanged.  Call before allow(); slices are multiples of 256 nonzeros (slice_len()).
  void pack_values(const double *data, const double *y, double *const *packSlot, uint64_t *const *maskSlot,
                   uint32_t *const *blkSlot, double *const *ySlot);
  int64_t slice_len(int64_t nnz) const { return ((nnz + T_ - 1) / T_ + 255) & ~(int64_t)255; }
  int threads() const { return T_; }
  ~HostStageTeam();                       // stops and joins the workers
  void allow(int64_t upTo);               // chunks < upTo may be written (their slot's last copy has completed)
  HostChunkInfo wait(int64_t c);          // blocks until chunk c is staged in slot c % kSlots
  static int default_threads(int nRanks); // 0: not enough host threads for this rank -> narrow on the device
  static int pageable_threads(int nRanks);   // team size when the values are staged too

 private:
  void work(int t);
  void part(int64_t c, int t, HostChunkInfo &info);
  const int T_;
  const int64_t *indices_, *indptr_;
  const int64_t d_;
  const std::vector<HostChunk> &chunks_;
  int32_t *const *idxSlot_;
  int64_t *const *ptrSlot_;
  const double *data_ = nullptr, *y_ = nullptr;
  double *const *dataSlot_ = nullptr;
  double *const *ySlot_ = nullptr;
  double *const *packSlot_ = nullptr;
  uint64_t *const *maskSlot_ = nullptr;
  uint32_t *const *blkSlot_ = nullptr;
  std::mutex mu_;
  std::condition_variable cv_;
  int64_t allowed_ = 0;
  bool stop_ = false;
  std::vector<int> done_;
  std::vector<HostChunkInfo> info_;
  std::vector<std::thread> threads_;
};

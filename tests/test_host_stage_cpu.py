"""The host staging team of the host-fed calls (nimfm_b200/csrc/host_stage.cpp: int64 -> int32 narrowing with
range check, indptr rebasing, longest / shortest row per chunk, the slot hand-over protocol) without a GPU:
the file is compiled with a small driver that plays the library's main loop (wait chunk c, "copy" it out of
its pinned slot, release the slot of chunk c-2) and the result is compared with numpy."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nimfm_b200", "csrc")

DRIVER = r"""
#include "host_stage.h"
#include <string.h>
#include <stdlib.h>
#include <algorithm>
extern "C" int run_team(int threads, const int64_t *indices, const int64_t *indptr, int64_t nRows, int64_t d,
                        int64_t chunkRows, int32_t *outIdx, int64_t *outPtr, int64_t *outMaxSeg, int64_t *outMinSeg,
                        int *outBad) {
  std::vector<HostChunk> chunks;
  size_t maxNnz = 1;
  for (int64_t r0 = 0; r0 < nRows; r0 += chunkRows) {
    HostChunk ch; ch.r0 = r0; ch.r1 = std::min(nRows, r0 + chunkRows);
    ch.base = indptr[ch.r0]; ch.nnz = indptr[ch.r1] - ch.base;
    maxNnz = std::max(maxNnz, (size_t)ch.nnz);
    chunks.push_back(ch);
  }
  int32_t *idxSlot[HostStageTeam::kSlots]; int64_t *ptrSlot[HostStageTeam::kSlots];
  for (int s = 0; s < HostStageTeam::kSlots; s++) {
    idxSlot[s] = (int32_t *)aligned_alloc(64, ((maxNnz * 4 + 63) / 64) * 64);
    ptrSlot[s] = (int64_t *)aligned_alloc(64, (((size_t)chunkRows + 1) * 8 + 63) / 64 * 64);
    memset(idxSlot[s], 0xff, ((maxNnz * 4 + 63) / 64) * 64);
  }
  const int64_t nChunks = (int64_t)chunks.size();
  *outBad = 0;
  {
    HostStageTeam team(threads, indices, indptr, d, chunks, idxSlot, ptrSlot);
    team.allow(std::min<int64_t>(nChunks, HostStageTeam::kSlots - 1));
    for (int64_t c = 0; c < nChunks; c++) {
      HostChunkInfo info = team.wait(c);
      const int s = (int)(c % HostStageTeam::kSlots);
      memcpy(outIdx + chunks[c].base - indptr[0], idxSlot[s], (size_t)chunks[c].nnz * 4);
      memcpy(outPtr + c * (chunkRows + 1), ptrSlot[s], (size_t)(chunks[c].r1 - chunks[c].r0 + 1) * 8);
      memset(idxSlot[s], 0xff, (size_t)chunks[c].nnz * 4);     // a late writer into a released slot would show
      outMaxSeg[c] = info.maxSeg; outMinSeg[c] = info.minSeg; *outBad |= info.bad;
      team.allow(c + 3);
    }
  }
  for (int s = 0; s < HostStageTeam::kSlots; s++) { free(idxSlot[s]); free(ptrSlot[s]); }
  return (int)nChunks;
}
// pack mode: the team packs the values of every chunk; the driver concatenates per chunk [mask words | blk offsets |
// packed region of the whole slot] so that Python can expand them the way expand_values_kernel does
extern "C" int run_pack(int threads, const double *data, const int64_t *indices, const int64_t *indptr, int64_t nRows,
                        int64_t d, int64_t chunkRows, double *outPacked, uint64_t *outMask, uint32_t *outBlk,
                        int64_t *outPackedCount) {
  std::vector<HostChunk> chunks;
  size_t maxNnz = 1;
  for (int64_t r0 = 0; r0 < nRows; r0 += chunkRows) {
    HostChunk ch; ch.r0 = r0; ch.r1 = std::min(nRows, r0 + chunkRows);
    ch.base = indptr[ch.r0]; ch.nnz = indptr[ch.r1] - ch.base;
    maxNnz = std::max(maxNnz, (size_t)ch.nnz);
    chunks.push_back(ch);
  }
  int32_t *idxSlot[HostStageTeam::kSlots]; int64_t *ptrSlot[HostStageTeam::kSlots];
  double *packSlot[HostStageTeam::kSlots]; uint64_t *maskSlot[HostStageTeam::kSlots]; uint32_t *blkSlot[HostStageTeam::kSlots];
  for (int s = 0; s < HostStageTeam::kSlots; s++) {
    idxSlot[s] = (int32_t *)aligned_alloc(64, ((maxNnz * 4 + 63) / 64) * 64);
    ptrSlot[s] = (int64_t *)aligned_alloc(64, (((size_t)chunkRows + 1) * 8 + 63) / 64 * 64);
    packSlot[s] = (double *)aligned_alloc(64, (((maxNnz + 8) * 8 + 63) / 64) * 64);
    maskSlot[s] = (uint64_t *)aligned_alloc(64, ((maxNnz / 64 + 8) * 8 + 63) / 64 * 64);
    blkSlot[s] = (uint32_t *)aligned_alloc(64, ((maxNnz / 256 + 8) * 4 + 63) / 64 * 64);
  }
  const int64_t nChunks = (int64_t)chunks.size();
  {
    HostStageTeam team(threads, indices, indptr, d, chunks, idxSlot, ptrSlot);
    team.pack_values(data, nullptr, packSlot, maskSlot, blkSlot, nullptr);
    team.allow(std::min<int64_t>(nChunks, HostStageTeam::kSlots - 1));
    for (int64_t c = 0; c < nChunks; c++) {
      HostChunkInfo info = team.wait(c);
      const int s = (int)(c % HostStageTeam::kSlots);
      const int64_t nnz = chunks[c].nnz, off = chunks[c].base - indptr[0], per = team.slice_len(nnz);
      // what the library copies: every thread's packed region, the mask words, the block offsets
      int64_t total = 0;
      for (int t = 0; t < team.threads(); t++) {
        const int64_t a0 = std::min(nnz, per * t);
        memcpy(outPacked + off + a0, packSlot[s] + a0, (size_t)info.packed[t] * 8);
        total += info.packed[t];
      }
      outPackedCount[c] = total;
      memcpy(outMask + c * (maxNnz / 64 + 8), maskSlot[s], (size_t)((nnz + 63) / 64) * 8);
      memcpy(outBlk + c * (maxNnz / 256 + 8), blkSlot[s], (size_t)((nnz + 255) / 256) * 4);
      team.allow(c + 3);
    }
  }
  for (int s = 0; s < HostStageTeam::kSlots; s++) { free(idxSlot[s]); free(ptrSlot[s]); free(packSlot[s]); free(maskSlot[s]); free(blkSlot[s]); }
  return (int)nChunks;
}
extern "C" int narrow_only(const int64_t *src, int32_t *dst, int64_t n, int64_t d) { return nimfm_host_narrow(src, dst, n, d); }
extern "C" int default_threads(int nRanks) { return HostStageTeam::default_threads(nRanks); }
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("host_stage")
    src = d / "driver.cpp"
    src.write_text(DRIVER)
    so = d / "libhs.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", CSRC, str(src),
                           os.path.join(CSRC, "host_stage.cpp"), "-o", str(so)])
    return C.CDLL(str(so))


def ragged(n, d, seed, zmax):
    rng = np.random.default_rng(seed)
    z = rng.integers(0, zmax + 1, n)
    z[::7] = 0
    indptr = np.concatenate([[0], np.cumsum(z)]).astype(np.int64)
    indices = rng.integers(0, d, int(indptr[-1])).astype(np.int64)
    return indices, indptr


@pytest.mark.parametrize("threads", [1, 2, 3, 8])
@pytest.mark.parametrize("chunk", [1, 7, 64, 5000])
def test_team_matches_numpy(lib, threads, chunk):
    n, d = 1000, 100_000
    indices, indptr = ragged(n, d, 3 + threads, 40)
    nChunks = (n + chunk - 1) // chunk
    outIdx = np.zeros(len(indices), np.int32)
    outPtr = np.zeros(nChunks * (chunk + 1), np.int64)
    mx, mn, bad = np.zeros(nChunks, np.int64), np.zeros(nChunks, np.int64), C.c_int()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    got = lib.run_team(threads, P(indices), P(indptr), C.c_int64(n), C.c_int64(d), C.c_int64(chunk), P(outIdx), P(outPtr),
                       P(mx), P(mn), C.byref(bad))
    assert got == nChunks and bad.value == 0
    assert np.array_equal(outIdx, indices.astype(np.int32))
    for c in range(nChunks):
        a, b = c * chunk, min(n, (c + 1) * chunk)
        assert np.array_equal(outPtr[c * (chunk + 1): c * (chunk + 1) + (b - a + 1)], indptr[a:b + 1] - indptr[a])
        lens = np.diff(indptr[a:b + 1])
        assert mx[c] == lens.max() and mn[c] == min(0, lens.min())


@pytest.mark.parametrize("pos,val", [(0, -1), (17, 100_000), (999, 2**40), (1234, -2**62), (4000, 100_001)])
def test_bad_index_is_flagged(lib, pos, val):
    n, d = 500, 100_000
    indices, indptr = ragged(n, d, 9, 30)
    indices = indices.copy()
    pos = pos % len(indices)
    indices[pos] = val
    outIdx = np.zeros(len(indices), np.int32)
    outPtr = np.zeros(8 * 65, np.int64)
    mx, mn, bad = np.zeros(8, np.int64), np.zeros(8, np.int64), C.c_int()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.run_team(4, P(indices), P(indptr), C.c_int64(n), C.c_int64(d), C.c_int64(64), P(outIdx), P(outPtr), P(mx), P(mn), C.byref(bad))
    assert bad.value == 1


def test_decreasing_indptr_is_reported(lib):
    indptr = np.array([0, 3, 2, 6], np.int64)
    indices = np.arange(6, dtype=np.int64)
    outIdx, outPtr = np.zeros(6, np.int32), np.zeros(8, np.int64)
    mx, mn, bad = np.zeros(1, np.int64), np.zeros(1, np.int64), C.c_int()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.run_team(2, P(indices), P(indptr), C.c_int64(3), C.c_int64(10), C.c_int64(4), P(outIdx), P(outPtr), P(mx), P(mn), C.byref(bad))
    assert mn[0] == -1          # the caller turns a negative row length into "indptr is not monotone"


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 31, 1000, 4099])
def test_narrow_tail_and_alignment(lib, n):
    rng = np.random.default_rng(n)
    src = rng.integers(0, 2**31 - 1, n + 3).astype(np.int64)[3:]        # an unaligned source is fine
    dst = np.zeros(n + 16, np.int32)
    off = (-dst.ctypes.data // 4) % 8                                    # the destination must be 32-byte aligned
    view = dst[off:off + n]
    assert lib.narrow_only(src.ctypes.data_as(C.c_void_p), view.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int64(2**31 - 1)) == 0
    assert np.array_equal(view, src.astype(np.int32))
    if n:
        assert lib.narrow_only(src.ctypes.data_as(C.c_void_p), view.ctypes.data_as(C.c_void_p), C.c_int64(n),
                               C.c_int64(int(src.max()))) == 1     # d == max id: that id is out of range


def test_default_threads_gate(lib, monkeypatch):
    monkeypatch.delenv("NIMFM_HOST_THREADS", raising=False)
    hw = len(os.sched_getaffinity(0))
    assert lib.default_threads(1) == (8 if hw >= 8 else 0)
    assert lib.default_threads(max(hw, 1)) == 0                          # one core per rank: narrow on the device
    monkeypatch.setenv("NIMFM_HOST_THREADS", "3")
    assert lib.default_threads(64) == 3


@pytest.mark.parametrize("threads", [1, 3, 8, 14])
@pytest.mark.parametrize("chunk,frac_ones", [(5000, 0.66), (300, 0.3), (64, 1.0), (5000, 0.0)])
def test_pack_values_roundtrip(lib, threads, chunk, frac_ones):
    """the lossless packed transport of the values (bit mask "== 1.0" + the other values + one offset per 256 nonzeros):
    expanding it the way expand_values_kernel does gives back every double bit for bit -- NaN, -0.0, denormals and
    values next to 1.0 included"""
    n, d = 2000, 100_000
    indices, indptr = ragged(n, d, 11 + threads, 40)
    nnz = int(indptr[-1])
    rng = np.random.default_rng(chunk)
    data = rng.standard_normal(nnz)
    data[rng.random(nnz) < frac_ones] = 1.0
    special = np.array([np.nan, -0.0, 5e-324, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), np.inf, -1.0])
    data[rng.integers(0, nnz, 50)] = special[rng.integers(0, len(special), 50)]
    nChunks = (n + chunk - 1) // chunk
    maxNnz = max(int(indptr[min(n, (c + 1) * chunk)] - indptr[c * chunk]) for c in range(nChunks))
    W, B = maxNnz // 64 + 8, maxNnz // 256 + 8
    outPacked = np.zeros(nnz + 16)
    outMask, outBlk = np.zeros(nChunks * W, np.uint64), np.zeros(nChunks * B, np.uint32)
    counts = np.zeros(nChunks, np.int64)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    got = lib.run_pack(threads, P(data), P(indices), P(indptr), C.c_int64(n), C.c_int64(d), C.c_int64(chunk), P(outPacked),
                       P(outMask), P(outBlk), P(counts))
    assert got == nChunks
    for c in range(nChunks):
        a, b = int(indptr[c * chunk]), int(indptr[min(n, (c + 1) * chunk)])
        want = data[a:b]
        m = b - a
        assert counts[c] == int(np.sum(~(want == 1.0)))
        words = outMask[c * W: c * W + (m + 63) // 64]
        bits = ((words[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool).ravel()[:m]
        assert np.array_equal(bits, want == 1.0)
        # expand: index of packed value q = blk[q // 256] + number of non-ones before q inside its block
        q = np.arange(m)
        nonone = ~bits
        before = np.cumsum(nonone) - nonone                   # non-ones before q in the chunk
        blk_start = (q // 256) * 256
        inblock = before - before[blk_start]                  # ... inside q's block (before[blk_start] counts up to the block)
        idx = outBlk[c * B + q // 256].astype(np.int64) + inblock
        out = np.where(bits, 1.0, outPacked[a + idx])
        assert np.array_equal(out.view(np.uint64), want.view(np.uint64))

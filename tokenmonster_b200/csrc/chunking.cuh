// chunking.cuh -- intra-document parallelism for few-but-huge inputs.
//
// One string is one lane in the tokenize kernel, so a single 1 MB document would run at
// single-lane speed.  When a batch is too small to fill the GPU, documents longer than
// kChunkMin bytes are cut into chunks of ~kChunkBytes:
//
//   plan     each chunk starts right after an ASCII character that is neither a letter, a digit,
//            a space nor an apostrophe.  There the capcode encoder is in its initial state
//            (javascript/tokenmonster.js:900-1005: inWord is false, rlast is "other", rlast2 only
//            matters behind an apostrophe), so the chunk's normalized text is exactly the
//            corresponding slice of the document's normalized text.
//   tokenize tokenize_thread_kernel<.., CHUNK = true>: every chunk assumes that a token starts at
//            its first byte (the ungreedy loop's state at a token start is a function of the
//            position and the forward-delete flag alone: the current token is the longest match
//            there, go/tokenmonster.go:1049,1068,1091), runs kChunkOverlap normalized bytes past
//            the start of the next chunk and logs (position, forward-delete) for every step.
//   stitch   per document, the step where chunk k's log meets chunk k+1's log (same position, same
//            flag) is the hand-over point: from there on both sequences are identical, so chunk k
//            keeps its ids before that step and chunk k+1 its ids from that step on.  If two logs
//            never meet inside the overlap (or a log overflowed) the document is tokenized again
//            by a single lane -- exactness never depends on the speculation.
//   gather   the kept id ranges are copied into the document's ordinary slab slot; scan + pack
//            then proceed as for any other string.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace tmgpu {

constexpr uint32_t kChunkBytes = 1024;     // nominal raw bytes per chunk
constexpr uint32_t kChunkMin = 4 * kChunkBytes;
constexpr uint32_t kChunkOverlap = 256;    // normalized bytes a chunk runs into its successor
// ids logged per chunk (a chunk spans up to 1.5 x kChunkBytes raw bytes; text that yields more ids
// than this -- alternating case under capcode -- sends the document to the one-lane path)
constexpr uint32_t kChunkLogCap = kChunkBytes + kChunkBytes / 2 + kChunkOverlap + 64;

__device__ __forceinline__ bool chunk_safe_after(uint32_t c) {
  // ASCII, not alphanumeric, not a space, not an apostrophe: the capcode encoder is "reset" behind it
  return c < 0x80u && !((c | 0x20u) - 'a' < 26u) && !(c - '0' < 10u) && c != ' ' && c != '\'';
}

// chunks per document: 0 = tokenize it as one string
__global__ void __launch_bounds__(256) chunk_count_kernel(const int64_t* __restrict__ off,
                                                          const int64_t* __restrict__ ends, uint32_t n,
                                                          uint32_t* __restrict__ n_chunks, uint8_t* __restrict__ skip) {
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x) {
    const int64_t len = (ends ? ends[d] : off[d + 1]) - off[d];
    const uint32_t k = len > int64_t(kChunkMin) ? uint32_t((len + kChunkBytes - 1) / kChunkBytes) : 0u;
    n_chunks[d] = k;
    skip[d] = k ? 1 : 0;
  }
}

__device__ __forceinline__ int64_t chunk_start(const uint8_t* data, int64_t b0, int64_t e0, uint32_t k) {
  if (k == 0) return b0;
  const int64_t nominal = b0 + int64_t(k) * kChunkBytes;
  if (nominal >= e0) return e0;
  const int64_t lim = min(nominal + int64_t(kChunkBytes / 2), e0);
  for (int64_t p = nominal; p < lim; p++)
    if (chunk_safe_after(data[p - 1])) return p;
  return -1;   // no safe boundary in reach: this document is not chunked
}

// one thread per chunk: its document (binary search in the exclusive scan of n_chunks) and bounds
__global__ void __launch_bounds__(256) chunk_fill_kernel(const uint8_t* __restrict__ data,
                                                         const int64_t* __restrict__ off,
                                                         const int64_t* __restrict__ ends, uint32_t n,
                                                         const uint32_t* __restrict__ n_chunks,
                                                         const uint64_t* __restrict__ chunk_base, uint32_t total,
                                                         ChunkItem* __restrict__ items, uint8_t* __restrict__ doc_fail) {
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = n;                     // last d with chunk_base[d] <= c and n_chunks[d] > 0
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) / 2;
      if (chunk_base[mid] <= c) lo = mid; else hi = mid;
    }
    const uint32_t d = lo;
    const uint32_t k = c - uint32_t(chunk_base[d]);
    const int64_t b0 = off[d], e0 = ends ? ends[d] : off[d + 1];
    const bool last = (k + 1 == n_chunks[d]);
    const int64_t s0 = chunk_start(data, b0, e0, k);
    const int64_t s1 = last ? e0 : chunk_start(data, b0, e0, k + 1);
    if (s0 < 0 || s1 < 0) doc_fail[d] = 1;
    ChunkItem it;
    it.start = s0 < 0 ? b0 : s0;
    it.next = s1 < 0 ? e0 : s1;
    it.doc = d;
    it.flags = last ? 1u : 0u;
    items[c] = it;
  }
}

// ---- stitch, in three small kernels --------------------------------------------------------------
// A  (one thread per chunk)  where this chunk's log meets its successor's
// B  (one warp per chunk)    the kept range [begin, end) of the chunk and its tallies
// C  (one thread per document) running offsets, totals, verdict
constexpr uint32_t kMeetNone = 0xFFFFFFFFu;   // the two logs never met
constexpr uint32_t kMeetEnd = 0xFFFFFFFEu;    // no successor needed (last chunk, or ran to the document's end)

__global__ void __launch_bounds__(128) chunk_meet_kernel(const StitchArgs A, uint32_t total_chunks,
                                                         const ChunkItem* __restrict__ items) {
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total_chunks; c += gridDim.x * blockDim.x) {
    const ChunkOut co = A.c_out[c];
    uint32_t ma = kMeetNone, mb = 0;
    if ((items[c].flags & 1u) || (co.flags & 4u)) {
      ma = kMeetEnd;
    } else if ((co.flags & 3u) == 0) {
      const ChunkOut cn = A.c_out[c + 1];
      if ((cn.flags & 3u) == 0) {
        const uint32_t* pos = A.c_pos + uint64_t(c) * A.c_cap;
        const uint32_t* pn = A.c_pos + uint64_t(c + 1) * A.c_cap;
        // the steps at or beyond the successor's start are at the tail of the log: walk back to the first
        uint32_t a = co.cnt;
        while (a > 0) {
          const uint32_t p = pos[a - 1];
          if (p != 0xFFFFFFFFu && (p & 0x7FFFFFFFu) < co.n_norm) break;
          a--;
        }
        uint32_t b = 0;
        while (a < co.cnt && b < cn.cnt) {   // two pointers over (position relative to the successor, fd)
          const uint32_t pa = pos[a], pb = pn[b];
          if (pa == 0xFFFFFFFFu) { a++; continue; }
          if (pb == 0xFFFFFFFFu) { b++; continue; }
          const uint32_t ra = (pa & 0x7FFFFFFFu) - co.n_norm, xb = pb & 0x7FFFFFFFu;
          if (ra == xb && (pa >> 31) == (pb >> 31)) { ma = a; mb = b; break; }
          if (ra < xb || (ra == xb && (pa >> 31) < (pb >> 31))) a++; else b++;
        }
      }
    }
    A.meet_a[c] = ma;
    A.meet_b[c] = mb;
  }
}

__global__ void __launch_bounds__(256) chunk_segment_kernel(const StitchArgs A, uint32_t total_chunks,
                                                            const ChunkItem* __restrict__ items) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < total_chunks; c += warps) {
    const ChunkOut co = A.c_out[c];
    const uint32_t d = items[c].doc;
    const bool first = (uint64_t(c) == A.chunk_base[d]);
    const uint32_t ma = A.meet_a[c];
    uint32_t flag = (co.flags & 1u) ? 1u : ((co.flags & 2u) ? 2u : 0u);   // 1 non-ASCII, 2 cannot stitch
    uint32_t begin = 0, end = co.cnt;
    if (!first) {
      if (A.meet_a[c - 1] == kMeetNone) flag |= 2u; else begin = A.meet_b[c - 1];
    }
    if (ma == kMeetNone) flag |= 2u;
    else if (ma != kMeetEnd) end = ma;
    if (end < begin) flag |= 2u;
    uint32_t dels = 0;
    const uint32_t* pos = A.c_pos + uint64_t(c) * A.c_cap;
    if (!(flag & 2u) && A.count_mode)
      for (uint32_t t = begin + lane; t < end; t += 32) dels += (pos[t] == 0xFFFFFFFFu);
#pragma unroll
    for (int o = 16; o; o >>= 1) dels += __shfl_xor_sync(0xFFFFFFFFu, dels, o);
    if (lane == 0) {
      const uint32_t* cm = A.c_miss + uint64_t(c) * A.c_cap;
      uint32_t miss = 0;
      if (!(flag & 2u)) miss = ((end < co.cnt) ? cm[end] : co.miss) - ((begin < co.cnt) ? cm[begin] : co.miss);
      A.seg_begin[c] = begin;
      A.seg_end[c] = (flag & 2u) ? begin : end;
      A.seg_dels[c] = dels;
      A.seg_miss[c] = miss;
      A.seg_flag[c] = flag | (ma == kMeetEnd ? 4u : 0u);
    }
  }
}

__global__ void __launch_bounds__(128) chunk_finish_kernel(const StitchArgs A) {
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < A.n; d += gridDim.x * blockDim.x) {
    const uint32_t nc = A.n_chunks[d];
    if (nc == 0) continue;
    const uint64_t c0 = A.chunk_base[d];
    bool fail = A.doc_fail[d] != 0, bad = false, done = false;
    uint64_t total = 0, dels = 0, miss = 0;
    for (uint32_t k = 0; k < nc; k++) {
      const uint64_t c = c0 + k;
      if (done || fail || bad) { A.seg_end[c] = A.seg_begin[c]; continue; }   // dropped: nothing to gather
      const uint32_t f = A.seg_flag[c];
      if (f & 1u) { bad = true; A.seg_end[c] = A.seg_begin[c]; continue; }
      if (f & 2u) { fail = true; A.seg_end[c] = A.seg_begin[c]; continue; }
      A.seg_dst[c] = uint32_t(total);
      total += A.seg_end[c] - A.seg_begin[c];
      dels += A.seg_dels[c];
      miss += A.seg_miss[c];
      if (f & 4u) done = true;
    }
    const int64_t len = (A.ends ? A.ends[d] : A.off[d + 1]) - A.off[d];
    const uint64_t cap = uint64_t(len) + A.slab_pad;
    const bool too_big = !A.count_mode && total > cap;   // more ids than the slab slot holds: exact-size re-run
    if (bad || fail || too_big) {
      for (uint32_t k = 0; k < nc; k++) A.seg_end[c0 + k] = A.seg_begin[c0 + k];
    }
    if (bad) {              // non-ASCII text: status 13 here, the non-ASCII pass picks the document up
      A.counts[d] = 0; A.need[d] = 0; A.missing[d] = 0; A.status[d] = 13;
      A.doc_fail[d] = 0;
    } else if (fail) {
      A.doc_fail[d] = 1;
    } else {
      A.doc_fail[d] = 0;
      const uint64_t value = A.count_mode ? total - dels : total;
      A.need[d] = uint32_t(value);
      A.missing[d] = uint32_t(miss);
      A.status[d] = 0;
      A.counts[d] = too_big ? uint32_t(cap) : uint32_t(value);
      if (too_big) atomicExch(A.overflow, 1u);
    }
  }
}

// one warp per chunk: its kept ids -> the document's slab slot
__global__ void __launch_bounds__(256) chunk_gather_kernel(const ChunkItem* __restrict__ items, uint32_t total,
                                                           const uint32_t* __restrict__ c_ids, uint32_t c_cap,
                                                           const uint32_t* __restrict__ seg_begin,
                                                           const uint32_t* __restrict__ seg_end,
                                                           const uint32_t* __restrict__ seg_dst,
                                                           const int64_t* __restrict__ off, uint32_t slab_pad,
                                                           uint32_t* __restrict__ slab) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < total; c += warps) {
    const uint32_t b = seg_begin[c], e = seg_end[c];
    if (e <= b) continue;
    const uint32_t d = items[c].doc;
    const uint32_t* src = c_ids + uint64_t(c) * c_cap;
    uint32_t* dst = slab + uint64_t(off[d] - off[0]) + uint64_t(slab_pad) * d + seg_dst[c];
    for (uint32_t t = b + lane; t < e; t += 32) dst[t - b] = src[t];
  }
}

// documents whose chunks could not be stitched: list them for a one-lane launch
__global__ void __launch_bounds__(256) chunk_failed_kernel(const uint8_t* __restrict__ doc_fail,
                                                           const uint32_t* __restrict__ n_chunks,
                                                           const int64_t* __restrict__ off,
                                                           const int64_t* __restrict__ ends, uint32_t n,
                                                           uint32_t* __restrict__ list, int64_t* __restrict__ f_begin,
                                                           int64_t* __restrict__ f_end, unsigned long long* __restrict__ ctr) {
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x) {
    if (!n_chunks[d] || !doc_fail[d]) continue;
    const unsigned long long k = atomicAdd(ctr, 1ull);
    list[k] = d;
    f_begin[k] = off[d];
    f_end[k] = ends ? ends[d] : off[d + 1];
  }
}

}  // namespace tmgpu

// kernels.cu -- device code of libtmgpu.so (sm_100a only) and its launchers.
//   K1/K3 prep / table / chain    the position-table pipeline (pipeline_kernels.cuh): capcode
//                                 encoder + reversed-key automaton right to left, then the ungreedy
//                                 loop; ids, counts, training scores; raw and pre-normalized text
//   tokenize_thread_kernel        round-1 trie-walk kernel, chunk launch for long documents only
//                                 (tokenize_thread_kernel.cuh, chunking.cuh)
//   unicode_normalize_kernel      NFD / accents / lowercase / capcode per code point
//                                 (unicode_normalize.cuh)
//   chunk_* kernels               plan, stitch and gather for long documents (chunking.cuh)
//   pack_tokens_kernel            slab of u32 ids -> dense 2/3/4-byte serialized ids
//                                 (the emit variants go/tokenmonster.go:1545/:1817/:2089)
//   wire_frame_kernel             reply framing of the server protocol
//   K2 decode_size / decode_copy  reverse[id] gather (go/tokenmonster.go:813-948)
//   capcode_decode_*_kernel       capcode level-2 decoder (javascript/tokenmonster.js:1007-1065):
//                                 ASCII strings as a warp scan over the decoder's transition monoid
//                                 (capcode_decode_scan.cuh), the rest per code point
//   stream_tail / copy_tail       the streaming decoder's incomplete-sequence remainder
#include "kernels.h"

#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "pipeline_kernels.cuh"
#include "chunk_kernel_helpers.cuh"
#include "tokenize_thread_kernel.cuh"
#include "unicode_normalize.cuh"
#include "chunking.cuh"
#include "capcode_decode_scan.cuh"

namespace tmgpu {

// ------------------------------------------------------------------------------------------
// pack: one warp per string copies its ids out of the slab into the dense serialized stream
// ------------------------------------------------------------------------------------------
template <int ENC>
__global__ void __launch_bounds__(256) pack_tokens_kernel(const uint32_t* __restrict__ slab,
                                                          const int64_t* __restrict__ off, uint32_t slab_pad,
                                                          const uint64_t* __restrict__ slab_off,
                                                          const uint32_t* __restrict__ counts,
                                                          const uint64_t* __restrict__ tok_off,
                                                          uint8_t* __restrict__ out, uint32_t n) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const uint64_t base = slab_off ? slab_off[s] : uint64_t(off[s] - off[0]) + uint64_t(slab_pad) * s;
    const uint32_t cnt = counts[s];
    const uint64_t o = tok_off[s];
    for (uint32_t k = lane; k < cnt; k += 32) {
      const uint32_t id = slab[base + k];
      if (ENC == 2) {
        reinterpret_cast<uint16_t*>(out)[o + k] = uint16_t(id);
      } else if (ENC == 4) {
        reinterpret_cast<uint32_t*>(out)[o + k] = id;
      } else {
        uint8_t* p = out + (o + k) * 3;
        p[0] = uint8_t(id);
        p[1] = uint8_t(id >> 8);
        p[2] = uint8_t(id >> 16);
      }
    }
  }
}

// Chunked host path: the ids of chunk k go to the host while chunk k+1 is still being tokenized, so
// every chunk turns its local offsets (exclusive scan of its own counts) into global ones itself:
// base_io[0] = ids before this chunk (written by its predecessor), base_io[1] = ids including it.
__global__ void __launch_bounds__(256) chunk_offsets_kernel(uint64_t* __restrict__ tok_off,
                                                            const uint32_t* __restrict__ counts, uint32_t n,
                                                            unsigned long long* __restrict__ base_io, int write_end) {
  const unsigned long long base = base_io[0];
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const unsigned long long local = tok_off[k];
    tok_off[k] = local + base;
    if (k == n - 1) {
      const unsigned long long end = base + local + counts[k];
      if (write_end) tok_off[n] = end;   // (an inner chunk's end is the next chunk's first entry, which that chunk writes)
      base_io[1] = end;
    }
  }
}

cudaError_t launch_chunk_offsets(uint64_t* tok_off, const uint32_t* counts, uint32_t n, unsigned long long* base_io,
                                 bool write_end, int sms, cudaStream_t st) {
  const long long want = (static_cast<long long>(n) + 255) / 256;
  chunk_offsets_kernel<<<int(want < sms * 4ll ? (want < 1 ? 1 : want) : sms * 4ll), 256, 0, st>>>(tok_off, counts, n, base_io,
                                                                                                 write_end ? 1 : 0);
  return cudaGetLastError();
}

// Reply framing of the server protocol (training/tokenmonsterserver.go:382-392, :435-445):
// [u32 n] then, per string, [u64 byte length][payload].  One warp per string; `unit` is the
// item size of src_off (the encoding length for ids, 1 for decoded bytes).
__global__ void __launch_bounds__(256) wire_frame_kernel(const uint8_t* __restrict__ src,
                                                         const uint64_t* __restrict__ src_off, uint32_t unit,
                                                         uint32_t n, uint8_t* __restrict__ dst) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w0 == 0 && lane < 4) dst[lane] = uint8_t(n >> (8 * lane));
  for (uint32_t s = w0; s < n; s += warps) {
    const uint64_t b0 = src_off[s] * unit, len = (src_off[s + 1] - src_off[s]) * unit;
    uint8_t* d = dst + 4 + 8ull * (s + 1) + b0;
    if (lane < 8) d[int(lane) - 8] = uint8_t(len >> (8 * lane));
    const uint8_t* a = src + b0;
    for (uint64_t k = lane; k < len; k += 32) d[k] = a[k];
  }
}

// ------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_id(const uint8_t* t, uint64_t k, int enc) {
  if (enc == 2) return reinterpret_cast<const uint16_t*>(t)[k];
  if (enc == 4) return reinterpret_cast<const uint32_t*>(t)[k];
  const uint8_t* p = t + k * 3;
  return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16);
}

// reverse[id] as one word: offset into rev_blob8 | length << 24; 0 for out-of-range / nil ids.  The
// 16-bit path compares against uint16(len(reverse) - 1) like go/tokenmonster.go:861-864.
__device__ __forceinline__ uint32_t rev_entry(const VocabDev& V, uint32_t id, int enc) {
  const bool ok = (enc == 2) ? (V.n_reverse != 0 && id <= ((V.n_reverse - 1) & 0xFFFFu)) : (id < V.n_reverse);
  return ok ? __ldg(&V.rev_pack[id]) : 0u;
}

__global__ void __launch_bounds__(256) decode_size_kernel(VocabDev V, const uint8_t* __restrict__ tokens,
                                                          const int64_t* __restrict__ tok_off, uint32_t n, int enc,
                                                          uint32_t* __restrict__ raw_len,
                                                          const uint64_t* __restrict__ prefix_off) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const uint64_t t0 = uint64_t(tok_off[s]), t1 = uint64_t(tok_off[s + 1]);
    uint32_t sum = 0;
    for (uint64_t k = t0 + lane; k < t1; k += 32) sum += rev_entry(V, load_id(tokens, k, enc), enc) >> 24;
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if (lane == 0) raw_len[s] = sum + (prefix_off ? uint32_t(prefix_off[s + 1] - prefix_off[s]) : 0u);
  }
}

// One warp per string, 32 ids per round: every lane looks its id up (one word), a warp prefix sum
// places the keys, each lane drops its key -- read 8 bytes at a time -- into a shared-memory window,
// and the warp writes the window out with consecutive lanes on consecutive bytes.
__global__ void __launch_bounds__(256) decode_copy_kernel(VocabDev V, const uint8_t* __restrict__ tokens,
                                                          const int64_t* __restrict__ tok_off, uint32_t n, int enc,
                                                          const uint64_t* __restrict__ raw_off,
                                                          uint8_t* __restrict__ out,
                                                          const uint64_t* __restrict__ prefix_off,
                                                          const uint8_t* __restrict__ prefix) {
  __shared__ uint8_t s_win[8][32 * 40];
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint8_t* win = s_win[threadIdx.x >> 5];
  for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const uint64_t t0 = uint64_t(tok_off[s]), t1 = uint64_t(tok_off[s + 1]);
    uint64_t at = raw_off[s];
    if (prefix_off) {  // streaming decoder: the stream's remainder comes first (go:717)
      const uint64_t p0 = prefix_off[s], pl = prefix_off[s + 1] - p0;
      for (uint64_t k = lane; k < pl; k += 32) out[at + k] = prefix[p0 + k];
      at += pl;
    }
    for (uint64_t kb = t0; kb < t1; kb += 32) {
      const uint64_t k = kb + lane;
      const uint32_t e = (k < t1) ? rev_entry(V, load_id(tokens, k, enc), enc) : 0u;
      const uint32_t l = e >> 24;
      uint32_t incl = l;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (l) {
        const unsigned long long* src = V.rev_blob8 + (e & 0xFFFFFFu);
        uint8_t* dst = win + (incl - l);
        for (uint32_t b = 0; b < l; b += 8) {
          unsigned long long w = __ldg(src + (b >> 3));
          const uint32_t m = l - b < 8u ? l - b : 8u;
          for (uint32_t j = 0; j < m; j++) {
            dst[b + j] = uint8_t(w);
            w >>= 8;
          }
        }
      }
      __syncwarp();
      for (uint32_t b = lane; b < total; b += 32) out[at + b] = win[b];
      __syncwarp();
      at += total;
    }
  }
}

// capcode level-2 decoder, one thread per string (javascript/tokenmonster.js:1007-1065; the
// stateless form used by Vocab.Decode, go/tokenmonster.go:788-796).  ASCII text never grows, so
// the result is written at the same offsets of `dst`.  With `wide` (vocabularies whose text may
// be non-ASCII) code points are classified and upper-cased through the Unicode tables (simple
// mappings, as Go's unicode.ToUpper); inside a W-word a 2-byte letter may become a 3-byte one,
// so string s is written at raw_off[s] + (raw_off[s]+1)/2, a region of 1.5x its length.
__global__ void __launch_bounds__(256) capcode_decode_kernel(const uint8_t* __restrict__ src,
                                                             const uint64_t* __restrict__ raw_off, uint32_t n,
                                                             uint8_t* __restrict__ dst,
                                                             uint32_t* __restrict__ out_len,
                                                             const uint32_t* __restrict__ len_in,
                                                             uint8_t* __restrict__ state, const UnicodeDev U,
                                                             int wide, int only_flagged) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (only_flagged && out_len[s] != 0xFFFFFFFFu) continue;   // the ASCII kernel has done this one
    const uint8_t* in = src + raw_off[s];
    uint8_t* out = dst + raw_off[s] + (wide ? ((raw_off[s] + 1) >> 1) : 0ull);
    const uint32_t len = len_in ? len_in[s] : uint32_t(raw_off[s + 1] - raw_off[s]);
    // streaming decoder: the four flags of capcode.Decoder persist across calls (js:1009-1012)
    const uint32_t st0 = state ? state[s] : 0u;
    bool in_word = st0 & 1u, in_char = st0 & 2u, del = st0 & 4u, ignore = st0 & 8u;
    uint32_t o = 0;
    for (uint32_t i = 0; i < len; i++) {
      uint32_t r = in[i];
      if (r == 'C') { in_char = true; in_word = false; continue; }
      if (r == 'W') { in_word = true; in_char = false; ignore = true; continue; }
      if (r == 'D') { del = true; continue; }
      if (r == ' ') {
        if (del) del = false;
        else { out[o++] = ' '; if (!ignore) in_word = false; }
      } else if (r >= 0x80) {
        // one code point (a malformed sequence is passed through byte by byte)
        uint32_t cp = 0;
        int l = wide ? utf8_decode(in + i, int64_t(len - i), &cp) : 0;
        if (l == 0) {
          l = 1;
          while (i + l < len && (in[i + l] & 0xC0) == 0x80) l++;
          cp = 0xFFFFFFFFu;
        }
        const uint32_t prop = (cp != 0xFFFFFFFFu) ? uni_prop(U, cp) : 0u;
        const uint32_t cl = (cp != 0xFFFFFFFFu) ? uni_class(cp, prop) : uint32_t(UC_O);
        bool raise = false;
        if (del) {
          del = false;                                   // js:1042-1043: the code point is swallowed
          i += l - 1;
          ignore = false;
          continue;
        } else if (in_char) {
          in_char = false;
          raise = true;                                  // js:1045-1046
        } else if (in_word) {
          if (cl == UC_LL || cl == UC_LU) raise = true;  // js:1048-1049
          else if (!(cl == UC_ND || cl == UC_APO || cl == UC_MN || cl == UC_MO)) in_word = false;
        }
        const uint32_t up = (raise && cp != 0xFFFFFFFFu) ? uni_upper(U, cp, prop) : cp;
        if (up != cp) {
          o += utf8_put(out + o, up);
        } else {
          for (int k = 0; k < l; k++) out[o++] = in[i + k];
        }
        i += l - 1;
      } else if (del) {
        del = false;
      } else if (in_char) {
        in_char = false;
        out[o++] = uint8_t((r - 'a' < 26u) ? r - 32 : r);
      } else if (in_word) {
        if (r - 'a' < 26u) out[o++] = uint8_t(r - 32);
        else if (r - 'A' < 26u) out[o++] = uint8_t(r);
        else {
          out[o++] = uint8_t(r);
          if (!((r - '0' < 10u) || r == '\'')) in_word = false;
        }
      } else {
        out[o++] = uint8_t(r);
      }
      ignore = false;
    }
    out_len[s] = o;
    if (state) state[s] = uint8_t((in_word ? 1u : 0u) | (in_char ? 2u : 0u) | (del ? 4u : 0u) | (ignore ? 8u : 0u));
  }
}

// Streaming decoder: how many bytes at the end of each stream's text are an incomplete UTF-8 /
// UTF-16 sequence and must wait for the next call (go/tokenmonster.go:149-219, including the
// quirk that a truncated UTF-8 sequence reports the MISSING byte count).
__global__ void __launch_bounds__(256) stream_tail_kernel(const uint8_t* __restrict__ raw,
                                                          const uint64_t* __restrict__ raw_off, uint32_t n,
                                                          int charset, uint32_t* __restrict__ keep_len,
                                                          uint32_t* __restrict__ rem_len) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const uint8_t* b = raw + raw_off[s];
    const long long len = (long long)(raw_off[s + 1] - raw_off[s]);
    long long inc = 0;
    if (len > 0 && charset == 1) {
      if (b[len - 1] & 0x80) {
        long long st = len - 1;
        while (st >= 0 && (b[st] & 0xC0) == 0x80) st--;
        if (st < 0) {
          inc = len;
        } else {
          const uint32_t f = b[st];
          int need;
          if ((f & 0x80) == 0) need = 1;
          else if ((f & 0xE0) == 0xC0) need = 2;
          else if ((f & 0xF0) == 0xE0) need = 3;
          else if ((f & 0xF8) == 0xF0) need = 4;
          else need = 0;
          const long long have = len - st;
          if (need == 0) inc = have;
          else if (have < need) inc = need - have;
          else if (need == 1 && (f & 0xC0) != 0) inc = len;
        }
      }
    } else if (len > 0 && charset == 2) {
      if (len & 1) {
        inc = 1;
        if (len >= 3) {
          const uint32_t u = b[len - 3] | (uint32_t(b[len - 2]) << 8);
          if (u >= 0xD800 && u <= 0xDBFF) inc = 3;
        }
      } else {
        const uint32_t last = b[len - 2] | (uint32_t(b[len - 1]) << 8);
        const uint32_t first = b[0] | (uint32_t(b[1]) << 8);
        if (last >= 0xD800 && last <= 0xDBFF) inc = 2;
        else if (first >= 0xDC00 && first <= 0xDFFF) inc = 2;
      }
    }
    if (inc > len) inc = len;
    keep_len[s] = uint32_t(len - inc);
    rem_len[s] = uint32_t(inc);
  }
}

// copy each stream's incomplete tail into the new remainder buffer
__global__ void __launch_bounds__(256) copy_tail_kernel(const uint8_t* __restrict__ raw,
                                                        const uint64_t* __restrict__ raw_off,
                                                        const uint32_t* __restrict__ keep_len,
                                                        const uint64_t* __restrict__ rem_off, uint32_t n,
                                                        uint8_t* __restrict__ rem) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const uint8_t* a = raw + raw_off[s] + keep_len[s];
    const uint64_t l = rem_off[s + 1] - rem_off[s];
    for (uint64_t k = 0; k < l; k++) rem[rem_off[s] + k] = a[k];
  }
}

// dense copy of variable-length segments: src segments at src_off[s] (length len[s]) -> dst_off[s].
// One warp per segment, one aligned 32-bit store per lane and step: the source word that lands in
// it straddles two aligned source words, joined by a funnel shift (source and destination are
// misaligned against each other by the bytes the earlier segments lost).  Only bytes of the
// segment itself are ever read.
__global__ void __launch_bounds__(256) compact_bytes_kernel(const uint8_t* __restrict__ src,
                                                            const uint64_t* __restrict__ src_off,
                                                            const uint32_t* __restrict__ len,
                                                            const uint64_t* __restrict__ dst_off, uint32_t n,
                                                            uint8_t* __restrict__ dst, int wide) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const uint8_t* a = src + src_off[s] + (wide ? ((src_off[s] + 1) >> 1) : 0ull);
    uint8_t* b = dst + dst_off[s];
    const uint32_t l = len[s];
    uint32_t head = uint32_t((4u - (reinterpret_cast<uintptr_t>(b) & 3u)) & 3u);   // bytes until b is word aligned
    head = head < l ? head : l;
    const uint8_t* sa = a + head;
    const uint32_t sh = uint32_t(reinterpret_cast<uintptr_t>(sa) & 3u);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa - sh);   // aligned word holding sa[0]
    uint32_t* dw = reinterpret_cast<uint32_t*>(b + head);
    // destination word i needs source bytes [4i, 4i + 4) of sa = aligned words i and (sh ? i + 1 : none):
    // whole words from the first whose aligned source word starts inside the segment to the last
    // whose second source word ends inside it; the (<= 7 + 7) bytes around them one by one
    const uint32_t rest = l - head;
    const uint32_t i0 = (sh > head) ? 1u : 0u;
    const uint32_t nw = sh ? (rest + sh >= 8u ? (rest + sh - 4u) >> 2 : 0u) : rest >> 2;
    uint32_t first_end = head + 4u * i0;
    first_end = first_end < l ? first_end : l;
    if (lane < first_end) b[lane] = a[lane];
    for (uint32_t i = i0 + lane; i < nw; i += 32) {
      const uint32_t w0 = __ldg(sw + i);
      const uint32_t w1 = sh ? __ldg(sw + i + 1) : 0u;
      dw[i] = __funnelshift_r(w0, w1, 8u * sh);
    }
    uint32_t last = head + 4u * nw;
    last = last > first_end ? last : first_end;
    for (uint32_t k = last + lane; k < l; k += 32) b[k] = a[k];
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// exclusive sum of u32 counts into u64 offsets (the accumulator type follows the u64 init value)
size_t scan_temp_bytes(uint32_t n_items) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveScan(nullptr, bytes, (const uint32_t*)nullptr, (uint64_t*)nullptr, ::cuda::std::plus<>{},
                                 uint64_t(0), int(n_items));
  return bytes;
}

cudaError_t exclusive_scan_u32_to_u64(void* temp, size_t temp_bytes, const uint32_t* in, uint64_t* out,
                                      uint32_t n_items, cudaStream_t st) {
  return cub::DeviceScan::ExclusiveScan(temp, temp_bytes, in, out, ::cuda::std::plus<>{}, uint64_t(0), int(n_items), st);
}

// one warp per string: any byte >= 0x80?
__global__ void __launch_bounds__(256) flag_non_ascii_kernel(const uint8_t* __restrict__ data,
                                                             const int64_t* __restrict__ off,
                                                             const int64_t* __restrict__ ends, uint32_t n,
                                                             uint8_t* __restrict__ skip, uint8_t* __restrict__ status,
                                                             uint32_t* __restrict__ counts,
                                                             uint32_t* __restrict__ missing) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const int64_t b0 = off[s], b1 = ends ? ends[s] : off[s + 1];
    bool hi = false;
    for (int64_t p = b0 + lane; p < b1; p += 32) hi |= (data[p] & 0x80u) != 0;
    hi = __any_sync(0xFFFFFFFFu, hi);
    if (lane == 0) {
      skip[s] = hi ? 1 : 0;
      if (hi) { status[s] = 13; counts[s] = 0; missing[s] = 0; }
    }
  }
}

cudaError_t launch_flag_non_ascii(const uint8_t* data, const int64_t* off, const int64_t* ends, uint32_t n,
                                  uint8_t* skip, uint8_t* status, uint32_t* counts, uint32_t* missing, int sms,
                                  cudaStream_t st) {
  flag_non_ascii_kernel<<<sms * 8, 256, 0, st>>>(data, off, ends, n, skip, status, counts, missing);
  return cudaGetLastError();
}

// persistent grid for a one-lane-per-string kernel: every SM gets its resident share of CTAs
template <class K>
static int pipe_grid(K kernel, uint32_t n, int sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPipeBlock, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const long long want = (static_cast<long long>(n) + kPipeBlock - 1) / kPipeBlock;
  const long long full = static_cast<long long>(sms) * per_sm;
  return int(want < full ? (want < 1 ? 1 : want) : full);
}

cudaError_t launch_prep(const TokenizeArgs& a, bool cap, bool slot_full, int sms, cudaStream_t st) {
  long long want = (static_cast<long long>(a.n) + 255) / 256;
  int grid = int(want < sms * 8ll ? (want < 1 ? 1 : want) : sms * 8ll);
  prep_kernel<<<grid, 256, 0, st>>>(a, cap ? 1 : 0, slot_full ? 1 : 0);
  return cudaGetLastError();
}

size_t sort_temp_bytes(uint32_t n_items) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                            (const uint32_t*)nullptr, (uint32_t*)nullptr, int(n_items), 0, 16);
  return bytes;
}

// items by descending length key (16 bits): long strings first, equal lengths next to each other
cudaError_t sort_items_by_length(void* temp, size_t temp_bytes, const uint32_t* key_in, uint32_t* key_out,
                                 const uint32_t* val_in, uint32_t* val_out, uint32_t n_items, cudaStream_t st) {
  return cub::DeviceRadixSort::SortPairsDescending(temp, temp_bytes, key_in, key_out, val_in, val_out, int(n_items), 0, 16,
                                                   st);
}

cudaError_t launch_table(const TokenizeArgs& a, bool cap, bool text, int sms, cudaStream_t st) {
  const uint32_t groups = (a.n + 31) / 32 * 32;
#define TM_TABLE(C_, T_) table_kernel<C_, T_><<<pipe_grid(table_kernel<C_, T_>, groups, sms), kPipeBlock, 0, st>>>(a)
  if (cap) { if (text) TM_TABLE(true, true); else TM_TABLE(true, false); }
  else { if (text) TM_TABLE(false, true); else TM_TABLE(false, false); }
#undef TM_TABLE
  return cudaGetLastError();
}

cudaError_t launch_select_overflowed(const uint8_t* status, const int64_t* off, const int64_t* ends, uint32_t n,
                                     uint32_t* list, int64_t* o_begin, int64_t* o_end, unsigned long long* ctr, int sms,
                                     cudaStream_t st) {
  select_overflowed_kernel<<<sms * 4, 256, 0, st>>>(status, off, ends, n, list, o_begin, o_end, ctr);
  return cudaGetLastError();
}

cudaError_t launch_chain(const TokenizeArgs& a, bool store, int sms, cudaStream_t st) {
  const uint32_t groups = (a.n + 31) / 32 * 32;
  if (store) chain_kernel<true><<<pipe_grid(chain_kernel<true>, groups, sms), kPipeBlock, 0, st>>>(a);
  else chain_kernel<false><<<pipe_grid(chain_kernel<false>, groups, sms), kPipeBlock, 0, st>>>(a);
  return cudaGetLastError();
}

uint32_t chunk_log_capacity() { return kChunkLogCap; }
uint32_t chunk_overlap() { return kChunkOverlap; }

cudaError_t launch_chunk_count(const int64_t* off, const int64_t* ends, uint32_t n, uint32_t* n_chunks, uint8_t* skip,
                               int sms, cudaStream_t st) {
  chunk_count_kernel<<<sms * 2, 256, 0, st>>>(off, ends, n, n_chunks, skip);
  return cudaGetLastError();
}

cudaError_t launch_chunk_fill(const uint8_t* data, const int64_t* off, const int64_t* ends, uint32_t n,
                              const uint32_t* n_chunks, const uint64_t* chunk_base, uint32_t total, ChunkItem* items,
                              uint8_t* doc_fail, int sms, cudaStream_t st) {
  chunk_fill_kernel<<<sms * 2, 256, 0, st>>>(data, off, ends, n, n_chunks, chunk_base, total, items, doc_fail);
  return cudaGetLastError();
}

// chunk launch: always the id-storing kernel (the stitcher needs the step log even for Count)
cudaError_t launch_tokenize_chunks(const TokenizeArgs& a, int sms, cudaStream_t st) {
  constexpr int B = 256;
  constexpr int smem = ThreadCfg<B>::SMEM;
  auto k0 = tokenize_thread_kernel<B, 0, true, true>;
  auto k2 = tokenize_thread_kernel<B, 2, true, true>;
  // function attributes are per device: set them on every launch (a process may hold several GPUs)
  int per_sm0 = 1, per_sm2 = 1;
  {
    cudaError_t e = cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm0, k0, B, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k2, B, smem);
    if (e != cudaSuccess) return e;
  }
  const bool cap2 = a.v.capcode == 2;
  const long long want = (static_cast<long long>(a.n) + B - 1) / B;
  const long long full = static_cast<long long>(sms) * (cap2 ? per_sm2 : per_sm0);
  const int grid = int(want < full ? (want < 1 ? 1 : want) : full);
  if (cap2) k2<<<grid, B, smem, st>>>(a);
  else k0<<<grid, B, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_chunk_stitch(const StitchArgs& a, uint32_t total_chunks, const ChunkItem* items, int sms,
                                cudaStream_t st) {
  chunk_meet_kernel<<<sms * 2, 128, 0, st>>>(a, total_chunks, items);
  chunk_segment_kernel<<<sms * 4, 256, 0, st>>>(a, total_chunks, items);
  chunk_finish_kernel<<<sms, 128, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_chunk_gather(const ChunkItem* items, uint32_t total, const uint32_t* c_ids, uint32_t c_cap,
                                const uint32_t* seg_begin, const uint32_t* seg_end, const uint32_t* seg_dst,
                                const int64_t* off, uint32_t slab_pad, uint32_t* slab, int sms, cudaStream_t st) {
  chunk_gather_kernel<<<sms * 8, 256, 0, st>>>(items, total, c_ids, c_cap, seg_begin, seg_end, seg_dst, off, slab_pad,
                                               slab);
  return cudaGetLastError();
}

cudaError_t launch_chunk_failed(const uint8_t* doc_fail, const uint32_t* n_chunks, const int64_t* off,
                                const int64_t* ends, uint32_t n, uint32_t* list, int64_t* f_begin, int64_t* f_end,
                                unsigned long long* ctr, int sms, cudaStream_t st) {
  chunk_failed_kernel<<<sms * 2, 256, 0, st>>>(doc_fail, n_chunks, off, ends, n, list, f_begin, f_end, ctr);
  return cudaGetLastError();
}

cudaError_t launch_pack(const uint32_t* slab, const int64_t* off, uint32_t slab_pad, const uint64_t* slab_off,
                        const uint32_t* counts, const uint64_t* tok_off, uint8_t* out, uint32_t n, int enc,
                        int sms, cudaStream_t st) {
  int grid = sms * 8;
  if (enc == 2) pack_tokens_kernel<2><<<grid, 256, 0, st>>>(slab, off, slab_pad, slab_off, counts, tok_off, out, n);
  else if (enc == 3) pack_tokens_kernel<3><<<grid, 256, 0, st>>>(slab, off, slab_pad, slab_off, counts, tok_off, out, n);
  else pack_tokens_kernel<4><<<grid, 256, 0, st>>>(slab, off, slab_pad, slab_off, counts, tok_off, out, n);
  return cudaGetLastError();
}

cudaError_t launch_select_refused(const uint8_t* status, const int64_t* off, const int64_t* ends, uint32_t n,
                                  uint32_t* list, int64_t* n_begin, unsigned long long* ctr, int sms,
                                  cudaStream_t st) {
  select_refused_kernel<<<sms * 4, 256, 0, st>>>(status, off, ends, n, list, n_begin, ctr);
  return cudaGetLastError();
}

cudaError_t launch_unicode_normalize(const NormalizeArgs& a, int sms, cudaStream_t st) {
  unicode_normalize_kernel<<<sms * 8, 128, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_wire_frame(const uint8_t* src, const uint64_t* src_off, uint32_t unit, uint32_t n, uint8_t* dst,
                              int sms, cudaStream_t st) {
  wire_frame_kernel<<<sms * 8, 256, 0, st>>>(src, src_off, unit, n, dst);
  return cudaGetLastError();
}

cudaError_t launch_decode_size(const VocabDev& v, const uint8_t* tokens, const int64_t* tok_off, uint32_t n, int enc,
                               uint32_t* raw_len, int sms, cudaStream_t st, const uint64_t* prefix_off) {
  decode_size_kernel<<<sms * 8, 256, 0, st>>>(v, tokens, tok_off, n, enc, raw_len, prefix_off);
  return cudaGetLastError();
}

cudaError_t launch_decode_copy(const VocabDev& v, const uint8_t* tokens, const int64_t* tok_off, uint32_t n, int enc,
                               const uint64_t* raw_off, uint8_t* out, int sms, cudaStream_t st,
                               const uint64_t* prefix_off, const uint8_t* prefix) {
  decode_copy_kernel<<<sms * 8, 256, 0, st>>>(v, tokens, tok_off, n, enc, raw_off, out, prefix_off, prefix);
  return cudaGetLastError();
}

cudaError_t launch_capcode_decode(const uint8_t* src, const uint64_t* raw_off, uint32_t n, uint8_t* dst,
                                  uint32_t* out_len, int sms, cudaStream_t st, const uint32_t* len_in,
                                  uint8_t* state, const UnicodeDev* uni, uint32_t* counter) {
  // ASCII strings by the warp-per-string scan kernel (strings are pulled from a work counter); the
  // rest (flagged out_len = 0xFFFFFFFF) per code point
  cudaError_t e = cudaMemsetAsync(counter, 0, 4, st);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, capcode_decode_scan_kernel, 256, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const long long want = (static_cast<long long>(n) + 7) / 8;   // eight warps = eight strings per CTA
  const long long full = static_cast<long long>(sms) * per_sm;
  const int grid = int(want < full ? (want < 1 ? 1 : want) : full);
  capcode_decode_scan_kernel<<<grid, 256, 0, st>>>(src, raw_off, n, dst, out_len, len_in, state, uni ? 1 : 0, counter);
  capcode_decode_kernel<<<sms * 8, 256, 0, st>>>(src, raw_off, n, dst, out_len, len_in, state,
                                                 uni ? *uni : UnicodeDev{}, uni ? 1 : 0, 1);
  return cudaGetLastError();
}

cudaError_t launch_stream_tail(const uint8_t* raw, const uint64_t* raw_off, uint32_t n, int charset,
                               uint32_t* keep_len, uint32_t* rem_len, int sms, cudaStream_t st) {
  stream_tail_kernel<<<sms * 4, 256, 0, st>>>(raw, raw_off, n, charset, keep_len, rem_len);
  return cudaGetLastError();
}

cudaError_t launch_copy_tail(const uint8_t* raw, const uint64_t* raw_off, const uint32_t* keep_len,
                             const uint64_t* rem_off, uint32_t n, uint8_t* rem, int sms, cudaStream_t st) {
  copy_tail_kernel<<<sms * 4, 256, 0, st>>>(raw, raw_off, keep_len, rem_off, n, rem);
  return cudaGetLastError();
}

cudaError_t launch_compact_bytes(const uint8_t* src, const uint64_t* src_off, const uint32_t* len,
                                 const uint64_t* dst_off, uint32_t n, uint8_t* dst, int sms, cudaStream_t st,
                                 int wide) {
  compact_bytes_kernel<<<sms * 8, 256, 0, st>>>(src, src_off, len, dst_off, n, dst, wide);
  return cudaGetLastError();
}

}  // namespace tmgpu

// pipeline_kernels.cuh -- K1/K3 of round 2: the position-table pipeline (pipeline_core.h has the
// per-string logic and the reference citations; vocab_tables.h the automaton).
//
//   encode_kernel<false>   capcode-2 vocabularies: normalized length of every string   } one lane
//   (scan)                 16-byte aligned text slots                                   } per
//   encode_kernel<true>    normalized text -> HBM                                       } string,
//   table_kernel           reversed-key automaton, right to left: one word / position   } lanes pull
//   chain_kernel<STORE>    the ungreedy loop over the position words -> ids / counts    } strings
//
// Why three passes over HBM instead of one fused kernel: the round-1 kernel kept everything in
// shared-memory rings and paid for it with ~12 dependent trie probes per emitted id under SIMT
// divergence (165 warp instructions per id, 13 of 32 lanes active, profiles/r01_*).  Here the
// automaton needs ONE table probe per byte whatever the token lengths, the chain stage has no
// inner loop at all, and every lane of a warp does the same thing in every iteration; the price is
// ~10 bytes of HBM traffic per input byte, on a GPU whose HBM was 99 % idle.
//
// All three kernels are persistent: a lane takes the next string from a global counter the moment
// its own string ends, so ragged batches balance themselves.  Per-lane loads are 16-byte vectors
// wherever the data is a stream (raw text, normalized text, position words are stored 16 bytes at
// a time): with one string per lane every load instruction touches 32 different lines, and the
// L1 pipeline charges per line, so the number of load INSTRUCTIONS per byte is what matters.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_tables.h"
#include "kernels.h"
#include "pipeline_core.h"

namespace tmgpu {

constexpr int kPipeImageBytes = 1536;   // dfa_d1[256] u32, class_code[256], 256 spare
constexpr int kPipeBlock = 256;

__device__ __forceinline__ uint32_t pipe_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void pipe_stage_tma(uint8_t* smem, const uint8_t* image, uint64_t* bar) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pipe_smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pipe_smem_u32(bar)),
                 "r"(kPipeImageBytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            pipe_smem_u32(smem)),
        "l"(image), "r"(kPipeImageBytes), "r"(pipe_smem_u32(bar))
        : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(pipe_smem_u32(bar))
        : "memory");
  }
}

// one 256-bit read-only load (SASS: LDG.E.256.CONSTANT): a whole 32-byte bucket / record per lane
__device__ __forceinline__ void pipe_load32(const uint4* p, uint4* a, uint4* b) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w)
               : "l"(p));
}

__device__ __forceinline__ uint32_t pipe_lower4(uint32_t w) {  // ASCII A-Z -> a-z in 4 packed bytes
  const uint32_t hi = w & 0x80808080u;
  const uint32_t x = w & 0x7F7F7F7Fu;
  const uint32_t ge_a = x + 0x3F3F3F3Fu;   // bit 7 set if byte >= 'A'
  const uint32_t gt_z = x + 0x25252525u;   // bit 7 set if byte >  'Z'
  const uint32_t is_up = (ge_a & ~gt_z & ~hi) & 0x80808080u;
  return w | (is_up >> 2);
}

// work item -> string.  false: nothing to do for this item (refused by the normalizer, or handled
// by another launch).
__device__ __forceinline__ bool pipe_fetch(const TokenizeArgs& A, uint32_t item, uint32_t* s, int64_t* o0, int* slen) {
  const int64_t b0 = A.off[item];
  const int64_t b1 = A.ends ? A.ends[item] : A.off[item + 1];
  if (A.list && b1 < b0) return false;       // the normalizer refused this string: status 13 stands
  if (A.skip && A.skip[item]) return false;  // a long document that another launch tokenizes in chunks
  *s = A.list ? A.list[item] : item;
  *o0 = b0;
  *slen = int(b1 - b0);
  return true;
}

// 16 raw bytes at aligned coordinate c0 of a string that ends at total_c; never reads past the end
__device__ __forceinline__ uint4 pipe_load16(const uint8_t* ga, int c0, int total_c) {
  if (c0 + 16 <= total_c) return __ldg(reinterpret_cast<const uint4*>(ga + c0));
  uint32_t t[4] = {0, 0, 0, 0};
  for (int k = 0; c0 + k < total_c; k++) t[k >> 2] |= uint32_t(ga[c0 + k]) << (8 * (k & 3));
  return make_uint4(t[0], t[1], t[2], t[3]);
}

__device__ __forceinline__ uint32_t pipe_word_of(const uint4& q, int wi) {
  return (wi == 0) ? q.x : (wi == 1) ? q.y : (wi == 2) ? q.z : q.w;
}

// ------------------------------------------------------------------------------------------
// stage 1: capcode level 2 (WRITE = false: lengths only)
// ------------------------------------------------------------------------------------------
template <bool WRITE>
__global__ void __launch_bounds__(kPipeBlock) encode_kernel(const TokenizeArgs A) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  __shared__ uint8_t s_cls[128];
  if (threadIdx.x < 128) s_cls[threadIdx.x] = uint8_t(pc_class(threadIdx.x));
  __syncthreads();
  const bool to_lower = (A.v.normflag & 2) != 0;
  uint32_t* counter = A.pipe_counters + (WRITE ? 1 : 0);

  bool running = false, no_more = false, bad = false;
  uint32_t item = 0, s = 0, ow = 0;
  int galign = 0, total_c = 0, rc = 0, qbase = -16, avail = 0, outpos = 0;
  uint64_t win = 0;
  uint4 q = make_uint4(0, 0, 0, 0);
  const uint8_t* ga = nullptr;
  uint32_t* dst = nullptr;
  CapState e;
  cap_reset(e);

  for (;;) {
    if (!running && !no_more) {
      item = atomicAdd(counter, 1u);
      if (item >= A.n) {
        no_more = true;
      } else {
        int64_t o0 = 0;
        int slen = 0;
        const bool ok = pipe_fetch(A, item, &s, &o0, &slen);
        if (!ok) {
          if (!WRITE) { A.npad[item] = 0; A.nlen[item] = 0; A.pflag[item] = 2; }
        } else if (WRITE && A.pflag[item] != 0) {
          // refused by the length pass
        } else {
          const uint8_t* g0 = A.data + o0;
          galign = int(reinterpret_cast<uintptr_t>(g0) & 15);
          ga = g0 - galign;
          total_c = galign + slen;
          rc = galign;
          qbase = -16;
          win = 0;
          avail = 0;
          outpos = 0;
          ow = 0;
          bad = false;
          cap_reset(e);
          if (WRITE) dst = reinterpret_cast<uint32_t*>(A.text + A.toff[item]);
          running = true;
        }
      }
    }
    if (__all_sync(FULL, !running)) break;
    if (!running) continue;
    if (rc >= total_c || bad) {   // the string is done
      if (WRITE) {
        if (outpos & 3) dst[outpos >> 2] = ow;
      } else {
        A.nlen[item] = bad ? 0u : uint32_t(outpos);
        A.npad[item] = bad ? 0u : (uint32_t(outpos) + 16u) & ~15u;
        A.pflag[item] = bad ? 1 : 0;
        if (bad) { A.counts[s] = 0; A.need[s] = 0; A.missing[s] = 0; A.status[s] = 13; }
      }
      running = false;
      continue;
    }
    // keep >= 5 raw bytes in the window: 4 for the fast path + 1 to peek at the next character
    while (avail < 5 && rc + avail < total_c) {
      const int nwc = rc + avail;            // first coordinate not yet in the window
      const int wc = nwc & ~3;
      const int cb = wc & ~15;
      if (cb != qbase) {
        q = pipe_load16(ga, cb, total_c);
        qbase = cb;
        uint32_t m0 = q.x, m1 = q.y, m2 = q.z, m3 = q.w;
        if (cb == 0 && galign) {             // bytes before the string start are foreign
          uint32_t t[4] = {m0, m1, m2, m3};
          for (int k = 0; k < galign; k++) t[k >> 2] &= ~(0xFFu << (8 * (k & 3)));
          m0 = t[0]; m1 = t[1]; m2 = t[2]; m3 = t[3];
        }
        if ((m0 | m1 | m2 | m3) & 0x80808080u) bad = true;
        if (to_lower) { q.x = pipe_lower4(q.x); q.y = pipe_lower4(q.y); q.z = pipe_lower4(q.z); q.w = pipe_lower4(q.w); }
      }
      const uint32_t wq = pipe_word_of(q, (wc >> 2) & 3);
      const int skip = nwc - wc;             // non-zero only for the first, unaligned word
      win |= uint64_t(wq >> (8 * skip)) << (8 * avail);
      avail += 4 - skip;
    }
    if (bad) continue;
    uint32_t bytes = 0, total = 0;
    bool fast = false;
    if (rc + 4 <= total_c) {
      const uint32_t c5 = (rc + 4 < total_c) ? (uint32_t(win >> 32) & 0xFFu) : 0u;
      fast = cap_fast4(e, uint32_t(win), c5);
    }
    if (fast) {
      bytes = uint32_t(win);
      total = 4;
      win >>= 32;
      avail -= 4;
      rc += 4;
    } else {
      const uint32_t c = uint32_t(win) & 0xFFu;
      const uint32_t cn = uint32_t(win >> 8) & 0xFFu;
      const int k = rc - galign, n = total_c - galign;
      const uint8_t* g0 = ga + galign;
      total = cap_step(e, s_cls, c, cn, k, n,
                       [&](int j) -> uint32_t {
                         uint32_t x = g0[j];
                         if (to_lower && x - 'A' < 26u) x |= 0x20u;
                         return x;
                       },
                       &bytes);
      win >>= 8;
      avail--;
      rc++;
    }
    // append `total` bytes through the partial word ow
    const uint32_t sh = (uint32_t(outpos) & 3u) * 8u;
    const uint64_t acc = uint64_t(ow) | (uint64_t(bytes) << sh);
    const int nf = outpos + int(total);
    if ((nf >> 2) != (outpos >> 2)) {
      if (WRITE) dst[outpos >> 2] = uint32_t(acc);
      ow = uint32_t(acc >> 32);
    } else {
      ow = uint32_t(acc);
    }
    outpos = nf;
  }
}

// ------------------------------------------------------------------------------------------
// stage 2: position words.  ENC: the text is what encode_kernel wrote (slot toff[item], length
// nlen[item]); otherwise the string itself (capcode 0, or text the device normalizer produced),
// checked for non-ASCII bytes / lower-cased on the fly when the vocabulary asks for it.
// ------------------------------------------------------------------------------------------
template <bool ENC>
__global__ void __launch_bounds__(kPipeBlock) table_kernel(const TokenizeArgs A) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  __shared__ __align__(128) uint8_t s_img[kPipeImageBytes];
  __shared__ __align__(8) uint64_t s_bar;
  pipe_stage_tma(s_img, A.v.pipe_image, &s_bar);
  __syncthreads();
  const uint32_t* s_d1 = reinterpret_cast<const uint32_t*>(s_img);
  const uint8_t* s_code = s_img + 1024;
  const DfaDev D = A.v.dfa;
  const uint32_t voff = uint32_t(A.v.off);
  const bool want_ascii = !ENC && !A.prenorm && (A.v.normflag != 0);
  const bool to_lower = !ENC && !A.prenorm && (A.v.normflag & 2) != 0;
  const uint32_t code0 = s_code[0];
  uint32_t* counter = A.pipe_counters + 2;
  const int64_t origin = (A.list && A.prenorm) ? 0 : (A.list ? A.raw_off[0] : A.off[0]);

  bool running = false, no_more = false, bad = false;
  uint32_t item = 0, s = 0, packed = 0, prev = 0, nbuf = 0;
  int p = -1, L = 0, ta = 0, qbase = -16, total_c = 0;
  const uint8_t* ga = nullptr;
  uint32_t* wdst = nullptr;
  uint4 q = make_uint4(0, 0, 0, 0), wbuf = make_uint4(0, 0, 0, 0);
  ClassTrail tr;
  trail_reset(tr, code0);

  for (;;) {
    if (!running && !no_more) {
      item = atomicAdd(counter, 1u);
      if (item >= A.n) {
        no_more = true;
      } else {
        int64_t o0 = 0;
        int slen = 0;
        bool ok = pipe_fetch(A, item, &s, &o0, &slen);
        if (ENC) {
          ok = ok && A.pflag[item] == 0;
        } else if (!ok) {
          A.pflag[item] = 2;
        }
        if (ok) {
          const uint8_t* g0;
          if (ENC) {
            g0 = A.text + A.toff[item];
            L = int(A.nlen[item]);
            wdst = A.words + A.toff[item];
          } else {
            g0 = A.data + o0;
            L = slen;
            // words of string s start at a multiple of 4: (offset rounded down) + 4 per string in
            // front for contiguous strings, (offset rounded up) inside the string's own slack for
            // the normalizer's scratch slots
            const uint64_t wb = (A.list && A.prenorm) ? uint64_t((o0 - origin + 3) & ~int64_t(3))
                                                      : uint64_t((o0 - origin) & ~int64_t(3)) + 4ull * (A.list ? s : item);
            wdst = A.words + wb;
            A.wbase[item] = wb;
          }
          ta = int(reinterpret_cast<uintptr_t>(g0) & 15);
          ga = g0 - ta;
          total_c = ta + L;
          p = L - 1;
          qbase = -16;
          bad = false;
          nbuf = 0;
          trail_reset(tr, code0);
          running = true;
        }
      }
    }
    if (__all_sync(FULL, !running)) break;
    if (!running) continue;
    if (p < 0 || bad) {
      if (!ENC) {
        A.pflag[item] = bad ? 1 : 0;
        if (bad) { A.counts[s] = 0; A.need[s] = 0; A.missing[s] = 0; A.status[s] = 13; }
      }
      running = false;
      continue;
    }
    const int c = ta + p;
    if ((c & ~15) != qbase) {
      qbase = c & ~15;
      q = pipe_load16(ga, qbase, total_c);
      if (want_ascii) {
        uint32_t m0 = q.x, m1 = q.y, m2 = q.z, m3 = q.w;
        if (qbase == 0 && ta) {
          uint32_t t[4] = {m0, m1, m2, m3};
          for (int k = 0; k < ta; k++) t[k >> 2] &= ~(0xFFu << (8 * (k & 3)));
          m0 = t[0]; m1 = t[1]; m2 = t[2]; m3 = t[3];
        }
        if (qbase + 16 > total_c) {   // (only the vector path can carry foreign bytes behind the end, and it never does)
        }
        if ((m0 | m1 | m2 | m3) & 0x80808080u) { bad = true; continue; }
      }
      if (to_lower) { q.x = pipe_lower4(q.x); q.y = pipe_lower4(q.y); q.z = pipe_lower4(q.z); q.w = pipe_lower4(q.w); }
    }
    const uint32_t b = (pipe_word_of(q, (c >> 2) & 3) >> (8 * (c & 3))) & 0xFFu;
    if (p == L - 1) {
      packed = s_d1[b];
    } else {
      packed = dfa_next(packed, prev, b, D.shift, D.mask,
                        [&](uint32_t i, uint4* k, uint4* v) { pipe_load32(D.buckets + 2 * i, k, v); },
                        [&](uint32_t i) { return __ldg(&D.d2[i]); });
    }
    const uint32_t code = s_code[b];
    const uint32_t w = position_word(packed, code, tr, voff);
    trail_push(tr, code);
    prev = b;
    wbuf.w = wbuf.z;
    wbuf.z = wbuf.y;
    wbuf.y = wbuf.x;
    wbuf.x = w;
    nbuf++;
    if ((p & 3) == 0) {
      uint32_t* d = wdst + p;
      if (nbuf == 4) {
        *reinterpret_cast<uint4*>(d) = wbuf;
      } else {
        d[0] = wbuf.x;
        if (nbuf > 1) d[1] = wbuf.y;
        if (nbuf > 2) d[2] = wbuf.z;
      }
      nbuf = 0;
    }
    p--;
  }
}

// ------------------------------------------------------------------------------------------
// stage 3: the ungreedy loop.  STORE = ids into the slab (K1), otherwise counts (K3, and the
// training worker's score histogram when A.scores is set).
// ------------------------------------------------------------------------------------------
template <bool STORE, bool ENC>
__global__ void __launch_bounds__(kPipeBlock) chain_kernel(const TokenizeArgs A) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  ChainTables T;
  T.rec = A.v.rec;
  T.hot = A.v.dfa.hot;
  T.delete_token = A.v.delete_token;
  T.unk = A.v.unk;
  T.off = uint32_t(A.v.off);
  T.can_retry = A.v.delete_token != 0xFFFFFFu;
  uint32_t* counter = A.pipe_counters + 3;
  const bool two_ids = STORE || A.scores != nullptr;

  bool running = false, no_more = false;
  uint32_t item = 0, s = 0, cap = 0;
  uint64_t slab_base = 0;
  const uint8_t* text = nullptr;
  ChainString S{nullptr, 0};
  ChainLane z;
  chain_reset(z);

  for (;;) {
    if (!running && !no_more) {
      item = atomicAdd(counter, 1u);
      if (item >= A.n) {
        no_more = true;
      } else if (A.pflag[item] == 0) {
        int64_t o0 = 0;
        int slen = 0;
        pipe_fetch(A, item, &s, &o0, &slen);
        if (ENC) {
          S.words = A.words + A.toff[item];
          S.len = int(A.nlen[item]);
          text = A.text + A.toff[item];
        } else {
          S.words = A.words + A.wbase[item];
          S.len = slen;
          text = A.data + o0;
        }
        if (A.list) {   // placement follows the string's own (raw) extent, like the first pass
          const int64_t r0 = A.raw_off[s];
          const int64_t r1 = A.raw_ends ? A.raw_ends[s] : A.raw_off[s + 1];
          slab_base = A.slab_off ? A.slab_off[s] : uint64_t(r0 - A.raw_off[0]) + uint64_t(A.slab_pad) * s;
          cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(r1 - r0) + A.slab_pad;
        } else {
          slab_base = A.slab_off ? A.slab_off[s] : uint64_t(o0 - A.off[0]) + uint64_t(A.slab_pad) * s;
          cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(slen) + A.slab_pad;
        }
        chain_reset(z);
        running = true;
      }
    }
    if (__all_sync(FULL, !running)) break;
    if (!running) continue;
    if (z.i >= S.len) {
      A.counts[s] = STORE ? min(z.cnt, cap) : z.cnt;
      A.need[s] = z.cnt;
      A.missing[s] = z.miss;
      A.status[s] = 0;
      if (STORE && z.cnt > cap) atomicExch(A.overflow, 1u);
      running = false;
      continue;
    }
    uint32_t stored = z.cnt;   // ids stored so far follow the lane's own count
    chain_step(
        T, S, z, two_ids, [&](uint32_t idx, uint4* r0, uint4* r1) { pipe_load32(T.rec + 2 * idx, r0, r1); },
        [&](uint32_t id, uint32_t, uint32_t adv) {
          if (!STORE && A.scores) atomicAdd(&A.scores[id], static_cast<unsigned long long>(adv));
          if (STORE && stored < cap) A.slab[slab_base + stored] = id;
          stored++;
        },
        [&](int pos) -> uint32_t {
          if (!STORE && A.scores) {   // trainvocab.go:1164-1171: counted, never scored
            uint32_t b = text[pos];
            if (!ENC && !A.prenorm && (A.v.normflag & 2) && b - 'A' < 26u) b |= 0x20u;
            atomicAdd(&A.miss_hist[b], 1ull);
            return 1u;
          }
          if (A.v.unk != 0xFFFFFFu) {
            if (STORE && stored < cap) A.slab[slab_base + stored] = A.v.unk;
            stored++;
            return 1u;
          }
          return 0u;
        });
  }
}

}  // namespace tmgpu

// tokenize_thread_kernel.cuh -- K1/K3, second mapping: ONE THREAD PER STRING.
//
// Why a second mapping: the lanes-over-branches kernel (tokenize_kernel.cuh) was measured on a
// B200 with ncu (profiles/r01_v1_group8_details.txt): 440 warp instructions per emitted id with
// 8.1 of 32 lanes active on average -- the per-string control flow keeps the groups of a warp
// from running in lock-step, so cooperative lanes buy nothing and the kernel is issue bound at
// 0.13 % of the HBM roofline.  Tokenising a string is a strict dependency chain (the branch
// chosen at step k fixes the position of step k+1, go/tokenmonster.go:1217-1262); the only
// parallelism that keeps all 32 lanes busy is ACROSS strings.  So here every lane runs the whole
// chain for its own string and the warp-level machinery is spent on what the memory system
// needs instead:
//   * normalized text lives in a shared-memory ring per string, interleaved word-by-word across
//     the lanes of a warp (word w of lane t at ring[w*32 + t]) so that 32 different strings read
//     their own bytes with zero bank conflicts, whatever their positions;
//   * capcode level 2 is encoded on the fly into that ring by a forward-looking restatement of
//     javascript/tokenmonster.js:900-1005: the marker of a capitalised run (W or C) and the
//     "D C ␠" respelling of a mixed-case run are decided when the run starts, by peeking at the
//     first lower-case / non-letter character after it, so nothing already emitted is rewritten;
//   * the hashed trie edges (8 bytes each) are read through a direct-mapped software cache in
//     shared memory shared by the whole CTA (hot edges hit there; misses fall through to the
//     L2-resident table and refill the slot), and the first-byte table + beginByte are staged
//     once per CTA by a TMA bulk copy;
//   * strings are pulled from a global work counter by each lane as it finishes, so ragged
//     batches keep the lanes of a warp occupied.
// Same reference semantics as tokenize_kernel.cuh: go/tokenmonster.go:242-253 + :1017-1279
// (ids) / :1281-1543 (count).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_tables.h"
#include "kernels.h"
#include "tokenize_kernel.cuh"  // stage_tables_tma, char_class, lower4, smem layout constants

namespace tmgpu {

constexpr int kRingBytes = 128;               // per string; >= 2*40+2 look-ahead + 16 refill + 15 align
constexpr int kRingWords = kRingBytes / 4;

template <int BLOCK>
struct ThreadCfg {
  static constexpr int WARPS = BLOCK / 32;
  static constexpr int RING_BYTES = WARPS * kRingWords * 32 * 4;   // interleaved rings
  static constexpr int CACHE_ENTRIES = (BLOCK >= 1024) ? 8192 : (BLOCK >= 512 ? 4096 : 2048);
  static constexpr int CACHE_BYTES = CACHE_ENTRIES * 8;
  static constexpr int CACHE_BASE = kSmemRingBase;
  static constexpr int RING_BASE = CACHE_BASE + CACHE_BYTES;
  static constexpr int SMEM = RING_BASE + RING_BYTES;
};

template <int BLOCK, int CAP, bool STORE>
__global__ void __launch_bounds__(BLOCK, 1) tokenize_thread_kernel(const TokenizeArgs A) {
  using Cfg = ThreadCfg<BLOCK>;
  extern __shared__ __align__(128) uint8_t smem[];
  // one 64-bit word per slot (key in the low half, value in the high half): read and written with
  // single 8-byte shared accesses so a concurrent refill can never be observed half-way
  volatile unsigned long long* s_cache = reinterpret_cast<volatile unsigned long long*>(smem + Cfg::CACHE_BASE);
  for (int k = threadIdx.x; k < Cfg::CACHE_ENTRIES; k += BLOCK) s_cache[k] = ~0ull;
  stage_tables_tma(smem, A.v.smem_image);   // contains a __syncthreads()
  __syncthreads();
  const uint32_t* s_root = reinterpret_cast<const uint32_t*>(smem);
  const uint8_t* s_begin = smem + 1024;
  const int lane = threadIdx.x & 31;
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem + Cfg::RING_BASE) + (threadIdx.x >> 5) * (kRingWords * 32) + lane;
  // word w of this string: ring[(w & 31) * 32]

  const VocabDev& V = A.v;
  const int M = V.max_len;
  const int LA = 2 * M + 2;
  const bool want_ascii = (V.normflag != 0) || (CAP != 0);
  const bool to_lower = (V.normflag & 2) != 0;
  constexpr uint32_t kCacheShift = 32 - (Cfg::CACHE_ENTRIES == 8192 ? 13 : (Cfg::CACHE_ENTRIES == 4096 ? 12 : 11));

  auto ring_byte = [&](int c) -> uint32_t {
    return (ring[((c >> 2) & (kRingWords - 1)) * 32] >> ((c & 3) * 8)) & 0xFFu;
  };

  // trie edge: CTA-shared direct-mapped cache in shared memory in front of the global table
  auto find_edge = [&](uint32_t key) -> uint32_t {
    const uint32_t slot = (key * 0x85EBCA6Bu) >> kCacheShift;
    const unsigned long long c = s_cache[slot];
    if (uint32_t(c) == key) return uint32_t(c >> 32);
    uint32_t h = (key * 0x9E3779B1u) >> V.edge_shift;
    uint32_t val;
    for (;;) {
      const uint2 e = __ldg(&V.edges[h]);
      if (e.x == key) { val = e.y; break; }
      if (e.x == kNoVal) { val = kNoVal; break; }
      h = (h + 1) & V.edge_mask;
    }
    s_cache[slot] = (static_cast<unsigned long long>(val) << 32) | key;
    return val;
  };

  // longest key prefixing normalized[c .. c+maxd) (ring coordinates); returns window bytes matched
  auto longest = [&](int c, int maxd, uint32_t start, uint32_t& index) -> int {
    int best_len = 0;
    uint32_t best = 0, node = 0, sig = 0;
    bool at_root = (start == kRootStart);
    if (!at_root) { node = start & 0xFFFFFFu; sig = start >> 24; }
    uint32_t w = ring[((c >> 2) & (kRingWords - 1)) * 32] >> ((c & 3) * 8);
    for (int d = 0; d < maxd; d++) {
      const uint32_t b = w & 0xFFu;
      const int cn = c + d + 1;
      w = (cn & 3) ? (w >> 8) : ring[((cn >> 2) & (kRingWords - 1)) * 32];
      uint32_t val;
      if (at_root) {
        at_root = false;
        val = s_root[b];
      } else {
        if (((sig >> sig_bit_of(b)) & 1u) == 0) break;
        val = find_edge((node << 8) | b);
      }
      if (val == kNoVal) break;
      node = val & 0xFFFFFFu;
      sig = val >> 24;
      if (node < V.n_info) { best = node; best_len = d + 1; }
    }
    index = best;
    return best_len;
  };

  // ---- per-lane string state (lives across iterations of the warp-synchronous loop below) ----
  bool running = false, no_more = false;
  uint32_t s = 0;
  int slen = 0, galign = 0, cshift = 0, total_c = 0;
  const uint8_t* ga = nullptr;
  uint64_t slab_base = 0;
  uint32_t cap = 0;
  int filled = 0;                                // ring coordinates [0, filled) are valid
  bool bad = false;
  // CAP 2 encoder: raw cursor + a 64-bit byte window refilled from 16-byte loads
  int rc = 0;                                    // aligned coordinate of the next raw byte
  uint4 q = make_uint4(0, 0, 0, 0);              // the 16-byte chunk at aligned coordinate qbase
  int qbase = -16;
  uint64_t win = 0;
  int avail = 0;                                 // bytes in win, starting at coordinate rc
  bool in_word = false, mixed = false;
  uint32_t rl1 = '.', rl2 = '.';
  uint32_t ow = 0;                               // partially assembled output word
  // consumer
  int i = 0, fd = 0, len = 0;
  uint32_t idx = 0;
  bool have = false;
  uint32_t cnt = 0, miss = 0;

  auto load16 = [&](int c0) -> uint4 {           // raw bytes [c0, c0+16) of the aligned stream
    if (c0 + 16 <= total_c) return __ldg(reinterpret_cast<const uint4*>(ga + c0));
    uint32_t t[4] = {0, 0, 0, 0};                // ragged tail: never read past the string
    for (int k = 0; c0 + k < total_c; k++) t[k >> 2] |= uint32_t(ga[c0 + k]) << (8 * (k & 3));
    return make_uint4(t[0], t[1], t[2], t[3]);
  };
  auto put = [&](uint32_t b) {                   // append one normalized byte (CAP 2)
    ow |= b << ((filled & 3) * 8);
    filled++;
    if ((filled & 3) == 0) {
      ring[(((filled - 1) >> 2) & (kRingWords - 1)) * 32] = ow;
      ow = 0;
    }
  };
  auto emit = [&](uint32_t id) {
    if (STORE && cnt < cap) A.slab[slab_base + cnt] = id;
    cnt++;
  };

  // One iteration = at most one token step for every lane.  The explicit __syncwarp() is what
  // keeps the 32 strings of a warp in lock-step: without it the lanes drift apart for good the
  // first time two strings end at different moments (measured: 2.0 of 32 lanes active).
  for (;;) {
    __syncwarp();
    if (!running && !no_more) {
      s = atomicAdd(A.work_counter, 1u);
      if (s >= A.n) {
        no_more = true;
      } else {
        const int64_t o0 = A.off[s];
        slen = int(A.off[s + 1] - o0);
        const uint8_t* g0 = A.data + o0;
        slab_base = A.slab_off ? A.slab_off[s] : uint64_t(o0 - A.off[0]) + uint64_t(A.slab_pad) * s;
        cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(slen) + A.slab_pad;
        galign = int(reinterpret_cast<uintptr_t>(g0) & 15);
        ga = g0 - galign;                        // 16-byte aligned
        cshift = (CAP == 0) ? galign : 0;        // ring coordinate of normalized byte 0
        total_c = galign + slen;                 // end of the raw data in aligned coordinates
        filled = 0; bad = false;
        rc = galign; qbase = -16; win = 0; avail = 0;
        in_word = false; mixed = false; rl1 = '.'; rl2 = '.'; ow = 0;
        i = 0; fd = 0; len = 0; idx = 0; have = false; cnt = 0; miss = 0;
        running = true;
      }
    }
    if (__all_sync(0xFFFFFFFFu, !running)) break;
    if (running) {
      // =========================== producer ==================================================
      if (CAP == 0) {
        while (filled < total_c && filled < i + cshift + LA) {
          uint4 w = load16(filled);
          if (want_ascii) {
            uint32_t m0 = w.x, m1 = w.y, m2 = w.z, m3 = w.w;
            if (filled < cshift) {                 // leading bytes of the first block are foreign
              uint32_t t[4] = {m0, m1, m2, m3};
              for (int k = 0; k < cshift; k++) t[k >> 2] &= ~(0xFFu << (8 * (k & 3)));
              m0 = t[0]; m1 = t[1]; m2 = t[2]; m3 = t[3];
            }
            if ((m0 | m1 | m2 | m3) & 0x80808080u) bad = true;
            if (to_lower) { w.x = lower4(w.x); w.y = lower4(w.y); w.z = lower4(w.z); w.w = lower4(w.w); }
          }
          const int w0 = filled >> 2;
          ring[((w0 + 0) & (kRingWords - 1)) * 32] = w.x;
          ring[((w0 + 1) & (kRingWords - 1)) * 32] = w.y;
          ring[((w0 + 2) & (kRingWords - 1)) * 32] = w.z;
          ring[((w0 + 3) & (kRingWords - 1)) * 32] = w.w;
          filled += 16;
        }
      } else {
        // capcode level 2 (javascript/tokenmonster.js:900-1005), one character at a time
        while (rc < total_c && filled < i + LA) {
          // keep >= 2 bytes in the window so the next character can be peeked
          while (avail < 2) {
            const int nwc = rc + avail;            // first coordinate not yet in the window
            if (nwc >= total_c) break;
            const int wc = nwc & ~3;
            const int cb = wc & ~15;
            if (cb != qbase) {
              q = load16(cb);
              qbase = cb;
              if (want_ascii) {
                uint32_t m0 = q.x, m1 = q.y, m2 = q.z, m3 = q.w;
                if (cb == 0 && galign) {           // bytes before the string start are foreign
                  uint32_t t[4] = {m0, m1, m2, m3};
                  for (int k = 0; k < galign; k++) t[k >> 2] &= ~(0xFFu << (8 * (k & 3)));
                  m0 = t[0]; m1 = t[1]; m2 = t[2]; m3 = t[3];
                }
                if ((m0 | m1 | m2 | m3) & 0x80808080u) bad = true;
                if (to_lower) { q.x = lower4(q.x); q.y = lower4(q.y); q.z = lower4(q.z); q.w = lower4(q.w); }
              }
            }
            const int wi = (wc >> 2) & 3;
            const uint32_t wq = (wi == 0) ? q.x : (wi == 1) ? q.y : (wi == 2) ? q.z : q.w;
            const int skip = nwc - wc;             // non-zero only for the first, unaligned word
            win |= uint64_t(wq >> (8 * skip)) << (8 * avail);
            avail += 4 - skip;
          }
          const uint32_t c = uint32_t(win) & 0xFFu;
          const bool has_next = (rc + 1 < total_c);
          const uint32_t cn = has_next ? (uint32_t(win >> 8) & 0xFFu) : 0u;
          win >>= 8;
          avail--;
          rc++;
          const uint32_t cl = char_class(c);
          const uint32_t cl_last = char_class(rl1);
          const bool last_letter = (cl_last == CL_U || cl_last == CL_L);
          const bool last_glue = last_letter || cl_last == CL_A;
          if (cl == CL_U) {
            if (!in_word) {
              // decide W (word stays capital) or C (mixed case) from the first L/O char after the run
              uint32_t k_cl = has_next ? char_class(cn) : CL_O;
              if (has_next && !(k_cl == CL_L || k_cl == CL_O)) {
                k_cl = CL_O;
                for (int j = rc + 1; j < total_c; j++) {   // rare: runs of 3+ capitals / digits
                  uint32_t cj = ga[j];
                  if (to_lower && cj - 'A' < 26u) cj += 32;
                  const uint32_t x = char_class(cj);
                  if (x == CL_L || x == CL_O) { k_cl = x; break; }
                }
              }
              mixed = (k_cl == CL_L);
              if (rl1 != ' ') put('D');
              put(mixed ? 'C' : 'W');
              put(' ');
              put(c | 0x20u);
              in_word = true;
            } else {
              if (mixed) { put('D'); put('C'); put(' '); }
              else if (!last_glue) { put('D'); put(' '); }
              put(c | 0x20u);
            }
          } else if (cl == CL_L) {
            const uint32_t cl_last2 = char_class(rl2);
            const bool joined = in_word ? last_glue
                                        : (rl1 == ' ' || last_letter ||
                                           ((cl_last2 == CL_U || cl_last2 == CL_L) && cl_last == CL_A));
            if (!joined) { put('D'); put(' '); }
            put(c);
            in_word = false;
          } else if (cl == CL_N) {
            const bool joined = in_word ? (cl_last == CL_N) : (rl1 == ' ' || cl_last == CL_N);
            if (!joined) { put('D'); put(' '); }
            put(c);
          } else {
            // a space right before the first letter of a capitalised word is replaced by the
            // marker (js:976-979): the capital emits "<marker>␠x", the space emits nothing
            const bool eaten = (c == ' ') && has_next && (cn - 'A' < 26u);
            if (!eaten) put(c);
            if (cl == CL_O) in_word = false;
          }
          rl2 = rl1;
          rl1 = c;
        }
        if (filled & 3) ring[((filled >> 2) & (kRingWords - 1)) * 32] = ow;   // publish the partial word
      }
      const bool exhausted = (CAP == 0) ? true : (rc >= total_c);
      const int len_data = (CAP == 0) ? slen : (exhausted ? filled : 0x3FFFFFFF);
      bool finished = bad || (i >= len_data);
      bool skip = finished;

      // =========================== consumer: one token step ==================================
      if (!skip && !have) {
        uint32_t x;
        const int l = longest(i + cshift, min(len_data - i, M), kRootStart, x);   // go:1049
        if (l == 0) {                                                             // go:1269-1276
          if (V.unk != 0xFFFFFFu) emit(V.unk);
          i++;
          miss++;
          fd = 0;
          skip = true;
        } else {
          idx = x;
          len = l;
          have = true;
        }
      }
      if (!skip) {
      const uint4 r0 = __ldg(&V.rec[2 * idx]);
      const uint4 r1 = __ldg(&V.rec[2 * idx + 1]);
      const uint32_t oflag = r0.x & 0xFFu;
      const int onw = int((r0.x >> 8) & 0xFFu);
      const int alen1 = int(r0.x >> 24), alen2 = int(r1.x & 0xFFu);
      const int i1 = i + len;
      const bool look = (i1 < len_data) && (((oflag & 32u) == 0) || s_begin[ring_byte(i1 + cshift)] != 12);   // go:1057

      int best = kNegScore, best_b = kNegScore;   // best plain / best retry score so far
      int w_plain = 0, w_retry = 0;               // branch that holds it (first wins on ties)
      uint32_t nx_plain = 0, nx_retry = 0;        // second token of that branch
      int nl_plain = 0, nl_retry = 0;
      int nbase_plain = 0, nbase_retry = 0;
      uint32_t x0 = 0;                            // primary plain look-ahead (next current token)
      int l0 = 0;
#pragma unroll
      for (int br = 0; br < 3; br++) {
        int base = len;
        uint32_t aflag = oflag;
        int anw = onw;
        bool exists = true;
        if (br == 1) { base = alen1 - fd; aflag = (r1.x >> 8) & 0xFFu; anw = int((r1.x >> 16) & 0xFFu); exists = alen1 != 0; }
        if (br == 2) { base = alen2 - fd; aflag = (r1.x >> 24) & 0xFFu; anw = int(r1.y & 0xFFu); exists = alen1 != 0 && alen2 != 0; }
        const int p = i + base;
        if (!(exists && (look || br == 0) && p <= len_data)) continue;
        const int room = len_data - p;
        uint32_t x;
        const int l = longest(p + cshift, min(room, M), kRootStart, x);
        if (br == 0) { x0 = x; l0 = l; }
        if (!look || l == 0) continue;
        const int nw = anw - fd;
        const uint32_t w0 = __ldg(&V.rec[2 * x]).x;
        const uint32_t bflag = w0 & 0xFFu;
        const int bnw = int((w0 >> 8) & 0xFFu);
        const int q1 = p + l;
        const int nb = s_begin[(q1 < len_data) ? ring_byte(q1 + cshift) : 0u];
        int sc = pair_score(base + l, aflag, nw, bflag, bnw, nb, false);
        if (br != 0) sc -= 100 * int(base + l < len) + 10000 * int(base + l == len);   // go:1132-1133
        if (sc > best) { best = sc; w_plain = br; nx_plain = x; nl_plain = l; nbase_plain = base; }
        // " "-prefixed retry when the second token starts inside a word (go:1088-1107)
        if (V.delete_token != 0xFFFFFFu && (bflag & 2u) && nb == 1 && bnw == 0 && V.space_start != kNoVal) {
          uint32_t xb;
          const int lb = longest(p + cshift, min(room, M - V.off), V.space_start, xb);
          if (lb + V.off > l + 1) {
            const uint32_t wb = __ldg(&V.rec[2 * xb]).x;
            const uint32_t bflag2 = wb & 0xFFu;
            const int bnw2 = int((wb >> 8) & 0xFFu);
            const int q2 = p + lb;
            const int nb2 = s_begin[(q2 < len_data) ? ring_byte(q2 + cshift) : 0u];
            int sb = pair_score(base + lb, aflag, nw, bflag2, bnw2, nb2, true);
            if (br != 0) sb -= 100 * int(base + lb < len) + 10000 * int(base + lb == len);
            if (sb > best_b) { best_b = sb; w_retry = br; nx_retry = xb; nl_retry = lb; nbase_retry = base; }
          }
        }
      }
      // tie order s1,s2,s3,s1b,s2b,s3b (go:1217-1262): strict > above keeps the first of each
      // kind; a plain score wins a tie against a retry score
      const uint32_t id0 = r0.y, id1 = r0.z, id2 = r0.w;
      if (best == kNegScore && best_b == kNegScore) {
        emit(id0);                                 // go:1265-1267
        i += len;
        fd = 0;
        have = (l0 > 0);
        idx = x0;
        len = l0;
      } else if (best >= best_b) {
        emit(w_plain == 0 ? id0 : (w_plain == 1 ? id1 : id2));
        i += nbase_plain;
        idx = nx_plain;
        len = nl_plain;
        fd = 0;
        have = true;
      } else {
        emit(w_retry == 0 ? id0 : (w_retry == 1 ? id1 : id2));
        if (STORE) emit(V.delete_token);           // Count() adds only 1 here (quirk Q2)
        i += nbase_retry;
        idx = nx_retry;
        len = nl_retry;
        fd = 1;
        have = true;
      }
      }  // !skip
      if (finished) {
        if (bad) { cnt = 0; miss = 0; }
        A.counts[s] = STORE ? min(cnt, cap) : cnt;
        A.need[s] = cnt;
        A.missing[s] = miss;
        A.status[s] = bad ? 13 : 0;
        if (STORE && cnt > cap) atomicExch(A.overflow, 1u);
        running = false;
      }
    }  // running
  }
}

}  // namespace tmgpu

// tokenize_kernel.cuh -- K1/K3, mapping 1 (G lanes per string): batched normalize -> capcode ->
// 6-branch ungreedy match.  This is the mapping the north star describes and the first one
// built; the default since is mapping 2 (tokenize_thread_kernel.cuh, one lane per string), which
// is 2x faster on large batches.  This one stays selectable (TMGPU_GROUP=4|8|16|32) and under
// test as an independent implementation of the same semantics; it does ASCII text only (the
// non-ASCII pass, chunking and scores always use mapping 2) and shares the TMA table staging,
// the ring constants and lower4() with it.
//
// Replaces, for a whole batch of strings at once,
//   /root/reference go/tokenmonster.go:242-253  normalize + applyCapcode
//   /root/reference go/tokenmonster.go:1017-1279 Vocab.tokenize (K1, ids) and :1281-1543
//   tokenizeCount (K3, STORE = false); the serialized widths are applied by pack_tokens_kernel.
//
// Mapping (not a port of the Go loop):
//   * a GROUP of G lanes owns one string; a warp holds 32/G strings; groups pull strings from
//     a global work counter, so a persistent grid of (SMs x occupancy) CTAs balances ragged
//     batches by itself.
//   * the normalized text never touches HBM: each group streams its string through a
//     shared-memory ring.  Producer = fill_*: coalesced 16-byte loads (capcode 0) or the
//     data-parallel capcode encoder (capcode 2: one input character per lane, emission offsets
//     by a group prefix sum).  Consumer = the match loop below.
//   * the three look-ahead positions of a step (after the longest token and after each of its two
//     alternatives: go:1068,1113,1165) are independent once (i, index, length, forwardDelete) is
//     known, so lanes 0..2 of the group walk the prefix trie for them at the same time; the
//     " "-prefixed retries (go:1088-1107) run as a second pass on the same lanes, only where the
//     plain result asks for one; the six scores live in registers and the winner is picked with
//     a group max + two ballots in the reference's tie order (go:1217-1262).
//   * the look-ahead token that wins becomes the next step's current token (the reference's
//     `goto checkpoint`), and so does the plain look-ahead after a step that takes no branch,
//     which is exactly the LongestSubstring the reference would recompute at the loop top.
//   * first-byte table + beginByte are staged once per CTA with one TMA bulk copy
//     (cp.async.bulk + mbarrier); trie edges and the 32-byte token records are read through
//     the read-only path and stay L1/L2 resident (a 32k vocabulary is ~3 MB of tables).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_tables.h"
#include "kernels.h"

namespace tmgpu {

constexpr int kBlockThreads = 256;
constexpr int kSmemTableBytes = 2304;  // root_child[256] u32 + begin_byte[256] + sp_child[256] u32
constexpr int kSmemRingBase = 2560;    // tables + mbarrier + class table, 512-aligned ring area

template <int G>
struct GroupCfg {
  static constexpr int FILL = 16 * G;                             // bytes per capcode-0 refill
  static constexpr int RB = (2 * FILL > 256) ? 2 * FILL : 256;    // ring bytes per group
  static constexpr int GROUPS = kBlockThreads / G;
  static constexpr int SMEM = kSmemRingBase + GROUPS * RB;
};

// ---- small PTX helpers -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void stage_tables_tma(uint8_t* smem, const uint8_t* image) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSmemTableBytes);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(kSmemTableBytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem)),
        "l"(image), "r"(kSmemTableBytes), "r"(smem_u32(bar))
        : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar))
        : "memory");
  }
}

__device__ __forceinline__ uint32_t lower4(uint32_t w) {  // ASCII A-Z -> a-z in 4 packed bytes
  uint32_t hi = w & 0x80808080u;
  uint32_t x = w & 0x7F7F7F7Fu;
  uint32_t ge_a = x + 0x3F3F3F3Fu;   // bit7 set if byte >= 'A'
  uint32_t gt_z = x + 0x25252525u;   // bit7 set if byte >  'Z'
  uint32_t is_up = (ge_a & ~gt_z & ~hi) & 0x80808080u;
  return w | (is_up >> 2);
}

// character classes of the capcode encoder (javascript/tokenmonster.js:880-898, ASCII)
enum : uint32_t { CL_O = 0, CL_U = 1, CL_L = 2, CL_N = 3, CL_A = 4 };
__device__ __forceinline__ uint32_t char_class(uint32_t c) {
  if (c - 'a' < 26u) return CL_L;
  if (c - 'A' < 26u) return CL_U;
  if (c - '0' < 10u) return CL_N;
  if (c == '\'') return CL_A;
  return CL_O;
}

template <int G, int CAP, bool STORE>
__global__ void __launch_bounds__(kBlockThreads) tokenize_kernel(const TokenizeArgs A) {
  using Cfg = GroupCfg<G>;
  constexpr int RB = Cfg::RB;
  constexpr int FILL = Cfg::FILL;
  extern __shared__ __align__(128) uint8_t smem[];
  stage_tables_tma(smem, A.v.smem_image);
  const uint32_t* s_root = reinterpret_cast<const uint32_t*>(smem);
  const uint8_t* s_begin = smem + 1024;

  const int lane = threadIdx.x & 31;
  const int gl = lane % G;                    // lane within the group
  const int gbase = lane - gl;                // first warp lane of the group
  const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << gbase);
  uint8_t* ring = smem + kSmemRingBase + (threadIdx.x / G) * RB;

  const VocabDev& V = A.v;
  const int M = V.max_len;
  const int LA = 2 * M + 2;                   // look-ahead the match loop may touch
  const bool want_ascii = (V.normflag != 0) || (CAP != 0);
  const bool to_lower = (V.normflag & 2) != 0;

  // per-lane branch role: lanes 0..2 of a group own the three look-ahead positions (after the
  // longest token, after alternative 1, after alternative 2); the same lane also does that
  // position's space-prefixed retry when it is needed
  const int br = gl;
  const bool is_branch_lane = (gl < 3);

  // ---- per-group string state (lives across iterations of the warp-synchronous loop) --------
  bool running = false, no_more = false;
  uint32_t s = 0;
  int slen = 0, cshift = 0, total_c = 0;
  const uint8_t* g0 = nullptr;
  const uint8_t* ga = nullptr;
  uint64_t slab_base = 0;
  uint32_t cap = 0;
  int filled = 0;                           // ring coordinates [0, filled) are valid
  int rpos = 0;                             // CAP 2: raw characters consumed
  bool in_word = false, run_mixed = false;  // CAP 2 encoder carry
  uint32_t rl1 = '.', rl2 = '.';            // CAP 2: previous two raw characters
  bool bad = false;                         // non-ASCII where ASCII is required
  int i = 0, fd = 0, len = 0;
  uint32_t idx = 0;
  bool have = false;
  uint32_t cnt = 0, miss = 0;

  auto ring_at = [&](int p) -> uint32_t { return ring[(p + cshift) & (RB - 1)]; };
  auto emit = [&](uint32_t id) {
    if (STORE && gl == 0 && cnt < cap) A.slab[slab_base + cnt] = id;
    cnt++;
  };

  // One iteration = at most one token step for every group of the warp.  The full-warp
  // __syncwarp() keeps the 32/G strings of a warp in lock-step; without it the groups drift
  // apart for good and every instruction issues for one group only (measured: 8.1 of 32 lanes
  // active at G = 8, profiles/r01_v1_group8_details.txt).
  for (;;) {
    __syncwarp();
    if (!running && !no_more) {
      if (gl == 0) s = atomicAdd(A.work_counter, 1u);
      s = __shfl_sync(gmask, s, 0, G);
      if (s >= A.n) {
        no_more = true;
      } else {
        const int64_t o0 = A.off[s];
        slen = int((A.ends ? A.ends[s] : A.off[s + 1]) - o0);
        if (A.watermark) {   // the batch is still streaming in: wait until this string has landed
          const unsigned long long need_bytes = static_cast<unsigned long long>(o0 + slen);
          while (*reinterpret_cast<const volatile unsigned long long*>(A.watermark) < need_bytes) __nanosleep(500);
        }
        g0 = A.data + o0;
        slab_base = A.slab_off ? A.slab_off[s] : uint64_t(o0 - A.off[0]) + uint64_t(A.slab_pad) * s;
        cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(slen) + A.slab_pad;
        cshift = (CAP == 0) ? int(reinterpret_cast<uintptr_t>(g0) & 15) : 0;
        ga = g0 - cshift;                       // 16-byte aligned
        filled = 0;
        total_c = cshift + slen;                // CAP 0: end coordinate of the data
        rpos = 0; in_word = false; run_mixed = false; rl1 = '.'; rl2 = '.'; bad = false;
        i = 0; fd = 0; len = 0; idx = 0; have = false; cnt = 0; miss = 0;
        running = true;
      }
    }
    if (__all_sync(0xFFFFFFFFu, !running)) break;
    if (running) {
      // =========================== producer: keep LA bytes ahead of i ================
      if (CAP == 0) {
        while (filled < total_c && filled < i + cshift + LA) {
          const int c0 = filled + 16 * gl;
          uint4 w = make_uint4(0, 0, 0, 0);
          if (c0 < total_c) {
            if (c0 + 16 <= total_c) {
              w = __ldg(reinterpret_cast<const uint4*>(ga + c0));
            } else {  // ragged tail: never read past the last byte of the string
              uint32_t t[4] = {0, 0, 0, 0};
              for (int k = 0; c0 + k < total_c; k++) t[k >> 2] |= uint32_t(ga[c0 + k]) << (8 * (k & 3));
              w = make_uint4(t[0], t[1], t[2], t[3]);
            }
            if (want_ascii) {
              // bytes before the string start (c0 < cshift) belong to the previous string
              uint32_t m0 = w.x, m1 = w.y, m2 = w.z, m3 = w.w;
              if (c0 < cshift) {
                int skip = cshift - c0;  // 1..15 leading bytes to ignore
                uint32_t t[4] = {m0, m1, m2, m3};
                for (int k = 0; k < skip; k++) t[k >> 2] &= ~(0xFFu << (8 * (k & 3)));
                m0 = t[0]; m1 = t[1]; m2 = t[2]; m3 = t[3];
              }
              if ((m0 | m1 | m2 | m3) & 0x80808080u) bad = true;
              if (to_lower) { w.x = lower4(w.x); w.y = lower4(w.y); w.z = lower4(w.z); w.w = lower4(w.w); }
            }
          }
          *reinterpret_cast<uint4*>(ring + (c0 & (RB - 1))) = w;
          filled += FILL;
        }
        if (want_ascii) bad = __any_sync(gmask, bad);
      } else {
        // ---- capcode level 2, data-parallel (javascript/tokenmonster.js:900-1005) ----
        // One raw character per lane.  inWord before a character = class of the last
        // U/L/O character before it; a capital run is "mixed" iff the first L/O after it
        // is a lower-case letter; each character then emits 0..4 bytes on its own and a
        // group prefix sum places them.
        while (rpos < slen && filled < i + LA) {
          const int k = rpos + gl;
          uint32_t c = (k < slen) ? uint32_t(g0[k]) : 0u;
          uint32_t cn = (k + 1 < slen) ? uint32_t(g0[k + 1]) : 0u;  // next raw character
          if ((c | cn) & 0x80u) bad = true;
          if (to_lower) { c = (c - 'A' < 26u) ? c + 32 : c; cn = (cn - 'A' < 26u) ? cn + 32 : cn; }
          const bool valid = (k < slen);
          const uint32_t cl = valid ? char_class(c) : CL_O;
          const unsigned shiftm = gbase;
          const uint32_t mU = (__ballot_sync(gmask, valid && cl == CL_U) & gmask) >> shiftm;
          const uint32_t mL = (__ballot_sync(gmask, valid && cl == CL_L) & gmask) >> shiftm;
          const uint32_t mO = (__ballot_sync(gmask, valid && cl == CL_O) & gmask) >> shiftm;
          const uint32_t mSig = mU | mL | mO;
          const uint32_t below = (gl == 0) ? 0u : (mSig & ((1u << gl) - 1u));
          const bool w_before = below ? (((mU >> (31 - __clz(below))) & 1u) != 0) : in_word;
          // terminator of the capital run this lane may belong to
          const uint32_t term = (mL | mO) & ~((2u << gl) - 1u);  // L/O strictly above this lane
          const bool has_term = term != 0;
          bool mixed = has_term ? (((mL >> (__ffs(term) - 1)) & 1u) != 0) : false;
          const bool open_u = valid && cl == CL_U && !has_term;  // run leaves the chunk
          const bool carry_ok = in_word && ((mL | mO) == 0);     // still the run we entered with
          const uint32_t need = __ballot_sync(gmask, open_u && !carry_ok) & gmask;
          bool open_mixed = run_mixed;
          if (need) {  // scan ahead for the first L/O character (group-uniform loop)
            open_mixed = false;
            for (int j = rpos + G; j < slen; j += G) {
              const int kk = j + gl;
              uint32_t cc = (kk < slen) ? uint32_t(g0[kk]) : 0u;
              if (to_lower && cc - 'A' < 26u) cc += 32;
              const uint32_t c2 = (kk < slen) ? char_class(cc) : CL_U;  // padding is neutral
              const uint32_t tL = (__ballot_sync(gmask, c2 == CL_L) & gmask) >> shiftm;
              const uint32_t tO = (__ballot_sync(gmask, c2 == CL_O) & gmask) >> shiftm;
              if (tL | tO) {
                open_mixed = ((tL >> (__ffs(tL | tO) - 1)) & 1u) != 0;
                break;
              }
            }
          }
          if (!has_term) mixed = open_mixed;

          uint32_t rlast = __shfl_up_sync(gmask, c, 1, G);
          uint32_t rlast2 = __shfl_up_sync(gmask, c, 2, G);
          if (gl == 0) { rlast = rl1; rlast2 = rl2; }
          if (gl == 1) rlast2 = rl1;
          const uint32_t cl_last = char_class(rlast), cl_last2 = char_class(rlast2);
          const bool last_letter = (cl_last == CL_U || cl_last == CL_L);
          const bool last_glue = last_letter || cl_last == CL_A;

          // emission of this character: up to 4 bytes, packed little endian in `ob`
          uint32_t ob = 0, ol = 0;
          auto put = [&](uint32_t b) { ob |= b << (8 * ol); ol++; };
          if (valid) {
            if (cl == CL_U) {
              if (!w_before) {  // first letter of a capitalised word: marker W or C
                if (rlast != ' ') put('D');
                put(mixed ? 'C' : 'W');
                put(' ');
                put(c | 0x20u);
              } else {
                if (mixed) { put('D'); put('C'); put(' '); }
                else if (!last_glue) { put('D'); put(' '); }
                put(c | 0x20u);
              }
            } else if (cl == CL_L) {
              const bool joined = w_before ? last_glue
                                           : (rlast == ' ' || last_letter ||
                                              ((cl_last2 == CL_U || cl_last2 == CL_L) && cl_last == CL_A));
              if (!joined) { put('D'); put(' '); }
              put(c);
            } else if (cl == CL_N) {
              const bool joined = w_before ? (cl_last == CL_N) : (rlast == ' ' || cl_last == CL_N);
              if (!joined) { put('D'); put(' '); }
              put(c);
            } else {
              // a space directly before the first letter of a capitalised word is replaced by
              // the marker (js:976-979): the capital emits "<marker>␠x", the space emits nothing
              const bool eaten = (c == ' ') && (k + 1 < slen) && (cn - 'A' < 26u);
              if (!eaten) put(c);
            }
          }
          // exclusive prefix sum of ol over the group
          uint32_t incl = ol;
#pragma unroll
          for (int d = 1; d < G; d <<= 1) {
            uint32_t t = __shfl_up_sync(gmask, incl, d, G);
            if (gl >= d) incl += t;
          }
          const uint32_t total = __shfl_sync(gmask, incl, G - 1, G);
          const uint32_t at = uint32_t(filled) + incl - ol;
          for (uint32_t q = 0; q < ol; q++) ring[(at + q) & (RB - 1)] = uint8_t(ob >> (8 * q));
          filled += int(total);
          // carries
          const int nvalid = min(G, slen - rpos);
          const uint32_t lastc = __shfl_sync(gmask, c, nvalid - 1, G);
          const uint32_t lastc2 = (nvalid >= 2) ? __shfl_sync(gmask, c, nvalid - 2, G) : rl1;
          rl2 = lastc2;
          rl1 = lastc;
          in_word = mSig ? (((mU >> (31 - __clz(mSig))) & 1u) != 0) : in_word;
          run_mixed = open_mixed;
          rpos += nvalid;
        }
        bad = __any_sync(gmask, bad);
      }
      __syncwarp(gmask);
      const int len_data = (CAP == 0) ? slen : ((rpos >= slen) ? filled : 0x3FFFFFFF);
      const bool finished = bad || (i >= len_data);
      bool skip = finished;

      // =========================== consumer: one token step ==========================
      if (!skip && !have) {
        // LongestSubstring(data[i : i + min(lenData - i, M)])  (go:1049)
        uint32_t x = 0;
        int l = 0;
        if (gl == 0) {
          const int maxd = min(len_data - i, M);
          l = trie_longest([&](int d) { return ring_at(i + d); }, [&](uint32_t b) { return s_root[b]; },
                           [&](uint32_t h) { return __ldg(&V.edges[h]); }, V.edge_mask, V.edge_shift, V.n_info,
                           maxd, kRootStart, &x);
        }
        l = __shfl_sync(gmask, l, 0, G);
        x = __shfl_sync(gmask, x, 0, G);
        if (l == 0) {  // no token covers this byte (go:1269-1276)
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

      // ---- "checkpoint": current token = (idx, len) at i (go:1051-1054) ----------------
      const uint4 r0 = __ldg(&V.rec[2 * idx]);
      const uint4 r1 = __ldg(&V.rec[2 * idx + 1]);
      const uint32_t oflag = r0.x & 0xFFu, onw = (r0.x >> 8) & 0xFFu;
      const int alen1 = int(r0.x >> 24), alen2 = int(r1.x & 0xFFu);
      const int i1 = i + len;
      // skip the look-ahead for a whole word followed by a space (go:1057)
      const bool look = (i1 < len_data) && (((oflag & 32u) == 0) || s_begin[ring_at(i1)] != 12);

      // this lane's candidate: first token a (flag, nWords), bytes it consumes, where b starts.
      // lanes 0,1,2 of the group = look-ahead after the longest token / alternative 1 / alternative 2
      int base = len;
      uint32_t aflag = oflag;
      int anw = int(onw);
      bool exists = is_branch_lane;
      if (br == 1) { base = alen1 - fd; aflag = (r1.x >> 8) & 0xFFu; anw = int((r1.x >> 16) & 0xFFu); exists = exists && alen1 != 0; }
      if (br == 2) { base = alen2 - fd; aflag = (r1.x >> 24) & 0xFFu; anw = int(r1.y & 0xFFu); exists = exists && alen1 != 0 && alen2 != 0; }
      const int nw = anw - fd;
      const int p = i + base;
      const int room = len_data - p;
      // lane 0's plain walk is always wanted: it is the next current token when no branch wins
      const bool walk = exists && (look || gl == 0) && (room >= 0);
      // ---- phase 1: plain look-ahead (go:1068, :1113, :1165) ------------------------------
      uint32_t x = 0;
      int l = 0;
      if (walk) {
        l = trie_longest([&](int d) { return ring_at(p + d); }, [&](uint32_t b) { return s_root[b]; },
                         [&](uint32_t h) { return __ldg(&V.edges[h]); }, V.edge_mask, V.edge_shift, V.n_info,
                         min(room, M), kRootStart, &x);
      }
      uint32_t bflag = 0;
      int bnw = 0, nb = 0;
      int score = kNegScore;
      if (look && l > 0) {
        const uint32_t w0 = __ldg(&V.rec[2 * x]).x;
        bflag = w0 & 0xFFu;
        bnw = int((w0 >> 8) & 0xFFu);
        const int q = p + l;  // q <= len_data; the byte AT len_data reads as 0 (cpp:1726)
        nb = s_begin[(q < len_data) ? ring_at(q) : 0u];
        score = pair_score(base + l, aflag, nw, bflag, bnw, nb, false);
        if (br != 0) score -= 100 * int(base + l < len) + 10000 * int(base + l == len);  // go:1132-1133
      }
      // ---- phase 2: the " "-prefixed retry, only where the second token starts inside a word
      //      (go:1088-1107); same lane, same position, trie walk started behind the virtual space
      const bool need = (score != kNegScore) && (V.delete_token != 0xFFFFFFu) && (V.space_start != kNoVal) &&
                        (bflag & 2u) && nb == 1 && bnw == 0;
      int score_b = kNegScore, lb = 0;
      uint32_t xb = 0;
      if (__any_sync(gmask, need)) {
        if (need) {
          lb = trie_longest([&](int d) { return ring_at(p + d); }, [&](uint32_t b) { return s_root[b]; },
                            [&](uint32_t h) { return __ldg(&V.edges[h]); }, V.edge_mask, V.edge_shift, V.n_info,
                            min(room, M - V.off), V.space_start, &xb);
          if (lb + V.off > l + 1) {
            const uint32_t w0 = __ldg(&V.rec[2 * xb]).x;
            const uint32_t bflag2 = w0 & 0xFFu;
            const int bnw2 = int((w0 >> 8) & 0xFFu);
            const int q = p + lb;
            const int nb2 = s_begin[(q < len_data) ? ring_at(q) : 0u];
            score_b = pair_score(base + lb, aflag, nw, bflag2, bnw2, nb2, true);
            if (br != 0) score_b -= 100 * int(base + lb < len) + 10000 * int(base + lb == len);
          }
        }
      }
      // group max, then the first holder in the tie order s1,s2,s3,s1b,s2b,s3b (go:1217-1262)
      int best = max(score, score_b);
#pragma unroll
      for (int d = 1; d < 4 && d < G; d <<= 1) best = max(best, __shfl_xor_sync(gmask, best, d, G));
      best = __shfl_sync(gmask, best, 0, G);  // lanes 0..3 hold the max; make it group-uniform
      const uint32_t plain_mask = (__ballot_sync(gmask, score == best && is_branch_lane) & gmask) >> gbase;
      const uint32_t retry_mask = (__ballot_sync(gmask, score_b == best && is_branch_lane) & gmask) >> gbase;
      const uint32_t id0 = r0.y, id1 = r0.z, id2 = r0.w;
      if (best == kNegScore) {
        emit(id0);                       // go:1265-1267
        i += len;
        fd = 0;
        const int nl = __shfl_sync(gmask, l, 0, G);  // lane 0's plain walk = next LongestSubstring
        const uint32_t nx = __shfl_sync(gmask, x, 0, G);
        have = (nl > 0);
        idx = nx;
        len = nl;
      } else {
        const bool retry = (plain_mask == 0);
        const int w = __ffs(retry ? retry_mask : plain_mask) - 1;
        emit(w == 0 ? id0 : (w == 1 ? id1 : id2));
        if (retry && STORE) emit(V.delete_token);   // Count() adds only 1 here (go:1281-1543, quirk Q2)
        i += __shfl_sync(gmask, base, w, G);
        idx = __shfl_sync(gmask, retry ? xb : x, w, G);
        len = __shfl_sync(gmask, retry ? lb : l, w, G);
        fd = retry ? 1 : 0;
        have = true;
      }
      }  // !skip
      if (finished) {
        if (gl == 0) {
          if (bad) { cnt = 0; miss = 0; }
          A.counts[s] = STORE ? min(cnt, cap) : cnt;
          A.need[s] = cnt;
          A.missing[s] = miss;
          A.status[s] = bad ? 13 : 0;
          if (STORE && cnt > cap) atomicExch(A.overflow, 1u);
        }
        running = false;
      }
    }  // running
  }
}

}  // namespace tmgpu

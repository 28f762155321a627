// tokenize_thread_kernel.cuh -- the round-1 trie-walk kernel, kept for ONE job: the speculative
// chunks of long documents (chunking.cuh; instantiated with CHUNK = true only).  Batches of
// ordinary strings run the position-table pipeline (pipeline_kernels.cuh), which is 2.8x faster
// per byte but walks a string with one lane from end to end; a batch of a few huge documents
// needs parallelism INSIDE a document, and this kernel's step log (position, forward-delete flag
// per emitted id) is what the stitcher joins chunks on.
//
// One lane per chunk, warp-synchronous phases per token step:
//     FILL   each lane tops up its normalized-text ring (capcode level 2 encoded on the fly by a
//            branch-free, forward-looking restatement of javascript/tokenmonster.js:900-1005)
//     WALK 1 one loop, one copy of the trie-step code: a lane pops its next look-ahead position
//            (after the longest token / alternative 1 / alternative 2: go:1068,1113,1165) the
//            moment its previous walk ends; the walks of all 32 lanes go into one pool that idle
//            lanes take from by ballot rank
//     SCORE  straight-line code for the three plain candidates (go:1075-1133)
//     WALK 2 same loop for the " "-prefixed retries that the plain results ask for (go:1088-1107)
//     PICK   first maximum in the order s1,s2,s3,s1b,s2b,s3b (go:1217-1262), emit, advance
//
// Memory system: normalized text lives in shared-memory rings interleaved word-by-word across the
// lanes of a warp; trie edges (8 bytes each, L2-resident hashed table at load factor <= 1/4) are
// read through the read-only path behind an 8-bit child-byte filter kept in the parent's value;
// first-byte table, beginByte and the character-class table are staged once per CTA by a TMA bulk
// copy; chunks are pulled from a global work counter by each lane.
// Reference semantics: go/tokenmonster.go:242-253 + :1017-1279 (ids) / :1281-1543 (count).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_tables.h"
#include "kernels.h"
#include "chunk_kernel_helpers.cuh"  // stage_tables_tma, lower4, smem layout constants

namespace tmgpu {

constexpr int kRingBytes = 128;               // per string; >= 2*40+2 look-ahead + 16 refill + 15 align
constexpr int kRingWords = kRingBytes / 4;

template <int BLOCK>
struct ThreadCfg {
  static constexpr int WARPS = BLOCK / 32;
  // one contiguous block per warp: interleaved rings, then the walk pool (96 tasks, 96 results,
  // counter) -- every pointer of a warp is a constant offset from one base register
  static constexpr int RING_WARP_BYTES = kRingWords * 32 * 4;
  static constexpr int TASK_OFF = RING_WARP_BYTES;
  static constexpr int RES_OFF = TASK_OFF + 96 * 8;
  static constexpr int CTR_OFF = RES_OFF + 96 * 4;
  static constexpr int WARP_BYTES = CTR_OFF + 16;
  static constexpr int WARP_BASE = kSmemRingBase;
  static constexpr int SMEM = WARP_BASE + WARPS * WARP_BYTES;
};

// CHUNK = true: the work items are speculative chunks of long documents (chunking.cuh).  A chunk
// starts at a byte offset where the capcode encoder's state is known to be the initial one (right
// after a non-alphanumeric, non-space, non-apostrophe ASCII character), assumes that a token
// starts there, and runs OV normalized bytes past the start of the next chunk; every emitted id
// is logged with the normalized position (and forward-delete flag) of the step that emitted it,
// so that a stitch kernel can find the step where this chunk's sequence joins the next chunk's.
template <int BLOCK, int CAP, bool STORE, bool CHUNK = false>
__global__ void __launch_bounds__(BLOCK, 1) tokenize_thread_kernel(const TokenizeArgs A) {
  using Cfg = ThreadCfg<BLOCK>;
  constexpr unsigned FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) uint8_t smem[];
  stage_tables_tma(smem, A.v.smem_image);   // root_child[256], begin_byte[256] (+ barrier inside)
  uint8_t* s_cls = smem + kSmemTableBytes + 16;   // 240 bytes are free behind the mbarrier: classes of 0..127
  if (threadIdx.x < 128) {
    const uint32_t c = threadIdx.x;
    s_cls[c] = uint8_t((c - 'a' < 26u) ? CL_L : (c - 'A' < 26u) ? CL_U : (c - '0' < 10u) ? CL_N : (c == '\'') ? CL_A : CL_O);
  }
  __syncthreads();
  const uint32_t* s_root = reinterpret_cast<const uint32_t*>(smem);
  const uint8_t* s_begin = smem + 1024;
  const uint32_t* s_sp = reinterpret_cast<const uint32_t*>(smem + 1280);   // child of the " " state by byte
  const int lane = threadIdx.x & 31;
  uint8_t* wbase = smem + Cfg::WARP_BASE + (threadIdx.x >> 5) * Cfg::WARP_BYTES;
  uint32_t* ring = reinterpret_cast<uint32_t*>(wbase) + lane;          // word w of this lane: ring[(w & 31) * 32]
  uint2* task = reinterpret_cast<uint2*>(wbase + Cfg::TASK_OFF);
  uint32_t* res = reinterpret_cast<uint32_t*>(wbase + Cfg::RES_OFF);

  const VocabDev& V = A.v;
  const int M = V.max_len;
  const int LA = 2 * M + 2;
  const bool want_ascii = !A.prenorm && ((V.normflag != 0) || (CAP != 0));
  const bool to_lower = !A.prenorm && (V.normflag & 2) != 0;
  const bool can_retry = (V.delete_token != 0xFFFFFFu) && (V.space_start != kNoVal);

  // byte c of a ring column: word (c >> 2) & 31 of the column sits 128 bytes after the previous one
  // (the 32 lanes' words are interleaved), so a single byte load does it
  const uint32_t rb = smem_u32(ring);
  auto col_byte = [](uint32_t colb, int c) -> uint32_t {
    uint32_t v;
    const uint32_t a = colb + ((uint32_t(c) << 5) & uint32_t((kRingWords - 1) << 7)) + (uint32_t(c) & 3u);
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
  };
  auto ring_byte = [&](int c) -> uint32_t { return col_byte(rb, c); };
  auto cls_of = [&](uint32_t c) -> uint32_t { return s_cls[c & 127u]; };

  // ---- per-lane string state ------------------------------------------------------------------
  bool running = false, no_more = false;
  uint32_t s = 0;
  int slen = 0, galign = 0, cshift = 0, total_c = 0;
  const uint8_t* ga = nullptr;
  uint64_t slab_base = 0;
  uint32_t cap = 0;
  int filled = 0;                  // ring coordinates [0, filled) hold normalized text
  bool bad = false;
  int rc = 0;                      // CAP 2: aligned coordinate of the next raw character
  uint4 q = make_uint4(0, 0, 0, 0);
  int qbase = -16;                 // CAP 2: q holds raw bytes [qbase, qbase+16)
  uint64_t win = 0;                // CAP 2: the next `avail` raw bytes, starting at rc
  int avail = 0;
  bool in_word = false, mixed = false;
  uint32_t cl1 = CL_O, cl2 = CL_O; // classes of the previous two raw characters
  bool sp1 = false;                // previous raw character was a space
  uint32_t ow = 0;                 // CAP 2: partially assembled ring word
  int i = 0, fd = 0, len = 0;
  uint32_t idx = 0;
  bool have = false;
  uint32_t cnt = 0, miss = 0;
  // CHUNK: aligned raw coordinate where the next chunk starts, the normalized length up to there
  // (-1 until the encoder has passed it), and whether this is the document's last chunk
  int cross_c = 0, n_norm = -1;
  bool last_chunk = false;

  auto load16 = [&](int c0) -> uint4 {           // raw bytes [c0, c0+16) of the aligned stream
    if (c0 + 16 <= total_c) return __ldg(reinterpret_cast<const uint4*>(ga + c0));
    uint32_t t[4] = {0, 0, 0, 0};                // ragged tail: never read past the string
    for (int k = 0; c0 + k < total_c; k++) t[k >> 2] |= uint32_t(ga[c0 + k]) << (8 * (k & 3));
    return make_uint4(t[0], t[1], t[2], t[3]);
  };
  auto check16 = [&](uint4& w, int c0) {         // non-ASCII detection + optional lower-casing
    uint32_t m0 = w.x, m1 = w.y, m2 = w.z, m3 = w.w;
    if (c0 == 0 && galign) {                     // bytes before the string start are foreign
      uint32_t t[4] = {m0, m1, m2, m3};
      for (int k = 0; k < galign; k++) t[k >> 2] &= ~(0xFFu << (8 * (k & 3)));
      m0 = t[0]; m1 = t[1]; m2 = t[2]; m3 = t[3];
    }
    if ((m0 | m1 | m2 | m3) & 0x80808080u) bad = true;
    if (to_lower) { w.x = lower4(w.x); w.y = lower4(w.y); w.z = lower4(w.z); w.w = lower4(w.w); }
  };
  // adv = characters the step advances by (what the training worker adds to scores[id],
  // training/trainvocab.go:1109-1162); only the count kernels carry the histogram
  auto emit = [&](uint32_t id, uint32_t pos_code = 0, uint32_t adv = 0) {
    if (!STORE && A.scores) atomicAdd(&A.scores[id], static_cast<unsigned long long>(adv));
    if (STORE && cnt < cap) {
      A.slab[slab_base + cnt] = id;
      if (CHUNK) {
        A.c_pos[slab_base + cnt] = pos_code;
        A.c_miss[slab_base + cnt] = miss;
      }
    }
    cnt++;
  };

  for (;;) {
    __syncwarp();
    // ============================ take a new string ============================================
    if (!running && !no_more) {
      const uint32_t item = atomicAdd(A.work_counter, 1u);
      if (item >= A.n) {
        no_more = true;
      } else if (A.list && A.ends[item] < A.off[item]) {
        // the normalizer refused this string: its status 13 from the first pass stands
      } else if (!CHUNK && A.skip && A.skip[item]) {
        // a long document that is tokenized in chunks by another launch
      } else if (CHUNK && A.skip[A.chunks[item].doc]) {
        // no safe chunk boundaries in this document (planning failed): it goes to the one-lane path
        ChunkOut co;
        co.cnt = 0; co.n_norm = 0; co.miss = 0; co.flags = 2u;
        A.c_out[item] = co;
      } else {
        s = A.list ? A.list[item] : item;
        int64_t o0 = A.off[CHUNK ? 0 : item];
        if (CHUNK) {
          const ChunkItem ci = A.chunks[item];
          o0 = ci.start;
          slen = int((A.ends ? A.ends[ci.doc] : A.off[ci.doc + 1]) - o0);
          last_chunk = (ci.flags & 1u) != 0;
          n_norm = (CAP == 0 && !last_chunk) ? int(ci.next - ci.start) : -1;
          cross_c = int(ci.next - ci.start);           // + galign below
        } else {
          slen = int((A.ends ? A.ends[item] : A.off[item + 1]) - o0);
        }
        if (A.watermark) {   // the batch is still streaming in: wait until this string has landed
          const unsigned long long need_bytes = static_cast<unsigned long long>(o0 + slen);
          while (*reinterpret_cast<const volatile unsigned long long*>(A.watermark) < need_bytes) __nanosleep(500);
        }
        const uint8_t* g0 = A.data + o0;
        if (CHUNK) {
          slab_base = uint64_t(item) * A.c_cap;
          cap = A.c_cap;
        } else if (A.list) {   // placement follows the string's own (raw) extent, like the first pass
          const int64_t r0 = A.raw_off[s];
          const int64_t r1 = A.raw_ends ? A.raw_ends[s] : A.raw_off[s + 1];
          slab_base = A.slab_off ? A.slab_off[s] : uint64_t(r0 - A.raw_off[0]) + uint64_t(A.slab_pad) * s;
          cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(r1 - r0) + A.slab_pad;
        } else {
          slab_base = A.slab_off ? A.slab_off[s] : uint64_t(o0 - A.off[0]) + uint64_t(A.slab_pad) * s;
          cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(slen) + A.slab_pad;
        }
        galign = int(reinterpret_cast<uintptr_t>(g0) & 15);
        ga = g0 - galign;
        if (CHUNK) cross_c += galign;
        cshift = (CAP == 0) ? galign : 0;
        total_c = galign + slen;
        filled = 0; bad = false;
        rc = galign; qbase = -16; win = 0; avail = 0;
        in_word = false; mixed = false; cl1 = CL_O; cl2 = CL_O; sp1 = false; ow = 0;
        i = 0; fd = 0; len = 0; idx = 0; have = false; cnt = 0; miss = 0;
        running = true;
      }
    }
    if (__all_sync(FULL, !running)) break;

    // ============================ FILL =========================================================
    if (CAP == 0) {
      for (;;) {
        // the loop runs while some lane NEEDS bytes; every lane with ring space joins in
        const bool must = running && filled < total_c && filled < i + cshift + LA;
        if (!__any_sync(FULL, must)) break;
        const bool want = running && filled < total_c && filled + 16 <= i + cshift + (kRingBytes - 16);
        if (want) {
          uint4 w = load16(filled);
          if (want_ascii) check16(w, filled);
          const int w0 = filled >> 2;
          ring[((w0 + 0) & (kRingWords - 1)) * 32] = w.x;
          ring[((w0 + 1) & (kRingWords - 1)) * 32] = w.y;
          ring[((w0 + 2) & (kRingWords - 1)) * 32] = w.z;
          ring[((w0 + 3) & (kRingWords - 1)) * 32] = w.w;
          filled += 16;
        }
      }
    } else {
      for (;;) {
        const bool must = running && rc < total_c && filled < i + LA;
        if (!__any_sync(FULL, must)) break;
        const bool want = running && rc < total_c && filled + 4 <= i + (kRingBytes - 4);
        if (want) {
          // keep >= 5 raw bytes in the window: 4 for the fast path + 1 to peek at the next character
          while (avail < 5 && rc + avail < total_c) {
            const int nwc = rc + avail;            // first coordinate not yet in the window
            {
              const int wc = nwc & ~3;
              const int cb = wc & ~15;
              if (cb != qbase) {
                q = load16(cb);
                qbase = cb;
                check16(q, cb);
              }
              const int wi = (wc >> 2) & 3;
              const uint32_t wq = (wi == 0) ? q.x : (wi == 1) ? q.y : (wi == 2) ? q.z : q.w;
              const int skip = nwc - wc;           // non-zero only for the first, unaligned word
              win |= uint64_t(wq >> (8 * skip)) << (8 * avail);
              avail += 4 - skip;
            }
          }
          // fast path: four characters that are all lower-case letters or spaces, after a letter or a
          // space, encode to themselves (js:969-974: no "D " is due; js:976-979: no capital follows
          // inside; the space rule only needs the 5th character)
          {
            const uint32_t w4 = uint32_t(win);
            const uint32_t low = (w4 + 0x1F1F1F1Fu) & ~(w4 + 0x05050505u);             // 'a'..'z'
            const uint32_t y = w4 ^ 0x20202020u;
            const uint32_t spc = ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y);              // == ' '
            const uint32_t c5 = uint32_t(win >> 32) & 0xFFu;
            const bool prev_ok = sp1 || cl1 == CL_U || cl1 == CL_L;
            const bool tail_ok = ((w4 >> 24) != ' ') || (rc + 4 >= total_c) || !(c5 - 'A' < 26u);
            if (prev_ok && tail_ok && rc + 4 <= total_c && (w4 & 0x80808080u) == 0 &&
                ((low | spc) & 0x80808080u) == 0x80808080u) {
              const uint32_t sh = (filled & 3) * 8;
              const uint64_t acc = uint64_t(ow) | (uint64_t(w4) << sh);
              ring[((filled >> 2) & (kRingWords - 1)) * 32] = uint32_t(acc);   // 4 bytes always complete a word
              ow = uint32_t(acc >> 32);
              filled += 4;
              win >>= 32;
              avail -= 4;
              rc += 4;
              const uint32_t b2 = (w4 >> 16) & 0xFFu, b3 = w4 >> 24;
              in_word = false;
              cl2 = (b2 == ' ') ? uint32_t(CL_O) : uint32_t(CL_L);
              cl1 = (b3 == ' ') ? uint32_t(CL_O) : uint32_t(CL_L);
              sp1 = (b3 == ' ');
              continue;
            }
          }
          const uint32_t c = uint32_t(win) & 0xFFu;
          const bool has_next = (rc + 1 < total_c);
          const uint32_t cn = has_next ? (uint32_t(win >> 8) & 0xFFu) : uint32_t('.');
          win >>= 8;
          avail--;
          rc++;
          const uint32_t cl = cls_of(c), cln = cls_of(cn);
          const bool letter1 = (cl1 == CL_U) || (cl1 == CL_L);
          const bool glue1 = letter1 || (cl1 == CL_A);
          const bool start = (cl == CL_U) && !in_word;
          if (start) {
            // W (the word stays capital) or C (mixed case): class of the first L/O character
            // after the run of capitals / digits / apostrophes that starts here
            uint32_t k_cl = cln;
            if (!(k_cl == CL_L || k_cl == CL_O)) {     // rare: a run of 2+ such characters
              k_cl = CL_O;
              for (int jj = rc + 1; jj < total_c; jj++) {
                uint32_t cj = ga[jj];
                if (to_lower && cj - 'A' < 26u) cj += 32;
                const uint32_t x = cls_of(cj);
                if (x == CL_L || x == CL_O) { k_cl = x; break; }
              }
            }
            mixed = (k_cl == CL_L);
          }
          // what goes in front of the character
          const bool joined_l = in_word ? glue1 : (sp1 || letter1 || (((cl2 == CL_U) || (cl2 == CL_L)) && cl1 == CL_A));
          const bool joined_n = in_word ? (cl1 == CL_N) : (sp1 || cl1 == CL_N);
          const bool k_dspace = (cl == CL_L && !joined_l) || (cl == CL_N && !joined_n) ||
                                (cl == CL_U && in_word && !mixed && !glue1);
          const bool k_dcs = (cl == CL_U) && in_word && mixed;
          const uint32_t marker = mixed ? 'C' : 'W';
          uint32_t pre = 0, np = 0;
          if (k_dspace) { pre = 'D' | (' ' << 8); np = 2; }
          if (k_dcs) { pre = 'D' | ('C' << 8) | (' ' << 16); np = 3; }
          if (start) {
            pre = sp1 ? (marker | (' ' << 8)) : ('D' | (marker << 8) | (' ' << 16));
            np = sp1 ? 2 : 3;
          }
          // a space right before the first letter of a capitalised word is replaced by the marker
          // (js:976-979): the capital emits "<marker> x" and the space emits nothing
          const bool eaten = (c == ' ') && has_next && (cln == CL_U);
          const uint32_t ch = (cl == CL_U) ? (c | 0x20u) : c;
          const uint64_t bytes = uint64_t(pre) | (eaten ? 0ull : (uint64_t(ch) << (8 * np)));
          const uint32_t total = np + (eaten ? 0u : 1u);
          // append `total` bytes to the ring through the partial word ow
          const uint32_t sh = (filled & 3) * 8;
          const uint64_t acc = uint64_t(ow) | (bytes << sh);
          const int nf = filled + int(total);
          if ((nf >> 2) != (filled >> 2)) {
            ring[((filled >> 2) & (kRingWords - 1)) * 32] = uint32_t(acc);
            ow = uint32_t(acc >> 32);
          } else {
            ow = uint32_t(acc);
          }
          filled = nf;
          if (CHUNK && rc == cross_c && n_norm < 0 && !last_chunk) n_norm = filled;
          // carries
          in_word = (cl == CL_U) ? true : ((cl == CL_L || cl == CL_O) ? false : in_word);
          cl2 = cl1;
          cl1 = cl;
          sp1 = (c == ' ');
        }
      }
      if (running && (filled & 3)) ring[((filled >> 2) & (kRingWords - 1)) * 32] = ow;   // publish the partial word
    }

    // ============================ step setup ===================================================
    const bool exhausted = (CAP == 0) ? true : (rc >= total_c);
    const int len_data = (CAP == 0) ? slen : (exhausted ? filled : 0x3FFFFFFF);
    const bool past = CHUNK && !last_chunk && n_norm >= 0 && i >= n_norm + int(A.c_ov);
    // the encoder went past the next chunk's start without landing on it: cannot happen behind a
    // safe boundary; stop rather than tokenize the rest of the document
    const bool lost = CHUNK && CAP != 0 && !last_chunk && n_norm < 0 && rc > cross_c;
    const bool finished = running && (bad || i >= len_data || past || lost);
    const bool stepping = running && !finished;
    const bool seek = stepping && !have;           // this iteration only finds the current token
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
    if (stepping && have) {
      r0 = __ldg(&V.rec[2 * idx]);
      r1 = __ldg(&V.rec[2 * idx + 1]);
    }
    const uint32_t oflag = r0.x & 0xFFu;
    const int alen1 = int(r0.x >> 24), alen2 = int(r1.x & 0xFFu);
    const int i1 = i + len;
    bool look = false;
    if (stepping && have)
      look = (i1 < len_data) && (((oflag & 32u) == 0) || s_begin[ring_byte(i1 + cshift)] != 12);   // go:1057
    // the (up to) three look-ahead positions; a seeking lane has the single position i
    const int base0 = seek ? 0 : len;
    const int base1 = alen1 - fd, base2 = alen2 - fd;
    uint32_t todo = 0;                              // bit j: walk at position j still to do
    if (stepping) {
      todo = 1u;                                    // position 0 is always wanted (next current token)
      if (have && look && alen1 != 0) todo |= 2u;
      if (have && look && alen1 != 0 && alen2 != 0) todo |= 4u;   // nested: go:1163
    }
    uint32_t x0 = 0, x1 = 0, x2 = 0;
    int l0 = 0, l1 = 0, l2 = 0;

    // ============================ WALK (two passes share the code) =============================
    // All walks of the warp's 32 strings go into one pool in shared memory and the 32 lanes take
    // them one at a time (shared-memory atomic counter), whoever the owner is: a long walk no
    // longer keeps 31 lanes waiting, the loop runs for ~(total trie steps)/32 iterations.
    // Task = {ring coordinate, owner lane | position j | max depth}; result = len<<24 | index.
    uint32_t need = 0;                              // bit j: retry wanted at position j
    uint32_t xb0 = 0, xb1 = 0, xb2 = 0;
    int lb0 = 0, lb1 = 0, lb2 = 0;
    uint32_t bflag0 = 0, bflag1 = 0, bflag2 = 0;
    int bnw0 = 0, bnw1 = 0, bnw2 = 0, nb0 = 0, nb1 = 0, nb2 = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      const uint32_t queue = (pass == 0) ? todo : need;
      // compact the tasks of all lanes: position j of lane t gets slot  (#tasks of lower j) + rank
      const uint32_t m0 = __ballot_sync(FULL, queue & 1u), m1 = __ballot_sync(FULL, queue & 2u),
                     m2 = __ballot_sync(FULL, queue & 4u);
      const int n0 = __popc(m0), n1 = __popc(m1), ntask = n0 + n1 + __popc(m2);
      if (ntask == 0) {
        if (pass == 1) break;
        continue;
      }
      const uint32_t lt = (1u << lane) - 1u;
      const int maxd_cap = (pass == 0) ? M : (M - V.off);
#pragma unroll
      for (int jj = 0; jj < 3; jj++) {
        res[jj * 32 + lane] = 0;
        if (queue & (1u << jj)) {
          const int slot = (jj == 0 ? 0 : (jj == 1 ? n0 : n0 + n1)) + __popc((jj == 0 ? m0 : (jj == 1 ? m1 : m2)) & lt);
          const int p = i + (jj == 0 ? base0 : (jj == 1 ? base1 : base2));
          task[slot] = make_uint2(uint32_t(p + cshift),
                                  uint32_t(lane) | (uint32_t(jj) << 5) | (uint32_t(max(0, min(len_data - p, maxd_cap))) << 8));
        }
      }
      __syncwarp();
      // The pool is handed out by rank among the idle lanes (one ballot, no atomics).  A lane that
      // takes a task does the first steps right away, without the hash table: the plain walk reads
      // the node two bytes below the root from a 64 K-entry table (the top of the trie is where
      // every walk goes), the retry walk the child of the " " state from shared memory.
      bool walking = false, dry = false;
      int c = 0, d = 0, maxd = 0, best_len = 0, slot_out = 0;
      uint32_t node = 0, best = 0, sig = 0;
      uint32_t next = 0;                            // tasks handed out so far (warp-uniform)
      uint32_t colb = rb;                           // ring column of the task's owner (shared address)
      for (;;) {
        const bool idle = !walking && !dry;
        const uint32_t im = __ballot_sync(FULL, idle);
        if (idle) {
          const uint32_t t = next + __popc(im & lt);
          if (t < uint32_t(ntask)) {
            const uint2 tk = task[t];
            const int owner = int(tk.y & 31u);
            c = int(tk.x);
            maxd = int(tk.y >> 8);
            slot_out = int((tk.y >> 5) & 3u) * 32 + owner;
            colb = rb + uint32_t(owner - lane) * 4u;
            best = 0;
            best_len = 0;
            d = 0;
            uint32_t val = kNoVal;
            if (maxd > 0) {
              const uint32_t b0 = col_byte(colb, c);
              if (pass == 0) {
                const uint32_t v1 = s_root[b0];
                if (v1 != kNoVal) {
                  d = 1;
                  val = v1;
                  if ((v1 & 0xFFFFFFu) < V.n_info) { best = v1 & 0xFFFFFFu; best_len = 1; }
                  if (maxd > 1) {
                    const uint32_t b1 = col_byte(colb, c + 1);
                    const uint32_t v2 = __ldg(&V.root2[(b0 << 8) | b1]);
                    if (v2 != kNoVal) {
                      d = 2;
                      val = v2;
                      if ((v2 & 0xFFFFFFu) < V.n_info) { best = v2 & 0xFFFFFFu; best_len = 2; }
                    } else {
                      val = kNoVal;                 // no node two bytes down: the walk ends at depth 1
                    }
                  }
                }
              } else {
                const uint32_t v1 = s_sp[b0];
                if (v1 != kNoVal) {
                  d = 1;
                  val = v1;
                  if ((v1 & 0xFFFFFFu) < V.n_info) { best = v1 & 0xFFFFFFu; best_len = 1; }
                }
              }
            }
            node = val & 0xFFFFFFu;
            sig = val >> 24;
            walking = (val != kNoVal) && (d < maxd) && (sig != 0u);
            if (!walking) res[slot_out] = best | (uint32_t(best_len) << 24);
          } else {
            dry = true;
          }
        }
        next += __popc(im);
        if (!__any_sync(FULL, walking)) {
          if (next >= uint32_t(ntask)) break;
          continue;
        }
        if (walking) {
          const uint32_t b = col_byte(colb, c + d);
          uint32_t val = kNoVal;
          if ((sig >> sig_bit_of(b)) & 1u) {        // else: no child of this node starts with b
            const uint32_t key = (node << 8) | b;
            uint32_t h = (key * 0x9E3779B1u) >> V.edge_shift;
            uint2 e = __ldg(&V.edges[h]);
            while (e.x != key && e.x != kNoVal) {    // linear probing, load factor <= 1/4
              h = (h + 1) & V.edge_mask;
              e = __ldg(&V.edges[h]);
            }
            if (e.x == key) val = e.y;
          }
          if (val != kNoVal) {
            node = val & 0xFFFFFFu;
            sig = val >> 24;
            d++;
            if (node < V.n_info) { best = node; best_len = d; }
          }
          if (val == kNoVal || d >= maxd || sig == 0u) {
            walking = false;
            res[slot_out] = best | (uint32_t(best_len) << 24);
          }
        }
      }
      __syncwarp();
      {
        const uint32_t ra = res[lane], rb = res[32 + lane], rcv = res[64 + lane];
        if (pass == 0) {
          x0 = ra & 0xFFFFFFu; l0 = int(ra >> 24);
          x1 = rb & 0xFFFFFFu; l1 = int(rb >> 24);
          x2 = rcv & 0xFFFFFFu; l2 = int(rcv >> 24);
        } else {
          xb0 = ra & 0xFFFFFFu; lb0 = int(ra >> 24);
          xb1 = rb & 0xFFFFFFu; lb1 = int(rb >> 24);
          xb2 = rcv & 0xFFFFFFu; lb2 = int(rcv >> 24);
        }
      }
      __syncwarp();
      if (pass == 0) {
        // second-token attributes of the plain results and which positions want the retry
        const bool sc = stepping && have && look;
        if (sc && l0 > 0) { const uint32_t w = __ldg(&V.rec[2 * x0]).x; bflag0 = w & 0xFFu; bnw0 = int((w >> 8) & 0xFFu); const int qq = i + base0 + l0; nb0 = s_begin[(qq < len_data) ? ring_byte(qq + cshift) : 0u]; }
        if (sc && l1 > 0) { const uint32_t w = __ldg(&V.rec[2 * x1]).x; bflag1 = w & 0xFFu; bnw1 = int((w >> 8) & 0xFFu); const int qq = i + base1 + l1; nb1 = s_begin[(qq < len_data) ? ring_byte(qq + cshift) : 0u]; }
        if (sc && l2 > 0) { const uint32_t w = __ldg(&V.rec[2 * x2]).x; bflag2 = w & 0xFFu; bnw2 = int((w >> 8) & 0xFFu); const int qq = i + base2 + l2; nb2 = s_begin[(qq < len_data) ? ring_byte(qq + cshift) : 0u]; }
        if (sc && can_retry) {
          if (l0 > 0 && (bflag0 & 2u) && nb0 == 1 && bnw0 == 0) need |= 1u;   // go:1088
          if (l1 > 0 && (bflag1 & 2u) && nb1 == 1 && bnw1 == 0) need |= 2u;
          if (l2 > 0 && (bflag2 & 2u) && nb2 == 1 && bnw2 == 0) need |= 4u;
        }
      }
    }

    // ============================ SCORE + PICK =================================================
    if (seek) {
      if (l0 == 0) {                                // no token covers this byte (go:1269-1276)
        if (!STORE && A.scores) {                   // trainvocab.go:1164-1171: counted, never scored
          cnt++;
          atomicAdd(&A.miss_hist[ring_byte(i + cshift)], 1ull);
        } else if (V.unk != 0xFFFFFFu) {
          emit(V.unk, uint32_t(i));
        }
        i++;
        miss++;
        fd = 0;
      } else {
        idx = x0;
        len = l0;
        have = true;
      }
    } else if (stepping) {
      const int onw = int((r0.x >> 8) & 0xFFu);
      const uint32_t aflag1 = (r1.x >> 8) & 0xFFu, aflag2 = (r1.x >> 24) & 0xFFu;
      const int anw1 = int((r1.x >> 16) & 0xFFu), anw2 = int(r1.y & 0xFFu);
      int best_s = kNegScore, pick = -1;            // pick: 0..2 plain, 3..5 retry
      if (look) {
        // plain candidates in the order s1, s2, s3 (strict > keeps the first maximum)
        if (l0 > 0) {
          const int sc = pair_score(base0 + l0, oflag, onw - fd, bflag0, bnw0, nb0, false);
          if (sc > best_s) { best_s = sc; pick = 0; }
        }
        if ((todo & 2u) && l1 > 0) {
          const int bl = base1 + l1;
          const int sc = pair_score(bl, aflag1, anw1 - fd, bflag1, bnw1, nb1, false) - 100 * int(bl < len) - 10000 * int(bl == len);
          if (sc > best_s) { best_s = sc; pick = 1; }
        }
        if ((todo & 4u) && l2 > 0) {
          const int bl = base2 + l2;
          const int sc = pair_score(bl, aflag2, anw2 - fd, bflag2, bnw2, nb2, false) - 100 * int(bl < len) - 10000 * int(bl == len);
          if (sc > best_s) { best_s = sc; pick = 2; }
        }
        // retries in the order s1b, s2b, s3b; a plain score keeps a tie
        if ((need & 1u) && lb0 + V.off > l0 + 1) {
          const uint32_t w = __ldg(&V.rec[2 * xb0]).x;
          const int qq = i + base0 + lb0;
          const int nb = s_begin[(qq < len_data) ? ring_byte(qq + cshift) : 0u];
          const int sc = pair_score(base0 + lb0, oflag, onw - fd, w & 0xFFu, int((w >> 8) & 0xFFu), nb, true);
          if (sc > best_s) { best_s = sc; pick = 3; }
        }
        if ((need & 2u) && lb1 + V.off > l1 + 1) {
          const uint32_t w = __ldg(&V.rec[2 * xb1]).x;
          const int bl = base1 + lb1, qq = i + bl;
          const int nb = s_begin[(qq < len_data) ? ring_byte(qq + cshift) : 0u];
          const int sc = pair_score(bl, aflag1, anw1 - fd, w & 0xFFu, int((w >> 8) & 0xFFu), nb, true) - 100 * int(bl < len) - 10000 * int(bl == len);
          if (sc > best_s) { best_s = sc; pick = 4; }
        }
        if ((need & 4u) && lb2 + V.off > l2 + 1) {
          const uint32_t w = __ldg(&V.rec[2 * xb2]).x;
          const int bl = base2 + lb2, qq = i + bl;
          const int nb = s_begin[(qq < len_data) ? ring_byte(qq + cshift) : 0u];
          const int sc = pair_score(bl, aflag2, anw2 - fd, w & 0xFFu, int((w >> 8) & 0xFFu), nb, true) - 100 * int(bl < len) - 10000 * int(bl == len);
          if (sc > best_s) { best_s = sc; pick = 5; }
        }
      }
      const uint32_t id0 = r0.y, id1 = r0.z, id2 = r0.w;
      if (pick < 0) {
        emit(id0, uint32_t(i) | (uint32_t(fd) << 31), uint32_t(len));   // go:1265-1267
        i += len;
        fd = 0;
        have = (l0 > 0);
        idx = x0;
        len = l0;
      } else {
        const int br = pick >= 3 ? pick - 3 : pick;
        emit(br == 0 ? id0 : (br == 1 ? id1 : id2), uint32_t(i) | (uint32_t(fd) << 31),
             uint32_t(br == 0 ? base0 : (br == 1 ? base1 : base2)));
        // Count() adds only 1 here (quirk Q2); the training worker counts and scores the delete token
        if (pick >= 3 && (STORE || A.scores)) emit(V.delete_token, 0xFFFFFFFFu, 1u);
        i += (br == 0 ? base0 : (br == 1 ? base1 : base2));
        if (pick >= 3) { idx = (br == 0 ? xb0 : (br == 1 ? xb1 : xb2)); len = (br == 0 ? lb0 : (br == 1 ? lb1 : lb2)); fd = 1; }
        else { idx = (br == 0 ? x0 : (br == 1 ? x1 : x2)); len = (br == 0 ? l0 : (br == 1 ? l1 : l2)); fd = 0; }
        have = true;
      }
    }
    if (finished && CHUNK) {
      ChunkOut co;
      co.cnt = min(cnt, cap);
      co.n_norm = uint32_t(n_norm < 0 ? 0 : n_norm);
      co.miss = miss;
      co.flags = (bad ? 1u : 0u) | ((cnt > cap || lost) ? 2u : 0u) | ((i >= len_data && !bad && !lost) ? 4u : 0u);
      A.c_out[s] = co;
      running = false;
    } else if (finished) {
      if (bad) { cnt = 0; miss = 0; }
      A.counts[s] = STORE ? min(cnt, cap) : cnt;
      A.need[s] = cnt;
      A.missing[s] = miss;
      A.status[s] = bad ? 13 : 0;
      if (STORE && cnt > cap) atomicExch(A.overflow, 1u);
      running = false;
    }
  }
}

}  // namespace tmgpu

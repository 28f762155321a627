// pipeline_kernels.cuh -- K1/K3 of round 2: the position-table pipeline (pipeline_core.h has the
// per-string logic and the reference citations; vocab_tables.h the automaton).
//
//   prep_kernel            per string: anything to do? slot size, sort key (length / 8)
//   (scan, radix sort)     slot offsets; strings by descending length
//   table_kernel           capcode encoder + reversed-key automaton in ONE right-to-left loop:
//                          one 32-bit word per normalized position goes to HBM, the normalized
//                          text itself never exists in memory
//   chain_kernel<STORE>    the ungreedy loop over the position words -> ids / counts
//
// Why passes over HBM instead of one fused kernel: the round-1 kernel kept everything in
// shared-memory rings and paid for it with ~12 dependent trie probes per emitted id under SIMT
// divergence (165 warp instructions per id, 13 of 32 lanes active, profiles/r01_*).  Here the
// automaton needs ONE table probe per byte whatever the token lengths, the chain stage has no
// inner loop at all, and every lane of a warp does the same thing in every iteration; the price is
// ~9 bytes of HBM traffic per input byte, on a GPU whose HBM was 99 % idle.
//
// One lane per string, 32 strings of EQUAL LENGTH per warp (the sort): the first version of this
// pipeline let every lane pull its next string the moment its own ended, and rare per-lane events
// (a 16-byte refill, a store, the fetch of a new string) then happen at a different iteration in
// every lane, so the warp executes their code almost every iteration for one or two lanes (ncu:
// 13-20 of 32 lanes active, profiles/r02_v1_*).  With equal lengths the lanes stay in phase: they
// start together, refill together and end together.  Per-lane loads are 16-byte vectors
// wherever the data is a stream, and what a lane produces (position words, ids) collects in a
// shared-memory row and leaves as whole 32-byte sectors: with one string per lane every memory
// instruction touches 32 different lines, and the L1 pipeline charges per line.
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
// prep: one thread per work item -- is there anything to do, how many words may it need, and a
// sort key (length / 8) so that the 32 strings a warp works on together have the same length
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_kernel(const TokenizeArgs A, int cap_mode, int slot_full) {
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < A.n; item += gridDim.x * blockDim.x) {
    uint32_t s = 0;
    int64_t o0 = 0;
    int slen = 0;
    const bool ok = pipe_fetch(A, item, &s, &o0, &slen);
    uint64_t slot = 0;
    if (ok) {
      const uint64_t l = uint64_t(slen);
      // capcode 2 grows ASCII text by a few percent (a capital costs up to 3 extra bytes): room for
      // 1.25 x + 48, the rare string that needs more is redone with 4 x + 32 (slot_full)
      slot = !cap_mode ? l : (slot_full ? 4 * l + 32 : l + (l >> 2) + 48);
      slot = (slot + 15) & ~uint64_t(15);
      if (slot > 0xFFFFFFF0ull) slot = 0xFFFFFFF0ull;
    }
    A.npad[item] = uint32_t(slot);
    A.sort_key[item] = ok ? min(uint32_t(slen >> 3) + 1u, 0xFFFFu) : 0u;
    A.sort_val[item] = item;
    A.pflag[item] = ok ? 0 : 2;
  }
}

// strings whose text did not fit the optimistic slot (status 14): listed with their extents and the
// words a full-size slot needs, for the second run
__global__ void __launch_bounds__(256) select_overflowed_kernel(const uint8_t* __restrict__ status,
                                                                const int64_t* __restrict__ off,
                                                                const int64_t* __restrict__ ends, uint32_t n,
                                                                uint32_t* __restrict__ list, int64_t* __restrict__ o_begin,
                                                                int64_t* __restrict__ o_end,
                                                                unsigned long long* __restrict__ ctr) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (status[s] != 14) continue;
    const int64_t b0 = off[s], b1 = ends ? ends[s] : off[s + 1];
    const unsigned long long k = atomicAdd(&ctr[0], 1ull);
    atomicAdd(&ctr[1], (4ull * static_cast<unsigned long long>(b1 - b0) + 32ull + 15ull) & ~15ull);
    list[k] = s;
    o_begin[k] = b0;
    o_end[k] = b1;
  }
}

// ------------------------------------------------------------------------------------------
// stages 1 + 2: capcode encoder and reversed-key automaton in one right-to-left loop.
// A warp takes 32 strings of (nearly) equal length; every iteration each lane consumes exactly one
// normalized byte: if its byte queue is empty it first encodes the next raw character (0..4 bytes,
// one table read decides), then the automaton takes one byte and the position word goes out (a
// whole 32-byte sector at a time, right aligned in the string's slot, so the slot's start need
// not be known).
// (Splitting the loop into two phases -- encode 8 characters into a shared-memory buffer, then the
// automaton over their bytes -- was measured slower: 6.6 against 5.6 ms.)
// Raw text comes in through a 32-byte ring per lane in shared memory, filled 16 bytes at a time
// by cp.async (SASS: LDGSTS) -- no registers, no wait: the copy is issued when the cursor is
// still 8+ characters away, and by all lanes of the warp in the same iteration.
// CAP: capcode level 2; otherwise the text is read as it is (capcode 0, or text the device
// normalizer produced), lower-cased / checked for non-ASCII bytes when the vocabulary asks for it.
// TEXT: also keep the normalized bytes (the training scores need the byte no token covers).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pipe_cp_async4(uint32_t smem_addr, const void* gptr, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes) : "memory");
}
// one 4-byte word of a string into the ring: asynchronous when the whole word belongs to the string,
// byte by byte for the last partial word (nothing behind the end of the string is ever read)
__device__ __forceinline__ void pipe_ring_word(uint32_t smem_addr, const uint8_t* g, int left) {
  if (left >= 4) {
    pipe_cp_async4(smem_addr, g, 4);
  } else {
    uint32_t w = 0;
    for (int t = 0; t < left; t++) w |= uint32_t(g[t]) << (8 * t);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_addr), "r"(w) : "memory");
  }
}
__device__ __forceinline__ void pipe_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared memory of one CTA, ONE object: every table is read with "ld.shared [base + constant +
// index]" from a single base register (reading separate __shared__ arrays through C++ made the
// compiler rebuild each array's window address -- S2R / MOV / LEA -- at every use in the loop).
struct __align__(128) TableShared {
  uint8_t img[kPipeImageBytes];      // dfa_d1[256] (u32), class_code[256], 256 spare: one TMA bulk copy
  uint32_t lut[kCapLutSize];         // revcap_lut_entry
  uint16_t chr[256];                 // raw byte -> byte to emit | class << 8 (revcap_char)
  uint32_t ring[kPipeBlock * 8];     // raw text: word w of lane l at [warp * 256 + w * 32 + l]
  uint32_t rows[kPipeBlock * 12];    // eight position words per lane waiting for their sector (48-byte rows)
  uint64_t bar;
};
constexpr uint32_t kTsD1 = 0, kTsCode = 1024;
constexpr uint32_t kTsLut = kPipeImageBytes;
constexpr uint32_t kTsChr = kTsLut + 4 * kCapLutSize;
constexpr uint32_t kTsRing = kTsChr + 2 * 256;
constexpr uint32_t kTsRows = kTsRing + 4 * kPipeBlock * 8;
constexpr uint32_t kTsRowBytes = 48;   // 3 x 16: the 128-bit reads of eight lanes fall into eight different bank groups
static_assert(kTsRing % 16 == 0 && kTsRows % 16 == 0, "ring / rows are read and written 16 bytes at a time");

template <uint32_t OFF>
__device__ __forceinline__ uint32_t pipe_lds_u32(uint32_t addr) {   // read-only tables: plain (movable) asm
  uint32_t v;
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <uint32_t OFF>
__device__ __forceinline__ uint32_t pipe_lds_u16(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <uint32_t OFF>
__device__ __forceinline__ uint32_t pipe_lds_u8(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}

template <bool CAP, bool TEXT>
__global__ void __launch_bounds__(kPipeBlock) table_kernel(const TokenizeArgs A) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  __shared__ TableShared sm;
  static_assert(offsetof(TableShared, lut) == kTsLut && offsetof(TableShared, chr) == kTsChr &&
                    offsetof(TableShared, ring) == kTsRing && offsetof(TableShared, rows) == kTsRows,
                "the constants above are the struct's layout");
  pipe_stage_tma(sm.img, A.v.pipe_image, &sm.bar);
  const bool want_ascii = CAP || (!A.prenorm && (A.v.normflag != 0));
  const bool to_lower = !A.prenorm && (A.v.normflag & 2) != 0;
  if (CAP)
    for (int i = threadIdx.x; i < kCapLutSize; i += kPipeBlock) sm.lut[i] = revcap_lut_entry(uint32_t(i));
  {
    uint32_t c = threadIdx.x & 255u;
    if (to_lower && c - 'A' < 26u) c |= 0x20u;
    sm.chr[threadIdx.x & 255] = uint16_t(CAP ? revcap_char(c) : c);
  }
  __syncthreads();
  const DfaDev D = A.v.dfa;
  const uint32_t voff = uint32_t(A.v.off);
  uint32_t sb = pipe_smem_u32(&sm);
  asm volatile("mov.u32 %0, %0;" : "+r"(sb));   // opaque: keeps the base in a register (the compiler would rebuild it at every use)
  const uint32_t code0 = pipe_lds_u8<kTsCode>(sb);
  const int lane = threadIdx.x & 31;
  uint32_t* counter = A.pipe_counters + 2;
  const uint32_t ring_b = sb + kTsRing + uint32_t((threadIdx.x >> 5) * 256 + lane) * 4u;   // byte address of word 0
  const uint32_t row_b = sb + kTsRows + uint32_t(threadIdx.x) * kTsRowBytes;

  for (;;) {
    uint32_t g = 0;
    if (lane == 0) g = atomicAdd(counter, 1u);
    g = __shfl_sync(FULL, g, 0);
    if (uint64_t(g) * 32u >= A.n) break;
    const uint32_t slot_i = g * 32u + uint32_t(lane);
    // ---- this lane's string ---------------------------------------------------------------
    bool active = false;
    uint32_t seen = 0;   // OR of the bytes read so far (bit 7: a byte >= 0x80)
    uint32_t item = 0, s = 0, slot = 0;
    int k = -1, galign = 0, total_c = 0;
    int fcur = -1;       // aligned coordinate of the next raw byte the read-ahead takes
    // The ring holds the 16-byte blocks lo_blk and lo_blk + 1.  Block lo_blk - 1 is fetched once the
    // cursor is 8 bytes into block lo_blk: trig = 16 * lo_blk + 8 while a lower block exists, else
    // far below any cursor (so "fcur <= trig" is the whole test).
    constexpr int kNoRefill = -(1 << 30);
    int trig = kNoRefill;
    const uint8_t* ga = nullptr;
    uint32_t* wend = nullptr;
    uint8_t* tend = nullptr;
    if (slot_i < A.n) {
      item = A.sort_order[slot_i];
      if (A.pflag[item] == 0) {
        int64_t o0 = 0;
        int slen = 0;
        pipe_fetch(A, item, &s, &o0, &slen);
        const uint8_t* g0 = A.data + o0;
        galign = int(reinterpret_cast<uintptr_t>(g0) & 15);
        ga = g0 - galign;
        total_c = galign + slen;
        k = slen - 1;
        const uint64_t e = A.toff[item + 1];
        slot = uint32_t(e - A.toff[item]);
        wend = A.words + e;
        if (TEXT) tend = A.text + e;
        active = true;
        if (e > A.words_cap) {   // the batch is larger than the workspace was created for: refuse
          A.nlen[item] = 0;
          A.pflag[item] = 1;
          A.counts[s] = 0; A.need[s] = 0; A.missing[s] = 0; A.status[s] = 13;
          active = false;
        }
      }
    }
    // 16 raw bytes of block blk -> the ring (bytes behind the end of the string read as 0)
    auto load_block = [&](int blk) {
      const int c0 = blk << 4;
      const uint32_t dst = ring_b + uint32_t((blk & 1) * 4) * 128u;
      if (c0 + 16 <= total_c) {   // the whole block belongs to the string: four asynchronous 4-byte copies
#pragma unroll
        for (int j = 0; j < 4; j++) pipe_cp_async4(dst + uint32_t(j) * 128u, ga + c0 + 4 * j, 4);
      } else {                    // the string's last block (once per string)
        for (int j = 0; j < 4; j++) pipe_ring_word(dst + uint32_t(j) * 128u, ga + c0 + 4 * j, total_c - (c0 + 4 * j));
      }
    };
    // the raw character at coordinate fcur: byte to emit | class << 8 (revcap_char); 0 = "no
    // character" when the string has none left (`has` false: nothing is read, byte 0 maps to 0)
    auto take = [&](bool has) -> uint32_t {
      const int x = fcur;
      fcur -= has ? 1 : 0;
      uint32_t b;
      // entering a block (x & 15 == 15): wait for its copy, which was issued long ago; then one byte
      asm volatile(
          "{\n\t.reg .pred p, h;\n\t"
          "setp.ne.u32 h, %3, 0;\n\t"
          "setp.eq.and.u32 p, %2, 15, h;\n\t"
          "@p cp.async.wait_all;\n\t"
          "mov.u32 %0, 0;\n\t"
          "@h ld.shared.u8 %0, [%1];\n\t}"
          : "=r"(b)
          : "r"(ring_b + uint32_t((x >> 2) & 7) * 128u + uint32_t(x & 3)), "r"(uint32_t(x) & 15u), "r"(uint32_t(has))
          : "memory");
      const uint32_t t = pipe_lds_u16<kTsChr>(sb + 2u * b);
      seen |= t;
      return t;
    };
    uint32_t x0 = 0;   // capcode 0: the character at k
    RevCap e;          // capcode 2: the encoder, with its read-ahead of the characters k, k-1, k-2
    revcap_reset(e);
    if (active && k >= 0) {
      const int top = (total_c - 1) >> 4;
      load_block(top);
      if (top > 0) load_block(top - 1);
      trig = top > 1 ? 16 * (top - 1) + 8 : kNoRefill;
      pipe_cp_async_wait();
      fcur = total_c - 1;
      x0 = take(true);
      if (CAP) {
        revcap_push(e, x0);
        revcap_push(e, take(k >= 1));
        revcap_push(e, take(k >= 2));
      }
    }
    ClassTrail tr;
    trail_reset(tr, code0);
    uint32_t packed = 0, qbytes = 0, qn = 0, dist = 0;

    while (__any_sync(FULL, active)) {
      // ---- ring refill, all lanes of the warp together ----------------------------------------
      if (__any_sync(FULL, active && fcur <= trig)) {
        if (active && fcur <= trig + 7) {   // every lane whose cursor is inside its lowest block joins in
          const int blk = (trig >> 4) - 1;
          trig = blk > 0 ? trig - 16 : kNoRefill;
          load_block(blk);
        }
      }
      if (active) {
        const bool bad = want_ascii & ((seen & 0x80u) != 0);
        if (qn == 0 || bad) {
          if (k < 0 || bad) {
            // ---- the string is done: the last (partial) group of words, the verdict ------------
            if (!bad) {
              const uint32_t r = dist & 7u;   // they wait in the row: word of distance d at row[(-d) & 7]
              for (uint32_t j = 0; j < r; j++) {
                const uint32_t d1 = dist - j;
                uint32_t w;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(row_b + ((0u - d1) & 7u) * 4u) : "memory");
                *(wend - d1) = w;
              }
            }
            A.nlen[item] = bad ? 0u : dist;
            A.pflag[item] = bad ? 1 : 0;
            if (bad) { A.counts[s] = 0; A.need[s] = 0; A.missing[s] = 0; A.status[s] = 13; }
            active = false;
          } else if (CAP) {
            qn = revcap_step(
                e, [&](uint32_t i) -> uint32_t { return pipe_lds_u32<kTsLut>(sb + 4u * i); }, k,
                [&](int j) -> uint32_t { return pipe_lds_u16<kTsChr>(sb + 2u * uint32_t(ga[galign + j])) >> 8; }, &qbytes);
            revcap_push(e, take(k >= 3));
            k--;
          } else {
            qbytes = x0 & 0xFFu;
            qn = 1;
            x0 = take(k >= 1);
            k--;
          }
        }
        if (active && qn != 0 && !bad) {
          const uint32_t b = qbytes & 0xFFu;   // the queue leaves lowest byte first
          qbytes >>= 8;
          qn--;
          if (dist >= slot) {   // more text than the slot holds: redone with a full-size slot
            A.nlen[item] = 0;
            A.pflag[item] = 3;
            A.counts[s] = 0; A.need[s] = 0; A.missing[s] = 0; A.status[s] = 14;
            active = false;
          } else {
            packed = dfa_next(packed, b, D.shift, D.mask,
                              [&](uint32_t i, uint4* kk, uint4* vv) { pipe_load32(D.buckets + 2 * i, kk, vv); },
                              [&](uint32_t c) { return pipe_lds_u32<kTsD1>(sb + 4u * c); });
            const uint32_t code = pipe_lds_u8<kTsCode>(sb + b);
            const uint32_t w = position_word(packed, code, tr, voff);
            trail_push(tr, code);
            dist++;
            if (TEXT) tend[-int64_t(dist)] = uint8_t(b);
            // The word of distance d waits at row[(-d) & 7] of the lane's shared-memory row; every
            // 8th word the row is a whole 32-byte sector and leaves with ONE store (SASS: STG.256 --
            // a narrower store makes L2 fetch the rest of its sector from DRAM).  Predicated, not a
            // branch: the lanes' phases differ.
            const uint32_t ph = (0u - dist) & 7u;
            asm volatile(
                "{\n\t.reg .pred q;\n\t.reg .b32 t0, t1, t2, t3, t4, t5, t6, t7;\n\t"
                "st.shared.u32 [%2], %1;\n\t"
                "setp.eq.u32 q, %3, 0;\n\t"
                "@q ld.shared.v4.u32 {t0, t1, t2, t3}, [%4];\n\t"
                "@q ld.shared.v4.u32 {t4, t5, t6, t7}, [%4+16];\n\t"
                "@q st.global.v8.u32 [%0], {t0, t1, t2, t3, t4, t5, t6, t7};\n\t}" ::"l"(wend - dist),
                "r"(w), "r"(row_b + ph * 4u), "r"(ph), "r"(row_b)
                : "memory");
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// stage 3: the ungreedy loop, 32 strings of equal length per warp.  STORE = ids into the slab
// (K1), otherwise counts (K3, and the training worker's score histogram when A.scores is set).
// ------------------------------------------------------------------------------------------
// (Five CTAs per SM instead of four was measured: the 51-register budget spills and the kernel is
// 0.3 ms slower on c2, 0.7 ms on c3.)
template <bool STORE>
__global__ void __launch_bounds__(kPipeBlock, 4) chain_kernel(const TokenizeArgs A) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  ChainTables T;
  T.crec = A.v.dfa.crec;
  T.hot = A.v.dfa.hot;
  T.delete_token = A.v.delete_token;
  T.unk = A.v.unk;
  T.off = uint32_t(A.v.off);
  T.can_retry = A.v.delete_token != 0xFFFFFFu;
  uint32_t* counter = A.pipe_counters + 3;
  const bool two_ids = STORE || A.scores != nullptr;
  const int lane = threadIdx.x & 31;
  // Ids leave in whole 32-byte sectors: the id with slab index g waits at word g & 15 of the lane's
  // shared-memory row (two sectors: a step emits up to two ids and may finish one sector and begin
  // the next) and a finished, aligned group of eight goes out with one STG.256 after the step (a
  // 4-byte store per id and lane was 20 of the ~100 L1 wavefronts of a step).  The ends of a string
  // that do not fill a sector go out one by one.
  constexpr uint32_t kRowBytes = 80;   // 16 words + 16: an odd number of 16-byte units
  __shared__ __align__(16) uint8_t s_rows[STORE ? kPipeBlock * kRowBytes : 16];
  const uint32_t row_b = pipe_smem_u32(s_rows) + (STORE ? uint32_t(threadIdx.x) * kRowBytes : 0u);
  auto row_word = [&](uint32_t g) -> uint32_t {
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(row_b + (g & 15u) * 4u) : "memory");
    return w;
  };

  for (;;) {
    uint32_t g = 0;
    if (lane == 0) g = atomicAdd(counter, 1u);
    g = __shfl_sync(FULL, g, 0);
    if (uint64_t(g) * 32u >= A.n) break;
    const uint32_t slot_i = g * 32u + uint32_t(lane);
    bool active = false;
    uint32_t s = 0, cap = 0;
    uint64_t slab_base = 0;
    const uint8_t* text = nullptr;
    ChainString S{nullptr, 0};
    ChainLane z;
    chain_reset(z);
    if (slot_i < A.n) {
      const uint32_t item = A.sort_order[slot_i];
      if (A.pflag[item] == 0) {
        int64_t o0 = 0;
        int slen = 0;
        pipe_fetch(A, item, &s, &o0, &slen);
        const uint64_t e = A.toff[item + 1];
        S.len = int(A.nlen[item]);
        S.words = A.words + e - S.len;
        if (A.text) text = A.text + e - S.len;
        if (A.list) {   // placement follows the string's own (raw) extent, like the first pass
          const int64_t r0 = A.raw_off[s];
          const int64_t r1 = A.raw_ends ? A.raw_ends[s] : A.raw_off[s + 1];
          slab_base = A.slab_off ? A.slab_off[s] : uint64_t(r0 - A.raw_off[0]) + uint64_t(A.slab_pad) * s;
          cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(r1 - r0) + A.slab_pad;
        } else {
          slab_base = A.slab_off ? A.slab_off[s] : uint64_t(o0 - A.off[0]) + uint64_t(A.slab_pad) * s;
          cap = A.slab_off ? uint32_t(A.slab_off[s + 1] - A.slab_off[s]) : uint32_t(slen) + A.slab_pad;
        }
        active = true;
      }
    }
    // position of the string's first id inside its 32-byte sector (the slab pointer of a chunk launch
    // is itself only 4-byte aligned, so this comes from the address, not from the index)
    const uint32_t base_lo = uint32_t(reinterpret_cast<uintptr_t>(A.slab + slab_base) >> 2) & 7u;
    auto put = [&](uint32_t id, uint32_t& stored) {
      if (STORE && stored < cap)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(row_b + ((base_lo + stored) & 15u) * 4u), "r"(id) : "memory");
      stored++;
    };
    while (__any_sync(FULL, active)) {
      if (!active) continue;
      if (z.i >= S.len) {
        if (STORE) {   // the ids behind the last complete group of eight
          const uint32_t cntc = min(z.cnt, cap), rem = (base_lo + cntc) & 7u;
          for (uint32_t k = cntc > rem ? cntc - rem : 0u; k < cntc; k++) A.slab[slab_base + k] = row_word(base_lo + k);
        }
        A.counts[s] = STORE ? min(z.cnt, cap) : z.cnt;
        A.need[s] = z.cnt;
        A.missing[s] = z.miss;
        A.status[s] = 0;
        if (STORE && z.cnt > cap) atomicExch(A.overflow, 1u);
        active = false;
        continue;
      }
      uint32_t stored = z.cnt;   // ids stored so far follow the lane's own count
      const uint32_t g_before = base_lo + min(stored, cap);
      chain_step(
          T, S, z, two_ids, [&](const uint4* tab, uint32_t idx, uint4* a, uint4* b) { pipe_load32(tab + 2 * idx, a, b); },
          [&](uint32_t id, uint32_t, uint32_t adv) {
            if (!STORE && A.scores) atomicAdd(&A.scores[id], static_cast<unsigned long long>(adv));
            put(id, stored);
          },
          [&](int pos) -> uint32_t {
            if (!STORE && A.scores) {   // trainvocab.go:1164-1171: counted, never scored
              atomicAdd(&A.miss_hist[text[pos]], 1ull);
              return 1u;
            }
            if (A.v.unk != 0xFFFFFFu) {
              put(A.v.unk, stored);
              return 1u;
            }
            return 0u;
          });
      if (STORE) {
        // did this step finish a group of eight?  (at most one: a step stages at most two ids)
        const uint32_t g_after = base_lo + min(stored, cap);
        const uint32_t m = g_after >> 3;                       // groups 0 .. m-1 are complete now
        const bool done = m != (g_before >> 3);
        const bool whole = done & ((m > 1u) | (base_lo == 0u));   // group m-1 starts inside this string's ids
        asm volatile(
            "{\n\t.reg .pred q;\n\t.reg .b32 t0, t1, t2, t3, t4, t5, t6, t7;\n\t"
            "setp.ne.u32 q, %1, 0;\n\t"
            "@q ld.shared.v4.u32 {t0, t1, t2, t3}, [%2];\n\t"
            "@q ld.shared.v4.u32 {t4, t5, t6, t7}, [%2+16];\n\t"
            "@q st.global.v8.u32 [%0], {t0, t1, t2, t3, t4, t5, t6, t7};\n\t}" ::"l"(A.slab + (slab_base - base_lo + 8u * (m - 1u))),
            "r"(uint32_t(whole)), "r"(row_b + ((m - 1u) & 1u) * 32u)
            : "memory");
        if (done && !whole)   // the string's first ids end a sector they do not fill (once per string)
          for (uint32_t k = 0; k < 8u - base_lo; k++) A.slab[slab_base + k] = row_word(base_lo + k);
      }
    }
  }
}

}  // namespace tmgpu

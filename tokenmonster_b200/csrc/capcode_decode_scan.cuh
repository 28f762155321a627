// capcode_decode_scan.cuh -- the capcode level-2 decoder (javascript/tokenmonster.js:1007-1065;
// stateless use go/tokenmonster.go:788-796, streaming use go:701-740) for ASCII text as a SCAN.
//
// The decoder is a 16-state transducer (inWord, inChar, delete, ignore) and the state-to-state
// functions its bytes generate form a monoid of 44 elements (gen_capdec_tables.py), so the effect
// of any run of bytes on the state is a one-byte id and two runs compose with one table read.
// One warp per string, 512 bytes per round:
//   1. every lane loads 16 bytes (one 128-bit load, the warp reads 512 consecutive bytes);
//   2. composes them into a monoid id (16 shared-memory reads);
//   3. a warp scan of the ids (5 shuffles + table reads) gives every lane the state its bytes start in;
//   4. the lane runs its 16 bytes from that state: emitted bytes (capitals restored) go to a
//      lane-private strip of shared memory, then -- after a warp prefix sum of the counts -- into
//      the warp's window, which the warp writes out with consecutive lanes on consecutive words.
// The lane-per-string kernel this replaces spent ~100 instructions per byte; this one spends ~30
// per byte-lane with all 32 lanes busy and coalesced traffic in both directions.
// Strings holding a byte >= 0x80 are left to capcode_decode_kernel (out_len = 0xFFFFFFFF), as before.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace tmgpu {

#define TM_CAPDEC_QUAL __device__ __align__(16) const
#include "capdec_tables.inc"
#undef TM_CAPDEC_QUAL

// The byte-indexed tables sit in shared memory with rows of 260 bytes instead of 256: text is mostly
// lower-case letters, whose 26 codes fall into 7 of the 32 banks, and with 256-byte rows every row
// starts in bank 0 -- lanes in different rows (states) then collide on those few banks.  The skew
// moves row r by r banks.
constexpr int kCapdecRow = 260;
constexpr int kCapdecCompBytes = TM_CAPDEC_ELEMS * kCapdecRow;
constexpr int kCapdecMulBytes = TM_CAPDEC_ELEMS * 64;
constexpr int kCapdecApplyBytes = TM_CAPDEC_ELEMS * 16;
constexpr int kCapdecStepBytes = 16 * kCapdecRow;
constexpr int kCapdecPrivStride = 20;   // bytes per lane in the private strip (5 words: conflict-free across lanes)

__device__ __forceinline__ void capdec_copy_table(uint8_t* dst, const uint8_t* src, int bytes) {
  for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}
// rows of 256 bytes -> rows of kCapdecRow bytes, word by word
__device__ __forceinline__ void capdec_copy_rows(uint8_t* dst, const uint8_t* src, int rows) {
  for (int i = threadIdx.x; i < rows * 64; i += blockDim.x)
    *reinterpret_cast<uint32_t*>(dst + (i >> 6) * kCapdecRow + (i & 63) * 4) = *reinterpret_cast<const uint32_t*>(src + i * 4);
}
// mask of the bytes [lo, hi) of a 16-byte block that fall into its 32-bit word wi (little endian)
__device__ __forceinline__ uint32_t capdec_word_mask(int lo, int hi, int wi) {
  int a = lo - 4 * wi, b = hi - 4 * wi;
  a = a < 0 ? 0 : a;
  b = b > 4 ? 4 : b;
  if (b <= a) return 0u;
  const uint32_t upto = b >= 4 ? 0xFFFFFFFFu : ((1u << (8 * b)) - 1u);
  return upto & (0xFFFFFFFFu << (8 * a));   // a < b <= 4
}

__device__ __forceinline__ uint32_t capdec_byte(const uint4& w, int j) {
  const uint32_t x = (j < 4) ? w.x : (j < 8) ? w.y : (j < 12) ? w.z : w.w;
  return (x >> (8 * (j & 3))) & 0xFFu;
}

__global__ void __launch_bounds__(256) capcode_decode_scan_kernel(const uint8_t* __restrict__ src,
                                                                  const uint64_t* __restrict__ raw_off, uint32_t n,
                                                                  uint8_t* __restrict__ dst,
                                                                  uint32_t* __restrict__ out_len,
                                                                  const uint32_t* __restrict__ len_in,
                                                                  uint8_t* __restrict__ state, int wide,
                                                                  uint32_t* __restrict__ counter) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  __shared__ __align__(16) uint8_t s_comp[kCapdecCompBytes];
  __shared__ __align__(16) uint8_t s_mul[kCapdecMulBytes];
  __shared__ __align__(16) uint8_t s_apply[kCapdecApplyBytes];
  __shared__ __align__(16) uint8_t s_step[kCapdecStepBytes];
  __shared__ __align__(16) uint8_t s_priv[8][32 * kCapdecPrivStride];
  __shared__ __align__(16) uint8_t s_win[8][512 + 16];
  capdec_copy_rows(s_comp, kCapdecComp, TM_CAPDEC_ELEMS);
  capdec_copy_table(s_mul, kCapdecMul, kCapdecMulBytes);
  capdec_copy_table(s_apply, kCapdecApply, kCapdecApplyBytes);
  capdec_copy_rows(s_step, kCapdecStep, 16);
  __syncthreads();
  // the text buffer is read 16 aligned bytes at a time wherever those 16 bytes lie inside it: before
  // a string's first byte and behind its last one they belong to its neighbours
  const uint64_t text_end = raw_off[n];
  const int lane = threadIdx.x & 31;
  uint8_t* priv = s_priv[threadIdx.x >> 5] + lane * kCapdecPrivStride;
  uint8_t* win = s_win[threadIdx.x >> 5];

  for (;;) {
    uint32_t s = 0;
    if (lane == 0) s = atomicAdd(counter, 1u);
    s = __shfl_sync(FULL, s, 0);
    if (s >= n) break;
    const uint64_t o0 = raw_off[s];
    const uint8_t* in = src + o0;
    const uint32_t len = len_in ? len_in[s] : uint32_t(raw_off[s + 1] - o0);
    const int galign = int(reinterpret_cast<uintptr_t>(in) & 15);
    const uint8_t* ga = in - galign;                 // 16-byte aligned; coordinates count from here
    const long long total_c = (long long)galign + (long long)len;
    uint8_t* out = dst + o0 + (wide ? ((o0 + 1) >> 1) : 0ull);
    uint32_t st = state ? uint32_t(state[s]) & 15u : 0u;   // streaming decoder: the flags persist (js:1009-1012)
    uint32_t o = 0;
    bool wide_text = false;

    for (long long base_c = 0; base_c < total_c; base_c += 512) {
      // ---- 1. this lane's 16 bytes; bytes outside the string read as 0xFF (= identity) ----------
      const long long c0 = base_c + lane * 16;
      uint4 w = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      bool hi = false;
      if (c0 >= galign && c0 + 16 <= total_c) {
        w = __ldg(reinterpret_cast<const uint4*>(ga + c0));
        hi = ((w.x | w.y | w.z | w.w) & 0x80808080u) != 0;
      } else if (c0 < total_c && o0 >= uint64_t(galign) && (o0 - uint64_t(galign)) + uint64_t(c0) + 16u <= text_end) {
        // the string's first / last block, its neighbours' bytes masked out (they read as 0xFF)
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(ga + c0));
        const int lo = galign > c0 ? int(galign - c0) : 0;
        const int hi_b = total_c - c0 < 16 ? int(total_c - c0) : 16;
        const uint32_t m0 = capdec_word_mask(lo, hi_b, 0), m1 = capdec_word_mask(lo, hi_b, 1);
        const uint32_t m2 = capdec_word_mask(lo, hi_b, 2), m3 = capdec_word_mask(lo, hi_b, 3);
        hi = (((v.x & m0) | (v.y & m1) | (v.z & m2) | (v.w & m3)) & 0x80808080u) != 0;
        w = make_uint4(v.x | ~m0, v.y | ~m1, v.z | ~m2, v.w | ~m3);
      } else if (c0 < total_c) {   // the last bytes of the whole buffer
        uint32_t t[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        for (int j = 0; j < 16; j++) {
          const long long c = c0 + j;
          if (c >= galign && c < total_c) {
            const uint32_t b = ga[c];
            hi |= (b & 0x80u) != 0;
            t[j >> 2] = (t[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (b << (8 * (j & 3)));
          }
        }
        w = make_uint4(t[0], t[1], t[2], t[3]);
      }
      if (__any_sync(FULL, hi)) {   // non-ASCII text: the per-code-point kernel takes the whole string
        wide_text = true;
        break;
      }
      // ---- 2. the effect of the 16 bytes on the state, as a monoid id ---------------------------
      uint32_t m = 0;   // identity
#pragma unroll
      for (int j = 0; j < 16; j++) m = s_comp[m * uint32_t(kCapdecRow) + capdec_byte(w, j)];
      // ---- 3. inclusive scan over the lanes: incl = (bytes of lanes 0..lane) -------------------
      uint32_t incl = m;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl = s_mul[t * 64u + incl];   // the lower lanes' run first, then ours
      }
      uint32_t excl = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) excl = 0;
      const uint32_t tile_m = __shfl_sync(FULL, incl, 31);
      // ---- 4. run the 16 bytes from the true start state ----------------------------------------
      uint32_t cur = s_apply[excl * 16u + st];
      uint32_t cnt = 0;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const uint32_t b = capdec_byte(w, j);
        const uint32_t e = s_step[cur * uint32_t(kCapdecRow) + b];
        cur = e & 15u;
        if (e & 16u) {
          priv[cnt] = uint8_t(b - (e & 32u));   // upper-case bit = 32 = 'a' - 'A'
          cnt++;
        }
      }
      uint32_t pre = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, pre, d);
        if (lane >= d) pre += t;
      }
      const uint32_t total = __shfl_sync(FULL, pre, 31);
      // the window starts at the output's own alignment, so that whole words can be stored
      const uint32_t a = uint32_t(reinterpret_cast<uintptr_t>(out + o) & 3u);
      uint8_t* wd = win + a + (pre - cnt);
      // Byte by byte into the window.  Lanes that emitted all 16 bytes sit 16 bytes apart, i.e. lanes
      // l, l+8, l+16, l+24 in the same banks: each group of eight starts at a different byte (0, 4,
      // 8, 12) and wraps, so that they write different banks at every step.
      uint32_t j = uint32_t(lane >> 3) * 4u;
      j = j < cnt ? j : 0u;
      for (uint32_t t = 0; t < cnt; t++) {
        wd[j] = priv[j];
        j = j + 1u < cnt ? j + 1u : 0u;
      }
      __syncwarp();
      uint8_t* gout = out + o - a;   // 4-byte aligned
      const uint32_t end = a + total;
      for (uint32_t q = uint32_t(lane) * 4u; q < end; q += 128u) {
        if (q >= a && q + 4u <= end) {
          *reinterpret_cast<uint32_t*>(gout + q) = *reinterpret_cast<const uint32_t*>(win + q);
        } else {
          for (uint32_t k = q; k < q + 4u; k++)
            if (k >= a && k < end) gout[k] = win[k];
        }
      }
      __syncwarp();
      o += total;
      st = s_apply[tile_m * 16u + st];
    }
    if (lane == 0) {
      if (wide_text) {
        out_len[s] = 0xFFFFFFFFu;
      } else {
        out_len[s] = o;
        if (state) state[s] = uint8_t(st);
      }
    }
  }
}

}  // namespace tmgpu

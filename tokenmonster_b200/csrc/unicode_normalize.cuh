// unicode_normalize.cuh -- the non-ASCII half of normalization, on the device.
//
// The tokenize kernels normalize ASCII text on the fly (their FILL phase) and give up on the first
// byte >= 0x80 (per-string status 13).  Those strings come here: one lane per string runs the whole
// reference pipeline (go/tokenmonster.go:242-253)
//     UTF-8 decode -> NFD (flag 1; UAX #15: full canonical decomposition + canonical ordering)
//                  -> drop Mn (flag 4, tokenmonster-cpp/src/tokenmonster.cpp:464-468)
//                  -> simple lower case (flag 2, cpp:469-472)
//                  -> capcode level 2 encode per code point (javascript/tokenmonster.js:900-1005,
//                     categories Lu / Ll / L / Nd / M, apostrophes ' and U+2019)
// and writes the normalized UTF-8 text to a scratch buffer; a second launch of the tokenize kernel
// then reads that text verbatim ("prenorm").  This is the slow path by design: it costs nothing
// for ASCII batches and keeps registers out of the hot kernel.
// Refused (status 13 stays): invalid UTF-8; more than 30 consecutive non-starters (x/text's NFD
// inserts U+034F there, ICU does not: there is no single reference answer).
// Character data: unicode_tables.inc, generated from Unicode 15.0 (gen_unicode_tables.py).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "unicode_dev.h"

namespace tmgpu {

// property word: bits 0-2 class, bit 3 decomposes, bit 4 has lower, bit 5 has upper, bits 8-15 ccc
enum : uint32_t { UC_O = 0, UC_LU = 1, UC_LL = 2, UC_LO = 3, UC_ND = 4, UC_MN = 5, UC_MO = 6, UC_APO = 7, UC_SP = 8 };

__device__ __forceinline__ uint32_t uni_prop(const UnicodeDev& U, uint32_t cp) {
  if (cp < 128u)   // ASCII needs no table: class + case bits, no decomposition, ccc 0
    return (cp - 'A' < 26u) ? (UC_LU | 16u) : (cp - 'a' < 26u) ? (UC_LL | 32u) : (cp - '0' < 10u) ? uint32_t(UC_ND) : 0u;
  if (cp >= 0x110000u) return 0;
  return U.stage2[uint32_t(U.stage1[cp >> 7]) * 128u + (cp & 127u)];
}

__device__ __forceinline__ int uni_find(const uint32_t* keys, uint32_t n, uint32_t cp) {
  int lo = 0, hi = int(n) - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const uint32_t k = keys[mid];
    if (k == cp) return mid;
    if (k < cp) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

__device__ __forceinline__ uint32_t uni_lower(const UnicodeDev& U, uint32_t cp, uint32_t prop) {
  if (cp < 128u) return (cp - 'A' < 26u) ? cp + 32u : cp;
  if (!(prop & 16u)) return cp;
  const int k = uni_find(U.lkey, U.n_lower, cp);
  return k < 0 ? cp : U.lval[k];
}

__device__ __forceinline__ uint32_t uni_upper(const UnicodeDev& U, uint32_t cp, uint32_t prop) {
  if (cp < 128u) return (cp - 'a' < 26u) ? cp - 32u : cp;
  if (!(prop & 32u)) return cp;
  const int k = uni_find(U.ukey, U.n_upper, cp);
  return k < 0 ? cp : U.uval[k];
}

// class used by capcode: the general-category class, with ' ' and the two apostrophes set apart
__device__ __forceinline__ uint32_t uni_class(uint32_t cp, uint32_t prop) {
  if (cp == ' ') return UC_SP;
  if (cp == '\'' || cp == 0x2019u) return UC_APO;
  return prop & 7u;
}

// strict UTF-8 decode (what Python's / Go's validating decoders accept).  Returns the sequence
// length, 0 if invalid.
__device__ __forceinline__ int utf8_decode(const uint8_t* p, int64_t left, uint32_t* cp) {
  const uint32_t b0 = p[0];
  if (b0 < 0x80u) { *cp = b0; return 1; }
  if (b0 < 0xC2u) return 0;
  if (b0 < 0xE0u) {
    if (left < 2 || (p[1] & 0xC0u) != 0x80u) return 0;
    *cp = ((b0 & 0x1Fu) << 6) | (p[1] & 0x3Fu);
    return 2;
  }
  if (b0 < 0xF0u) {
    if (left < 3 || (p[1] & 0xC0u) != 0x80u || (p[2] & 0xC0u) != 0x80u) return 0;
    const uint32_t c = ((b0 & 0x0Fu) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3Fu);
    if (c < 0x800u || (c >= 0xD800u && c <= 0xDFFFu)) return 0;
    *cp = c;
    return 3;
  }
  if (b0 < 0xF5u) {
    if (left < 4 || (p[1] & 0xC0u) != 0x80u || (p[2] & 0xC0u) != 0x80u || (p[3] & 0xC0u) != 0x80u) return 0;
    const uint32_t c = ((b0 & 0x07u) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3Fu);
    if (c < 0x10000u || c > 0x10FFFFu) return 0;
    *cp = c;
    return 4;
  }
  return 0;
}

__device__ __forceinline__ int utf8_len(uint32_t cp) { return cp < 0x80u ? 1 : cp < 0x800u ? 2 : cp < 0x10000u ? 3 : 4; }

__device__ __forceinline__ int utf8_put(uint8_t* o, uint32_t cp) {
  if (cp < 0x80u) { o[0] = uint8_t(cp); return 1; }
  if (cp < 0x800u) { o[0] = uint8_t(0xC0u | (cp >> 6)); o[1] = uint8_t(0x80u | (cp & 0x3Fu)); return 2; }
  if (cp < 0x10000u) {
    o[0] = uint8_t(0xE0u | (cp >> 12)); o[1] = uint8_t(0x80u | ((cp >> 6) & 0x3Fu)); o[2] = uint8_t(0x80u | (cp & 0x3Fu));
    return 3;
  }
  o[0] = uint8_t(0xF0u | (cp >> 18)); o[1] = uint8_t(0x80u | ((cp >> 12) & 0x3Fu));
  o[2] = uint8_t(0x80u | ((cp >> 6) & 0x3Fu)); o[3] = uint8_t(0x80u | (cp & 0x3Fu));
  return 4;
}

// ---- capcode level-2 encoder over code points, writing UTF-8 (js:900-1005) -------------------
struct CapcodeWriter {
  uint8_t* o;            // output of this string
  int64_t pos = 0;
  int64_t word_tok = 0, goback = 0;
  uint32_t cl1 = UC_O, cl2 = UC_O;   // classes of the previous two code points ('.' initially)
  bool in_word = false, multi = false;

  __device__ __forceinline__ void put(uint32_t cp) { pos += utf8_put(o + pos, cp); }
  __device__ __forceinline__ void dspace() { o[pos++] = 'D'; o[pos++] = ' '; }

  // js:924-951: the word turned out mixed case -- every lower-case code point written since
  // gobackPos gets a "DC " in front (its "D " prefix, if any, becomes "DC ").  Done in place,
  // back to front; the segment holds only lower-cased letters, digits, apostrophes, marks and
  // "D " pairs, so a space can only be the second half of such a pair.
  __device__ void respell(const UnicodeDev& U) {
    int64_t extra = 0;
    for (int phase = 0; phase < 2; phase++) {     // 0: count the growth, 1: move
      int64_t r = pos, w = pos + extra;
      while (r > goback) {
        int64_t a = r - 1;
        while (a > goback && (o[a] & 0xC0u) == 0x80u) a--;
        uint32_t cp = 0;
        utf8_decode(o + a, r - a, &cp);
        const bool lower = (uni_prop(U, cp) & 7u) == UC_LL;
        const bool after_d = lower && a - 2 >= goback && o[a - 2] == 'D' && o[a - 1] == ' ';
        if (phase == 1) {
          for (int64_t k = r - 1; k >= a; k--) o[--w] = o[k];
          if (lower) { o[--w] = ' '; o[--w] = 'C'; o[--w] = 'D'; }
        } else if (lower) {
          extra += after_d ? 1 : 3;
        }
        r = after_d ? a - 2 : a;
      }
      if (phase == 0 && extra == 0) return;
    }
    pos += extra;
  }

  __device__ void feed(const UnicodeDev& U, uint32_t cp, uint32_t prop) {
    const uint32_t cl = uni_class(cp, prop);
    const bool letter1 = cl1 == UC_LU || cl1 == UC_LL || cl1 == UC_LO;
    const bool mod1 = cl1 == UC_MN || cl1 == UC_MO;
    const bool glue1 = letter1 || cl1 == UC_APO || mod1;
    if (in_word) {
      if (cl == UC_LU) {
        if (!glue1) dspace();
        multi = true;
        put(uni_lower(U, cp, prop));
      } else {
        if (cl == UC_LL) {
          in_word = false;
          o[word_tok] = 'C';
          if (multi) respell(U);
          if (!glue1) dspace();
        } else if (cl == UC_ND) {
          if (cl1 != UC_ND) dspace();
        } else if (!(cl == UC_APO || cl == UC_MN || cl == UC_MO)) {
          in_word = false;
        }
        put(cp);
      }
    } else if (cl == UC_LL) {
      const bool letter2 = cl2 == UC_LU || cl2 == UC_LL || cl2 == UC_LO;
      if (!(cl1 == UC_SP || letter1 || (letter2 && cl1 == UC_APO) || mod1)) dspace();
      put(cp);
    } else if (cl == UC_LU) {
      if (cl1 == UC_SP) {
        word_tok = pos - 1;
        o[word_tok] = 'W';
        o[pos++] = ' ';
      } else {
        o[pos++] = 'D';
        word_tok = pos;
        o[pos++] = 'W';
        o[pos++] = ' ';
      }
      put(uni_lower(U, cp, prop));
      goback = pos;
      multi = false;
      in_word = true;
    } else if (cl == UC_ND) {
      if (!(cl1 == UC_SP || cl1 == UC_ND)) dspace();
      put(cp);
    } else {
      put(cp);
    }
    cl2 = cl1;
    cl1 = cl;
  }
};

constexpr int kMaxNonStarters = 30;

__global__ void __launch_bounds__(128) unicode_normalize_kernel(const NormalizeArgs A) {
  const UnicodeDev& U = A.U;
  const uint32_t n = *A.n_list;
  const bool do_nfd = (A.normflag & 5u) != 0;      // accents implies NFD (cpp:464)
  const bool strip = (A.normflag & 4u) != 0;
  const bool lower = (A.normflag & 2u) != 0;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const uint32_t s = A.list[k];
    const int64_t b0 = A.off[s], b1 = A.ends ? A.ends[s] : A.off[s + 1];
    const uint8_t* in = A.data + b0;
    const int64_t len = b1 - b0;
    CapcodeWriter W;
    W.o = A.scratch + A.n_begin[k];
    uint32_t run_cp[kMaxNonStarters];
    uint8_t run_cc[kMaxNonStarters] = {};
    int nrun = 0;
    bool refused = false;
    auto out_cp = [&](uint32_t cp, uint32_t prop) {   // after NFD: accents, lowercase, capcode
      if (strip && (prop & 7u) == UC_MN) return;
      if (lower) {
        const uint32_t l = uni_lower(U, cp, prop);
        if (l != cp) { cp = l; prop = uni_prop(U, cp); }
      }
      if (A.capcode == 2) W.feed(U, cp, prop);
      else W.put(cp);
    };
    auto flush_run = [&]() {
      for (int t = 0; t < nrun; t++) out_cp(run_cp[t], uni_prop(U, run_cp[t]));
      nrun = 0;
    };
    auto nfd_cp = [&](uint32_t cp) {                // one fully decomposed code point
      const uint32_t prop = uni_prop(U, cp);
      const uint32_t cc = prop >> 8;
      if (cc == 0) {
        flush_run();
        out_cp(cp, prop);
      } else {
        if (nrun >= kMaxNonStarters) { refused = true; return; }
        int t = nrun;                               // stable insertion by combining class
        while (t > 0 && run_cc[t - 1] > cc) { run_cp[t] = run_cp[t - 1]; run_cc[t] = run_cc[t - 1]; t--; }
        run_cp[t] = cp;
        run_cc[t] = uint8_t(cc);
        nrun++;
      }
    };
    int64_t p = 0;
    while (p < len && !refused) {
      uint32_t cp = 0;
      const int l = utf8_decode(in + p, len - p, &cp);
      if (l == 0) { refused = true; break; }
      p += l;
      if (!do_nfd) {
        out_cp(cp, uni_prop(U, cp));
      } else if (cp - 0xAC00u < 11172u) {           // Hangul syllable: algorithmic (UAX #15)
        const uint32_t si = cp - 0xAC00u;
        nfd_cp(0x1100u + si / 588u);
        nfd_cp(0x1161u + (si % 588u) / 28u);
        if (si % 28u) nfd_cp(0x11A7u + si % 28u);
      } else {
        const uint32_t prop = uni_prop(U, cp);
        if (prop & 8u) {
          const int d = uni_find(U.dkey, U.n_decomp, cp);
          for (uint32_t t = U.didx[d]; t < U.didx[d + 1]; t++) nfd_cp(U.ddata[t]);
        } else {
          nfd_cp(cp);
        }
      }
    }
    if (!refused) flush_run();
    A.n_end[k] = refused ? A.n_begin[k] - 1 : A.n_begin[k] + W.pos;
  }
}

// strings of status 13 -> list + scratch placement (capacity 4 * len + 16, 16-byte aligned)
__global__ void __launch_bounds__(256) select_refused_kernel(const uint8_t* __restrict__ status,
                                                             const int64_t* __restrict__ off,
                                                             const int64_t* __restrict__ ends, uint32_t n,
                                                             uint32_t* __restrict__ list, int64_t* __restrict__ n_begin,
                                                             unsigned long long* __restrict__ ctr) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (status[s] != 13) continue;
    const unsigned long long len = static_cast<unsigned long long>((ends ? ends[s] : off[s + 1]) - off[s]);
    const unsigned long long cap = (4ull * len + 16ull + 15ull) & ~15ull;
    const unsigned long long k = atomicAdd(&ctr[0], 1ull);
    const unsigned long long b = atomicAdd(&ctr[1], cap);
    list[k] = s;
    n_begin[k] = static_cast<int64_t>(b);
  }
}

}  // namespace tmgpu

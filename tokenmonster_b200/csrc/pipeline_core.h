// pipeline_core.h -- the per-string logic of the position-table pipeline, written once as
// host+device inline functions: the sm_100a kernels (pipeline_kernels.cuh) run it one lane per
// string, and tests/host_sim.cpp compiles the same functions with g++ to check them against the
// oracle on machines without a GPU (test infrastructure; the product has no CPU path).
//
// The pipeline replaces the trie-walk kernel of round 1.  Three stages per string:
//   1 ENCODE  capcode level 2 (javascript/tokenmonster.js:900-1005, the only in-repo restatement
//             of capcode.Encode used by go/tokenmonster.go:233-240), left to right, into HBM text
//   2 TABLE   the reversed-key automaton (vocab_tables.h) runs right to left over the normalized
//             text and leaves ONE 32-bit word per position:
//               state(20) | best_len(6) << 20 | nb(2) << 26 | cls(2) << 28 | nbb(2) << 30
//             best_len = LongestSubstring at that position (go:1049,1068,1113,1165), nb = class
//             of the byte behind that match (beginByte, go:1073), cls = class of the byte AT the
//             position (go:1057), nbb = class of the byte behind the space-prefixed match
//             (go:1094) -- everything the scoring needs about the text itself
//   3 CHAIN   the ungreedy loop (go:1017-1279): per step three word loads, three state records,
//             one token record, the scores -- no search of any kind
#pragma once
#include <cstdint>
#include <vector_types.h>

#include "device_tables.h"

namespace tmgpu {

// ------------------------------------------------------------------------------------------
// stage 1: capcode level 2 encoder
// ------------------------------------------------------------------------------------------
enum : uint32_t { PC_O = 0, PC_U = 1, PC_L = 2, PC_N = 3, PC_A = 4 };   // other, upper, lower, digit, apostrophe

TM_HD uint32_t pc_class(uint32_t c) {
  if (c - 'a' < 26u) return PC_L;
  if (c - 'A' < 26u) return PC_U;
  if (c - '0' < 10u) return PC_N;
  if (c == '\'') return PC_A;
  return PC_O;
}

// ------------------------------------------------------------------------------------------
// stage 1, right to left: the encoder (a forward-looking restatement of js:900-1005: the marker
// of a capitalised run -- W = stays capital, C = mixed case -- is decided where the run starts, so
// nothing already emitted is ever rewritten) driven from the END of the string, one raw character
// per call.  This is what the table kernel runs (the automaton consumes the normalized text back
// to front, so producing it back to front lets both live in one loop and the normalized text never
// exists in memory).  Everything the encoder knows about a character is local:
//   * the two characters to its LEFT are read directly (c1, c2);
//   * "inside a capitalised run" = the nearest character to the left that is not a digit or an
//     apostrophe is a capital: it is the left neighbour itself unless that is a digit / apostrophe,
//     in which case one look-back over the run answers for every character of the run;
//   * "the run is mixed case" = the nearest lower-case-or-other character to the RIGHT is
//     lower-case, which a right-to-left pass carries along (nlo);
//   * the class of the character to the right (clr) decides whether a space is replaced by a marker.
// ------------------------------------------------------------------------------------------
struct RevCap {
  uint32_t clr;          // class of the character to the right
  uint32_t nlo;          // class of the nearest lower-case-or-other character to the right (other if none)
  bool has_right;        // there is a character to the right
  bool iw_valid, iw;     // cached answer for the current run of digits / apostrophes
};

TM_HD void revcap_reset(RevCap& e) {
  e.clr = PC_O;
  e.nlo = PC_O;
  e.has_right = false;
  e.iw_valid = false;
  e.iw = false;
}

// The encoder's decision is a function of ten bits, so the kernels read it from a table:
//   index = cl | cl1 << 3 | (cl2 is a letter) << 6 | in_word << 7 | mixed << 8 | (a capital follows) << 9
//   (classes as PC_*, with the space as a class of its own, PC_S)
//   entry = bytes in front of the character (<= 3, little endian) | count << 24 | eaten << 26 | lower << 27
constexpr uint32_t PC_S = 5;       // ' ' -- behaves as PC_O except where the encoder asks for a space
constexpr int kCapLutSize = 1024;

TM_HD uint32_t pc_class_sp(uint32_t c) { return c == ' ' ? PC_S : pc_class(c); }

TM_HD uint32_t revcap_lut_entry(uint32_t idx) {
  const uint32_t cl = idx & 7u, cl1 = (idx >> 3) & 7u;
  const bool cl2_letter = (idx >> 6) & 1u, in_word = (idx >> 7) & 1u, mixed = (idx >> 8) & 1u, cap_follows = (idx >> 9) & 1u;
  const bool sp1 = (cl1 == PC_S);
  const bool letter1 = (cl1 == PC_U) || (cl1 == PC_L);
  const bool glue1 = letter1 || (cl1 == PC_A);
  const bool start = (cl == PC_U) && !in_word;
  const bool joined_l = in_word ? glue1 : (sp1 || letter1 || (cl2_letter && cl1 == PC_A));
  const bool joined_n = in_word ? (cl1 == PC_N) : (sp1 || cl1 == PC_N);
  const bool k_dspace = (cl == PC_L && !joined_l) || (cl == PC_N && !joined_n) ||
                        (cl == PC_U && in_word && !mixed && !glue1);
  const bool k_dcs = (cl == PC_U) && in_word && mixed;
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
  const bool eaten = (cl == PC_S) && cap_follows;
  return pre | (np << 24) | (uint32_t(eaten) << 26) | (uint32_t(cl == PC_U) << 27);
}

// Encodes the raw character at index k.  x0 / x1 / x2 = character | class << 8 (pc_class_sp) of the
// characters at k / k-1 / k-2 (x1, x2 ignored where they do not exist); class_at(j) = class of
// character j (rare deeper look-back); lut = the kCapLutSize entries of revcap_lut_entry().
// Returns the number of bytes (0..4) placed in *out in FORWARD order, little endian -- the caller
// feeds them to the automaton last byte first (checked against the oracle's encoder,
// tests/test_host_sim.py).
template <class ClassAt>
TM_HD uint32_t revcap_step(RevCap& e, const uint32_t* lut, uint32_t x0, uint32_t x1, uint32_t x2, int k, ClassAt class_at,
                           uint32_t* out) {
  const uint32_t cl = x0 >> 8;
  const uint32_t cl1 = (k >= 1) ? (x1 >> 8) : uint32_t(PC_O);
  const uint32_t cl2 = (k >= 2) ? (x2 >> 8) : uint32_t(PC_O);
  // inside a capitalised run?  Decided by the left neighbour unless that is a digit / apostrophe;
  // then one look-back (rare) answers for the whole run of digits / apostrophes
  const bool run1 = (cl1 == PC_N) | (cl1 == PC_A);
  if (run1 & !e.iw_valid) {
    uint32_t x = cl2;
    int j = k - 2;
    while ((x == PC_N || x == PC_A) && j > 0) {
      j--;
      x = class_at(j);
    }
    e.iw = (k >= 2) && (x == PC_U);
  }
  e.iw_valid = run1;
  const bool in_word = run1 ? e.iw : (cl1 == PC_U);
  const uint32_t idx = cl | (cl1 << 3) | (uint32_t(cl2 == PC_U || cl2 == PC_L) << 6) | (uint32_t(in_word) << 7) |
                       (uint32_t(e.nlo == PC_L) << 8) | (uint32_t(e.has_right && e.clr == PC_U) << 9);
  const uint32_t a = lut[idx];
  const uint32_t np = (a >> 24) & 3u, eaten = (a >> 26) & 1u;
  const uint32_t ch = (x0 & 0xFFu) | (((a >> 27) & 1u) << 5);
  *out = (a & 0xFFFFFFu) | (eaten ? 0u : (ch << (8u * np)));
  e.clr = cl;
  e.has_right = true;
  e.nlo = (cl == PC_L) ? uint32_t(PC_L) : ((cl == PC_O) | (cl == PC_S) ? uint32_t(PC_O) : e.nlo);
  return np + 1u - eaten;
}

// ------------------------------------------------------------------------------------------
// stage 2: the position word
// ------------------------------------------------------------------------------------------
TM_HD uint32_t pw_state(uint32_t w) { return w & kDfaStateMask; }
TM_HD uint32_t pw_len(uint32_t w) { return (w >> 20) & 63u; }
TM_HD uint32_t pw_nb(uint32_t w) { return (w >> 26) & 3u; }
TM_HD uint32_t pw_cls(uint32_t w) { return (w >> 28) & 3u; }
TM_HD uint32_t pw_nbb(uint32_t w) { return w >> 30; }
// beginByte value of a 2-bit class code (vocab_tables.h: class_code)
TM_HD int class_value(uint32_t code) { return int((0x0C0A0100u >> (8 * code)) & 0xFFu); }

// Classes of the bytes to the right of the current position, 2 bits each, nearest first; bytes
// behind the end of the string are the pad byte 0 (defined as 0).  48 positions
// are kept (a key has at most 40 bytes).
struct ClassTrail {
  unsigned long long lo;   // positions +1..+32
  uint32_t hi;             // positions +33..+48
};
TM_HD void trail_reset(ClassTrail& t, uint32_t code0) {
  t.lo = code0 * 0x5555555555555555ull;
  t.hi = code0 * 0x55555555u;
}
TM_HD void trail_push(ClassTrail& t, uint32_t code) {
  t.hi = (t.hi << 2) | uint32_t(t.lo >> 62);
  t.lo = (t.lo << 2) | code;
}
// class of the byte d positions to the right (d >= 1)
TM_HD uint32_t trail_at(const ClassTrail& t, uint32_t d) {
  const uint32_t k = d - 1u;
  return (k < 32u ? uint32_t(t.lo >> (2u * k)) : (t.hi >> (2u * (k - 32u)))) & 3u;
}

// The word of the position whose byte (class code `code`) was just consumed into `packed`.
// `off` = bytes of the virtual space (1, or 2 for UTF-16).
TM_HD uint32_t position_word(uint32_t packed, uint32_t code, const ClassTrail& t, uint32_t off) {
  const uint32_t l = dfa_len(packed), sl = dfa_splen(packed);
  const uint32_t nb = l ? trail_at(t, l) : 0u;
  const uint32_t nbb = (sl > off) ? trail_at(t, sl - off) : 0u;
  return (packed & (kDfaStateMask | (63u << 20))) | (nb << 26) | (code << 28) | (nbb << 30);
}

// ------------------------------------------------------------------------------------------
// stage 3: one step of the ungreedy loop, branch free.
// The score of "first token a, then second token b" (go:1075-1084 plain, :1096-1105 retry;
// pair_score() in device_tables.h is the direct restatement) is split at table-build time into
//   A(a, fd)   = (a.flag>>7) + max0(nw-1) + 100*nw,  nw = a.nWords - fd          (token record)
//   B(b, nb)   = (b.flag>>7) + max0(b.nWords-1) + ((nb>>2)&1) + 100*(b.nWords + (nb>>3))
//                - 3*(b.flag & 1 & nb) [+ ((b.flag>>2)&1) plain | - 1 retry]      (state record,
//                one value per class nb of the following byte)
//   and the two interaction bits: -100*(a.flag>>3 & b.flag>>4 & 1), -103*(a.flag & 1 [& b.flag>>1])
// so a candidate costs one add chain instead of ~30 instructions, and all six candidates are
// evaluated unconditionally (invalid ones score -1000000): 32 lanes execute the same instructions.
// ------------------------------------------------------------------------------------------
struct ChainTables {
  const uint4* crec;     // 2 x uint4 per dictionary index (vocab_tables.h: chain record)
  const uint4* hot;      // 2 x uint4 per automaton state  (vocab_tables.h: state record)
  uint32_t delete_token, unk;
  uint32_t off;          // bytes of the virtual space
  bool can_retry;        // a delete token exists (go:1088)
};
struct ChainString {
  const uint32_t* words; // position words of this string
  int len;               // normalized length
};
struct ChainLane {
  int i, fd, tok_len;
  uint32_t idx;
  bool have;
  uint32_t cnt, miss;
};

TM_HD void chain_reset(ChainLane& z) {
  z.i = 0; z.fd = 0; z.tok_len = 0; z.idx = 0; z.have = false; z.cnt = 0; z.miss = 0;
}

TM_HD int i16_lo(uint32_t x) { return int(int16_t(x & 0xFFFFu)); }
TM_HD int i16_hi(uint32_t x) { return int(int16_t(x >> 16)); }
// value number `code` (0..3) of four int16 packed in two words
TM_HD int i16_sel(uint32_t w01, uint32_t w23, uint32_t code) {
  const uint32_t w = (code & 2u) ? w23 : w01;
  return int(int16_t((w >> (16u * (code & 1u))) & 0xFFFFu));
}

// emit(id, pos_code, adv): pos_code = position | forward-delete << 31 of the step (0xFFFFFFFF for
// the delete token of a two-id step), adv = characters the step advances by.
// on_miss(position) -> ids added: a byte no token covers (go:1269-1276); the caller emits unk.
// two_ids: false for Vocab.Count, which adds 1 where tokenize emits 2 ids.
// load32(table, index, &a, &b): one 32-byte record (a 256-bit load on the device).
template <class Load32, class Emit, class Miss>
TM_HD void chain_step(const ChainTables& T, const ChainString& S, ChainLane& z, bool two_ids, Load32 load32, Emit emit,
                      Miss on_miss) {
  if (!z.have) {   // loop top: the longest token at i (go:1049)
    const uint32_t w = S.words[z.i];
    const uint32_t l = pw_len(w);
    if (l == 0) {
      z.cnt += on_miss(z.i);   // ids the caller emitted for it (unk) / counted
      z.i++;
      z.miss++;
      z.fd = 0;
      return;
    }
    uint4 ha, hb;
    load32(T.hot, pw_state(w), &ha, &hb);
    z.idx = ha.x & 0xFFFFFFu;
    z.tok_len = int(l);
    z.have = true;
  }
  uint4 ra, rb;
  load32(T.crec, z.idx, &ra, &rb);
  const uint32_t bits = ra.x >> 24;
  const int alen1 = int(ra.y >> 24), alen2 = int(ra.z >> 24);
  const int i = z.i, fd = z.fd, len = z.tok_len, L = S.len;
  const int i1 = i + len;
  const int base1 = alen1 - fd, base2 = alen2 - fd;
  // the three look-ahead positions (go:1068, 1113, 1165); alternative 2 only with alternative 1 (Q10)
  const bool in0 = i1 < L;
  const uint32_t w0 = in0 ? S.words[i1] : 0u;
  const bool look = in0 & (((bits & 4u) == 0) | (pw_cls(w0) != 3u));   // go:1057 (class 3 = 12 = space)
  const bool has1 = look & (alen1 != 0), has2 = has1 & (alen2 != 0);
  const uint32_t w1 = (has1 & (i + base1 < L)) ? S.words[i + base1] : 0u;
  const uint32_t w2 = (has2 & (i + base2 < L)) ? S.words[i + base2] : 0u;
  const int l0 = int(pw_len(w0)), l1 = int(pw_len(w1)), l2 = int(pw_len(w2));
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 h0a = zero, h0b = zero, h1a = zero, h1b = zero, h2a = zero, h2b = zero;
  if (l0) load32(T.hot, pw_state(w0), &h0a, &h0b);
  if (l1) load32(T.hot, pw_state(w1), &h1a, &h1b);
  if (l2) load32(T.hot, pw_state(w2), &h2a, &h2b);
  // first-token parts: the current token, its alternative 1, its alternative 2
  const int A0 = fd ? i16_hi(ra.w) : i16_lo(ra.w);
  const int A1 = fd ? i16_hi(rb.x) : i16_lo(rb.x);
  const int A2 = fd ? i16_hi(rb.y) : i16_lo(rb.y);
  const int a1_0 = int(bits & 1u), a8_0 = int((bits >> 1) & 1u);
  const int a1_1 = int((bits >> 3) & 1u), a8_1 = int((bits >> 4) & 1u);
  const int a1_2 = int((bits >> 5) & 1u), a8_2 = int((bits >> 6) & 1u);
  const bool v0 = look & (l0 > 0), v1 = has1 & (l1 > 0), v2 = has2 & (l2 > 0);
  const int bl1 = base1 + l1, bl2 = base2 + l2;
  // plain candidates s1, s2, s3
  const int s0 = v0 ? (len + l0) + A0 + i16_sel(h0b.x, h0b.y, pw_nb(w0)) - 100 * (a8_0 & int((h0a.z >> 1) & 1u)) -
                          103 * (a1_0 & int(h0a.z & 1u))
                    : kNegScore;
  const int s1 = v1 ? bl1 + A1 + i16_sel(h1b.x, h1b.y, pw_nb(w1)) - 100 * (a8_1 & int((h1a.z >> 1) & 1u)) -
                          103 * (a1_1 & int(h1a.z & 1u)) - 100 * int(bl1 < len) - 10000 * int(bl1 == len)
                    : kNegScore;
  const int s2 = v2 ? bl2 + A2 + i16_sel(h2b.x, h2b.y, pw_nb(w2)) - 100 * (a8_2 & int((h2a.z >> 1) & 1u)) -
                          103 * (a1_2 & int(h2a.z & 1u)) - 100 * int(bl2 < len) - 10000 * int(bl2 == len)
                    : kNegScore;
  // retries s1b, s2b, s3b (go:1088-1107): wanted when the second token begins with a letter, has
  // no word boundary inside and a letter follows it; accepted when the space-prefixed key is more
  // than one byte longer than the plain one
  const int off = int(T.off);
  const int sl0 = int(h0a.y >> 24), sl1 = int(h1a.y >> 24), sl2 = int(h2a.y >> 24);
  const bool r0 = T.can_retry & v0 & (((h0a.z >> 2) & 1u) != 0) & (pw_nb(w0) == 1u) & (sl0 > l0 + 1);
  const bool r1 = T.can_retry & v1 & (((h1a.z >> 2) & 1u) != 0) & (pw_nb(w1) == 1u) & (sl1 > l1 + 1);
  const bool r2 = T.can_retry & v2 & (((h2a.z >> 2) & 1u) != 0) & (pw_nb(w2) == 1u) & (sl2 > l2 + 1);
  const int bb1 = base1 + sl1 - off, bb2 = base2 + sl2 - off;
  const int s3 = r0 ? (len + sl0 - off) + A0 + i16_sel(h0b.z, h0b.w, pw_nbb(w0)) - 100 * (a8_0 & int((h0a.z >> 3) & 1u)) - 103 * a1_0
                    : kNegScore;
  const int s4 = r1 ? bb1 + A1 + i16_sel(h1b.z, h1b.w, pw_nbb(w1)) - 100 * (a8_1 & int((h1a.z >> 3) & 1u)) - 103 * a1_1 -
                          100 * int(bb1 < len) - 10000 * int(bb1 == len)
                    : kNegScore;
  const int s5 = r2 ? bb2 + A2 + i16_sel(h2b.z, h2b.w, pw_nbb(w2)) - 100 * (a8_2 & int((h2a.z >> 3) & 1u)) - 103 * a1_2 -
                          100 * int(bb2 < len) - 10000 * int(bb2 == len)
                    : kNegScore;
  // first maximum in the order s1, s2, s3, s1b, s2b, s3b (go:1217-1262)
  int best_s = kNegScore, pick = -1;
  if (s0 > best_s) { best_s = s0; pick = 0; }
  if (s1 > best_s) { best_s = s1; pick = 1; }
  if (s2 > best_s) { best_s = s2; pick = 2; }
  if (s3 > best_s) { best_s = s3; pick = 3; }
  if (s4 > best_s) { best_s = s4; pick = 4; }
  if (s5 > best_s) { best_s = s5; pick = 5; }
  const uint32_t pos_code = uint32_t(i) | (uint32_t(fd) << 31);
  const int br = pick >= 3 ? pick - 3 : (pick < 0 ? 0 : pick);
  const bool retry = pick >= 3;
  const uint32_t id = br == 0 ? ra.x : (br == 1 ? ra.y : ra.z);
  const int adv = br == 0 ? len : (br == 1 ? base1 : base2);   // no branch taken: the token itself (go:1265-1267)
  emit(id & 0xFFFFFFu, pos_code, uint32_t(adv));
  z.cnt++;
  if (retry & two_ids) {
    emit(T.delete_token, 0xFFFFFFFFu, 1u);
    z.cnt++;
  }
  const uint32_t hx = br == 0 ? h0a.x : (br == 1 ? h1a.x : h2a.x);
  const uint32_t hy = br == 0 ? h0a.y : (br == 1 ? h1a.y : h2a.y);
  const int lx = br == 0 ? l0 : (br == 1 ? l1 : l2);
  z.idx = (retry ? hy : hx) & 0xFFFFFFu;
  z.tok_len = retry ? int(hy >> 24) - off : lx;
  z.fd = retry ? 1 : 0;
  z.i = i + adv;
  z.have = (pick >= 0) | (l0 > 0);
}

}  // namespace tmgpu

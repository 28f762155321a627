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
// character classes.  The numbering is chosen for the encoder's table index: letters are the
// classes with bit 2 set.  other, digit, apostrophe, space (behaves as "other" except where the
// encoder asks for a space), upper, lower
enum : uint32_t { PC_O = 0, PC_N = 1, PC_A = 2, PC_S = 3, PC_U = 4, PC_L = 5 };

TM_HD uint32_t pc_class(uint32_t c) {
  if (c - 'a' < 26u) return PC_L;
  if (c - 'A' < 26u) return PC_U;
  if (c - '0' < 10u) return PC_N;
  if (c == '\'') return PC_A;
  return PC_O;
}
TM_HD uint32_t pc_class_sp(uint32_t c) { return c == ' ' ? uint32_t(PC_S) : pc_class(c); }
// what the encoder keeps of a raw character: the byte it will emit for it (capitals lower-cased)
// | class << 8.  The table kernel reads this from a 256-entry shared-memory table.
TM_HD uint32_t revcap_char(uint32_t c) { return (c - 'A' < 26u ? (c | 0x20u) : c) | (pc_class_sp(c) << 8); }

// ------------------------------------------------------------------------------------------
// stage 1, right to left: the encoder (a forward-looking restatement of js:900-1005: the marker
// of a capitalised run -- W = stays capital, C = mixed case -- is decided where the run starts, so
// nothing already emitted is ever rewritten) driven from the END of the string, one raw character
// per call.  This is what the table kernel runs (the automaton consumes the normalized text back
// to front, so producing it back to front lets both live in one loop and the normalized text never
// exists in memory).  Everything the encoder knows about a character is local:
//   * the two characters to its LEFT are read ahead (their classes sit in a shift register);
//   * "inside a capitalised run" = the nearest character to the left that is not a digit or an
//     apostrophe is a capital: it is the left neighbour itself unless that is a digit / apostrophe,
//     in which case one look-back over the run answers for every character of the run (iw);
//   * "the run is mixed case" = the nearest lower-case-or-other character to the RIGHT is
//     lower-case, which a right-to-left pass carries along (mixed);
//   * whether the character to the right is a capital decides whether a space is replaced by a
//     marker (capf).
// The decision AND the carried bits of the next step are one table read:
//   index = (class k-2 is a letter) | class k-1 << 1 | class k << 4 | iw << 7 | mixed << 8 | capf << 9
//   entry = bytes to emit (<= 4) | iw << 3 | next mixed << 4 | next capf << 5 | (k-1 is a digit /
//           apostrophe) << 6 | the bytes IN FRONT of the character, last one first, from bit 8
// so that the emitted bytes, character included, leave the queue lowest byte first -- the order the
// right-to-left automaton wants them in.
// ------------------------------------------------------------------------------------------
constexpr int kCapLutSize = 1024;

TM_HD uint32_t revcap_lut_entry(uint32_t idx) {
  const bool cl2_letter = idx & 1u;
  const uint32_t cl1 = (idx >> 1) & 7u, cl = (idx >> 4) & 7u;
  const bool iw = (idx >> 7) & 1u, mixed = (idx >> 8) & 1u, cap_follows = (idx >> 9) & 1u;
  const bool run1 = (cl1 == PC_N) || (cl1 == PC_A);
  const bool in_word = run1 ? iw : (cl1 == PC_U);
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
  // the bytes in front of the character, LAST ONE FIRST from bit 8 (the order they leave the queue)
  uint32_t rev = 0, np = 0;
  if (k_dspace) { rev = (uint32_t(' ') << 8) | (uint32_t('D') << 16); np = 2; }
  if (k_dcs) { rev = (uint32_t(' ') << 8) | (uint32_t('C') << 16) | (uint32_t('D') << 24); np = 3; }
  if (start) {
    if (sp1) { rev = (uint32_t(' ') << 8) | (marker << 16); np = 2; }
    else { rev = (uint32_t(' ') << 8) | (marker << 16) | (uint32_t('D') << 24); np = 3; }
  }
  // a space right before the first letter of a capitalised word is replaced by the marker
  // (js:976-979): the capital emits "<marker> x" and the space emits nothing
  const bool eaten = (cl == PC_S) && cap_follows;
  const bool next_mixed = (cl == PC_L) ? true : ((cl == PC_O || cl == PC_S) ? false : mixed);
  return (eaten ? 0u : np + 1u) | (uint32_t(iw) << 3) | (uint32_t(next_mixed) << 4) | (uint32_t(cl == PC_U) << 5) |
         (uint32_t(run1) << 6) | rev;
}

struct RevCap {
  uint32_t cls;     // classes of the characters k-2 (bits 0-2), k-1 (3-5), k (6-8); older ones above
  uint32_t chr;     // their bytes (capitals lower-cased): k (bits 0-7), k-1, k-2
  uint32_t carry;   // iw << 7 | mixed << 8 | capf << 9: the table index bits carried from the right
  bool iw_valid;    // iw answers for the current run of digits / apostrophes
};

TM_HD void revcap_reset(RevCap& e) {
  e.cls = 0;
  e.chr = 0;
  e.carry = 0;
  e.iw_valid = false;
}
// The read-ahead takes one more character to the left: x = revcap_char() of it, 0 when the string
// has none left (class "other").  Three pushes position the first character.
TM_HD void revcap_push(RevCap& e, uint32_t x) {
  e.cls = (e.cls << 3) + (x >> 8);
  e.chr = (e.chr >> 8) | ((x & 0xFFu) << 16);
}

// Encodes the raw character at index k (the lowest of the three characters pushed last).
// class_at(j) = class of character j (rare deeper look-back); lut(i) = entry i of the kCapLutSize
// entries of revcap_lut_entry().  Returns the number of bytes (0..4) placed in *out, FIRST TO LEAVE in the
// lowest byte: the character itself, then the bytes in front of it back to front (checked against
// the oracle's encoder, tests/test_host_sim.py).  The caller pushes the next character afterwards.
template <class Lut, class ClassAt>
TM_HD uint32_t revcap_step(RevCap& e, Lut lut, int k, ClassAt class_at, uint32_t* out) {
  const uint32_t base = (e.cls >> 2) & 0x7Fu;
  uint32_t a = lut(base | e.carry);
  // the left neighbour is a digit / apostrophe and nothing is known about the run yet: one
  // look-back (rare) answers for the whole run -- is the nearest other character a capital?
  if (((a >> 6) & 1u) && !e.iw_valid) {
    uint32_t x = e.cls & 7u;
    int j = k - 2;
    while ((x == PC_N || x == PC_A) && j > 0) {
      j--;
      x = class_at(j);
    }
    e.carry = (e.carry & ~0x80u) | (((k >= 2) && (x == PC_U)) ? 0x80u : 0u);
    a = lut(base | e.carry);
  }
  e.iw_valid = ((a >> 6) & 1u) != 0;
  *out = (a & 0xFFFFFF00u) | (e.chr & 0xFFu);
  e.carry = (a << 4) & 0x380u;
  return a & 7u;
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

// Classes of the bytes to the right of the current position, 2 bits each: the byte at distance d
// (1 <= d <= 47; a key has at most 40 bytes) sits at bits 2d, 2d+1 of a 96-bit shift register, and
// bits 0, 1 stay zero, so that "distance 0" (no match) reads as class 0 without a test.  Bytes
// behind the end of the string are the pad byte 0 (defined as 0).
struct ClassTrail {
  uint32_t w0, w1, w2;   // distances 0..15, 16..31, 32..47
};
TM_HD void trail_reset(ClassTrail& t, uint32_t code0) {
  t.w0 = code0 * 0x55555554u;
  t.w1 = code0 * 0x55555555u;
  t.w2 = code0 * 0x55555555u;
}
TM_HD void trail_push(ClassTrail& t, uint32_t code) {
  t.w2 = (t.w2 << 2) | (t.w1 >> 30);
  t.w1 = (t.w1 << 2) | (t.w0 >> 30);
  t.w0 = (t.w0 << 2) | (code << 2);
}
// class of the byte d positions to the right; 0 for d == 0
TM_HD uint32_t trail_at(const ClassTrail& t, uint32_t d) {
  const uint32_t w = d < 16u ? t.w0 : (d < 32u ? t.w1 : t.w2);
  return (w >> ((2u * d) & 31u)) & 3u;
}

// The word of the position whose byte (class code `code`) was just consumed into `packed`.
// `off` = bytes of the virtual space (1, or 2 for UTF-16).
TM_HD uint32_t position_word(uint32_t packed, uint32_t code, const ClassTrail& t, uint32_t off) {
  const uint32_t l = dfa_len(packed), sl = dfa_splen(packed);
  const uint32_t nb = trail_at(t, l);                              // l == 0: no match, class 0
  const uint32_t nbb = trail_at(t, sl - (sl < off ? sl : off));    // distance 0 unless sl > off
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

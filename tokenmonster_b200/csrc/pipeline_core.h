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
// stage 1: capcode level 2 encoder, one raw character per call
// ------------------------------------------------------------------------------------------
enum : uint32_t { PC_O = 0, PC_U = 1, PC_L = 2, PC_N = 3, PC_A = 4 };   // other, upper, lower, digit, apostrophe

TM_HD uint32_t pc_class(uint32_t c) {
  if (c - 'a' < 26u) return PC_L;
  if (c - 'A' < 26u) return PC_U;
  if (c - '0' < 10u) return PC_N;
  if (c == '\'') return PC_A;
  return PC_O;
}

struct CapState {
  uint32_t cl1, cl2;        // classes of the previous two raw characters
  bool in_word, mixed, sp1; // inside a capitalised run; that run is mixed case; previous char was ' '
};

TM_HD void cap_reset(CapState& e) {
  e.cl1 = PC_O;
  e.cl2 = PC_O;
  e.in_word = false;
  e.mixed = false;
  e.sp1 = false;
}

// Encodes raw character c (index k of n; next character cn, '.' behind the end).  `cls` maps an
// ASCII byte to its class; `raw_at(j)` reads raw character j (only used by the rare look-ahead
// over a run of capitals / digits / apostrophes).  Returns the number of bytes (0..4) placed in
// *out, little endian.  Forward-looking restatement of js:900-1005: the marker of a capitalised
// run (W = stays capital, C = mixed case) is decided when the run starts, from the class of the
// first lower-case-or-other character behind it, so nothing already emitted is ever rewritten.
template <class RawAt>
TM_HD uint32_t cap_step(CapState& e, const uint8_t* cls, uint32_t c, uint32_t cn, int k, int n, RawAt raw_at,
                        uint32_t* out) {
  const bool has_next = (k + 1 < n);
  const uint32_t cl = cls[c & 127u], cln = has_next ? cls[cn & 127u] : uint32_t(PC_O);
  const bool letter1 = (e.cl1 == PC_U) || (e.cl1 == PC_L);
  const bool glue1 = letter1 || (e.cl1 == PC_A);
  const bool start = (cl == PC_U) && !e.in_word;
  if (start) {
    uint32_t k_cl = cln;
    if (!(k_cl == PC_L || k_cl == PC_O)) {   // rare: a run of 2+ capitals / digits / apostrophes
      k_cl = PC_O;
      for (int j = k + 2; j < n; j++) {
        const uint32_t x = cls[raw_at(j) & 127u];
        if (x == PC_L || x == PC_O) { k_cl = x; break; }
      }
    }
    e.mixed = (k_cl == PC_L);
  }
  const bool joined_l = e.in_word ? glue1 : (e.sp1 || letter1 || (((e.cl2 == PC_U) || (e.cl2 == PC_L)) && e.cl1 == PC_A));
  const bool joined_n = e.in_word ? (e.cl1 == PC_N) : (e.sp1 || e.cl1 == PC_N);
  const bool k_dspace = (cl == PC_L && !joined_l) || (cl == PC_N && !joined_n) ||
                        (cl == PC_U && e.in_word && !e.mixed && !glue1);
  const bool k_dcs = (cl == PC_U) && e.in_word && e.mixed;
  const uint32_t marker = e.mixed ? 'C' : 'W';
  uint32_t pre = 0, np = 0;
  if (k_dspace) { pre = 'D' | (' ' << 8); np = 2; }
  if (k_dcs) { pre = 'D' | ('C' << 8) | (' ' << 16); np = 3; }
  if (start) {
    pre = e.sp1 ? (marker | (' ' << 8)) : ('D' | (marker << 8) | (' ' << 16));
    np = e.sp1 ? 2 : 3;
  }
  // a space right before the first letter of a capitalised word is replaced by the marker
  // (js:976-979): the capital emits "<marker> x" and the space emits nothing
  const bool eaten = (c == ' ') && has_next && (cln == PC_U);
  const uint32_t ch = (cl == PC_U) ? (c | 0x20u) : c;
  *out = eaten ? pre : (pre | (ch << (8 * np)));
  e.in_word = (cl == PC_U) ? true : ((cl == PC_L || cl == PC_O) ? false : e.in_word);
  e.cl2 = e.cl1;
  e.cl1 = cl;
  e.sp1 = (c == ' ');
  return np + (eaten ? 0u : 1u);
}

// Four raw characters at once: all lower-case letters or spaces, behind a letter or a space, encode
// to themselves (js:969-974: no "D " is due; js:976-979: no capital follows inside; the rule for
// a trailing space only needs the fifth character c5, 0 behind the end).
TM_HD bool cap_fast4(CapState& e, uint32_t w4, uint32_t c5) {
  const uint32_t low = (w4 + 0x1F1F1F1Fu) & ~(w4 + 0x05050505u);             // 'a'..'z'
  const uint32_t y = w4 ^ 0x20202020u;
  const uint32_t spc = ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y);              // == ' '
  const bool prev_ok = e.sp1 || e.cl1 == PC_U || e.cl1 == PC_L;
  const bool tail_ok = ((w4 >> 24) != ' ') || !(c5 - 'A' < 26u);
  if (!(prev_ok && tail_ok && (w4 & 0x80808080u) == 0 && ((low | spc) & 0x80808080u) == 0x80808080u)) return false;
  const uint32_t b2 = (w4 >> 16) & 0xFFu, b3 = w4 >> 24;
  e.in_word = false;
  e.cl2 = (b2 == ' ') ? uint32_t(PC_O) : uint32_t(PC_L);
  e.cl1 = (b3 == ' ') ? uint32_t(PC_O) : uint32_t(PC_L);
  e.sp1 = (b3 == ' ');
  return true;
}

// ------------------------------------------------------------------------------------------
// stage 1, right to left: the same encoder driven from the END of the string, one raw character
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

// Encodes raw character c0 at index k; c1 / c2 = the characters at k-1 / k-2 (ignored where they do
// not exist); raw_at(j) reads character j (rare deeper look-back).  Returns the number of bytes
// (0..4) placed in *out in FORWARD order, little endian -- the caller feeds them to the automaton
// last byte first.  Same output as cap_step() run left to right (tests/test_host_sim.py).
template <class RawAt>
TM_HD uint32_t revcap_step(RevCap& e, const uint8_t* cls, uint32_t c0, uint32_t c1, uint32_t c2, int k, RawAt raw_at,
                           uint32_t* out) {
  const uint32_t cl = cls[c0 & 127u];
  const uint32_t cl1 = (k >= 1) ? uint32_t(cls[c1 & 127u]) : uint32_t(PC_O);
  const uint32_t cl2 = (k >= 2) ? uint32_t(cls[c2 & 127u]) : uint32_t(PC_O);
  const bool sp1 = (k >= 1) && (c1 == ' ');
  bool in_word;
  if (cl1 != PC_N && cl1 != PC_A) {        // also k == 0: no left neighbour counts as "other"
    in_word = (cl1 == PC_U);
    e.iw_valid = false;
  } else {
    if (!e.iw_valid) {
      uint32_t x = cl2;
      int j = k - 2;
      while ((x == PC_N || x == PC_A) && j > 0) {
        j--;
        x = cls[raw_at(j) & 127u];
      }
      e.iw = (k >= 2) && (x == PC_U);
      e.iw_valid = true;
    }
    in_word = e.iw;
  }
  const bool mixed = (e.nlo == PC_L);
  const bool letter1 = (cl1 == PC_U) || (cl1 == PC_L);
  const bool glue1 = letter1 || (cl1 == PC_A);
  const bool start = (cl == PC_U) && !in_word;
  const bool joined_l = in_word ? glue1 : (sp1 || letter1 || (((cl2 == PC_U) || (cl2 == PC_L)) && cl1 == PC_A));
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
  const bool eaten = (c0 == ' ') && e.has_right && (e.clr == PC_U);
  const uint32_t ch = (cl == PC_U) ? (c0 | 0x20u) : c0;
  *out = eaten ? pre : (pre | (ch << (8 * np)));
  e.clr = cl;
  e.has_right = true;
  if (cl == PC_L || cl == PC_O) e.nlo = cl;
  return np + (eaten ? 0u : 1u);
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
  uint32_t a, b, c;   // positions +1..+16, +17..+32, +33..+48
};
TM_HD void trail_reset(ClassTrail& t, uint32_t code0) {
  t.a = t.b = t.c = code0 * 0x55555555u;
}
TM_HD void trail_push(ClassTrail& t, uint32_t code) {
  t.c = (t.c << 2) | (t.b >> 30);
  t.b = (t.b << 2) | (t.a >> 30);
  t.a = (t.a << 2) | code;
}
// class of the byte d positions to the right (d >= 1)
TM_HD uint32_t trail_at(const ClassTrail& t, uint32_t d) {
  const uint32_t k = d - 1u;
  const uint32_t w = k < 16u ? t.a : (k < 32u ? t.b : t.c);
  return (w >> (2u * (k & 15u))) & 3u;
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
// stage 3: one step of the ungreedy loop
// ------------------------------------------------------------------------------------------
struct ChainTables {
  const uint4* rec;      // 2 x uint4 per dictionary index (vocab_tables.h: TokenRecord)
  const uint4* hot;      // per automaton state
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

// emit(id, pos_code, adv): pos_code = position | forward-delete << 31 of the step (0xFFFFFFFF for
// the delete token of a two-id step), adv = characters the step advances by.
// on_miss(position) -> ids added: a byte no token covers (go:1269-1276); the caller emits unk.
// two_ids: false for Vocab.Count, which adds 1 where tokenize emits 2 ids.
// load_rec(index, &r0, &r1): the 32-byte token record (one 256-bit load on the device).
template <class LoadRec, class Emit, class Miss>
TM_HD void chain_step(const ChainTables& T, const ChainString& S, ChainLane& z, bool two_ids, LoadRec load_rec, Emit emit,
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
    z.idx = T.hot[pw_state(w)].x & 0xFFFFFFu;
    z.tok_len = int(l);
    z.have = true;
  }
  uint4 r0, r1;
  load_rec(z.idx, &r0, &r1);
  const uint32_t oflag = r0.x & 0xFFu;
  const int onw = int((r0.x >> 8) & 0xFFu);
  const int alen1 = int(r0.x >> 24), alen2 = int(r1.x & 0xFFu);
  const int i = z.i, fd = z.fd, len = z.tok_len, L = S.len;
  const int i1 = i + len;
  const int base1 = alen1 - fd, base2 = alen2 - fd;
  // the three look-ahead positions (go:1068, 1113, 1165); alternative 2 only with alternative 1 (Q10)
  const uint32_t w0 = (i1 < L) ? S.words[i1] : 0u;
  const bool look = (i1 < L) && (((oflag & 32u) == 0) || pw_cls(w0) != 3u);   // go:1057 (class 3 = 12 = space)
  const bool has1 = look && alen1 != 0, has2 = has1 && alen2 != 0;
  const uint32_t w1 = (has1 && i + base1 < L) ? S.words[i + base1] : 0u;
  const uint32_t w2 = (has2 && i + base2 < L) ? S.words[i + base2] : 0u;
  const int l0 = int(pw_len(w0)), l1 = int(pw_len(w1)), l2 = int(pw_len(w2));
  uint4 h0 = make_uint4(0, 0, 0, 0), h1 = h0, h2 = h0;
  if (l0) h0 = T.hot[pw_state(w0)];
  if (l1) h1 = T.hot[pw_state(w1)];
  if (l2) h2 = T.hot[pw_state(w2)];
  int best_s = kNegScore, pick = -1;   // pick: 0..2 plain, 3..5 retry
  if (look) {
    const uint32_t aflag1 = (r1.x >> 8) & 0xFFu, aflag2 = (r1.x >> 24) & 0xFFu;
    const int anw1 = int((r1.x >> 16) & 0xFFu), anw2 = int(r1.y & 0xFFu);
    // plain candidates in the order s1, s2, s3 (strict > keeps the first maximum, go:1217-1262)
    const int nb0 = class_value(pw_nb(w0)), nb1 = class_value(pw_nb(w1)), nb2 = class_value(pw_nb(w2));
    if (l0 > 0) {
      const int sc = pair_score(len + l0, oflag, onw - fd, h0.z & 0xFFu, int((h0.z >> 8) & 0xFFu), nb0, false);
      if (sc > best_s) { best_s = sc; pick = 0; }
    }
    if (has1 && l1 > 0) {
      const int bl = base1 + l1;
      const int sc = pair_score(bl, aflag1, anw1 - fd, h1.z & 0xFFu, int((h1.z >> 8) & 0xFFu), nb1, false) -
                     100 * int(bl < len) - 10000 * int(bl == len);
      if (sc > best_s) { best_s = sc; pick = 1; }
    }
    if (has2 && l2 > 0) {
      const int bl = base2 + l2;
      const int sc = pair_score(bl, aflag2, anw2 - fd, h2.z & 0xFFu, int((h2.z >> 8) & 0xFFu), nb2, false) -
                     100 * int(bl < len) - 10000 * int(bl == len);
      if (sc > best_s) { best_s = sc; pick = 2; }
    }
    if (T.can_retry) {
      // retries in the order s1b, s2b, s3b (go:1088-1107): wanted when the second token begins
      // with a letter, has no word boundary inside and a letter follows it; accepted when the
      // space-prefixed key is more than one byte longer than the plain one
      const int off = int(T.off);
      if (l0 > 0 && (h0.z & 2u) && nb0 == 1 && ((h0.z >> 8) & 0xFFu) == 0 && int(h0.y >> 24) > l0 + 1) {
        const int lb = int(h0.y >> 24) - off;
        const int sc = pair_score(len + lb, oflag, onw - fd, (h0.z >> 16) & 0xFFu, int(h0.z >> 24), class_value(pw_nbb(w0)), true);
        if (sc > best_s) { best_s = sc; pick = 3; }
      }
      if (has1 && l1 > 0 && (h1.z & 2u) && nb1 == 1 && ((h1.z >> 8) & 0xFFu) == 0 && int(h1.y >> 24) > l1 + 1) {
        const int bl = base1 + int(h1.y >> 24) - off;
        const int sc = pair_score(bl, aflag1, anw1 - fd, (h1.z >> 16) & 0xFFu, int(h1.z >> 24), class_value(pw_nbb(w1)), true) -
                       100 * int(bl < len) - 10000 * int(bl == len);
        if (sc > best_s) { best_s = sc; pick = 4; }
      }
      if (has2 && l2 > 0 && (h2.z & 2u) && nb2 == 1 && ((h2.z >> 8) & 0xFFu) == 0 && int(h2.y >> 24) > l2 + 1) {
        const int bl = base2 + int(h2.y >> 24) - off;
        const int sc = pair_score(bl, aflag2, anw2 - fd, (h2.z >> 16) & 0xFFu, int(h2.z >> 24), class_value(pw_nbb(w2)), true) -
                       100 * int(bl < len) - 10000 * int(bl == len);
        if (sc > best_s) { best_s = sc; pick = 5; }
      }
    }
  }
  const uint32_t pos_code = uint32_t(i) | (uint32_t(fd) << 31);
  if (pick < 0) {   // no branch: emit the token, its successor is the longest token behind it (go:1265-1267)
    emit(r0.y, pos_code, uint32_t(len));
    z.cnt++;
    z.i = i1;
    z.fd = 0;
    z.have = (l0 > 0);
    z.idx = h0.x & 0xFFFFFFu;
    z.tok_len = l0;
    return;
  }
  const int br = pick >= 3 ? pick - 3 : pick;
  const int adv = br == 0 ? len : (br == 1 ? base1 : base2);
  emit(br == 0 ? r0.y : (br == 1 ? r0.z : r0.w), pos_code, uint32_t(adv));
  z.cnt++;
  const uint4 hh = br == 0 ? h0 : (br == 1 ? h1 : h2);
  if (pick >= 3) {
    if (two_ids) {
      emit(T.delete_token, 0xFFFFFFFFu, 1u);
      z.cnt++;
    }
    z.idx = hh.y & 0xFFFFFFu;
    z.tok_len = int(hh.y >> 24) - int(T.off);
    z.fd = 1;
  } else {
    z.idx = hh.x & 0xFFFFFFu;
    z.tok_len = br == 0 ? l0 : (br == 1 ? l1 : l2);
    z.fd = 0;
  }
  z.i = i + adv;
  z.have = true;
}

}  // namespace tmgpu

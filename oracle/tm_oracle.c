/*
 * tm_oracle.c -- CPU oracle (plain C restatement) of the TokenMonster hot path.
 * TEST INFRASTRUCTURE ONLY -- see tm_oracle.h for scope, citations and pinning.
 *
 * Structure differs deliberately from the reference: one generic loop with an emit
 * sink replaces the five textual copies of the tokenizer (go/tokenmonster.go:1017-2359),
 * one branch evaluator replaces the six inlined score expressions, and the dictionary
 * is a single open-addressed (length, bytes) hash probed from the longest length down
 * instead of pansearch's bloom/bucket cascade (tokenmonster-cpp/src/tokenmonster.cpp:491-1280).
 */
#include "tm_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ vocabulary */

typedef struct {
  uint8_t flag, n_words;
  uint8_t len;       /* key length */
  uint8_t alt_len;   /* key length of alternative 1 (0 if none) */
  uint8_t alt_len2;  /* key length of alternative 2 */
  uint32_t alt, alt2; /* dictionary indexes or TMO_DNE */
  uint32_t id, id1, id2;
  uint32_t key_off;  /* offset into keys blob */
} tmo_rec;

struct tmo_vocab {
  uint8_t capcode, charset, normflag, level, reserve;
  uint32_t unk, vocab_size, n_reverse, n_info, delete_token;
  int max_len;
  tmo_rec* rec;
  uint8_t* keys;
  uint8_t begin_byte[256];
  /* reverse[id] -> (offset,len) of the LAST record carrying that id (go:2715) */
  uint32_t* rev_off;
  uint8_t* rev_len; /* 0xFF = nil entry */
  /* open-addressed hash over (len, bytes) -> index */
  uint32_t* ht;
  uint32_t ht_mask;
};

static uint64_t hash_bytes(const uint8_t* p, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 0x100000001B3ull;
    h ^= h >> 29;
  }
  return h;
}

void tmo_free(void* p) { free(p); }

void tmo_vocab_free(tmo_vocab* v) {
  if (!v) return;
  free(v->rec);
  free(v->keys);
  free(v->rev_off);
  free(v->rev_len);
  free(v->ht);
  free(v);
}

static uint32_t rd24(const uint8_t* p) { return p[0] | (p[1] << 8) | ((uint32_t)p[2] << 16); }

int tmo_longest_prefix(const tmo_vocab* v, const uint8_t* w, size_t n, uint32_t* index) {
  size_t top = n < (size_t)v->max_len ? n : (size_t)v->max_len;
  for (size_t len = top; len >= 1; len--) {
    uint32_t slot = (uint32_t)hash_bytes(w, len) & v->ht_mask;
    for (;;) {
      uint32_t idx = v->ht[slot];
      if (idx == 0xFFFFFFFFu) break;
      const tmo_rec* r = &v->rec[idx];
      if (r->len == len && memcmp(v->keys + r->key_off, w, len) == 0) {
        *index = idx;
        return (int)len;
      }
      slot = (slot + 1) & v->ht_mask;
    }
  }
  return 0;
}

/* Load: go/tokenmonster.go:2656-2736; layout worked through in tests/unit.cpp:12-79 */
tmo_vocab* tmo_vocab_load(const uint8_t* f, size_t n, int* status) {
  int st = TMO_ERR_FORMAT;
  tmo_vocab* v = NULL;
  if (n < 8 + 15 + 1) goto fail;
  v = (tmo_vocab*)calloc(1, sizeof(*v));
  v->capcode = f[0];
  v->charset = f[1];
  v->normflag = f[2];
  v->level = f[3];
  v->reserve = f[4];
  if (v->charset > 2 || v->capcode > 2) goto fail;
  size_t at = 8;
  v->unk = rd24(f + at); at += 3;
  v->vocab_size = rd24(f + at); at += 3;
  v->n_reverse = rd24(f + at); at += 3;
  v->n_info = rd24(f + at); at += 3;
  v->delete_token = rd24(f + at); at += 3;
  v->max_len = f[at++];
  v->rec = (tmo_rec*)calloc(v->n_info ? v->n_info : 1, sizeof(tmo_rec));
  v->keys = (uint8_t*)malloc((size_t)v->n_info * 40 + 1);
  v->rev_off = (uint32_t*)calloc(v->n_reverse ? v->n_reverse : 1, sizeof(uint32_t));
  v->rev_len = (uint8_t*)malloc(v->n_reverse ? v->n_reverse : 1);
  memset(v->rev_len, 0xFF, v->n_reverse ? v->n_reverse : 1);
  uint32_t koff = 0;
  for (uint32_t i = 0; i < v->n_info; i++) {
    if (at + 1 > n) goto fail;
    uint32_t kl = f[at++];
    if (kl > 40 || at + kl + 15 > n) goto fail;
    tmo_rec* r = &v->rec[i];
    r->len = (uint8_t)kl;
    r->key_off = koff;
    memcpy(v->keys + koff, f + at, kl);
    koff += kl;
    at += kl;
    r->flag = f[at++];
    r->n_words = f[at++];
    r->alt = rd24(f + at); at += 3;
    r->alt2 = rd24(f + at); at += 3;
    r->id = rd24(f + at); at += 3;
    at += 4; /* float32 score: unused on the hot path */
    /* alternatives are resolved against EARLIER records only (go:2703-2711, :2636) */
    if (r->alt != TMO_DNE) {
      if (r->alt >= i) goto fail;
      r->alt_len = v->rec[r->alt].len;
      r->id1 = v->rec[r->alt].id;
    }
    if (r->alt2 != TMO_DNE) {
      if (r->alt2 >= i) goto fail;
      r->alt_len2 = v->rec[r->alt2].len;
      r->id2 = v->rec[r->alt2].id;
    }
    if (r->id >= v->n_reverse) goto fail;
    v->rev_off[r->id] = r->key_off;
    v->rev_len[r->id] = r->len;
    /* file order must be (length, bytes) ascending: cpp:1352-1357 */
    if (i > 0) {
      const tmo_rec* p = &v->rec[i - 1];
      int bad = p->len > r->len ||
                (p->len == r->len && memcmp(v->keys + p->key_off, v->keys + r->key_off, kl) >= 0);
      if (bad) { st = TMO_ERR_ORDER; goto fail; }
    }
  }
  if (at + 256 + 3 > n) goto fail;
  memcpy(v->begin_byte, f + at, 256);
  at += 256;
  uint32_t n_del = rd24(f + at); at += 3;
  for (uint32_t i = 0; i < n_del; i++) {
    if (at + 1 > n) goto fail;
    uint32_t kl = f[at++];
    if (at + kl + 7 > n) goto fail;
    at += kl + 7;
  }
  if (at != n) goto fail; /* go:2731 */

  uint32_t cap = 16;
  while (cap < v->n_info * 2u + 2u) cap <<= 1;
  v->ht = (uint32_t*)malloc(cap * sizeof(uint32_t));
  memset(v->ht, 0xFF, cap * sizeof(uint32_t));
  v->ht_mask = cap - 1;
  for (uint32_t i = 0; i < v->n_info; i++) {
    const tmo_rec* r = &v->rec[i];
    uint32_t slot = (uint32_t)hash_bytes(v->keys + r->key_off, r->len) & v->ht_mask;
    while (v->ht[slot] != 0xFFFFFFFFu) slot = (slot + 1) & v->ht_mask;
    v->ht[slot] = i;
  }
  if (status) *status = TMO_OK;
  return v;
fail:
  tmo_vocab_free(v);
  if (status) *status = st;
  return NULL;
}

int tmo_vocab_info(const tmo_vocab* v, uint32_t out[12]) {
  out[0] = v->capcode; out[1] = v->charset; out[2] = v->normflag; out[3] = v->level;
  out[4] = v->reserve; out[5] = v->unk; out[6] = v->vocab_size; out[7] = v->n_reverse;
  out[8] = v->n_info; out[9] = v->delete_token; out[10] = (uint32_t)v->max_len; out[11] = 0;
  return TMO_OK;
}

/* ------------------------------------------------------------------ capcode (level 2) */
/* javascript/tokenmonster.js:872-1065.  Character classes restricted to ASCII in this
 * build (Lu = A-Z, Ll = a-z, Nd = 0-9, M = none); callers reject non-ASCII input. */

static int is_up(int c) { return c >= 'A' && c <= 'Z'; }
static int is_lo(int c) { return c >= 'a' && c <= 'z'; }
static int is_let(int c) { return is_up(c) || is_lo(c); }
static int is_num(int c) { return c >= '0' && c <= '9'; }
static int glue(int c) { return is_let(c) || c == '\''; } /* letter / apostrophe / (modifier) */

typedef struct { uint8_t* p; size_t n, cap; } obuf;
static void ob_put(obuf* b, uint8_t c) {
  if (b->n == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 64;
    b->p = (uint8_t*)realloc(b->p, b->cap);
  }
  b->p[b->n++] = c;
}

/* Re-mark a capitalised run that turned out to be mixed case (js:924-951): every letter
 * that was upper-case in the input gets its own "D C ␠" prefix. */
static void respell_run(obuf* b, size_t from) {
  size_t seg = b->n - from;
  uint8_t* tmp = (uint8_t*)malloc(seg + 1);
  memcpy(tmp, b->p + from, seg);
  b->n = from;
  for (size_t k = 0; k < seg;) {
    if (tmp[k] == 'D' && k + 2 < seg && tmp[k + 1] == ' ') {
      if (is_lo(tmp[k + 2])) {
        ob_put(b, 'D'); ob_put(b, 'C'); ob_put(b, ' '); ob_put(b, tmp[k + 2]);
      } else {
        ob_put(b, tmp[k]); ob_put(b, tmp[k + 1]); ob_put(b, tmp[k + 2]);
      }
      k += 3;
    } else if (is_lo(tmp[k])) {
      ob_put(b, 'D'); ob_put(b, 'C'); ob_put(b, ' '); ob_put(b, tmp[k]);
      k += 1;
    } else {
      ob_put(b, tmp[k]);
      k += 1;
    }
  }
  free(tmp);
}

int tmo_capcode_encode(const uint8_t* in, size_t n, uint8_t** out, size_t* out_n) {
  obuf b = {0, 0, 0};
  b.cap = n + n / 2 + 16;
  b.p = (uint8_t*)malloc(b.cap);
  int rlast = '.', rlast2 = '.';
  int in_word = 0, multi = 0;
  size_t word_tok = 0, goback = 0;
  for (size_t i = 0; i < n; i++) {
    int r = in[i];
    if (r >= 0x80) { free(b.p); return TMO_ERR_NORMALIZE; }
    if (in_word) {
      if (is_up(r)) {
        if (!glue(rlast)) { ob_put(&b, 'D'); ob_put(&b, ' '); }
        multi = 1;
        ob_put(&b, (uint8_t)(r | 0x20));
      } else {
        if (is_lo(r)) {
          in_word = 0;
          b.p[word_tok] = 'C';
          if (multi) respell_run(&b, goback);
          if (!glue(rlast)) { ob_put(&b, 'D'); ob_put(&b, ' '); }
        } else if (is_num(r)) {
          if (!is_num(rlast)) { ob_put(&b, 'D'); ob_put(&b, ' '); }
        } else if (r != '\'') {
          in_word = 0;
        }
        ob_put(&b, (uint8_t)r);
      }
    } else if (is_lo(r)) {
      if (!(rlast == ' ' || is_let(rlast) || (is_let(rlast2) && rlast == '\''))) {
        ob_put(&b, 'D'); ob_put(&b, ' ');
      }
      ob_put(&b, (uint8_t)r);
    } else if (is_up(r)) {
      if (rlast == ' ') {
        word_tok = b.n - 1;
        b.p[word_tok] = 'W';
        ob_put(&b, ' ');
      } else {
        ob_put(&b, 'D');
        word_tok = b.n;
        ob_put(&b, 'W');
        ob_put(&b, ' ');
      }
      ob_put(&b, (uint8_t)(r | 0x20));
      goback = b.n;
      multi = 0;
      in_word = 1;
    } else if (is_num(r)) {
      if (!(rlast == ' ' || is_num(rlast))) { ob_put(&b, 'D'); ob_put(&b, ' '); }
      ob_put(&b, (uint8_t)r);
    } else {
      ob_put(&b, (uint8_t)r);
    }
    rlast2 = rlast;
    rlast = r;
  }
  *out = b.p;
  *out_n = b.n;
  return TMO_OK;
}

size_t tmo_capcode_decode(tmo_capcode_decoder* s, const uint8_t* in, size_t n, uint8_t* out) {
  size_t o = 0;
  for (size_t i = 0; i < n; i++) {
    int r = in[i];
    if (r == 'C') { s->in_char = 1; s->in_word = 0; continue; }
    if (r == 'W') { s->in_word = 1; s->in_char = 0; s->ignore = 1; continue; }
    if (r == 'D') { s->del = 1; continue; }
    if (r == ' ') {
      if (s->del) s->del = 0;
      else { out[o++] = ' '; if (!s->ignore) s->in_word = 0; }
    } else if (r >= 0x80) {
      /* ASCII build: non-ASCII code points pass through; a pending delete swallows the
       * whole UTF-8 sequence.  Case mapping of non-ASCII letters is NOT applied. */
      if (s->del) {
        while (i + 1 < n && (in[i + 1] & 0xC0) == 0x80) i++;
        s->del = 0;
      } else {
        s->in_char = 0;
        out[o++] = (uint8_t)r;
        while (i + 1 < n && (in[i + 1] & 0xC0) == 0x80) out[o++] = in[++i];
      }
    } else if (s->del) {
      s->del = 0;
    } else if (s->in_char) {
      s->in_char = 0;
      out[o++] = (uint8_t)(is_lo(r) ? r - 32 : r);
    } else if (s->in_word) {
      if (is_let(r)) out[o++] = (uint8_t)(is_lo(r) ? r - 32 : r);
      else {
        out[o++] = (uint8_t)r;
        if (!(is_num(r) || r == '\'')) s->in_word = 0;
      }
    } else {
      out[o++] = (uint8_t)r;
    }
    s->ignore = 0;
  }
  return o;
}

/* ------------------------------------------------------------------ normalize */
/* go:242-253 -> norm.Normalizer.Normalize then applyCapcode.  Flag bits 1 (NFD),
 * 2 (lowercase) and 4 (accents) are supported for ASCII input, where NFD and accent
 * removal are the identity (cpp:191-198).  Other bits and non-ASCII input with any of
 * those bits set return TMO_ERR_NORMALIZE in this build (scope row a1). */
int tmo_normalize(const tmo_vocab* v, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n) {
  uint8_t flag = v->normflag;
  if (flag & ~7u) return TMO_ERR_NORMALIZE;
  if (v->capcode == 1) return TMO_ERR_NORMALIZE; /* NoCapcodeEncode: no specification exists */
  if (flag != 0 || v->capcode != 0) {
    for (size_t i = 0; i < n; i++)
      if (in[i] & 0x80) return TMO_ERR_NORMALIZE;
  }
  uint8_t* tmp = (uint8_t*)malloc(n + 1);
  memcpy(tmp, in, n);
  if (flag & 2)
    for (size_t i = 0; i < n; i++)
      if (is_up(tmp[i])) tmp[i] |= 0x20;
  if (v->capcode == 2) {
    int st = tmo_capcode_encode(tmp, n, out, out_n);
    free(tmp);
    return st;
  }
  *out = tmp;
  *out_n = n;
  return TMO_OK;
}

/* ------------------------------------------------------------------ tokenize */

typedef struct {
  int mode; /* 0 ids, 1 count, 2/3/4 serialized, 5 training scores (training/trainvocab.go:932-1170) */
  uint8_t* p;
  size_t n_ids, bytes, cap;
  uint64_t* scores;    /* mode 5: characters covered per token id */
  uint64_t* miss_hist; /* mode 5: occurrences of each byte value no token covers */
} sink;

/* adv = characters the step advances by: what the training worker adds to scores[id]
 * (trainvocab.go:1109-1162; 1 for the delete token, :1134) */
static void emit(sink* s, uint32_t id, int adv) {
  s->n_ids++;
  if (s->mode == 5) {
    if (id != TMO_DNE) s->scores[id] += (uint64_t)adv;
    return;
  }
  if (s->mode == 1) return;
  size_t w = s->mode == 0 ? 4 : (size_t)s->mode;
  if (s->bytes + 4 > s->cap) {
    s->cap = s->cap ? s->cap * 2 : 256;
    s->p = (uint8_t*)realloc(s->p, s->cap);
  }
  s->p[s->bytes] = (uint8_t)id;
  s->p[s->bytes + 1] = (uint8_t)(id >> 8);
  if (w > 2) s->p[s->bytes + 2] = (uint8_t)(id >> 16);
  if (w > 3) s->p[s->bytes + 3] = (uint8_t)(id >> 24);
  s->bytes += w;
}

#define NEG (-1000000)

/* One candidate continuation: first token `a` (flag/nWords; nWords already reduced by the
 * forward-delete), second token `b` found at the look-ahead position, `nb` = beginByte of
 * the byte after it.  SURVEY Appendix A.4 / go:1075-1084 (plain) and :1096-1105 (retry). */
static int score_pair(int branch_len, uint8_t aflag, int nw, uint8_t bflag, int bnw, int nb,
                      int retry) {
  int s = branch_len + (aflag >> 7) + (bflag >> 7) + (nw > 1 ? nw - 1 : 0) +
          (bnw > 1 ? bnw - 1 : 0) + ((nb >> 2) & 1) + 100 * (nw + bnw + (nb >> 3));
  s -= 100 * ((aflag >> 3) & 1 & (bflag >> 4)) + 3 * (bflag & 1 & nb);
  if (retry) s -= 103 * (aflag & 1) + 1;
  else { s += (bflag >> 2) & 1; s -= 103 * (aflag & 1 & (bflag >> 1)); }
  return s;
}

typedef struct {
  int score, score_b;       /* plain / space-prefixed retry */
  uint32_t idx, idx_b;      /* second token */
  int len, len_b;           /* data bytes the second token covers */
} cand;

/* Evaluate the continuation that starts at data position p after a first token with
 * (aflag, nw).  base = bytes the first token consumes; ref_len < 0 for the primary branch,
 * otherwise the primary length used for the alt penalties (go:1132-1133). */
/* Diagnostics for the design notes (tools/walk_reuse.py): how many look-ahead searches the loop
 * performs and how many of the plain ones repeat a position searched during the previous steps. */
static uint64_t g_dbg[4];        /* plain searches, retry searches, plain repeats, steps */
static int g_recent[8], g_recent_n;
void tmo_debug_walk_stats(uint64_t out[4], int reset) {
  for (int k = 0; k < 4; k++) out[k] = g_dbg[k];
  if (reset) { memset(g_dbg, 0, sizeof g_dbg); g_recent_n = 0; }
}

static void eval_branch(const tmo_vocab* v, const uint8_t* data, int len_data, int p, int base,
                        uint8_t aflag, int nw, int ref_len, int off, cand* c) {
  c->score = c->score_b = NEG;
  int room = len_data - p;
  if (room < 0) return;
  g_dbg[0]++;
  for (int k = 0; k < g_recent_n; k++)
    if (g_recent[k] == p) { g_dbg[2]++; break; }
  if (g_recent_n < 8) g_recent[g_recent_n++] = p;
  else { memmove(g_recent, g_recent + 1, 7 * sizeof(int)); g_recent[7] = p; }
  uint32_t x;
  int l = tmo_longest_prefix(v, data + p, (size_t)(room < v->max_len ? room : v->max_len), &x);
  if (l == 0) return;
  const tmo_rec* b = &v->rec[x];
  int nb = v->begin_byte[data[p + l]];
  int bl = base + l;
  c->idx = x;
  c->len = l;
  c->score = score_pair(bl, aflag, nw, b->flag, b->n_words, nb, 0);
  if (ref_len >= 0) c->score -= 100 * (bl < ref_len) + 10000 * (bl == ref_len);
  if (v->delete_token != TMO_DNE && (b->flag & 2) && nb == 1 && b->n_words == 0) {
    /* retry with a virtual space (2 bytes 0x20 0x00 in UTF-16) in front: go:1088-1107 */
    uint8_t lil[48];
    int mws = v->max_len - off;
    int take = room < mws ? room : mws;
    if (take < 0) take = 0;
    lil[0] = 32;
    lil[1] = 0;
    memcpy(lil + off, data + p, (size_t)take);
    uint32_t xb = 0;
    g_dbg[1]++;
    int lb = tmo_longest_prefix(v, lil, (size_t)(take + off), &xb);
    if (lb > l + 1) {
      lb -= off;
      const tmo_rec* b2 = &v->rec[xb];
      int nb2 = v->begin_byte[data[p + lb]];
      int bl2 = base + lb;
      c->idx_b = xb;
      c->len_b = lb;
      c->score_b = score_pair(bl2, aflag, nw, b2->flag, b2->n_words, nb2, 1);
      if (ref_len >= 0) c->score_b -= 100 * (bl2 < ref_len) + 10000 * (bl2 == ref_len);
    }
  }
}

/* go:1017-1279.  data must have one readable byte past len_data; it is defined as 0
 * (cpp:1726; the Go slice reads whatever follows -- SURVEY quirk Q1). */
static void tokenize_core(const tmo_vocab* v, const uint8_t* data, int len_data, sink* out,
                          int64_t* missing_out) {
  int64_t missing = 0;
  int off = v->charset == 2 ? 2 : 1;
  int i = 0, fd = 0;
  while (i < len_data) {
    uint32_t index;
    int room = len_data - i;
    int length = tmo_longest_prefix(v, data + i, (size_t)(room < v->max_len ? room : v->max_len), &index);
    if (length == 0) {
      if (out->mode == 5) { /* trainvocab.go:1164-1171: counted, never scored */
        out->n_ids++;
        out->miss_hist[data[i]]++;
      } else if (v->unk != TMO_DNE) {
        emit(out, v->unk, 0);
      }
      i++;
      missing++;
      fd = 0;
      continue;
    }
    for (;;) { /* "checkpoint" */
      g_dbg[3]++;
      const tmo_rec* o = &v->rec[index];
      int i1 = i + length;
      int advanced = 0;
      if (i1 < len_data && ((o->flag & 32) == 0 || v->begin_byte[data[i1]] != 12)) {
        cand c1, c2, c3;
        c2.score = c2.score_b = c3.score = c3.score_b = NEG;
        eval_branch(v, data, len_data, i1, length, o->flag, (int)o->n_words - fd, -1, off, &c1);
        if (o->alt != TMO_DNE) {
          const tmo_rec* a = &v->rec[o->alt];
          eval_branch(v, data, len_data, i + o->alt_len - fd, o->alt_len - fd, a->flag,
                      (int)a->n_words - fd, length, off, &c2);
          if (o->alt2 != TMO_DNE) { /* nested: alt2 only when alt1 exists (go:1163) */
            const tmo_rec* a2 = &v->rec[o->alt2];
            eval_branch(v, data, len_data, i + o->alt_len2 - fd, o->alt_len2 - fd, a2->flag,
                        (int)a2->n_words - fd, length, off, &c3);
          }
        }
        int best = c1.score;
        if (c1.score_b > best) best = c1.score_b;
        if (c2.score > best) best = c2.score;
        if (c2.score_b > best) best = c2.score_b;
        if (c3.score > best) best = c3.score;
        if (c3.score_b > best) best = c3.score_b;
        /* first match wins in this order: go:1217-1262 */
        if (best != NEG) {
          advanced = 1;
          if (best == c1.score) {
            emit(out, o->id, length); i += length; index = c1.idx; length = c1.len; fd = 0;
          } else if (best == c2.score) {
            emit(out, o->id1, o->alt_len - fd); i += o->alt_len - fd; index = c2.idx; length = c2.len; fd = 0;
          } else if (best == c3.score) {
            emit(out, o->id2, o->alt_len2 - fd); i += o->alt_len2 - fd; index = c3.idx; length = c3.len; fd = 0;
          } else if (best == c1.score_b) {
            emit(out, o->id, length); emit(out, v->delete_token, 1); if (out->mode == 1) out->n_ids--;
            i += length; index = c1.idx_b; length = c1.len_b; fd = 1;
          } else if (best == c2.score_b) {
            emit(out, o->id1, o->alt_len - fd); emit(out, v->delete_token, 1); if (out->mode == 1) out->n_ids--;
            i += o->alt_len - fd; index = c2.idx_b; length = c2.len_b; fd = 1;
          } else {
            emit(out, o->id2, o->alt_len2 - fd); emit(out, v->delete_token, 1); if (out->mode == 1) out->n_ids--;
            i += o->alt_len2 - fd; index = c3.idx_b; length = c3.len_b; fd = 1;
          }
        }
      }
      if (advanced) continue;
      emit(out, o->id, length);
      i += length;
      fd = 0;
      break;
    }
  }
  *missing_out = missing;
}

int tmo_tokenize(const tmo_vocab* v, const uint8_t* in, size_t n, int normalized, int mode,
                 void** out, size_t* n_tokens, int64_t* missing) {
  if (!(mode == 0 || mode == 1 || mode == 2 || mode == 3 || mode == 4)) return TMO_ERR_ARG;
  *out = NULL;
  *n_tokens = 0;
  *missing = 0;
  if (v->max_len == 0) return TMO_OK; /* go:960-962 */
  uint8_t* norm = NULL;
  size_t nn = n;
  if (!normalized) {
    int st = tmo_normalize(v, in, n, &norm, &nn);
    if (st != TMO_OK) return st;
  } else {
    norm = (uint8_t*)malloc(n + 1);
    memcpy(norm, in, n);
  }
  norm = (uint8_t*)realloc(norm, nn + 1);
  norm[nn] = 0; /* defined look-ahead byte */
  sink s = {mode, NULL, 0, 0, 0, NULL, NULL};
  tokenize_core(v, norm, (int)nn, &s, missing);
  free(norm);
  *out = s.p;
  *n_tokens = s.n_ids;
  return TMO_OK;
}

/* The training worker's tokenization loop (training/trainvocab.go:932-1170): the same branch
 * decisions as Vocab.tokenize (the two loops are textual copies), but instead of ids it
 * accumulates scores[id] += characters covered (1 for the delete token), tokensInText (a
 * two-id step counts 2, a byte without a token counts 1) and the bytes without a token.
 * scores has n_reverse entries, miss_hist 256; both are ADDED to. */
int tmo_score(const tmo_vocab* v, const uint8_t* in, size_t n, int normalized, uint64_t* scores,
              uint64_t* miss_hist, uint64_t* tokens_in_text) {
  if (v->max_len == 0) return TMO_OK;
  uint8_t* norm = NULL;
  size_t nn = n;
  if (!normalized) {
    int st = tmo_normalize(v, in, n, &norm, &nn);
    if (st != TMO_OK) return st;
  } else {
    norm = (uint8_t*)malloc(n + 1);
    memcpy(norm, in, n);
  }
  norm = (uint8_t*)realloc(norm, nn + 1);
  norm[nn] = 0;
  sink s = {5, NULL, 0, 0, 0, scores, miss_hist};
  int64_t missing = 0;
  tokenize_core(v, norm, (int)nn, &s, &missing);
  free(norm);
  *tokens_in_text += s.n_ids;
  return TMO_OK;
}

/* ------------------------------------------------------------------ decode */

static uint32_t read_id(const uint8_t* p, int enc) {
  uint32_t x = p[0] | (p[1] << 8);
  if (enc > 2) x |= (uint32_t)p[2] << 16;
  if (enc > 3) x |= (uint32_t)p[3] << 24;
  return x;
}

/* go:813-948: concat reverse[id] for in-range ids (nil entries contribute nothing);
 * the 16-bit serialized path compares against uint16(len(reverse)-1) (go:861-864). */
static uint8_t* decode_raw(const tmo_vocab* v, const void* tokens, size_t count, int enc,
                           const uint8_t* prefix, size_t prefix_n, size_t* out_n) {
  const uint8_t* tb = (const uint8_t*)tokens;
  size_t w = enc == 0 ? 4 : (size_t)enc;
  size_t n_ids = enc == 0 ? count : count / w;
  size_t total = prefix_n;
  if (enc == 2 && v->n_reverse == 0) n_ids = 0;
  uint32_t lim16 = (uint32_t)((v->n_reverse - 1) & 0xFFFF);
  for (int pass = 0; pass < 2; pass++) {
    uint8_t* dst = NULL;
    if (pass == 1) {
      dst = (uint8_t*)malloc(total + 1);
      memcpy(dst, prefix, prefix_n);
      total = prefix_n;
    }
    for (size_t k = 0; k < n_ids; k++) {
      uint32_t id = enc == 0 ? ((const uint32_t*)tokens)[k] : read_id(tb + k * w, enc);
      int ok = enc == 2 ? (id <= lim16) : (id < v->n_reverse);
      if (!ok || v->rev_len[id] == 0xFF) continue;
      if (pass == 1) memcpy(dst + total, v->keys + v->rev_off[id], v->rev_len[id]);
      total += v->rev_len[id];
    }
    if (pass == 1) { *out_n = total; return dst; }
  }
  return NULL;
}

int tmo_decode(const tmo_vocab* v, const void* tokens, size_t count, int enc_len, int raw,
               uint8_t** out, size_t* out_n) {
  if (!(enc_len == 0 || enc_len == 2 || enc_len == 3 || enc_len == 4)) return TMO_ERR_ARG;
  size_t n = 0;
  uint8_t* d = decode_raw(v, tokens, count, enc_len, NULL, 0, &n);
  if (!raw && v->capcode == 2) {
    tmo_capcode_decoder st = {0, 0, 0, 0};
    uint8_t* o = (uint8_t*)malloc(n + 1);
    n = tmo_capcode_decode(&st, d, n, o);
    free(d);
    d = o;
  }
  *out = d;
  *out_n = n;
  return TMO_OK;
}

/* go:149-191 -- note the truncated-sequence case returns the MISSING byte count (quirk Q7) */
int tmo_incomplete_utf8(const uint8_t* b, size_t n) {
  if (n == 0 || (b[n - 1] & 0x80) == 0) return 0;
  long s = (long)n - 1;
  while (s >= 0 && (b[s] & 0xC0) == 0x80) s--;
  if (s < 0) return (int)n;
  int need;
  uint8_t f = b[s];
  if ((f & 0x80) == 0) need = 1;
  else if ((f & 0xE0) == 0xC0) need = 2;
  else if ((f & 0xF0) == 0xE0) need = 3;
  else if ((f & 0xF8) == 0xF0) need = 4;
  else return (int)(n - (size_t)s);
  int have = (int)(n - (size_t)s);
  if (have < need) return need - have;
  if (need == 1 && (f & 0xC0) != 0) return (int)n;
  return 0;
}

/* go:193-219 */
int tmo_incomplete_utf16(const uint8_t* b, size_t n) {
  if (n == 0) return 0;
  if (n & 1) {
    if (n >= 3) {
      uint32_t u = b[n - 3] | (b[n - 2] << 8);
      if (u >= 0xD800 && u <= 0xDBFF) return 3;
    }
    return 1;
  }
  uint32_t last = b[n - 2] | (b[n - 1] << 8);
  if (last >= 0xD800 && last <= 0xDBFF) return 2;
  uint32_t first = b[0] | (b[1] << 8);
  if (first >= 0xDC00 && first <= 0xDFFF) return 2;
  return 0;
}

tmo_decoder* tmo_decoder_new(const tmo_vocab* v) {
  tmo_decoder* d = (tmo_decoder*)calloc(1, sizeof(*d));
  d->vocab = v;
  return d;
}

void tmo_decoder_free(tmo_decoder* d) {
  if (!d) return;
  free(d->remainder);
  free(d);
}

/* go:701-740 (ids) and :466-698 (serialized) */
int tmo_decoder_decode(tmo_decoder* d, const void* tokens, size_t count, int enc_len,
                       uint8_t** out, size_t* out_n) {
  const tmo_vocab* v = d->vocab;
  if (!(enc_len == 0 || enc_len == 2 || enc_len == 3 || enc_len == 4)) return TMO_ERR_ARG;
  size_t n = 0;
  if (v->charset == 0) { /* returns before remainder / capcode handling (quirk Q6) */
    *out = decode_raw(v, tokens, count, enc_len, NULL, 0, &n);
    *out_n = n;
    return TMO_OK;
  }
  uint8_t* data = decode_raw(v, tokens, count, enc_len, d->remainder, d->remainder_len, &n);
  int inc = v->charset == 1 ? tmo_incomplete_utf8(data, n) : tmo_incomplete_utf16(data, n);
  size_t keep = n - (size_t)inc;
  free(d->remainder);
  d->remainder_len = n - keep;
  d->remainder = (uint8_t*)malloc(d->remainder_len + 1);
  memcpy(d->remainder, data + keep, d->remainder_len);
  n = keep;
  if (v->capcode == 2) {
    uint8_t* o = (uint8_t*)malloc(n + 1);
    n = tmo_capcode_decode(&d->cc, data, n, o);
    free(data);
    data = o;
  }
  *out = data;
  *out_n = n;
  return TMO_OK;
}

int tmo_decoder_flush(tmo_decoder* d, uint8_t** out, size_t* out_n) {
  *out = d->remainder ? d->remainder : (uint8_t*)malloc(1);
  *out_n = d->remainder_len;
  d->remainder = NULL;
  d->remainder_len = 0;
  return TMO_OK;
}

/* Checksum of a batch result: per-string FNV-1a over the ids, mixed with the string index, summed
 * (the fold oracle/ref_driver.cpp's tmr_tokenize_batch applies to the reference's own output, after
 * tokenmonster-cpp/tests/bench.cpp:30-37).  tokens = dense ids of enc_len (2, 3, 4) bytes each. */
uint64_t tmo_ids_checksum(const uint8_t* tokens, int enc_len, const uint64_t* tok_off, uint32_t n) {
  uint64_t c = 0;
  for (uint32_t s = 0; s < n; s++) {
    uint64_t h = 14695981039346656037ULL;
    for (uint64_t k = tok_off[s]; k < tok_off[s + 1]; k++) {
      const uint8_t* p = tokens + k * (uint64_t)enc_len;
      uint32_t id = p[0] | ((uint32_t)p[1] << 8);
      if (enc_len > 2) id |= (uint32_t)p[2] << 16;
      if (enc_len > 3) id |= (uint32_t)p[3] << 24;
      h = (h ^ id) * 1099511628211ULL;
    }
    c += h ^ ((uint64_t)s * 0x9E3779B97F4A7C15ULL);
  }
  return c;
}

/*
 * tm_oracle.h -- CPU oracle for the TokenMonster tokenize / detokenize hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it is
 * imported / linked only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, and only as the checker.
 *
 * This is a plain-C restatement (not a copy) of the reference algorithm:
 *   go/tokenmonster.go:242-253   normalize + applyCapcode
 *   go/tokenmonster.go:1017-1279 Vocab.tokenize (and its 4 emit variants :1281-2359)
 *   go/tokenmonster.go:788-948   Decode / DecodeSerialized
 *   go/tokenmonster.go:701-740   Decoder.Decode (streaming)
 *   go/tokenmonster.go:2656-2736 Load (.vocab binary format)
 *   javascript/tokenmonster.js:900-1065 capcode encode / decoder (the only in-repo
 *       restatement of github.com/alasdairforsythe/capcode, which is NOT vendored)
 *   pansearch.Fast.LongestSubstring (third-party, not vendored; semantics taken from
 *       tokenmonster-cpp/src/tokenmonster.cpp:786-877 = "longest key that prefixes the
 *       window", result index = rank in (length, bytes) order = record order in the file)
 *
 * Pinning: tests/test_oracle.py checks this file against (1) every assertion of the
 * reference's tests/unit.cpp:87-119 and (2) oracle/_ref (the reference's own C++ runtime
 * compiled unmodified from /root/reference) on randomized vocabularies and inputs.
 * Unpinned (no golden vectors exist anywhere in the reference): capcode byte output vs
 * the Go capcode package, non-ASCII normalization.  See DESIGN.md "Oracle".
 */
#ifndef TM_ORACLE_H
#define TM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMO_DNE 0xFFFFFFu /* DOES_NOT_EXIST, go/tokenmonster.go:32 */

/* status codes */
#define TMO_OK 0
#define TMO_ERR_FORMAT 1      /* not a valid vocabulary */
#define TMO_ERR_ORDER 2       /* records not sorted by (length, bytes) -- cpp:1352-1357 */
#define TMO_ERR_NORMALIZE 13  /* normalization failed / unsupported (server status 13) */
#define TMO_ERR_ARG 3

typedef struct tmo_vocab tmo_vocab;

typedef struct tmo_capcode_decoder {
  uint8_t in_word, in_char, del, ignore;
} tmo_capcode_decoder;

typedef struct tmo_decoder {
  const tmo_vocab* vocab;
  uint8_t* remainder;
  size_t remainder_len;
  tmo_capcode_decoder cc;
} tmo_decoder;

/* ---- vocabulary ---- */
tmo_vocab* tmo_vocab_load(const uint8_t* file, size_t n, int* status);
void tmo_vocab_free(tmo_vocab* v);
int tmo_vocab_info(const tmo_vocab* v, uint32_t out[12]);
/* out: capcode, charset, normflag, level, reserve, unk, vocab_size, n_reverse, n_info,
 *      delete_token, max_token_length, 0 */

/* ---- buffers returned by the oracle are malloc'd; release with tmo_free ---- */
void tmo_free(void* p);

/* ---- normalize: norm flag then capcode (go:242-253) ---- */
int tmo_normalize(const tmo_vocab* v, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n);

/* ---- capcode level 2 (js:900-1065). ASCII classes only in this build. ---- */
int tmo_capcode_encode(const uint8_t* in, size_t n, uint8_t** out, size_t* out_n);
/* stateful decode; out buffer must hold n bytes; returns bytes written */
size_t tmo_capcode_decode(tmo_capcode_decoder* st, const uint8_t* in, size_t n, uint8_t* out);

/* ---- tokenize.  mode: 0 = ids (u32), 2/3/4 = serialized LE bytes, 1 = count only ----
 * normalized != 0: input is already normalized (Vocab.tokenize);
 * normalized == 0: Vocab.Tokenize / Count / TokenizeToSerialized.
 * n_tokens receives the number of ids emitted (mode 1: the reference's Count value,
 * which counts 1 for the branches that emit 2 ids -- go:1281-1543 vs :1017-1279). */
int tmo_tokenize(const tmo_vocab* v, const uint8_t* in, size_t n, int normalized, int mode,
                 void** out, size_t* n_tokens, int64_t* missing);

/* ---- decode ---- */
/* raw != 0 skips the capcode decode (Vocab.decode, go:813). enc_len 0 => u32 ids. */
/* training worker loop (training/trainvocab.go:932-1170): scores[id] += characters covered,
 * *tokens_in_text += tokens, miss_hist[byte] += bytes without a token; all three accumulate */
int tmo_score(const tmo_vocab* v, const uint8_t* in, size_t n, int normalized, uint64_t* scores,
              uint64_t* miss_hist, uint64_t* tokens_in_text);

/* diagnostics (not thread safe): look-ahead searches {plain, retry, plain repeats of one of the
 * previous 8 positions, steps} since the last reset */
void tmo_debug_walk_stats(uint64_t out[4], int reset);

int tmo_decode(const tmo_vocab* v, const void* tokens, size_t n_tokens_or_bytes, int enc_len,
               int raw, uint8_t** out, size_t* out_n);

/* ---- streaming decoder (go:450-740) ---- */
tmo_decoder* tmo_decoder_new(const tmo_vocab* v);
void tmo_decoder_free(tmo_decoder* d);
int tmo_decoder_decode(tmo_decoder* d, const void* tokens, size_t n_tokens_or_bytes, int enc_len,
                       uint8_t** out, size_t* out_n);
int tmo_decoder_flush(tmo_decoder* d, uint8_t** out, size_t* out_n);

/* helpers exposed for tests (go:149-219) */
int tmo_incomplete_utf8(const uint8_t* b, size_t n);
int tmo_incomplete_utf16(const uint8_t* b, size_t n);

/* longest-prefix match over the dictionary, exposed for table tests:
 * returns length (0 = not found), *index = dictionary index */
int tmo_longest_prefix(const tmo_vocab* v, const uint8_t* w, size_t n, uint32_t* index);

/* checksum of a batch result (same fold as oracle/ref_driver.cpp tmr_tokenize_batch) */
uint64_t tmo_ids_checksum(const uint8_t* tokens, int enc_len, const uint64_t* tok_off, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif

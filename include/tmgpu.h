/*
 * tmgpu.h -- C ABI of the B200 batched tokenize / detokenize library (libtmgpu.so).
 *
 * This is the drop-in boundary for TokenMonster's hot path.  The reference has no FFI
 * boundary of its own; every entry point below names the reference
 * interface it stands in for, and INTEGRATION.md shows the cgo stub that binds it from
 * go/tokenmonster.go and the patch to training/tokenmonsterserver.go.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA / torch types in any signature
 *     (streams are passed as void* = cudaStream_t; NULL = the CUDA default stream)
 *   - every function returns an int status: 0 = OK, otherwise a TMGPU_E_* code;
 *     tmgpu_last_error() gives the message for the calling thread
 *   - per-string failures do not fail the batch: see status[] (the reference server only
 *     reports results[0].err, training/tokenmonsterserver.go:379-381)
 *   - inputs are read-only for the duration of the call (unlike the Go API the library never
 *     mutates caller memory, go/README.md:32); cgo rules: one flat blob + offsets, no Go
 *     pointer is retained after return
 *   - outputs of the host-buffer calls live in pinned host memory owned by the library and
 *     are released with tmgpu_result_free()
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns
 *     TMGPU_E_CUDA
 */
#ifndef TMGPU_H
#define TMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMGPU_DNE 0xFFFFFFu /* DOES_NOT_EXIST, go/tokenmonster.go:32 */

/* library-level status codes (numbers 12/13 match training/tokenmonsterserver.go:38-50) */
#define TMGPU_OK 0
#define TMGPU_E_ARG 1           /* bad argument (enc_len not 2/3/4, null pointer, ...) */
#define TMGPU_E_CUDA 2          /* CUDA runtime error or no device: message has the detail */
#define TMGPU_E_NOMEM 3
#define TMGPU_E_VOCAB_OPEN 12   /* cannot open vocabulary file */
#define TMGPU_E_VOCAB_FORMAT 14 /* not a valid TokenMonster vocabulary (go:2675,2696,2731) */
#define TMGPU_E_VOCAB_ORDER 15  /* records not sorted by (length, bytes) (cpp:1352-1357) */
#define TMGPU_E_UNSUPPORTED 17  /* capcode level 1 / UTF-16 capcode / unsupported norm flag */

/* per-string status[] values */
#define TMGPU_S_OK 0
#define TMGPU_S_NORMALIZE 13 /* normalization failed (server status 13, go:963-966); the result for
                                that string is empty.  Raised for: invalid UTF-8 or more than 30
                                consecutive combining marks under a normalizing UTF-8 vocabulary;
                                any byte >= 0x80 under a normalizing vocabulary whose charset is
                                not UTF-8.  Non-ASCII text is otherwise normalized on the device
                                (NFD, accents, lowercase, capcode; Unicode 15.0 tables). */

typedef struct tmgpu_vocab tmgpu_vocab;

typedef struct tmgpu_vocab_info {
  uint8_t capcode;       /* 0, 1, 2   (Vocab.Capcode, go:2377-2597 getters) */
  uint8_t charset;       /* 0 none, 1 UTF-8, 2 UTF-16 */
  uint8_t normflag;      /* norm.Normalizer.Flag */
  uint8_t level;         /* mode */
  uint8_t reserve;
  uint8_t max_token_length;
  uint8_t pad[2];
  uint32_t unk_token;    /* id or TMGPU_DNE */
  uint32_t delete_token; /* id or TMGPU_DNE */
  uint32_t vocab_size;   /* Vocab.Len */
  uint32_t n_reverse;    /* highest id + 1 */
  uint32_t n_info;       /* dictionary entries incl. "D "-duplicates */
  uint32_t n_trie_nodes; /* device prefix-trie size */
} tmgpu_vocab_info;

/* Result of a host-buffer tokenize call.  All arrays are pinned host memory. */
typedef struct tmgpu_result {
  uint32_t n;         /* strings */
  uint8_t enc_len;    /* bytes per id in `tokens`: 2, 3 or 4 (little endian) */
  uint64_t n_tokens;  /* total ids */
  uint8_t* tokens;    /* n_tokens * enc_len bytes, strings back to back */
  uint64_t* tok_off;  /* n + 1 offsets, in ids */
  uint32_t* missing;  /* n: characters with no token (2nd return of Vocab.Tokenize) */
  uint8_t* status;    /* n: TMGPU_S_* */
  void* priv;
} tmgpu_result;

/* Result of a host-buffer decode call. */
typedef struct tmgpu_bytes {
  uint32_t n;
  uint64_t n_bytes;
  uint8_t* bytes;     /* n_bytes, strings back to back */
  uint64_t* off;      /* n + 1 byte offsets */
  void* priv;
} tmgpu_bytes;

/* ---- library ------------------------------------------------------------------------ */
const char* tmgpu_last_error(void);
int tmgpu_device_count(void); /* number of CUDA devices, 0 if none */
const char* tmgpu_version(void);

/* ---- vocabulary: replaces Load (go/tokenmonster.go:2656-2736; C++ Vocab::load cpp:1287) --
 * Parses the .vocab binary layout, builds the device tables
 * (prefix trie, 32-byte token records, reverse table) and uploads them to `device`.
 * The vocabulary is immutable afterwards; concurrent batch calls on one vocab are
 * serialised on its stream.  After a mutation on the Go side (PrivateGenerateVocab
 * callers, go:2777-2851) destroy and re-create it. */
int tmgpu_vocab_create(const uint8_t* file_bytes, size_t n, int device, tmgpu_vocab** out);
int tmgpu_vocab_load_file(const char* path, int device, tmgpu_vocab** out);
void tmgpu_vocab_destroy(tmgpu_vocab* v);
int tmgpu_vocab_info_get(const tmgpu_vocab* v, tmgpu_vocab_info* out);

/* ---- batched Vocab.Tokenize / TokenizeToSerialized (go:959-968, :986-1014) --------------
 * data: all strings back to back; off[n+1]: byte offsets (off[0] may be non-zero).
 * enc_len 2, 3 or 4 selects the serialized width (the server uses 2 or 4,
 * training/tokenmonsterserver.go:350-353); 0 picks 2 if n_reverse <= 65536 else 3 (go:990-996).
 * This is the call training/tokenmonsterserver.go:363-377 (job 1, goroutine per item) makes
 * instead of spawning goroutines. */
int tmgpu_tokenize_batch(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n,
                         uint8_t enc_len, tmgpu_result** out);
void tmgpu_result_free(tmgpu_result* r);

/* ---- batched Vocab.Count (go:971-980; server job 20, :773-787) ---------------------------
 * counts[n] receive the reference's Count value (which counts 1 for the branches that emit
 * two ids, go:1281-1543 -- SURVEY quirk Q2, replicated). missing/status may be NULL. */
int tmgpu_count_batch(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n,
                      uint64_t* counts, uint32_t* missing, uint8_t* status);

/* ---- training worker loop (training/trainvocab.go:932-1170) ---------------------------------
 * What `worker()` computes for one candidate vocabulary over its data strips: the same branch
 * decisions as Vocab.tokenize (the two loops are textual copies), accumulated instead of emitted:
 *   scores[id] += characters the chosen token covers (scores[deleteToken]++ for a retry branch,
 *                 :1109-1162); n_reverse entries, overwritten
 *   *tokens_in_text = tokens used (a retry branch counts 2, a byte without a token 1, :1164-1171)
 *   miss_hist[b]    = how often byte value b had no token (the worker's missingList, as a
 *                     histogram); 256 entries, may be NULL
 * The candidate is passed as a vocabulary (tmgpu_vocab_create on the .vocab layout of its
 * tokenInfo table); strings are the data strips.  status[] as for tmgpu_tokenize_batch (NULL ok). */
int tmgpu_score_batch(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n,
                      uint64_t* scores, uint64_t* miss_hist, uint64_t* tokens_in_text, uint8_t* status);

/* ---- batched Vocab.Decode / DecodeSerialized (go:788-811; server jobs 2-4, :420-433) -----
 * tokens: ids packed enc_len (2/3/4) bytes each, strings back to back; tok_off[n+1] in ids.
 * Out-of-range ids and nil reverse entries decode to nothing (go:813-834). */
int tmgpu_decode_batch(tmgpu_vocab* v, const uint8_t* tokens, const uint64_t* tok_off, uint32_t n,
                       uint8_t enc_len, tmgpu_bytes** out);
void tmgpu_bytes_free(tmgpu_bytes* b);

/* ---- streaming decode: Vocab.NewDecoder / Decoder.Decode / DecodeSerialized / Flush ----------
 *      (go/tokenmonster.go:450-740; python Vocab.decoder(), py:273) for a BATCH of n streams.
 * Stream s consumes ids [tok_off[s], tok_off[s+1]) of this call; its incomplete UTF-8 / UTF-16
 * tail (go:149-219) and its capcode decoder flags persist on the device until the next call. */
typedef struct tmgpu_decoder tmgpu_decoder;
int tmgpu_decoder_create(tmgpu_vocab* v, uint32_t n_streams, tmgpu_decoder** out);
void tmgpu_decoder_destroy(tmgpu_decoder* d);
int tmgpu_decoder_decode(tmgpu_decoder* d, const uint8_t* tokens, const uint64_t* tok_off,
                         uint8_t enc_len, tmgpu_bytes** out);
int tmgpu_decoder_flush(tmgpu_decoder* d, tmgpu_bytes** out);

/* ---- server wire format: the payloads of tokenmonsterserver's jobs, taken and produced as-is ----
 *      (training/tokenmonsterserver.go:189-209 protocol comment; python client py:341-552)
 * Request payload of jobs 1, 2-4 and 20:  [u32 n] then n x ([u64 len][len bytes]), little endian.
 * The request goes to the GPU verbatim (no host repack; the kernel skips the 8-byte headers) and
 * the reply payload is framed on the device, so the server writes `payload` straight to stdout
 * behind its 9-byte header {status_code, u64 len}.
 *   tmgpu_wire_tokenize  job 1 (server:339-394): reply [u32 n] n x ([u64 bytes][ids]); ids are 2
 *                        bytes each when the vocabulary has <= 65536 ids, else 4 (server:351-354,
 *                        the width python's deserialize_tokens expects).  A string that fails
 *                        normalization yields an empty entry; status_code is 13 only if string 0
 *                        failed, like the server (:379-381), else 0.
 *   tmgpu_wire_count     job 20 as python's tokenize_count sends it (py:485-552): reply [u32 n]
 *                        n x [u64 count].  (The reference server has no case 20 -- SURVEY quirk Q3.)
 *   tmgpu_wire_decode    jobs 2/3/4 (server:396-446), enc_len = the job number: entries hold
 *                        serialized ids; reply [u32 n] n x ([u64 len][decoded bytes]).  Every
 *                        entry is decoded from its own body (the reference's multi-entry branch
 *                        passes the wrong slice, server:429). */
typedef struct tmgpu_wire_reply {
  uint8_t status_code; /* first byte of the server's response header */
  uint64_t len;        /* payload bytes */
  uint8_t* payload;    /* pinned host memory owned by the library */
  void* priv;
} tmgpu_wire_reply;
int tmgpu_wire_tokenize(tmgpu_vocab* v, const uint8_t* payload, uint64_t len, tmgpu_wire_reply** out);
int tmgpu_wire_count(tmgpu_vocab* v, const uint8_t* payload, uint64_t len, tmgpu_wire_reply** out);
int tmgpu_wire_decode(tmgpu_vocab* v, uint8_t enc_len, const uint8_t* payload, uint64_t len,
                      tmgpu_wire_reply** out);
void tmgpu_wire_reply_free(tmgpu_wire_reply* r);

/* Deserialize (go/tokenmonster.go:749-784): little-endian ids of enc_len (2, 3, 4; 0 = by the
 * vocabulary, 2 if len(reverse) <= 65536 else 3) bytes -> uint32.  `out` has room for
 * n_bytes / enc_len ids; a ragged tail is ignored.  Host-side byte shuffling, no device work;
 * v may be NULL when enc_len != 0.  Returns the number of ids, or -1 for an invalid enc_len. */
int64_t tmgpu_deserialize(const tmgpu_vocab* v, const uint8_t* data, uint64_t n_bytes, uint8_t enc_len,
                          uint32_t* out);

/* ---- device-resident path (inputs already in HBM; used for the kernel roofline figure and
 *      by the sharded multi-GPU host) -------------------------------------------------------
 * A workspace owns the device output buffers for batches up to (max_strings, max_bytes).
 * d_data needs no particular alignment, but the 16-byte block around its first byte must be
 * readable (true for any pointer into a cudaMalloc / torch allocation). */
typedef struct tmgpu_workspace tmgpu_workspace;

typedef struct tmgpu_device_view {
  const uint8_t* tokens;    /* device: dense ids, enc_len bytes each */
  const uint64_t* tok_off;  /* device: n + 1 */
  const uint32_t* missing;  /* device: n */
  const uint8_t* status;    /* device: n */
  const uint64_t* n_tokens; /* device: 1 element (valid after the stream is synchronised) */
  const uint32_t* counts;   /* device: n (ids per string; Count value in count mode) */
} tmgpu_device_view;

int tmgpu_workspace_create(tmgpu_vocab* v, uint32_t max_strings, uint64_t max_bytes,
                           tmgpu_workspace** out);
void tmgpu_workspace_destroy(tmgpu_workspace* ws);
/* Launches the tokenize kernels on `stream` (void* = cudaStream_t; NULL = default stream).
 * mode: 0 = ids (enc_len 2/3/4), 1 = count only.  Returns with the scan / pack kernels still in
 * flight; under a vocabulary that normalizes UTF-8 text the call waits once, after the tokenize
 * kernel, for a 16-byte read that says whether any string needs the non-ASCII pass. */
int tmgpu_tokenize_device(tmgpu_workspace* ws, const uint8_t* d_data, const int64_t* d_off,
                          uint32_t n, uint8_t enc_len, int mode, void* stream,
                          tmgpu_device_view* view);
/* Durations (ms) of the stages of the most recent tokenize run on this workspace, from CUDA events
 * on the launching stream: out[0] prep + slot scan + length sort, out[1] capcode encoder +
 * position table, out[2] chain, out[3] 0.  Returns 4, or 0 before the first run. */
int tmgpu_workspace_stage_times(tmgpu_workspace* ws, float out[4]);
/* Number of kernels the last tmgpu_tokenize_device / decode call launched (bench accounting). */
int tmgpu_last_launch_count(const tmgpu_workspace* ws);
/* What the most recent tokenize / count call did besides the plain launch: out[0] strings handled by
 * the non-ASCII normalizer, out[1] chunks launched for long documents (a batch too small to fill
 * the GPU has its documents above 4 KB cut into ~1 KB speculative chunks that are stitched
 * exactly), out[2] chunked documents that had to be redone by one lane,
 * out[3] kernels launched.  ws == NULL reports the vocabulary's internal workspace (host calls). */
int tmgpu_last_stats(const tmgpu_vocab* v, const tmgpu_workspace* ws, uint32_t out[4]);
/* Device durations (ms, CUDA events on the launching stream) of the tokenize kernel for the
 * launches since the previous call, oldest first (ring of 64).  Synchronises on the last one. */
int tmgpu_workspace_kernel_times(tmgpu_workspace* ws, float* ms_out, int max_out);
/* Device decode: d_tokens packed enc_len bytes per id, d_tok_off[n+1] in ids.  Output bytes and
 * offsets are workspace-owned device buffers. */
int tmgpu_decode_device(tmgpu_workspace* ws, const uint8_t* d_tokens, const int64_t* d_tok_off,
                        uint32_t n, uint8_t enc_len, void* stream, const uint8_t** d_bytes,
                        const uint64_t** d_off, const uint64_t** d_total);

#ifdef __cplusplus
}
#endif
#endif /* TMGPU_H */

// kernels.h -- launcher interface between the C-ABI host code (abi*.cu) and the device code.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "device_tables.h"
#include "unicode_dev.h"

namespace tmgpu {

// ---- long documents cut into speculative chunks (chunking.cuh) ---------------------------------
struct ChunkItem {
  int64_t start;     // byte offset of the chunk's first byte in `data`
  int64_t next;      // start of the document's next chunk (document end for the last one)
  uint32_t doc;      // string index
  uint32_t flags;    // bit 0: last chunk of its document
};
struct ChunkOut {
  uint32_t cnt;      // ids logged (capped at the chunk's capacity)
  uint32_t n_norm;   // normalized bytes between this chunk's start and the next chunk's start
  uint32_t miss;     // characters without a token over the whole run
  uint32_t flags;    // 1 non-ASCII seen, 2 log overflow, 4 ran to the end of the document
};

struct TokenizeArgs {
  VocabDev v;
  const uint8_t* data;       // all strings back to back (16-byte aligned base)
  const int64_t* off;        // n + 1 byte offsets
  const int64_t* ends;       // optional n end offsets when strings are not back to back (the server's
                             // wire payload keeps an 8-byte length header in front of each string)
  uint32_t n;
  uint32_t* slab;            // u32 ids; string s at (off[s]-off[0]) + slab_pad*s unless slab_off
  uint32_t slab_pad;         // extra capacity per string (ids)
  const uint64_t* slab_off;  // optional exact offsets (overflow re-run), n + 1 entries
  uint32_t* counts;          // n: ids stored, capped at the slab share (the Count value when !STORE)
  uint32_t* need;            // n: uncapped id count (sizes the overflow re-run)
  uint32_t* missing;         // n
  uint8_t* status;           // n
  uint32_t* work_counter;    // zeroed before launch
  uint32_t* overflow;        // set to 1 if any string exceeded its slab capacity
  // second pass over strings the device normalizer rewrote (unicode_normalize.cuh): work item k
  // reads its text at off[k]..ends[k] and reports as string list[k]
  const uint32_t* list;      // optional: work item -> string index
  const int64_t* raw_off;    // with list: the strings' own offsets (slab placement and capacity)
  const int64_t* raw_ends;
  uint32_t prenorm;          // text is already normalized: no ASCII check, no case folding, CAP = 0
  const uint8_t* skip;       // optional: strings another launch handles (documents tokenized in chunks)
  // training worker loop (training/trainvocab.go:932-1170), count kernels only: characters covered
  // per token id, and occurrences of each byte value that no token covers
  unsigned long long* scores;
  unsigned long long* miss_hist;
  // chunk launch (tokenize_thread_kernel<.., CHUNK = true>): n = number of chunks
  const ChunkItem* chunks;
  uint32_t* c_pos;           // per logged id: normalized position | forward-delete << 31 of its step
  uint32_t* c_miss;          // per logged id: characters without a token before its step
  ChunkOut* c_out;
  uint32_t c_cap, c_ov;      // log capacity per chunk; how far a chunk runs past the next one's start
  const unsigned long long* watermark;  // (chunk kernel only) unused
  // position-table pipeline scratch, per launch, indexed by work item
  uint8_t* text;             // training scores only: the normalized bytes, laid out like `words`
  uint32_t* words;           // one word per normalized position (pipeline_core.h), right aligned in the slot
  uint64_t words_cap;        // entries
  uint32_t* nlen;            // normalized length
  uint32_t* npad;            // n + 1: slot size in words (0 for refused / skipped items)
  uint64_t* toff;            // n + 1: exclusive scan of npad; the slot of item i ends at toff[i + 1]
  uint32_t* sort_key;        // length / 8 (0: nothing to do)
  uint32_t* sort_val;        // item
  const uint32_t* sort_order;  // items by descending length: a warp works on 32 strings of equal length
  uint8_t* pflag;            // 0 run, 1 refused (status 13), 2 nothing to do, 3 slot too small (status 14)
  uint32_t* pipe_counters;   // 4 zeroed work counters (2: table, 3: chain)
};

size_t scan_temp_bytes(uint32_t n_items);
cudaError_t exclusive_scan_u32_to_u64(void* temp, size_t temp_bytes, const uint32_t* in, uint64_t* out,
                                      uint32_t n_items, cudaStream_t st);
// the position-table pipeline (pipeline_kernels.cuh); enc = the vocabulary uses capcode 2 and the
// text is raw (not pre-normalized); store = ids (K1) vs count only (K3)
cudaError_t launch_prep(const TokenizeArgs& a, bool cap, bool slot_full, int sms, cudaStream_t st);
size_t sort_temp_bytes(uint32_t n_items);
cudaError_t sort_items_by_length(void* temp, size_t temp_bytes, const uint32_t* key_in, uint32_t* key_out,
                                 const uint32_t* val_in, uint32_t* val_out, uint32_t n_items, cudaStream_t st);
cudaError_t launch_table(const TokenizeArgs& a, bool cap, bool text, int sms, cudaStream_t st);
cudaError_t launch_select_overflowed(const uint8_t* status, const int64_t* off, const int64_t* ends, uint32_t n,
                                     uint32_t* list, int64_t* o_begin, int64_t* o_end, unsigned long long* ctr, int sms,
                                     cudaStream_t st);
cudaError_t launch_chain(const TokenizeArgs& a, bool store, int sms, cudaStream_t st);
cudaError_t launch_pack(const uint32_t* slab, const int64_t* off, uint32_t slab_pad, const uint64_t* slab_off,
                        const uint32_t* counts, const uint64_t* tok_off, uint8_t* out, uint32_t n, int enc,
                        int sms, cudaStream_t st);
cudaError_t launch_chunk_offsets(uint64_t* tok_off, const uint32_t* counts, uint32_t n, unsigned long long* base_io,
                                 bool write_end, int sms, cudaStream_t st);
// reply framing of the server protocol: [u32 n] then per string [u64 byte length][unit * items]
cudaError_t launch_wire_frame(const uint8_t* src, const uint64_t* src_off, uint32_t unit, uint32_t n, uint8_t* dst,
                              int sms, cudaStream_t st);
// long documents in chunks (chunking.cuh)
struct StitchArgs {
  const int64_t* off;            // documents
  const int64_t* ends;
  uint32_t n;
  const uint32_t* n_chunks;      // per document
  const uint64_t* chunk_base;    // exclusive scan of n_chunks
  const ChunkOut* c_out;
  const uint32_t* c_pos;
  const uint32_t* c_miss;
  uint32_t c_cap;
  uint32_t slab_pad;
  uint32_t count_mode;           // Vocab.Count: the delete token of a two-id step is not counted
  uint8_t* doc_fail;             // in: planning failed; out: 1 = tokenize this document with one lane
  // per chunk: the kept range of its log and where it goes in the document's slab slot
  uint32_t* meet_a;              // step of this chunk's log where it meets its successor's ...
  uint32_t* meet_b;              // ... and the successor's step there
  uint32_t* seg_begin;
  uint32_t* seg_end;
  uint32_t* seg_dst;
  uint32_t* seg_dels;            // delete-token entries in the kept range (not counted by Vocab.Count)
  uint32_t* seg_miss;            // characters without a token in the kept range
  uint32_t* seg_flag;            // 1 non-ASCII, 2 cannot stitch, 4 document ends with this chunk
  // per document
  uint32_t* counts;
  uint32_t* need;
  uint32_t* missing;
  uint8_t* status;
  uint32_t* overflow;
};

uint32_t chunk_log_capacity();
uint32_t chunk_overlap();
cudaError_t launch_chunk_count(const int64_t* off, const int64_t* ends, uint32_t n, uint32_t* n_chunks, uint8_t* skip,
                               int sms, cudaStream_t st);
cudaError_t launch_chunk_fill(const uint8_t* data, const int64_t* off, const int64_t* ends, uint32_t n,
                              const uint32_t* n_chunks, const uint64_t* chunk_base, uint32_t total, ChunkItem* items,
                              uint8_t* doc_fail, int sms, cudaStream_t st);
cudaError_t launch_tokenize_chunks(const TokenizeArgs& a, int sms, cudaStream_t st);
cudaError_t launch_chunk_stitch(const StitchArgs& a, uint32_t total_chunks, const ChunkItem* items, int sms,
                                cudaStream_t st);
cudaError_t launch_chunk_gather(const ChunkItem* items, uint32_t total, const uint32_t* c_ids, uint32_t c_cap,
                                const uint32_t* seg_begin, const uint32_t* seg_end, const uint32_t* seg_dst,
                                const int64_t* off, uint32_t slab_pad, uint32_t* slab, int sms, cudaStream_t st);
cudaError_t launch_chunk_failed(const uint8_t* doc_fail, const uint32_t* n_chunks, const int64_t* off,
                                const int64_t* ends, uint32_t n, uint32_t* list, int64_t* f_begin, int64_t* f_end,
                                unsigned long long* ctr, int sms, cudaStream_t st);
// strings with a byte >= 0x80: status 13, counts 0, skip 1 (so that a pass with side effects --
// the score histogram -- never starts on them; the non-ASCII pass then takes them whole)
cudaError_t launch_flag_non_ascii(const uint8_t* data, const int64_t* off, const int64_t* ends, uint32_t n,
                                  uint8_t* skip, uint8_t* status, uint32_t* counts, uint32_t* missing, int sms,
                                  cudaStream_t st);
cudaError_t launch_select_refused(const uint8_t* status, const int64_t* off, const int64_t* ends, uint32_t n,
                                  uint32_t* list, int64_t* n_begin, unsigned long long* ctr, int sms,
                                  cudaStream_t st);
cudaError_t launch_unicode_normalize(const NormalizeArgs& a, int sms, cudaStream_t st);
cudaError_t launch_decode_size(const VocabDev& v, const uint8_t* tokens, const int64_t* tok_off, uint32_t n, int enc,
                               uint32_t* raw_len, int sms, cudaStream_t st, const uint64_t* prefix_off = nullptr);
cudaError_t launch_decode_copy(const VocabDev& v, const uint8_t* tokens, const int64_t* tok_off, uint32_t n, int enc,
                               const uint64_t* raw_off, uint8_t* out, int sms, cudaStream_t st,
                               const uint64_t* prefix_off = nullptr, const uint8_t* prefix = nullptr);
// scratch for ordering work items by length (a warp wants 32 strings of equal length)
cudaError_t launch_capcode_decode(const uint8_t* src, const uint64_t* raw_off, uint32_t n, uint8_t* dst,
                                  uint32_t* out_len, int sms, cudaStream_t st, const uint32_t* len_in,
                                  uint8_t* state, const UnicodeDev* uni, uint32_t* counter /* one work counter, zeroed by the launcher */);
cudaError_t launch_stream_tail(const uint8_t* raw, const uint64_t* raw_off, uint32_t n, int charset,
                               uint32_t* keep_len, uint32_t* rem_len, int sms, cudaStream_t st);
cudaError_t launch_copy_tail(const uint8_t* raw, const uint64_t* raw_off, const uint32_t* keep_len,
                             const uint64_t* rem_off, uint32_t n, uint8_t* rem, int sms, cudaStream_t st);
cudaError_t launch_compact_bytes(const uint8_t* src, const uint64_t* src_off, const uint32_t* len,
                                 const uint64_t* dst_off, uint32_t n, uint8_t* dst, int sms, cudaStream_t st,
                                 int wide = 0);  // wide: segment s sits at src_off[s] + (src_off[s]+1)/2

}  // namespace tmgpu

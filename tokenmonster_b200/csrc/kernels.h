// kernels.h -- launcher interface between the C-ABI host code (abi.cu) and the device code.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "device_tables.h"
#include "unicode_dev.h"

namespace tmgpu {

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
  const unsigned long long* watermark;  // optional: input bytes that have landed in `data` so far (host path
                             // streams the batch in while the kernel runs); NULL = everything is there
};

size_t scan_temp_bytes(uint32_t n_items);
cudaError_t exclusive_scan_u32_to_u64(void* temp, size_t temp_bytes, const uint32_t* in, uint64_t* out,
                                      uint32_t n_items, cudaStream_t st);
// group = lanes per string (32, 16 or 8); store = ids (K1) vs count only (K3)
cudaError_t launch_tokenize(const TokenizeArgs& a, int group, bool store, int sms, cudaStream_t st, int* grid_out);
cudaError_t launch_pack(const uint32_t* slab, const int64_t* off, uint32_t slab_pad, const uint64_t* slab_off,
                        const uint32_t* counts, const uint64_t* tok_off, uint8_t* out, uint32_t n, int enc,
                        int sms, cudaStream_t st);
// reply framing of the server protocol: [u32 n] then per string [u64 byte length][unit * items]
cudaError_t launch_wire_frame(const uint8_t* src, const uint64_t* src_off, uint32_t unit, uint32_t n, uint8_t* dst,
                              int sms, cudaStream_t st);
cudaError_t launch_select_refused(const uint8_t* status, const int64_t* off, const int64_t* ends, uint32_t n,
                                  uint32_t* list, int64_t* n_begin, unsigned long long* ctr, int sms,
                                  cudaStream_t st);
cudaError_t launch_unicode_normalize(const NormalizeArgs& a, int sms, cudaStream_t st);
cudaError_t launch_decode_size(const VocabDev& v, const uint8_t* tokens, const int64_t* tok_off, uint32_t n, int enc,
                               uint32_t* raw_len, int sms, cudaStream_t st, const uint64_t* prefix_off = nullptr);
cudaError_t launch_decode_copy(const VocabDev& v, const uint8_t* tokens, const int64_t* tok_off, uint32_t n, int enc,
                               const uint64_t* raw_off, uint8_t* out, int sms, cudaStream_t st,
                               const uint64_t* prefix_off = nullptr, const uint8_t* prefix = nullptr);
cudaError_t launch_capcode_decode(const uint8_t* src, const uint64_t* raw_off, uint32_t n, uint8_t* dst,
                                  uint32_t* out_len, int sms, cudaStream_t st, const uint32_t* len_in = nullptr,
                                  uint8_t* state = nullptr, const UnicodeDev* uni = nullptr);
cudaError_t launch_stream_tail(const uint8_t* raw, const uint64_t* raw_off, uint32_t n, int charset,
                               uint32_t* keep_len, uint32_t* rem_len, int sms, cudaStream_t st);
cudaError_t launch_copy_tail(const uint8_t* raw, const uint64_t* raw_off, const uint32_t* keep_len,
                             const uint64_t* rem_off, uint32_t n, uint8_t* rem, int sms, cudaStream_t st);
cudaError_t launch_compact_bytes(const uint8_t* src, const uint64_t* src_off, const uint32_t* len,
                                 const uint64_t* dst_off, uint32_t n, uint8_t* dst, int sms, cudaStream_t st,
                                 int wide = 0);  // wide: segment s sits at src_off[s] + (src_off[s]+1)/2

}  // namespace tmgpu

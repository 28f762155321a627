// abi_internal.h -- what the translation units behind include/tmgpu.h share: the vocabulary and
// workspace objects, the error string, allocation helpers and the internal entry points that one
// unit offers to the others (abi.cu: library / vocabulary / workspace; abi_tokenize.cu: tokenize,
// count, score; abi_decode.cu: decode and the streaming decoder; abi_wire.cu: server payloads).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/tmgpu.h"
#include "kernels.h"
#include "unicode_dev.h"
#include "unicode_tables.h"
#include "vocab_tables.h"

using namespace tmgpu;

extern thread_local std::string g_err;   // message behind tmgpu_last_error()

#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(e__);                        \
      return TMGPU_E_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// Pinned host blocks handed out as results (cudaMallocHost per call is far too slow).  The pool is
// shared between the vocabulary and every result / bytes / reply object cut from it, so a result
// may outlive its vocabulary: the blocks are freed when the last owner goes.
struct PinnedPool {
  struct Block { void* p; size_t bytes; bool used; };
  std::vector<Block> blocks;
  std::mutex mu;
  bool orphaned = false;   // the vocabulary is gone: blocks are freed as soon as they come back
  ~PinnedPool() {
    for (auto& b : blocks)
      if (b.p) cudaFreeHost(b.p);
  }
};

// Helper threads that copy slices of a pageable host buffer into pinned staging memory (a cgo
// caller passes ordinary Go slices; cudaMemcpyAsync from pageable memory goes through one driver
// thread at a fraction of the PCIe rate).
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; i++) th_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // dst[0..len) = src[0..len), split over the workers and the calling thread
  void copy(uint8_t* dst, const uint8_t* src, size_t len) {
    const size_t parts = th_.size() + 1, slice = (len / parts + 4095) & ~size_t(4095);
    {
      std::lock_guard<std::mutex> l(mu_);
      dst_ = dst; src_ = src; len_ = len; slice_ = slice ? slice : len;
      next_ = 1;                 // part 0 is the caller's
      parts_ = slice_ ? (len + slice_ - 1) / slice_ : 0;
      pending_ = parts_ > 1 ? parts_ - 1 : 0;
      gen_++;
    }
    cv_.notify_all();
    if (len) std::memcpy(dst, src, slice_ < len ? slice_ : len);
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      size_t part;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || (gen_ != seen && next_ < parts_) || (gen_ != seen && next_ >= parts_); });
        if (stop_) return;
        if (next_ >= parts_) { seen = gen_; continue; }
        part = next_++;
      }
      const size_t a = part * slice_, b = a + slice_ < len_ ? a + slice_ : len_;
      std::memcpy(dst_ + a, src_ + a, b - a);
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  uint8_t* dst_ = nullptr;
  const uint8_t* src_ = nullptr;
  size_t len_ = 0, slice_ = 0, parts_ = 0, next_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

struct tmgpu_vocab {
  HostTables host;
  int device = -1;  // -1: host tables only (parser / table self-tests without a GPU)
  int sms = 0;
  VocabDev dev{};
  bool unicode = false;           // non-ASCII strings go through the device normalizer (else status 13)
  UnicodeDev uni{};
  std::vector<void*> allocs;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  tmgpu_workspace* ws = nullptr;  // internal workspace of the host-buffer calls
  uint8_t* d_in = nullptr;        // staging of host inputs
  size_t d_in_cap = 0;
  int64_t* d_off = nullptr;       // d_off_cap begin offsets, then d_off_cap end offsets (wire payloads)
  size_t d_off_cap = 0;
  int64_t* h_off = nullptr;       // pinned, same layout
  size_t h_off_cap = 0;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t d2h_stream = nullptr;
  double ids_per_byte = 0;        // of the previous host-buffer call (sizes the streamed-out id buffer)
  cudaEvent_t chunk_ev[32] = {};
  cudaEvent_t done_ev[32] = {};
  uint64_t* d_run = nullptr;      // device: running id total + per-chunk {base, total}
  uint64_t* h_info = nullptr;     // pinned mirror of the per-chunk {base, total}
  std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
  // pageable inputs: pinned staging ring + helper threads (created on first use)
  static constexpr int kStageSlots = 3;
  static constexpr size_t kStageBytes = size_t(16) << 20;
  uint8_t* stage_buf[kStageSlots] = {};
  cudaEvent_t stage_free[kStageSlots] = {};
  std::unique_ptr<CopyPool> copy_pool;
};

struct tmgpu_workspace {
  tmgpu_vocab* v = nullptr;
  uint32_t max_strings = 0;
  uint64_t max_bytes = 0;
  uint32_t slab_pad = 16;
  uint32_t* slab = nullptr;
  uint64_t slab_ids = 0;
  uint32_t* counts = nullptr;   // n + 1 (last = 0 for the scan)
  uint32_t* need = nullptr;     // n + 1: uncapped counts (overflow re-run)
  uint32_t* missing = nullptr;
  uint8_t* status = nullptr;
  uint64_t* tok_off = nullptr;  // n + 1
  uint8_t* tokens = nullptr;    // dense serialized ids
  uint32_t* work_counter = nullptr;
  uint32_t* overflow = nullptr;
  void* scan_temp = nullptr;
  size_t scan_bytes = 0;
  // position-table pipeline (pipeline_kernels.cuh): normalized text, one word per position, and
  // per work item: normalized length, slot size, slot offset, first word, verdict
  uint8_t* p_text = nullptr;             // training scores only
  uint64_t p_text_cap = 0;
  uint32_t* p_words = nullptr;
  uint64_t p_words_cap = 0;              // entries
  uint32_t* p_nlen = nullptr;            // n + 1
  uint32_t* p_npad = nullptr;            // n + 1
  uint64_t* p_toff = nullptr;            // n + 1
  uint32_t* p_key = nullptr;             // n + 1: sort keys / values in, out
  uint32_t* p_key2 = nullptr;
  uint32_t* p_val = nullptr;
  uint32_t* p_order = nullptr;
  void* p_sort_temp = nullptr;
  size_t p_sort_bytes = 0;
  uint8_t* p_flag = nullptr;             // n + 1
  uint32_t* p_counters = nullptr;        // 4
  // strings whose normalized text did not fit the optimistic slot: redone with full-size slots
  uint32_t* ov_list = nullptr;           // n
  int64_t* ov_begin = nullptr;           // n
  int64_t* ov_end = nullptr;             // n
  unsigned long long* ov_ctr = nullptr;  // {strings, words needed}
  uint32_t last_overflowed = 0;
  // device normalizer (non-ASCII strings): work list, placement, scratch text
  uint32_t* uni_list = nullptr;          // n
  int64_t* uni_begin = nullptr;          // n
  int64_t* uni_end = nullptr;            // n
  uint8_t* uni_text = nullptr;
  uint64_t uni_text_cap = 0;
  uint32_t last_unicode = 0;             // strings the normalizer handled in the most recent launch
  // long documents tokenized in chunks (chunking.cuh): per document ...
  uint32_t* ck_n = nullptr;              // n + 1: chunks of the document (0 = not chunked)
  uint64_t* ck_base = nullptr;           // n + 1: exclusive scan
  uint8_t* ck_skip = nullptr;            // n
  uint8_t* ck_fail = nullptr;            // n
  uint32_t* ck_list = nullptr;           // n: documents to redo with one lane
  int64_t* ck_fb = nullptr;              // n
  int64_t* ck_fe = nullptr;              // n
  unsigned long long* ck_ctr = nullptr;
  // ... and per chunk (grown on demand)
  ChunkItem* ck_items = nullptr;
  ChunkOut* ck_out = nullptr;
  uint32_t *ck_ids = nullptr, *ck_pos = nullptr, *ck_miss = nullptr;   // chunks x log capacity
  uint32_t *ck_sb = nullptr, *ck_se = nullptr, *ck_sd = nullptr;
  uint32_t *ck_ma = nullptr, *ck_mb = nullptr, *ck_dl = nullptr, *ck_sm = nullptr, *ck_fl = nullptr;
  uint64_t ck_chunks_cap = 0;
  std::vector<void*> ck_allocs;
  uint32_t last_chunked = 0, last_chunk_redo = 0;   // chunks launched / documents redone, last launch
  unsigned long long* score_buf = nullptr;          // training scores: n_reverse ids, then 256 byte values
  // decode
  uint32_t* raw_len = nullptr;  // n + 1
  uint64_t* raw_off = nullptr;  // n + 1
  uint32_t* out_len = nullptr;  // n + 1
  uint64_t* out_off = nullptr;  // n + 1
  uint8_t* raw = nullptr;
  uint8_t* dec = nullptr;
  uint8_t* dense = nullptr;
  uint64_t dec_cap = 0;
  std::vector<void*> allocs;
  int last_launches = 0;
  // CUDA events around the dominant kernel (tokenize) of the most recent launches
  static constexpr int kMaxChunks = 32;
  static constexpr int kEv = 64;
  cudaEvent_t ev0[kEv] = {}, ev1[kEv] = {};
  int ev_head = 0, ev_count = 0;
  // stage boundaries of the most recent pipeline run: start, prep + scan + sort, table, chain
  cudaEvent_t stage_ev[4] = {};
  bool stage_valid = false;
};

// Debug aid (TMGPU_POISON=1): every device allocation is filled with 0xA5, so that a kernel that
// reads memory nobody wrote gives the same wrong answer on every run instead of a flaky one.
inline cudaError_t tm_malloc(void** p, size_t bytes) {
  static const bool poison = getenv("TMGPU_POISON") != nullptr;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaSuccess && poison) {
    e = cudaMemset(*p, 0xA5, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  return e;
}
template <class T>
inline cudaError_t tm_malloc(T** p, size_t bytes) {
  return tm_malloc(reinterpret_cast<void**>(p), bytes);
}

template <class T>
inline int dev_alloc(std::vector<void*>& pool, T** out, size_t count) {
  void* p = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = 16;
  cudaError_t e = tm_malloc(&p, bytes + 16);
  if (e != cudaSuccess) {
    g_err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
    return TMGPU_E_NOMEM;
  }
  pool.push_back(p);
  *out = static_cast<T*>(p);
  return TMGPU_OK;
}

// A host-buffer batch is tokenized in chunks of strings so that the H2D copy of chunk k+1 overlaps
// the kernel of chunk k: `ready` is recorded on the copy stream after the chunk's bytes landed.
struct ChunkPlan {
  uint32_t s0, s1;       // strings [s0, s1)
  uint64_t slab_base;    // (off[s0] - off[0]) + slab_pad * s0
  cudaEvent_t ready;     // may be null
  std::function<int()> stage;   // optional: called right before the chunk's kernels are queued (pageable
                                // input: stages + uploads this chunk, so that the host-side memcpy of
                                // chunk k+1 overlaps the kernels of chunk k)
};

// Host destination for ids that leave the GPU chunk by chunk (tokenize_on_stream, host path).
struct StreamOut {
  uint8_t* tokens = nullptr;   // pinned
  uint64_t cap_bytes = 0;
  bool ok = false;             // every chunk fitted and no later pass changed the counts
  uint64_t total_ids = 0;
};

struct ResultPriv {
  std::shared_ptr<PinnedPool> pool;
  std::vector<void*> blocks;
};

// pinned host memory from the vocabulary's pool (cudaMallocHost is far too slow to pay per call)
template <class T>
inline int pinned_alloc(ResultPriv* p, T** out, size_t count) {
  size_t bytes = count * sizeof(T);
  if (bytes < 64) bytes = 64;
  PinnedPool* v = p->pool.get();
  std::lock_guard<std::mutex> lock(v->mu);
  auto& pinned = v->blocks;
  int best = -1;
  for (size_t i = 0; i < pinned.size(); i++) {
    auto& b = pinned[i];
    if (b.p && !b.used && b.bytes >= bytes && (best < 0 || b.bytes < pinned[best].bytes)) best = int(i);
  }
  if (best < 0) {
    void* q = nullptr;
    size_t cap = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&q, cap);
    if (e != cudaSuccess) {
      g_err = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
      return TMGPU_E_NOMEM;
    }
    pinned.push_back({q, cap, false});
    best = int(pinned.size()) - 1;
  }
  pinned[best].used = true;
  p->blocks.push_back(pinned[best].p);
  *out = static_cast<T*>(pinned[best].p);
  return TMGPU_OK;
}

// ---- internal entry points shared between the units -----------------------------------------
void release_priv(ResultPriv* p);
int ensure_host_path(tmgpu_vocab* v, uint32_t n, uint64_t bytes);
int stage_input(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n, uint64_t* total,
                std::vector<ChunkPlan>* plan);
int stage_copy(tmgpu_vocab* v, const uint8_t* src, uint32_t n, uint64_t total, bool gapped,
               std::vector<ChunkPlan>* plan);
int rerun_if_overflow(tmgpu_vocab* v, uint32_t n, uint8_t enc_len, std::vector<void*>* temp, uint8_t** dense,
                      const int64_t* d_ends = nullptr);
int tokenize_on_stream(tmgpu_workspace* ws, const uint8_t* d_data, const int64_t* d_off, uint32_t n,
                       uint8_t enc_len, int mode, cudaStream_t st, const uint64_t* slab_off, uint32_t* slab,
                       uint8_t* dense, const std::vector<ChunkPlan>* chunks = nullptr,
                       const int64_t* d_ends = nullptr, StreamOut* so = nullptr);
int ensure_decode_cap(tmgpu_workspace* ws, uint64_t bytes);
int decode_on_stream(tmgpu_workspace* ws, const uint8_t* d_tokens, const int64_t* d_tok_off, uint32_t n, int enc,
                     cudaStream_t st, const uint8_t** d_bytes, const uint64_t** d_off, uint64_t* total_out);

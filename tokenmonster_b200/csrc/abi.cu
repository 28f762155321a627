// abi.cu -- the C ABI of libtmgpu.so (include/tmgpu.h): vocabulary upload, workspaces, and the
// host-buffer / device-buffer batch calls.  Host orchestration only; all per-byte work is in
// kernels.cu.  There is no CPU compute path here: without a CUDA device every compute entry
// fails with TMGPU_E_CUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tmgpu.h"
#include "kernels.h"
#include "unicode_dev.h"
#include "unicode_tables.h"
#include "vocab_tables.h"

using namespace tmgpu;

static thread_local std::string g_err;

#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(e__);                        \
      return TMGPU_E_CUDA;                                                                \
    }                                                                                     \
  } while (0)

struct tmgpu_vocab {
  HostTables host;
  int device = -1;  // -1: host tables only (parser / table self-tests without a GPU)
  int sms = 0;
  int group = 1;    // 1 = lane-per-string kernel (default); 4/8/16/32 = lanes-over-branches kernel
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
  cudaEvent_t chunk_ev[32] = {};
  cudaEvent_t done_ev[32] = {};
  uint64_t* d_run = nullptr;      // device: running id total + per-chunk {base, total}
  uint64_t* h_info = nullptr;     // pinned mirror of the per-chunk {base, total}
  // pinned host blocks handed out as results and taken back by tmgpu_result_free / bytes_free
  struct PinnedBlock { void* p; size_t bytes; bool used; };
  std::vector<PinnedBlock> pinned;
  std::mutex pin_mu;
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
  // device normalizer (non-ASCII strings): work list, placement, scratch text
  uint32_t* uni_list = nullptr;          // n
  int64_t* uni_begin = nullptr;          // n
  int64_t* uni_end = nullptr;            // n
  unsigned long long* uni_ctr = nullptr; // {strings, scratch bytes}
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
};

extern "C" const char* tmgpu_last_error(void) { return g_err.c_str(); }
extern "C" const char* tmgpu_version(void) { return "tokenmonster-b200 0.1 (sm_100a)"; }

extern "C" int tmgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <class T>
static int dev_alloc(std::vector<void*>& pool, T** out, size_t count) {
  void* p = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&p, bytes + 16);
  if (e != cudaSuccess) {
    g_err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
    return TMGPU_E_NOMEM;
  }
  pool.push_back(p);
  *out = static_cast<T*>(p);
  return TMGPU_OK;
}

template <class T>
static int upload(tmgpu_vocab* v, const T* src, size_t count, const T** out) {
  T* p = nullptr;
  int rc = dev_alloc(v->allocs, &p, count);
  if (rc) return rc;
  if (count) CU(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *out = p;
  return TMGPU_OK;
}

extern "C" int tmgpu_vocab_create(const uint8_t* file_bytes, size_t n, int device, tmgpu_vocab** out) {
  if (!file_bytes || !out) {
    g_err = "tmgpu_vocab_create: null argument";
    return TMGPU_E_ARG;
  }
  *out = nullptr;
  auto* v = new tmgpu_vocab();
  std::string err;
  int rc = parse_vocab(file_bytes, n, &v->host, &err);
  if (rc) {
    g_err = err;
    delete v;
    return rc;
  }
  const HostTables& h = v->host;
  if (h.capcode == 1) {
    g_err = "capcode level 1 (NoCapcodeEncode) has no specification in the reference; unsupported";
    delete v;
    return TMGPU_E_UNSUPPORTED;
  }
  if (h.capcode == 2 && h.charset == 2) {
    g_err = "capcode with a UTF-16 vocabulary is unsupported";
    delete v;
    return TMGPU_E_UNSUPPORTED;
  }
  if (h.normflag & ~7u) {
    g_err = "normalization flags other than nfd/lowercase/accents are unsupported on the device path";
    delete v;
    return TMGPU_E_UNSUPPORTED;
  }
  if (const char* g = getenv("TMGPU_GROUP")) {
    int gi = atoi(g);
    if (gi == 1 || gi == 4 || gi == 8 || gi == 16 || gi == 32) v->group = gi;
  }
  v->device = device;
  if (device < 0) {  // host tables only
    *out = v;
    return TMGPU_OK;
  }
  if (device >= tmgpu_device_count()) {
    g_err = "no CUDA device " + std::to_string(device) + " (this library has no CPU path)";
    delete v;
    return TMGPU_E_CUDA;
  }
  auto fail = [&](int code) {
    tmgpu_vocab_destroy(v);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) {
    g_err = "cudaSetDevice failed";
    return fail(TMGPU_E_CUDA);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    g_err = "cudaGetDeviceProperties failed";
    return fail(TMGPU_E_CUDA);
  }
  v->sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_err = "cudaStreamCreate failed";
    return fail(TMGPU_E_CUDA);
  }
  if (cudaStreamCreateWithFlags(&v->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_err = "cudaStreamCreate failed";
    return fail(TMGPU_E_CUDA);
  }
  for (int k = 0; k < 32; k++) {
    if (cudaEventCreateWithFlags(&v->chunk_ev[k], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&v->done_ev[k], cudaEventDisableTiming) != cudaSuccess) {
      g_err = "cudaEventCreate failed";
      return fail(TMGPU_E_CUDA);
    }
  }
  if (cudaMalloc(&v->d_run, 8 * 80) != cudaSuccess || cudaMallocHost(&v->h_info, 8 * 80) != cudaSuccess) {
    g_err = "allocation of the chunk bookkeeping failed";
    return fail(TMGPU_E_NOMEM);
  }
  VocabDev& d = v->dev;
  std::vector<uint8_t> image(2304);
  std::memcpy(image.data(), h.root_child, 1024);
  std::memcpy(image.data() + 1024, h.begin_byte, 256);
  std::memcpy(image.data() + 1280, h.sp_child, 1024);
  const uint64_t* edges64 = nullptr;
  if ((rc = upload(v, h.edges.data(), h.edges.size(), &edges64))) return fail(rc);
  d.edges = reinterpret_cast<const uint2*>(edges64);
  d.edge_mask = uint32_t(h.edges.size() - 1);
  d.edge_shift = 32u - h.edge_bits;
  const TokenRecord* recs = nullptr;
  if ((rc = upload(v, h.rec.data(), h.rec.size(), &recs))) return fail(rc);
  d.rec = reinterpret_cast<const uint4*>(recs);
  if ((rc = upload(v, image.data(), image.size(), &d.smem_image))) return fail(rc);
  if ((rc = upload(v, h.root2.data(), h.root2.size(), &d.root2))) return fail(rc);
  if ((rc = upload(v, h.rev_off.data(), h.rev_off.size(), &d.rev_off))) return fail(rc);
  if ((rc = upload(v, h.rev_len.data(), h.rev_len.size(), &d.rev_len))) return fail(rc);
  if ((rc = upload(v, h.rev_blob.data(), h.rev_blob.size(), &d.rev_blob))) return fail(rc);
  d.n_info = h.n_info;
  d.n_reverse = h.n_reverse;
  d.unk = h.unk;
  d.delete_token = h.delete_token;
  d.space_start = h.space_start;
  d.max_len = h.max_len;
  d.off = h.charset == 2 ? 2 : 1;
  d.capcode = h.capcode;
  d.normflag = h.normflag;
  d.charset = h.charset;
  // UTF-8 vocabularies that normalize: non-ASCII strings are normalized on the device too
  if (h.charset == 1 && ((h.normflag & 7u) != 0 || h.capcode == 2)) {
    const UnicodeHost& t = unicode_host_tables();
    UnicodeDev& u = v->uni;
    if ((rc = upload(v, t.stage1, t.n_stage1, &u.stage1)) || (rc = upload(v, t.stage2, t.n_stage2, &u.stage2)) ||
        (rc = upload(v, t.dkey, t.n_decomp, &u.dkey)) || (rc = upload(v, t.didx, t.n_decomp + 1, &u.didx)) ||
        (rc = upload(v, t.ddata, t.n_ddata, &u.ddata)) || (rc = upload(v, t.lkey, t.n_lower, &u.lkey)) ||
        (rc = upload(v, t.lval, t.n_lower, &u.lval)) || (rc = upload(v, t.ukey, t.n_upper, &u.ukey)) ||
        (rc = upload(v, t.uval, t.n_upper, &u.uval)))
      return fail(rc);
    u.n_decomp = uint32_t(t.n_decomp);
    u.n_lower = uint32_t(t.n_lower);
    u.n_upper = uint32_t(t.n_upper);
    v->unicode = true;
  }
  *out = v;
  return TMGPU_OK;
}

extern "C" int tmgpu_vocab_load_file(const char* path, int device, tmgpu_vocab** out) {
  std::ifstream in(path ? path : "", std::ios::binary);
  if (!in) {
    g_err = std::string("cannot open vocabulary file: ") + (path ? path : "(null)");
    return TMGPU_E_VOCAB_OPEN;
  }
  std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  return tmgpu_vocab_create(bytes.data(), bytes.size(), device, out);
}

extern "C" void tmgpu_vocab_destroy(tmgpu_vocab* v) {
  if (!v) return;
  if (v->device >= 0) cudaSetDevice(v->device);
  if (v->ws) tmgpu_workspace_destroy(v->ws);
  for (void* p : v->allocs) cudaFree(p);
  if (v->d_in) cudaFree(v->d_in);
  if (v->d_off) cudaFree(v->d_off);
  if (v->h_off) cudaFreeHost(v->h_off);
  if (v->stream) cudaStreamDestroy(v->stream);
  if (v->copy_stream) cudaStreamDestroy(v->copy_stream);
  if (v->d2h_stream) cudaStreamDestroy(v->d2h_stream);
  for (auto& e : v->chunk_ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : v->done_ev)
    if (e) cudaEventDestroy(e);
  if (v->d_run) cudaFree(v->d_run);
  if (v->h_info) cudaFreeHost(v->h_info);
  for (auto& b : v->pinned) cudaFreeHost(b.p);
  delete v;
}

extern "C" int tmgpu_vocab_info_get(const tmgpu_vocab* v, tmgpu_vocab_info* o) {
  if (!v || !o) {
    g_err = "tmgpu_vocab_info_get: null argument";
    return TMGPU_E_ARG;
  }
  const HostTables& h = v->host;
  std::memset(o, 0, sizeof(*o));
  o->capcode = h.capcode;
  o->charset = h.charset;
  o->normflag = h.normflag;
  o->level = h.level;
  o->reserve = h.reserve;
  o->max_token_length = uint8_t(h.max_len);
  o->unk_token = h.unk;
  o->delete_token = h.delete_token;
  o->vocab_size = h.vocab_size;
  o->n_reverse = h.n_reverse;
  o->n_info = h.n_info;
  o->n_trie_nodes = h.n_nodes;
  return TMGPU_OK;
}

// Table self-check (host): walks the same prefix trie the kernels walk, for a single window.
// Not a tokenizer and not a fallback -- it exists so that CPU-only CI can verify the table
// builder against the oracle's dictionary.  Returns matched length in window bytes.
// spaced: bit 0 = window sits behind the virtual space, bit 1 = first steps through the shortcut
// tables (root2 / sp_child) the lane-per-string kernel uses.
extern "C" int tmgpu_selftest_trie_lookup(const tmgpu_vocab* v, const uint8_t* window, int n, int spaced,
                                          uint32_t* index) {
  return host_trie_lookup(v->host, window, n, spaced, index);
}

// Table self-check (host): what the generated Unicode data says about one code point, so that
// CPU-only CI can compare the whole table with an independent source (Python's unicodedata).
// Returns the property word; lower/upper = simple case mappings; decomp gets the full canonical
// decomposition (up to 4 code points, Hangul excluded), *n_decomp its length.
extern "C" uint32_t tmgpu_selftest_unicode(uint32_t cp, uint32_t* lower, uint32_t* upper, uint32_t* decomp,
                                           int* n_decomp) {
  const UnicodeHost& t = unicode_host_tables();
  if (n_decomp) *n_decomp = 0;
  if (lower) *lower = cp;
  if (upper) *upper = cp;
  if (cp >= 0x110000u) return 0;
  const uint32_t prop = t.stage2[size_t(t.stage1[cp >> 7]) * 128u + (cp & 127u)];
  auto find = [](const uint32_t* keys, size_t n, uint32_t c) -> long {
    long lo = 0, hi = long(n) - 1;
    while (lo <= hi) {
      long mid = (lo + hi) / 2;
      if (keys[mid] == c) return mid;
      if (keys[mid] < c) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
  };
  if ((prop & 16u) && lower) { long k = find(t.lkey, t.n_lower, cp); if (k >= 0) *lower = t.lval[k]; }
  if ((prop & 32u) && upper) { long k = find(t.ukey, t.n_upper, cp); if (k >= 0) *upper = t.uval[k]; }
  if ((prop & 8u) && decomp && n_decomp) {
    long k = find(t.dkey, t.n_decomp, cp);
    if (k >= 0) for (uint32_t j = t.didx[k]; j < t.didx[k + 1] && *n_decomp < 4; j++) decomp[(*n_decomp)++] = t.ddata[j];
  }
  return prop;
}

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
extern "C" void tmgpu_workspace_destroy(tmgpu_workspace* ws) {
  if (!ws) return;
  if (ws->v && ws->v->device >= 0) cudaSetDevice(ws->v->device);
  for (void* p : ws->allocs) cudaFree(p);
  if (ws->uni_text) cudaFree(ws->uni_text);
  for (void* p : ws->ck_allocs) cudaFree(p);
  for (int i = 0; i < tmgpu_workspace::kEv; i++) {
    if (ws->ev0[i]) cudaEventDestroy(ws->ev0[i]);
    if (ws->ev1[i]) cudaEventDestroy(ws->ev1[i]);
  }
  delete ws;
}

extern "C" int tmgpu_workspace_create(tmgpu_vocab* v, uint32_t max_strings, uint64_t max_bytes,
                                      tmgpu_workspace** out) {
  if (!v || !out) {
    g_err = "tmgpu_workspace_create: null argument";
    return TMGPU_E_ARG;
  }
  if (v->device < 0) {
    g_err = "vocabulary was created without a device (this library has no CPU path)";
    return TMGPU_E_CUDA;
  }
  CU(cudaSetDevice(v->device));
  auto* ws = new tmgpu_workspace();
  ws->v = v;
  ws->max_strings = max_strings;
  ws->max_bytes = max_bytes;
  const size_t n1 = size_t(max_strings) + 1;
  ws->slab_ids = max_bytes + uint64_t(ws->slab_pad) * max_strings + 64;
  int rc = 0;
  auto A = [&](auto** p, size_t count) {
    if (!rc) rc = dev_alloc(ws->allocs, p, count);
  };
  A(&ws->slab, ws->slab_ids);
  A(&ws->counts, n1);
  A(&ws->need, n1);
  A(&ws->missing, n1);
  A(&ws->status, n1);
  A(&ws->tok_off, n1);
  A(&ws->tokens, ws->slab_ids * 4);
  A(&ws->work_counter, tmgpu_workspace::kMaxChunks);
  A(&ws->overflow, 4);
  ws->scan_bytes = scan_temp_bytes(uint32_t(n1)) + 256;
  A(reinterpret_cast<uint8_t**>(&ws->scan_temp), ws->scan_bytes);
  if (v->unicode) {
    A(&ws->uni_list, n1);
    A(&ws->uni_begin, n1);
    A(&ws->uni_end, n1);
    A(&ws->uni_ctr, 2);
  }
  A(&ws->ck_n, n1);
  A(&ws->ck_base, n1);
  A(&ws->ck_skip, n1);
  A(&ws->ck_fail, n1);
  A(&ws->ck_list, n1);
  A(&ws->ck_fb, n1);
  A(&ws->ck_fe, n1);
  A(&ws->ck_ctr, 2);
  A(&ws->score_buf, size_t(v->host.n_reverse) + 256 + 8);
  A(&ws->raw_len, n1);
  A(&ws->raw_off, n1);
  A(&ws->out_len, n1);
  A(&ws->out_off, n1);
  if (rc) {
    tmgpu_workspace_destroy(ws);
    return rc;
  }
  for (int i = 0; i < tmgpu_workspace::kEv; i++) {
    CU(cudaEventCreate(&ws->ev0[i]));
    CU(cudaEventCreate(&ws->ev1[i]));
  }
  *out = ws;
  return TMGPU_OK;
}

// Durations (ms) of the tokenize kernel in the launches since the previous call, oldest first.
// Synchronises on the last recorded event.  Returns the number of values written.
extern "C" int tmgpu_workspace_kernel_times(tmgpu_workspace* ws, float* ms_out, int max_out) {
  if (!ws || !ms_out) return 0;
  int n = ws->ev_count < max_out ? ws->ev_count : max_out;
  int first = (ws->ev_head - ws->ev_count + 2 * tmgpu_workspace::kEv) % tmgpu_workspace::kEv;
  int skip = ws->ev_count - n;
  for (int k = 0; k < n; k++) {
    int i = (first + skip + k) % tmgpu_workspace::kEv;
    if (cudaEventSynchronize(ws->ev1[i]) != cudaSuccess ||
        cudaEventElapsedTime(&ms_out[k], ws->ev0[i], ws->ev1[i]) != cudaSuccess) {
      cudaGetLastError();
      ms_out[k] = -1.f;
    }
  }
  ws->ev_count = 0;
  return n;
}

extern "C" int tmgpu_last_launch_count(const tmgpu_workspace* ws) { return ws ? ws->last_launches : 0; }

// What the most recent tokenize / count call did besides the plain launch: out[0] = strings that
// went through the non-ASCII normalizer, out[1] = chunks launched for long documents, out[2] =
// chunked documents redone by a single lane, out[3] = kernels launched.  `ws` may be NULL: then
// the vocabulary's internal workspace (host-buffer calls) is reported.
extern "C" int tmgpu_last_stats(const tmgpu_vocab* v, const tmgpu_workspace* ws, uint32_t out[4]) {
  if (!ws && v) ws = v->ws;
  if (!ws || !out) {
    g_err = "tmgpu_last_stats: no workspace yet";
    return TMGPU_E_ARG;
  }
  out[0] = ws->last_unicode;
  out[1] = ws->last_chunked;
  out[2] = ws->last_chunk_redo;
  out[3] = uint32_t(ws->last_launches);
  return TMGPU_OK;
}

// A host-buffer batch is tokenized in chunks of strings so that the H2D copy of chunk k+1 overlaps
// the kernel of chunk k: `ready` is recorded on the copy stream after the chunk's bytes landed.
struct ChunkPlan {
  uint32_t s0, s1;       // strings [s0, s1)
  uint64_t slab_base;    // (off[s0] - off[0]) + slab_pad * s0
  cudaEvent_t ready;     // may be null
};

// ---- long documents in chunks (chunking.cuh) -----------------------------------------------------
// Phase 1 (before the main launch): decide which documents are chunked.  Returns the number of
// chunks (0 = nothing to do); the main launch then skips those documents via ws->ck_skip.
static int chunk_plan(tmgpu_workspace* ws, const TokenizeArgs& a, cudaStream_t st, uint32_t* total_out) {
  tmgpu_vocab* v = ws->v;
  *total_out = 0;
  ws->last_chunked = 0;
  ws->last_chunk_redo = 0;
  static const bool off = getenv("TMGPU_NO_CHUNKS") != nullptr;   // diagnostic: one lane per document always
  const HostTables& h = v->host;
  if (off || a.n == 0 || a.n >= uint32_t(v->sms) * 512u || h.max_len == 0 || h.charset == 2 || a.slab_off || a.scores)
    return TMGPU_OK;   // (overlapping chunks would be scored twice)
  CU(launch_chunk_count(a.off, a.ends, a.n, ws->ck_n, ws->ck_skip, v->sms, st));
  CU(cudaMemsetAsync(ws->ck_n + a.n, 0, 4, st));
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->ck_n, ws->ck_base, a.n + 1, st));
  uint64_t total = 0;
  CU(cudaMemcpyAsync(&total, ws->ck_base + a.n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (total == 0 || total > (1ull << 19)) return TMGPU_OK;   // nothing long, or far too much for the logs
  if (ws->ck_chunks_cap < total) {
    for (void* p : ws->ck_allocs) cudaFree(p);
    ws->ck_allocs.clear();
    ws->ck_chunks_cap = 0;
    const uint64_t cap = total + total / 4 + 64, log = cap * chunk_log_capacity();
    int rc = 0;
    auto A = [&](auto** p, size_t count) {
      if (!rc) rc = dev_alloc(ws->ck_allocs, p, count);
    };
    A(&ws->ck_items, cap);
    A(&ws->ck_out, cap);
    A(&ws->ck_ids, log);
    A(&ws->ck_pos, log);
    A(&ws->ck_miss, log);
    A(&ws->ck_sb, cap);
    A(&ws->ck_se, cap);
    A(&ws->ck_sd, cap);
    A(&ws->ck_ma, cap);
    A(&ws->ck_mb, cap);
    A(&ws->ck_dl, cap);
    A(&ws->ck_sm, cap);
    A(&ws->ck_fl, cap);
    if (rc) return rc;
    ws->ck_chunks_cap = cap;
  }
  *total_out = uint32_t(total);
  return TMGPU_OK;
}

// Phase 2 (after the main launch): tokenize the chunks, stitch, gather; documents whose chunks do
// not meet are tokenized again by one lane each.
static int chunk_run(tmgpu_workspace* ws, const TokenizeArgs& first, uint32_t total, int mode, cudaStream_t st) {
  tmgpu_vocab* v = ws->v;
  CU(cudaMemsetAsync(ws->ck_fail, 0, first.n, st));
  CU(launch_chunk_fill(first.data, first.off, first.ends, first.n, ws->ck_n, ws->ck_base, total, ws->ck_items,
                       ws->ck_fail, v->sms, st));
  TokenizeArgs c = first;
  c.n = total;
  c.slab = ws->ck_ids;
  c.skip = ws->ck_fail;      // per document: planning found no safe boundary, its chunks do nothing
  c.watermark = nullptr;
  c.chunks = ws->ck_items;
  c.c_pos = ws->ck_pos;
  c.c_miss = ws->ck_miss;
  c.c_out = ws->ck_out;
  c.c_cap = chunk_log_capacity();
  c.c_ov = chunk_overlap();
  c.work_counter = ws->work_counter + (tmgpu_workspace::kMaxChunks - 2);
  CU(cudaMemsetAsync(c.work_counter, 0, 4, st));
  static const bool trace = getenv("TMGPU_TRACE") != nullptr;   // diagnostic: phase times on stderr
  cudaEvent_t te[4] = {};
  if (trace) {
    for (auto& e : te) cudaEventCreate(&e);
    cudaEventRecord(te[0], st);
  }
  CU(launch_tokenize_chunks(c, v->sms, st));
  if (trace) cudaEventRecord(te[1], st);
  StitchArgs sa{};
  sa.off = first.off;
  sa.ends = first.ends;
  sa.n = first.n;
  sa.n_chunks = ws->ck_n;
  sa.chunk_base = ws->ck_base;
  sa.c_out = ws->ck_out;
  sa.c_pos = ws->ck_pos;
  sa.c_miss = ws->ck_miss;
  sa.c_cap = c.c_cap;
  sa.slab_pad = first.slab_pad;
  sa.count_mode = mode == 1;
  sa.doc_fail = ws->ck_fail;
  sa.seg_begin = ws->ck_sb;
  sa.seg_end = ws->ck_se;
  sa.seg_dst = ws->ck_sd;
  sa.meet_a = ws->ck_ma;
  sa.meet_b = ws->ck_mb;
  sa.seg_dels = ws->ck_dl;
  sa.seg_miss = ws->ck_sm;
  sa.seg_flag = ws->ck_fl;
  sa.counts = first.counts;
  sa.need = first.need;
  sa.missing = first.missing;
  sa.status = first.status;
  sa.overflow = first.overflow;
  CU(launch_chunk_stitch(sa, total, ws->ck_items, v->sms, st));
  if (trace) cudaEventRecord(te[2], st);
  if (mode == 0)
    CU(launch_chunk_gather(ws->ck_items, total, ws->ck_ids, c.c_cap, ws->ck_sb, ws->ck_se, ws->ck_sd, first.off,
                           first.slab_pad, first.slab, v->sms, st));
  if (trace) {
    cudaEventRecord(te[3], st);
    cudaEventSynchronize(te[3]);
    float t0 = 0, t1 = 0, t2 = 0;
    cudaEventElapsedTime(&t0, te[0], te[1]);
    cudaEventElapsedTime(&t1, te[1], te[2]);
    cudaEventElapsedTime(&t2, te[2], te[3]);
    fprintf(stderr, "[tmgpu] chunks: %u, tokenize %.3f ms, stitch %.3f ms, gather %.3f ms\n", total, t0, t1, t2);
    for (auto& e : te) cudaEventDestroy(e);
  }
  CU(cudaMemsetAsync(ws->ck_ctr, 0, 8, st));
  CU(launch_chunk_failed(ws->ck_fail, ws->ck_n, first.off, first.ends, first.n, ws->ck_list, ws->ck_fb, ws->ck_fe,
                         ws->ck_ctr, v->sms, st));
  unsigned long long redo = 0;
  CU(cudaMemcpyAsync(&redo, ws->ck_ctr, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  ws->last_launches += 7 + (mode == 0 ? 1 : 0);
  ws->last_chunked = total;
  ws->last_chunk_redo = uint32_t(redo);
  if (redo) {   // one lane per document, the ordinary way
    TokenizeArgs b = first;
    b.off = ws->ck_fb;
    b.ends = ws->ck_fe;
    b.n = uint32_t(redo);
    b.list = ws->ck_list;
    b.raw_off = first.off;
    b.raw_ends = first.ends;
    b.skip = nullptr;
    b.watermark = nullptr;
    b.work_counter = ws->work_counter + (tmgpu_workspace::kMaxChunks - 2);
    CU(cudaMemsetAsync(b.work_counter, 0, 4, st));
    CU(launch_tokenize(b, 1, mode == 0, v->sms, st, nullptr));
    ws->last_launches++;
  }
  return TMGPU_OK;
}

// Strings the tokenize kernel gave up on (first byte >= 0x80 under a normalizing vocabulary) are
// normalized by the device normalizer and tokenized by a second launch that reads the normalized
// text verbatim.  One small D2H tells the host whether there are any; ASCII batches pay only that.
static int unicode_pass(tmgpu_workspace* ws, const TokenizeArgs& first, bool store, cudaStream_t st) {
  tmgpu_vocab* v = ws->v;
  ws->last_unicode = 0;
  if (!v->unicode || first.n == 0 || v->host.max_len == 0) return TMGPU_OK;
  CU(cudaMemsetAsync(ws->uni_ctr, 0, 16, st));
  CU(launch_select_refused(first.status, first.off, first.ends, first.n, ws->uni_list, ws->uni_begin, ws->uni_ctr,
                           v->sms, st));
  unsigned long long h[2] = {0, 0};
  CU(cudaMemcpyAsync(h, ws->uni_ctr, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  ws->last_launches++;
  if (h[0] == 0) return TMGPU_OK;
  if (ws->uni_text_cap < h[1] + 64) {
    if (ws->uni_text) cudaFree(ws->uni_text);
    ws->uni_text = nullptr;
    ws->uni_text_cap = h[1] + h[1] / 4 + 4096;
    cudaError_t e = cudaMalloc(&ws->uni_text, ws->uni_text_cap);
    if (e != cudaSuccess) {
      ws->uni_text_cap = 0;
      g_err = std::string("cudaMalloc(normalizer scratch): ") + cudaGetErrorString(e);
      return TMGPU_E_NOMEM;
    }
  }
  NormalizeArgs na{};
  na.U = v->uni;
  na.data = first.data;
  na.off = first.off;
  na.ends = first.ends;
  na.list = ws->uni_list;
  na.n_list = reinterpret_cast<const uint32_t*>(ws->uni_ctr);
  na.scratch = ws->uni_text;
  na.n_begin = ws->uni_begin;
  na.n_end = ws->uni_end;
  na.normflag = v->host.normflag;
  na.capcode = v->host.capcode;
  static const bool trace = getenv("TMGPU_TRACE") != nullptr;   // diagnostic: phase times on stderr
  cudaEvent_t te[3] = {};
  if (trace) {
    for (auto& e : te) cudaEventCreate(&e);
    cudaEventRecord(te[0], st);
  }
  CU(launch_unicode_normalize(na, v->sms, st));
  if (trace) cudaEventRecord(te[1], st);
  TokenizeArgs b = first;
  b.data = ws->uni_text;
  b.off = ws->uni_begin;
  b.ends = ws->uni_end;
  b.n = uint32_t(h[0]);
  b.list = ws->uni_list;
  b.raw_off = first.off;
  b.raw_ends = first.ends;
  b.prenorm = 1;
  b.skip = nullptr;
  b.watermark = nullptr;
  b.work_counter = ws->work_counter + (tmgpu_workspace::kMaxChunks - 1);
  CU(cudaMemsetAsync(b.work_counter, 0, 4, st));
  CU(launch_tokenize(b, 1, store, v->sms, st, nullptr));
  if (trace) {
    cudaEventRecord(te[2], st);
    cudaEventSynchronize(te[2]);
    float a = 0, c = 0;
    cudaEventElapsedTime(&a, te[0], te[1]);
    cudaEventElapsedTime(&c, te[1], te[2]);
    fprintf(stderr, "[tmgpu] non-ASCII pass: %llu strings, scratch %llu B, normalize %.3f ms, tokenize %.3f ms\n", h[0],
            h[1], a, c);
    for (auto& e : te) cudaEventDestroy(e);
  }
  ws->last_launches += 2;
  ws->last_unicode = uint32_t(h[0]);
  return TMGPU_OK;
}

static int tokenize_on_stream(tmgpu_workspace* ws, const uint8_t* d_data, const int64_t* d_off, uint32_t n,
                              uint8_t enc_len, int mode, cudaStream_t st, const uint64_t* slab_off,
                              uint32_t* slab, uint8_t* dense, const std::vector<ChunkPlan>* chunks = nullptr,
                              const int64_t* d_ends = nullptr) {
  tmgpu_vocab* v = ws->v;
  ws->last_launches = 0;
  CU(cudaMemsetAsync(ws->work_counter, 0, 4 * tmgpu_workspace::kMaxChunks, st));
  CU(cudaMemsetAsync(ws->overflow, 0, 4, st));
  CU(cudaMemsetAsync(ws->counts + n, 0, 4, st));
  TokenizeArgs a{};
  a.v = v->dev;
  a.data = d_data;
  a.off = d_off;
  a.ends = d_ends;
  a.n = n;
  a.slab = slab;
  a.slab_pad = ws->slab_pad;
  a.slab_off = slab_off;
  a.counts = ws->counts;
  a.need = ws->need;
  a.missing = ws->missing;
  a.status = ws->status;
  a.work_counter = ws->work_counter;
  a.overflow = ws->overflow;
  if (mode == 2 && n > 0 && v->host.max_len != 0) {
    // training scores (trainvocab.go:932-1170): a histogram cannot be rolled back, so strings the
    // first pass would abandon at a non-ASCII byte are set aside before it starts
    a.scores = ws->score_buf;
    a.miss_hist = ws->score_buf + v->host.n_reverse;
    CU(cudaMemsetAsync(ws->score_buf, 0, (size_t(v->host.n_reverse) + 256) * 8, st));
    if (v->unicode) {
      if (chunks)
        for (const ChunkPlan& c : *chunks)
          if (c.ready) CU(cudaStreamWaitEvent(st, c.ready, 0));
      CU(launch_flag_non_ascii(d_data, d_off, d_ends, n, ws->ck_skip, ws->status, ws->counts, ws->missing, v->sms, st));
      a.skip = ws->ck_skip;
      ws->last_launches++;
    }
  }
  uint32_t ck_total = 0;   // long documents of a small batch are tokenized in chunks (chunking.cuh)
  if (n > 0 && (!chunks || chunks->size() == 1)) {
    if (chunks && (*chunks)[0].ready) CU(cudaStreamWaitEvent(st, (*chunks)[0].ready, 0));
    int rc = chunk_plan(ws, a, st, &ck_total);
    if (rc) return rc;
    if (ck_total) a.skip = ws->ck_skip;
  }
  if (v->host.max_len == 0) {  // empty vocabulary: empty result, not an error (go:960-962)
    CU(cudaMemsetAsync(ws->counts, 0, (size_t(n) + 1) * 4, st));
    CU(cudaMemsetAsync(ws->missing, 0, size_t(n) * 4, st));
    CU(cudaMemsetAsync(ws->status, 0, n, st));
  } else if (n > 0 && chunks && !slab_off) {
    int ci = 0;
    for (const ChunkPlan& c : *chunks) {
      if (c.s1 <= c.s0) continue;
      if (c.ready) CU(cudaStreamWaitEvent(st, c.ready, 0));
      TokenizeArgs b = a;
      b.off = d_off + c.s0;
      if (d_ends) b.ends = d_ends + c.s0;
      if (a.skip) b.skip = a.skip + c.s0;
      b.n = c.s1 - c.s0;
      b.slab = slab + c.slab_base;
      b.counts = ws->counts + c.s0;
      b.need = ws->need + c.s0;
      b.missing = ws->missing + c.s0;
      b.status = ws->status + c.s0;
      b.work_counter = ws->work_counter + (ci++ % tmgpu_workspace::kMaxChunks);
      CU(launch_tokenize(b, v->group, mode == 0, v->sms, st, nullptr));
      ws->last_launches++;
    }
  } else if (n > 0) {
    const int e = ws->ev_head;
    CU(cudaEventRecord(ws->ev0[e], st));
    CU(launch_tokenize(a, v->group, mode == 0, v->sms, st, nullptr));
    CU(cudaEventRecord(ws->ev1[e], st));
    ws->ev_head = (e + 1) % tmgpu_workspace::kEv;
    if (ws->ev_count < tmgpu_workspace::kEv) ws->ev_count++;
    ws->last_launches++;
  }
  if (ck_total) {
    int rc = chunk_run(ws, a, ck_total, mode, st);
    if (rc) return rc;
    a.skip = nullptr;
  }
  {
    int rc = unicode_pass(ws, a, mode == 0, st);
    if (rc) return rc;
  }
  if (mode == 0) {
    CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->counts, ws->tok_off, n + 1, st));
    ws->last_launches += 2;  // cub::DeviceScan = init kernel + scan kernel
    if (n > 0) {
      CU(launch_pack(a.slab, d_off, ws->slab_pad, slab_off, ws->counts, ws->tok_off, dense, n, enc_len, v->sms, st));
      ws->last_launches++;
    }
  }
  return TMGPU_OK;
}

extern "C" int tmgpu_tokenize_device(tmgpu_workspace* ws, const uint8_t* d_data, const int64_t* d_off, uint32_t n,
                                     uint8_t enc_len, int mode, void* stream, tmgpu_device_view* view) {
  if (!ws || !d_off || (!d_data && n) || !(mode == 0 || mode == 1)) {
    g_err = "tmgpu_tokenize_device: bad argument";
    return TMGPU_E_ARG;
  }
  if (enc_len == 0) enc_len = ws->v->host.n_reverse <= 65536 ? 2 : 3;  // go:990-996
  if (mode == 0 && !(enc_len == 2 || enc_len == 3 || enc_len == 4)) {
    g_err = "invalid encoding length";  // go:1012
    return TMGPU_E_ARG;
  }
  if (n > ws->max_strings) {
    g_err = "batch has more strings than the workspace was created for";
    return TMGPU_E_ARG;
  }
  if (reinterpret_cast<uintptr_t>(d_data) & 15) {
    g_err = "d_data must be 16-byte aligned";
    return TMGPU_E_ARG;
  }
  CU(cudaSetDevice(ws->v->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);  // NULL = the CUDA default stream
  int rc = tokenize_on_stream(ws, d_data, d_off, n, enc_len, mode, st, nullptr, ws->slab, ws->tokens);
  if (rc) return rc;
  if (view) {
    view->tokens = ws->tokens;
    view->tok_off = ws->tok_off;
    view->missing = ws->missing;
    view->status = ws->status;
    view->n_tokens = ws->tok_off + n;
    view->counts = ws->counts;
  }
  return TMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer calls
// ------------------------------------------------------------------------------------------
static int ensure_host_path(tmgpu_vocab* v, uint32_t n, uint64_t bytes) {
  if (v->device < 0) {
    g_err = "vocabulary was created without a device (this library has no CPU path)";
    return TMGPU_E_CUDA;
  }
  CU(cudaSetDevice(v->device));
  if (!v->ws || v->ws->max_strings < n || v->ws->max_bytes < bytes) {
    if (v->ws) {
      CU(cudaStreamSynchronize(v->stream));
      tmgpu_workspace_destroy(v->ws);
      v->ws = nullptr;
    }
    uint32_t ms = n < 1024 ? 1024 : n + n / 8;
    uint64_t mb = bytes < (1u << 20) ? (1u << 20) : bytes + bytes / 8;
    int rc = tmgpu_workspace_create(v, ms, mb, &v->ws);
    if (rc) return rc;
  }
  if (v->d_in_cap < bytes + 64) {
    if (v->d_in) cudaFree(v->d_in);
    v->d_in = nullptr;
    v->d_in_cap = bytes + bytes / 8 + 4096;
    CU(cudaMalloc(&v->d_in, v->d_in_cap));
  }
  if (v->d_off_cap < size_t(n) + 1) {
    if (v->d_off) cudaFree(v->d_off);
    if (v->h_off) cudaFreeHost(v->h_off);
    v->d_off = nullptr;
    v->h_off = nullptr;
    v->d_off_cap = size_t(n) + n / 8 + 1024;
    CU(cudaMalloc(&v->d_off, v->d_off_cap * 16));
    CU(cudaMallocHost(&v->h_off, v->d_off_cap * 16));
  }
  return TMGPU_OK;
}

struct ResultPriv {
  tmgpu_vocab* v = nullptr;
  std::vector<void*> blocks;
};

// pinned host memory from the vocabulary's pool (cudaMallocHost is far too slow to pay per call)
template <class T>
static int pinned_alloc(ResultPriv* p, T** out, size_t count) {
  size_t bytes = count * sizeof(T);
  if (bytes < 64) bytes = 64;
  tmgpu_vocab* v = p->v;
  std::lock_guard<std::mutex> lock(v->pin_mu);
  int best = -1;
  for (size_t i = 0; i < v->pinned.size(); i++) {
    auto& b = v->pinned[i];
    if (!b.used && b.bytes >= bytes && (best < 0 || b.bytes < v->pinned[best].bytes)) best = int(i);
  }
  if (best < 0) {
    void* q = nullptr;
    size_t cap = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&q, cap);
    if (e != cudaSuccess) {
      g_err = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
      return TMGPU_E_NOMEM;
    }
    v->pinned.push_back({q, cap, false});
    best = int(v->pinned.size()) - 1;
  }
  v->pinned[best].used = true;
  p->blocks.push_back(v->pinned[best].p);
  *out = static_cast<T*>(v->pinned[best].p);
  return TMGPU_OK;
}

static void release_priv(ResultPriv* p) {
  if (!p) return;
  if (p->v) {
    std::lock_guard<std::mutex> lock(p->v->pin_mu);
    for (void* q : p->blocks)
      for (auto& b : p->v->pinned)
        if (b.p == q) b.used = false;
  }
  delete p;
}

extern "C" void tmgpu_result_free(tmgpu_result* r) {
  if (!r) return;
  release_priv(static_cast<ResultPriv*>(r->priv));
  delete r;
}

extern "C" void tmgpu_bytes_free(tmgpu_bytes* b) {
  if (!b) return;
  release_priv(static_cast<ResultPriv*>(b->priv));
  delete b;
}

// stage host strings: offsets rebased to 0 go first, then the bytes in chunks on the copy stream;
// each chunk records an event the compute stream waits on, so copy k+1 overlaps kernel k.
static int stage_copy(tmgpu_vocab* v, const uint8_t* src, uint32_t n, uint64_t total, bool gapped,
                      std::vector<ChunkPlan>* plan);

static int stage_input(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n, uint64_t* total,
                       std::vector<ChunkPlan>* plan) {
  const uint64_t base = off[0];
  for (uint32_t i = 0; i <= n; i++) {
    if (off[i] < base || (i && off[i] < off[i - 1])) {
      g_err = "offsets must be non-decreasing";
      return TMGPU_E_ARG;
    }
    if (i && off[i] - off[i - 1] > 0x3FFFFFFFull) {
      g_err = "a single string may not exceed 1 GiB";
      return TMGPU_E_ARG;
    }
    v->h_off[i] = int64_t(off[i] - base);
  }
  *total = off[n] - base;
  return stage_copy(v, data + base, n, *total, false, plan);
}

// h_off[0..n] (and, if gapped, the end offsets behind them) are filled in; `src` holds `total`
// bytes that go to d_in verbatim.
static int stage_copy(tmgpu_vocab* v, const uint8_t* src, uint32_t n, uint64_t total_bytes, bool gapped,
                      std::vector<ChunkPlan>* plan) {
  const uint64_t* total = &total_bytes;
  // watermark = 0, then the offsets; the compute stream may start once both are on the device
  v->h_info[40] = 0;
  CU(cudaMemcpyAsync(v->d_run + 4, v->h_info + 40, 8, cudaMemcpyHostToDevice, v->copy_stream));
  CU(cudaMemcpyAsync(v->d_off, v->h_off, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, v->copy_stream));
  if (gapped && n)
    CU(cudaMemcpyAsync(v->d_off + v->d_off_cap, v->h_off + v->d_off_cap, size_t(n) * 8, cudaMemcpyHostToDevice,
                       v->copy_stream));
  CU(cudaEventRecord(v->chunk_ev[31], v->copy_stream));
  // chunks of ~equal bytes; after each one the watermark (bytes landed) moves forward
  int want = int(*total / (16ull << 20)) + 1;
  if (want > 24) want = 24;
  if (n < uint32_t(v->sms) * 800u) want = 1;   // small batches: one copy, then the plain path
  if (plan) plan->clear();
  uint32_t s0 = 0;
  const uint32_t pad = v->ws->slab_pad;
  for (int c = 0; c < want && s0 < n; c++) {
    uint32_t s1 = n;
    if (c + 1 < want) {
      const uint64_t target = *total * uint64_t(c + 1) / uint64_t(want);
      uint32_t lo = s0, hi = n;  // first string whose end offset passes the target
      while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (uint64_t(v->h_off[mid + 1]) < target) lo = mid + 1; else hi = mid;
      }
      s1 = lo + 1 > n ? n : lo + 1;
    }
    const uint64_t b0 = uint64_t(v->h_off[s0]), b1 = uint64_t(v->h_off[s1]);
    if (b1 > b0) CU(cudaMemcpyAsync(v->d_in + b0, src + b0, b1 - b0, cudaMemcpyHostToDevice, v->copy_stream));
    v->h_info[41 + c] = b1;
    CU(cudaMemcpyAsync(v->d_run + 4, v->h_info + 41 + c, 8, cudaMemcpyHostToDevice, v->copy_stream));
    CU(cudaEventRecord(v->chunk_ev[c], v->copy_stream));
    if (plan) plan->push_back({s0, s1, (b0 - uint64_t(v->h_off[0])) + uint64_t(pad) * s0, v->chunk_ev[c]});
    s0 = s1;
  }
  if (!plan) CU(cudaStreamSynchronize(v->copy_stream));
  return TMGPU_OK;
}

// If a string produced more ids than its slab share (len + 16), run again with exact sizes.
// On success *dense points at the serialized ids to copy out; temp holds what to cudaFree.
static int rerun_if_overflow(tmgpu_vocab* v, uint32_t n, uint8_t enc_len, std::vector<void*>* temp,
                             uint8_t** dense, const int64_t* d_ends = nullptr) {
  tmgpu_workspace* ws = v->ws;
  uint32_t flag = 0;
  CU(cudaMemcpyAsync(&flag, ws->overflow, 4, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  if (!flag) return TMGPU_OK;
  CU(cudaMemsetAsync(ws->need + n, 0, 4, v->stream));
  uint64_t* slab_off = nullptr;
  int rc = dev_alloc(*temp, &slab_off, size_t(n) + 1);
  if (rc) return rc;
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->need, slab_off, n + 1, v->stream));
  uint64_t total = 0;
  CU(cudaMemcpyAsync(&total, slab_off + n, 8, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  uint32_t* big = nullptr;
  if ((rc = dev_alloc(*temp, &big, total + 16))) return rc;
  uint8_t* d = nullptr;
  if ((rc = dev_alloc(*temp, &d, (total + 16) * 4))) return rc;
  if ((rc = tokenize_on_stream(ws, v->d_in, v->d_off, n, enc_len, 0, v->stream, slab_off, big, d, nullptr, d_ends)))
    return rc;
  *dense = d;
  return TMGPU_OK;
}

// Large host batches: ONE launch of the tokenize kernel runs while the strings are still being
// copied in.  stage_input() issues the H2D copies in chunks on the copy stream and, after each,
// an 8-byte copy that advances a device "watermark" (bytes landed); a lane that pulls a string
// beyond the watermark sleeps until its bytes are there.  No SM is needed to move the watermark
// (it is a DMA write), so the persistent kernel can own the whole GPU.
static int tokenize_streamed(tmgpu_vocab* v, uint32_t n, uint8_t enc_len, cudaEvent_t offsets_ready,
                             const int64_t* d_ends = nullptr) {
  tmgpu_workspace* ws = v->ws;
  cudaStream_t st = v->stream;
  ws->last_launches = 0;
  CU(cudaMemsetAsync(ws->work_counter, 0, 4 * tmgpu_workspace::kMaxChunks, st));
  CU(cudaMemsetAsync(ws->overflow, 0, 4, st));
  CU(cudaMemsetAsync(ws->counts + n, 0, 4, st));
  CU(cudaStreamWaitEvent(st, offsets_ready, 0));
  TokenizeArgs a{};
  a.v = v->dev;
  a.data = v->d_in;
  a.off = v->d_off;
  a.ends = d_ends;
  a.n = n;
  a.slab = ws->slab;
  a.slab_pad = ws->slab_pad;
  a.counts = ws->counts;
  a.need = ws->need;
  a.missing = ws->missing;
  a.status = ws->status;
  a.work_counter = ws->work_counter;
  a.overflow = ws->overflow;
  a.watermark = reinterpret_cast<const unsigned long long*>(v->d_run + 4);
  CU(launch_tokenize(a, v->group, true, v->sms, st, nullptr));
  {
    TokenizeArgs whole = a;
    whole.watermark = nullptr;
    int rc = unicode_pass(ws, whole, true, st);
    if (rc) return rc;
  }
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->counts, ws->tok_off, n + 1, st));
  CU(launch_pack(ws->slab, v->d_off, ws->slab_pad, nullptr, ws->counts, ws->tok_off, ws->tokens, n, enc_len, v->sms, st));
  ws->last_launches = 4;
  return TMGPU_OK;
}

extern "C" int tmgpu_tokenize_batch(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n,
                                    uint8_t enc_len, tmgpu_result** out) {
  if (!v || !off || !out || (!data && n && off[n] != off[0])) {
    g_err = "tmgpu_tokenize_batch: null argument";
    return TMGPU_E_ARG;
  }
  if (enc_len == 0) enc_len = v->host.n_reverse <= 65536 ? 2 : 3;
  if (!(enc_len == 2 || enc_len == 3 || enc_len == 4)) {
    g_err = "invalid encoding length";
    return TMGPU_E_ARG;
  }
  std::lock_guard<std::mutex> lock(v->mu);
  int rc = ensure_host_path(v, n, off[n] - off[0]);
  if (rc) return rc;
  uint64_t total = 0;
  std::vector<ChunkPlan> plan;
  if ((rc = stage_input(v, data, off, n, &total, &plan))) return rc;
  tmgpu_workspace* ws = v->ws;
  if (plan.size() > 1 && v->host.max_len != 0) {
    // streamed: the kernel is already chasing the copies (stage_input armed the watermark)
    if ((rc = tokenize_streamed(v, n, enc_len, v->chunk_ev[31]))) return rc;
  } else
  if ((rc = tokenize_on_stream(ws, v->d_in, v->d_off, n, enc_len, 0, v->stream, nullptr, ws->slab, ws->tokens, &plan)))
    return rc;
  std::vector<void*> temp;
  uint8_t* dense = ws->tokens;
  auto cleanup = [&]() {
    for (void* p : temp) cudaFree(p);
    temp.clear();
  };
  if ((rc = rerun_if_overflow(v, n, enc_len, &temp, &dense))) {
    cleanup();
    return rc;
  }
  auto* priv = new ResultPriv();
  priv->v = v;
  auto* r = new tmgpu_result();
  std::memset(r, 0, sizeof(*r));
  r->priv = priv;
  r->n = n;
  r->enc_len = enc_len;
  auto bail = [&](int code) {
    cleanup();
    tmgpu_result_free(r);
    return code;
  };
  if ((rc = pinned_alloc(priv, &r->tok_off, size_t(n) + 1))) return bail(rc);
  if ((rc = pinned_alloc(priv, &r->missing, size_t(n) + 1))) return bail(rc);
  if ((rc = pinned_alloc(priv, &r->status, size_t(n) + 1))) return bail(rc);
  cudaError_t e;
  e = cudaMemcpyAsync(r->tok_off, ws->tok_off, (size_t(n) + 1) * 8, cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r->missing, ws->missing, size_t(n) * 4, cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r->status, ws->status, size_t(n), cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
  if (e != cudaSuccess) {
    g_err = std::string("tokenize: ") + cudaGetErrorString(e);
    return bail(TMGPU_E_CUDA);
  }
  r->n_tokens = r->tok_off[n];
  if ((rc = pinned_alloc(priv, &r->tokens, r->n_tokens * enc_len + 8))) return bail(rc);
  if (r->n_tokens) {
    e = cudaMemcpyAsync(r->tokens, dense, r->n_tokens * enc_len, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    if (e != cudaSuccess) {
      g_err = std::string("tokenize D2H: ") + cudaGetErrorString(e);
      return bail(TMGPU_E_CUDA);
    }
  }
  cleanup();
  *out = r;
  return TMGPU_OK;
}

extern "C" int tmgpu_count_batch(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n,
                                 uint64_t* counts, uint32_t* missing, uint8_t* status) {
  if (!v || !off || !counts) {
    g_err = "tmgpu_count_batch: null argument";
    return TMGPU_E_ARG;
  }
  std::lock_guard<std::mutex> lock(v->mu);
  int rc = ensure_host_path(v, n, off[n] - off[0]);
  if (rc) return rc;
  uint64_t total = 0;
  std::vector<ChunkPlan> plan;
  if ((rc = stage_input(v, data, off, n, &total, &plan))) return rc;
  tmgpu_workspace* ws = v->ws;
  if ((rc = tokenize_on_stream(ws, v->d_in, v->d_off, n, 0, 1, v->stream, nullptr, ws->slab, ws->tokens, &plan)))
    return rc;
  std::vector<uint32_t> c(n + 1), m(n + 1);
  std::vector<uint8_t> s(n + 1);
  CU(cudaMemcpyAsync(c.data(), ws->counts, size_t(n) * 4, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaMemcpyAsync(m.data(), ws->missing, size_t(n) * 4, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaMemcpyAsync(s.data(), ws->status, size_t(n), cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  for (uint32_t i = 0; i < n; i++) {
    counts[i] = c[i];
    if (missing) missing[i] = m[i];
    if (status) status[i] = s[i];
  }
  return TMGPU_OK;
}

// The tokenization loop of the training worker (training/trainvocab.go:932-1170) for one candidate
// vocabulary over a batch of data strips.
extern "C" int tmgpu_score_batch(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n,
                                 uint64_t* scores, uint64_t* miss_hist, uint64_t* tokens_in_text, uint8_t* status) {
  if (!v || !off || !scores || !tokens_in_text) {
    g_err = "tmgpu_score_batch: null argument";
    return TMGPU_E_ARG;
  }
  std::lock_guard<std::mutex> lock(v->mu);
  int rc = ensure_host_path(v, n, off[n] - off[0]);
  if (rc) return rc;
  uint64_t total = 0;
  std::vector<ChunkPlan> plan;
  if ((rc = stage_input(v, data, off, n, &total, &plan))) return rc;
  tmgpu_workspace* ws = v->ws;
  if ((rc = tokenize_on_stream(ws, v->d_in, v->d_off, n, 0, 2, v->stream, nullptr, ws->slab, ws->tokens, &plan)))
    return rc;
  const size_t nr = v->host.n_reverse;
  std::vector<uint32_t> c(size_t(n) + 1);
  std::vector<uint8_t> s(size_t(n) + 1);
  std::vector<unsigned long long> hist(256);
  if (v->host.max_len == 0) {
    std::memset(scores, 0, nr * 8);
    if (miss_hist) std::memset(miss_hist, 0, 256 * 8);
  } else {
    CU(cudaMemcpyAsync(scores, ws->score_buf, nr * 8, cudaMemcpyDeviceToHost, v->stream));
    CU(cudaMemcpyAsync(hist.data(), ws->score_buf + nr, 256 * 8, cudaMemcpyDeviceToHost, v->stream));
  }
  CU(cudaMemcpyAsync(c.data(), ws->counts, size_t(n) * 4, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaMemcpyAsync(s.data(), ws->status, size_t(n), cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  uint64_t t = 0;
  for (uint32_t i = 0; i < n; i++) {
    t += c[i];
    if (status) status[i] = s[i];
  }
  *tokens_in_text = t;
  if (miss_hist && v->host.max_len != 0)
    for (int b = 0; b < 256; b++) miss_hist[b] = hist[b];
  return TMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------
static int ensure_decode_cap(tmgpu_workspace* ws, uint64_t bytes) {
  if (ws->dec_cap >= bytes + 64) return TMGPU_OK;
  uint64_t cap = bytes + bytes / 8 + 4096;
  // capcode decode of non-ASCII text may grow a W-word by half (2-byte letter -> 3-byte capital)
  const uint64_t wide = ws->v->unicode ? cap + cap / 2 + 64 : cap;
  int rc = dev_alloc(ws->allocs, &ws->raw, cap);
  if (!rc) rc = dev_alloc(ws->allocs, &ws->dec, wide);
  if (!rc) rc = dev_alloc(ws->allocs, &ws->dense, wide);
  if (rc) return rc;
  ws->dec_cap = cap;  // earlier (smaller) buffers stay in the pool until the workspace dies
  return TMGPU_OK;
}

// Device decode.  Needs one host round trip for the byte total (output sizing).
static int decode_on_stream(tmgpu_workspace* ws, const uint8_t* d_tokens, const int64_t* d_tok_off, uint32_t n,
                            int enc, cudaStream_t st, const uint8_t** d_bytes, const uint64_t** d_off,
                            uint64_t* total_out) {
  tmgpu_vocab* v = ws->v;
  ws->last_launches = 0;
  CU(cudaMemsetAsync(ws->raw_len + n, 0, 4, st));
  if (n) {
    CU(launch_decode_size(v->dev, d_tokens, d_tok_off, n, enc, ws->raw_len, v->sms, st));
    ws->last_launches++;
  }
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->raw_len, ws->raw_off, n + 1, st));
  ws->last_launches++;
  uint64_t raw_total = 0;
  CU(cudaMemcpyAsync(&raw_total, ws->raw_off + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  int rc = ensure_decode_cap(ws, raw_total);
  if (rc) return rc;
  if (n) {
    CU(launch_decode_copy(v->dev, d_tokens, d_tok_off, n, enc, ws->raw_off, ws->raw, v->sms, st));
    ws->last_launches++;
  }
  if (v->host.capcode != 2) {
    *d_bytes = ws->raw;
    *d_off = ws->raw_off;
    *total_out = raw_total;
    return TMGPU_OK;
  }
  CU(cudaMemsetAsync(ws->out_len + n, 0, 4, st));
  if (n) {
    CU(launch_capcode_decode(ws->raw, ws->raw_off, n, ws->dec, ws->out_len, v->sms, st, nullptr, nullptr,
                             v->unicode ? &v->uni : nullptr));
    ws->last_launches++;
  }
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->out_len, ws->out_off, n + 1, st));
  ws->last_launches++;
  if (n) {
    CU(launch_compact_bytes(ws->dec, ws->raw_off, ws->out_len, ws->out_off, n, ws->dense, v->sms, st,
                            v->unicode ? 1 : 0));
    ws->last_launches++;
  }
  uint64_t total = 0;
  CU(cudaMemcpyAsync(&total, ws->out_off + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *d_bytes = ws->dense;
  *d_off = ws->out_off;
  *total_out = total;
  return TMGPU_OK;
}

extern "C" int tmgpu_decode_device(tmgpu_workspace* ws, const uint8_t* d_tokens, const int64_t* d_tok_off, uint32_t n,
                                   uint8_t enc_len, void* stream, const uint8_t** d_bytes, const uint64_t** d_off,
                                   const uint64_t** d_total) {
  if (!ws || !d_tok_off || !d_bytes || !d_off) {
    g_err = "tmgpu_decode_device: null argument";
    return TMGPU_E_ARG;
  }
  if (enc_len == 0) enc_len = ws->v->host.n_reverse <= 65536 ? 2 : 3;
  if (!(enc_len == 2 || enc_len == 3 || enc_len == 4) || n > ws->max_strings) {
    g_err = "tmgpu_decode_device: bad enc_len or batch larger than the workspace";
    return TMGPU_E_ARG;
  }
  CU(cudaSetDevice(ws->v->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);  // NULL = the CUDA default stream
  uint64_t total = 0;
  int rc = decode_on_stream(ws, d_tokens, d_tok_off, n, enc_len, st, d_bytes, d_off, &total);
  if (rc) return rc;
  if (d_total) *d_total = *d_off + n;
  return TMGPU_OK;
}

extern "C" int tmgpu_decode_batch(tmgpu_vocab* v, const uint8_t* tokens, const uint64_t* tok_off, uint32_t n,
                                  uint8_t enc_len, tmgpu_bytes** out) {
  if (!v || !tok_off || !out) {
    g_err = "tmgpu_decode_batch: null argument";
    return TMGPU_E_ARG;
  }
  if (enc_len == 0) enc_len = v->host.n_reverse <= 65536 ? 2 : 3;
  if (!(enc_len == 2 || enc_len == 3 || enc_len == 4)) {
    g_err = "invalid encoding length";
    return TMGPU_E_ARG;
  }
  std::lock_guard<std::mutex> lock(v->mu);
  const uint64_t n_ids = tok_off[n] - tok_off[0];
  int rc = ensure_host_path(v, n, n_ids * enc_len + 64);
  if (rc) return rc;
  for (uint32_t i = 0; i <= n; i++) {
    if (tok_off[i] < tok_off[0] || (i && tok_off[i] < tok_off[i - 1])) {
      g_err = "token offsets must be non-decreasing";
      return TMGPU_E_ARG;
    }
    v->h_off[i] = int64_t(tok_off[i] - tok_off[0]);
  }
  if (n_ids)
    CU(cudaMemcpyAsync(v->d_in, tokens + tok_off[0] * enc_len, n_ids * enc_len, cudaMemcpyHostToDevice, v->stream));
  CU(cudaMemcpyAsync(v->d_off, v->h_off, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, v->stream));
  const uint8_t* d_bytes = nullptr;
  const uint64_t* d_off = nullptr;
  uint64_t total = 0;
  if ((rc = decode_on_stream(v->ws, v->d_in, v->d_off, n, enc_len, v->stream, &d_bytes, &d_off, &total))) return rc;
  auto* priv = new ResultPriv();
  priv->v = v;
  auto* b = new tmgpu_bytes();
  std::memset(b, 0, sizeof(*b));
  b->priv = priv;
  b->n = n;
  b->n_bytes = total;
  if ((rc = pinned_alloc(priv, &b->off, size_t(n) + 1)) || (rc = pinned_alloc(priv, &b->bytes, total + 8))) {
    tmgpu_bytes_free(b);
    return rc;
  }
  cudaError_t e = cudaMemcpyAsync(b->off, d_off, (size_t(n) + 1) * 8, cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess && total) e = cudaMemcpyAsync(b->bytes, d_bytes, total, cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
  if (e != cudaSuccess) {
    g_err = std::string("decode D2H: ") + cudaGetErrorString(e);
    tmgpu_bytes_free(b);
    return TMGPU_E_CUDA;
  }
  *out = b;
  return TMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// server wire format (training/tokenmonsterserver.go:189-209): request payloads in, reply
// payloads out.  The request is copied to the GPU verbatim; the reply is framed on the device.
// ------------------------------------------------------------------------------------------
static inline uint64_t rd_u64(const uint8_t* p) {
  uint64_t x = 0;
  for (int k = 7; k >= 0; k--) x = (x << 8) | p[k];
  return x;
}
static inline void wr_u64(uint8_t* p, uint64_t x) {
  for (int k = 0; k < 8; k++) p[k] = uint8_t(x >> (8 * k));
}

// count the entries and check that they fit the payload (the server trusts its client; a
// library cannot)
static int wire_scan(const uint8_t* payload, uint64_t len, uint32_t* n_out, uint64_t* longest) {
  if (!payload || len < 4) {
    g_err = "wire payload shorter than its 4-byte entry count";
    return TMGPU_E_ARG;
  }
  const uint32_t n = uint32_t(payload[0]) | uint32_t(payload[1]) << 8 | uint32_t(payload[2]) << 16 |
                     uint32_t(payload[3]) << 24;
  uint64_t at = 4, big = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (len - at < 8) {
      g_err = "wire payload truncated inside an entry header";
      return TMGPU_E_ARG;
    }
    const uint64_t l = rd_u64(payload + at);
    at += 8;
    if (l > len - at) {
      g_err = "wire payload truncated inside an entry body";
      return TMGPU_E_ARG;
    }
    if (l > big) big = l;
    at += l;
  }
  *n_out = n;
  *longest = big;
  return TMGPU_OK;
}

static tmgpu_wire_reply* new_reply(tmgpu_vocab* v) {
  auto* priv = new ResultPriv();
  priv->v = v;
  auto* r = new tmgpu_wire_reply();
  std::memset(r, 0, sizeof(*r));
  r->priv = priv;
  return r;
}

extern "C" void tmgpu_wire_reply_free(tmgpu_wire_reply* r) {
  if (!r) return;
  release_priv(static_cast<ResultPriv*>(r->priv));
  delete r;
}

// fill h_off / h_end from the payload and start the copies; the kernel may chase them
static int wire_stage(tmgpu_vocab* v, const uint8_t* payload, uint64_t len, uint32_t n,
                      std::vector<ChunkPlan>* plan) {
  int64_t* begins = v->h_off;
  int64_t* ends = v->h_off + v->d_off_cap;
  uint64_t at = 4;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t l = rd_u64(payload + at);
    at += 8;
    begins[i] = int64_t(at);
    ends[i] = int64_t(at + l);
    at += l;
  }
  begins[n] = int64_t(len);
  return stage_copy(v, payload, n, len, true, plan);
}

static int wire_tokenize_impl(tmgpu_vocab* v, const uint8_t* payload, uint64_t len, bool count_only,
                              tmgpu_wire_reply** out) {
  if (!v || !out) {
    g_err = "tmgpu_wire_tokenize: null argument";
    return TMGPU_E_ARG;
  }
  uint32_t n = 0;
  uint64_t longest = 0;
  int rc = wire_scan(payload, len, &n, &longest);
  if (rc) return rc;
  if (longest > 0x3FFFFFFFull) {
    g_err = "a single string may not exceed 1 GiB";
    return TMGPU_E_ARG;
  }
  const uint8_t enc = v->host.vocab_size > 65536 ? 4 : 2;  // server:351-354
  std::lock_guard<std::mutex> lock(v->mu);
  if ((rc = ensure_host_path(v, n, len))) return rc;
  std::vector<ChunkPlan> plan;
  if ((rc = wire_stage(v, payload, len, n, &plan))) return rc;
  tmgpu_workspace* ws = v->ws;
  const int64_t* d_ends = v->d_off + v->d_off_cap;
  if (!count_only && plan.size() > 1 && v->host.max_len != 0) {
    if ((rc = tokenize_streamed(v, n, enc, v->chunk_ev[31], d_ends))) return rc;
  } else if ((rc = tokenize_on_stream(ws, v->d_in, v->d_off, n, enc, count_only ? 1 : 0, v->stream, nullptr,
                                      ws->slab, ws->tokens, &plan, d_ends))) {
    return rc;
  }
  tmgpu_wire_reply* r = new_reply(v);
  auto* priv = static_cast<ResultPriv*>(r->priv);
  std::vector<void*> temp;
  auto bail = [&](int code) {
    for (void* p : temp) cudaFree(p);
    tmgpu_wire_reply_free(r);
    return code;
  };
  uint8_t st0 = 0;
  if (count_only) {
    r->len = 4 + 8ull * n;
    uint32_t* c32 = nullptr;
    if ((rc = pinned_alloc(priv, &r->payload, r->len + 8)) || (rc = pinned_alloc(priv, &c32, size_t(n) + 1)))
      return bail(rc);
    cudaError_t e = cudaMemcpyAsync(c32, ws->counts, size_t(n) * 4, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(&st0, ws->status, 1, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    if (e != cudaSuccess) {
      g_err = std::string("wire count D2H: ") + cudaGetErrorString(e);
      return bail(TMGPU_E_CUDA);
    }
    std::memcpy(r->payload, payload, 4);
    for (uint32_t i = 0; i < n; i++) wr_u64(r->payload + 4 + 8ull * i, c32[i]);
  } else {
    uint8_t* dense = ws->tokens;
    if ((rc = rerun_if_overflow(v, n, enc, &temp, &dense, d_ends))) return bail(rc);
    uint64_t total = 0;
    cudaError_t e = cudaMemcpyAsync(&total, ws->tok_off + n, 8, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(&st0, ws->status, 1, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    if (e != cudaSuccess) {
      g_err = std::string("wire tokenize: ") + cudaGetErrorString(e);
      return bail(TMGPU_E_CUDA);
    }
    r->len = 4 + 8ull * n + total * enc;
    // the id slab is free again once pack has run: frame the reply there if it fits
    uint8_t* framed = reinterpret_cast<uint8_t*>(ws->slab);
    if (r->len + 16 > ws->slab_ids * 4 && (rc = dev_alloc(temp, &framed, r->len + 16))) return bail(rc);
    if ((rc = pinned_alloc(priv, &r->payload, r->len + 8))) return bail(rc);
    e = launch_wire_frame(dense, ws->tok_off, enc, n, framed, v->sms, v->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->payload, framed, r->len, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    if (e != cudaSuccess) {
      g_err = std::string("wire tokenize D2H: ") + cudaGetErrorString(e);
      return bail(TMGPU_E_CUDA);
    }
    ws->last_launches++;
  }
  for (void* p : temp) cudaFree(p);
  r->status_code = (n && st0) ? 13 : 0;  // ERROR_NORMALIZATION_FAILED, results[0] only (server:379-381)
  *out = r;
  return TMGPU_OK;
}

extern "C" int tmgpu_wire_tokenize(tmgpu_vocab* v, const uint8_t* payload, uint64_t len, tmgpu_wire_reply** out) {
  return wire_tokenize_impl(v, payload, len, false, out);
}

extern "C" int tmgpu_wire_count(tmgpu_vocab* v, const uint8_t* payload, uint64_t len, tmgpu_wire_reply** out) {
  return wire_tokenize_impl(v, payload, len, true, out);
}

extern "C" int tmgpu_wire_decode(tmgpu_vocab* v, uint8_t enc_len, const uint8_t* payload, uint64_t len,
                                 tmgpu_wire_reply** out) {
  if (!v || !out) {
    g_err = "tmgpu_wire_decode: null argument";
    return TMGPU_E_ARG;
  }
  if (!(enc_len == 2 || enc_len == 3 || enc_len == 4)) {
    g_err = "invalid encoding length";
    return TMGPU_E_ARG;
  }
  uint32_t n = 0;
  uint64_t longest = 0;
  int rc = wire_scan(payload, len, &n, &longest);
  if (rc) return rc;
  if (longest > 0xFFFFFFFFull) {
    g_err = "a single entry may not exceed 4 GiB";
    return TMGPU_E_ARG;
  }
  std::lock_guard<std::mutex> lock(v->mu);
  if ((rc = ensure_host_path(v, n, len))) return rc;
  tmgpu_workspace* ws = v->ws;
  tmgpu_wire_reply* r = new_reply(v);
  auto* priv = static_cast<ResultPriv*>(r->priv);
  std::vector<void*> temp;
  auto bail = [&](int code) {
    for (void* p : temp) cudaFree(p);
    tmgpu_wire_reply_free(r);
    return code;
  };
  // entry s: whole ids only (go:845: len(b) >> 1 and its 3/4-byte siblings)
  uint64_t* h_src = nullptr;
  uint64_t* h_dst = nullptr;
  uint32_t* h_len = nullptr;
  if ((rc = pinned_alloc(priv, &h_src, size_t(n) + 1)) || (rc = pinned_alloc(priv, &h_dst, size_t(n) + 1)) ||
      (rc = pinned_alloc(priv, &h_len, size_t(n) + 1)))
    return bail(rc);
  uint64_t at = 4, ids = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t l = rd_u64(payload + at);
    at += 8;
    const uint64_t k = l / enc_len;
    h_src[i] = at;
    h_len[i] = uint32_t(k * enc_len);
    h_dst[i] = ids * enc_len;
    v->h_off[i] = int64_t(ids);
    ids += k;
    at += l;
  }
  v->h_off[n] = int64_t(ids);
  h_dst[n] = ids * enc_len;
  cudaStream_t st = v->stream;
  cudaError_t e = cudaMemcpyAsync(v->d_in, payload, len, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(v->d_off, v->h_off, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ws->out_off, h_src, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ws->raw_off, h_dst, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ws->out_len, h_len, (size_t(n) + 1) * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n)
    e = launch_compact_bytes(v->d_in, ws->out_off, ws->out_len, ws->raw_off, n, ws->tokens, v->sms, st);
  if (e != cudaSuccess) {
    g_err = std::string("wire decode staging: ") + cudaGetErrorString(e);
    return bail(TMGPU_E_CUDA);
  }
  const uint8_t* d_bytes = nullptr;
  const uint64_t* d_off = nullptr;
  uint64_t total = 0;
  if ((rc = decode_on_stream(ws, ws->tokens, v->d_off, n, enc_len, st, &d_bytes, &d_off, &total))) return bail(rc);
  r->len = 4 + 8ull * n + total;
  uint8_t* framed = reinterpret_cast<uint8_t*>(ws->slab);
  if (r->len + 16 > ws->slab_ids * 4 && (rc = dev_alloc(temp, &framed, r->len + 16))) return bail(rc);
  if ((rc = pinned_alloc(priv, &r->payload, r->len + 8))) return bail(rc);
  e = launch_wire_frame(d_bytes, d_off, 1, n, framed, v->sms, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r->payload, framed, r->len, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    g_err = std::string("wire decode D2H: ") + cudaGetErrorString(e);
    return bail(TMGPU_E_CUDA);
  }
  for (void* p : temp) cudaFree(p);
  r->status_code = 0;
  *out = r;
  return TMGPU_OK;
}

// ------------------------------------------------------------------------------------------
// streaming decoder: a batch of n independent Decoder objects (go/tokenmonster.go:450-740)
// ------------------------------------------------------------------------------------------
struct tmgpu_decoder {
  tmgpu_vocab* v = nullptr;
  uint32_t n = 0;
  uint64_t* rem_off = nullptr;   // device, n + 1: each stream's held-back bytes (go Decoder.remainder)
  uint8_t* rem = nullptr;        // device, dense
  uint64_t rem_cap = 0;
  uint8_t* cc_state = nullptr;   // device, n: capcode.Decoder flags (inWord, inChar, delete, ignore)
  uint32_t* keep_len = nullptr;  // device, n + 1
  uint32_t* rem_len = nullptr;   // device, n + 1
  uint64_t* new_off = nullptr;   // device, n + 1
  uint8_t* new_rem = nullptr;
  uint64_t new_cap = 0;
  std::vector<void*> allocs;
};

extern "C" void tmgpu_decoder_destroy(tmgpu_decoder* d) {
  if (!d) return;
  if (d->v && d->v->device >= 0) cudaSetDevice(d->v->device);
  for (void* p : d->allocs) cudaFree(p);
  if (d->rem) cudaFree(d->rem);
  if (d->new_rem) cudaFree(d->new_rem);
  delete d;
}

extern "C" int tmgpu_decoder_create(tmgpu_vocab* v, uint32_t n_streams, tmgpu_decoder** out) {
  if (!v || !out) {
    g_err = "tmgpu_decoder_create: null argument";
    return TMGPU_E_ARG;
  }
  if (v->device < 0) {
    g_err = "vocabulary was created without a device (this library has no CPU path)";
    return TMGPU_E_CUDA;
  }
  CU(cudaSetDevice(v->device));
  auto* d = new tmgpu_decoder();
  d->v = v;
  d->n = n_streams;
  const size_t n1 = size_t(n_streams) + 1;
  int rc = dev_alloc(d->allocs, &d->rem_off, n1);
  if (!rc) rc = dev_alloc(d->allocs, &d->cc_state, n1);
  if (!rc) rc = dev_alloc(d->allocs, &d->keep_len, n1);
  if (!rc) rc = dev_alloc(d->allocs, &d->rem_len, n1);
  if (!rc) rc = dev_alloc(d->allocs, &d->new_off, n1);
  if (rc) {
    tmgpu_decoder_destroy(d);
    return rc;
  }
  CU(cudaMemset(d->rem_off, 0, n1 * 8));
  CU(cudaMemset(d->cc_state, 0, n1));
  *out = d;
  return TMGPU_OK;
}

static int grow(uint8_t** p, uint64_t* cap, uint64_t need) {
  if (*cap >= need + 16) return TMGPU_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = need + need / 4 + 4096;
  CU(cudaMalloc(p, *cap));
  return TMGPU_OK;
}

// Decoder.Decode / Decoder.DecodeSerialized for every stream (go:701-740, :466-698): prepend the
// remainder, decode, hold back an incomplete UTF-8/16 tail, then the stateful capcode decode.
extern "C" int tmgpu_decoder_decode(tmgpu_decoder* d, const uint8_t* tokens, const uint64_t* tok_off, uint8_t enc_len,
                                    tmgpu_bytes** out) {
  if (!d || !tok_off || !out) {
    g_err = "tmgpu_decoder_decode: null argument";
    return TMGPU_E_ARG;
  }
  tmgpu_vocab* v = d->v;
  const uint32_t n = d->n;
  if (enc_len == 0) enc_len = v->host.n_reverse <= 65536 ? 2 : 3;
  if (!(enc_len == 2 || enc_len == 3 || enc_len == 4)) {
    g_err = "invalid encoding length";
    return TMGPU_E_ARG;
  }
  std::lock_guard<std::mutex> lock(v->mu);
  const uint64_t n_ids = tok_off[n] - tok_off[0];
  int rc = ensure_host_path(v, n, n_ids * enc_len + 64);
  if (rc) return rc;
  tmgpu_workspace* ws = v->ws;
  cudaStream_t st = v->stream;
  for (uint32_t i = 0; i <= n; i++) {
    if (tok_off[i] < tok_off[0] || (i && tok_off[i] < tok_off[i - 1])) {
      g_err = "token offsets must be non-decreasing";
      return TMGPU_E_ARG;
    }
    v->h_off[i] = int64_t(tok_off[i] - tok_off[0]);
  }
  if (n_ids) CU(cudaMemcpyAsync(v->d_in, tokens + tok_off[0] * enc_len, n_ids * enc_len, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(v->d_off, v->h_off, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, st));
  const bool stateful = v->host.charset != 0;  // charset 0 returns before remainder / capcode handling (quirk Q6)
  const uint64_t* prefix_off = stateful ? d->rem_off : nullptr;
  CU(cudaMemsetAsync(ws->raw_len + n, 0, 4, st));
  if (n) CU(launch_decode_size(v->dev, v->d_in, v->d_off, n, enc_len, ws->raw_len, v->sms, st, prefix_off));
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->raw_len, ws->raw_off, n + 1, st));
  uint64_t raw_total = 0;
  CU(cudaMemcpyAsync(&raw_total, ws->raw_off + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if ((rc = ensure_decode_cap(ws, raw_total))) return rc;
  if (n) CU(launch_decode_copy(v->dev, v->d_in, v->d_off, n, enc_len, ws->raw_off, ws->raw, v->sms, st, prefix_off, d->rem));
  const uint8_t* src = ws->raw;
  const uint32_t* lens = ws->raw_len;
  if (stateful && n) {
    CU(cudaMemsetAsync(d->rem_len + n, 0, 4, st));
    CU(launch_stream_tail(ws->raw, ws->raw_off, n, v->host.charset, d->keep_len, d->rem_len, v->sms, st));
    CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, d->rem_len, d->new_off, n + 1, st));
    uint64_t rem_total = 0;
    CU(cudaMemcpyAsync(&rem_total, d->new_off + n, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if ((rc = grow(&d->new_rem, &d->new_cap, rem_total))) return rc;
    CU(launch_copy_tail(ws->raw, ws->raw_off, d->keep_len, d->new_off, n, d->new_rem, v->sms, st));
    lens = d->keep_len;
    if (v->host.capcode == 2) {
      CU(launch_capcode_decode(ws->raw, ws->raw_off, n, ws->dec, ws->out_len, v->sms, st, d->keep_len, d->cc_state,
                               v->unicode ? &v->uni : nullptr));
      src = ws->dec;
      lens = ws->out_len;
    }
  }
  // dense output
  if (lens != ws->out_len && n) CU(cudaMemcpyAsync(ws->out_len, lens, size_t(n) * 4, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemsetAsync(ws->out_len + n, 0, 4, st));
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, ws->out_len, ws->out_off, n + 1, st));
  if (n)
    CU(launch_compact_bytes(src, ws->raw_off, ws->out_len, ws->out_off, n, ws->dense, v->sms, st,
                            (src == ws->dec && v->unicode) ? 1 : 0));
  uint64_t total = 0;
  CU(cudaMemcpyAsync(&total, ws->out_off + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (stateful && n) {  // the new remainders become the state
    std::swap(d->rem, d->new_rem);
    std::swap(d->rem_cap, d->new_cap);
    std::swap(d->rem_off, d->new_off);
  }
  auto* priv = new ResultPriv();
  priv->v = v;
  auto* b = new tmgpu_bytes();
  std::memset(b, 0, sizeof(*b));
  b->priv = priv;
  b->n = n;
  b->n_bytes = total;
  if ((rc = pinned_alloc(priv, &b->off, size_t(n) + 1)) || (rc = pinned_alloc(priv, &b->bytes, total + 8))) {
    tmgpu_bytes_free(b);
    return rc;
  }
  cudaError_t e = cudaMemcpyAsync(b->off, ws->out_off, (size_t(n) + 1) * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && total) e = cudaMemcpyAsync(b->bytes, ws->dense, total, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    g_err = std::string("decoder D2H: ") + cudaGetErrorString(e);
    tmgpu_bytes_free(b);
    return TMGPU_E_CUDA;
  }
  *out = b;
  return TMGPU_OK;
}

// Decoder.Flush (go:456): hand back every stream's remainder and clear it (the capcode state stays).
extern "C" int tmgpu_decoder_flush(tmgpu_decoder* d, tmgpu_bytes** out) {
  if (!d || !out) {
    g_err = "tmgpu_decoder_flush: null argument";
    return TMGPU_E_ARG;
  }
  tmgpu_vocab* v = d->v;
  std::lock_guard<std::mutex> lock(v->mu);
  CU(cudaSetDevice(v->device));
  const uint32_t n = d->n;
  auto* priv = new ResultPriv();
  priv->v = v;
  auto* b = new tmgpu_bytes();
  std::memset(b, 0, sizeof(*b));
  b->priv = priv;
  b->n = n;
  int rc = pinned_alloc(priv, &b->off, size_t(n) + 1);
  if (rc) {
    tmgpu_bytes_free(b);
    return rc;
  }
  cudaError_t e = cudaMemcpy(b->off, d->rem_off, (size_t(n) + 1) * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) {
    b->n_bytes = b->off[n];
    rc = pinned_alloc(priv, &b->bytes, b->n_bytes + 8);
    if (!rc && b->n_bytes) e = cudaMemcpy(b->bytes, d->rem, b->n_bytes, cudaMemcpyDeviceToHost);
  }
  if (rc || e != cudaSuccess) {
    if (e != cudaSuccess) g_err = std::string("decoder flush: ") + cudaGetErrorString(e);
    tmgpu_bytes_free(b);
    return rc ? rc : TMGPU_E_CUDA;
  }
  CU(cudaMemset(d->rem_off, 0, (size_t(n) + 1) * 8));
  *out = b;
  return TMGPU_OK;
}

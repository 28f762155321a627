// abi.cu -- library, vocabulary and workspace entry points of libtmgpu.so (include/tmgpu.h).
// Host orchestration only; all per-byte work is in kernels.cu.  There is no CPU compute path
// anywhere behind this ABI: without a CUDA device every compute entry fails with TMGPU_E_CUDA.
#include "abi_internal.h"

thread_local std::string g_err;

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
  if ((rc = upload(v, h.rev_pack.data(), h.rev_pack.size(), &d.rev_pack))) return fail(rc);
  {
    const uint64_t* b8 = nullptr;
    if ((rc = upload(v, h.rev_blob8.data(), h.rev_blob8.size(), &b8))) return fail(rc);
    d.rev_blob8 = reinterpret_cast<const unsigned long long*>(b8);
  }
  {
    const uint32_t* bk = nullptr;
    const uint32_t* hot = nullptr;
    if ((rc = upload(v, h.dfa_buckets.data(), h.dfa_buckets.size(), &bk))) return fail(rc);
    if ((rc = upload(v, h.dfa_hot.data(), h.dfa_hot.size(), &hot))) return fail(rc);
    const uint32_t* crec = nullptr;
    if ((rc = upload(v, h.chain_rec.data(), h.chain_rec.size(), &crec))) return fail(rc);
    d.dfa.crec = reinterpret_cast<const uint4*>(crec);
    d.dfa.buckets = reinterpret_cast<const uint4*>(bk);
    d.dfa.hot = reinterpret_cast<const uint4*>(hot);
    d.dfa.shift = 32u - h.dfa_bits;
    d.dfa.mask = (1u << h.dfa_bits) - 1u;
    d.dfa.n_states = h.dfa_states;
    std::vector<uint8_t> pimg(1536, 0);
    std::memcpy(pimg.data(), h.dfa_d1, 1024);
    std::memcpy(pimg.data() + 1024, h.class_code, 256);
    if ((rc = upload(v, pimg.data(), pimg.size(), &d.pipe_image))) return fail(rc);
  }
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
  v->copy_pool.reset();
  for (auto& b : v->stage_buf)
    if (b) cudaFreeHost(b);
  for (auto& e : v->stage_free)
    if (e) cudaEventDestroy(e);
  {  // results cut from the pool may still be alive: free what is idle, leave the rest to them
    std::lock_guard<std::mutex> lock(v->pool->mu);
    v->pool->orphaned = true;
    for (auto& b : v->pool->blocks)
      if (b.p && !b.used) {
        cudaFreeHost(b.p);
        b.p = nullptr;
      }
  }
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

extern "C" int64_t tmgpu_deserialize(const tmgpu_vocab* v, const uint8_t* data, uint64_t n_bytes, uint8_t enc_len,
                                     uint32_t* out) {
  if (enc_len == 0 && v) enc_len = v->host.n_reverse <= 65536 ? 2 : 3;   // go:751-757
  if (!(enc_len == 2 || enc_len == 3 || enc_len == 4) || (!data && n_bytes) || (!out && n_bytes >= enc_len)) {
    g_err = "tmgpu_deserialize: invalid encoding length or null buffer";
    return -1;
  }
  const uint64_t n = n_bytes / enc_len;
  for (uint64_t k = 0; k < n; k++) {
    const uint8_t* p = data + k * enc_len;
    uint32_t id = uint32_t(p[0]) | (uint32_t(p[1]) << 8);
    if (enc_len > 2) id |= uint32_t(p[2]) << 16;
    if (enc_len > 3) id |= uint32_t(p[3]) << 24;
    out[k] = id;
  }
  return int64_t(n);
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
  if (ws->p_text) cudaFree(ws->p_text);
  if (ws->p_words) cudaFree(ws->p_words);
  for (void* p : ws->ck_allocs) cudaFree(p);
  for (int i = 0; i < tmgpu_workspace::kEv; i++) {
    if (ws->ev0[i]) cudaEventDestroy(ws->ev0[i]);
    if (ws->ev1[i]) cudaEventDestroy(ws->ev1[i]);
  }
  for (auto& e : ws->stage_ev)
    if (e) cudaEventDestroy(e);
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
  A(&ws->p_nlen, n1);
  A(&ws->p_npad, n1);
  A(&ws->p_toff, n1);
  A(&ws->p_key, n1);
  A(&ws->p_key2, n1);
  A(&ws->p_val, n1);
  A(&ws->p_order, n1);
  ws->p_sort_bytes = sort_temp_bytes(uint32_t(n1)) + 256;
  A(reinterpret_cast<uint8_t**>(&ws->p_sort_temp), ws->p_sort_bytes);
  A(&ws->p_flag, n1);
  A(&ws->p_counters, 4);
  A(&ws->ov_list, n1);
  A(&ws->ov_begin, n1);
  A(&ws->ov_end, n1);
  A(&ws->ov_ctr, 4);   // non-ASCII {strings, scratch bytes}, overflowed {strings, words}
  {  // position words of a full batch: 1.25 x + 64 per string under capcode 2, else 1 x + 16
    const uint64_t words = v->host.capcode == 2 ? max_bytes + max_bytes / 4 + 64ull * max_strings + 4096
                                                : max_bytes + 16ull * max_strings + 4096;
    if (!rc) {
      cudaError_t e = tm_malloc(&ws->p_words, words * 4);
      if (e != cudaSuccess) {
        g_err = std::string("cudaMalloc(position words, ") + std::to_string(words * 4) + "): " + cudaGetErrorString(e);
        rc = TMGPU_E_NOMEM;
      } else {
        ws->p_words_cap = words;
      }
    }
  }
  A(&ws->overflow, 4);
  ws->scan_bytes = scan_temp_bytes(uint32_t(n1)) + 256;
  A(reinterpret_cast<uint8_t**>(&ws->scan_temp), ws->scan_bytes);
  if (v->unicode) {
    A(&ws->uni_list, n1);
    A(&ws->uni_begin, n1);
    A(&ws->uni_end, n1);
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
  for (auto& e : ws->stage_ev) CU(cudaEventCreate(&e));
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

// Durations (ms) of the stages of the most recent tokenize pipeline run on this workspace:
// out[0] prep + slot scan + length sort, out[1] encoder + position table, out[2] chain, out[3] 0.
// Returns 4, or 0 when nothing has run yet.  Synchronises on the last event.
extern "C" int tmgpu_workspace_stage_times(tmgpu_workspace* ws, float out[4]) {
  if (!ws || !out || !ws->stage_valid) return 0;
  if (cudaEventSynchronize(ws->stage_ev[3]) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  out[3] = 0.f;
  for (int k = 0; k < 3; k++) {
    out[k] = 0.f;
    if (cudaEventElapsedTime(&out[k], ws->stage_ev[k], ws->stage_ev[k + 1]) != cudaSuccess) {
      cudaGetLastError();
      out[k] = -1.f;
    }
  }
  return 4;
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

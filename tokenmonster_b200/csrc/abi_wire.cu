// abi_wire.cu -- the payloads of tokenmonsterserver's jobs, taken and produced as they travel.
#include "abi_internal.h"

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
  priv->pool = v->pool;
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
  if ((rc = tokenize_on_stream(ws, v->d_in, v->d_off, n, enc, count_only ? 1 : 0, v->stream, nullptr,
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

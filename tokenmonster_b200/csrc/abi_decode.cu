// abi_decode.cu -- decode entry points of libtmgpu.so: batched Vocab.Decode / DecodeSerialized and
// the streaming Decoder as a batch of streams with device-resident state.
#include "abi_internal.h"

// ------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------
int ensure_decode_cap(tmgpu_workspace* ws, uint64_t bytes) {
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
int decode_on_stream(tmgpu_workspace* ws, const uint8_t* d_tokens, const int64_t* d_tok_off, uint32_t n,
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
                             v->unicode ? &v->uni : nullptr,
                             ws->p_counters));
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
  for (uint32_t i = 1; i <= n; i++) {   // before any buffer is sized from them
    if (tok_off[i] < tok_off[i - 1]) {
      g_err = "token offsets must be non-decreasing";
      return TMGPU_E_ARG;
    }
  }
  const uint64_t n_ids = tok_off[n] - tok_off[0];
  int rc = ensure_host_path(v, n, n_ids * enc_len + 64);
  if (rc) return rc;
  for (uint32_t i = 0; i <= n; i++) v->h_off[i] = int64_t(tok_off[i] - tok_off[0]);
  if (n_ids)
    CU(cudaMemcpyAsync(v->d_in, tokens + tok_off[0] * enc_len, n_ids * enc_len, cudaMemcpyHostToDevice, v->stream));
  CU(cudaMemcpyAsync(v->d_off, v->h_off, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, v->stream));
  const uint8_t* d_bytes = nullptr;
  const uint64_t* d_off = nullptr;
  uint64_t total = 0;
  if ((rc = decode_on_stream(v->ws, v->d_in, v->d_off, n, enc_len, v->stream, &d_bytes, &d_off, &total))) return rc;
  auto* priv = new ResultPriv();
  priv->pool = v->pool;
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
  // the decoder's kernels run on v->stream (non-blocking: not ordered against the legacy stream)
  CU(cudaMemsetAsync(d->rem_off, 0, n1 * 8, v->stream));
  CU(cudaMemsetAsync(d->cc_state, 0, n1, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  *out = d;
  return TMGPU_OK;
}

static int grow(uint8_t** p, uint64_t* cap, uint64_t need) {
  if (*cap >= need + 16) return TMGPU_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = need + need / 4 + 4096;
  CU(tm_malloc(p, *cap));
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
  for (uint32_t i = 1; i <= n; i++) {   // before any buffer is sized from them
    if (tok_off[i] < tok_off[i - 1]) {
      g_err = "token offsets must be non-decreasing";
      return TMGPU_E_ARG;
    }
  }
  const uint64_t n_ids = tok_off[n] - tok_off[0];
  int rc = ensure_host_path(v, n, n_ids * enc_len + 64);
  if (rc) return rc;
  tmgpu_workspace* ws = v->ws;
  cudaStream_t st = v->stream;
  for (uint32_t i = 0; i <= n; i++) v->h_off[i] = int64_t(tok_off[i] - tok_off[0]);
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
                               v->unicode ? &v->uni : nullptr,
                             ws->p_counters));
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
  priv->pool = v->pool;
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
  priv->pool = v->pool;
  auto* b = new tmgpu_bytes();
  std::memset(b, 0, sizeof(*b));
  b->priv = priv;
  b->n = n;
  int rc = pinned_alloc(priv, &b->off, size_t(n) + 1);
  if (rc) {
    tmgpu_bytes_free(b);
    return rc;
  }
  cudaError_t e = cudaMemcpyAsync(b->off, d->rem_off, (size_t(n) + 1) * 8, cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
  if (e == cudaSuccess) {
    b->n_bytes = b->off[n];
    rc = pinned_alloc(priv, &b->bytes, b->n_bytes + 8);
    if (!rc && b->n_bytes) {
      e = cudaMemcpyAsync(b->bytes, d->rem, b->n_bytes, cudaMemcpyDeviceToHost, v->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    }
  }
  if (rc || e != cudaSuccess) {
    if (e != cudaSuccess) g_err = std::string("decoder flush: ") + cudaGetErrorString(e);
    tmgpu_bytes_free(b);
    return rc ? rc : TMGPU_E_CUDA;
  }
  CU(cudaMemsetAsync(d->rem_off, 0, (size_t(n) + 1) * 8, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  *out = b;
  return TMGPU_OK;
}

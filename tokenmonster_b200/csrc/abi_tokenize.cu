// abi_tokenize.cu -- tokenize / count / score entry points of libtmgpu.so: the device-resident
// path, the host-buffer path (chunked upload chased by one kernel launch), the passes for long
// documents (chunking.cuh) and non-ASCII strings (unicode_normalize.cuh), the overflow re-run.
#include "abi_internal.h"

#include <chrono>

// ---- the position-table pipeline (pipeline_kernels.cuh) ------------------------------------------
static int grow_device(void** p, uint64_t* cap, uint64_t want, size_t unit, const char* what) {
  if (*cap >= want) return TMGPU_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const uint64_t c = want + want / 8 + 4096;
  cudaError_t e = tm_malloc(p, c * unit);
  if (e != cudaSuccess) {
    g_err = std::string("cudaMalloc(") + what + ", " + std::to_string(c * unit) + "): " + cudaGetErrorString(e);
    return TMGPU_E_NOMEM;
  }
  *cap = c;
  return TMGPU_OK;
}

// Runs the stages for the work items of `a` (a.n <= ws->max_strings): prep + slot scan + length
// sort, encoder + position table, chain.  No host round trip.  `need_words` = a bound on the sum
// of the items' slots (0: a full batch, which the workspace was created for); the stream must be
// idle when the buffers have to grow (every caller with need_words != 0 has just synchronised).
static int run_pipeline(tmgpu_workspace* ws, TokenizeArgs a, bool store, cudaStream_t st, uint64_t need_words,
                        bool slot_full = false) {
  tmgpu_vocab* v = ws->v;
  if (a.n == 0) return TMGPU_OK;
  const bool cap = !a.prenorm && v->host.capcode == 2;
  const bool text = !store && a.scores != nullptr;
  int rc = grow_device(reinterpret_cast<void**>(&ws->p_words), &ws->p_words_cap, need_words, 4, "position words");
  if (!rc && text) {
    if (ws->p_text_cap < ws->p_words_cap) CU(cudaStreamSynchronize(st));
    rc = grow_device(reinterpret_cast<void**>(&ws->p_text), &ws->p_text_cap, ws->p_words_cap, 1, "normalized text");
  }
  if (rc) return rc;
  a.text = text ? ws->p_text : nullptr;
  a.words = ws->p_words;
  a.words_cap = ws->p_words_cap;
  a.nlen = ws->p_nlen;
  a.npad = ws->p_npad;
  a.toff = ws->p_toff;
  a.sort_key = ws->p_key;
  a.sort_val = ws->p_val;
  a.sort_order = ws->p_order;
  a.pflag = ws->p_flag;
  a.pipe_counters = ws->p_counters;
  CU(cudaMemsetAsync(a.pipe_counters, 0, 16, st));
  ws->stage_valid = true;
  CU(cudaEventRecord(ws->stage_ev[0], st));
  CU(launch_prep(a, cap, slot_full, v->sms, st));
  CU(cudaMemsetAsync(a.npad + a.n, 0, 4, st));
  CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, a.npad, a.toff, a.n + 1, st));
  CU(sort_items_by_length(ws->p_sort_temp, ws->p_sort_bytes, ws->p_key, ws->p_key2, ws->p_val, ws->p_order, a.n, st));
  CU(cudaEventRecord(ws->stage_ev[1], st));
  CU(launch_table(a, cap, text, v->sms, st));
  CU(cudaEventRecord(ws->stage_ev[2], st));
  CU(launch_chain(a, store, v->sms, st));
  CU(cudaEventRecord(ws->stage_ev[3], st));
  ws->last_launches += 8;   // prep, scan x2, sort (histogram + onesweep passes ~ 3), table, chain
  return TMGPU_OK;
}

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
    int rc = run_pipeline(ws, b, mode == 0, st, 0);
    if (rc) return rc;
  }
  return TMGPU_OK;
}

// After the main launch: strings the table kernel refused at a byte >= 0x80 (status 13) are
// normalized by the device normalizer and run through the pipeline again, reading the normalized
// text verbatim; strings whose capcode text outgrew the optimistic slot (status 14) are run again
// with full-size slots.  One 32-byte D2H tells the host whether there are any of either; plain
// batches pay only that.
static int side_passes(tmgpu_workspace* ws, const TokenizeArgs& first, bool store, cudaStream_t st) {
  tmgpu_vocab* v = ws->v;
  ws->last_unicode = 0;
  ws->last_overflowed = 0;
  const bool need_uni = v->unicode, need_ov = v->host.capcode == 2;
  if ((!need_uni && !need_ov) || first.n == 0 || v->host.max_len == 0) return TMGPU_OK;
  unsigned long long* uni_ctr = ws->ov_ctr;        // {strings, scratch bytes}
  unsigned long long* ov_ctr = ws->ov_ctr + 2;     // {strings, words}
  CU(cudaMemsetAsync(ws->ov_ctr, 0, 32, st));
  if (need_uni)
    CU(launch_select_refused(first.status, first.off, first.ends, first.n, ws->uni_list, ws->uni_begin, uni_ctr, v->sms, st));
  if (need_ov)
    CU(launch_select_overflowed(first.status, first.off, first.ends, first.n, ws->ov_list, ws->ov_begin, ws->ov_end, ov_ctr,
                                v->sms, st));
  unsigned long long h[4] = {0, 0, 0, 0};
  CU(cudaMemcpyAsync(h, ws->ov_ctr, 32, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  ws->last_launches += (need_uni ? 1 : 0) + (need_ov ? 1 : 0);
  if (h[2]) {   // full-size slots for the strings that outgrew theirs
    TokenizeArgs b = first;
    b.off = ws->ov_begin;
    b.ends = ws->ov_end;
    b.n = uint32_t(h[2]);
    b.list = ws->ov_list;
    b.raw_off = first.off;
    b.raw_ends = first.ends;
    b.skip = nullptr;
    int rc = run_pipeline(ws, b, store, st, h[3] + 64, true);
    if (rc) return rc;
    ws->last_overflowed = uint32_t(h[2]);
    if (need_uni) {   // a redone string may turn out to hold a byte >= 0x80: list the refused ones again
      CU(cudaMemsetAsync(uni_ctr, 0, 16, st));
      CU(launch_select_refused(first.status, first.off, first.ends, first.n, ws->uni_list, ws->uni_begin, uni_ctr, v->sms, st));
      CU(cudaMemcpyAsync(h, uni_ctr, 16, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
    }
  }
  if (h[0] == 0) return TMGPU_OK;
  if (ws->uni_text_cap < h[1] + 64) {
    if (ws->uni_text) cudaFree(ws->uni_text);
    ws->uni_text = nullptr;
    ws->uni_text_cap = h[1] + h[1] / 4 + 4096;
    cudaError_t e = tm_malloc(&ws->uni_text, ws->uni_text_cap);
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
  na.n_list = reinterpret_cast<const uint32_t*>(uni_ctr);
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
  {
    int rc = run_pipeline(ws, b, store, st, h[1] + 16ull * h[0] + 64);
    if (rc) return rc;
  }
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
  ws->last_launches += 1;
  ws->last_unicode = uint32_t(h[0]);
  return TMGPU_OK;
}

int tokenize_on_stream(tmgpu_workspace* ws, const uint8_t* d_data, const int64_t* d_off, uint32_t n,
                              uint8_t enc_len, int mode, cudaStream_t st, const uint64_t* slab_off,
                              uint32_t* slab, uint8_t* dense, const std::vector<ChunkPlan>* chunks,
                              const int64_t* d_ends, StreamOut* so) {
  tmgpu_vocab* v = ws->v;
  ws->last_launches = 0;
  // ids stream out chunk by chunk only on the plain chunked host path
  const bool stream_out = so && mode == 0 && chunks && chunks->size() > 1 && !slab_off && n > 0 && v->host.max_len != 0;
  auto stage_all = [&]() -> int {   // upload whatever the plan has not uploaded yet
    for (const ChunkPlan& c : *chunks)
      if (c.stage) {
        int rc = c.stage();
        const_cast<ChunkPlan&>(c).stage = nullptr;
        if (rc) return rc;
      }
    return TMGPU_OK;
  };
  int streamed_chunks = 0;
  if (so) so->ok = false;
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
      if (chunks) {
        int rc = stage_all();
        if (rc) return rc;
        for (const ChunkPlan& c : *chunks)
          if (c.ready) CU(cudaStreamWaitEvent(st, c.ready, 0));
      }
      CU(launch_flag_non_ascii(d_data, d_off, d_ends, n, ws->ck_skip, ws->status, ws->counts, ws->missing, v->sms, st));
      a.skip = ws->ck_skip;
      ws->last_launches++;
    }
  }
  uint32_t ck_total = 0;   // long documents of a small batch are tokenized in chunks (chunking.cuh)
  if (n > 0 && (!chunks || chunks->size() == 1)) {
    if (chunks) {
      int rc = stage_all();
      if (rc) return rc;
    }
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
    // One compute stream: alternating chunks between two streams (so that the tail of chunk k
    // overlaps the head of chunk k+1) was measured SLOWER -- 29.5 ms against 16.0 ms per 512 MB call
    // on the same box: two persistent kernels fight over the SM slots and the chunk that is needed
    // first (its ids are waited for) finishes later.
    for (const ChunkPlan& c : *chunks) {
      if (c.stage) {   // (not yet uploaded: pageable input is staged chunk by chunk)
        int rc = c.stage();
        const_cast<ChunkPlan&>(c).stage = nullptr;
        if (rc) return rc;
      }
      if (c.s1 <= c.s0) continue;
      cudaStream_t cs = st;
      if (c.ready) CU(cudaStreamWaitEvent(cs, c.ready, 0));
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
      int rc = run_pipeline(ws, b, mode == 0, cs, 0);
      if (rc) return rc;
      if (stream_out) {
        // this chunk's offsets, dense ids and running id total, while the next chunk is still copying in
        const int ci = streamed_chunks++;
        if (ci == 0) CU(cudaMemsetAsync(v->d_run, 0, 8, cs));
        CU(exclusive_scan_u32_to_u64(ws->scan_temp, ws->scan_bytes, b.counts, ws->tok_off + c.s0, b.n, cs));
        CU(launch_chunk_offsets(ws->tok_off + c.s0, b.counts, b.n, reinterpret_cast<unsigned long long*>(v->d_run + ci),
                                c.s1 == n, v->sms, cs));
        CU(launch_pack(b.slab, b.off, ws->slab_pad, nullptr, b.counts, ws->tok_off + c.s0, dense, b.n, enc_len, v->sms, cs));
        CU(cudaMemcpyAsync(v->h_info + ci + 1, v->d_run + ci + 1, 8, cudaMemcpyDeviceToHost, cs));
        CU(cudaEventRecord(v->done_ev[ci], cs));
        ws->last_launches += 4;
      }
    }
    if (stream_out) {   // everything is queued: hand each chunk's ids to the copy-out stream as it completes
      bool fits = true;
      uint64_t prev = 0;
      for (int ci = 0; ci < streamed_chunks; ci++) {
        CU(cudaEventSynchronize(v->done_ev[ci]));
        const uint64_t cum = v->h_info[ci + 1];
        if (cum * enc_len > so->cap_bytes) fits = false;
        if (fits && cum > prev)
          CU(cudaMemcpyAsync(so->tokens + prev * enc_len, dense + prev * enc_len, (cum - prev) * enc_len,
                             cudaMemcpyDeviceToHost, v->d2h_stream));
        prev = cum;
      }
      so->ok = fits;
      so->total_ids = prev;
    }
  } else if (n > 0) {
    const int e = ws->ev_head;
    CU(cudaEventRecord(ws->ev0[e], st));
    {
      int rc = run_pipeline(ws, a, mode == 0, st, 0);
      if (rc) return rc;
    }
    CU(cudaEventRecord(ws->ev1[e], st));
    ws->ev_head = (e + 1) % tmgpu_workspace::kEv;
    if (ws->ev_count < tmgpu_workspace::kEv) ws->ev_count++;
  }
  if (ck_total) {
    int rc = chunk_run(ws, a, ck_total, mode, st);
    if (rc) return rc;
    a.skip = nullptr;
  }
  {
    int rc = side_passes(ws, a, mode == 0, st);
    if (rc) return rc;
  }
  if (so && so->ok && (ws->last_unicode || ws->last_overflowed || ck_total)) so->ok = false;   // counts changed after the chunks went out
  if (mode == 0 && !(stream_out && so->ok)) {
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
int ensure_host_path(tmgpu_vocab* v, uint32_t n, uint64_t bytes) {
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
    CU(tm_malloc(&v->d_in, v->d_in_cap));
  }
  if (v->d_off_cap < size_t(n) + 1) {
    if (v->d_off) cudaFree(v->d_off);
    if (v->h_off) cudaFreeHost(v->h_off);
    v->d_off = nullptr;
    v->h_off = nullptr;
    v->d_off_cap = size_t(n) + n / 8 + 1024;
    CU(tm_malloc(&v->d_off, v->d_off_cap * 16));
    CU(cudaMallocHost(&v->h_off, v->d_off_cap * 16));
  }
  return TMGPU_OK;
}

void release_priv(ResultPriv* p) {
  if (!p) return;
  if (p->pool) {
    std::lock_guard<std::mutex> lock(p->pool->mu);
    for (void* q : p->blocks)
      for (auto& b : p->pool->blocks)
        if (b.p == q) {
          b.used = false;
          if (p->pool->orphaned) {   // nobody will ask for it again
            cudaFreeHost(b.p);
            b.p = nullptr;
          }
        }
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
int stage_input(tmgpu_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n, uint64_t* total,
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
int stage_copy(tmgpu_vocab* v, const uint8_t* src, uint32_t n, uint64_t total_bytes, bool gapped,
                      std::vector<ChunkPlan>* plan) {
  const uint64_t* total = &total_bytes;
  CU(cudaMemcpyAsync(v->d_off, v->h_off, (size_t(n) + 1) * 8, cudaMemcpyHostToDevice, v->copy_stream));
  if (gapped && n)
    CU(cudaMemcpyAsync(v->d_off + v->d_off_cap, v->h_off + v->d_off_cap, size_t(n) * 8, cudaMemcpyHostToDevice,
                       v->copy_stream));
  CU(cudaEventRecord(v->chunk_ev[31], v->copy_stream));
  // Chunks of ~equal bytes: the kernels of chunk k run while chunk k+1 is still on the wire.  One
  // lane works on one string and a warp's 32 strings end together, so a chunk must be more than a
  // "wave" of the GPU (1024 resident lanes per SM) for the longest-first order to balance it, while
  // the call ends with the last chunk's kernels and ids, which want it small.  Measured per 512 MB
  // call (1 M strings, pinned input) with 2048 / 1536 / 1280 / 1024 strings per SM and chunk, i.e.
  // 3 / 4 / 5 / 6 chunks: 15.8 / 15.4 / 14.95 ms on one box, 15.4 (1280) / 16.5-17.1 (1024) on another.
  static const uint32_t lanes_per_sm = getenv("TMGPU_CHUNK_LANES") ? uint32_t(atoi(getenv("TMGPU_CHUNK_LANES"))) : 1280u;   // diagnostic override
  int want = int(n / (uint32_t(v->sms) * (lanes_per_sm ? lanes_per_sm : 1280u)));
  if (want > 8) want = 8;
  if (want < 1) want = 1;
  if (plan) plan->clear();
  // is the source ordinary (pageable) memory?  Then big batches are staged through pinned memory
  bool pageable = false;
  if (plan && total_bytes >= (8u << 20)) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, src) != cudaSuccess) {
      cudaGetLastError();
      pageable = true;
    } else {
      pageable = attr.type == cudaMemoryTypeUnregistered;
    }
    if (pageable && !v->copy_pool) {
      for (int k = 0; k < tmgpu_vocab::kStageSlots; k++) {
        CU(cudaMallocHost(&v->stage_buf[k], tmgpu_vocab::kStageBytes));
        CU(cudaEventCreateWithFlags(&v->stage_free[k], cudaEventDisableTiming));
      }
      unsigned hw = std::thread::hardware_concurrency();
      v->copy_pool.reset(new CopyPool(hw >= 16 ? 7 : (hw >= 4 ? 3 : 1)));
    }
  }
  uint64_t piece = 0;
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
    cudaEvent_t ready = v->chunk_ev[c];
    if (!pageable || !plan) {
      if (b1 > b0) CU(cudaMemcpyAsync(v->d_in + b0, src + b0, b1 - b0, cudaMemcpyHostToDevice, v->copy_stream));
      CU(cudaEventRecord(ready, v->copy_stream));
      if (plan) plan->push_back({s0, s1, (b0 - uint64_t(v->h_off[0])) + uint64_t(pad) * s0, ready, nullptr});
    } else {
      // pageable source: helper threads copy 16 MB pieces into a pinned ring, each piece goes up as
      // soon as it is staged; done lazily, chunk by chunk, from the launch loop
      auto stage = [v, src, b0, b1, ready, piece]() mutable -> int {
        for (uint64_t at = b0; at < b1; at += tmgpu_vocab::kStageBytes) {
          const uint64_t len = b1 - at < tmgpu_vocab::kStageBytes ? b1 - at : tmgpu_vocab::kStageBytes;
          const int slot = int(piece++ % tmgpu_vocab::kStageSlots);
          CU(cudaEventSynchronize(v->stage_free[slot]));
          v->copy_pool->copy(v->stage_buf[slot], src + at, len);
          CU(cudaMemcpyAsync(v->d_in + at, v->stage_buf[slot], len, cudaMemcpyHostToDevice, v->copy_stream));
          CU(cudaEventRecord(v->stage_free[slot], v->copy_stream));
        }
        CU(cudaEventRecord(ready, v->copy_stream));
        return TMGPU_OK;
      };
      piece += (b1 - b0 + tmgpu_vocab::kStageBytes - 1) / tmgpu_vocab::kStageBytes;
      plan->push_back({s0, s1, (b0 - uint64_t(v->h_off[0])) + uint64_t(pad) * s0, ready, stage});
    }
    s0 = s1;
  }
  if (!plan) CU(cudaStreamSynchronize(v->copy_stream));
  return TMGPU_OK;
}

// If a string produced more ids than its slab share (len + 16), run again with exact sizes.
// On success *dense points at the serialized ids to copy out; temp holds what to cudaFree.
int rerun_if_overflow(tmgpu_vocab* v, uint32_t n, uint8_t enc_len, std::vector<void*>* temp,
                             uint8_t** dense, const int64_t* d_ends) {
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
  static const bool trace = getenv("TMGPU_TRACE") != nullptr;   // diagnostic: host-side phase times on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  int rc = ensure_host_path(v, n, off[n] - off[0]);
  if (rc) return rc;
  uint64_t total = 0;
  std::vector<ChunkPlan> plan;
  if ((rc = stage_input(v, data, off, n, &total, &plan))) return rc;
  const double t_staged = since();
  tmgpu_workspace* ws = v->ws;
  auto* priv = new ResultPriv();
  priv->pool = v->pool;
  auto* r = new tmgpu_result();
  std::memset(r, 0, sizeof(*r));
  r->priv = priv;
  r->n = n;
  r->enc_len = enc_len;
  std::vector<void*> temp;
  auto cleanup = [&]() {
    for (void* p : temp) cudaFree(p);
    temp.clear();
  };
  auto bail = [&](int code) {
    cleanup();
    cudaStreamSynchronize(v->d2h_stream);
    tmgpu_result_free(r);
    return code;
  };
  // Big batches: the ids leave the GPU chunk by chunk while later chunks are still being tokenized.
  // The pinned destination has to exist before the id count is known: room for what the previous
  // call produced per input byte (+ 12 %), else one id per 3 bytes; a batch that needs more falls
  // back to one copy at the end.
  StreamOut so;
  if (plan.size() > 1) {
    const double per_byte = v->ids_per_byte > 0 ? v->ids_per_byte * 1.12 : 1.0 / 3.0;
    so.cap_bytes = (uint64_t(double(total) * per_byte) + n + 4096) * enc_len;
    if ((rc = pinned_alloc(priv, &so.tokens, so.cap_bytes))) return bail(rc);
  }
  if ((rc = tokenize_on_stream(ws, v->d_in, v->d_off, n, enc_len, 0, v->stream, nullptr, ws->slab, ws->tokens, &plan, nullptr,
                               so.tokens ? &so : nullptr)))
    return bail(rc);
  const double t_queued = since();
  uint8_t* dense = ws->tokens;
  if ((rc = rerun_if_overflow(v, n, enc_len, &temp, &dense))) return bail(rc);
  const double t_computed = since();
  const bool streamed = so.tokens && so.ok && dense == ws->tokens;
  if ((rc = pinned_alloc(priv, &r->tok_off, size_t(n) + 1))) return bail(rc);
  if ((rc = pinned_alloc(priv, &r->missing, size_t(n) + 1))) return bail(rc);
  if ((rc = pinned_alloc(priv, &r->status, size_t(n) + 1))) return bail(rc);
  cudaError_t e;
  e = cudaMemcpyAsync(r->tok_off, ws->tok_off, (size_t(n) + 1) * 8, cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r->missing, ws->missing, size_t(n) * 4, cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r->status, ws->status, size_t(n), cudaMemcpyDeviceToHost, v->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v->d2h_stream);
  if (e != cudaSuccess) {
    g_err = std::string("tokenize: ") + cudaGetErrorString(e);
    return bail(TMGPU_E_CUDA);
  }
  r->n_tokens = r->tok_off[n];
  if (total) v->ids_per_byte = double(r->n_tokens) / double(total);
  if (streamed) {
    r->tokens = so.tokens;
  } else {
    if ((rc = pinned_alloc(priv, &r->tokens, r->n_tokens * enc_len + 8))) return bail(rc);
    if (r->n_tokens) {
      e = cudaMemcpyAsync(r->tokens, dense, r->n_tokens * enc_len, cudaMemcpyDeviceToHost, v->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
      if (e != cudaSuccess) {
        g_err = std::string("tokenize D2H: ") + cudaGetErrorString(e);
        return bail(TMGPU_E_CUDA);
      }
    }
  }
  cleanup();
  if (trace)
    fprintf(stderr, "[tmgpu] tokenize_batch: %u strings, %.1f MB, %zu chunk(s)%s: offsets + copies queued %.2f ms, kernels queued / "
            "chunks drained %.2f ms, compute done %.2f ms, results on the host %.2f ms\n", n, double(total) / 1e6, plan.size(),
            streamed ? " (ids streamed out)" : "", t_staged, t_queued, t_computed, since());
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

// ref_driver.cpp -- C ABI around the REFERENCE's own C++ runtime.  TEST INFRASTRUCTURE ONLY.
//
// oracle/Makefile compiles this file together with
//   /root/reference/tokenmonster-cpp/src/tokenmonster.cpp   (unmodified, in place)
// against oracle/shims/ into oracle/_ref/libtmref.so.  No reference source is copied
// into this repository.  The driver exposes the reference's public API
// (tokenmonster-cpp/include/tokenmonster/tokenmonster.hpp:52-112) plus the private
// tokenize_normalized (hpp:147) the reference's own tests/bench.cpp:13-15 reaches the same way.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#define private public
#include <tokenmonster/tokenmonster.hpp>
#undef private

namespace tkm = tokenmonster;

struct tmr_vocab {
  tkm::Vocab v;
};
struct tmr_decoder {
  tkm::Decoder d;
};

static thread_local std::string g_err;

template <class T>
static T* dup_out(const std::vector<T>& v) {
  T* p = static_cast<T*>(std::malloc(v.size() * sizeof(T) + 1));
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

extern "C" {

const char* tmr_last_error() { return g_err.c_str(); }
void tmr_free(void* p) { std::free(p); }

tmr_vocab* tmr_vocab_load(const uint8_t* file, size_t n, int* status) {
  char path[] = "/tmp/tmref_vocab_XXXXXX";
  int fd = mkstemp(path);
  if (fd < 0) { *status = 1; return nullptr; }
  size_t done = 0;
  while (done < n) {
    ssize_t w = write(fd, file + done, n - done);
    if (w <= 0) break;
    done += static_cast<size_t>(w);
  }
  close(fd);
  tmr_vocab* out = nullptr;
  try {
    out = new tmr_vocab{tkm::Vocab::load(path)};
    *status = 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    *status = std::strstr(e.what(), "order") ? 2 : 1;
  }
  unlink(path);
  return out;
}

void tmr_vocab_free(tmr_vocab* v) { delete v; }

int tmr_vocab_info(const tmr_vocab* v, uint32_t out[12]) {
  const auto& x = v->v;
  out[0] = x.using_capcode_; out[1] = x.charset_; out[2] = x.normalizer_flag_; out[3] = x.level_;
  out[4] = x.reserve_; out[5] = x.unk_token_; out[6] = static_cast<uint32_t>(x.vocab_size_);
  out[7] = static_cast<uint32_t>(x.reverse_.size()); out[8] = static_cast<uint32_t>(x.info_.size());
  out[9] = x.delete_token_; out[10] = static_cast<uint32_t>(x.max_token_length_); out[11] = 0;
  return 0;
}

int tmr_normalize(const tmr_vocab* v, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n) {
  try {
    auto r = v->v.normalize({in, n});
    *out = dup_out(r);
    *out_n = r.size();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 13;
  }
}

// mode: 0 ids, 1 count, 2/3/4 serialized.  normalized != 0 -> tokenize_normalized (ids only).
int tmr_tokenize(const tmr_vocab* v, const uint8_t* in, size_t n, int normalized, int mode,
                 void** out, size_t* n_tokens, int64_t* missing) {
  *out = nullptr;
  *n_tokens = 0;
  *missing = 0;
  try {
    std::span<const uint8_t> s{in, n};
    if (mode == 0) {
      auto r = normalized ? v->v.tokenize_normalized(s) : v->v.tokenize(s);
      *out = dup_out(r.tokens);
      *n_tokens = r.tokens.size();
      *missing = r.missing;
    } else if (mode == 1) {
      auto r = normalized ? v->v.tokenize_count_normalized(s) : v->v.count(s);
      *n_tokens = static_cast<size_t>(r.tokens);
      *missing = r.missing;
    } else {
      if (normalized) {
        auto r = mode == 2   ? v->v.tokenize_to_serialized16(s)
                 : mode == 3 ? v->v.tokenize_to_serialized24(s)
                             : v->v.tokenize_to_serialized32(s);
        *out = dup_out(r.first);
        *n_tokens = r.first.size() / static_cast<size_t>(mode);
        *missing = r.second;
      } else {
        auto r = v->v.tokenize_serialized(s, static_cast<uint8_t>(mode));
        *out = dup_out(r.bytes);
        *n_tokens = r.bytes.size() / static_cast<size_t>(mode);
        *missing = r.missing;
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 13;
  }
}

int tmr_decode(const tmr_vocab* v, const void* tokens, size_t count, int enc_len, int raw,
               uint8_t** out, size_t* out_n) {
  try {
    std::vector<uint8_t> r;
    if (enc_len == 0) {
      std::span<const uint32_t> s{static_cast<const uint32_t*>(tokens), count};
      r = raw ? v->v.decode_raw(s) : v->v.decode(s);
    } else {
      std::span<const uint8_t> s{static_cast<const uint8_t*>(tokens), count};
      r = raw ? v->v.decode_serialized_raw(s, static_cast<uint8_t>(enc_len))
              : v->v.decode_serialized(s, static_cast<uint8_t>(enc_len));
    }
    *out = dup_out(r);
    *out_n = r.size();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 13;
  }
}

tmr_decoder* tmr_decoder_new(const tmr_vocab* v) { return new tmr_decoder{v->v.new_decoder()}; }
void tmr_decoder_free(tmr_decoder* d) { delete d; }

int tmr_decoder_decode(tmr_decoder* d, const void* tokens, size_t count, int enc_len,
                       uint8_t** out, size_t* out_n) {
  try {
    std::vector<uint8_t> r;
    if (enc_len == 0)
      r = d->d.decode({static_cast<const uint32_t*>(tokens), count});
    else
      r = d->d.decode_serialized({static_cast<const uint8_t*>(tokens), count},
                                 static_cast<uint8_t>(enc_len));
    *out = dup_out(r);
    *out_n = r.size();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 13;
  }
}

int tmr_decoder_flush(tmr_decoder* d, uint8_t** out, size_t* out_n) {
  auto r = d->d.flush();
  *out = dup_out(r);
  *out_n = r.size();
  return 0;
}

// Batch tokenize on `threads` host threads, one contiguous slice of strings per thread --
// the analogue of the server's goroutine-per-item loop (training/tokenmonsterserver.go:364-377).
// Used only as the timed CPU baseline.  Returns total ids emitted; per-string counts optional.
int64_t tmr_tokenize_batch(const tmr_vocab* v, const uint8_t* data, const uint64_t* off, uint32_t n,
                           int threads, uint32_t* counts, uint64_t* checksum) {
  if (threads < 1) threads = 1;
  std::vector<int64_t> tot(static_cast<size_t>(threads), 0);
  std::vector<uint64_t> chk(static_cast<size_t>(threads), 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      uint32_t lo = static_cast<uint32_t>(static_cast<uint64_t>(n) * t / threads);
      uint32_t hi = static_cast<uint32_t>(static_cast<uint64_t>(n) * (t + 1) / threads);
      for (uint32_t s = lo; s < hi; s++) {
        auto r = v->v.tokenize({data + off[s], static_cast<size_t>(off[s + 1] - off[s])});
        tot[t] += static_cast<int64_t>(r.tokens.size());
        if (counts) counts[s] = static_cast<uint32_t>(r.tokens.size());
        // per-string FNV-1a over the ids (fold of tokenmonster-cpp/tests/bench.cpp:30-37), mixed with
        // the string index and SUMMED: the same value whatever the thread count or the order
        uint64_t h = 14695981039346656037ULL;
        for (auto id : r.tokens) h = (h ^ id) * 1099511628211ULL;
        chk[t] += h ^ (static_cast<uint64_t>(s) * 0x9E3779B97F4A7C15ULL);
      }
    });
  }
  for (auto& th : pool) th.join();
  int64_t total = 0;
  uint64_t c = 0;
  for (int t = 0; t < threads; t++) { total += tot[t]; c += chk[t]; }
  if (checksum) *checksum = c;
  return total;
}

}  // extern "C"

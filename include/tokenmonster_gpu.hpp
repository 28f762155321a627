// tokenmonster_gpu.hpp -- header-only C++ mirror of the reference's C++ runtime API
// (tokenmonster-cpp/include/tokenmonster/tokenmonster.hpp:52-112) on top of the C ABI in tmgpu.h.
// Same method names, argument meaning and error behaviour (tokenmonster::Error is thrown where the
// reference throws), but batch-first: every single-string method has a *_batch sibling, which is
// what the GPU path is for.  Link with -ltmgpu.
#pragma once

#include <cstdint>
#include <filesystem>
#include <span>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tmgpu.h"

namespace tokenmonster_gpu {

class Error : public std::runtime_error {
 public:
  Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

struct TokenizeResult {  // hpp:26-29
  std::vector<std::uint32_t> tokens;
  int missing = 0;
};
struct CountResult {  // hpp:31-34
  int tokens = 0;
  int missing = 0;
};
struct SerializedResult {  // hpp:36-40
  std::vector<std::uint8_t> bytes;
  std::uint8_t encoding_length = 0;
  int missing = 0;
};

inline void check(int rc) {
  if (rc != TMGPU_OK) throw Error(rc, tmgpu_last_error());
}

class Vocab;

// tokenmonster::Decoder (hpp:52-70) for n parallel streams.
class Decoder {
 public:
  Decoder(Decoder&& o) noexcept : h_(std::exchange(o.h_, nullptr)), n_(o.n_) {}
  Decoder(const Decoder&) = delete;
  ~Decoder() { tmgpu_decoder_destroy(h_); }

  // chunks[s] = the next ids of stream s
  std::vector<std::vector<std::uint8_t>> decode(const std::vector<std::vector<std::uint32_t>>& chunks) {
    if (chunks.size() != n_) throw Error(TMGPU_E_ARG, "one chunk per stream");
    std::vector<std::uint32_t> flat;
    std::vector<std::uint64_t> off(n_ + 1, 0);
    for (std::size_t s = 0; s < n_; s++) {
      flat.insert(flat.end(), chunks[s].begin(), chunks[s].end());
      off[s + 1] = flat.size();
    }
    tmgpu_bytes* b = nullptr;
    check(tmgpu_decoder_decode(h_, reinterpret_cast<const std::uint8_t*>(flat.data()), off.data(), 4, &b));
    return take(b);
  }
  std::vector<std::vector<std::uint8_t>> flush() {
    tmgpu_bytes* b = nullptr;
    check(tmgpu_decoder_flush(h_, &b));
    return take(b);
  }

 private:
  friend class Vocab;
  Decoder(tmgpu_decoder* h, std::uint32_t n) : h_(h), n_(n) {}
  std::vector<std::vector<std::uint8_t>> take(tmgpu_bytes* b) {
    std::vector<std::vector<std::uint8_t>> out(n_);
    for (std::size_t s = 0; s < n_; s++) out[s].assign(b->bytes + b->off[s], b->bytes + b->off[s + 1]);
    tmgpu_bytes_free(b);
    return out;
  }
  tmgpu_decoder* h_ = nullptr;
  std::uint32_t n_ = 0;
};

class Vocab {
 public:
  Vocab(Vocab&& o) noexcept : h_(std::exchange(o.h_, nullptr)), info_(o.info_) {}
  Vocab(const Vocab&) = delete;
  ~Vocab() { tmgpu_vocab_destroy(h_); }

  // tokenmonster::Vocab::load (hpp:81, cpp:1287)
  static Vocab load(const std::filesystem::path& path, int device = 0) {
    tmgpu_vocab* h = nullptr;
    check(tmgpu_vocab_load_file(path.string().c_str(), device, &h));
    return Vocab(h);
  }

  int size() const { return int(info_.vocab_size); }
  int max_token_length() const { return info_.max_token_length; }
  std::uint8_t charset() const { return info_.charset; }
  std::uint8_t capcode() const { return info_.capcode; }
  std::uint8_t mode() const { return info_.level; }
  std::uint8_t normalization_code() const { return info_.normflag; }
  bool has_unk() const { return info_.unk_token != TMGPU_DNE; }
  std::uint32_t unk() const { return info_.unk_token; }
  int highest_token_id() const { return int(info_.n_reverse) - 1; }

  // ---- batch entry points -----------------------------------------------------------------
  std::vector<TokenizeResult> tokenize_batch(const std::vector<std::span<const std::uint8_t>>& texts) const {
    auto [blob, off] = flatten(texts);
    tmgpu_result* r = nullptr;
    check(tmgpu_tokenize_batch(h_, blob.data(), off.data(), std::uint32_t(texts.size()), 4, &r));
    std::vector<TokenizeResult> out(texts.size());
    const auto* ids = reinterpret_cast<const std::uint32_t*>(r->tokens);
    for (std::size_t s = 0; s < texts.size(); s++) {
      if (r->status[s] != TMGPU_S_OK) {
        tmgpu_result_free(r);
        throw Error(TMGPU_S_NORMALIZE, "normalization error");  // cpp:202
      }
      out[s].tokens.assign(ids + r->tok_off[s], ids + r->tok_off[s + 1]);
      out[s].missing = int(r->missing[s]);
    }
    tmgpu_result_free(r);
    return out;
  }
  std::vector<SerializedResult> tokenize_serialized_batch(const std::vector<std::span<const std::uint8_t>>& texts,
                                                          std::uint8_t encoding_length = 0) const {
    auto [blob, off] = flatten(texts);
    tmgpu_result* r = nullptr;
    check(tmgpu_tokenize_batch(h_, blob.data(), off.data(), std::uint32_t(texts.size()), encoding_length, &r));
    std::vector<SerializedResult> out(texts.size());
    const std::uint64_t w = r->enc_len;
    for (std::size_t s = 0; s < texts.size(); s++) {
      out[s].bytes.assign(r->tokens + r->tok_off[s] * w, r->tokens + r->tok_off[s + 1] * w);
      out[s].encoding_length = r->enc_len;
      out[s].missing = int(r->missing[s]);
    }
    tmgpu_result_free(r);
    return out;
  }
  std::vector<CountResult> count_batch(const std::vector<std::span<const std::uint8_t>>& texts) const {
    auto [blob, off] = flatten(texts);
    std::vector<std::uint64_t> c(texts.size() + 1);
    std::vector<std::uint32_t> m(texts.size() + 1);
    check(tmgpu_count_batch(h_, blob.data(), off.data(), std::uint32_t(texts.size()), c.data(), m.data(), nullptr));
    std::vector<CountResult> out(texts.size());
    for (std::size_t s = 0; s < texts.size(); s++) out[s] = {int(c[s]), int(m[s])};
    return out;
  }
  std::vector<std::vector<std::uint8_t>> decode_batch(const std::vector<std::span<const std::uint32_t>>& ids) const {
    std::vector<std::uint32_t> flat;
    std::vector<std::uint64_t> off(ids.size() + 1, 0);
    for (std::size_t s = 0; s < ids.size(); s++) {
      flat.insert(flat.end(), ids[s].begin(), ids[s].end());
      off[s + 1] = flat.size();
    }
    tmgpu_bytes* b = nullptr;
    check(tmgpu_decode_batch(h_, reinterpret_cast<const std::uint8_t*>(flat.data()), off.data(),
                             std::uint32_t(ids.size()), 4, &b));
    std::vector<std::vector<std::uint8_t>> out(ids.size());
    for (std::size_t s = 0; s < ids.size(); s++) out[s].assign(b->bytes + b->off[s], b->bytes + b->off[s + 1]);
    tmgpu_bytes_free(b);
    return out;
  }

  // ---- the reference's single-string methods (hpp:84-93) --------------------------------------
  TokenizeResult tokenize(std::span<const std::uint8_t> data) const { return tokenize_batch({data})[0]; }
  CountResult count(std::span<const std::uint8_t> data) const { return count_batch({data})[0]; }
  SerializedResult tokenize_serialized(std::span<const std::uint8_t> data, std::uint8_t encoding_length = 0) const {
    return tokenize_serialized_batch({data}, encoding_length)[0];
  }
  std::vector<std::uint8_t> decode(std::span<const std::uint32_t> tokens) const { return decode_batch({tokens})[0]; }
  Decoder new_decoder(std::uint32_t n_streams = 1) const {
    tmgpu_decoder* d = nullptr;
    check(tmgpu_decoder_create(h_, n_streams, &d));
    return Decoder(d, n_streams);
  }

 private:
  explicit Vocab(tmgpu_vocab* h) : h_(h) { check(tmgpu_vocab_info_get(h_, &info_)); }
  static std::pair<std::vector<std::uint8_t>, std::vector<std::uint64_t>> flatten(
      const std::vector<std::span<const std::uint8_t>>& texts) {
    std::vector<std::uint8_t> blob;
    std::vector<std::uint64_t> off(texts.size() + 1, 0);
    for (std::size_t s = 0; s < texts.size(); s++) {
      blob.insert(blob.end(), texts[s].begin(), texts[s].end());
      off[s + 1] = blob.size();
    }
    blob.push_back(0);
    return {std::move(blob), std::move(off)};
  }
  tmgpu_vocab* h_ = nullptr;
  tmgpu_vocab_info info_{};
};

}  // namespace tokenmonster_gpu

// host_sim.cpp -- TEST INFRASTRUCTURE.  Compiles the product's own table builder
// (csrc/vocab_tables.cpp) and per-string pipeline logic (csrc/pipeline_core.h: the functions the
// sm_100a kernels call per lane) with g++, and runs the three stages for ONE string on the host so
// that tests/test_host_sim.py can compare tables + logic with the oracle on a machine without a
// GPU.  Never linked into libtmgpu.so; the product has no CPU path.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <vector_functions.h>

#include "../tokenmonster_b200/csrc/pipeline_core.h"
#include "../tokenmonster_b200/csrc/vocab_tables.h"

using namespace tmgpu;

namespace {
struct Sim {
  HostTables t;
  uint8_t cls[128];
};

// stage 1 as the encode kernel does it (fast path of 4, else one character)
std::vector<uint8_t> encode(const Sim& S, const uint8_t* raw, int n, bool to_lower, bool* bad) {
  std::vector<uint8_t> out;
  *bad = false;
  for (int k = 0; k < n; k++)
    if (raw[k] & 0x80) { *bad = true; return out; }
  auto at = [&](int j) -> uint32_t {
    uint32_t c = raw[j];
    if (to_lower && c - 'A' < 26u) c |= 0x20u;
    return c;
  };
  CapState e;
  cap_reset(e);
  int k = 0;
  while (k < n) {
    if (k + 4 <= n) {
      const uint32_t w4 = at(k) | (at(k + 1) << 8) | (at(k + 2) << 16) | (at(k + 3) << 24);
      const uint32_t c5 = (k + 4 < n) ? at(k + 4) : 0u;
      if (cap_fast4(e, w4, c5)) {
        for (int j = 0; j < 4; j++) out.push_back(uint8_t(at(k + j)));
        k += 4;
        continue;
      }
    }
    uint32_t bytes = 0;
    const uint32_t nb = cap_step(e, S.cls, at(k), (k + 1 < n) ? at(k + 1) : uint32_t('.'), k, n, at, &bytes);
    for (uint32_t j = 0; j < nb; j++) out.push_back(uint8_t(bytes >> (8 * j)));
    k++;
  }
  return out;
}

// stage 2 as the table kernel does it
std::vector<uint32_t> table(const Sim& S, const uint8_t* text, int L, bool to_lower) {
  const HostTables& t = S.t;
  std::vector<uint32_t> words(size_t(L) + 1, 0);
  const uint4* bk = reinterpret_cast<const uint4*>(t.dfa_buckets.data());
  const uint32_t off = t.charset == 2 ? 2u : 1u;
  ClassTrail tr;
  trail_reset(tr, t.class_code[0]);
  uint32_t packed = 0, prev = 0;
  for (int p = L - 1; p >= 0; p--) {
    uint32_t b = text[p];
    if (to_lower && b - 'A' < 26u) b |= 0x20u;
    if (p == L - 1) packed = t.dfa_d1[b];
    else
      packed = dfa_next(packed, prev, b, 32u - t.dfa_bits, (1u << t.dfa_bits) - 1u,
                        [&](uint32_t i, uint4* k, uint4* v) { *k = bk[2 * i]; *v = bk[2 * i + 1]; },
                        [&](uint32_t i) { return t.dfa_d2[i]; });
    const uint32_t code = t.class_code[b];
    words[p] = position_word(packed, code, tr, off);
    trail_push(tr, code);
    prev = b;
  }
  return words;
}
}  // namespace

extern "C" {

void* hs_load(const uint8_t* vocab, size_t n, char* err, int errcap) {
  auto* S = new Sim();
  std::string e;
  int rc = parse_vocab(vocab, n, &S->t, &e);
  if (rc) {
    if (err && errcap > 0) {
      std::strncpy(err, e.c_str(), size_t(errcap) - 1);
      err[errcap - 1] = 0;
    }
    delete S;
    return nullptr;
  }
  for (uint32_t c = 0; c < 128; c++) S->cls[c] = uint8_t(pc_class(c));
  return S;
}

void hs_free(void* h) { delete static_cast<Sim*>(h); }

void hs_info(void* h, uint32_t out[4]) {
  const HostTables& t = static_cast<Sim*>(h)->t;
  out[0] = t.dfa_states;
  out[1] = t.dfa_bits;
  out[2] = uint32_t(t.dfa_buckets.size() * 4);
  out[3] = t.n_info;
}

// best match at window[0]: returns length, *index; *sp_len = key length of the space-prefixed match
int hs_dfa_lookup(void* h, const uint8_t* window, int n, uint32_t* index, uint32_t* sp_index, int* sp_len) {
  return host_dfa_lookup(static_cast<Sim*>(h)->t, window, n, index, sp_index, sp_len);
}

// normalized text of an ASCII string (capcode 2 and / or lower-casing); -1 = refused (byte >= 0x80)
long hs_normalize(void* h, const uint8_t* data, size_t n, uint8_t* out, size_t cap) {
  const Sim& S = *static_cast<Sim*>(h);
  const bool to_lower = (S.t.normflag & 2) != 0;
  std::vector<uint8_t> text;
  if (S.t.capcode == 2) {
    bool bad = false;
    text = encode(S, data, int(n), to_lower, &bad);
    if (bad) return -1;
  } else {
    text.assign(data, data + n);
    for (auto& c : text) {
      if ((S.t.normflag != 0) && (c & 0x80)) return -1;
      if (to_lower && uint32_t(c) - 'A' < 26u) c |= 0x20;
    }
  }
  if (text.size() > cap) return -2;
  if (!text.empty()) std::memcpy(out, text.data(), text.size());
  return long(text.size());
}

// ids of one string.  prenorm: the text is already normalized.  count_mode: Vocab.Count (quirk Q2).
// Returns the id count (ids beyond cap are dropped), -1 = refused (status 13).
long hs_tokenize(void* h, const uint8_t* data, size_t n, int prenorm, int count_mode, uint32_t* ids, size_t cap,
                 uint32_t* missing) {
  const Sim& S = *static_cast<Sim*>(h);
  const HostTables& t = S.t;
  *missing = 0;
  if (t.max_len == 0) return 0;
  const bool to_lower = !prenorm && (t.normflag & 2) != 0;
  std::vector<uint8_t> text;
  const uint8_t* tp = data;
  int L = int(n);
  if (!prenorm && t.capcode == 2) {
    bool bad = false;
    text = encode(S, data, int(n), to_lower, &bad);
    if (bad) return -1;
    tp = text.data();
    L = int(text.size());
  } else if (!prenorm && t.normflag != 0) {
    for (size_t k = 0; k < n; k++)
      if (data[k] & 0x80) return -1;
  }
  const std::vector<uint32_t> words = table(S, tp, L, to_lower && t.capcode != 2);
  ChainTables T;
  T.rec = reinterpret_cast<const uint4*>(t.rec.data());
  T.hot = reinterpret_cast<const uint4*>(t.dfa_hot.data());
  T.delete_token = t.delete_token;
  T.unk = t.unk;
  T.off = t.charset == 2 ? 2u : 1u;
  T.can_retry = t.delete_token != 0xFFFFFFu;
  ChainString cs{words.data(), L};
  ChainLane z;
  chain_reset(z);
  size_t cnt = 0;
  auto emit = [&](uint32_t id, uint32_t, uint32_t) {
    if (cnt < cap && ids) ids[cnt] = id;
    cnt++;
  };
  auto on_miss = [&](int) -> uint32_t {
    if (t.unk != 0xFFFFFFu) {
      emit(t.unk, 0, 0);
      return 1u;
    }
    return 0u;
  };
  auto load_rec = [&](uint32_t idx, uint4* r0, uint4* r1) { *r0 = T.rec[2 * idx]; *r1 = T.rec[2 * idx + 1]; };
  while (z.i < L) chain_step(T, cs, z, count_mode == 0, load_rec, emit, on_miss);
  *missing = z.miss;
  return long(count_mode ? z.cnt : cnt);
}

}  // extern "C"

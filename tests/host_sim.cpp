// host_sim.cpp -- TEST INFRASTRUCTURE.  Compiles the product's own table builder
// (csrc/vocab_tables.cpp) and per-string pipeline logic (csrc/pipeline_core.h: the functions the
// sm_100a kernels call per lane) with g++, and runs the three stages for ONE string on the host so
// that tests/test_host_sim.py can compare tables + logic with the oracle on a machine without a
// GPU.  Never linked into libtmgpu.so; the product has no CPU path.
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <vector_functions.h>

#include "../tokenmonster_b200/csrc/pipeline_core.h"
#include "../tokenmonster_b200/csrc/vocab_tables.h"
#include "../tokenmonster_b200/csrc/capdec_tables.inc"   // the decode kernel's transducer tables (host copy)

using namespace tmgpu;

namespace {
struct Sim {
  HostTables t;
  uint8_t cls[128];
  uint32_t lut[kCapLutSize];
};

// stages 1 + 2 as the table kernel does them: from the end of the string, the encoder yields the
// bytes of one raw character at a time (forward order), a byte queue hands them to the automaton
// last byte first, one byte per step.  Returns the position words (index = normalized position)
// and the normalized text.
struct Tabled {
  std::vector<uint32_t> words;
  std::vector<uint8_t> text;
  bool bad = false;
};

Tabled table(const Sim& S, const uint8_t* raw, int n, bool cap, bool to_lower, bool want_ascii) {
  const HostTables& t = S.t;
  Tabled R;
  if (want_ascii)
    for (int k = 0; k < n; k++)
      if (raw[k] & 0x80) { R.bad = true; return R; }
  auto at = [&](int j) -> uint32_t {
    uint32_t c = raw[j];
    if (to_lower && c - 'A' < 26u) c |= 0x20u;
    return c;
  };
  const uint4* bk = reinterpret_cast<const uint4*>(t.dfa_buckets.data());
  const uint32_t off = t.charset == 2 ? 2u : 1u;
  ClassTrail tr;
  trail_reset(tr, t.class_code[0]);
  RevCap e;
  revcap_reset(e);
  uint32_t packed = 0, qbytes = 0, qn = 0;
  int k = n - 1;
  auto xc = [&](int j) -> uint32_t { return j >= 0 ? revcap_char(at(j)) : 0u; };   // the kernel's s_chr[] entry
  if (cap) {   // the read-ahead: characters k, k-1, k-2
    revcap_push(e, xc(k));
    revcap_push(e, xc(k - 1));
    revcap_push(e, xc(k - 2));
  }
  while (k >= 0 || qn) {
    if (qn == 0) {
      if (cap) {
        qn = revcap_step(e, [&](uint32_t i) -> uint32_t { return S.lut[i]; }, k,
                         [&](int j) -> uint32_t { return revcap_char(at(j)) >> 8; }, &qbytes);
        revcap_push(e, xc(k - 3));
      } else {
        qbytes = at(k);
        qn = 1;
      }
      k--;
      if (qn == 0) continue;   // a space replaced by a marker emits nothing
    }
    const uint32_t b = qbytes & 0xFFu;   // the queue leaves lowest byte first
    qbytes >>= 8;
    qn--;
    packed = dfa_next(packed, b, 32u - t.dfa_bits, (1u << t.dfa_bits) - 1u,
                      [&](uint32_t i, uint4* kk, uint4* v) { *kk = bk[2 * i]; *v = bk[2 * i + 1]; },
                      [&](uint32_t c) { return t.dfa_d1[c]; });
    const uint32_t code = t.class_code[b];
    R.words.push_back(position_word(packed, code, tr, off));
    R.text.push_back(uint8_t(b));
    trail_push(tr, code);
  }
  std::reverse(R.words.begin(), R.words.end());
  std::reverse(R.text.begin(), R.text.end());
  return R;
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
  for (uint32_t i = 0; i < uint32_t(kCapLutSize); i++) S->lut[i] = revcap_lut_entry(i);
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
  const HostTables& t = S.t;
  const Tabled R = table(S, data, int(n), t.capcode == 2, (t.normflag & 2) != 0, t.normflag != 0 || t.capcode == 2);
  if (R.bad) return -1;
  if (R.text.size() > cap) return -2;
  if (!R.text.empty()) std::memcpy(out, R.text.data(), R.text.size());
  return long(R.text.size());
}

// position words + normalized text of one string (analysis / tests); returns the normalized length
long hs_table(void* h, const uint8_t* data, size_t n, int prenorm, uint32_t* words, uint8_t* text, size_t cap) {
  const Sim& S = *static_cast<Sim*>(h);
  const HostTables& t = S.t;
  const Tabled R = table(S, data, int(n), !prenorm && t.capcode == 2, !prenorm && (t.normflag & 2) != 0,
                         !prenorm && (t.normflag != 0 || t.capcode == 2));
  if (R.bad) return -1;
  if (R.words.size() > cap) return -2;
  for (size_t k = 0; k < R.words.size(); k++) {
    words[k] = R.words[k];
    text[k] = R.text[k];
  }
  return long(R.words.size());
}

// ids of one string.  prenorm: the text is already normalized.  count_mode: Vocab.Count (quirk Q2).
// Returns the id count (ids beyond cap are dropped), -1 = refused (status 13).
long hs_tokenize(void* h, const uint8_t* data, size_t n, int prenorm, int count_mode, uint32_t* ids, size_t cap,
                 uint32_t* missing) {
  const Sim& S = *static_cast<Sim*>(h);
  const HostTables& t = S.t;
  *missing = 0;
  if (t.max_len == 0) return 0;
  const Tabled R = table(S, data, int(n), !prenorm && t.capcode == 2, !prenorm && (t.normflag & 2) != 0,
                         !prenorm && (t.normflag != 0 || t.capcode == 2));
  if (R.bad) return -1;
  const int L = int(R.words.size());
  ChainTables T;
  T.crec = reinterpret_cast<const uint4*>(t.chain_rec.data());
  T.hot = reinterpret_cast<const uint4*>(t.dfa_hot.data());
  T.delete_token = t.delete_token;
  T.unk = t.unk;
  T.off = t.charset == 2 ? 2u : 1u;
  T.can_retry = t.delete_token != 0xFFFFFFu;
  ChainString cs{R.words.data(), L};
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
  auto load32 = [&](const uint4* tab, uint32_t idx, uint4* a, uint4* b) { *a = tab[2 * idx]; *b = tab[2 * idx + 1]; };
  while (z.i < L) chain_step(T, cs, z, count_mode == 0, load32, emit, on_miss);
  *missing = z.miss;
  return long(count_mode ? z.cnt : cnt);
}

// The warp algorithm of capcode_decode_scan_kernel (csrc/capcode_decode_scan.cuh) for one string,
// lane by lane, on the kernel's own tables: 512-byte rounds from the 16-byte aligned address below
// the string (galign = the string's offset from it), 16 bytes per lane, bytes outside the string
// read as 0xFF; a monoid id per lane, the scan, the second pass from the scanned start state.
// Returns the output length, -1 for text with a byte >= 0x80; *st_io = decoder flags in / out.
long hs_capdec_scan(const uint8_t* in, size_t len, int galign, uint8_t* st_io, uint8_t* out) {
  const long long total_c = (long long)galign + (long long)len;
  uint32_t st = *st_io & 15u;
  size_t o = 0;
  for (long long base_c = 0; base_c < total_c; base_c += 512) {
    uint8_t bytes[32][16];
    uint32_t m[32], incl[32];
    for (int lane = 0; lane < 32; lane++) {
      m[lane] = 0;
      for (int j = 0; j < 16; j++) {
        const long long c = base_c + lane * 16 + j;
        uint8_t b = 0xFF;
        if (c >= galign && c < total_c) {
          b = in[c - galign];
          if (b & 0x80) return -1;
        }
        bytes[lane][j] = b;
        m[lane] = kCapdecComp[m[lane] * 256u + b];
      }
    }
    for (int lane = 0; lane < 32; lane++) incl[lane] = m[lane];
    for (int d = 1; d < 32; d <<= 1) {   // the shuffle scan, all lanes at once
      uint32_t nxt[32];
      for (int lane = 0; lane < 32; lane++)
        nxt[lane] = lane >= d ? kCapdecMul[incl[lane - d] * 64u + incl[lane]] : incl[lane];
      for (int lane = 0; lane < 32; lane++) incl[lane] = nxt[lane];
    }
    for (int lane = 0; lane < 32; lane++) {
      const uint32_t excl = lane ? incl[lane - 1] : 0u;
      uint32_t cur = kCapdecApply[excl * 16u + st];
      for (int j = 0; j < 16; j++) {
        const uint32_t b = bytes[lane][j];
        const uint32_t e = kCapdecStep[cur * 256u + b];
        cur = e & 15u;
        if (e & 16u) out[o++] = uint8_t(b - (e & 32u));
      }
    }
    st = kCapdecApply[incl[31] * 16u + st];
  }
  *st_io = uint8_t(st);
  return long(o);
}

}  // extern "C"

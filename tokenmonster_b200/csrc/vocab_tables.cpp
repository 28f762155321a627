// vocab_tables.cpp -- .vocab parser and device-table builder (host side, no CUDA).
// Format: /root/reference go/tokenmonster.go:2602-2653 (writer), :2656-2736 (reader);
// worked example tokenmonster-cpp/tests/unit.cpp:12-79.
#include "vocab_tables.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string_view>
#include <unordered_map>

#include "../../include/tmgpu.h"
#include "device_tables.h"

namespace tmgpu {

namespace {
struct Cursor {
  const uint8_t* p;
  size_t n, at = 0;
  bool ok = true;
  bool need(size_t k) {
    if (at + k > n) ok = false;
    return ok;
  }
  uint32_t u8() { return need(1) ? p[at++] : 0; }
  uint32_t u24() {
    if (!need(3)) return 0;
    uint32_t v = p[at] | (p[at + 1] << 8) | (uint32_t(p[at + 2]) << 16);
    at += 3;
    return v;
  }
};
}  // namespace

struct KeyRef {
  uint32_t off;
  uint8_t len;
};

// Reversed-key Aho-Corasick automaton in DFA form (vocab_tables.h describes the tables).
// score parts (pipeline_core.h, stage 3)
static int score_a(uint32_t aflag, int nwords, int fd) {
  const int nw = nwords - fd;
  return int(aflag >> 7) + (nw > 1 ? nw - 1 : 0) + 100 * nw;
}
static int score_b(uint32_t bflag, int bnw, int nb, bool retry) {
  int s = int(bflag >> 7) + (bnw > 1 ? bnw - 1 : 0) + ((nb >> 2) & 1) + 100 * (bnw + (nb >> 3)) - 3 * int(bflag & 1u & uint32_t(nb));
  return retry ? s - 1 : s + int((bflag >> 2) & 1u);
}
static uint32_t pack_i16(int lo, int hi) { return (uint32_t(lo) & 0xFFFFu) | (uint32_t(hi) << 16); }

static int build_dfa(HostTables* t, const std::vector<KeyRef>& keys, const std::vector<uint8_t>& blob,
                     const std::vector<uint8_t>& flag, const std::vector<uint8_t>& nwords,
                     const std::vector<uint32_t>& alt, const std::vector<uint32_t>& alt2, const std::vector<uint32_t>& id,
                     std::string* err) {
  for (int b = 0; b < 256; b++) {
    const uint8_t v = t->begin_byte[b];   // go:3780-3787 writes only these four values
    if (v == 0) t->class_code[b] = 0;
    else if (v == 1) t->class_code[b] = 1;
    else if (v == 10) t->class_code[b] = 2;
    else if (v == 12) t->class_code[b] = 3;
    else {
      *err = "beginByte table holds a value other than 0, 1, 10, 12";
      return TMGPU_E_UNSUPPORTED;
    }
  }
  // ---- trie of the reversed keys ---------------------------------------------------------
  struct Node {
    uint32_t fail = 0, term = kNone, child_begin = 0, child_end = 0;
    uint8_t depth = 0;
  };
  std::vector<Node> nodes(1);
  std::unordered_map<uint32_t, uint32_t> child_of;   // node<<8 | byte -> child
  child_of.reserve(size_t(t->n_info) * 4);
  struct Edge {
    uint32_t parent, child;
    uint8_t byte;
  };
  std::vector<Edge> edges;
  for (uint32_t i = 0; i < t->n_info; i++) {
    uint32_t at = 0;
    for (int k = int(keys[i].len) - 1; k >= 0; k--) {
      const uint8_t b = blob[keys[i].off + k];
      if (nodes.size() >= (1u << kDfaStateBits)) {
        *err = "vocabulary too large for the 20-bit automaton state space";
        return TMGPU_E_UNSUPPORTED;
      }
      auto ins = child_of.emplace((at << 8) | b, uint32_t(nodes.size()));
      if (ins.second) {
        Node nd;
        nd.depth = uint8_t(nodes[at].depth + 1);
        nodes.push_back(nd);
        edges.push_back({at, ins.first->second, b});
      }
      at = ins.first->second;
    }
    nodes[at].term = i;
  }
  const uint32_t N = uint32_t(nodes.size());
  t->dfa_states = N;
  std::sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) {
    return a.parent != b.parent ? a.parent < b.parent : a.byte < b.byte;
  });
  for (size_t e = 0; e < edges.size(); e++) {
    Node& p = nodes[edges[e].parent];
    if (p.child_end == 0) p.child_begin = uint32_t(e);
    p.child_end = uint32_t(e) + 1;
  }
  // ---- failure links, rows of the DFA restricted to targets at depth >= 2 ------------------
  // row entry: byte << 24 | target, sorted by byte; row(s) = row(fail(s)) overridden by s's children
  std::vector<std::vector<uint32_t>> row(N);
  uint32_t d1[256];
  for (int b = 0; b < 256; b++) {
    auto it = child_of.find(uint32_t(b));
    d1[b] = it == child_of.end() ? 0u : it->second;
  }
  auto delta = [&](uint32_t s, uint32_t b) -> uint32_t {
    const std::vector<uint32_t>& r = row[s];
    size_t lo = 0, hi = r.size();
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if ((r[mid] >> 24) < b) lo = mid + 1; else hi = mid;
    }
    if (lo < r.size() && (r[lo] >> 24) == b) return r[lo] & 0xFFFFFFu;
    return d1[b];
  };
  std::vector<uint32_t> order;
  order.reserve(N);
  order.push_back(0);
  std::vector<uint32_t> best_idx(N, kNone);
  std::vector<uint8_t> best_len(N, 0);
  for (size_t q = 0; q < order.size(); q++) {
    const uint32_t s = order[q];
    Node& nd = nodes[s];
    if (s != 0) {
      const std::vector<uint32_t>& base = row[nd.fail];
      std::vector<uint32_t>& r = row[s];
      r.reserve(base.size() + (nd.child_end - nd.child_begin));
      size_t a = 0;
      uint32_t e = nd.child_begin;
      while (a < base.size() || e < nd.child_end) {
        const uint32_t ba = a < base.size() ? (base[a] >> 24) : 256u;
        const uint32_t be = e < nd.child_end ? edges[e].byte : 256u;
        if (be <= ba) {
          r.push_back((be << 24) | edges[e].child);
          e++;
          if (be == ba) a++;
        } else {
          r.push_back(base[a++]);
        }
      }
      if (nd.term != kNone) { best_idx[s] = nd.term; best_len[s] = nd.depth; }
      else { best_idx[s] = best_idx[nd.fail]; best_len[s] = best_len[nd.fail]; }
    }
    for (uint32_t e = nd.child_begin; e < nd.child_end; e++) {
      const uint32_t c = edges[e].child;
      nodes[c].fail = (s == 0) ? 0u : delta(nd.fail, edges[e].byte);
      order.push_back(c);
    }
  }
  // ---- the space-prefixed retry: best key ending at the virtual space ----------------------
  std::vector<uint32_t> sp_idx(N, kNone);
  std::vector<uint8_t> sp_len(N, 0);
  for (uint32_t s = 0; s < N; s++) {
    uint32_t x = s;
    if (t->charset == 2) x = delta(x, 0x00);   // the virtual space is 0x20 0x00, consumed back to front
    x = delta(x, 0x20);
    sp_idx[s] = best_idx[x];
    sp_len[s] = best_len[x];
  }
  auto pack = [&](uint32_t s) { return s | (uint32_t(best_len[s]) << 20) | (uint32_t(sp_len[s]) << 26); };
  for (int b = 0; b < 256; b++) t->dfa_d1[b] = pack(d1[b]);
  static const int kClassValue[4] = {0, 1, 10, 12};
  t->dfa_hot.assign(size_t(N) * 8, 0);
  for (uint32_t s = 0; s < N; s++) {
    uint32_t* h = &t->dfa_hot[size_t(s) * 8];
    const uint32_t bi = best_len[s] ? best_idx[s] : 0u, si = sp_len[s] ? sp_idx[s] : 0u;
    const uint32_t bf = best_len[s] ? flag[bi] : 0u, sf = sp_len[s] ? flag[si] : 0u;
    const int bn = best_len[s] ? nwords[bi] : 0, sn = sp_len[s] ? nwords[si] : 0;
    h[0] = bi | (uint32_t(best_len[s]) << 24);
    h[1] = si | (uint32_t(sp_len[s]) << 24);
    h[2] = ((bf >> 1) & 1u) | (((bf >> 4) & 1u) << 1) | ((((bf & 2u) != 0 && bn == 0) ? 1u : 0u) << 2) | (((sf >> 4) & 1u) << 3);
    int bp[4], br[4];
    for (int c = 0; c < 4; c++) {
      bp[c] = score_b(bf, bn, kClassValue[c], false);
      br[c] = score_b(sf, sn, kClassValue[c], true);
    }
    h[4] = pack_i16(bp[0], bp[1]);
    h[5] = pack_i16(bp[2], bp[3]);
    h[6] = pack_i16(br[0], br[1]);
    h[7] = pack_i16(br[2], br[3]);
  }
  // chain records: ids, alternative lengths, first-token score parts (forward delete 0 / 1)
  t->chain_rec.assign(size_t(t->n_info) * 8, 0);
  for (uint32_t i = 0; i < t->n_info; i++) {
    uint32_t* r = &t->chain_rec[size_t(i) * 8];
    const bool h1 = alt[i] != kDNE, h2 = alt2[i] != kDNE;
    const uint32_t f0 = flag[i], f1 = h1 ? flag[alt[i]] : 0u, f2 = h2 ? flag[alt2[i]] : 0u;
    const int n0 = nwords[i], n1 = h1 ? nwords[alt[i]] : 0, n2 = h2 ? nwords[alt2[i]] : 0;
    const uint32_t bits = (f0 & 1u) | (((f0 >> 3) & 1u) << 1) | (((f0 >> 5) & 1u) << 2) | ((f1 & 1u) << 3) |
                          (((f1 >> 3) & 1u) << 4) | ((f2 & 1u) << 5) | (((f2 >> 3) & 1u) << 6);
    r[0] = id[i] | (bits << 24);
    r[1] = (h1 ? id[alt[i]] : 0u) | (uint32_t(h1 ? keys[alt[i]].len : 0) << 24);
    r[2] = (h2 ? id[alt2[i]] : 0u) | (uint32_t(h2 ? keys[alt2[i]].len : 0) << 24);
    r[3] = pack_i16(score_a(f0, n0, 0), score_a(f0, n0, 1));
    r[4] = pack_i16(score_a(f1, n1, 0), score_a(f1, n1, 1));
    r[5] = pack_i16(score_a(f2, n2, 0), score_a(f2, n2, 1));
  }
  // ---- hashed transitions to depth >= 2 (every row entry) -----------------------------------
  size_t n_deep = 0;
  for (uint32_t s = 1; s < N; s++) n_deep += row[s].size();
  uint32_t bits = 4;
  // on average <= 2 of 4 slots used.  (Twice the buckets -- <= 1 of 4 -- halves the second probes, yet
  // measured no faster on c2, 4.75 against 4.71 ms, and 9.3 against 6.3 ms on c5, whose table then
  // misses L2.)
  while ((size_t(1) << bits) * 2 < n_deep) bits++;
  t->dfa_bits = bits;
  const uint32_t mask = (1u << bits) - 1u, shift = 32u - bits;
  t->dfa_buckets.assign((size_t(1) << bits) * 8, 0);
  for (size_t b = 0; b < (size_t(1) << bits); b++)
    for (int k = 0; k < 4; k++) t->dfa_buckets[b * 8 + k] = kDfaEmpty;
  // Insert the transitions to SHALLOW targets first: a full bucket pushes later keys into the next
  // one (a second dependent probe for the whole warp), and short suffixes are what text visits most,
  // so the keys that end up displaced should be the deep, rare ones.
  std::vector<std::pair<uint32_t, uint32_t>> trans;   // key, packed target -- bucketed by target depth
  trans.reserve(n_deep);
  for (uint32_t depth = 2; depth <= 41; depth++)
    for (uint32_t s = 1; s < N; s++)
      for (uint32_t e : row[s]) {
        const uint32_t tgt = e & 0xFFFFFFu;
        if (nodes[tgt].depth == depth) trans.emplace_back((s << 8) | (e >> 24), pack(tgt));
      }
  for (const auto& kv : trans) {
    const uint32_t key = kv.first;
    uint32_t h = dfa_hash(key, shift);
    for (;;) {
      uint32_t* bk = &t->dfa_buckets[size_t(h) * 8];
      int slot = -1;
      for (int k = 0; k < 4; k++)
        if ((bk[k] & ~kDfaOverflow) == kDfaEmpty) { slot = k; break; }
      if (slot >= 0) {
        bk[slot] = (bk[slot] & kDfaOverflow) | key;
        bk[4 + slot] = kv.second;
        break;
      }
      bk[0] |= kDfaOverflow;
      h = (h + 1) & mask;
    }
  }
  return TMGPU_OK;
}

int host_dfa_lookup(const HostTables& h, const uint8_t* window, int n, uint32_t* index, uint32_t* sp_index,
                    int* sp_len) {
  const uint4* bk = reinterpret_cast<const uint4*>(h.dfa_buckets.data());
  uint32_t w = 0;
  for (int p = n - 1; p >= 0; p--) {
    w = dfa_next(w, window[p], 32u - h.dfa_bits, (1u << h.dfa_bits) - 1u,
                 [&](uint32_t i, uint4* k, uint4* v) { *k = bk[2 * i]; *v = bk[2 * i + 1]; },
                 [&](uint32_t b) { return h.dfa_d1[b]; });
  }
  const uint32_t* hot = &h.dfa_hot[size_t(w & kDfaStateMask) * 8];
  if (index) *index = hot[0] & 0xFFFFFFu;
  if (sp_index) *sp_index = hot[1] & 0xFFFFFFu;
  if (sp_len) *sp_len = int(dfa_splen(w));
  return n > 0 ? int(dfa_len(w)) : 0;
}

int parse_vocab(const uint8_t* file, size_t n, HostTables* t, std::string* err) {
  Cursor c{file, n};
  if (n < 24) {
    *err = "not a valid TokenMonster vocabulary (truncated header)";
    return TMGPU_E_VOCAB_FORMAT;
  }
  t->capcode = (uint8_t)c.u8();
  t->charset = (uint8_t)c.u8();
  t->normflag = (uint8_t)c.u8();
  t->level = (uint8_t)c.u8();
  t->reserve = (uint8_t)c.u8();
  c.at = 8;
  if (t->charset > 2 || t->capcode > 2) {  // go:2675
    *err = "not a valid TokenMonster vocabulary (charset/capcode byte)";
    return TMGPU_E_VOCAB_FORMAT;
  }
  t->unk = c.u24();
  t->vocab_size = c.u24();
  t->n_reverse = c.u24();
  t->n_info = c.u24();
  t->delete_token = c.u24();
  t->max_len = (int)c.u8();

  using Key = KeyRef;
  std::vector<Key> keys(t->n_info);
  std::vector<uint8_t> blob;
  blob.reserve(size_t(t->n_info) * 8);
  std::vector<uint8_t> flag(t->n_info), nwords(t->n_info);
  std::vector<uint32_t> alt(t->n_info), alt2(t->n_info), id(t->n_info);
  t->rev_off.assign(t->n_reverse, 0);
  t->rev_len.assign(t->n_reverse, 0xFF);

  for (uint32_t i = 0; i < t->n_info; i++) {
    uint32_t kl = c.u8();
    if (!c.ok || kl == 0 || kl > 40 || !c.need(kl + 15)) {  // go:2695
      *err = "not a valid TokenMonster vocabulary (token record)";
      return TMGPU_E_VOCAB_FORMAT;
    }
    keys[i] = {uint32_t(blob.size()), uint8_t(kl)};
    blob.insert(blob.end(), file + c.at, file + c.at + kl);
    c.at += kl;
    flag[i] = (uint8_t)c.u8();
    nwords[i] = (uint8_t)c.u8();
    alt[i] = c.u24();
    alt2[i] = c.u24();
    id[i] = c.u24();
    c.at += 4;  // float32 score: not used on the hot path
    // alternatives always point at EARLIER records (go:2636 enforces it at save time,
    // go:2703-2711 resolves them against what has been read so far)
    if ((alt[i] != kDNE && alt[i] >= i) || (alt2[i] != kDNE && alt2[i] >= i) || id[i] >= t->n_reverse) {
      *err = "not a valid TokenMonster vocabulary (alternative index / id out of range)";
      return TMGPU_E_VOCAB_FORMAT;
    }
    if (i > 0) {  // the dictionary index IS the record order: cpp:1352-1357
      const Key& a = keys[i - 1];
      const Key& b = keys[i];
      bool bad = a.len > b.len || (a.len == b.len && std::memcmp(&blob[a.off], &blob[b.off], b.len) >= 0);
      if (bad) {
        *err = "dictionary order does not match vocabulary info order";
        return TMGPU_E_VOCAB_ORDER;
      }
    }
  }
  if (!c.need(256 + 3)) {
    *err = "not a valid TokenMonster vocabulary (beginByte table)";
    return TMGPU_E_VOCAB_FORMAT;
  }
  std::memcpy(t->begin_byte, file + c.at, 256);
  c.at += 256;
  uint32_t n_deleted = c.u24();
  for (uint32_t i = 0; i < n_deleted && c.ok; i++) {
    uint32_t kl = c.u8();
    if (c.need(kl + 7)) c.at += kl + 7;
  }
  if (!c.ok || c.at != n) {  // go:2731
    *err = "not a valid TokenMonster vocabulary (trailing bytes)";
    return TMGPU_E_VOCAB_FORMAT;
  }

  // ---- token records -------------------------------------------------------------------
  t->rec.resize(t->n_info);
  for (uint32_t i = 0; i < t->n_info; i++) {
    TokenRecord& r = t->rec[i];
    uint32_t al = alt[i] != kDNE ? keys[alt[i]].len : 0;
    uint32_t al2 = alt2[i] != kDNE ? keys[alt2[i]].len : 0;
    uint32_t f1 = alt[i] != kDNE ? flag[alt[i]] : 0, n1 = alt[i] != kDNE ? nwords[alt[i]] : 0;
    uint32_t f2 = alt2[i] != kDNE ? flag[alt2[i]] : 0, n2 = alt2[i] != kDNE ? nwords[alt2[i]] : 0;
    r.w[0] = flag[i] | (uint32_t(nwords[i]) << 8) | (uint32_t(keys[i].len) << 16) | (al << 24);
    r.w[1] = id[i];
    r.w[2] = alt[i] != kDNE ? id[alt[i]] : 0;
    r.w[3] = alt2[i] != kDNE ? id[alt2[i]] : 0;
    r.w[4] = al2 | (f1 << 8) | (n1 << 16) | (f2 << 24);
    r.w[5] = n2;
    r.w[6] = alt[i];
    r.w[7] = alt2[i];
    t->rev_off[id[i]] = keys[i].off;  // later records overwrite earlier ones (go:2715)
    t->rev_len[id[i]] = keys[i].len;
  }
  t->rev_blob = blob;
  t->rev_pack.assign(t->n_reverse, 0);
  for (uint32_t id = 0; id < t->n_reverse; id++) {
    const uint32_t l = t->rev_len[id];
    if (l == 0xFF || l == 0) continue;
    const uint32_t at = uint32_t(t->rev_blob8.size());
    t->rev_blob8.resize(at + (l + 7) / 8, 0);
    std::memcpy(&t->rev_blob8[at], &blob[t->rev_off[id]], l);
    t->rev_pack[id] = at | (l << 24);
  }
  t->rev_blob8.push_back(0);

  // ---- prefix trie ---------------------------------------------------------------------
  std::unordered_map<std::string_view, uint32_t> node_of;
  node_of.reserve(size_t(t->n_info) * 4);
  for (uint32_t i = 0; i < t->n_info; i++)
    node_of.emplace(std::string_view((const char*)&blob[keys[i].off], keys[i].len), i);
  struct Edge {
    uint32_t parent, child;
    uint8_t byte;
  };
  std::vector<Edge> edge_list;
  edge_list.reserve(size_t(t->n_info) * 3);
  uint32_t next_id = t->n_info;
  constexpr uint32_t kRoot = 0xFFFFFFu;
  // every proper prefix of every key becomes an interior node unless it is itself a key
  for (uint32_t i = 0; i < t->n_info; i++) {
    const char* k = (const char*)&blob[keys[i].off];
    for (uint32_t l = 1; l < keys[i].len; l++) {
      auto ins = node_of.emplace(std::string_view(k, l), next_id);
      if (ins.second) next_id++;
    }
  }
  // one edge per node: (node of the prefix minus its last byte, last byte) -> node
  for (const auto& kv : node_of) {
    std::string_view p = kv.first;
    uint32_t parent = p.size() == 1 ? kRoot : node_of.find(p.substr(0, p.size() - 1))->second;
    edge_list.push_back({parent, kv.second, (uint8_t)p.back()});
  }
  t->n_nodes = next_id;
  if (next_id >= 0xFFFFF0u) {
    *err = "vocabulary too large for the 24-bit trie node space";
    return TMGPU_E_UNSUPPORTED;
  }
  std::vector<uint8_t> sig(next_id, 0);
  for (const Edge& e : edge_list)
    if (e.parent != kRoot) sig[e.parent] |= uint8_t(1u << sig_bit(e.byte));
  uint32_t bits = 4;
  while ((1ull << bits) < edge_list.size() * 4 + 2) bits++;  // load factor <= 1/4
  t->edge_bits = bits;
  t->edges.assign(size_t(1) << bits, ~0ull);
  for (int b = 0; b < 256; b++) t->root_child[b] = kNone;
  uint32_t mask = (1u << bits) - 1;
  for (const Edge& e : edge_list) {
    uint32_t val = e.child | (uint32_t(sig[e.child]) << 24);
    if (e.parent == kRoot) {
      t->root_child[e.byte] = val;
      continue;
    }
    uint32_t key = (e.parent << 8) | e.byte;
    uint32_t h = edge_hash(key, bits);
    while (t->edges[h] != ~0ull) h = (h + 1) & mask;
    t->edges[h] = uint64_t(key) | (uint64_t(val) << 32);
  }
  // start state of the space-prefixed retry (go:1029-1036 lilbuf): after " " (+0x00 in UTF-16)
  t->space_start = t->root_child[0x20];
  if (t->charset == 2 && t->space_start != kNone) {
    uint32_t node = t->space_start & 0xFFFFFFu;
    uint32_t key = (node << 8) | 0u, h = edge_hash(key, bits);
    t->space_start = kNone;
    while (t->edges[h] != ~0ull) {
      if (uint32_t(t->edges[h]) == key) {
        t->space_start = uint32_t(t->edges[h] >> 32);
        break;
      }
      h = (h + 1) & mask;
    }
  }
  auto child_of = [&](uint32_t node, uint32_t byte) -> uint32_t {
    uint32_t key = (node << 8) | byte, h = edge_hash(key, bits);
    while (t->edges[h] != ~0ull) {
      if (uint32_t(t->edges[h]) == key) return uint32_t(t->edges[h] >> 32);
      h = (h + 1) & mask;
    }
    return kNone;
  };
  t->root2.assign(65536, kNone);
  for (uint32_t b0 = 0; b0 < 256; b0++) {
    if (t->root_child[b0] == kNone) continue;
    const uint32_t n1 = t->root_child[b0] & 0xFFFFFFu;
    for (uint32_t b1 = 0; b1 < 256; b1++) t->root2[b0 * 256 + b1] = child_of(n1, b1);
  }
  for (uint32_t b = 0; b < 256; b++)
    t->sp_child[b] = t->space_start == kNone ? kNone : child_of(t->space_start & 0xFFFFFFu, b);
  return build_dfa(t, keys, blob, flag, nwords, alt, alt2, id, err);
}

// spaced: 0 plain window, 1 window behind the virtual space; +2 = take the first steps from the
// shortcut tables the lane-per-string kernel uses (root2 / sp_child) instead of the hashed edges
int host_trie_lookup(const HostTables& h, const uint8_t* window, int n, int spaced, uint32_t* index) {
  const bool shortcut = (spaced & 2) != 0;
  spaced &= 1;
  uint32_t start = spaced ? h.space_start : kRootStart;
  if (spaced && start == kNoVal) return 0;
  const uint2* edges = reinterpret_cast<const uint2*>(h.edges.data());
  uint32_t idx = 0;
  if (shortcut && n > 0) {   // mirrors the task fetch of tokenize_thread_kernel
    uint32_t best = 0, val = kNoVal;
    int best_len = 0, d = 0;
    const uint32_t v1 = spaced ? h.sp_child[window[0]] : h.root_child[window[0]];
    if (v1 != kNoVal) {
      d = 1;
      val = v1;
      if ((v1 & 0xFFFFFFu) < h.n_info) { best = v1 & 0xFFFFFFu; best_len = 1; }
      if (!spaced && n > 1) {
        const uint32_t v2 = h.root2[size_t(window[0]) * 256 + window[1]];
        if (v2 != kNoVal) {
          d = 2;
          val = v2;
          if ((v2 & 0xFFFFFFu) < h.n_info) { best = v2 & 0xFFFFFFu; best_len = 2; }
        } else {
          val = kNoVal;
        }
      }
    }
    if (val != kNoVal && d < n) {   // the rest of the walk through the hashed edges
      uint32_t sub = 0;
      int l = trie_longest([&](int k) { return uint32_t(window[d + k]); }, [&](uint32_t) { return kNoVal; },
                           [&](uint32_t s) { return edges[s]; }, uint32_t(h.edges.size() - 1), 32u - h.edge_bits,
                           h.n_info, n - d, val, &sub);
      if (l > 0) { best = sub; best_len = d + l; }
    }
    if (index) *index = best;
    return best_len;
  }
  int l = trie_longest([&](int d) { return uint32_t(window[d]); }, [&](uint32_t b) { return h.root_child[b]; },
                       [&](uint32_t s) { return edges[s]; }, uint32_t(h.edges.size() - 1), 32u - h.edge_bits,
                       h.n_info, n, start, &idx);
  if (index) *index = idx;
  return l;
}

}  // namespace tmgpu

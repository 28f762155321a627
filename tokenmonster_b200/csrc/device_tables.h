// device_tables.h -- table layout shared by the host table builder and the sm_100a kernels,
// plus the two pure functions both sides agree on: the trie walk and the pair score.
#pragma once
#include <cstdint>
#include <vector_types.h>

#if defined(__CUDACC__)
#define TM_HD __host__ __device__ __forceinline__
#else
#define TM_HD inline
#endif

namespace tmgpu {

constexpr uint32_t kNoVal = 0xFFFFFFFFu;   // "no child" / empty edge key
constexpr uint32_t kRootStart = 0xFFFFFFFEu;  // walk starts at the trie root
constexpr int kNegScore = -1000000;        // go/tokenmonster.go:1059-1065 sentinel

// ---- reversed-key automaton (vocab_tables.h has the table description) ------------------------
constexpr uint32_t kDfaStateBits = 20;
constexpr uint32_t kDfaStateMask = (1u << kDfaStateBits) - 1u;
constexpr uint32_t kDfaEmpty = 0x7FFFFFFFu;
constexpr uint32_t kDfaOverflow = 0x80000000u;
TM_HD uint32_t dfa_len(uint32_t packed) { return (packed >> 20) & 63u; }
TM_HD uint32_t dfa_splen(uint32_t packed) { return packed >> 26; }
TM_HD uint32_t dfa_hash(uint32_t key, uint32_t shift) { return (key * 0x9E3779B1u) >> shift; }

struct DfaDev {
  const uint4* buckets;      // 2 x uint4 per bucket: keys, values
  const uint4* hot;          // 2 per state
  const uint4* crec;         // 2 per dictionary index
  uint32_t shift;            // 32 - log2(buckets)
  uint32_t mask;             // buckets - 1
  uint32_t n_states;
};

// One transition: the packed state after consuming byte b from the state in `packed`.
// load_bucket(h, &keys, &vals) reads the 32-byte bucket h; d1(b) = the packed state one byte below
// the root (every transition whose target is at depth <= 1 depends on the byte alone).
// Straight-line code: one probe, the four key compares are selects; the loop only runs for a key
// that was pushed into the next bucket at build time (~1 % of the buckets are full).  The root
// (state 0) probes like any other state: it has no transition to depth >= 2, and its keys (< 256)
// equal no stored key (>= 256), so the probe misses and d1 stands.
template <class LoadBucket, class LoadD1>
TM_HD uint32_t dfa_next(uint32_t packed, uint32_t b, uint32_t shift, uint32_t mask, LoadBucket load_bucket, LoadD1 d1) {
  const uint32_t key = ((packed & kDfaStateMask) << 8) | b;
  uint32_t h = dfa_hash(key, shift);
  uint32_t res = d1(b);
  bool more;
  do {
    uint4 k, v;
    load_bucket(h, &k, &v);
    const bool h0 = (k.x & ~kDfaOverflow) == key, h1 = k.y == key, h2 = k.z == key, h3 = k.w == key;
    res = h0 ? v.x : (h1 ? v.y : (h2 ? v.z : (h3 ? v.w : res)));
    more = !(h0 | h1 | h2 | h3) & ((k.x & kDfaOverflow) != 0);
    h = (h + 1) & mask;
  } while (more);
  return res;
}

// Everything a kernel needs about one vocabulary (device pointers + scalars).
struct VocabDev {
  const uint2* edges;          // hashed trie edges {key = parent<<8|byte, val = child|sig<<24}
  uint32_t edge_mask;          // capacity - 1
  uint32_t edge_shift;         // 32 - log2(capacity)
  const uint4* rec;            // 2 x uint4 per dictionary index (TokenRecord)
  const uint8_t* smem_image;   // 2304 B: root_child[256] (u32), begin_byte[256], sp_child[256] (u32); staged by TMA
  const uint32_t* root2;       // 65536: val of the node two bytes below the root (first two steps in one load)
  const uint32_t* rev_off;     // decode: reverse[id] offset into rev_blob
  const uint8_t* rev_len;      // 0xFF = nil
  const uint8_t* rev_blob;
  const uint32_t* rev_pack;    // decode: offset into rev_blob8 (8-byte units) | length << 24, 0 = nothing
  const unsigned long long* rev_blob8;
  uint32_t n_info, n_reverse;
  uint32_t unk, delete_token;  // ids or 0xFFFFFF
  uint32_t space_start;        // trie state after the virtual space prefix, or kNoVal
  int max_len;                 // maxTokenLength
  int off;                     // 1, or 2 for UTF-16 (lilbufOffset, go:1031-1034)
  uint8_t capcode, normflag, charset, pad;
  DfaDev dfa;
  const uint8_t* pipe_image;   // 1536 B staged per CTA: dfa_d1[256] (u32), begin_byte[256], class_code[256]
};

TM_HD uint32_t sig_bit_of(uint32_t b) { return (b ^ (b >> 3)) & 7u; }

// Longest key that prefixes window[0 .. maxd): the semantics of pansearch.Fast.LongestSubstring
// (call sites go/tokenmonster.go:1049,1068,1091,...; restated cpp:786-877).  `byte_at(d)` yields
// window byte d, `root_child` is the 256-entry first-byte table, `load_edge(h)` reads edge slot h.
// start = kRootStart for a plain window, or a trie state for the space-prefixed retry.
// Returns the number of WINDOW bytes matched (0 = not found) and the dictionary index.
template <class ByteAt, class RootChild, class LoadEdge>
TM_HD int trie_longest(ByteAt byte_at, RootChild root_child, LoadEdge load_edge, uint32_t edge_mask,
                       uint32_t edge_shift, uint32_t n_info, int maxd, uint32_t start,
                       uint32_t* index_out) {
  int best_len = 0;
  uint32_t best = 0;
  uint32_t node = 0, sig = 0;
  bool at_root = (start == kRootStart);
  if (!at_root) {
    node = start & 0xFFFFFFu;
    sig = start >> 24;
  }
  for (int d = 0; d < maxd; d++) {
    uint32_t b = byte_at(d);
    uint32_t val;
    if (at_root) {
      at_root = false;
      val = root_child(b);
    } else {
      if (((sig >> sig_bit_of(b)) & 1u) == 0) break;  // no child can start with this byte
      uint32_t key = (node << 8) | b;
      uint32_t h = (key * 0x9E3779B1u) >> edge_shift;
      for (;;) {
        uint2 e = load_edge(h);
        if (e.x == key) { val = e.y; break; }
        if (e.x == kNoVal) { val = kNoVal; break; }
        h = (h + 1) & edge_mask;
      }
    }
    if (val == kNoVal) break;
    node = val & 0xFFFFFFu;
    sig = val >> 24;
    if (node < n_info) {
      best = node;
      best_len = d + 1;
    }
  }
  *index_out = best;
  return best_len;
}

// Score of "first token a, then second token b" (go:1075-1084 plain,
// :1096-1105 retry).  nw = a.nWords - forwardDelete, nb = beginByte of the byte after b.
TM_HD int pair_score(int branch_len, uint32_t aflag, int nw, uint32_t bflag, int bnw, int nb, bool retry) {
  int s = branch_len + int(aflag >> 7) + int(bflag >> 7) + (nw > 1 ? nw - 1 : 0) + (bnw > 1 ? bnw - 1 : 0) +
          ((nb >> 2) & 1) + 100 * (nw + bnw + (nb >> 3));
  s -= 100 * int((aflag >> 3) & 1u & (bflag >> 4)) + 3 * int(bflag & 1u & uint32_t(nb));
  if (retry) s -= 103 * int(aflag & 1u) + 1;
  else s += int((bflag >> 2) & 1u) - 103 * int(aflag & 1u & (bflag >> 1));
  return s;
}

}  // namespace tmgpu

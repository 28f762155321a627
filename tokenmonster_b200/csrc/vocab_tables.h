// vocab_tables.h -- host-side: parse a TokenMonster .vocab file and flatten it into the
// tables the sm_100a kernels consume.  Replaces the products of Load
// (/root/reference go/tokenmonster.go:2656-2736): info[], reverse[][], beginByte[256] and the
// pansearch dictionary.  The dictionary is NOT pansearch's bloom/bucket cascade
// (tokenmonster-cpp/src/tokenmonster.cpp:491-1280, a CPU-cache design) but a prefix trie with
// hashed edges: the path to a node is the key, so a walk proves the longest key that prefixes a
// window exactly, with no key-byte compares, and the " "-prefixed retry of go:1088-1107 is just a
// different start node.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tmgpu {

constexpr uint32_t kDNE = 0xFFFFFFu;
constexpr uint32_t kNone = 0xFFFFFFFFu;

// 32-byte token record (one L2 sector), indexed by dictionary index.
//  w0: flag | nWords<<8 | key_len<<16 | alt_len<<24      (alt_len 0 = no alternative 1)
//  w1: id   w2: id1 (alternative 1)   w3: id2 (alternative 2)
//  w4: alt_len2 | alt1.flag<<8 | alt1.nWords<<16 | alt2.flag<<24
//  w5: alt2.nWords | (spare)
//  w6: alt index  w7: alt2 index (diagnostics)
struct TokenRecord {
  uint32_t w[8];
};

struct HostTables {
  // header
  uint8_t capcode = 0, charset = 0, normflag = 0, level = 0, reserve = 0;
  uint32_t unk = kDNE, vocab_size = 0, n_reverse = 0, n_info = 0, delete_token = kDNE;
  int max_len = 0;
  uint8_t begin_byte[256] = {};

  std::vector<TokenRecord> rec;  // n_info
  // trie: node ids < n_info are terminal (id == dictionary index); others are interior.
  // edge table: open addressing, key = parent<<8 | byte, val = child | sig(child)<<24
  std::vector<uint64_t> edges;   // lo 32 = key, hi 32 = val; empty = all ones
  uint32_t edge_bits = 0;        // log2(capacity)
  uint32_t root_child[256];      // val or kNone
  uint32_t space_start = kNone;  // trie state after the virtual " " (0x20 0x00 in UTF-16) prefix
  // shortcuts for the first steps of a walk (the top of the trie is where every walk goes):
  std::vector<uint32_t> root2;   // 65536: val of the node two bytes below the root, or kNone
  uint32_t sp_child[256];        // val of space_start's child by byte, or kNone
  uint32_t n_nodes = 0;

  // ---- reversed-key automaton (position-table pipeline, stage 2) -------------------------------
  // Aho-Corasick over the REVERSED keys, run right to left over the normalized text: after the byte
  // at position p has been consumed the state names the longest suffix-of-a-key that starts at p,
  // and its first terminal on the failure chain is the longest KEY that starts at p -- the result
  // of pansearch.Fast.LongestSubstring for the window at p (call sites go/tokenmonster.go:1049,
  // 1068,...) with ONE transition per byte instead of one trie walk per window.
  //   transition value ("packed state"): state | best_len << 20 | sp_len << 26
  //     best_len = length of the longest key starting here (0 = none)
  //     sp_len   = length of the longest key that prefixes  virtual space + text from here
  //                (the lilbuf retry of go:1088-1107; includes the `off` prefix bytes; 0 = none)
  //   dfa_d1[b]            default: every transition whose target is at depth <= 1 depends only on
  //                        the byte (the node one byte below the root, or the root)
  //   dfa_buckets          hashed (state<<8|b) -> packed target for targets at depth >= 2;
  //                        32-byte buckets of 4 keys + 4 values, bit 31 of key 0 = "overflowed
  //                        into the next bucket", empty key = kDfaEmpty
  //   dfa_hot[state]       state record, what the chain stage reads per look-ahead position (8 x u32):
  //                        w0 = best index | best_len<<24, w1 = sp index | sp_len<<24,
  //                        w2 = bit 0 best.flag>>1, bit 1 best.flag>>4, bit 2 retry eligible (best
  //                        begins with a letter and has no word boundary, go:1088), bit 3 sp.flag>>4,
  //                        w4,w5 = B(best, nb) for the four classes nb, int16 each,
  //                        w6,w7 = B(sp, nb) of the retry, int16 each (pipeline_core.h, stage 3)
  //   chain_rec[index]     chain record (8 x u32): w0 = id | bits<<24 (flag&1, flag>>3, flag>>5 of the
  //                        token, then flag&1, flag>>3 of alternative 1 and 2), w1 = id1 | alt_len1<<24,
  //                        w2 = id2 | alt_len2<<24, w3..w5 = A(token / alt 1 / alt 2, forward delete
  //                        0 | 1) as int16 pairs
  uint32_t dfa_states = 0;
  uint32_t dfa_bits = 0;
  std::vector<uint32_t> dfa_buckets;
  uint32_t dfa_d1[256];
  std::vector<uint32_t> dfa_hot;
  std::vector<uint32_t> chain_rec;
  uint8_t class_code[256];       // begin_byte as a 2-bit code: 0 -> 0, 1 -> 1, 10 -> 2, 12 -> 3

  // reverse table for decode: reverse[id] = LAST record with that id (go:2715)
  std::vector<uint32_t> rev_off; // n_reverse, offset into rev_blob
  std::vector<uint8_t> rev_len;  // n_reverse, 0xFF = nil
  std::vector<uint8_t> rev_blob;
  // the same, laid out for the decode kernels: one word per id = offset into rev_blob8 in 8-byte
  // units | length << 24 (0 = nil / no bytes); every key starts on an 8-byte boundary
  std::vector<uint32_t> rev_pack;
  std::vector<uint64_t> rev_blob8;
};

// returns 0 or a TMGPU_E_* code; err receives a message
int parse_vocab(const uint8_t* file, size_t n, HostTables* out, std::string* err);

// host walk of the device trie for one window (table self-check; see tmgpu_selftest_trie_lookup)
int host_trie_lookup(const HostTables& h, const uint8_t* window, int n, int spaced, uint32_t* index);

// host run of the reversed automaton over a window: best match at window[0] (table self-check)
int host_dfa_lookup(const HostTables& h, const uint8_t* window, int n, uint32_t* index, uint32_t* sp_index,
                    int* sp_len);

inline uint32_t sig_bit(uint32_t b) { return (b ^ (b >> 3)) & 7u; }
inline uint32_t edge_hash(uint32_t key, uint32_t bits) { return (key * 0x9E3779B1u) >> (32u - bits); }

}  // namespace tmgpu

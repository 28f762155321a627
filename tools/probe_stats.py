"""How many hash-table loads a trie step needs on real walks (CPU model of the device edge table).

For the c2 vocabulary: rebuilds the hashed edge table exactly like csrc/vocab_tables.cpp (same key,
same multiplicative hash, linear probing at load factor <= 1/4), replays the walks the tokenizer
performs on sample text (start of every token, hashed steps only: depth >= 2) and reports the
distribution of loads per step, for the current hash and for a murmur-style finalizer.

    python tools/probe_stats.py
"""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from workloads import configs, corpus  # noqa: E402

M32 = 0xFFFFFFFF


def h_mul(key, bits):
    return ((key * 0x9E3779B1) & M32) >> (32 - bits)


def h_mix(key, bits):
    k = key
    k ^= k >> 16
    k = (k * 0x85EBCA6B) & M32
    k ^= k >> 13
    k = (k * 0xC2B2AE35) & M32
    k ^= k >> 16
    return k >> (32 - bits)


def main():
    vb = configs.vocab_bytes("c2")
    n_info = vb[17] | vb[18] << 8 | vb[19] << 16
    p, keys = 24, []
    for _ in range(n_info):
        l = vb[p]
        keys.append(vb[p + 1:p + 1 + l])
        p += 1 + l + 2 + 9 + 4
    node = {k: i for i, k in enumerate(keys)}
    nxt = n_info
    for k in keys:
        for d in range(1, len(k)):
            if k[:d] not in node:
                node[k[:d]] = nxt
                nxt += 1
    edges = [((node[k[:-1]] << 8) | k[-1], v) for k, v in node.items() if len(k) > 1]
    bits = 4
    while (1 << bits) < len(edges) * 4 + 2:
        bits += 1
    o = oracle.OracleVocab(vb, "port")
    text = o.normalize(bytes(corpus.base_text(5, 1 << 19, 0.3)))
    ids, _ = o.tokenize(text, normalized=True)
    starts, pos = [], 0
    for t in ids:
        starts.append(pos)
        pos += len(o.decode([int(t)], raw=True))
    for name, h in (("multiplicative (current)", h_mul), ("murmur finalizer", h_mix)):
        table = {}
        for key, v in edges:
            s = h(key, bits)
            while s in table:
                s = (s + 1) & ((1 << bits) - 1)
            table[s] = key
        loads = collections.Counter()
        for a in starts:
            cur, d = b"", 0
            while a + d < len(text) and d < 40:
                nk = text[a:a + d + 1]
                if nk not in node:
                    break
                if d >= 2:                                   # depth 0 and 1 come from the root tables
                    key = (node[text[a:a + d]] << 8) | text[a + d]
                    s, n = h(key, bits), 1
                    while table.get(s) != key:
                        s = (s + 1) & ((1 << bits) - 1)
                        n += 1
                    loads[n] += 1
                d += 1
        tot = sum(loads.values())
        print(f"{name:28s} steps {tot}: " + ", ".join(f"{k} load(s) {100 * v / tot:.1f}%" for k, v in sorted(loads.items())[:5]))


if __name__ == "__main__":
    main()

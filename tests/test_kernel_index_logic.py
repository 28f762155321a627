"""Index arithmetic of two kernels whose data movement cannot run without a GPU, restated in Python
and checked exhaustively-at-random for the invariants the kernels rely on (every element written
exactly once, nothing outside the segment read or written, vector stores aligned).  The restatements
follow csrc/kernels.cu: compact_bytes_kernel and csrc/pipeline_kernels.cuh: chain_kernel (the
sector-wise id stores); the GPU suite checks the kernels themselves."""
import random


def _compact(aoff, boff, l, src):
    """compact_bytes_kernel for one segment: aligned 32-bit destination words, each the funnel shift of
    two aligned source words; the bytes around them one by one."""
    reads, out = set(), {}
    head = min((4 - (boff & 3)) & 3, l)
    sa = aoff + head
    sh = sa & 3
    sw = sa - sh
    rest = l - head
    i0 = 1 if sh > head else 0
    nw = ((rest + sh - 4) >> 2 if rest + sh >= 8 else 0) if sh else rest >> 2
    first_end = min(head + 4 * i0, l)
    for k in range(first_end):
        reads.add(aoff + k)
        out[boff + k] = src[aoff + k]
    for i in range(i0, nw):
        words = [sw + 4 * i + t for t in range(4)] + ([sw + 4 * i + 4 + t for t in range(4)] if sh else [])
        reads.update(words)
        byts = [src[x] for x in words] + ([] if sh else [0] * 4)
        assert (boff + head + 4 * i) % 4 == 0
        for t in range(4):
            d = boff + head + 4 * i + t
            assert d not in out
            out[d] = byts[sh + t]
    for k in range(max(head + 4 * nw, first_end), l):
        reads.add(aoff + k)
        assert boff + k not in out
        out[boff + k] = src[aoff + k]
    assert all(aoff <= r < aoff + l for r in reads)
    assert len(out) == l and all(out[boff + k] == src[aoff + k] for k in range(l))


def test_compact_bytes_word_copy():
    rng = random.Random(1)
    buf = [rng.randrange(256) for _ in range(2000)]
    for _ in range(40000):
        l = rng.randrange(0, 40) if rng.random() < 0.8 else rng.randrange(0, 600)
        _compact(rng.randrange(0, 64) + 100, rng.randrange(0, 64), l, buf)


def _chain_ids(base_addr_words, steps, cap):
    """chain_kernel<STORE>: ids staged in a 16-word row, whole aligned groups of eight stored after the
    step that finishes them, the ends of the string one by one."""
    base_lo = base_addr_words & 7
    row, slab, emitted = [None] * 16, {}, []
    stored = 0
    for n_ids in steps:
        g_before = base_lo + min(stored, cap)
        for _ in range(n_ids):
            if stored < cap:
                row[(base_lo + stored) & 15] = len(emitted)
            emitted.append(len(emitted))
            stored += 1
        g_after = base_lo + min(stored, cap)
        m = g_after >> 3
        done = m != (g_before >> 3)
        whole = done and (m > 1 or base_lo == 0)
        if whole:
            start = base_addr_words - base_lo + 8 * (m - 1)
            assert start % 8 == 0
            for j in range(8):
                k = start + j - base_addr_words
                assert 0 <= k < min(stored, cap) and start + j not in slab
                slab[start + j] = row[((m - 1) & 1) * 8 + j]
        if done and not whole:
            for k in range(8 - base_lo):
                assert base_addr_words + k not in slab
                slab[base_addr_words + k] = row[(base_lo + k) & 15]
    cntc = min(stored, cap)
    rem = (base_lo + cntc) & 7
    for k in range(cntc - rem if cntc > rem else 0, cntc):
        assert base_addr_words + k not in slab
        slab[base_addr_words + k] = row[(base_lo + k) & 15]
    assert len(slab) == cntc and all(slab[base_addr_words + k] == emitted[k] for k in range(cntc))


def test_chain_sector_stores():
    rng = random.Random(3)
    for _ in range(40000):
        steps = [rng.choice([1, 1, 1, 2]) for _ in range(rng.randrange(0, 40))]
        tot = sum(steps)
        cap = rng.choice([tot + 16, max(0, tot - rng.randrange(0, 5)), rng.randrange(0, 20)])
        _chain_ids(rng.randrange(0, 64) + 64, steps, cap)

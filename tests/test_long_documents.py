"""Few-but-huge inputs: documents of a small batch are cut into speculative
chunks, stitched exactly, and must equal the oracle's sequential run bit for bit."""
import os

import numpy as np
import pytest

from oracle import oracle
from workloads import configs, corpus, vocabgen


def flat(texts):
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    return np.frombuffer(b"".join(texts), np.uint8), off


def doc(seed, nbytes, code=0.4):
    return bytes(corpus.base_text(seed, nbytes, code))


def check(v, chk, texts, enc=2):
    data, off = flat(texts)
    r = v.tokenize_batch(data, off, enc)
    stats = v.last_stats()
    counts, cmiss, cstat = v.count_batch(data, off)
    for s, t in enumerate(texts):
        try:
            want, miss = chk.tokenize(t)
            wc = chk.tokenize(t, mode=1)[0]
            st = 0
        except oracle.OracleError as e:
            want, miss, wc, st = np.zeros(0, np.uint32), 0, 0, e.status
        if st == 13 and any(b >= 0x80 for b in t):
            continue                                   # non-ASCII: covered by tests/test_unicode.py
        got = r[s].astype(np.uint32)
        assert len(got) == len(want), (s, len(got), len(want))
        assert np.array_equal(got, want), (s, int(np.argmax(got != want)))
        assert int(r.missing[s]) == miss and int(r.status[s]) == st
        assert int(counts[s]) == wc and int(cmiss[s]) == miss
    return r, stats


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "c2", "c4", "c3"])
def test_single_megabyte_document(built, name):
    from tokenmonster_b200 import Vocab
    vb = configs.vocab_bytes(name)
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    text = doc(41, 1 << 20, 0.3 if name != "c4" else 0.9)
    r, stats = check(v, chk, [text])
    assert stats["chunks"] >= 800 and stats["chunk_redo"] == 0, stats
    back, boff = v.decode_batch(r.raw, r.tok_off, 2)
    assert back.tobytes() == text
    v.close()


@pytest.mark.gpu
def test_mixed_lengths_fallbacks_and_unk(built):
    """Short and long documents together; a document without any safe boundary (planning fails ->
    one lane), one with a non-ASCII byte in the middle (-> the non-ASCII pass, whole document),
    and a vocabulary with an unk token and bytes it does not cover (missing counts are stitched)."""
    from tokenmonster_b200 import Vocab
    vb = configs.vocab_bytes("c2")
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    long_a, long_b = doc(1, 300_000), doc(2, 50_000, 0.9)
    texts = [b"", b"short one", doc(3, 9_000), long_a, b"x" * 30_000, long_b, doc(4, 4_097),
             long_b[:20_000] + "é".encode() + long_b[20_000:40_000], doc(5, 4_096), b"A" * 20_000 + b"b",
             (b"HTTP/2 " * 4000), b"\n" * 10_000, b"a.b" * 10_000]
    r, stats = check(v, chk, texts)
    assert stats["chunks"] > 100 and stats["chunk_redo"] >= 1, stats
    assert int(r.status[7]) == 0 and stats["unicode_strings"] == 1
    v.close()
    # unk token + uncovered bytes: capcode 0, only lower-case letters and space as single bytes
    toks = [b" the", b"the", b" a", b"an", b" and", b"nd", b" qu", b"ick", b" brown", b"fox", b"o", b"x"]
    vb2 = vocabgen.generate_vocab(toks, capcode=0, charset=1, normflag=0, level=2, reserve=32, unk=True)
    v2, chk2 = Vocab(vb2, device=0), oracle.OracleVocab(vb2, "port")
    rng = np.random.default_rng(7)
    words = [b"the", b" quick", b" brown", b" fox", b" and", b" an", b"?", b" Z", b"\n", b"."]
    big = b"".join(words[k] for k in rng.integers(0, len(words), 60_000))
    r2, stats2 = check(v2, chk2, [big, big[:5000], big[1234:100_000]])
    assert stats2["chunks"] > 50 and int(r2.missing[0]) > 1000
    v2.close()


@pytest.mark.gpu
def test_long_documents_through_wire_and_device_paths(built):
    import struct
    import torch
    from tokenmonster_b200 import Vocab
    from tokenmonster_b200.device import DeviceWorkspace
    vb = configs.vocab_bytes("c2")
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    texts = [doc(11, 200_000), b"tiny", doc(12, 70_000, 0.8)]
    req = struct.pack("<I", len(texts)) + b"".join(struct.pack("<Q", len(t)) + t for t in texts)
    st, reply = v.wire_tokenize(req)
    assert st == 0 and v.last_stats()["chunks"] > 50
    at = 4
    for t in texts:
        l = struct.unpack_from("<Q", reply, at)[0]
        at += 8
        assert np.array_equal(np.frombuffer(reply[at:at + l], np.uint16).astype(np.uint32), chk.tokenize(t)[0])
        at += l
    data, off = flat(texts)
    ws = DeviceWorkspace(v, len(texts), len(data))
    ws.launch_tokenize(torch.from_numpy(data.copy()).cuda(), torch.from_numpy(off.astype(np.int64)).cuda(), 4, 0)
    res = ws.results()
    ids = res["tokens_u8"].cpu().numpy().view(np.uint32)
    assert np.array_equal(ids, np.concatenate([chk.tokenize(t)[0] for t in texts]))
    ws.close()
    v.close()

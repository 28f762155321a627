"""Non-ASCII text: the device normalizer (csrc/unicode_normalize.cuh) against the Python oracle
(oracle/unicode_oracle.py = unicodedata NFD + a per-code-point restatement of
javascript/tokenmonster.js:900-1065), and the generated character tables against unicodedata.
Parity for non-ASCII is unpinned against the reference itself (its normalizer and capcode
packages are not in /root/reference); see the oracle's header."""
import ctypes as C
import os
import random
import re
import unicodedata as ud
from collections import Counter

import numpy as np
import pytest

from oracle import oracle
from oracle import unicode_oracle as uo
from workloads import capcode_py, corpus, vocabgen

SAMPLES = [
    "Le café de l'École polytechnique ÉTAIT fermé, naïve façade. ",
    "Москва — СТОЛИЦА России; ДомДОМ дом Ёлка ёж. ",
    "Ελληνικά ΚΕΦΑΛΑΙΑ καὶ ᾅδης ΐ ΰ ᾂ ᾷ. ",
    "日本語のテキストと한국어 텍스트, 中文文本。",
    "emoji 😀🎉 and ZWJ 👩‍💻 flags 🇫🇷. ",
    "Ångström Å Ω K ǅ ǆ Ǖ ǖ ß ſ İ ı ŉ. ",
    "don’t Can’t O’NEIL x’Y it's IT'S. ",
    "ạ́ ạ́ q̣̇ ̈́ ṩ ṩ. ",
    "HTTPСервер ÉCOLEs ÀB1C d2É ÉÉé 12Éa É1 ΑΒγ. ",
    "१२३ ٤٥٦ x१ y٤z ① ½. ",
    "tab\there\nnew line\r\n  double  space. ",
]


def unicode_texts(seed, n, lo=0, hi=300):
    rng = random.Random(seed)
    base = bytes(corpus.base_text(seed, 1 << 16, 0.3)).decode("ascii")
    out = []
    for _ in range(n):
        parts, size, want = [], 0, rng.randrange(lo, hi + 1)
        while size < want:
            if rng.random() < 0.5:
                a = rng.randrange(0, len(base) - 80)
                p = base[a:a + rng.randrange(1, 80)]
            else:
                smp = rng.choice(SAMPLES)
                a = rng.randrange(0, len(smp))
                p = smp[a:a + rng.randrange(1, 40)]
            parts.append(p)
            size += len(p.encode())
        out.append("".join(parts).encode())
    return out


def make_unicode_vocab(capcode, normflag, n_tokens=3000):
    """A stand-in vocabulary whose tokens come from normalized non-ASCII text (multi-byte keys)."""
    sample = b"".join(unicode_texts(99, 600, 100, 400))
    norm = uo.normalize(sample, normflag, capcode).decode()
    pieces = re.findall(r" ?\w+| ?[^\w\s]+|\s+|[DCW] ", norm)
    cnt = Counter()
    for p in pieces:
        b = p.encode()
        if len(b) <= 32:
            cnt[b] += len(b)
        for k in range(2, min(len(p), 6)):
            cnt[p[:k].encode()] += 1
    for a, b in zip(pieces, pieces[1:]):
        ab = (a + b).encode()
        if len(ab) <= 40:
            cnt[ab] += 1
    toks = dict(cnt.most_common(n_tokens))
    return vocabgen.generate_vocab(list(toks), capcode=capcode, charset=1, normflag=normflag, level=2,
                                   reserve=1, scores={k: float(v) for k, v in toks.items()})


# ---------------------------------------------------------------------------------------------
# CPU: oracle self-consistency and the generated tables
# ---------------------------------------------------------------------------------------------
def test_unicode_oracle_agrees_with_ascii_oracles_and_is_lossless():
    rng = random.Random(5)
    base = bytes(corpus.base_text(3, 1 << 16, 0.4))
    for _ in range(200):
        a = rng.randrange(0, len(base) - 300)
        s = base[a:a + rng.randrange(0, 300)]
        e = uo.capcode_encode(s.decode()).encode()
        assert e == capcode_py.encode(s) == oracle.capcode_encode(s)
        assert uo.capcode_decode(e.decode()).encode() == s
    for t in unicode_texts(7, 300):
        nfd = uo.normalize_flags(t, 1)
        assert uo.capcode_decode(uo.capcode_encode(nfd)) == nfd          # capcode is lossless on NFD text
        assert ud.normalize("NFD", nfd) == nfd
    # hand-checked vectors + their non-ASCII siblings
    assert uo.normalize(b"CamelCase", 1, 2) == b"DC camelDC case"
    assert uo.normalize("É".encode(), 1, 2) == "DW é".encode()
    assert uo.normalize(" ÉCOLE d'Été".encode(), 1, 2) == "W école d'DC été".encode()
    assert uo.normalize("ДОМ Дом".encode(), 1, 2) == "DW домC дом".encode()
    assert uo.normalize("Été".encode(), 7, 0) == b"ete"
    with pytest.raises(uo.Refused):
        uo.normalize(b"\xff", 1, 2)
    with pytest.raises(uo.Refused):
        uo.normalize(("a" + "́" * 31).encode(), 1, 0)


def test_generated_unicode_tables_match_unicodedata(built):
    from tokenmonster_b200 import _lib
    lib = _lib.lib()
    lo, up, n = C.c_uint32(), C.c_uint32(), C.c_int()
    dec = (C.c_uint32 * 4)()
    cls = {"Lu": 1, "Ll": 2, "Lt": 3, "Lm": 3, "Lo": 3, "Nd": 4, "Mn": 5, "Mc": 6, "Me": 6}
    step = 1 if os.environ.get("TM_FULL_UNICODE") else 3          # every 3rd code point + all of the BMP's first 12k
    for cp in list(range(0, 0x3000)) + list(range(0x3000, 0x110000, step)):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        ch = chr(cp)
        prop = lib.tmgpu_selftest_unicode(cp, C.byref(lo), C.byref(up), dec, C.byref(n))
        assert prop & 7 == cls.get(ud.category(ch), 0), hex(cp)
        assert prop >> 8 == ud.combining(ch), hex(cp)
        assert chr(lo.value) == uo._simple_lower(ch) and chr(up.value) == uo._simple_upper(ch), hex(cp)
        d = ud.decomposition(ch)
        if d and not d.startswith("<") and not (0xAC00 <= cp < 0xAC00 + 11172):
            assert "".join(chr(x) for x in dec[:n.value]) == ud.normalize("NFD", ch), hex(cp)
        else:
            assert n.value == 0


# ---------------------------------------------------------------------------------------------
# GPU: tokenize through the C ABI, oracle = python normalization + C tokenizer on normalized text
# ---------------------------------------------------------------------------------------------
def expect(chk, raw, normflag, capcode):
    try:
        norm = uo.normalize(raw, normflag, capcode)
    except uo.Refused:
        return np.zeros(0, np.uint32), 0, 13
    ids, miss = chk.tokenize(norm, normalized=True)
    return ids, miss, 0


EDGE = [b"", b"\xff", b"ok \xc3", b"\xc0\xaf", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", "é".encode(), "É".encode(),
        ("a" + "́" * 30 + "b").encode(), ("a" + "́" * 31 + "b").encode(), "̣́".encode(),
        " É".encode(), "x É".encode(), "’".encode(), "Ǆ".encode(), "ǅ".encode(), "İstanbul".encode(),
        "ΑΒΓδ".encode(), "한".encode(), "각".encode(), "\U0001D15E".encode(), "ÀÉÎ".encode() * 50]


@pytest.mark.gpu
@pytest.mark.parametrize("capcode,normflag", [(2, 1), (0, 1), (2, 7), (0, 3), (2, 5)])
def test_non_ascii_strings_against_oracle(built, capcode, normflag):
    from tokenmonster_b200 import Vocab
    vb = make_unicode_vocab(capcode, normflag)
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    texts = unicode_texts(11 + capcode + normflag, 500) + EDGE + [s.encode() for s in SAMPLES]
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    data = np.frombuffer(b"".join(texts), np.uint8)
    r = v.tokenize_batch(data, off, 4)
    counts, cmiss, cstat = v.count_batch(data, off)
    n_uni = 0
    for s, t in enumerate(texts):
        want, miss, st = expect(chk, t, normflag, capcode)
        assert int(r.status[s]) == st, (s, t)
        assert np.array_equal(r[s].astype(np.uint32), want), (s, t)
        assert int(r.missing[s]) == miss and int(cstat[s]) == st and int(cmiss[s]) == miss
        if st == 0:
            norm = uo.normalize(t, normflag, capcode)
            assert int(counts[s]) == chk.tokenize(norm, normalized=True, mode=1)[0]
            n_uni += any(b >= 0x80 for b in t)
    assert n_uni > 200
    v.close()


@pytest.mark.gpu
def test_non_ascii_batch_large_and_streamed(built):
    """Every third string non-ASCII in a batch big enough for the chunked-upload path; the ASCII
    strings must be untouched by the second pass and the rest must match the oracle (sampled)."""
    from tokenmonster_b200 import Vocab
    vb = make_unicode_vocab(2, 1)
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    pool = unicode_texts(5, 3000, 20, 400)
    asc = [bytes(t) for t in unicode_texts(6, 3000, 20, 400)]
    asc = [bytes(b for b in t if b < 0x80) for t in asc]
    rng = random.Random(2)
    texts = [(pool if k % 3 == 0 else asc)[rng.randrange(3000)] for k in range(150000)]
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    data = np.frombuffer(b"".join(texts), np.uint8)
    r = v.tokenize_batch(data, off, 2)
    cache = {}
    for s in list(range(0, 3000)) + [rng.randrange(len(texts)) for _ in range(3000)]:
        t = texts[s]
        if t not in cache:
            cache[t] = expect(chk, t, 1, 2)
        want, miss, st = cache[t]
        assert int(r.status[s]) == st and np.array_equal(r[s].astype(np.uint32), want), s
    v.close()


@pytest.mark.gpu
@pytest.mark.parametrize("capcode,normflag", [(2, 1), (0, 1), (2, 7)])
def test_non_ascii_decode_and_round_trip(built, capcode, normflag):
    """Decode (capcode decoder with Unicode classes and simple upper-casing) against the oracle, the
    NFD round trip, and the streaming decoder fed one id at a time."""
    from tokenmonster_b200 import Vocab
    vb = make_unicode_vocab(capcode, normflag)
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    grow = [" ⱯⱯⱯⱯⱯⱯⱯⱯ ⱭⱭⱭ".encode(), "ȾȾȾ ⱯⱭ x".encode(), " ǅ ǄǄ ǆ".encode()]      # capitals longer than their lower case
    texts = [t for t in unicode_texts(31 + capcode, 300) + EDGE + grow if expect(chk, t, normflag, capcode)[2] == 0]
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    r = v.tokenize_batch(np.frombuffer(b"".join(texts), np.uint8), off, 4)
    back, boff = v.decode_batch(r.raw, r.tok_off, 4)
    for s, t in enumerate(texts):
        got = back[int(boff[s]):int(boff[s + 1])].tobytes()
        keys = chk.decode(r[s], raw=True)                       # the concatenated token keys
        want = uo.capcode_decode(keys.decode("utf-8")).encode() if capcode == 2 else keys
        assert got == want, (s, t)
        if normflag == 1:
            assert got.decode() == ud.normalize("NFD", t.decode()), (s, t)   # lossless up to NFD
    # streaming: one id per call, state (partial UTF-8 + capcode flags) carried on the device
    some = [s for s in range(len(texts)) if 0 < len(r[s]) <= 60][:48]
    dec = v.decoder(len(some))
    outs = [b""] * len(some)
    for step in range(max(len(r[s]) for s in some)):
        chunk = [r[s][step:step + 1] for s in some]
        for k, piece in enumerate(dec.decode_batch(chunk)):
            outs[k] += piece
    tail = dec.flush()
    for k, s in enumerate(some):
        assert outs[k] + tail[k] == back[int(boff[s]):int(boff[s + 1])].tobytes(), (k, texts[s])
    dec.close()
    v.close()


# ---------------------------------------------------------------------------------------------
# committed fixtures (tests/golden/unicode_cases.json, frozen oracle answers on Unicode 15.0)
# ---------------------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _unicode_fixture():
    import json
    with open(os.path.join(GOLD, "unicode_cases.json")) as f:
        cases = json.load(f)
    with open(os.path.join(GOLD, "unicode_cap2.vocab"), "rb") as f:
        vb = f.read()
    return cases, vb


def test_oracle_reproduces_committed_unicode_fixtures():
    cases, vb = _unicode_fixture()
    assert cases["unicode"] == "15.0.0" and make_unicode_vocab(2, 1) == vb     # the stand-in is deterministic
    chk = oracle.OracleVocab(vb, "port")
    for row in cases["rows"]:
        t = bytes.fromhex(row["in"])
        ids, miss, st = expect(chk, t, 1, 2)
        assert st == row["status"]
        if st == 0:
            assert uo.normalize(t, 1, 2).hex() == row["norm"]
            assert ids.tolist() == row["ids"] and miss == row["missing"]


@pytest.mark.gpu
def test_device_matches_committed_unicode_fixtures(built):
    from tokenmonster_b200 import Vocab
    cases, vb = _unicode_fixture()
    rows = cases["rows"]
    texts = [bytes.fromhex(r["in"]) for r in rows]
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    data = np.frombuffer(b"".join(texts), np.uint8)
    v = Vocab(vb, device=0)
    r = v.tokenize_batch(data, off, 4)
    counts, _, _ = v.count_batch(data, off)
    back, boff = v.decode_batch(r.raw, r.tok_off, 4)
    scores, hist, tit, status = v.score_batch(data, off)
    want_scores = np.zeros_like(scores)
    want_tit = 0
    for s, row in enumerate(rows):
        assert int(r.status[s]) == row["status"] == int(status[s]), s
        if row["status"]:
            assert len(r[s]) == 0
            continue
        assert r[s].astype(np.uint32).tolist() == row["ids"], s
        assert int(r.missing[s]) == row["missing"] and int(counts[s]) == row["count"], s
        assert back[int(boff[s]):int(boff[s + 1])].tobytes().hex() == row["decoded"], s
        for k, val in row["scores"].items():
            want_scores[int(k)] += val
        want_tit += row["tokens_in_text"]
    assert np.array_equal(scores, want_scores) and tit == want_tit
    v.close()

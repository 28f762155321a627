"""Pins the CPU oracle (oracle/tm_oracle.c) -- no GPU needed.

1. every assertion of the reference's tests/unit.cpp:87-119
2. the committed golden fixtures, which are outputs of the reference's own C++ runtime
   (tests/golden/make_golden.py, oracle/_ref)
3. when oracle/_ref is available (build container): randomized differential runs of the port
   against the reference itself, including freshly generated vocabularies
4. capcode vectors of SURVEY.md Appendix E/F and the second (Python) restatement
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle
from workloads import capcode_py, corpus, vocabgen

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KINDS = ["port"] + (["reference"] if oracle.have_reference() else [])


def gold_vocab(name):
    if name == "unit_cpp":
        return vocabgen.unit_cpp_vocab()
    with open(os.path.join(GOLD, name + ".vocab"), "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def cases(built):
    with open(os.path.join(GOLD, "cases.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("kind", KINDS)
def test_unit_cpp_known_answers(built, kind):
    """tokenmonster-cpp/tests/unit.cpp:87-119, assertion by assertion."""
    v = oracle.OracleVocab(vocabgen.unit_cpp_vocab(), kind)
    assert v.vocab_size == 4 and v.max_token_length == 2
    ids, missing = v.tokenize(b"ab a z")
    assert ids.tolist() == [3, 0, 1, 0] and missing == 1
    assert v.decode(ids) == b"ab a "
    ser, _ = v.tokenize(b"ab", mode=2)
    assert ser == bytes([3, 0])
    assert v.decode(ser, enc_len=2) == b"ab"
    assert v.tokenize(b"ab a z", mode=1) == (4, 1)
    d = v.new_decoder()
    assert d.decode(np.array([3, 0, 1], np.uint32)) == b"ab a"
    assert d.decode(bytes([0, 0]), enc_len=2) == b" "
    assert d.flush() == b""


@pytest.mark.parametrize("name", ["unit_cpp", "small_cap2", "small_nocap", "gpt2"])
def test_port_matches_reference_golden(cases, name):
    v = oracle.OracleVocab(gold_vocab(name), "port")
    info = cases[name]["info"]
    assert [v.capcode, v.charset, v.normflag, v.vocab_size, v.n_reverse, v.n_info, v.max_token_length,
            v.delete_token, v.unk] == info
    for row in cases[name]["rows"]:
        data = bytes.fromhex(row["in"])
        ids, miss = v.tokenize(data)
        assert ids.tolist() == row["ids"] and miss == row["missing"]
        assert v.tokenize(data, mode=1) == (row["count"], row["missing"])
        assert v.normalize(data).hex() == row["normalized"]
        assert v.tokenize(bytes.fromhex(row["normalized"]), normalized=True)[0].tolist() == row["ids"]
        assert v.decode(ids).hex() == row["decoded"]
        assert v.decode(ids, raw=True).hex() == row["decoded_raw"]
        for enc in (2, 3, 4):
            if f"ser{enc}" in row:
                assert v.tokenize(data, mode=enc)[0].hex() == row[f"ser{enc}"]
                assert v.decode(bytes.fromhex(row[f"ser{enc}"]), enc_len=enc).hex() == row[f"dec_ser{enc}"]
        d = v.new_decoder()
        k1, k2 = len(ids) // 3, (2 * len(ids)) // 3
        got = [d.decode(ids[:k1]).hex(), d.decode(ids[k1:k2]).hex(), d.decode(ids[k2:]).hex(), d.flush().hex()]
        assert got == row["stream"]


def _random_vocab(rng, capcode):
    text = corpus.base_text(rng.randrange(1 << 30), 60_000, rng.choice([0.0, 0.4, 0.9]), pool=800)
    if capcode == 2:
        text = capcode_py.encode(text)
    toks = corpus.train_tokens(text, rng.choice([150, 400, 900]), max_len=rng.choice([6, 12, 30]))
    reserve = rng.choice([1, 8, 4 | 32, 16 | 32]) if capcode else rng.choice([1, 8, 4])
    return vocabgen.generate_vocab(list(toks), capcode=capcode, charset=1, normflag=rng.choice([0, 1]) if not capcode else 1,
                                   level=3, reserve=reserve, scores=toks, unk=rng.random() < 0.5)


@pytest.mark.skipif(not oracle.have_reference(), reason="needs oracle/_ref (reference sources present)")
@pytest.mark.parametrize("seed", range(6))
def test_port_vs_reference_random_vocabs(built, seed):
    rng = random.Random(1000 + seed)
    capcode = 2 if seed % 2 else 0
    vb = _random_vocab(rng, capcode)
    port, ref = oracle.OracleVocab(vb, "port"), oracle.OracleVocab(vb, "reference")
    text = corpus.base_text(rng.randrange(1 << 30), 80_000, 0.5, pool=1200)
    for _ in range(120):
        a = rng.randrange(len(text) - 900)
        data = text[a:a + rng.randrange(0, 800)]
        ip, mp = port.tokenize(data)
        ir, mr = ref.tokenize(data)
        assert np.array_equal(ip, ir) and mp == mr
        assert port.tokenize(data, mode=1) == ref.tokenize(data, mode=1)
        assert port.decode(ip) == ref.decode(ir)
        enc = rng.choice([2, 3, 4])
        assert port.tokenize(data, mode=enc) == ref.tokenize(data, mode=enc)


def test_capcode_vectors(built):
    """Examples derived in SURVEY.md Appendix E/F from javascript/tokenmonster.js:900-1005."""
    vec = {b"hello": b"D hello", b" hello": b" hello", b"Hello": b"DC hello", b" Hello": b"C hello",
           b" HELLO ": b"W hello ", b"x1": b"D xD 1", b"CamelCase": b"DC camelDC case",
           b" MIXEDCaseABC": b"C mDC iDC xDC eDC dDC caseDW abc", b"HTTP/2": b"DW http/D 2"}
    for src, want in vec.items():
        assert oracle.capcode_encode(src) == want
        assert capcode_py.encode(src) == want
        assert capcode_py.decode(want) == src


def test_capcode_two_restatements_agree(built):
    rng = random.Random(5)
    alphabet = b"aAbBzZ 019'.,-_\n\tXy"
    for _ in range(3000):
        s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
        e = oracle.capcode_encode(s)
        assert e == capcode_py.encode(s)
        assert capcode_py.decode(e) == s   # lossless


def test_incomplete_utf8_and_utf16(built):
    """go/tokenmonster.go:149-219 (note quirk Q7: a truncated sequence reports the MISSING count)."""
    assert oracle.incomplete_utf8(b"") == 0
    assert oracle.incomplete_utf8(b"abc") == 0
    assert oracle.incomplete_utf8("é".encode()) == 0
    assert oracle.incomplete_utf8("é".encode()[:1]) == 1
    assert oracle.incomplete_utf8("€".encode()[:1]) == 2
    assert oracle.incomplete_utf8("€".encode()[:2]) == 1
    assert oracle.incomplete_utf8(b"\x80\x80") == 2
    assert oracle.incomplete_utf8(b"a\xff") == 1
    assert oracle.incomplete_utf16(b"a") == 1
    assert oracle.incomplete_utf16(b"a\x00") == 0
    assert oracle.incomplete_utf16(b"\x00\xd8") == 2
    assert oracle.incomplete_utf16(b"\x00\xd8a") == 3
    assert oracle.incomplete_utf16(b"\x00\xdca\x00") == 2


@pytest.mark.parametrize("kind", KINDS)
def test_vocab_loader_rejects_bad_files(built, kind):
    good = gold_vocab("small_nocap")
    for bad in (good[:100], good + b"x", b"\x09" + good[1:]):
        with pytest.raises(oracle.OracleError):
            oracle.OracleVocab(bad, kind)
    # records out of (length, bytes) order: tokenmonster-cpp/src/tokenmonster.cpp:1352-1357
    swapped = vocabgen.raw_vocab([(b"b", 0, 0, 0, vocabgen.DNE, vocabgen.DNE), (b"a", 1, 0, 0, vocabgen.DNE, vocabgen.DNE)])
    with pytest.raises(oracle.OracleError) as e:
        oracle.OracleVocab(swapped, kind)
    assert e.value.status == 2


def test_non_ascii_under_nfd_is_a_normalization_error(built):
    v = oracle.OracleVocab(gold_vocab("small_cap2"), "port")
    with pytest.raises(oracle.OracleError) as e:
        v.tokenize("café".encode())
    assert e.value.status == 13

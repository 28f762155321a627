"""The product's table builder and per-string pipeline logic (csrc/vocab_tables.cpp +
csrc/pipeline_core.h -- the functions the sm_100a kernels call per lane), compiled with g++ by
tests/host_sim.cpp and compared with the oracle.  Runs without a GPU."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import oracle
from workloads import configs, corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SO = os.path.join(ROOT, "tests", "_build", "libhostsim.so")


@pytest.fixture(scope="module")
def sim():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "host_sim.cpp"), os.path.join(ROOT, "tokenmonster_b200", "csrc", "vocab_tables.cpp")]
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I/usr/local/cuda/include", *srcs, "-o", SO], check=True)
    lib = C.CDLL(SO)
    lib.hs_load.restype = C.c_void_p
    lib.hs_load.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int]
    lib.hs_free.argtypes = [C.c_void_p]
    lib.hs_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    lib.hs_dfa_lookup.restype = C.c_int
    lib.hs_dfa_lookup.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    lib.hs_normalize.restype = C.c_long
    lib.hs_normalize.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.hs_tokenize.restype = C.c_long
    lib.hs_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
    return lib


class Sim:
    def __init__(self, lib, vb):
        self.lib = lib
        err = C.create_string_buffer(256)
        self.h = lib.hs_load(vb, len(vb), err, 256)
        assert self.h, err.value

    def close(self):
        self.lib.hs_free(self.h)

    def lookup(self, w):
        idx, spi, spl = C.c_uint32(), C.c_uint32(), C.c_int()
        l = self.lib.hs_dfa_lookup(self.h, w, len(w), C.byref(idx), C.byref(spi), C.byref(spl))
        return (idx.value if l else None, l), (spi.value if spl.value else None, spl.value)

    def normalize(self, data):
        out = C.create_string_buffer(4 * len(data) + 16)
        n = self.lib.hs_normalize(self.h, data, len(data), out, len(out))
        return None if n < 0 else out.raw[:n]

    def tokenize(self, data, prenorm=False, count=False):
        cap = 2 * len(data) + 64
        ids = np.zeros(cap, np.uint32)
        miss = C.c_uint32()
        n = self.lib.hs_tokenize(self.h, data, len(data), int(prenorm), int(count), ids.ctypes.data, cap, C.byref(miss))
        if n < 0:
            return None, None
        assert n <= cap
        return (n if count else ids[:n].copy()), miss.value


def _vocab(name):
    if name in configs.CONFIGS:
        return configs.vocab_bytes(name)
    with open(os.path.join(GOLD, name + ".vocab"), "rb") as f:
        return f.read()


VOCABS = ["small_cap2", "small_nocap", "gpt2", "unicode_cap2", "c1", "c2", "c4"]


@pytest.mark.parametrize("name", VOCABS)
def test_reversed_automaton_equals_oracle_dictionary(sim, name):
    """One transition per byte, right to left, must name the longest key at the window start and
    the longest key behind the virtual space (go/tokenmonster.go:1049, :1088-1093)."""
    vb = _vocab(name)
    s, o = Sim(sim, vb), oracle.OracleVocab(vb, "port")
    text = corpus.base_text(3, 120_000, 0.5, pool=3000)
    if o.capcode == 2:
        text = o.normalize(text)
    rng = random.Random(11)
    m = o.max_token_length
    for _ in range(3000):
        a = rng.randrange(len(text) - 64)
        w = text[a:a + rng.randrange(1, 64)]
        (gi, gl), (si, sl) = s.lookup(w)
        assert (gi, gl) == o.longest_prefix(w[:m])
        wi, wl = o.longest_prefix((b" " + w)[:m])
        if wl > 1:
            assert (si, sl) == (wi, wl)
        else:
            assert sl <= 1
    s.close()


@pytest.mark.parametrize("name", ["small_cap2", "unicode_cap2", "c1", "c2"])
def test_capcode_encoder_equals_oracle(sim, name):
    vb = _vocab(name)
    s, o = Sim(sim, vb), oracle.OracleVocab(vb, "port")
    rng = random.Random(5)
    base = corpus.base_text(17, 300_000, 0.4)
    tricky = [b"", b"A", b"a", b"Ab", b"AbAbAb", b"A1B2C3", b"HTTP/2", b"CamelCase", b" MIXEDCaseABC", b"x1", b" HELLO ",
              b"don't DON'T Don'T", b"A'B'c", b"I'M 5'6\" TALL", b"ABC123def", b"ABC 123 DEF", b"  A  B  ", b"Z" * 70,
              b"aZ" * 40, b"A" * 33 + b"b", b"9A9a", b"'A'a'", b"O'Neil's", b"USA's GDP"]
    for t in tricky:
        assert s.normalize(t) == o.normalize(t), t
    for _ in range(600):
        a = rng.randrange(len(base) - 2000)
        t = bytes(base[a:a + rng.randrange(0, 900)])
        assert s.normalize(t) == o.normalize(t)
    alpha = b"aA bB'1 .Z"
    for _ in range(3000):
        t = bytes(rng.choice(alpha) for _ in range(rng.randrange(0, 24)))
        assert s.normalize(t) == o.normalize(t), t
    s.close()


@pytest.mark.parametrize("name", VOCABS + ["c5"])
def test_pipeline_logic_equals_oracle(sim, name):
    """encode -> position table -> chain, per string, against the oracle's Vocab.tokenize / Count."""
    vb = _vocab(name)
    s, o = Sim(sim, vb), oracle.OracleVocab(vb, "port")
    rng = random.Random(23)
    base = corpus.base_text(29, 400_000, 0.5 if name in ("gpt2", "c4") else 0.3)
    cases = [b"", b"a", b" ", b"A", b"hello", b"Hello World", b"x" * 100, b"The Quick BROWN fox's 123 jumps.\n\tint x = a[i]+1;"]
    for _ in range(250):
        a = rng.randrange(len(base) - 3000)
        cases.append(bytes(base[a:a + rng.randrange(1, 1500)]))
    alpha = b"aA bB'1 .Z\n_e t"
    for _ in range(400):
        cases.append(bytes(rng.choice(alpha) for _ in range(rng.randrange(1, 40))))
    for t in cases:
        want, wmiss = o.tokenize(t)
        got, gmiss = s.tokenize(t)
        assert got is not None and np.array_equal(got, want) and gmiss == wmiss, t
    for t in cases[:150]:
        wc, _ = o.tokenize(t, mode=1)
        gc, _ = s.tokenize(t, count=True)
        assert gc == wc, t
    s.close()


def test_pipeline_logic_on_normalized_non_ascii_text(sim):
    """The second pass reads text the device normalizer produced (bytes >= 0x80 included) verbatim."""
    from oracle import unicode_oracle as uo
    vb = _vocab("unicode_cap2")
    s, o = Sim(sim, vb), oracle.OracleVocab(vb, "port")
    texts = ["Le café de l'École ÉTAIT fermé — Москва, ΑΒγ.", "naïve façade ÀÉÎ straße 東京 x", "Ünïcödé ÇA va? ¿Qué tal?"]
    for t in texts:
        norm = uo.normalize(t.encode(), o.normflag, o.capcode)
        want, wmiss = o.tokenize(norm, normalized=True)
        got, gmiss = s.tokenize(norm, prenorm=True)
        assert np.array_equal(got, want) and gmiss == wmiss
    s.close()


def test_capcode_decode_scan_equals_oracle(sim):
    """The warp-scan capcode decoder (csrc/capcode_decode_scan.cuh: transition-monoid tables from
    gen_capdec_tables.py + the lane / round decomposition, simulated by tests/host_sim.cpp) against the
    oracle's sequential decoder: encoder output, adversarial marker soup, every start state (the
    streaming decoder resumes in any of them), every alignment, lengths across several rounds."""
    sim.hs_capdec_scan.restype = C.c_long
    sim.hs_capdec_scan.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint8), C.c_char_p]

    def run(data, galign, st):
        out = C.create_string_buffer(len(data) + 1)
        s = C.c_uint8(st)
        n = sim.hs_capdec_scan(data, len(data), galign, C.byref(s), out)
        return (None, None) if n < 0 else (out.raw[:n], s.value)

    rng = random.Random(77)
    base = corpus.base_text(41, 300_000, 0.4)
    cases = [b"", b"C", b"W", b"D", b"DC", b"DW hello", b"Wabc def", b"Chello", b"D Chello Wworld's 12ab, x", b"DDD   ", b"WCWCD a"]
    for _ in range(200):
        a = rng.randrange(len(base) - 3000)
        cases.append(oracle.capcode_encode(bytes(base[a:a + rng.randrange(0, 2500)])))
    alpha = b"CWD  aazZ9'.,\n"
    for _ in range(1500):
        cases.append(bytes(rng.choice(alpha) for _ in range(rng.randrange(0, 1200 if rng.random() < 0.1 else 40))))
    for t in cases:
        for st in (0, rng.randrange(16)):
            want, wst = oracle.capcode_decode(t, st)
            got, gst = run(t, rng.randrange(16), st)
            assert got == want and gst == wst, (t[:60], st)
    for st in range(16):   # every state x every byte, one step
        for b in range(128):
            want, wst = oracle.capcode_decode(bytes([b]), st)
            got, gst = run(bytes([b]), 0, st)
            assert got == want and gst == wst, (b, st)
    assert run(b"abc\xc3\xa9", 3, 0) == (None, None)   # non-ASCII text is left to the per-code-point kernel


def test_committed_capdec_tables_match_their_generator():
    """csrc/capdec_tables.inc (compiled into the decode kernel) is what gen_capdec_tables.py produces:
    44 monoid elements, composition / application / step tables entry by entry."""
    import importlib.util
    import re
    src = os.path.join(ROOT, "tokenmonster_b200", "csrc")
    spec = importlib.util.spec_from_file_location("gen_capdec_tables", os.path.join(src, "gen_capdec_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    n, comp, mul, apply_, stp = gen.tables()
    text = open(os.path.join(src, "capdec_tables.inc")).read()
    assert int(re.search(r"#define TM_CAPDEC_ELEMS (\d+)", text).group(1)) == n == 44
    for name, want in (("kCapdecComp", comp), ("kCapdecMul", mul), ("kCapdecApply", apply_), ("kCapdecStep", stp)):
        body = re.search(name + r"\[\d+\] = \{([^}]*)\}", text).group(1)
        got = [int(x) for x in body.replace("\n", " ").split(",") if x.strip()]
        assert got == want, name
    # the monoid really is closed and the identity is element 0
    assert all(mul[0 * 64 + b] == b and mul[b * 64 + 0] == b for b in range(n))
    for a in range(n):
        for b in range(n):
            ab = mul[a * 64 + b]
            assert all(apply_[ab * 16 + s] == apply_[b * 16 + apply_[a * 16 + s]] for s in range(16))

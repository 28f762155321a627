"""The C ABI without a GPU: the library loads, exports what include/tmgpu.h declares, parses
vocabularies, builds correct device tables, and refuses to compute without a device."""
import os
import random
import re

import numpy as np
import pytest

from oracle import oracle
from workloads import corpus, vocabgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol(built):
    from tokenmonster_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tmgpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tmgpu_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/tmgpu.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert b"sm_100a" in lib.tmgpu_version()


def test_header_is_plain_c(tmp_path):
    """cgo (and any C FFI) compiles include/tmgpu.h as C, not C++."""
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "tmgpu.h"\nint main(void) { tmgpu_vocab* v = 0; (void)v; return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "h.o")], check=True)


def test_host_only_vocab_and_no_cpu_path(built):
    from tokenmonster_b200 import TmgpuError, Vocab
    with open(os.path.join(GOLD, "gpt2.vocab"), "rb") as f:
        vb = f.read()
    v = Vocab(vb, device=-1)
    assert len(v) == 50238 and v.charset() == "UTF-8" and v.capcode() == 0 and v.max_token_length() == 40
    assert v.highest_token_id() == 50256 and v.unk_token_id() is None and v.default_encoding_length() == 2
    one = (np.frombuffer(b"hello", np.uint8), np.array([0, 5], np.uint64))
    for call in (lambda: v.tokenize("hello"), lambda: v.tokenize_count("hello"), lambda: v.decode([1, 2, 3]),
                 lambda: v.score_batch(*one), lambda: v.decoder(1)):
        with pytest.raises(TmgpuError) as e:
            call()
        assert e.value.code == 2 and "no CPU path" in str(e.value)


def test_wire_payload_validation_needs_no_gpu(built):
    """Truncated server payloads are refused before any device work (the server trusts its client,
    training/tokenmonsterserver.go:355-376; a library cannot)."""
    import struct
    from tokenmonster_b200 import TmgpuError, Vocab
    with open(os.path.join(GOLD, "gpt2.vocab"), "rb") as f:
        v = Vocab(f.read(), device=-1)
    good = struct.pack("<I", 1) + struct.pack("<Q", 5) + b"hello"
    for bad in (b"", b"\x01\x00", good[:9], good[:-1], struct.pack("<I", 2) + good[4:]):
        for call in (v.wire_tokenize, v.wire_count, lambda p: v.wire_decode(p, 2)):
            with pytest.raises(TmgpuError) as e:
                call(bad)
            assert e.value.code == 1
    with pytest.raises(TmgpuError) as e:
        v.wire_decode(good, 5)
    assert e.value.code == 1
    with pytest.raises(TmgpuError) as e:
        v.wire_tokenize(good)          # well-formed, but this vocabulary has no device
    assert e.value.code == 2


def test_vocab_error_codes(built):
    from tokenmonster_b200 import TmgpuError, Vocab
    with open(os.path.join(GOLD, "small_nocap.vocab"), "rb") as f:
        good = f.read()
    for bad in (good[:100], good + b"x", b"\x09" + good[1:]):
        with pytest.raises(TmgpuError) as e:
            Vocab(bad, device=-1)
        assert e.value.code == 14
    swapped = vocabgen.raw_vocab([(b"b", 0, 0, 0, vocabgen.DNE, vocabgen.DNE), (b"a", 1, 0, 0, vocabgen.DNE, vocabgen.DNE)])
    with pytest.raises(TmgpuError) as e:
        Vocab(swapped, device=-1)
    assert e.value.code == 15
    level1 = b"\x01" + good[1:]
    with pytest.raises(TmgpuError) as e:
        Vocab(level1, device=-1)
    assert e.value.code == 17
    with pytest.raises(TmgpuError) as e:
        Vocab.load("/nonexistent/x.vocab", device=-1)
    assert e.value.code == 12


@pytest.mark.parametrize("name", ["small_cap2", "small_nocap", "gpt2"])
def test_device_trie_equals_oracle_dictionary(built, name):
    """Walk the tables the kernels walk (host copy) against the oracle's longest-prefix search."""
    from tokenmonster_b200 import Vocab
    with open(os.path.join(GOLD, name + ".vocab"), "rb") as f:
        vb = f.read()
    v, o = Vocab(vb, device=-1), oracle.OracleVocab(vb, "port")
    text = corpus.base_text(3, 120_000, 0.5, pool=3000)
    if o.capcode == 2:
        text = o.normalize(text)
    rng = random.Random(9)
    m = o.max_token_length
    for _ in range(4000):
        a = rng.randrange(len(text) - 64)
        w = text[a:a + rng.randrange(1, m + 1)]
        assert v._trie_lookup(w) == o.longest_prefix(w) == v._trie_lookup(w, shortcut=True)
        wi, wl = o.longest_prefix((b" " + w)[:m])
        gi, gl = v._trie_lookup(w[:m - 1], spaced=True)
        assert (gi, gl) == v._trie_lookup(w[:m - 1], spaced=True, shortcut=True)
        if wl <= 1:
            assert gl == 0 or gl + 1 == wl
        else:
            assert (gi, gl + 1) == (wi, wl)


def _build_cpp_smoke(tmp_path):
    import subprocess
    exe = str(tmp_path / "cpp_smoke")
    so_dir = os.path.join(ROOT, "tokenmonster_b200")
    subprocess.run(["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp_smoke.cpp"), "-L" + so_dir, "-ltmgpu",
                    "-Wl,-rpath," + so_dir, "-o", exe], check=True)
    voc = str(tmp_path / "unit.vocab")
    with open(voc, "wb") as f:
        f.write(vocabgen.unit_cpp_vocab())
    return exe, voc


def test_cpp_mirror_header_compiles_and_refuses_without_gpu(built, tmp_path):
    """include/tokenmonster_gpu.hpp (mirror of tokenmonster-cpp's hpp:52-112) links against the C ABI."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu variant")
    exe, voc = _build_cpp_smoke(tmp_path)
    r = subprocess.run([exe, voc], capture_output=True, text=True)
    assert r.returncode == 0 and "no CPU path" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_runs_unit_cpp_assertions_on_gpu(built, tmp_path):
    import subprocess
    exe, voc = _build_cpp_smoke(tmp_path)
    r = subprocess.run([exe, voc, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


def test_deserialize_matches_reference_serializations(built):
    """Vocab.Deserialize (go/tokenmonster.go:749-784) of the reference's own 16/24/32-bit streams
    (tests/golden/cases.json, produced by oracle/_ref) gives back the reference's ids."""
    import json
    from tokenmonster_b200 import Vocab
    from tokenmonster_b200.vocab import TmgpuError
    with open(os.path.join(GOLD, "cases.json")) as f:
        cases = json.load(f)
    for name in ("small_cap2", "gpt2"):
        with open(os.path.join(GOLD, name + ".vocab"), "rb") as f:
            v = Vocab(f.read(), device=-1)
        for row in cases[name]["rows"][:60]:
            for enc in (2, 3, 4):
                if f"ser{enc}" in row:
                    got = v.deserialize_tokens(bytes.fromhex(row[f"ser{enc}"]), enc)
                    assert got.tolist() == row["ids"]
            assert v.deserialize_tokens(bytes.fromhex(row["ser2"]), 0).tolist() == row["ids"]   # both have <= 65536 ids
            assert v.deserialize_tokens(bytes.fromhex(row["ser4"]) + b"\x01\x02", 4).tolist() == row["ids"]   # ragged tail
        with pytest.raises(TmgpuError):
            v.deserialize_tokens(b"\x00\x00", 5)
        v.close()

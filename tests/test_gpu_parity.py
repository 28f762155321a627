"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the
committed reference fixtures.  Bit-exact bar (integer / byte / index work)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from workloads import configs, corpus, vocabgen

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
KERNELS = ["1"]   # one tokenize path since round 2 (the position-table pipeline)


def gold_vocab(name):
    if name == "unit_cpp":
        return vocabgen.unit_cpp_vocab()
    with open(os.path.join(GOLD, name + ".vocab"), "rb") as f:
        return f.read()


def make_vocab(vb, group="1", cta=None):
    from tokenmonster_b200 import Vocab
    return Vocab(vb, device=0)


def flat(texts):
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    return np.frombuffer(b"".join(texts), np.uint8), off


def oracle_tokenize(v, chk, raw):
    """(ids, missing, count, status).  ASCII: the C oracle end to end.  Non-ASCII under a normalizing
    UTF-8 vocabulary: oracle/unicode_oracle.py normalizes, the C oracle tokenizes the result."""
    from oracle import unicode_oracle as uo
    normalizes = v.charset() == "UTF-8" and (v.capcode() == 2 or (v.normalization_code() & 7))
    try:
        if normalizes and any(b >= 0x80 for b in raw):
            norm = uo.normalize(raw, v.normalization_code(), v.capcode())
            want, miss = chk.tokenize(norm, normalized=True)
            return want, miss, chk.tokenize(norm, normalized=True, mode=1)[0], 0
        want, miss = chk.tokenize(raw)
        return want, miss, chk.tokenize(raw, mode=1)[0], 0
    except (oracle.OracleError, uo.Refused):
        return np.zeros(0, np.uint32), 0, 0, 13


def check_batch(v, chk, data, off, enc):
    r = v.tokenize_batch(data, off, enc)
    counts, cmiss, cstat = v.count_batch(data, off)
    for s in range(len(off) - 1):
        raw = data[int(off[s]):int(off[s + 1])].tobytes()
        want, miss, wc, st = oracle_tokenize(v, chk, raw)
        assert np.array_equal(r[s].astype(np.uint32), want), f"string {s} ids differ"
        assert int(r.missing[s]) == miss and int(r.status[s]) == st
        assert int(counts[s]) == wc and int(cmiss[s]) == miss and int(cstat[s]) == st
    return r


@pytest.mark.parametrize("group", KERNELS)
@pytest.mark.parametrize("name", ["unit_cpp", "small_cap2", "small_nocap", "gpt2"])
def test_reference_golden_vectors(built, name, group):
    """Outputs of the reference's own C++ runtime (tests/golden/make_golden.py)."""
    with open(os.path.join(GOLD, "cases.json")) as f:
        rows = json.load(f)[name]["rows"]
    v = make_vocab(gold_vocab(name), group)
    data, off = flat([bytes.fromhex(r["in"]) for r in rows])
    encs = [4] if v.info.n_reverse > 65536 else [2, 3, 4]
    for enc in encs:
        r = v.tokenize_batch(data, off, enc)
        assert not r.status.any()
        for s, row in enumerate(rows):
            assert r[s].astype(np.uint32).tolist() == row["ids"]
            assert int(r.missing[s]) == row["missing"]
            ser = r.raw[int(r.tok_off[s]) * enc:int(r.tok_off[s + 1]) * enc].tobytes()
            assert ser.hex() == row[f"ser{enc}"]
        back, boff = v.decode_batch(r.raw, r.tok_off, enc)
        for s, row in enumerate(rows):
            assert back[int(boff[s]):int(boff[s + 1])].tobytes().hex() == row["decoded"]
    counts, _, _ = v.count_batch(data, off)
    assert counts.tolist() == [row["count"] for row in rows]
    v.close()


@pytest.mark.parametrize("group", KERNELS)
@pytest.mark.parametrize("name", ["c3", "c2", "c4", "c1", "c5"])
def test_seeded_batches_against_oracle(built, name, group):
    vb = configs.vocab_bytes(name)
    v, chk = make_vocab(vb, group), oracle.OracleVocab(vb, "port")
    base = corpus.base_text(31 + len(name), 1 << 19, configs.CONFIGS[name]["code"])
    data, off = corpus.make_batch(base, 1500, 0, 900, 5, "cpu")
    data, off = data.numpy(), off.numpy().astype(np.uint64)
    enc = 2 if configs.CONFIGS[name]["ids"] <= 65536 else 4
    r = check_batch(v, chk, data, off, enc)
    # decode round trip on the device, compared with the oracle's decode
    back, boff = v.decode_batch(r.raw, r.tok_off, enc)
    for s in range(0, len(off) - 1, 7):
        assert back[int(boff[s]):int(boff[s + 1])].tobytes() == chk.decode(r[s].astype(np.uint32))
    v.close()


@pytest.mark.parametrize("name", ["small_cap2", "c2"])
def test_capcode_text_that_outgrows_its_slot(built, name):
    """The table kernel gives every string room for 1.25 x + 48 normalized bytes; alternating case and
    digits grow 2.5-4 x and must come back through the full-size second run -- also when the string
    turns out to be non-ASCII only then, and when such strings sit among ordinary ones."""
    vb = configs.vocab_bytes(name) if name in configs.CONFIGS else gold_vocab(name)
    v, chk = make_vocab(vb), oracle.OracleVocab(vb, "port")
    base = bytes(corpus.base_text(77, 1 << 16, 0.3))
    texts = [b"A1B2C3D4" * 60, b"AbAbAb" * 200, base[:700], b"X" + b"1a" * 300, b"Q.R.S.T." * 90, base[700:1300],
             "é".encode() + b"A1" * 400, b"Z9" * 30, b"", b"A1B2" * 3000, base[2000:2100], b"I" * 500 + b"b"]
    for pad in (0, 5):
        data, off = flat([b"#" * pad] + texts)
        check_batch(v, chk, data, off, 2)
    v.close()


@pytest.mark.parametrize("group,cta", [("1", None)])
def test_edge_cases(built, group, cta):
    """Empty / 1-byte / unaligned / non-ASCII / look-ahead at the very end / offsets not starting at 0."""
    vb = gold_vocab("small_cap2")
    v, chk = make_vocab(vb, group, cta), oracle.OracleVocab(vb, "port")
    texts = [b"", b"a", b"A", b" ", b"", b"ab", b"Hello", b"x" * 3, "café".encode(), b"The End", b"\n", b"A" * 70 + b"b",
             b"HTTP" * 40, b"9" * 33, b"don't", b"", b"a" * 1000, b"\xffbad utf-8", b"ok after bad"]
    for pad in range(0, 17, 3):   # shift every string through all 16-byte alignments
        data, off = flat([b"#" * pad] + texts)
        check_batch(v, chk, data, off, 2)
    data, off = flat(texts)
    r = v.tokenize_batch(np.concatenate([np.zeros(5, np.uint8), data]), off + 5, 2)   # off[0] != 0
    assert np.array_equal(r.tokens, v.tokenize_batch(data, off, 2).tokens)
    assert r.status[8] == 0 and r.status[17] == 13 and int(r.tok_off[18] - r.tok_off[17]) == 0
    v.close()


def test_long_documents(built):
    """64 KiB documents (BASELINE configs[3] shape): tokenize + detokenize round trip is the identity."""
    vb = configs.vocab_bytes("c4")
    v, chk = make_vocab(vb, "1"), oracle.OracleVocab(vb, "port")
    base = corpus.base_text(404, 1 << 20, 0.9)
    data, off = corpus.make_batch(base, 24, 65536, 65536, 3, "cpu")
    data, off = data.numpy(), off.numpy().astype(np.uint64)
    r = v.tokenize_batch(data, off, 2)
    for s in (0, 11, 23):
        want, _ = chk.tokenize(data[int(off[s]):int(off[s + 1])].tobytes())
        assert np.array_equal(r[s].astype(np.uint32), want)
    back, boff = v.decode_batch(r.raw, r.tok_off, 2)
    assert np.array_equal(boff, off) and np.array_equal(back, data)   # NFD on ASCII + capcode 0 = identity
    v.close()


def test_slab_overflow_rerun(built):
    """More ids than bytes (+16): single-byte vocabulary under capcode 2, alternating case.
    The library notices the overflow and re-runs with exact sizes."""
    vb = vocabgen.generate_vocab([], capcode=2, charset=1, normflag=1, level=3, reserve=8)
    v, chk = make_vocab(vb, "1"), oracle.OracleVocab(vb, "port")
    texts = [b"aAbBcCdD" * 20, b"plain lower case text", b"xY" * 200, b""]
    data, off = flat(texts)
    r = check_batch(v, chk, data, off, 2)
    assert int(r.tok_off[1]) > len(texts[0]) + 16
    v.close()


def test_device_resident_path_matches_host_path(built):
    import torch
    from tokenmonster_b200.device import DeviceWorkspace
    vb = configs.vocab_bytes("c2")
    v = make_vocab(vb, "1")
    base = corpus.base_text(9, 1 << 19, 0.3)
    data, off = corpus.make_batch(base, 4000, 16, 700, 8, "cuda")
    ws = DeviceWorkspace(v, 4000, int(off[-1]))
    ws.launch_tokenize(data, off, 2, 0)
    res = ws.results()
    host = v.tokenize_batch(data.cpu().numpy(), off.cpu().numpy().astype(np.uint64), 2)
    assert res["n_tokens"] == host.n_tokens
    assert np.array_equal(res["tokens_u8"].cpu().numpy(), host.raw)
    assert np.array_equal(res["tok_off"].cpu().numpy().astype(np.uint64), host.tok_off)
    dbytes, doff = ws.decode(res["tokens_u8"], res["tok_off"], 2)
    back, boff = v.decode_batch(host.raw, host.tok_off, 2)
    assert np.array_equal(dbytes.cpu().numpy(), back) and np.array_equal(doff.cpu().numpy().astype(np.uint64), boff)
    assert len(ws.kernel_times_ms()) == 1 and ws.launches() >= 1
    ws.close()
    v.close()


def test_full_size_properties(built):
    """BASELINE configs[1] at full size (1,000,000 strings): size-independent properties --
    per-string id counts equal the Count kernel except where a retry emitted the delete token,
    decode(tokenize(x)) == x for ASCII text, and a strided sample equals the oracle."""
    import torch
    from tokenmonster_b200.device import DeviceWorkspace
    vb = configs.vocab_bytes("c2")
    v, chk = make_vocab(vb, "1"), oracle.OracleVocab(vb, "port")
    cfg = configs.CONFIGS["c2"]
    base = corpus.base_text(cfg["seed"], 16 << 20, cfg["code"])
    data, off = corpus.make_batch(base, 1_000_000, cfg["lo"], cfg["hi"], cfg["seed"] + 1, "cuda")
    ws = DeviceWorkspace(v, 1_000_000, int(off[-1]))
    ws.launch_tokenize(data, off, 2, 0)
    res = ws.results()
    assert int((res["status"] != 0).sum()) == 0
    toks, toff = res["tokens_u8"].clone(), res["tok_off"].clone()
    dbytes, doff = ws.decode(toks, toff, 2)
    assert torch.equal(doff, off) and torch.equal(dbytes, data)      # lossless: capcode + NFD round trip
    h_data, h_off = data.cpu().numpy(), off.cpu().numpy()
    ids = toks.cpu().numpy().view(np.uint16)
    t = toff.cpu().numpy()
    for s in range(0, 1_000_000, 4999):
        want, _ = chk.tokenize(h_data[int(h_off[s]):int(h_off[s + 1])].tobytes())
        assert np.array_equal(ids[int(t[s]):int(t[s + 1])].astype(np.uint32), want)
    ws.close()
    v.close()


@pytest.mark.parametrize("name", ["unit_cpp", "small_cap2", "small_nocap", "gpt2"])
def test_streaming_decoder_golden(built, name):
    """Decoder.Decode in three uneven pieces + Flush, against the reference's outputs (go:701-740)."""
    with open(os.path.join(GOLD, "cases.json")) as f:
        rows = json.load(f)[name]["rows"]
    v = make_vocab(gold_vocab(name), "1")
    d = v.decoder(len(rows))
    ids = [np.array(r["ids"], np.uint32) for r in rows]
    cuts = [(len(x) // 3, (2 * len(x)) // 3) for x in ids]
    got = [d.decode_batch([x[:a] for x, (a, b) in zip(ids, cuts)]),
           d.decode_batch([x[a:b] for x, (a, b) in zip(ids, cuts)]),
           d.decode_batch([x[b:] for x, (a, b) in zip(ids, cuts)]),
           d.flush()]
    for s, row in enumerate(rows):
        assert [g[s].hex() for g in got] == row["stream"], f"stream {s}"
    d.close()
    v.close()


def test_streaming_decoder_splits_utf8_and_capcode_state(built):
    """Ids that cut multi-byte characters (gpt2 byte-level pieces) and capcode markers split across
    calls, one id at a time, against the oracle's Decoder."""
    for name, texts in (("gpt2", ["naïve café — 日本語 😀 done", "€uro", "x"]),
                        ("small_cap2", [b"Hello WORLD and CamelCase HTTP/2 don't", b"THE END", b""])):
        vb = gold_vocab(name)
        v, chk = make_vocab(vb, "1"), oracle.OracleVocab(vb, "port")
        texts = [t.encode() if isinstance(t, str) else t for t in texts]
        ids = [chk.tokenize(t)[0] for t in texts]
        d, od = v.decoder(len(texts)), [chk.new_decoder() for _ in texts]
        steps = max(len(x) for x in ids)
        for k in range(steps):
            chunk = [x[k:k + 1] for x in ids]
            got = d.decode_batch(chunk)
            want = [o.decode(c) for o, c in zip(od, chunk)]
            assert got == want, f"{name} step {k}"
        assert d.flush() == [o.flush() for o in od]
        d.close()
        v.close()


# ---- server wire format (training/tokenmonsterserver.go jobs 1, 2-4; python client job 20) ----------
def wire_request(items):
    import struct
    return struct.pack("<I", len(items)) + b"".join(struct.pack("<Q", len(x)) + x for x in items)


def wire_entries(reply):
    import struct
    n = struct.unpack_from("<I", reply, 0)[0]
    at, out = 4, []
    for _ in range(n):
        l = struct.unpack_from("<Q", reply, at)[0]
        at += 8
        out.append(reply[at:at + l])
        at += l
    assert at == len(reply)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2", "c3", "gpt2"])
def test_wire_tokenize_count_decode_against_oracle(built, name):
    import struct
    vb = gold_vocab(name) if name == "gpt2" else configs.vocab_bytes(name)
    v, chk = make_vocab(vb, "1"), oracle.OracleVocab(vb, "port")
    base = corpus.base_text(21, 1 << 18, 0.4)
    d, o = corpus.make_batch(base, 700, 0, 500, 4, "cpu")
    d, o = d.numpy(), o.numpy()
    texts = [d[int(o[s]):int(o[s + 1])].tobytes() for s in range(len(o) - 1)] + [b"", b"x", b""]
    enc = 4 if len(v) > 65536 else 2
    st, reply = v.wire_tokenize(wire_request(texts))
    got = wire_entries(reply)
    assert st == 0 and len(got) == len(texts)
    dt = np.uint16 if enc == 2 else np.uint32
    for s, t in enumerate(texts):
        want, _ = chk.tokenize(t)
        assert np.array_equal(np.frombuffer(got[s], dt).astype(np.uint32), want), f"entry {s}"
    st, creply = v.wire_count(wire_request(texts))
    assert st == 0 and struct.unpack_from("<I", creply, 0)[0] == len(texts) and len(creply) == 4 + 8 * len(texts)
    for s, t in enumerate(texts):
        assert struct.unpack_from("<Q", creply, 4 + 8 * s)[0] == chk.tokenize(t, mode=1)[0]
    # decode: the reply of job 1 is, entry for entry, a valid request of job `enc`
    st, dreply = v.wire_decode(reply, enc)
    back = wire_entries(dreply)
    for s, t in enumerate(texts):
        want, _ = chk.tokenize(t)
        assert back[s] == chk.decode(want), f"decode entry {s}"
    # 3-byte ids and a dangling byte (whole ids only, go:845)
    ids = [chk.tokenize(t)[0] for t in texts[:50]]
    e3 = [b"".join(int(i).to_bytes(3, "little") for i in a) + b"\x07" for a in ids]
    st, d3 = v.wire_decode(wire_request(e3), 3)
    assert [chk.decode(a) for a in ids] == wire_entries(d3)
    v.close()


@pytest.mark.gpu
def test_wire_tokenize_streamed_batch_equals_flat_batch(built):
    """A payload big enough for the chunked upload + watermark path (>16 MB, >118k strings)."""
    vb = configs.vocab_bytes("c2")
    v = make_vocab(vb, "1")
    base = corpus.base_text(33, 1 << 20, 0.4)
    d, o = corpus.make_batch(base, 160000, 0, 300, 11, "cpu")
    d, o = d.numpy(), o.numpy().astype(np.uint64)
    flat_res = v.tokenize_batch(d, o, 2)
    lens = np.diff(o).astype("<u8")
    # build the request with numpy: [u32 n] then [u64 len][bytes] per string
    pos = 4 + 8 * np.arange(1, len(lens) + 1, dtype=np.uint64) + o[:-1]
    req = np.zeros(4 + 8 * len(lens) + len(d), np.uint8)
    req[:4] = np.frombuffer(np.uint32(len(lens)).tobytes(), np.uint8)
    hdr = (pos - 8)[:, None] + np.arange(8, dtype=np.uint64)[None, :]
    req[hdr.reshape(-1)] = lens.view(np.uint8)
    idx = np.repeat(pos - o[:-1], lens.astype(np.int64)) + np.arange(len(d), dtype=np.uint64)
    req[idx] = d
    st, reply = v.wire_tokenize(req)
    assert st == 0
    rep = np.frombuffer(reply, np.uint8)
    tl = np.diff(flat_res.tok_off).astype(np.uint64) * 2
    assert len(rep) == 4 + 8 * len(tl) + int(tl.sum())
    tpos = 4 + 8 * np.arange(1, len(tl) + 1, dtype=np.uint64) + np.concatenate([[0], np.cumsum(tl)[:-1]]).astype(np.uint64)
    thdr = (tpos - 8)[:, None] + np.arange(8, dtype=np.uint64)[None, :]
    assert np.array_equal(rep[thdr.reshape(-1)], tl.astype("<u8").view(np.uint8))
    tidx = np.repeat(tpos, tl.astype(np.int64)) + (np.arange(int(tl.sum()), dtype=np.uint64) -
                                                     np.repeat(np.cumsum(tl) - tl, tl.astype(np.int64)))
    assert np.array_equal(rep[tidx], flat_res.raw)
    v.close()


@pytest.mark.gpu
def test_wire_status_code_follows_entry_zero(built):
    vb = configs.vocab_bytes("c2")
    v = make_vocab(vb, "1")
    st, reply = v.wire_tokenize(wire_request([b"caf\xe9", b"plain"]))       # invalid UTF-8
    got = wire_entries(reply)
    assert st == 13 and got[0] == b"" and len(got[1]) > 0          # server:379-381
    st, reply = v.wire_tokenize(wire_request([b"plain", b"caf\xe9"]))
    assert st == 0 and wire_entries(reply)[1] == b""
    st, reply = v.wire_tokenize(wire_request([]))
    assert st == 0 and reply == b"\x00\x00\x00\x00"
    v.close()


def test_poisoned_workspace_decode_twice(built):
    """Every device allocation pre-filled with 0xA5 (TMGPU_POISON): tokenize, decode twice and the
    streaming decoder must still equal the oracle -- no kernel may depend on memory nobody wrote
    (one of three driver runs of smoke() decoded wrong bytes)."""
    import subprocess
    import sys
    env = dict(os.environ, TMGPU_POISON="1")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "gpu_poison_check.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    assert "__POISON_OK__" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_host_path_streams_ids_out_in_chunks(built):
    """A batch big enough for the chunked host path (several chunks of 1280 strings per SM): the ids leave the GPU
    chunk by chunk.  First call sizes the pinned buffer by guess, the second by the first's yield;
    a non-ASCII string forces the fallback (one copy at the end).  All three must equal the
    device-resident single launch and the oracle."""
    import torch
    from tokenmonster_b200.device import DeviceWorkspace
    vb = configs.vocab_bytes("c2")
    v, chk = make_vocab(vb), oracle.OracleVocab(vb, "port")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n = 4 * sms * 1024 + 777
    base = corpus.base_text(41, 1 << 20, 0.3)
    data, off = corpus.make_batch(base, n, 0, 160, 9, "cpu")
    data, off = data.numpy(), off.numpy().astype(np.uint64)
    ws = DeviceWorkspace(v, n, int(off[-1]))
    ws.launch_tokenize(torch.from_numpy(data).cuda(), torch.from_numpy(off.astype(np.int64)).cuda(), 2, 0)
    ref = ws.results()
    want_tok = ref["tokens_u8"].cpu().numpy().view(np.uint16)
    want_off = ref["tok_off"].cpu().numpy().astype(np.uint64)
    for rep in range(2):
        r = v.tokenize_batch(data, off, 2)
        assert np.array_equal(r.tok_off, want_off) and np.array_equal(r.tokens, want_tok), f"call {rep}"
        assert not r.status.any()
    for s in range(0, n, 4099):
        want, miss = chk.tokenize(data[int(off[s]):int(off[s + 1])].tobytes())
        assert np.array_equal(r[s].astype(np.uint32), want) and int(r.missing[s]) == miss
    # one non-ASCII string in the middle: its ids change the offsets of everything behind it
    mid = n // 2
    extra = "Æther café".encode()
    d2 = np.concatenate([data[:int(off[mid])], np.frombuffer(extra, np.uint8), data[int(off[mid]):]])
    o2 = np.concatenate([off[:mid + 1], off[mid:] + len(extra)])
    r2 = v.tokenize_batch(d2, o2, 2)
    assert v.last_stats()["unicode_strings"] == 1
    assert np.array_equal(r2.tokens[:int(r2.tok_off[mid])], want_tok[:int(want_off[mid])])
    k = int(r2.tok_off[mid + 1])
    assert np.array_equal(r2.tokens[k:], want_tok[int(want_off[mid]):])
    ws.close()
    v.close()


def test_two_devices_in_one_process(built):
    """What a cgo server with several GPUs does: vocabularies on device 0 and device 1 in ONE process,
    used alternately (kernel attributes / occupancy are per device, nothing may be cached per process)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tokenmonster_b200 import Vocab
    vb = configs.vocab_bytes("c2")
    chk = oracle.OracleVocab(vb, "port")
    base = corpus.base_text(55, 1 << 18, 0.3)
    data, off = corpus.make_batch(base, 3000, 0, 700, 4, "cpu")
    data, off = data.numpy(), off.numpy().astype(np.uint64)
    v0, v1 = Vocab(vb, device=0), Vocab(vb, device=1)
    for rep in range(2):
        r0, r1 = v0.tokenize_batch(data, off, 2), v1.tokenize_batch(data, off, 2)
        assert np.array_equal(r0.tokens, r1.tokens) and np.array_equal(r0.tok_off, r1.tok_off)
        b1, o1 = v1.decode_batch(r1.raw, r1.tok_off, 2)
        b0, o0 = v0.decode_batch(r0.raw, r0.tok_off, 2)
        assert np.array_equal(b0, b1) and np.array_equal(o0, o1)
    for s in range(0, 3000, 101):
        want, _ = chk.tokenize(data[int(off[s]):int(off[s + 1])].tobytes())
        assert np.array_equal(r1[s].astype(np.uint32), want)
    long_doc = bytes(corpus.base_text(6, 60000, 0.3))     # the chunk kernel needs its shared-memory opt-in on both
    o2 = np.array([0, len(long_doc)], np.uint64)
    a = v0.tokenize_batch(np.frombuffer(long_doc, np.uint8), o2, 4)
    b = v1.tokenize_batch(np.frombuffer(long_doc, np.uint8), o2, 4)
    assert np.array_equal(a.tokens, b.tokens) and v1.last_stats()["chunks"] > 10
    v0.close()
    v1.close()

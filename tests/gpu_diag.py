"""Diagnostic matrix for the GPU path (not a pytest file): python tests/gpu_diag.py [quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from workloads import configs, corpus, vocabgen  # noqa: E402


def run(name, vb, data, off, group, enc=4, max_report=3):
    os.environ["TMGPU_GROUP"] = str(group)
    from tokenmonster_b200 import Vocab
    v = Vocab(vb, device=0)
    chk = oracle.OracleVocab(vb, "port")
    t = time.time()
    r = v.tokenize_batch(data, off, enc)
    dt = time.time() - t
    bad = 0
    n = len(off) - 1
    for s in range(n):
        raw = data[int(off[s]):int(off[s + 1])].tobytes()
        try:
            want, miss = chk.tokenize(raw)
            st = 0
        except oracle.OracleError as e:
            want, miss, st = np.zeros(0, np.uint32), 0, e.status
        got = r[s].astype(np.uint32)
        if not np.array_equal(got, want) or int(r.missing[s]) != miss or int(r.status[s]) != st:
            bad += 1
            if bad <= max_report:
                k = 0
                while k < min(len(got), len(want)) and got[k] == want[k]:
                    k += 1
                print(f"  MISMATCH {name} G={group} s={s} len={len(raw)} ntok got {len(got)} want {len(want)} first diff @{k} "
                      f"miss {int(r.missing[s])}/{miss} status {int(r.status[s])}/{st}")
                print("   got ", got[max(0, k - 3):k + 5], " want", want[max(0, k - 3):k + 5])
                pre = chk.decode(want[:k], raw=True) if k else b""
                print("   raw", raw[:80], "... prefix-decoded tail", pre[-40:])
    # decode round trip through the GPU
    back, boff = v.decode_batch(r.tokens, r.tok_off, 4 if enc != 2 else 2) if enc != 3 else v.decode_batch(r.tokens, r.tok_off, 3)
    dbad = 0
    for s in range(n):
        want = chk.decode(r[s].astype(np.uint32))
        if back[int(boff[s]):int(boff[s + 1])].tobytes() != want:
            dbad += 1
            if dbad <= 2:
                print(f"  DECODE MISMATCH {name} s={s}", back[int(boff[s]):int(boff[s + 1])].tobytes()[:60], want[:60])
    c, m, stt = v.count_batch(data, off)
    cbad = 0
    for s in range(0, n, max(1, n // 200)):
        raw = data[int(off[s]):int(off[s + 1])].tobytes()
        try:
            wc, wm = chk.tokenize(raw, mode=1)
        except oracle.OracleError:
            wc, wm = 0, 0
        if int(c[s]) != wc or int(m[s]) != wm:
            cbad += 1
    print(f"{name:28s} G={group:2d} enc={enc} n={n} ids={int(r.tok_off[-1])} tok_bad={bad} dec_bad={dbad} cnt_bad={cbad} "
          f"e2e {dt * 1e3:.1f} ms")
    v.close()
    return bad + dbad + cbad


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    total = 0
    # unit.cpp vocabulary
    vb = vocabgen.unit_cpp_vocab()
    texts = [b"ab a z", b"", b"z", b"ab", b"abababab a", b"  ", b"bbbb"]
    data = np.frombuffer(b"".join(texts), np.uint8)
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    groups = [int(x) for x in os.environ.get('DIAG_GROUPS', '1,8,32').split(',')]
    for g in groups:
        total += run("unit.cpp", vb, data, off, g)
    for name in ("c3", "c4", "c2", "c1"):
        vb = configs.vocab_bytes(name)
        base = corpus.base_text(4242, 1 << 20, configs.CONFIGS[name]["code"])
        n = 300 if quick else 3000
        d, o = corpus.make_batch(base, n, 0, 900, 11, "cpu")
        d, o = d.numpy(), o.numpy().astype(np.uint64)
        for g in groups:
            total += run(name, vb, d, o, g, enc=4 if g != 8 else 2 if configs.CONFIGS[name]["ids"] <= 65536 else 3)
    print("TOTAL BAD", total)
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())

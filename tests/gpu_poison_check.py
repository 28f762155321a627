"""Run with TMGPU_POISON=1 (every device allocation pre-filled with 0xA5): tokenize -> decode twice
-> streaming decoder on capcode-2 and capcode-0 batches, everything compared with the oracle.
A kernel that reads memory nobody wrote fails here on every run, not once in three."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from tokenmonster_b200 import Vocab  # noqa: E402
from workloads import configs, corpus  # noqa: E402


def main():
    assert os.environ.get("TMGPU_POISON"), "set TMGPU_POISON=1"
    for name in ("c3", "c2", "c4"):
        vb = configs.vocab_bytes(name)
        chk = oracle.OracleVocab(vb, "port")
        base = corpus.base_text(99, 1 << 18, 0.4)
        for rep, (n, lo, hi) in enumerate(((256, 16, 700), (900, 0, 300), (64, 2000, 6000))):
            v = Vocab(vb, device=0)   # a fresh (poisoned) workspace every time
            data, off = corpus.make_batch(base, n, lo, hi, 7 + rep, "cpu")
            data, off = data.numpy(), off.numpy().astype(np.uint64)
            r = v.tokenize_batch(data, off, 4)
            for s in range(n):
                want, miss = chk.tokenize(data[int(off[s]):int(off[s + 1])].tobytes())
                assert np.array_equal(r[s], want), f"{name}/{rep}: string {s} differs"
                assert int(r.missing[s]) == miss
            want_dec = [chk.decode(r[s]) for s in range(n)]
            for again in range(2):
                back, boff = v.decode_batch(r.tokens, r.tok_off, 4)
                for s in range(n):
                    assert back[int(boff[s]):int(boff[s + 1])].tobytes() == want_dec[s], f"{name}/{rep}: decode {s} (pass {again})"
            dec = v.decoder(n)
            od = [chk.new_decoder() for _ in range(n)]
            half = [r[s][: len(r[s]) // 2] for s in range(n)], [r[s][len(r[s]) // 2:] for s in range(n)]
            for part in half:
                got = dec.decode_batch(part)
                for s in range(n):
                    assert got[s] == od[s].decode(part[s]), f"{name}/{rep}: stream {s}"
            fl = dec.flush()
            for s in range(n):
                assert fl[s] == od[s].flush()
            dec.close()
            v.close()
        print(f"poison check {name}: ok")
    print("__POISON_OK__")


if __name__ == "__main__":
    main()

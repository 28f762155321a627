"""Small driver for compute-sanitizer (memcheck / initcheck / racecheck / synccheck): every kernel
once on tiny batches -- tokenize (both text modes), count, scores, long documents in chunks,
non-ASCII pass, decode, streaming decoder, wire framing.

    compute-sanitizer --tool initcheck python tests/gpu_sanitize.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from workloads import configs, corpus  # noqa: E402


def main():
    from tokenmonster_b200 import Vocab
    for name in ("c2", "c4"):
        vb = configs.vocab_bytes(name)
        base = corpus.base_text(5, 1 << 17, 0.5)
        d, o = corpus.make_batch(base, 300, 0, 400, 2, "cpu")
        d, o = d.numpy(), o.numpy().astype(np.uint64)
        v = Vocab(vb, device=0)
        r = v.tokenize_batch(d, o, 2)
        c, _, _ = v.count_batch(d, o)
        sc = v.score_batch(d, o)
        back, boff = v.decode_batch(r.raw, r.tok_off, 2)
        back2, _ = v.decode_batch(r.raw, r.tok_off, 2)
        assert np.array_equal(back, back2)
        dec = v.decoder(len(o) - 1)
        dec.decode_batch([r[s] for s in range(len(o) - 1)])
        dec.flush()
        # side paths: one long document (chunks) and one non-ASCII string (device normalizer)
        texts = [bytes(corpus.base_text(6, 30000, 0.3)), "Le café ÉTAIT fermé — Москва".encode()]
        o2 = np.array([0, len(texts[0]), len(texts[0]) + len(texts[1])], np.uint64)
        r2 = v.tokenize_batch(np.frombuffer(b"".join(texts), np.uint8), o2, 4)
        print(name, int(r.tok_off[-1]), int(c.sum()), len(back), int(sc[2]), int(r2.tok_off[-1]), v.last_stats())
        dec.close()
        v.close()
    print("__SANITIZE_DONE__")


if __name__ == "__main__":
    main()

"""Small driver for compute-sanitizer (memcheck / racecheck): every kernel once on tiny batches."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from workloads import configs, corpus, vocabgen  # noqa: E402


def main():
    groups = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "8"]
    for name in ("c2", "c4"):
        vb = configs.vocab_bytes(name)
        base = corpus.base_text(5, 1 << 17, 0.5)
        d, o = corpus.make_batch(base, 300, 0, 400, 2, "cpu")
        d, o = d.numpy(), o.numpy().astype(np.uint64)
        for g in groups:
            os.environ["TMGPU_GROUP"] = g
            from tokenmonster_b200 import Vocab
            v = Vocab(vb, device=0)
            r = v.tokenize_batch(d, o, 2)
            c, _, _ = v.count_batch(d, o)
            back, boff = v.decode_batch(r.raw, r.tok_off, 2)
            dec = v.decoder(len(o) - 1)
            dec.decode_batch([r[s] for s in range(len(o) - 1)])
            dec.flush()
            print(name, g, int(r.tok_off[-1]), int(c.sum()), len(back))
            dec.close()
            v.close()


if __name__ == "__main__":
    main()

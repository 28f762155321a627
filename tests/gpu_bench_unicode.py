"""Cost of the non-ASCII pass (diagnostic): config c2's vocabulary, 300k strings, a share of which
carry accented words / other scripts.  Prints device time per launch for each share."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from workloads import configs, corpus  # noqa: E402
from tokenmonster_b200 import Vocab  # noqa: E402
from tokenmonster_b200.device import DeviceWorkspace  # noqa: E402
import test_unicode as tu  # noqa: E402


def main():
    n = 300000
    v = Vocab(configs.vocab_bytes("c2"), device=0)
    base = corpus.base_text(5, 1 << 22, 0.3)
    d, o = corpus.make_batch(base, n, 64, 960, 2, "cpu")
    d, o = d.numpy(), o.numpy()
    asc = [d[int(o[s]):int(o[s + 1])].tobytes() for s in range(n)]
    uni = tu.unicode_texts(3, 4000, 64, 960)
    for share in (0.0, 0.01, 0.1, 0.5, 1.0):
        rng = np.random.default_rng(1)
        pick = rng.random(n) < share
        texts = [uni[s % len(uni)] if pick[s] else asc[s] for s in range(n)]
        off = np.zeros(n + 1, np.int64)
        off[1:] = np.cumsum([len(t) for t in texts])
        data = torch.from_numpy(np.frombuffer(b"".join(texts), np.uint8).copy()).cuda()
        doff = torch.from_numpy(off).cuda()
        ws = DeviceWorkspace(v, n, int(off[-1]))
        for _ in range(2):
            ws.launch_tokenize(data, doff, 2, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 4
        for _ in range(reps):
            ws.launch_tokenize(data, doff, 2, 0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        res = ws.results()
        bad = int((res["status"] != 0).sum())
        print(f"non-ASCII share {share:4.2f}: {int(off[-1]) / 1e6:6.1f} MB in {dt * 1e3:6.2f} ms = "
              f"{int(off[-1]) / dt / 1e9:5.1f} GB/s   ids {res['n_tokens'] / 1e6:.1f} M   refused {bad}", flush=True)
        ws.close()
    v.close()


if __name__ == "__main__":
    main()

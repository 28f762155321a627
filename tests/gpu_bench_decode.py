"""Device-resident decode throughput on configuration c2 (diagnostic, not the bench line)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from workloads import configs, corpus  # noqa: E402
from tokenmonster_b200 import Vocab  # noqa: E402
from tokenmonster_b200.device import DeviceWorkspace  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    for name in ("c2", "c4"):
        v = Vocab(configs.vocab_bytes(name), device=0)
        base = corpus.base_text(5, 1 << 22, 0.3)
        data, off = corpus.make_batch(base, n, 64, 960, 2, "cuda")
        ws = DeviceWorkspace(v, n, int(off[-1]))
        ws.launch_tokenize(data, off, 2, 0)
        res = ws.results()
        toks, toff = res["tokens_u8"].clone(), res["tok_off"].clone()
        for _ in range(2):
            out, ooff = ws.decode(toks, toff, 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            out, ooff = ws.decode(toks, toff, 2)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"{name}: decode {out.numel() / 1e6:.1f} MB from {res['n_tokens'] / 1e6:.1f} M ids in {dt * 1e3:.2f} ms"
              f" = {out.numel() / dt / 1e9:.1f} GB/s of text (launches {ws.launches()})", flush=True)
        assert out.numel() == data.numel() and bool((out == data).all())
        ws.close()
        v.close()


if __name__ == "__main__":
    main()

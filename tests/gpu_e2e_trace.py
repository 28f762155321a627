"""Host-path phase times (TMGPU_TRACE=1): pinned and pageable input, config c2 shape."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TMGPU_TRACE"] = "1"
from tokenmonster_b200 import Vocab  # noqa: E402
from workloads import configs, corpus  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    cfg = configs.CONFIGS[name]
    v = Vocab(configs.vocab_bytes(name), device=0)
    base = corpus.base_text(cfg["seed"], 16 << 20, cfg["code"])
    data, off = corpus.make_batch(base, n, cfg["lo"], cfg["hi"], cfg["seed"] + 1, "cuda")
    total = int(off[-1].item())
    pinned = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(data)
    h_off = off.cpu().numpy().astype(np.uint64)
    pageable = np.empty(total, np.uint8)
    pageable[:] = pinned.numpy()
    enc = 2 if cfg["ids"] <= 65536 else 4
    for label, arr in (("pinned", pinned.numpy()), ("pageable", pageable)):
        r = None
        for rep in range(4):
            r = None
            t = time.perf_counter()
            r = v.tokenize_batch(arr, h_off, enc)
            dt = time.perf_counter() - t
            print(f"{label} call {rep}: {dt * 1e3:.2f} ms = {total / dt / 1e9:.1f} GB/s, ids {int(r.tok_off[-1])}", flush=True)
    v.close()


if __name__ == "__main__":
    main()

"""Profiling driver (run under ncu): the device-resident tokenize path on one configuration.
    python tests/gpu_profile_pipeline.py [config] [strings] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tokenmonster_b200 import Vocab  # noqa: E402
from tokenmonster_b200.device import DeviceWorkspace  # noqa: E402
from workloads import configs, corpus  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    cfg = configs.CONFIGS[name]
    dev = torch.device("cuda", 0)
    vocab = Vocab(configs.vocab_bytes(name), device=0)
    base = corpus.base_text(cfg["seed"], 16 << 20, cfg["code"])
    data, off = corpus.make_batch(base, n, cfg["lo"], cfg["hi"], cfg["seed"] + 1, dev)
    total = int(off[-1].item())
    ws = DeviceWorkspace(vocab, n, total)
    enc = 2 if cfg["ids"] <= 65536 else 4
    for _ in range(reps):
        ws.launch_tokenize(data, off, enc, 0)
    torch.cuda.synchronize()
    t = ws.kernel_times_ms()
    print(name, n, total, "bytes; pipeline ms:", t, "stages ms:", ws.stage_times_ms(), "ids:", ws.results()["n_tokens"])


if __name__ == "__main__":
    main()

"""Few-but-huge inputs (diagnostic): config c1's 1 MB document and a handful of large documents,
with and without chunking (TMGPU_NO_CHUNKS=1 in a child process), host-buffer C ABI call."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from workloads import configs, corpus  # noqa: E402


def run():
    from tokenmonster_b200 import Vocab
    from oracle import oracle
    for name, ndocs, size in (("c1", 1, 1 << 20), ("c2", 1, 1 << 20), ("c2", 16, 1 << 20), ("c4", 64, 1 << 16)):
        vb = configs.vocab_bytes(name)
        v = Vocab(vb, device=0)
        texts = [bytes(corpus.base_text(100 + k, size, 0.3)) for k in range(ndocs)]
        off = np.zeros(ndocs + 1, np.uint64)
        off[1:] = np.cumsum([len(t) for t in texts])
        data = np.frombuffer(b"".join(texts), np.uint8)
        for _ in range(2):
            r = v.tokenize_batch(data, off, 2)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            r = v.tokenize_batch(data, off, 2)
        dt = (time.perf_counter() - t0) / reps
        st = v.last_stats()
        chk = oracle.OracleVocab(vb, "reference" if oracle.have_reference() else "port")
        t1 = time.perf_counter()
        want, _ = chk.tokenize(texts[0])
        cpu = time.perf_counter() - t1
        ok = np.array_equal(r[0].astype(np.uint32), want)
        print(f"{name} {ndocs} x {size >> 10} KB: {dt * 1e3:8.2f} ms = {len(data) / dt / 1e6:8.1f} MB/s  chunks {st['chunks']} redo "
              f"{st['chunk_redo']}  | CPU 1 thread, one document: {cpu * 1e3:.1f} ms = {size / cpu / 1e6:.1f} MB/s  exact={ok}",
              flush=True)
        v.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for env in ({}, {"TMGPU_NO_CHUNKS": "1"}):
            print("chunking", "off" if env else "on", flush=True)
            subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env}, check=True)

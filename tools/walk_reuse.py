"""How many dictionary searches the ungreedy loop performs per emitted id, and how many of them
repeat a position it searched in one of the previous steps (oracle instrumentation, CPU only).

    python tools/walk_reuse.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from workloads import configs, corpus  # noqa: E402


def main():
    lib = C.CDLL(oracle.PORT_SO)
    out = (C.c_uint64 * 4)()
    for name in ("c2", "c3", "c4"):
        o = oracle.OracleVocab(configs.vocab_bytes(name), "port")
        text = bytes(corpus.base_text(5, 1 << 20, 0.3 if name != "c4" else 0.9))
        lib.tmo_debug_walk_stats(out, 1)
        ids, _ = o.tokenize(text)
        lib.tmo_debug_walk_stats(out, 1)
        plain, retry, rep, steps = (int(x) for x in out)
        n = len(ids)
        print(f"{name}: {n} ids, {steps} steps; per id: {plain / n:.2f} plain searches (+1 for the next current token "
              f"when no branch is taken), {retry / n:.2f} retry searches; {100 * rep / max(1, plain):.1f}% of the plain "
              f"searches repeat one of the previous 8 positions")


if __name__ == "__main__":
    main()

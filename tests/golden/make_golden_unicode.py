#!/usr/bin/env python3
"""Generates tests/golden/unicode_cases.json + unicode_cap2.vocab: non-ASCII inputs with what the
ORACLE (oracle/unicode_oracle.py on CPython's unicodedata 15.0 + oracle/tm_oracle.c) says about
them -- normalized text, ids, decoded text, training scores.  These are not reference outputs (the
reference's normalizer cannot run here: parity for non-ASCII is unpinned); they freeze
the oracle's answer so that a different interpreter / Unicode version cannot move the target.

    python tests/golden/make_golden_unicode.py
"""
import json
import os
import sys
import unicodedata

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle  # noqa: E402
from oracle import unicode_oracle as uo  # noqa: E402
import test_unicode as tu  # noqa: E402


def main():
    assert unicodedata.unidata_version == "15.0.0", unicodedata.unidata_version
    vb = tu.make_unicode_vocab(2, 1)
    with open(os.path.join(HERE, "unicode_cap2.vocab"), "wb") as f:
        f.write(vb)
    chk = oracle.OracleVocab(vb, "port")
    texts = [s.encode() for s in tu.SAMPLES] + tu.EDGE + tu.unicode_texts(2024, 40, 1, 120)
    rows = []
    for t in texts:
        row = {"in": t.hex()}
        try:
            norm = uo.normalize(t, 1, 2)
        except uo.Refused:
            row["status"] = 13
            rows.append(row)
            continue
        ids, miss = chk.tokenize(norm, normalized=True)
        keys = chk.decode(ids, raw=True)
        scores, hist, tit = chk.score([norm], normalized=True)
        nz = {int(k): int(scores[k]) for k in scores.nonzero()[0]}
        row.update(status=0, norm=norm.hex(), ids=[int(x) for x in ids], missing=int(miss),
                   count=int(chk.tokenize(norm, normalized=True, mode=1)[0]),
                   decoded=uo.capcode_decode(keys.decode("utf-8")).encode().hex(),
                   tokens_in_text=int(tit), scores=nz)
        rows.append(row)
    with open(os.path.join(HERE, "unicode_cases.json"), "w") as f:
        json.dump({"unicode": unicodedata.unidata_version, "capcode": 2, "normflag": 1, "rows": rows}, f, indent=0)
    print(len(rows), "rows", sum(r["status"] == 0 for r in rows), "accepted")


if __name__ == "__main__":
    main()

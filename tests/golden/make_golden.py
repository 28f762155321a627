"""Regenerates the golden fixtures from the REFERENCE's own C++ runtime (oracle/_ref).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Writes
  tests/golden/gpt2.vocab          gpt2 imported exactly as yaml_guide/convert_gpt2tokenizer.py:37-60
  tests/golden/small_cap2.vocab    a 600-id capcode-2 / NFD stand-in (vocabgen)
  tests/golden/small_nocap.vocab   a 700-id capcode-0 stand-in with alternatives
  tests/golden/cases.json          inputs + the reference's outputs (ids, missing, count, serialized
                                   16/24/32, decoded, streaming decode) for every vocabulary
The fixtures travel to the GPU box, where /root/reference does not exist.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from workloads import capcode_py, corpus, vocabgen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def small_vocab(capcode: int, n: int, seed: int) -> bytes:
    text = corpus.base_text(seed, 400_000, 0.4, pool=3000)
    if capcode == 2:
        text = capcode_py.encode(text)
    toks = corpus.train_tokens(text, n, max_len=16)
    return vocabgen.generate_vocab(list(toks), capcode=capcode, charset=1, normflag=1, level=3,
                                   reserve=(4 | 32) if capcode else 1, scores=toks)


def main():
    assert oracle.have_reference(), "oracle/_ref is required (run make -C oracle)"
    vocabs = {
        "unit_cpp": vocabgen.unit_cpp_vocab(),
        "small_cap2": small_vocab(2, 600, 11),
        "small_nocap": small_vocab(0, 700, 12),
        "gpt2": vocabgen.gpt2_vocab("/root/reference/yaml_guide/gpt2.json"),
    }
    for name in ("small_cap2", "small_nocap", "gpt2"):
        with open(os.path.join(HERE, name + ".vocab"), "wb") as f:
            f.write(vocabs[name])
    base = corpus.base_text(777, 60_000, 0.5, pool=3000)
    inputs = {
        "unit_cpp": [b"ab a z", b"ab", b"", b"z", b"abab  b", b"zzz"],
        "text": [b"", b"a", b" ", b"Hello World", b"The Quick Brown Fox tests TokenMonster tokenization.\n",
                 b"function tokenize(input) { return input.split(/\\s+/).map(x => x.toUpperCase()); }\n",
                 b"HTTP/2 status=200 path=/v1/chat/completions request_id=req_1234567890abcdef.\n",
                 b"CamelCaseWords snake_case_words kebab-case-words MIXEDCaseABC xyzXYZ 1234567890.\n\n",
                 b"don't DON'T Don't I'm I'M  X  Y a1B2 AB12", b"   leading and trailing   ", b"\t\tif (x == 42) {\n"]
        + [base[k:k + 37 + 91 * (k % 7)] for k in range(0, 50_000, 1700)],
    }
    cases = {}
    for name, vb in vocabs.items():
        ref = oracle.OracleVocab(vb, "reference")
        rows = []
        for data in inputs["unit_cpp" if name == "unit_cpp" else "text"]:
            ids, miss = ref.tokenize(data)
            cnt, cmiss = ref.tokenize(data, mode=1)
            row = {"in": data.hex(), "ids": [int(x) for x in ids], "missing": int(miss), "count": int(cnt),
                   "normalized": ref.normalize(data).hex(),
                   "decoded": ref.decode(ids).hex(), "decoded_raw": ref.decode(ids, raw=True).hex()}
            for enc in (2, 3, 4):
                if enc == 2 and ref.n_reverse > 65536:
                    continue
                ser, _ = ref.tokenize(data, mode=enc)
                row[f"ser{enc}"] = ser.hex()
                row[f"dec_ser{enc}"] = ref.decode(ser, enc_len=enc).hex()
            # streaming decoder fed in three uneven pieces (go:701-740)
            dec = ref.new_decoder()
            k1, k2 = len(ids) // 3, (2 * len(ids)) // 3
            row["stream"] = [dec.decode(ids[:k1]).hex(), dec.decode(ids[k1:k2]).hex(), dec.decode(ids[k2:]).hex(),
                             dec.flush().hex()]
            rows.append(row)
        cases[name] = {"info": [ref.capcode, ref.charset, ref.normflag, ref.vocab_size, ref.n_reverse, ref.n_info,
                                ref.max_token_length, ref.delete_token, ref.unk], "rows": rows}
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(cases), "vocabularies,", sum(len(c["rows"]) for c in cases.values()), "cases")


if __name__ == "__main__":
    main()

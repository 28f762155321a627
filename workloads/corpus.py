"""Synthetic text batches of the shapes BASELINE.json names.

Deterministic, offline, independent of /root/reference (which does not exist on the GPU
box).  Text = pseudo-English prose (Zipf word frequencies, sentence capitals, ALLCAPS and
CamelCase words, digits, apostrophes) mixed with code-like lines; pure ASCII by default.

    base_text(seed, nbytes, code_frac)        -> bytes      (a long stream of lines)
    make_batch(base, n, lo, hi, seed, device) -> (data u8, offsets i64) torch tensors;
                                                 each string = a slice of `base` starting at a
                                                 line boundary, length ~ U[lo, hi]
    bench_cpp_corpus(n)                       -> the tests/bench.cpp:39-54 style corpus
                                                 (our own six lines, same categories)

Workload / fixture infrastructure; the product never imports this module.
"""
from __future__ import annotations

import numpy as np

_COMMON = (
    "the of and to a in is that for it as was with be by on not he this are or his from at which but "
    "have an had they you were their one all we can her has there been if more when will would who so no "
    "out up into than them its only time new some could these two may then do first any my now such like our "
    "over man me even most made after also did many before must through back years where much your way well "
    "down should because each just those people how too little state good very make world still own see men "
    "work long get here between both life being under never day same another know while last might us great "
    "old year off come since against go came right used take three states himself few house use during without "
    "again place around however home small found thought went say part once general high upon school every don "
    "does got united left number course war until always away something fact though water less public put think "
    "almost hand enough far took head yet government system better set told nothing night end why called didn "
    "eyes find going look asked later knew point next program city business give group toward young days let room "
    "function return value index string number array object result error data file name type list self class "
    "import const let var int void public private static def for while if else elif switch case break continue "
    "true false null none print length size count key item node next prev left right root map filter reduce "
    "user request response status path server client token vocab encode decode buffer offset stream model batch"
).split()

_ONSET = ["", "b", "c", "d", "f", "g", "h", "j", "k", "l", "m", "n", "p", "r", "s", "t", "v", "w", "br", "ch", "cl",
          "cr", "dr", "fl", "fr", "gl", "gr", "pl", "pr", "sh", "sl", "sp", "st", "str", "th", "tr", "wh", "qu"]
_NUC = ["a", "e", "i", "o", "u", "ai", "ea", "ee", "io", "ou", "oo", "ie", "au"]
_CODA = ["", "", "n", "r", "s", "t", "l", "m", "d", "ng", "nt", "st", "rs", "ck", "ll", "ss", "ch", "sh", "th", "ly",
         "er", "ed", "es", "ing", "ion", "al", "ic", "ous", "ive", "ment", "ness", "able", "ful"]


def word_pool(n: int = 30000, seed: int = 1234) -> list:
    """n distinct lower-case words: common English/code words first, then pseudo-words."""
    rng = np.random.Generator(np.random.PCG64(seed))
    words, seen = [], set()
    for w in _COMMON:
        if w not in seen:
            seen.add(w)
            words.append(w)
    while len(words) < n:
        k = int(rng.integers(1, 4))
        w = ""
        for _ in range(k):
            w += _ONSET[int(rng.integers(len(_ONSET)))] + _NUC[int(rng.integers(len(_NUC)))]
        w += _CODA[int(rng.integers(len(_CODA)))]
        if 2 <= len(w) <= 14 and w not in seen:
            seen.add(w)
            words.append(w)
    return words[:n]


def _zipf_sampler(n: int, rng, s: float = 1.0):
    p = 1.0 / np.power(np.arange(n) + 2.7, s)
    cdf = np.cumsum(p / p.sum())

    def draw(k):
        return np.minimum(np.searchsorted(cdf, rng.random(k)), n - 1)

    return draw


_OPS = [" = ", " == ", " != ", " + ", " - ", " * ", " / ", " < ", " > ", " <= ", " >= ", " && ", " || ", " += ", " -> ", " => "]


def base_text(seed: int, nbytes: int, code_frac: float = 0.3, accents: float = 0.0, pool: int = 30000) -> bytes:
    """A stream of lines, about `nbytes` long (cut at a line end)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    words = word_pool(pool)
    draw = _zipf_sampler(len(words), rng)
    out = []
    size = 0
    buf_ids = draw(1 << 16)
    buf_at = 0
    rnd = rng.random(1 << 16)
    rnd_at = 0

    def W():
        nonlocal buf_ids, buf_at
        if buf_at >= len(buf_ids):
            buf_ids = draw(1 << 16)
            buf_at = 0
        w = words[int(buf_ids[buf_at])]
        buf_at += 1
        return w

    def R():
        nonlocal rnd, rnd_at
        if rnd_at >= len(rnd):
            rnd = rng.random(1 << 16)
            rnd_at = 0
        v = rnd[rnd_at]
        rnd_at += 1
        return v

    def ident():
        r = R()
        if r < 0.4:
            return W()
        if r < 0.7:
            return W() + W().capitalize() + (W().capitalize() if R() < 0.3 else "")
        if r < 0.9:
            return W() + "_" + W()
        return W().upper() + ("_" + W().upper() if R() < 0.5 else "")

    def number():
        r = R()
        if r < 0.6:
            return str(int(R() * 100))
        if r < 0.85:
            return str(int(R() * 100000))
        return "%.2f" % (R() * 1000)

    def prose_line():
        n = 5 + int(R() * 16)
        parts = []
        for k in range(n):
            r = R()
            if r < 0.02:
                w = W().upper()
            elif r < 0.05:
                w = W().capitalize() + W().capitalize()
            elif r < 0.08:
                w = number()
            elif r < 0.10:
                w = W() + ("'s" if R() < 0.6 else "n't")
            elif r < 0.16:
                w = W().capitalize()
            else:
                w = W()
            if accents and R() < accents:
                w = w.replace("e", "é", 1) if "e" in w else w + "ñ"
            if k == 0:
                w = w[:1].upper() + w[1:]
            r = R()
            if r < 0.08 and k < n - 1:
                w += ","
            elif r < 0.10:
                w = "(" + w + ")"
            elif r < 0.115:
                w = '"' + w + '"'
            parts.append(w)
        end = ".?!:;"[min(int(R() * 7), 4)] if R() < 0.9 else "."
        return " ".join(parts) + end

    def code_line():
        ind = ["", "", "    ", "        ", "\t", "\t\t", "  "][int(R() * 7)]
        r = R()
        if r < 0.15:
            return ind + "if (" + ident() + _OPS[1 + int(R() * 10)] + (number() if R() < 0.5 else ident()) + ") {"
        if r < 0.30:
            return ind + ["const ", "let ", "var ", "int ", "auto "][int(R() * 5)] + ident() + " = " + ident() + "." + ident() + "(" + ident() + ", " + number() + ");"
        if r < 0.42:
            return ind + "return " + ident() + "[" + ident() + "]" + _OPS[3 + int(R() * 4)] + number() + ";"
        if r < 0.52:
            return ind + "def " + W() + "_" + W() + "(self, " + W() + ", " + W() + "=None):"
        if r < 0.60:
            return ind + "for (int i = 0; i < " + ident() + "; i++) {"
        if r < 0.70:
            return ind + ident() + ' = {"' + W() + '": ' + number() + ', "' + W() + '": "' + W() + '"}'
        if r < 0.80:
            return ind + "// " + " ".join(W() for _ in range(3 + int(R() * 6)))
        if r < 0.86:
            return ind + "}"
        if r < 0.92:
            return ind + ident() + "(" + ", ".join(ident() for _ in range(1 + int(R() * 3))) + ");"
        return ind + "# " + W().upper() + ": " + ident() + " " + number()

    while size < nbytes:
        is_code = R() < code_frac
        for _ in range(2 + int(R() * 8)):
            line = (code_line() if is_code else prose_line()) + "\n"
            out.append(line)
            size += len(line)
        out.append("\n")
        size += 1
    return "".join(out).encode("utf-8")


def make_batch(base: bytes, n_strings: int, min_len: int, max_len: int, seed: int, device="cpu"):
    """(data uint8 [total], offsets int64 [n+1]) as torch tensors on `device`.

    String s = base[start_s : start_s + len_s], start_s a line start drawn uniformly,
    len_s ~ U[min_len, max_len] (clipped so the slice stays inside `base`).
    Built with torch ops so the 512 MB configurations are generated on the GPU.
    """
    import torch

    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    b = torch.frombuffer(bytearray(base), dtype=torch.uint8)
    nl = torch.nonzero(b == 10).flatten() + 1
    nl = nl[nl < len(base) - max_len - 1]
    if nl.numel() == 0:
        nl = torch.zeros(1, dtype=torch.int64)
    starts = nl[torch.randint(0, nl.numel(), (n_strings,), generator=g)]
    lens = torch.randint(min_len, max_len + 1, (n_strings,), generator=g)
    lens = torch.minimum(lens, torch.tensor(len(base)) - starts)
    off = torch.zeros(n_strings + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lens, 0)
    total = int(off[-1])
    dev = torch.device(device)
    if dev.type == "cpu":  # plain slicing is far faster than index arithmetic on the host
        st, ln = starts.tolist(), lens.tolist()
        blob = b"".join(base[a:a + l] for a, l in zip(st, ln))
        return torch.frombuffer(bytearray(blob), dtype=torch.uint8), off
    b, starts, lens, off_d = b.to(dev), starts.to(dev), lens.to(dev), off.to(dev)
    data = torch.empty(total, dtype=torch.uint8, device=dev)
    chunk = max(1, min(1 << 16, (128 << 20) // max(1, max_len)))  # bounds the index temporaries (~128 MB of positions)
    for lo in range(0, n_strings, chunk):
        hi = min(lo + chunk, n_strings)
        l = lens[lo:hi]
        sid = torch.repeat_interleave(torch.arange(hi - lo, device=dev), l)
        o0 = off_d[lo:hi] - off_d[lo]
        pos = torch.arange(int(l.sum()), device=dev) - o0[sid] + starts[lo:hi][sid]
        data[int(off[lo]):int(off[hi])] = b[pos]
    return data, off_d


def bench_cpp_corpus(nbytes: int) -> bytes:
    """Deterministic six-line corpus in the style of tokenmonster-cpp/tests/bench.cpp:39-54
    (prose / JS / HTTP / SQL / JSON / mixed-case), repeated and truncated to nbytes."""
    seed = (
        "A Lazy Grey Wolf checks TokenMonster encoding and decoding round trips.\n"
        "function splitWords(text) { return text.split(/\\s+/).map(w => w.toLowerCase()); }\n"
        "HTTP/1.1 status=404 path=/v2/embeddings trace_id=trace_abcdef0123456789.\n"
        "SQL: SELECT order_id, updated_at FROM orders WHERE kind = 'RoundTrip' ORDER BY 1 ASC;\n"
        "JSON: {\"vocab\":\"english-1024-clean-v1\",\"top_p\":1,\"echo\":true}\n"
        "PascalCaseWords lower_snake_words dash-joined-words SHOUTINGCaseXYZ abcABC 9876543210.\n\n"
    ).encode()
    reps = nbytes // len(seed) + 1
    return (seed * reps)[:nbytes]


# ----------------------------------------------------------------------------- vocabulary "training"
def train_tokens(encoded_sample: bytes, n_tokens: int, max_len: int = 32) -> dict:
    """Pick ~n_tokens frequent substrings of an (already normalized / capcoded) text sample.

    A frequency heuristic, not TokenMonster's distillation trainer (training/trainvocab.go is
    out of scope); its only job is a stand-in vocabulary with realistic structure: space-prefixed
    words, sub-word pieces, multi-word tokens, capcode-marker tokens.  Returns {token: score}.
    """
    import re
    from collections import Counter

    text = encoded_sample.decode("latin-1")
    pre = re.findall(r" ?[a-z]+(?:'[a-z]+)?| ?[0-9]+|[DCW] | ?[^\sa-z0-9]+|\s+", text)
    cnt = Counter(pre)
    gain = Counter()
    for tok, c in cnt.items():
        if len(tok) <= max_len:
            gain[tok] += c * len(tok)
        if len(tok) > 3 and c > 1:  # sub-word pieces
            for k in range(2, min(len(tok), 9)):
                gain[tok[:k]] += c * k * 0.35
                gain[tok[-k:]] += c * k * 0.25
    big = Counter()
    for a, b in zip(pre, pre[1:]):
        if len(a) + len(b) <= max_len:
            big[a + b] += 1
    for tok, c in big.items():
        if c >= 4:
            gain[tok] += c * len(tok) * 0.5
    tri = Counter()
    for a, b, c3 in zip(pre, pre[1:], pre[2:]):
        if len(a) + len(b) + len(c3) <= max_len:
            tri[a + b + c3] += 1
    for tok, c in tri.items():
        if c >= 8:
            gain[tok] += c * len(tok) * 0.4
    out = {}
    for tok, g in gain.most_common():
        b = tok.encode("latin-1")
        if len(b) >= 2 and len(b) <= 40:
            out[b] = float(g)
            if len(out) >= n_tokens:
                break
    return out

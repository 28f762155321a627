"""Vocabulary factory: builds TokenMonster ``.vocab`` files offline.

No pretrained vocabulary exists on this machine or on the GPU box (the Python binding
downloads them from Hugging Face: python/tokenmonster.py:200-206), so the named
configurations are manufactured here.  This module is a Python restatement of

    go/tokenmonster.go:2898-3801  PrivateGenerateVocab  (flags, nWords, alternatives,
                                   "D "-duplicates, beginByte, ids)
    go/tokenmonster.go:2602-2653  Save                   (binary layout)

for the "fresh vocabulary from a token list" case (no resize / delete / special tokens /
modification of an existing vocabulary -- those are offline tooling, out of scope).
It is workload / fixture infrastructure: the product never calls it.  Every file it emits
is validated by loading it with the reference's own C++ loader (oracle/_ref), whose
order self-check (tokenmonster-cpp/src/tokenmonster.cpp:1352-1357) catches ordering errors.
"""
from __future__ import annotations

import struct
import unicodedata
from typing import Dict, Iterable, List, Optional

DNE = 0xFFFFFF
RUNE_ERROR = 0xFFFD


# ----------------------------------------------------------------------------- runes
def decode_rune(b: bytes):
    """utf8.DecodeRune: (rune, size); invalid -> (U+FFFD, 1); empty -> (U+FFFD, 0)."""
    if not b:
        return RUNE_ERROR, 0
    b0 = b[0]
    if b0 < 0x80:
        return b0, 1
    if b0 < 0xC2:
        return RUNE_ERROR, 1
    if b0 < 0xE0:
        if len(b) < 2 or (b[1] & 0xC0) != 0x80:
            return RUNE_ERROR, 1
        return ((b0 & 0x1F) << 6) | (b[1] & 0x3F), 2
    if b0 < 0xF0:
        if len(b) < 3 or (b[1] & 0xC0) != 0x80 or (b[2] & 0xC0) != 0x80:
            return RUNE_ERROR, 1
        if b0 == 0xE0 and b[1] < 0xA0:
            return RUNE_ERROR, 1
        if b0 == 0xED and b[1] >= 0xA0:
            return RUNE_ERROR, 1
        return ((b0 & 0x0F) << 12) | ((b[1] & 0x3F) << 6) | (b[2] & 0x3F), 3
    if b0 < 0xF5:
        if len(b) < 4 or any((x & 0xC0) != 0x80 for x in b[1:4]):
            return RUNE_ERROR, 1
        if b0 == 0xF0 and b[1] < 0x90:
            return RUNE_ERROR, 1
        if b0 == 0xF4 and b[1] >= 0x90:
            return RUNE_ERROR, 1
        return ((b0 & 0x07) << 18) | ((b[1] & 0x3F) << 12) | ((b[2] & 0x3F) << 6) | (b[3] & 0x3F), 4
    return RUNE_ERROR, 1


def decode_last_rune(b: bytes) -> int:
    """utf8.DecodeLastRune (rune only)."""
    end = len(b)
    if end == 0:
        return RUNE_ERROR
    if b[end - 1] < 0x80:
        return b[end - 1]
    lim = max(end - 4, 0)
    start = end - 2
    while start >= lim:
        if (b[start] & 0xC0) != 0x80:
            break
        start -= 1
    if start < 0:
        start = 0
    r, size = decode_rune(b[start:end])
    if start + size != end:
        return RUNE_ERROR
    return r


def _cat(r: int) -> str:
    return unicodedata.category(chr(r)) if 0 <= r <= 0x10FFFF and not (0xD800 <= r <= 0xDFFF) else "Cn"


def is_letter(r: int, capcode: int) -> bool:  # go:359-361
    c = _cat(r)
    return (c[0] == "L" and (capcode != 2 or r not in (0x57, 0x43, 0x44))) or c in ("Mn", "Mc", "Me")


def is_number(r: int) -> bool:  # unicode.IsNumber
    return _cat(r)[0] == "N"


def is_alnum(r: int, capcode: int) -> bool:  # go:363-365
    return is_letter(r, capcode) or is_number(r)


def is_capcode(r: int, capcode: int) -> bool:  # go:367-369
    return (capcode == 1 and r == 0x7F) or (capcode == 2 and r in (0x43, 0x57, 0x44))


def is_space(r: int) -> bool:  # unicode.IsSpace
    if r <= 0xFF:
        return r in (0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0)
    return _cat(r) in ("Zs", "Zl", "Zp")


# ----------------------------------------------------------------------------- reserve sets (go:301-357)
def _not_capital(i: int, capcode: int) -> bool:
    return capcode != 2 or not (0x41 <= i <= 0x5A and i not in (0x43, 0x57, 0x44))


def reserved_bytes(reserve: int, capcode: int, lowercase: bool = False) -> List[bool]:
    t = [False] * 256

    def ascii_():
        for i in range(32, 127):
            if _not_capital(i, capcode):
                t[i] = True
        t[9] = t[10] = t[13] = True
        if capcode == 1:
            t[127] = True

    if reserve & 1:
        for i in range(256):
            if _not_capital(i, capcode):
                t[i] = True
    if reserve & 2:
        for i in range(128):
            if _not_capital(i, capcode):
                t[i] = True
    if reserve & 4:
        ascii_()
        for i in range(0x80, 0xC0):
            t[i] = True
        for i in range(0xC2, 0xF5):
            t[i] = True
    if reserve & 8:
        ascii_()
    if reserve & 16:
        s = "£€©®™°%¢¥—–•‘’“”áéíóúýàèìòùâêîôûäëïöüñãõçåæœ"
        if capcode != 2 and not lowercase:
            s += "ÁÉÍÓÚÝÀÈÌÒÙÂÊÎÔÛÄËÏÖÜÑÃÕÇÅÆŒ"
        for b in unicodedata.normalize("NFD", s).encode():  # callers use NFD or none
            t[b] = True
        ascii_()
    return t


# ----------------------------------------------------------------------------- generator
class _Rec:
    __slots__ = ("key", "flag", "nwords", "alt", "alt2", "alen", "alen2", "id", "score")


def _has_suffix_pos(key: bytes, capcode: int) -> int:  # go:287-299
    for suf in (b"'s", "’s".encode()):
        if key.endswith(suf) and len(suf) < len(key):
            if is_letter(decode_last_rune(key[: len(key) - len(suf)]), capcode):
                return len(key) - len(suf)
    return -1


def generate_vocab(
    tokens: Iterable[bytes],
    *,
    capcode: int = 0,
    charset: int = 1,
    normflag: int = 0,
    level: int = 5,
    reserve: int = 0,
    ids: Optional[Dict[bytes, int]] = None,
    scores: Optional[Dict[bytes, float]] = None,
    unk: bool = False,
) -> bytes:
    """Return the bytes of a ``.vocab`` file.  ``tokens`` are already normalized/capcoded."""
    assert charset in (0, 1), "UTF-16 vocabularies are not manufactured here"
    ids = dict(ids or {})
    scores = dict(scores or {})
    used = set(ids.values())
    exclude_other = bool(reserve & 32)
    table = reserved_bytes(reserve, capcode, bool(normflag & 2))
    multi = set()
    for t in tokens:
        if 0 < len(t) <= 40:  # go:3204
            if len(t) == 1:
                if not exclude_other:
                    table[t[0]] = True
            else:
                multi.add(bytes(t))
    singles = [bytes([i]) for i in range(256) if table[i]]
    dic1 = sorted(set(singles) | multi, key=lambda k: (len(k), k))
    total = len(dic1)

    unk_token = DNE
    if unk:  # go:3382-3398
        if total not in used:
            unk_token = total
        else:
            unk_token = 0
            while unk_token in used:
                unk_token += 1
        if not ((len(singles) < 256 and capcode != 2) or len(singles) < 233):  # go:438-443
            unk_token = DNE
    if unk_token != DNE:
        total += 1
        used.add(unk_token)
    vocab_size = total
    max_id = vocab_size
    for v in ids.values():
        max_id = max(max_id, v + 1)
    if unk_token != DNE:
        max_id = max(max_id, unk_token + 1)

    # ids + the second dictionary with "D "-prefixed duplicates (go:3423-3471)
    add = b"D " if capcode == 2 else b"\x7f "
    nxt = 0
    while nxt in used:
        nxt += 1
    keys: Dict[bytes, int] = {}
    for tok in dic1:
        if tok in ids:
            idx = ids[tok]
        else:
            idx = nxt
            ids[tok] = idx
            nxt += 1
            while nxt in used:
                nxt += 1
        keys[tok] = idx
        r, _ = decode_rune(tok)
        if capcode != 0 and is_alnum(r, capcode):
            dup = add + tok
            if len(dup) <= 40:
                keys[dup] = idx
                ids[dup] = idx
                scores[dup] = -1.0
    order = sorted(keys, key=lambda k: (len(k), k))
    index_of = {k: i for i, k in enumerate(order)}
    max_len = len(order[-1]) if order else 0

    begin = [[0, 0, 0, 0] for _ in range(256)]
    recs: List[_Rec] = []
    for tok in order:
        rec = _Rec()
        rec.key = tok
        rec.id = keys[tok]
        rec.score = float(scores.get(tok, 0.0))
        rec.alt = rec.alt2 = DNE
        rec.alen = rec.alen2 = 0
        flag = 0
        nwords = 0
        min_alt = 1
        only_letter_space = only_number_space = only_punc = False
        r, n = decode_rune(tok)
        r2, n2 = decode_rune(tok[n:])
        if r == 0x20:
            flag = 4
            begin[tok[0]][0] += 1
            if is_alnum(r2, capcode):
                nwords += 1
                min_alt = 2
        elif is_letter(r, capcode):
            flag = 2
            begin[tok[0]][1] += 1
        elif is_capcode(r, capcode):
            if r in (0x43, 0x57):
                flag = 4
            flag |= 16
            begin[tok[0]][3] += 1
        elif is_number(r):
            begin[tok[0]][2] += 1
        else:
            begin[tok[0]][3] += 1
        if len(tok) == 1:
            only_punc = True
        else:
            if (r == 0x20 or is_letter(r, capcode)) and is_letter(r2, capcode):
                only_letter_space = True
            elif (r == 0x20 or is_number(r)) and is_number(r2):
                only_number_space = True
            elif not is_alnum(r, capcode) and not is_alnum(r2, capcode):
                only_punc = True
            i = n + n2
            while i < len(tok):
                r = r2
                r2, n2 = decode_rune(tok[i:])
                if r == 0x20 and is_alnum(r2, capcode):
                    nwords += 1
                if is_letter(r2, capcode):
                    only_punc = False
                    only_number_space = False
                elif is_number(r2):
                    only_punc = False
                    only_letter_space = False
                elif r2 != 0x20:
                    only_letter_space = False
                    only_number_space = False
                i += n2
        nwords &= 0xFF
        rl = decode_last_rune(tok)
        if min_alt == 2 and is_letter(rl, capcode) and only_letter_space and nwords == 1:
            flag |= 32
        if min_alt == 2 and nwords <= 1:
            min_alt = 1
        if is_capcode(rl, capcode):
            flag |= 8
        if is_letter(rl, capcode):
            flag |= 1
        if only_letter_space or only_number_space or only_punc:
            flag |= 128

        suffix_at = _has_suffix_pos(tok, capcode)
        p1 = p2 = 0

        def assign(prio: int, index: int, length: int):
            nonlocal p1, p2
            if p1 < p2 or (p1 == p2 and rec.alen <= rec.alen2):
                if p1 < prio:
                    rec.alt, rec.alen, p1 = index, length, prio
            else:
                if p2 < prio:
                    rec.alt2, rec.alen2, p2 = index, length, prio

        length = len(tok) - 1
        while length >= min_alt:
            sub = tok[:length]
            index = index_of.get(sub)
            if index is not None:
                done = False
                if length <= len(tok) - 2 and tok[length] == 0x20:
                    rr, _ = decode_rune(tok[length + 1:])
                    if is_letter(rr, capcode) or is_number(rr):
                        assign(10, index, length)
                        done = True
                if not done:
                    ra = decode_last_rune(sub)
                    rb, _ = decode_rune(tok[length:])
                    la, lb = is_letter(ra, capcode), is_letter(rb, capcode)
                    na, nb = is_number(ra), is_number(rb)
                    if capcode == 0 and (((not la and ra != 0x5F) and (lb or rb == 0x5F)) or (not na and nb)):
                        assign(9, index, length)
                    elif ((la or ra == 0x5F) and (not lb and rb != 0x5F)) or (na and not nb):
                        assign(9, index, length)
                    elif is_space(ra) and not is_space(rb):
                        assign(7, index, length)
                    elif not is_space(ra) and is_space(rb):
                        assign(8, index, length)
                    elif is_capcode(rb, capcode):
                        assign(9, index, length)
                    elif length == suffix_at:
                        assign(8, index, length)
                        break  # Go's `break` here leaves the prefix loop (go:3734)
                    else:
                        assign(1, index, length)
            length -= 1

        if rec.alen2 > 0 and (p2 > p1 or (p2 == p1 and rec.alen2 > rec.alen)):
            rec.alt, rec.alt2 = rec.alt2, rec.alt
            rec.alen, rec.alen2 = rec.alen2, rec.alen
        rec.flag = flag
        rec.nwords = nwords
        recs.append(rec)

    begin_byte = bytearray(256)
    for i in range(256):  # go:3780-3787
        sp, le, nu, ot = begin[i]
        if le > sp and le > nu and le > ot and le > 2:
            begin_byte[i] = 1
        elif sp > le and sp > nu and sp > ot and sp > 2:
            begin_byte[i] = 12
        elif ot > sp and ot > le and ot > nu and ot > 2:
            begin_byte[i] = 10

    delete_token = DNE
    marker = b"D" if capcode == 2 else (b"\x7f" if capcode == 1 else None)
    if marker is not None and marker in index_of:
        delete_token = recs[index_of[marker]].id

    return save_vocab(
        recs, capcode=capcode, charset=charset, normflag=normflag, level=level, reserve=reserve,
        unk=unk_token, vocab_size=vocab_size, n_reverse=max_id, delete_token=delete_token,
        max_len=max_len, begin_byte=bytes(begin_byte),
    )


def _u24(x: int) -> bytes:
    return bytes((x & 0xFF, (x >> 8) & 0xFF, (x >> 16) & 0xFF))


def save_vocab(recs, *, capcode, charset, normflag, level, reserve, unk, vocab_size, n_reverse,
               delete_token, max_len, begin_byte: bytes) -> bytes:
    """go/tokenmonster.go:2602-2653."""
    out = bytearray()
    out += bytes((capcode, charset, normflag, level, reserve, 0, 0, 0))
    out += _u24(unk) + _u24(vocab_size) + _u24(n_reverse) + _u24(len(recs)) + _u24(delete_token)
    out.append(max_len)
    for i, r in enumerate(recs):
        assert (r.alt == DNE or r.alt < i) and (r.alt2 == DNE or r.alt2 < i), "corrupt (go:2636)"
        out.append(len(r.key))
        out += r.key
        out.append(r.flag)
        out.append(r.nwords)
        out += _u24(r.alt) + _u24(r.alt2) + _u24(r.id)
        out += struct.pack("<f", r.score)
    out += begin_byte
    out += _u24(0)
    return bytes(out)


def raw_vocab(records, *, capcode=0, charset=0, normflag=0, level=5, reserve=0, unk=DNE,
              delete_token=DNE, begin_byte: Optional[bytes] = None, vocab_size=None, n_reverse=None) -> bytes:
    """Hand-assembled vocabulary for edge-case tests (the tests/unit.cpp:50-79 recipe).

    records: iterable of (key, id, flag, nwords, alt_index, alt_index2), already sorted by (len, key).
    """
    recs = []
    for key, id_, flag, nwords, alt, alt2 in records:
        r = _Rec()
        r.key, r.id, r.flag, r.nwords, r.alt, r.alt2, r.score = bytes(key), id_, flag, nwords, alt, alt2, 1.0
        r.alen = r.alen2 = 0
        recs.append(r)
    nrev = (max((r.id for r in recs), default=-1) + 1) if n_reverse is None else n_reverse
    return save_vocab(
        recs, capcode=capcode, charset=charset, normflag=normflag, level=level, reserve=reserve, unk=unk,
        vocab_size=len({r.id for r in recs}) if vocab_size is None else vocab_size, n_reverse=nrev,
        delete_token=delete_token, max_len=max((len(r.key) for r in recs), default=0),
        begin_byte=begin_byte if begin_byte is not None else bytes(256),
    )


def unit_cpp_vocab() -> bytes:
    """The 4-token vocabulary of the reference's tests/unit.cpp:50-79."""
    return raw_vocab([(b" ", 0, 0, 0, DNE, DNE), (b"a", 1, 0, 0, DNE, DNE),
                      (b"b", 2, 0, 0, DNE, DNE), (b"ab", 3, 0, 0, DNE, DNE)], level=5)


def gpt2_tokens(json_path: str):
    """yaml_guide/convert_gpt2tokenizer.py:37-60: Ġ/Ċ/č/ĉ mapped back to bytes, ids preserved."""
    import json

    with open(json_path, "r", encoding="utf-8") as f:
        table = json.load(f)
    toks, ids = [], {}
    for original, id_ in table.items():
        t = original.replace("Ġ", " ").replace("Ċ", "\n").replace("č", "\r").replace("ĉ", "\t")
        b = t.encode("utf-8")
        toks.append(b)
        if 0 < len(b):
            ids[b] = id_
    return toks, ids


def gpt2_vocab(json_path: str) -> bytes:
    toks, ids = gpt2_tokens(json_path)
    ids = {k: v for k, v in ids.items() if len(k) <= 40}
    return generate_vocab(toks, capcode=0, charset=1, normflag=0, level=5, reserve=0, ids=ids)

"""oracle/unicode_oracle.py -- TEST INFRASTRUCTURE: the non-ASCII half of the normalization oracle.

Python restatement (over `unicodedata`, independent of the product's generated tables) of what
`Vocab.Tokenize` does to text before the match loop (go/tokenmonster.go:242-253):
  norm.Normalizer.Normalize   order of tokenmonster-cpp/src/tokenmonster.cpp:428-475, flags
                              nfd (1), lowercase (2), accents (4)          -> normalize_flags()
  capcode.Encode (level 2)    javascript/tokenmonster.js:900-1005, per code point, Unicode
                              categories Lu / Ll / L / Nd / M              -> capcode_encode()
  capcode decode              javascript/tokenmonster.js:1007-1065         -> capcode_decode()
Case mappings are the SIMPLE (single code point) ones, as Go's unicode.ToLower / ToUpper; the
JS restatement uses full mappings, which differ only for U+0130 and the multi-character
upper-casings ("unpinned").

PARITY UNPINNED for non-ASCII: the reference's own normalizer (golang.org/x/text NFD, ICU in the
C++ build) and the Go capcode package are not in /root/reference and cannot run here; this
module restates the published algorithms (UAX #15 NFD = unicodedata.normalize) and is checked by
losslessness and against the ASCII oracle (oracle/tm_oracle.c) on ASCII input.
Strings this path refuses (status 13, like the product): invalid UTF-8, and more than 30
consecutive non-starters (x/text inserts U+034F there; ICU does not -- no defined answer).
"""
from __future__ import annotations

import unicodedata as ud

MAX_NONSTARTERS = 30


class Refused(Exception):
    """normalization failed (server status 13)"""


def _simple_lower(ch: str) -> str:
    s = ch.lower()
    if len(s) == 1:
        return s
    return "i" if ch == "İ" else ch


def _simple_upper(ch: str) -> str:
    s = ch.upper()
    if len(s) == 1:
        return s
    t = ch.title()
    return t if len(t) == 1 else ch


def normalize_flags(data: bytes, normflag: int) -> str:
    """cpp:428-475 restricted to flags 1|2|4; returns code points."""
    try:
        s = data.decode("utf-8")
    except UnicodeDecodeError:
        raise Refused("invalid UTF-8")
    if normflag & ~7:
        raise Refused("flag not covered")
    if normflag & 4:                       # accents: NFD, drop Mn, optional lowercase (cpp:464-468)
        s = ud.normalize("NFD", s)
        _check_runs(s)
        s = "".join(c for c in s if ud.category(c) != "Mn")
        if normflag & 2:
            s = "".join(_simple_lower(c) for c in s)
    elif normflag & 2:                     # optional NFD, then lowercase (cpp:469-472)
        if normflag & 1:
            s = ud.normalize("NFD", s)
            _check_runs(s)
        s = "".join(_simple_lower(c) for c in s)
    elif normflag & 1:
        s = ud.normalize("NFD", s)
        _check_runs(s)
    return s


def _check_runs(s: str) -> None:
    run = 0
    for c in s:
        run = run + 1 if ud.combining(c) else 0
        if run > MAX_NONSTARTERS:
            raise Refused("more than 30 consecutive non-starters")


def _up(c): return ud.category(c) == "Lu"
def _lo(c): return ud.category(c) == "Ll"
def _let(c): return ud.category(c)[0] == "L"
def _num(c): return ud.category(c) == "Nd"
def _mod(c): return ud.category(c)[0] == "M"
def _apo(c): return c == "'" or c == "’"


def capcode_encode(s: str) -> str:
    """js:900-1005, one list element per code point (the JS buffer holds code points too)."""
    buf = []
    rlast = rlast2 = "."
    in_word = multi = False
    word_tok = goback = 0
    for r in s:
        if in_word:
            if _up(r):
                if not (_let(rlast) or _apo(rlast) or _mod(rlast)):
                    buf += ["D", " "]
                multi = True
                buf.append(_simple_lower(r))
            else:
                if _lo(r):
                    in_word = False
                    buf[word_tok] = "C"
                    if multi:                                   # js:924-951
                        seg, out, k = buf[goback:], [], 0
                        while k < len(seg):
                            if seg[k] == "D" and k + 1 < len(seg) and seg[k + 1] == " ":
                                nxt = seg[k + 2] if k + 2 < len(seg) else None
                                if nxt is not None and _lo(nxt):
                                    out += ["D", "C", " ", nxt]
                                    k += 3                       # js: i2++ then i2 += 2, loop i2++
                                else:
                                    out += ["D", " "]
                                    k += 2
                                    if nxt is not None:
                                        out.append(nxt)
                                        k += 1
                            elif _lo(seg[k]):
                                out += ["D", "C", " ", seg[k]]
                                k += 1
                            else:
                                out.append(seg[k])
                                k += 1
                        buf[goback:] = out
                    if not (_let(rlast) or _apo(rlast) or _mod(rlast)):
                        buf += ["D", " "]
                elif _num(r):
                    if not _num(rlast):
                        buf += ["D", " "]
                elif not (_apo(r) or _mod(r)):
                    in_word = False
                buf.append(r)
        elif _lo(r):
            if not (rlast == " " or _let(rlast) or (_let(rlast2) and _apo(rlast)) or _mod(rlast)):
                buf += ["D", " "]
            buf.append(r)
        elif _up(r):
            if rlast == " ":
                word_tok = len(buf) - 1
                buf[word_tok] = "W"
                buf.append(" ")
            else:
                buf.append("D")
                word_tok = len(buf)
                buf += ["W", " "]
            buf.append(_simple_lower(r))
            goback = len(buf)
            multi = False
            in_word = True
        elif _num(r):
            if not (rlast == " " or _num(rlast)):
                buf += ["D", " "]
            buf.append(r)
        else:
            buf.append(r)
        rlast2, rlast = rlast, r
    return "".join(buf)


def capcode_decode(s: str, state=None) -> str:
    """js:1007-1065; `state` = [inWord, inChar, delete, ignore] persists for streaming decoders."""
    st = state if state is not None else [False, False, False, False]
    out = []
    for r in s:
        if r == "C":
            st[1], st[0] = True, False
            continue
        if r == "W":
            st[0], st[1], st[3] = True, False, True
            continue
        if r == "D":
            st[2] = True
            continue
        if r == " ":
            if st[2]:
                st[2] = False
            else:
                out.append(" ")
                if not st[3]:
                    st[0] = False
        elif st[2]:
            st[2] = False
        elif st[1]:
            st[1] = False
            out.append(_simple_upper(r))
        elif st[0]:
            if _lo(r) or _up(r):
                out.append(_simple_upper(r))
            else:
                out.append(r)
                if not (_num(r) or _apo(r) or _mod(r)):
                    st[0] = False
        else:
            out.append(r)
        st[3] = False
    return "".join(out)


def normalize(data: bytes, normflag: int, capcode: int) -> bytes:
    """Everything Vocab.Tokenize does before the match loop, for UTF-8 vocabularies."""
    s = normalize_flags(data, normflag) if normflag else _decode(data)
    if capcode == 2:
        s = capcode_encode(s)
    elif capcode != 0:
        raise Refused("capcode level 1 has no specification")
    return s.encode("utf-8")


def _decode(data: bytes) -> str:
    try:
        return data.decode("utf-8")
    except UnicodeDecodeError:
        raise Refused("invalid UTF-8")

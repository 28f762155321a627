"""Pure-Python capcode (level 2) for ASCII text -- workload tooling.

Used to prepare the text sample a stand-in vocabulary is "trained" on, and by the tests as a
second, independent restatement of javascript/tokenmonster.js:900-1065 to cross-check the C
oracle against.  The product never imports this module.
"""
from __future__ import annotations


def _up(c): return 65 <= c <= 90
def _lo(c): return 97 <= c <= 122
def _let(c): return _up(c) or _lo(c)
def _num(c): return 48 <= c <= 57
def _glue(c): return _let(c) or c == 39


def encode(data: bytes) -> bytes:
    out = bytearray()
    rlast, rlast2 = 46, 46
    in_word = multi = False
    word_tok = goback = 0
    for r in data:
        if r >= 0x80:
            raise ValueError("capcode_py: ASCII only")
        if in_word:
            if _up(r):
                if not _glue(rlast):
                    out += b"D "
                multi = True
                out.append(r | 0x20)
            else:
                if _lo(r):
                    in_word = False
                    out[word_tok] = 67  # 'C'
                    if multi:
                        seg = bytes(out[goback:])
                        del out[goback:]
                        k = 0
                        while k < len(seg):
                            if seg[k] == 68 and k + 2 < len(seg) and seg[k + 1] == 32:
                                out += (b"DC " if _lo(seg[k + 2]) else b"D ") + seg[k + 2:k + 3]
                                k += 3
                            elif _lo(seg[k]):
                                out += b"DC " + seg[k:k + 1]
                                k += 1
                            else:
                                out.append(seg[k])
                                k += 1
                    if not _glue(rlast):
                        out += b"D "
                elif _num(r):
                    if not _num(rlast):
                        out += b"D "
                elif r != 39:
                    in_word = False
                out.append(r)
        elif _lo(r):
            if not (rlast == 32 or _let(rlast) or (_let(rlast2) and rlast == 39)):
                out += b"D "
            out.append(r)
        elif _up(r):
            if rlast == 32:
                word_tok = len(out) - 1
                out[word_tok] = 87  # 'W'
                out.append(32)
            else:
                out.append(68)
                word_tok = len(out)
                out += b"W "
            out.append(r | 0x20)
            goback = len(out)
            multi = False
            in_word = True
        elif _num(r):
            if not (rlast == 32 or _num(rlast)):
                out += b"D "
            out.append(r)
        else:
            out.append(r)
        rlast2, rlast = rlast, r
    return bytes(out)


def decode(data: bytes) -> bytes:
    out = bytearray()
    in_word = in_char = dele = ignore = False
    for r in data:
        if r == 67:
            in_char, in_word = True, False
            continue
        if r == 87:
            in_word, in_char, ignore = True, False, True
            continue
        if r == 68:
            dele = True
            continue
        if r == 32:
            if dele:
                dele = False
            else:
                out.append(32)
                if not ignore:
                    in_word = False
        elif dele:
            dele = False
        elif in_char:
            in_char = False
            out.append(r - 32 if _lo(r) else r)
        elif in_word:
            if _let(r):
                out.append(r - 32 if _lo(r) else r)
            else:
                out.append(r)
                if not (_num(r) or r == 39):
                    in_word = False
        else:
            out.append(r)
        ignore = False
    return bytes(out)

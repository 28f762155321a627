"""ctypes front end for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two interchangeable back ends behind one class:

* ``port``      -- oracle/_build/libtmoracle.so, the plain-C restatement (tm_oracle.c)
* ``reference`` -- oracle/_ref/libtmref.so, the reference's own C++ runtime
                   (tokenmonster-cpp/src/tokenmonster.cpp compiled unmodified, in place)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker (never as a product path).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libtmoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtmref.so")


def build(force: bool = False) -> None:
    """Compile the C restatement and (when /root/reference exists) oracle/_ref."""
    need = force or not os.path.exists(PORT_SO)
    if os.path.isdir("/root/reference/tokenmonster-cpp") and not os.path.exists(REF_SO):
        need = True
    if need:
        subprocess.run(["make", "-C", HERE, "-s"], check=True)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


class OracleError(RuntimeError):
    def __init__(self, status: int, what: str = ""):
        super().__init__(f"oracle status {status} {what}")
        self.status = status


def _bind(path: str, p: str):
    lib = C.CDLL(path)
    u8p = C.POINTER(C.c_uint8)

    def f(name, res, args):
        fn = getattr(lib, p + name)
        fn.restype = res
        fn.argtypes = args
        return fn

    api = {
        "vocab_load": f("vocab_load", C.c_void_p, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]),
        "vocab_free": f("vocab_free", None, [C.c_void_p]),
        "vocab_info": f("vocab_info", C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
        "free": f("free", None, [C.c_void_p]),
        "normalize": f("normalize", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        "tokenize": f("tokenize", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]),
        "decode": f("decode", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        "decoder_new": f("decoder_new", C.c_void_p, [C.c_void_p]),
        "decoder_free": f("decoder_free", None, [C.c_void_p]),
        "decoder_decode": f("decoder_decode", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        "decoder_flush": f("decoder_flush", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    }
    if p == "tmr_":
        api["tokenize_batch"] = f("tokenize_batch", C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)])
    else:
        api["score"] = f("score", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_uint64)])
        api["longest_prefix"] = f("longest_prefix", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32)])
        api["capcode_encode"] = f("capcode_encode", C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)])
        api["capcode_decode"] = f("capcode_decode", C.c_size_t, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p])
        api["incomplete_utf8"] = f("incomplete_utf8", C.c_int, [C.c_char_p, C.c_size_t])
        api["incomplete_utf16"] = f("incomplete_utf16", C.c_int, [C.c_char_p, C.c_size_t])
        api["ids_checksum"] = f("ids_checksum", C.c_uint64, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32])
    return api


_APIS: dict[str, dict] = {}


def _api(kind: str):
    if kind not in _APIS:
        build()
        if kind == "port":
            _APIS[kind] = _bind(PORT_SO, "tmo_")
        elif kind == "reference":
            if not have_reference():
                raise FileNotFoundError(REF_SO)
            _APIS[kind] = _bind(REF_SO, "tmr_")
        else:
            raise ValueError(kind)
    return _APIS[kind]


def _as_tokens_arg(tokens, enc_len: int):
    if enc_len == 0:
        a = np.ascontiguousarray(tokens, dtype=np.uint32)
        return a, a.ctypes.data_as(C.c_void_p), a.size
    b = bytes(tokens)
    buf = np.frombuffer(b, dtype=np.uint8) if b else np.zeros(0, np.uint8)
    return buf, buf.ctypes.data_as(C.c_void_p), len(b)


class OracleVocab:
    """CPU oracle vocabulary.  kind = 'port' (C restatement) or 'reference' (oracle/_ref)."""

    def __init__(self, file_bytes: bytes, kind: str = "port"):
        self.kind = kind
        self.api = _api(kind)
        st = C.c_int(0)
        self.h = self.api["vocab_load"](file_bytes, len(file_bytes), C.byref(st))
        if not self.h:
            raise OracleError(st.value, "vocab_load")
        info = (C.c_uint32 * 12)()
        self.api["vocab_info"](self.h, info)
        (self.capcode, self.charset, self.normflag, self.level, self.reserve, self.unk,
         self.vocab_size, self.n_reverse, self.n_info, self.delete_token,
         self.max_token_length) = list(info)[:11]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.api["vocab_free"](self.h)
                self.h = None
        except Exception:
            pass

    def _take(self, ptr, nbytes: int) -> bytes:
        out = C.string_at(ptr, nbytes) if (ptr and nbytes) else b""
        if ptr:
            self.api["free"](ptr)
        return out

    def normalize(self, data: bytes) -> bytes:
        p, n = C.c_void_p(), C.c_size_t()
        st = self.api["normalize"](self.h, data, len(data), C.byref(p), C.byref(n))
        if st:
            raise OracleError(st, "normalize")
        return self._take(p.value, n.value)

    def tokenize(self, data: bytes, normalized: bool = False, mode: int = 0):
        """mode 0 -> (uint32 array, missing); 1 -> (count, missing); 2/3/4 -> (bytes, missing)."""
        p, n, miss = C.c_void_p(), C.c_size_t(), C.c_int64()
        st = self.api["tokenize"](self.h, data, len(data), int(normalized), mode, C.byref(p), C.byref(n), C.byref(miss))
        if st:
            raise OracleError(st, "tokenize")
        if mode == 1:
            if p.value:
                self.api["free"](p.value)
            return n.value, miss.value
        width = 4 if mode == 0 else mode
        raw = self._take(p.value, n.value * width)
        if mode == 0:
            return np.frombuffer(raw, dtype=np.uint32).copy(), miss.value
        return raw, miss.value

    def score(self, texts, normalized: bool = False):
        """The training worker's loop (training/trainvocab.go:932-1170) over a list of byte strings
        -> (scores uint64[n_reverse], miss_hist uint64[256], tokens_in_text).  Port only."""
        scores = np.zeros(max(int(self.n_reverse), 1), np.uint64)
        hist = np.zeros(256, np.uint64)
        total = C.c_uint64(0)
        for t in texts:
            st = self.api["score"](self.h, t, len(t), int(normalized), scores.ctypes.data, hist.ctypes.data,
                                   C.byref(total))
            if st:
                raise OracleError(st, "score")
        return scores, hist, int(total.value)

    def decode(self, tokens, enc_len: int = 0, raw: bool = False) -> bytes:
        keep, ptr, cnt = _as_tokens_arg(tokens, enc_len)
        p, n = C.c_void_p(), C.c_size_t()
        st = self.api["decode"](self.h, ptr, cnt, enc_len, int(raw), C.byref(p), C.byref(n))
        if st:
            raise OracleError(st, "decode")
        return self._take(p.value, n.value)

    def new_decoder(self) -> "OracleDecoder":
        return OracleDecoder(self)

    def longest_prefix(self, window: bytes):
        idx = C.c_uint32()
        l = self.api["longest_prefix"](self.h, window, len(window), C.byref(idx))
        return (idx.value, l) if l else (None, 0)

    def tokenize_batch_timed(self, data: np.ndarray, off: np.ndarray, threads: int):
        """reference kind only: whole batch on `threads` host threads; returns (total ids, checksum)."""
        chk = C.c_uint64()
        tot = self.api["tokenize_batch"](self.h, data.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                         len(off) - 1, threads, None, C.byref(chk))
        return tot, chk.value


class OracleDecoder:
    def __init__(self, vocab: OracleVocab):
        self.v = vocab
        self.h = vocab.api["decoder_new"](vocab.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.v.api["decoder_free"](self.h)
                self.h = None
        except Exception:
            pass

    def decode(self, tokens, enc_len: int = 0) -> bytes:
        keep, ptr, cnt = _as_tokens_arg(tokens, enc_len)
        p, n = C.c_void_p(), C.c_size_t()
        st = self.v.api["decoder_decode"](self.h, ptr, cnt, enc_len, C.byref(p), C.byref(n))
        if st:
            raise OracleError(st, "decoder_decode")
        return self.v._take(p.value, n.value)

    def flush(self) -> bytes:
        p, n = C.c_void_p(), C.c_size_t()
        self.v.api["decoder_flush"](self.h, C.byref(p), C.byref(n))
        return self.v._take(p.value, n.value)


def capcode_encode(data: bytes) -> bytes:
    api = _api("port")
    p, n = C.c_void_p(), C.c_size_t()
    st = api["capcode_encode"](data, len(data), C.byref(p), C.byref(n))
    if st:
        raise OracleError(st, "capcode_encode")
    out = C.string_at(p.value, n.value) if n.value else b""
    api["free"](p.value)
    return out


def capcode_decode(data: bytes, state: int = 0) -> tuple[bytes, int]:
    """Stateful capcode level-2 decode of one piece of text (tmo_capcode_decode).  `state` packs the
    decoder's flags as the device does: 1 inWord, 2 inChar, 4 delete, 8 ignore.  Returns (bytes, state)."""
    api = _api("port")
    st = (C.c_uint8 * 4)(state & 1, (state >> 1) & 1, (state >> 2) & 1, (state >> 3) & 1)
    out = C.create_string_buffer(len(data) + 1)
    n = api["capcode_decode"](st, data, len(data), out)
    return out.raw[:n], int(st[0]) | (int(st[1]) << 1) | (int(st[2]) << 2) | (int(st[3]) << 3)


def ids_checksum(tokens_u8: np.ndarray, enc_len: int, tok_off: np.ndarray, n: int) -> int:
    """Checksum of the first n strings of a batch result, comparable with OracleVocab.tokenize_batch_timed()."""
    t = np.ascontiguousarray(tokens_u8).view(np.uint8)
    o = np.ascontiguousarray(tok_off, dtype=np.uint64)
    return int(_api("port")["ids_checksum"](t.ctypes.data, enc_len, o.ctypes.data, n))


def incomplete_utf8(b: bytes) -> int:
    return _api("port")["incomplete_utf8"](b, len(b))


def incomplete_utf16(b: bytes) -> int:
    return _api("port")["incomplete_utf16"](b, len(b))

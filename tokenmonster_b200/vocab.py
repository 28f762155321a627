"""Host-side mirror of the reference interface for the hot path, on top of the C ABI.

Mirrors, with the same names / argument meaning / error behaviour,
  * python/tokenmonster.py:  load (:25), Vocab.tokenize (:410), tokenize_count (:485),
    decode (:341), __len__ (:270), capcode/charset/mode/normalization (:279-339),
    unk_token_id (:660), deserialize_tokens / serialize_tokens (:916-943)
  * go/tokenmonster.go:  Load (:2656), Vocab.Tokenize (:959), Count (:971),
    TokenizeToSerialized (:986), Decode (:788), DecodeSerialized (:803)
Every compute call goes through libtmgpu.so (include/tmgpu.h); nothing here tokenizes on
the CPU.  torch is used only to hand device buffers to the device-resident entry points.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Union

import numpy as np

from . import _lib
from ._lib import TmgpuError, check, lib

Text = Union[str, bytes, bytearray, memoryview]


class TokenizeResult:
    """Batch result: ids of string s are tokens[tok_off[s]:tok_off[s+1]].

    The arrays are zero-copy views of the pinned host buffers the library returned; they stay
    valid for the lifetime of this object (the buffers go back to the library's pool on __del__).
    """

    def __init__(self, res, n: int):
        self._res = res
        r = res.contents
        nt, enc = int(r.n_tokens), int(r.enc_len)
        self.enc_len = enc
        self.n_tokens = nt
        raw = np.ctypeslib.as_array(r.tokens, shape=(nt * enc,)) if nt else np.zeros(0, np.uint8)
        self.raw = raw                                   # little-endian enc_len-byte stream
        self.tok_off = np.ctypeslib.as_array(r.tok_off, shape=(n + 1,))
        self.missing = np.ctypeslib.as_array(r.missing, shape=(n,)) if n else np.zeros(0, np.uint32)
        self.status = np.ctypeslib.as_array(r.status, shape=(n,)) if n else np.zeros(0, np.uint8)
        if enc == 2:
            self.tokens = raw.view(np.uint16)
        elif enc == 4:
            self.tokens = raw.view(np.uint32)
        else:
            self.tokens = _widen24(raw)                  # 24-bit ids are widened (one copy)

    @property
    def serialized(self) -> bytes:
        """The TokenizeToSerialized byte stream (go/tokenmonster.go:986)."""
        return self.raw.tobytes()

    def __len__(self):
        return len(self.tok_off) - 1

    def __getitem__(self, s: int) -> np.ndarray:
        return self.tokens[int(self.tok_off[s]):int(self.tok_off[s + 1])]

    def close(self):
        if getattr(self, "_res", None) is not None:
            lib().tmgpu_result_free(self._res)
            self._res = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _flatten(texts: Sequence[Text]):
    parts = []
    for t in texts:
        if isinstance(t, str):
            t = t.encode("utf-8")
        elif not isinstance(t, (bytes, bytearray, memoryview)):
            raise ValueError("TokenMonster: Input to tokenize must be a string or a list of strings.")
        parts.append(bytes(t))
    off = np.zeros(len(parts) + 1, dtype=np.uint64)
    if parts:
        off[1:] = np.cumsum([len(p) for p in parts], dtype=np.uint64)
    return b"".join(parts), off


def _widen24(raw: np.ndarray) -> np.ndarray:
    b = raw.reshape(-1, 3).astype(np.uint32)
    return b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)


class Vocab:
    """A vocabulary resident on one B200 (the GPU counterpart of tokenmonster.Vocab)."""

    def __init__(self, file_bytes: bytes, device: int = 0):
        self._h = C.c_void_p()
        self.device = device
        buf = bytes(file_bytes)
        check(lib().tmgpu_vocab_create(buf, len(buf), device, C.byref(self._h)))
        info = _lib.VocabInfo()
        check(lib().tmgpu_vocab_info_get(self._h, C.byref(info)))
        self.info = info

    @classmethod
    def load(cls, path: str, device: int = 0) -> "Vocab":
        """go/tokenmonster.go:2656 Load / python tokenmonster.load: read a .vocab file."""
        try:
            with open(path, "rb") as f:
                data = f.read()
        except OSError as e:
            raise TmgpuError(12, f"cannot open vocabulary file: {e}")  # server status 12
        return cls(data, device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().tmgpu_vocab_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (go:2377-2597) ------------------------------------------------------
    def __len__(self) -> int:
        return int(self.info.vocab_size)

    def capcode(self) -> int:
        return int(self.info.capcode)

    def charset(self) -> str:
        return ("None", "UTF-8", "UTF-16")[self.info.charset]

    def mode(self) -> int:
        return int(self.info.level)

    def normalization_code(self) -> int:
        return int(self.info.normflag)

    def unk_token_id(self):
        return None if self.info.unk_token == _lib.DNE else int(self.info.unk_token)

    def max_token_length(self) -> int:
        return int(self.info.max_token_length)

    def highest_token_id(self) -> int:
        return int(self.info.n_reverse) - 1

    def default_encoding_length(self) -> int:
        return 2 if self.info.n_reverse <= 65536 else 3  # go:990-996

    # ---- tokenize --------------------------------------------------------------------
    def tokenize_batch(self, data, offsets, enc_len: int = 0) -> TokenizeResult:
        """Batched Vocab.TokenizeToSerialized over strings data[offsets[s]:offsets[s+1]]."""
        data = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(off) - 1
        if enc_len not in (0, 2, 3, 4):
            raise TmgpuError(1, "Invalid encoding length")
        res = C.POINTER(_lib.Result)()
        check(lib().tmgpu_tokenize_batch(self._h, data.ctypes.data if data.size else None, off.ctypes.data, n,
                                         enc_len, C.byref(res)))
        return TokenizeResult(res, n)

    def tokenize(self, text):
        """python/tokenmonster.py:410 -- a string -> ids; a list of strings -> list of id arrays."""
        single = isinstance(text, (str, bytes, bytearray, memoryview))
        texts = [text] if single else list(text)
        if single and len(texts[0]) == 0:
            return None  # the reference returns None for an empty string (py:431-432)
        data, off = _flatten(texts)
        r = self.tokenize_batch(data, off, 2 if self.info.n_reverse <= 65536 else 4)  # server:350-353
        bad = np.nonzero(r.status)[0]
        if bad.size:
            raise TmgpuError(13, "TokenMonster: Normalization failed")  # py:1082, server:379-381
        out = [r[s].astype(np.uint32) for s in range(len(r))]
        return out[0] if single else out

    def tokenize_count(self, text):
        """python/tokenmonster.py:485 / go Count :971 (the reference's Count value, quirk Q2)."""
        single = isinstance(text, (str, bytes, bytearray, memoryview))
        texts = [text] if single else list(text)
        data, off = _flatten(texts)
        counts, _, status = self.count_batch(data, off)
        if np.any(status):
            raise TmgpuError(13, "TokenMonster: Normalization failed")
        return int(counts[0]) if single else [int(c) for c in counts]

    def count_batch(self, data, offsets):
        data = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(off) - 1
        counts = np.zeros(n, np.uint64)
        missing = np.zeros(n, np.uint32)
        status = np.zeros(n, np.uint8)
        check(lib().tmgpu_count_batch(self._h, data.ctypes.data if data.size else None, off.ctypes.data, n,
                                      counts.ctypes.data, missing.ctypes.data, status.ctypes.data))
        return counts, missing, status

    # ---- decode ----------------------------------------------------------------------
    def decode_batch(self, tokens, tok_off, enc_len: int = 4):
        """Batched Vocab.DecodeSerialized: returns (bytes array uint8, offsets uint64[n+1]).

        tokens: bytes / uint8 array = the serialized enc_len-byte stream; any other integer array =
        ids, which are packed to enc_len bytes here."""
        if enc_len not in (2, 3, 4):
            raise TmgpuError(1, "Invalid encoding length")
        if isinstance(tokens, (bytes, bytearray, memoryview)):
            raw = np.frombuffer(tokens, dtype=np.uint8)        # already the serialized stream
        elif isinstance(tokens, np.ndarray) and tokens.dtype == np.uint8:
            raw = tokens                                        # serialized stream as a byte array
        else:
            t = np.ascontiguousarray(tokens)
            if enc_len == 2:
                raw = t.astype(np.uint16).view(np.uint8)
            elif enc_len == 4:
                raw = t.astype(np.uint32).view(np.uint8)
            else:
                t = t.astype(np.uint32)
                raw = np.stack([t & 0xFF, (t >> 8) & 0xFF, (t >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(-1)
        raw = np.ascontiguousarray(raw)
        off = np.ascontiguousarray(tok_off, dtype=np.uint64)
        n = len(off) - 1
        res = C.POINTER(_lib.Bytes)()
        check(lib().tmgpu_decode_batch(self._h, raw.ctypes.data if raw.size else None, off.ctypes.data, n, enc_len,
                                       C.byref(res)))
        try:
            b = res.contents
            nb = int(b.n_bytes)
            data = np.ctypeslib.as_array(b.bytes, shape=(nb,)).copy() if nb else np.zeros(0, np.uint8)
            boff = np.ctypeslib.as_array(b.off, shape=(n + 1,)).copy()
        finally:
            lib().tmgpu_bytes_free(res)
        return data, boff

    def score_batch(self, data, off):
        """The training worker's loop (training/trainvocab.go:932-1170) over a batch of data strips:
        returns (scores uint64[n_reverse], miss_hist uint64[256], tokens_in_text, status uint8[n])."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        scores = np.zeros(max(int(self.info.n_reverse), 1), np.uint64)
        hist = np.zeros(256, np.uint64)
        status = np.zeros(max(n, 1), np.uint8)
        total = C.c_uint64(0)
        check(lib().tmgpu_score_batch(self._h, data.ctypes.data if data.size else None, off.ctypes.data, n,
                                      scores.ctypes.data, hist.ctypes.data, C.byref(total), status.ctypes.data))
        return scores, hist, int(total.value), status[:n]

    def last_stats(self) -> dict:
        """What the latest tokenize / count call did: strings through the non-ASCII normalizer,
        chunks launched for long documents, chunked documents redone by one lane, kernel launches."""
        out = (C.c_uint32 * 4)()
        check(lib().tmgpu_last_stats(self._h, None, C.byref(out)))
        return dict(unicode_strings=int(out[0]), chunks=int(out[1]), chunk_redo=int(out[2]), launches=int(out[3]))

    # ---- server wire format ------------------------------------------------------------
    def _wire(self, fn, payload, *lead):
        buf = np.frombuffer(payload, dtype=np.uint8) if not isinstance(payload, np.ndarray) else payload
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        res = C.POINTER(_lib.WireReply)()
        check(fn(self._h, *lead, buf.ctypes.data if buf.size else None, buf.size, C.byref(res)))
        try:
            r = res.contents
            n = int(r.len)
            body = np.ctypeslib.as_array(r.payload, shape=(n,)).tobytes() if n else b""
            return int(r.status_code), body
        finally:
            lib().tmgpu_wire_reply_free(res)

    def wire_tokenize(self, payload):
        """Job 1 of tokenmonsterserver (training/tokenmonsterserver.go:339-394): the request payload
        python's Vocab.tokenize builds (py:410-483) -> (status_code, reply payload)."""
        return self._wire(lib().tmgpu_wire_tokenize, payload)

    def wire_count(self, payload):
        """Job 20 as python's tokenize_count sends it (py:485-552) -> (status_code, reply payload)."""
        return self._wire(lib().tmgpu_wire_count, payload)

    def wire_decode(self, payload, enc_len: int):
        """Jobs 2/3/4 (server:396-446; py:341-408) -> (status_code, reply payload)."""
        return self._wire(lib().tmgpu_wire_decode, payload, enc_len)

    def decode(self, tokens, as_bytes: bool = True):
        """python/tokenmonster.py:341 -- ids -> text; a list of id lists -> list of texts."""
        def is_seq_of_seq(x):
            return len(x) > 0 and not np.isscalar(x[0]) and not isinstance(x[0], (int, np.integer))
        if isinstance(tokens, np.ndarray) and tokens.ndim == 1:
            batch, single = [tokens], True
        elif isinstance(tokens, (list, tuple)) and is_seq_of_seq(tokens):
            batch, single = [np.asarray(t, dtype=np.uint32) for t in tokens], False
        else:
            batch, single = [np.asarray(tokens, dtype=np.uint32)], True
        off = np.zeros(len(batch) + 1, np.uint64)
        off[1:] = np.cumsum([len(t) for t in batch], dtype=np.uint64)
        flat = np.concatenate(batch) if batch else np.zeros(0, np.uint32)
        data, boff = self.decode_batch(flat.astype(np.uint32), off, 4)
        outs = [data[int(boff[s]):int(boff[s + 1])].tobytes() for s in range(len(batch))]
        if not as_bytes:
            outs = [o.decode("utf-8", "replace") for o in outs]
        return outs[0] if single else outs

    def decoder(self, n_streams: int = 1) -> "Decoder":
        """python/tokenmonster.py:273 Vocab.decoder() / go NewDecoder :450 -- for n parallel streams."""
        return Decoder(self, n_streams)

    # ---- (de)serialization helpers (py:916-943, go Deserialize :749-784) ----------------
    def deserialize_tokens(self, binary: bytes, enc_len: int = 0) -> np.ndarray:
        """Vocab.Deserialize (go:749-784) through the C ABI (tmgpu_deserialize)."""
        raw = np.frombuffer(bytes(binary), dtype=np.uint8)
        enc = enc_len or self.default_encoding_length()
        out = np.zeros(len(raw) // max(enc, 1) if enc in (2, 3, 4) else 0, np.uint32)
        n = lib().tmgpu_deserialize(self._h, raw.ctypes.data if raw.size else None, len(raw), enc_len,
                                    out.ctypes.data if out.size else None)
        if n < 0:
            raise TmgpuError(1, "Invalid encoding length")
        return out[:n]

    # ---- table self-check (host walk of the device trie; CPU-only CI) ---------------------
    def _trie_lookup(self, window: bytes, spaced: bool = False, shortcut: bool = False):
        """Table self-check: host walk of the device trie; shortcut = first steps through the
        root2 / sp_child tables, the way the lane-per-string kernel starts a walk."""
        idx = C.c_uint32()
        l = lib().tmgpu_selftest_trie_lookup(self._h, window, len(window), int(spaced) | (2 if shortcut else 0),
                                             C.byref(idx))
        return (int(idx.value), l) if l else (None, 0)


class Decoder:
    """A batch of independent streaming decoders (go/tokenmonster.go:57-61, :450-740).

    Each stream keeps, on the device, the incomplete UTF-8/16 tail of what it has decoded so far and
    the capcode decoder flags, exactly like one reference `Decoder` object.  With n_streams == 1 the
    methods take / return a single sequence, like python/tokenmonster.py:136."""

    def __init__(self, vocab: Vocab, n_streams: int = 1):
        self.vocab = vocab
        self.n = n_streams
        self._h = C.c_void_p()
        check(lib().tmgpu_decoder_create(vocab._h, n_streams, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().tmgpu_decoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _take(self, res):
        try:
            b = res.contents
            nb = int(b.n_bytes)
            data = np.ctypeslib.as_array(b.bytes, shape=(nb,)).copy() if nb else np.zeros(0, np.uint8)
            off = np.ctypeslib.as_array(b.off, shape=(self.n + 1,)).copy()
        finally:
            lib().tmgpu_bytes_free(res)
        return [data[int(off[s]):int(off[s + 1])].tobytes() for s in range(self.n)]

    def decode_batch(self, chunks) -> List[bytes]:
        """chunks[s] = the next ids of stream s (may be empty)."""
        if len(chunks) != self.n:
            raise ValueError("one chunk per stream")
        arrs = [np.asarray(c, dtype=np.uint32).reshape(-1) for c in chunks]
        off = np.zeros(self.n + 1, np.uint64)
        off[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
        flat = np.ascontiguousarray(np.concatenate(arrs) if arrs else np.zeros(0, np.uint32), dtype=np.uint32)
        res = C.POINTER(_lib.Bytes)()
        check(lib().tmgpu_decoder_decode(self._h, flat.ctypes.data if flat.size else None, off.ctypes.data, 4,
                                         C.byref(res)))
        return self._take(res)

    def decode(self, tokens):
        out = self.decode_batch([tokens] if self.n == 1 else tokens)
        return out[0] if self.n == 1 else out

    def flush(self):
        res = C.POINTER(_lib.Bytes)()
        check(lib().tmgpu_decoder_flush(self._h, C.byref(res)))
        out = self._take(res)
        return out[0] if self.n == 1 else out


def load(path: str, device: int = 0) -> Vocab:
    """tokenmonster.load(path) for a local .vocab file (python/tokenmonster.py:25)."""
    return Vocab.load(path, device)

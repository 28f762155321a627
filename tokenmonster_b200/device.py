"""Device-resident batch path: inputs already in HBM, outputs stay in HBM.

torch is plumbing only (device memory, streams): the tensors are handed to libtmgpu.so as raw
pointers and the library-owned outputs are wrapped back into torch tensors without a copy.
This is what bench.py times for the kernel/roofline figure and what the sharded multi-GPU
host (dist.py) runs on every rank.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib
from .vocab import Vocab


class _CAI:
    """Zero-copy view of library-owned device memory (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _view(ptr, count: int, typestr: str, device):
    if count == 0 or not ptr:
        dt = {"|u1": torch.uint8, "<i4": torch.int32, "<i8": torch.int64}[typestr]
        return torch.empty(0, dtype=dt, device=device)
    return torch.as_tensor(_CAI(ptr, (count,), typestr), device=device)


class DeviceWorkspace:
    """Owns the device output buffers for batches up to (max_strings, max_bytes)."""

    def __init__(self, vocab: Vocab, max_strings: int, max_bytes: int):
        self.vocab = vocab
        self.max_strings, self.max_bytes = int(max_strings), int(max_bytes)
        self.device = torch.device("cuda", vocab.device)
        self._h = C.c_void_p()
        check(lib().tmgpu_workspace_create(vocab._h, max_strings, max_bytes, C.byref(self._h)))
        self.view = _lib.DeviceView()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().tmgpu_workspace_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_tokenize(self, data: torch.Tensor, off: torch.Tensor, enc_len: int = 0, mode: int = 0, stream=None):
        """Enqueue the tokenize kernels on `stream` (default: torch's current stream); no sync."""
        assert data.dtype == torch.uint8 and off.dtype == torch.int64 and data.is_cuda and off.is_cuda
        n = off.numel() - 1
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        check(lib().tmgpu_tokenize_device(self._h, data.data_ptr(), off.data_ptr(), n, enc_len, mode, st,
                                          C.byref(self.view)))
        self._n = n
        self._enc = enc_len or self.vocab.default_encoding_length()
        return n

    def launches(self) -> int:
        return lib().tmgpu_last_launch_count(self._h)

    def kernel_times_ms(self):
        """CUDA-event durations of the tokenize kernel for the launches since the last call."""
        buf = (C.c_float * 64)()
        k = lib().tmgpu_workspace_kernel_times(self._h, buf, 64)
        return [float(buf[i]) for i in range(k)]

    def stage_times_ms(self):
        """CUDA-event durations of the last run's stages: prep + scan + sort, table, chain (, 0)."""
        buf = (C.c_float * 4)()
        k = lib().tmgpu_workspace_stage_times(self._h, buf)
        return [float(buf[i]) for i in range(k)]

    def results(self):
        """Synchronise and wrap the outputs of the last launch_tokenize (mode 0) as torch tensors."""
        torch.cuda.current_stream(self.device).synchronize()
        n = self._n
        tok_off = _view(self.view.tok_off, n + 1, "<i8", self.device)
        total = int(tok_off[-1].item()) if n >= 0 else 0
        tokens = _view(self.view.tokens, total * self._enc, "|u1", self.device)
        return {
            "tokens_u8": tokens,
            "tok_off": tok_off,
            "missing": _view(self.view.missing, n, "<i4", self.device),
            "status": _view(self.view.status, n, "|u1", self.device),
            "counts": _view(self.view.counts, n, "<i4", self.device),
            "n_tokens": total,
            "enc_len": self._enc,
        }

    def raw_views(self, n: int, token_bytes: int):
        """Byte views of the last launch's outputs WITHOUT synchronising (valid in stream order):
        the 8-byte id total, counts / missing as bytes, status, and the first `token_bytes` bytes of
        the dense id buffer (which may reach beyond the ids actually written)."""
        v = self.view
        return {
            "n_tokens_u8": _view(v.n_tokens, 8, "|u1", self.device),
            "counts_u8": _view(v.counts, 4 * n, "|u1", self.device),
            "missing_u8": _view(v.missing, 4 * n, "|u1", self.device),
            "status": _view(v.status, n, "|u1", self.device),
            "tokens_u8": _view(v.tokens, token_bytes, "|u1", self.device),
        }

    def decode(self, tokens_u8: torch.Tensor, tok_off: torch.Tensor, enc_len: int, stream=None):
        """Device decode; returns (bytes u8 tensor, offsets u64 tensor) views (synchronises once)."""
        n = tok_off.numel() - 1
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        d_bytes, d_off, d_total = C.c_void_p(), C.c_void_p(), C.c_void_p()
        off64 = tok_off.to(torch.int64) if tok_off.dtype != torch.int64 else tok_off
        check(lib().tmgpu_decode_device(self._h, tokens_u8.data_ptr(), off64.data_ptr(), n, enc_len, st,
                                        C.byref(d_bytes), C.byref(d_off), C.byref(d_total)))
        off = _view(d_off.value, n + 1, "<i8", self.device)
        total = int(off[-1].item())
        return _view(d_bytes.value, total, "|u1", self.device), off

"""tokenmonster_b200 -- B200-native batched tokenize / detokenize behind TokenMonster's API.

Package layout (hot path only):
  csrc/        sm_100a CUDA kernels + the C ABI (include/tmgpu.h) -> libtmgpu.so
  _lib.py      ctypes binding (fails loudly if the library is missing; no fallback)
  vocab.py     host-side mirror of the reference interface (load / tokenize / decode ...)
  device.py    device-resident batch path (torch tensors in, torch views out)
  dist.py      sharded multi-GPU host (one process per GPU)
"""
from ._lib import TmgpuError, lib  # noqa: F401
from .vocab import Decoder, TokenizeResult, Vocab, load  # noqa: F401

__all__ = ["Vocab", "Decoder", "load", "TokenizeResult", "TmgpuError", "lib"]

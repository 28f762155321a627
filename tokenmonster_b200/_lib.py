"""ctypes binding of libtmgpu.so (include/tmgpu.h).

The library is built in-tree by tokenmonster_b200/build.sh (nvcc, sm_100a).  There is no
fallback of any kind: if the shared object is missing, importing a compute entry raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libtmgpu.so")

DNE = 0xFFFFFF
S_OK = 0
S_NORMALIZE = 13
E_CUDA = 2

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class VocabInfo(C.Structure):
    _fields_ = [("capcode", C.c_uint8), ("charset", C.c_uint8), ("normflag", C.c_uint8), ("level", C.c_uint8),
                ("reserve", C.c_uint8), ("max_token_length", C.c_uint8), ("pad", C.c_uint8 * 2),
                ("unk_token", C.c_uint32), ("delete_token", C.c_uint32), ("vocab_size", C.c_uint32),
                ("n_reverse", C.c_uint32), ("n_info", C.c_uint32), ("n_trie_nodes", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("n", C.c_uint32), ("enc_len", C.c_uint8), ("n_tokens", C.c_uint64), ("tokens", u8p),
                ("tok_off", u64p), ("missing", u32p), ("status", u8p), ("priv", C.c_void_p)]


class Bytes(C.Structure):
    _fields_ = [("n", C.c_uint32), ("n_bytes", C.c_uint64), ("bytes", u8p), ("off", u64p), ("priv", C.c_void_p)]


class WireReply(C.Structure):
    _fields_ = [("status_code", C.c_uint8), ("len", C.c_uint64), ("payload", u8p), ("priv", C.c_void_p)]


class DeviceView(C.Structure):
    _fields_ = [("tokens", C.c_void_p), ("tok_off", C.c_void_p), ("missing", C.c_void_p), ("status", C.c_void_p),
                ("n_tokens", C.c_void_p), ("counts", C.c_void_p)]


# every symbol include/tmgpu.h declares, with its prototype
PROTOTYPES = {
    "tmgpu_last_error": (C.c_char_p, []),
    "tmgpu_device_count": (C.c_int, []),
    "tmgpu_version": (C.c_char_p, []),
    "tmgpu_vocab_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "tmgpu_vocab_load_file": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "tmgpu_vocab_destroy": (None, [C.c_void_p]),
    "tmgpu_vocab_info_get": (C.c_int, [C.c_void_p, C.POINTER(VocabInfo)]),
    "tmgpu_tokenize_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8,
                                       C.POINTER(C.POINTER(Result))]),
    "tmgpu_result_free": (None, [C.POINTER(Result)]),
    "tmgpu_count_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "tmgpu_score_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_uint64), C.c_void_p]),
    "tmgpu_decode_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8,
                                     C.POINTER(C.POINTER(Bytes))]),
    "tmgpu_bytes_free": (None, [C.POINTER(Bytes)]),
    "tmgpu_decoder_create": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "tmgpu_decoder_destroy": (None, [C.c_void_p]),
    "tmgpu_decoder_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint8, C.POINTER(C.POINTER(Bytes))]),
    "tmgpu_decoder_flush": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(Bytes))]),
    "tmgpu_wire_tokenize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.POINTER(WireReply))]),
    "tmgpu_wire_count": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.POINTER(WireReply))]),
    "tmgpu_wire_decode": (C.c_int, [C.c_void_p, C.c_uint8, C.c_void_p, C.c_uint64,
                                    C.POINTER(C.POINTER(WireReply))]),
    "tmgpu_wire_reply_free": (None, [C.POINTER(WireReply)]),
    "tmgpu_workspace_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]),
    "tmgpu_workspace_destroy": (None, [C.c_void_p]),
    "tmgpu_tokenize_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_int,
                                        C.c_void_p, C.POINTER(DeviceView)]),
    "tmgpu_last_launch_count": (C.c_int, [C.c_void_p]),
    "tmgpu_last_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32 * 4)]),
    "tmgpu_workspace_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "tmgpu_decode_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_void_p,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "tmgpu_deserialize": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint8, C.c_void_p]),
    "tmgpu_workspace_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "tmgpu_selftest_trie_lookup": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, u32p]),
    "tmgpu_selftest_unicode": (C.c_uint32, [C.c_uint32, u32p, u32p, u32p, C.POINTER(C.c_int)]),
}

_lib = None


class TmgpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"tmgpu error {code}: {message}")
        self.code = code


def lib():
    """The loaded library (raises if it has not been built -- never falls back)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise FileNotFoundError(
                f"{SO_PATH} is missing: run tokenmonster_b200/build.sh (or __graft_entry__.build()). "
                "tokenmonster_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise TmgpuError(rc, lib().tmgpu_last_error().decode("utf-8", "replace"))

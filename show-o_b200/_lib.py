"""ctypes binding of libshowo_b200.so (include/showo_b200.h).  No torch types cross this boundary: only raw
device/host pointers, sizes and a cudaStream_t.  There is no CPU fallback: if the library is missing or no sm_100
device is visible the compute entry points raise."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libshowo_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "showo_b200.h")


class ShowoError(RuntimeError):
    pass


class SeqMask(C.Structure):
    _fields_ = [("pad_end", C.c_int32), ("full_begin", C.c_int32), ("full_end", C.c_int32),
                ("win_begin", C.c_int32), ("win_end", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32),
                ("ffn", C.c_int32), ("rotary_dim", C.c_int32), ("max_pos", C.c_int32), ("ln_eps", C.c_float),
                ("rope_theta", C.c_float), ("llm_vocab_size", C.c_int32), ("num_new_special_tokens", C.c_int32),
                ("codebook_size", C.c_int32)]


class ClipConfig(C.Structure):
    _fields_ = [("image_size", C.c_int32), ("patch_size", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32),
                ("n_heads", C.c_int32), ("ffn", C.c_int32), ("ln_eps", C.c_float)]


_P = C.c_void_p
_I = C.c_int
_I64 = C.c_int64
_F = C.c_float

_PROTOS = {
    "showo_last_error": (C.c_char_p, []),
    "showo_abi_version": (_I, []),
    "showo_device_count": (_I, []),
    "showo_engine_create": (_I, [C.POINTER(Config), _I, C.POINTER(_P)]),
    "showo_engine_destroy": (_I, [_P]),
    "showo_load_weight": (_I, [_P, C.c_char_p, _P, _I64, _I]),
    "showo_weights_complete": (_I, [_P]),
    "showo_forward": (_I, [_P, _P, _P, _I, _I, C.POINTER(SeqMask), _P, _P]),
    "showo_forward_fp32": (_I, [_P, _P, _P, _I, _I, C.POINTER(SeqMask), _P, _P]),
    "showo_t2i_logits": (_I, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(SeqMask), _P, _P]),
    "showo_t2i_generate": (_I, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(SeqMask), _I, _F, C.POINTER(C.c_int32),
                                C.POINTER(C.c_float), _P, _P, C.c_uint64, _P, _P]),
    "showo_sampler_step": (_I, [_P, _P, _I, _I, _I, _F, _P, _I64, _I, _I, _I, _I, _F, _P, _P, C.c_uint64,
                                C.c_uint32, _P, _P, _P]),
    "showo_mmu_generate": (_I, [_P, _P, _P, _I, _I, C.POINTER(SeqMask), _I, _I, _F, _I64, C.c_uint64, _P, _P, _P, _P]),
    "showo_cross_entropy": (_I, [_P, _P, _I64, _I, _I, _I, _I, _I, _I, _I64, _P, _P]),
    "showo_mmu_sample": (_I, [_P, _I64, _I, _I, _F, _I, _P, C.c_uint64, C.c_uint32, _P, _P]),
    "showo_mask_descriptors": (_I, [_P, _I, _I, _I, _I64, C.POINTER(SeqMask), C.POINTER(C.c_int32), _P]),
    "showo_train_forward": (_I, [_P, _P, _P, _I, _I, C.POINTER(SeqMask), _P, C.POINTER(C.c_int32), _I64, _P, _P, _P]),
    "showo_backward": (_I, [_P, _P, _P, _P]),
    "showo_backward_phase": (_I, [_P, _I, _P, _P, _P]),
    "showo_grad_buffer": (_I, [_P, C.POINTER(_P), C.POINTER(_I64)]),
    "showo_grad_range": (_I, [_P, _I, C.POINTER(_I64), C.POINTER(_I64)]),
    "showo_optimizer_enable": (_I, [_P]),
    "showo_adamw_step": (_I, [_P, _F, _F, _F, _F, _F, _P]),
    "showo_read_param": (_I, [_P, C.c_char_p, _P, _I64, _P]),
    "showo_read_grad": (_I, [_P, C.c_char_p, _P, _I64, _P]),
    "showo_attention_bwd_test": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, C.POINTER(SeqMask), _P]),
    "showo_t2i_train_prep": (_I, [_P, _I, _I, _P, _P, _I64, _I, C.POINTER(C.c_int64), _F, _F, _I, _F, _P, _P, _P, C.c_uint64, _I,
                                  _P, _P, _P, _P, _P, _P, _P, _P]),
    "showo_mm_projector": (_I, [_P, _P, _I64, _P, _P]),
    "showo_mm_projector_backward": (_I, [_P, _P, _I64, _P]),
    "showo_mm_projector_grad_buffer": (_I, [_P, C.POINTER(_P), C.POINTER(_I64)]),
    "showo_embed_tokens": (_I, [_P, _P, _I64, _P, _P]),
    "showo_set_rng_row_base": (_I, [_P, _I64]),
    "showo_kernel_launches": (_I64, [_P]),
    "magvit_engine_create": (_I, [_I, C.POINTER(_P)]),
    "magvit_engine_destroy": (_I, [_P]),
    "magvit_load_weight": (_I, [_P, C.c_char_p, _P, _I64, _I]),
    "magvit_weights_complete": (_I, [_P]),
    "magvit_decode_code": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "magvit_decode_code_u8": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "magvit_get_code": (_I, [_P, _P, _I, _I, _P, _P]),
    "magvit_decode_code_fp32": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "magvit_get_code_fp32": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "magvit_kernel_launches": (_I64, [_P]),
    "clip_engine_create": (_I, [C.POINTER(ClipConfig), _I, C.POINTER(_P)]),
    "clip_engine_destroy": (_I, [_P]),
    "clip_load_weight": (_I, [_P, C.c_char_p, _P, _I64, _I]),
    "clip_weights_complete": (_I, [_P]),
    "clip_forward": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "clip_kernel_launches": (_I64, [_P]),
    "showo_gemm_bf16": (_I, [_P, _I64, _P, _I64, _I, _I, _I, _P, _I64, _P, _P, _I64, _I, _I, _I, _P]),
    "showo_attention_test": (_I, [_P, _I64, _I, _I, _I, _I, _P, _P, _P, _P, _F, _F, _I, _P, _P, _I, _I,
                                  C.POINTER(SeqMask), _P]),
    "showo_attention_run": (_I, [_P, _I64, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I64, _P]),
    "showo_layernorm_test": (_I, [_P, _P, _P, _F, _P, _I, _I, _P]),
    "showo_conv_test": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function include/showo_b200.h declares (used by the CPU symbol-export test)."""
    src = open(HEADER_PATH).read()
    return re.findall(r"^SHOWO_API\s+[\w\s\*]+?\b(\w+)\(", src, flags=re.M)


def load(build_if_missing: bool = True):
    """dlopen the library (building it in-tree first if needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise ShowoError(f"{LIB_PATH} is missing: run `python show-o_b200/build.py` (no CPU fallback exists)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().showo_last_error()
        raise ShowoError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def require_gpu():
    lib = load()
    if lib.showo_device_count() <= 0:
        raise ShowoError("no sm_100 (B200) CUDA device visible: show-o_b200 has no CPU fallback")
    return lib


def masks_array(descs) -> C.Array:
    """descs: iterable of 5-tuples (pad_end, full_begin, full_end, win_begin, win_end)."""
    descs = list(descs)
    arr = (SeqMask * len(descs))()
    for i, d in enumerate(descs):
        arr[i] = SeqMask(*[int(v) for v in d])
    return arr


def ptr(t):
    """device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)

"""ctypes binding of the C ABI in include/attnshift_b200.h.

PyTorch is plumbing only: tensors own the device memory, ``data_ptr()`` and the current
CUDA stream cross the boundary as plain pointers.  There is no fallback: if the shared
library is missing, ``load()`` raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libattnshift_b200.so')

_vp, _i, _ll, _f, _d, _sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double,
                             ctypes.c_size_t)

# name -> (restype, argtypes); MUST mirror include/attnshift_b200.h (tests/test_abi.py checks the symbol list)
SIGNATURES = {
    'as_linear_f16': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'as_qkv_proj_f16': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'as_layernorm_f16': (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    'as_patch_im2col_f16': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'as_assemble_tokens': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'as_mhsa_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'as_attn_headmean': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    'as_mean_shift_workspace': (_sz, [_i, _i, _i, _i, _i]),
    'as_grid_seeds': (_i, [_vp, _f, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'as_mean_shift': (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _sz, _vp]),
}

_lib = None


class AttnShiftError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AttnShiftError(
                f'{LIB_PATH} not found -- build it with `python -m attentionshift_b200.build` '
                '(there is no CPU / PyTorch fallback for the hot path)')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'C ABI takes contiguous device tensors'
    return ctypes.c_void_p(t.data_ptr())


def check(rc, what):
    if rc != 0:
        msg = {10001: 'bad argument', 10002: 'CUDA driver entry point unavailable', 10003: 'TMA descriptor encode failed'}
        raise AttnShiftError(f'{what} failed with code {rc} ({msg.get(rc, "cudaError_t")})')

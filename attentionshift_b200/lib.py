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
    'as_linear_tn_f16': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    'as_qkv_proj_f16': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'as_layernorm_f16': (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    'as_patch_im2col_f16': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'as_assemble_tokens': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'as_mhsa_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'as_mhsa_bwd': (_i, [_vp] * 13 + [_i, _i, _i, _i, _vp]),
    'as_mhsa_bwd_ex': (_i, [_vp] * 14 + [_i, _i, _i, _i, _vp]),
    'as_colsum_workspace': (_sz, [_i]),
    'as_colsum': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    'as_gelu_bwd_f16': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    'as_layernorm_bwd_workspace': (_sz, [_i]),
    'as_layernorm_bwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    'as_attn_bwd_prep': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    'as_transpose_pad_f16': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'as_mhsa_small': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'as_mhsa_set_variant': (_i, [_i]),
    'as_mt19937_draws': (_i, [_vp, _i, _i, _vp]),
    'as_timer_slots': (_i, []),
    'as_timer_record': (_i, [_i, _i, _vp]),
    'as_timer_elapsed': (_i, [_i, _vp]),
    'as_attn_headmean': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _f, _i, _i, _i, _vp]),
    'as_attn_headmean_ex': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _f, _i, _i, _i, _i, _vp]),
    'as_bgemm_f16_f32': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _ll, _f, _vp]),
    'as_rollout_tc_workspace': (_sz, [_i, _i, _i]),
    'as_rollout_rows_tc': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _vp, _sz, _vp]),
    'as_mean_shift_workspace': (_sz, [_i, _i, _i, _i, _i]),
    'as_grid_seeds': (_i, [_vp, _f, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'as_mean_shift': (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _sz, _vp]),
    'as_mean_shift_tc_workspace': (_sz, [_i, _i, _i, _i, _i, _i]),
    'as_mean_shift_tc': (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _sz, _vp]),
    'as_mean_shift_fused_workspace': (_sz, [_i, _i, _i]),
    'as_mean_shift_fused_debug': (None, [_vp]),
    'as_mean_shift_fused_occupancy': (_i, []),
    'as_mean_shift_fused': (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _sz, _vp]),
    'as_mean_shift_v2_supported': (_i, [_i, _i, _i, _i]),
    'as_mean_shift_v2_workspace': (_sz, [_i, _i, _i, _i, _i]),
    'as_mean_shift_v2_debug': (None, [_vp]),
    'as_mean_shift_v2_occupancy': (_i, [_vp, _vp, _vp]),
    'as_mean_shift_v2': (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _sz, _vp]),
    'as_roi_align_tokens': (_i, [_vp, _ll, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    'as_hungarian_points': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'as_cosine_maps_workspace': (_sz, [_i, _i, _i, _i, _i]),
    'as_cosine_maps': (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp, _sz, _vp]),
    'as_rollout_workspace': (_sz, [_i, _i, _i]),
    'as_rollout_rows': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    'as_cam_gather': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    'as_cam_minmax': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    'as_cam_bbox_workspace': (_sz, [_i, _i, _i]),
    'as_cam_bbox': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    'as_norm_rowcount': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    'as_norm_select': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'as_seed_proto': (_i, [_vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    'as_refine_threshold': (_i, [_vp, _i, _i, _f, _vp, _vp]),
    'as_weighted_centroid_workspace': (_sz, [_i, _i, _i, _i]),
    'as_weighted_centroid': (_i, [_vp, _ll, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    'as_refine_select': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    'as_fuse_instance_maps': (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    'as_mask_candidates_workspace': (_sz, [_i, _i, _i]),
    'as_mask_candidates': (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    'as_mask_select': (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'as_erode_downsample': (_i, [_vp, _i, _i, _i, _f, _i, _vp, _vp, _vp]),
    'as_filter_seeds': (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    'as_merge_prototypes': (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    'as_part_centers': (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class _Timers:
    """Optional per-entry-point device timing (bench.py).  Every C-ABI call is bracketed by one of the library's timing slots
    (as_timer_record) on the launching stream.  Calls made while the stream is being captured keep their slot for good -- the
    replayed graph re-records it every step -- while the slots of live calls are recycled at every ``begin_step()``.
    ``summary()`` adds up the last step: device time per entry point = the kernels that call enqueued."""

    def __init__(self):
        self.on = False
        self.captured = []          # (name, slot) of calls that live inside a CUDA graph
        self.live = []              # (name, slot) of this step's eager calls
        self.next_captured = 0
        self.next_live = 0

    def enable(self):
        self.on = True
        self.captured, self.live = [], []
        self.next_captured = 0
        load()
        self.n_slots = _lib._cdll.as_timer_slots()
        self.next_live = self.n_slots // 2

    def disable(self):
        self.on = False

    def begin_step(self):
        self.live = []
        self.next_live = self.n_slots // 2

    def wrap(self, name, fn):
        def call(*a):
            if not self.on or name.startswith('as_timer') or name == 'as_mt19937_draws':
                return fn(*a)
            capturing = torch.cuda.is_current_stream_capturing()
            if capturing:
                slot = self.next_captured
                self.next_captured += 1
                assert slot < self.n_slots // 2, 'timing slots exhausted'
                self.captured.append((name, slot))
            else:
                slot = self.next_live
                self.next_live += 1
                if slot >= self.n_slots:
                    return fn(*a)
                self.live.append((name, slot))
            sp = stream_ptr()
            _lib._cdll.as_timer_record(slot, 0, sp)
            r = fn(*a)
            _lib._cdll.as_timer_record(slot, 1, sp)
            return r
        return call

    def summary(self):
        """{entry point: dict(ms = device time in the last step, n = calls)}."""
        torch.cuda.synchronize()
        out = {}
        ms = ctypes.c_float()
        for name, slot in self.captured + self.live:
            if _lib._cdll.as_timer_elapsed(slot, ctypes.byref(ms)) == 0:
                d = out.setdefault(name, dict(ms=0.0, n=0))
                d['ms'] += ms.value
                d['n'] += 1
        return out


TIMERS = _Timers()


class _Lib:
    pass


class AttnShiftError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AttnShiftError(
                f'{LIB_PATH} not found -- build it with `python -m attentionshift_b200.build` '
                '(there is no CPU / PyTorch fallback for the hot path)')
        cdll = ctypes.CDLL(LIB_PATH)
        lib = _Lib()
        lib._cdll = cdll
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(cdll, name)            # AttributeError here = header / library drift
            fn.restype = res
            fn.argtypes = args
            setattr(lib, name, fn if name.endswith(('_workspace', '_supported', '_debug', '_occupancy')) else TIMERS.wrap(name, fn))
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

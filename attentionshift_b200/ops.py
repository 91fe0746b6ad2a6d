"""Thin tensor-level wrappers over the C ABI (one python function per exported kernel family).

Every function takes/returns CUDA tensors, launches on the current stream and never
synchronises.  Shapes follow the reference (RH = stdroi_point_deform_attn_reppoints.py,
VT = models/vision_transformer.py)."""
import os

import torch

from . import lib as _l

EPI_F16, EPI_GELU_F16, EPI_RESID_F32, EPI_F32 = 0, 1, 2, 4
TIMERS = _l.TIMERS


def linear_f16(x, w, bias=None, mode=EPI_F16, resid=None):
    """y = x @ w.T (+bias) on tcgen05 tensor cores.  x [M,K] fp16, w [N,K] fp16, bias [N] fp32.
    mode: EPI_F16 -> fp16, EPI_GELU_F16 -> gelu -> fp16, EPI_RESID_F32 -> fp32 resid + y, EPI_F32 -> fp32."""
    L = _l.load()
    M, K = x.shape
    N = w.shape[0]
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and w.shape[1] == K
    out = torch.empty(M, N, device=x.device, dtype=torch.float16 if mode in (EPI_F16, EPI_GELU_F16) else torch.float32)
    _l.check(L.as_linear_f16(_l.ptr(x), _l.ptr(w), _l.ptr(bias), _l.ptr(out), _l.ptr(resid), M, N, K, mode,
                             _l.stream_ptr()), 'as_linear_f16')
    return out


def linear_tn_f16(a, b):
    """a [R,M], b [R,N] fp16 -> a^T b [M,N] fp32 (the weight gradient dY^T X) without transposed copies of the operands."""
    L = _l.load()
    R, M = a.shape
    N = b.shape[1]
    assert a.dtype == torch.float16 and b.dtype == torch.float16 and b.shape[0] == R
    out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _l.check(L.as_linear_tn_f16(_l.ptr(a), _l.ptr(b), _l.ptr(out), R, M, N, _l.stream_ptr()), 'as_linear_tn_f16')
    return out


def qkv_proj(x, w, bias, B, T, heads, Tpad):
    """VT:76 qkv Linear + head split.  x [B*T,C] fp16 -> q,k [B,h,T,64] fp16, vT [B,h,64,Tpad] fp16 (zero padded)."""
    L = _l.load()
    q = torch.empty(B, heads, T, 64, device=x.device, dtype=torch.float16)
    k = torch.empty_like(q)
    vt = torch.empty(B, heads, 64, Tpad, device=x.device, dtype=torch.float16)
    if Tpad > T:
        vt[..., T:].zero_()             # only the padding columns: P V multiplies them (by zero probabilities), NaNs must not sit there
    _l.check(L.as_qkv_proj_f16(_l.ptr(x), _l.ptr(w), _l.ptr(bias), _l.ptr(q), _l.ptr(k), _l.ptr(vt), B, T, Tpad, heads,
                               _l.stream_ptr()), 'as_qkv_proj_f16')
    return q, k, vt


def layernorm_f16(x, gamma, beta, eps=1e-6):
    """VT:110/114 LayerNorm (eps 1e-6, VT:146) fused with the fp16 cast of the next GEMM operand.  x [M,C] fp32."""
    L = _l.load()
    M, C = x.shape
    y = torch.empty(M, C, device=x.device, dtype=torch.float16)
    _l.check(L.as_layernorm_f16(_l.ptr(x), _l.ptr(gamma), _l.ptr(beta), _l.ptr(y), M, C, float(eps), _l.stream_ptr()),
             'as_layernorm_f16')
    return y


def patch_im2col_f16(img):
    """img [B,3,H,W] fp32 -> [B*hp*wp, 768] fp16 patch rows (the 16x16/16 conv of VT:136 becomes a GEMM)."""
    L = _l.load()
    B, _, H, W = img.shape
    cols = torch.empty(B * (H // 16) * (W // 16), 768, device=img.device, dtype=torch.float16)
    _l.check(L.as_patch_im2col_f16(_l.ptr(img), _l.ptr(cols), B, H, W, _l.stream_ptr()), 'as_patch_im2col_f16')
    return cols


def assemble_tokens(emb, cls, pos, ptok, B, N):
    """VTD:203-213: [cls+pos0 | patches+pos | point tokens].  emb [B*N,C], cls [C], pos [1+N,C], ptok [Tp,C] -> [B,T,C]."""
    L = _l.load()
    C = emb.shape[1]
    Tp = ptok.shape[0]
    x = torch.empty(B, 1 + N + Tp, C, device=emb.device, dtype=torch.float32)
    _l.check(L.as_assemble_tokens(_l.ptr(emb), _l.ptr(cls), _l.ptr(pos), _l.ptr(ptok), _l.ptr(x), B, N, Tp, C,
                                  _l.stream_ptr()), 'as_assemble_tokens')
    return x


def mhsa_fwd(q, k, vt, T):
    """VT:79-83 without materialising attn.  q,k [B,h,T,64], vt [B,h,64,Tpad] fp16
    -> (o [B,T,h*64] fp16, m [B,h,T], l [B,h,T] fp32 softmax row statistics, log2 domain)."""
    L = _l.load()
    B, heads = q.shape[0], q.shape[1]
    Tpad = vt.shape[3]
    o = torch.empty(B, T, heads * 64, device=q.device, dtype=torch.float16)
    m = torch.empty(B, heads, T, device=q.device, dtype=torch.float32)
    l = torch.empty_like(m)
    _l.check(L.as_mhsa_fwd(_l.ptr(q), _l.ptr(k), _l.ptr(vt), _l.ptr(o), _l.ptr(m), _l.ptr(l), B, T, Tpad, heads,
                           _l.stream_ptr()), 'as_mhsa_fwd')
    return o, m, l


T_SCALE = 16384.0     # 2^14: head-mean probabilities (<= 1) as split fp16 without touching the subnormal range


HEADMEAN_SLICES = 1 if os.environ.get('AS_HEADMEAN_VARIANT', '2') == '1' else 4


def attn_headmean(q, k, m, l, T, want_rowsum=True, want_transposed=True, slices=None, want_map=True, row0=0):
    """VTD:236/242 attn.mean(1) recomputed from (q, k, m, l).  -> mean [B,T,T] (view of a row-padded buffer).  The
    tensor carries what the roll-out slab needs as attributes: ``_as_rowsum_part`` [B,T,slices*ceil(T/128)] and, when
    ``want_transposed``, ``_as_t16`` = (hi, lo) split-fp16 transposed maps [B,Tpad,Tpad] scaled by T_SCALE.
    ``slices`` = 4 (default): persistent kernel; 1: one CTA per tile (env AS_HEADMEAN_VARIANT=1).
    Roll-out-only production (persistent kernel): ``want_map=False`` skips the fp32 map (the returned tensor is an EMPTY
    placeholder that only carries the attributes plus ``_as_shape`` = (B, T, ld)); ``row0`` > 0 computes only the query tiles
    holding rows [row0, T) -- the rest of the returned map is uninitialised memory (``_as_valid_from`` = first valid row)."""
    L = _l.load()
    slices = HEADMEAN_SLICES if slices is None else slices
    B, heads = q.shape[0], q.shape[1]
    ld = (T + 127) // 128 * 128
    nt = (T + 127) // 128
    lean = (not want_map) or row0 > 0
    if lean and slices != 4:
        raise ValueError('roll-out-only head-mean production needs the persistent schedule (slices=4)')
    if row0 > 0 and want_transposed:
        raise ValueError('row0 > 0 produces a partial map: it cannot come with the transposed GEMM operand')
    buf = torch.empty(B, T, ld, device=q.device, dtype=torch.float32) if want_map else None
    part = torch.empty(B, T, slices * nt, device=q.device, dtype=torch.float32) if want_rowsum else None
    thi = torch.empty(B, ld, ld, device=q.device, dtype=torch.float16) if want_transposed else None
    tlo = torch.empty(B, ld, ld, device=q.device, dtype=torch.float16) if want_transposed else None
    if lean:
        _l.check(L.as_attn_headmean_ex(_l.ptr(q), _l.ptr(k), _l.ptr(m), _l.ptr(l), _l.ptr(buf), ld, _l.ptr(part), slices, _l.ptr(thi),
                                       _l.ptr(tlo), ld, T_SCALE, B, T, heads, int(row0), _l.stream_ptr()), 'as_attn_headmean_ex')
    else:
        _l.check(L.as_attn_headmean(_l.ptr(q), _l.ptr(k), _l.ptr(m), _l.ptr(l), _l.ptr(buf), ld, _l.ptr(part), slices, _l.ptr(thi),
                                    _l.ptr(tlo), ld, T_SCALE, B, T, heads, _l.stream_ptr()), 'as_attn_headmean')
    out = buf[:, :, :T] if want_map else torch.empty(0, device=q.device, dtype=torch.float32)
    out._as_rowsum_part = part
    out._as_shape = (B, T, ld)
    out._as_valid_from = (int(row0) // 128) * 128
    if want_transposed:
        out._as_t16 = (thi, tlo)
    return out, part


def _feat_args(feats):
    """feats [n_img, N, C] fp32, last two dims contiguous (image stride free)."""
    assert feats.dtype == torch.float32 and feats.stride(2) == 1 and feats.stride(1) == feats.shape[2]
    return feats.stride(0)


def grid_seeds(maps, feats, obj_img, rois, wp, S, thr=0.35):
    """RH:1786-1810.  maps [n_tot,N] fp32, feats [n_img,N,C], obj_img [n_tot] int32, rois [n_tot,4] fp32.
    -> (seed token ids [n_tot,S] int32, prototypes [n_tot,S,C] fp32)."""
    L = _l.load()
    n_tot, N = maps.shape
    C = feats.shape[2]
    tok = torch.empty(n_tot, S, device=maps.device, dtype=torch.int32)
    proto = torch.empty(n_tot, S, C, device=maps.device, dtype=torch.float32)
    _l.check(L.as_grid_seeds(_l.ptr(maps), float(thr), ctypes_ptr(feats), _feat_args(feats), _l.ptr(obj_img), _l.ptr(rois),
                             n_tot, N, C, wp, S, _l.ptr(tok), _l.ptr(proto), _l.stream_ptr()), 'as_grid_seeds')
    return tok, proto


def ctypes_ptr(t):
    import ctypes
    assert t.is_cuda
    return ctypes.c_void_p(t.data_ptr())


_SMS = {}


def _num_sms(dev):
    i = dev.index if dev.index is not None else torch.cuda.current_device()
    if i not in _SMS:
        _SMS[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _SMS[i]


_GROUPS = {}


def _group_index(n_per_img, dev):
    """(first instance, instance count) per image as int32 device tensors, cached per batch composition (read-only)."""
    key = (tuple(n_per_img), str(dev))
    hit = _GROUPS.get(key)
    if hit is None:
        if len(_GROUPS) > 64:
            _GROUPS.clear()
        first = [0]
        for k in list(n_per_img)[:-1]:
            first.append(first[-1] + k)
        hit = (torch.tensor(first, dtype=torch.int32).to(dev), torch.tensor(list(n_per_img), dtype=torch.int32).to(dev))
        _GROUPS[key] = hit
    return hit


def mean_shift(proto, feats, obj_img, rois, hp, wp, n_shift, tau=0.1, temp=0.1, clamp0=True, want_trace=False,
               n_per_img=None, use_tensor_cores=True, impl=None):
    """RH:830-854 + RH:882-908 on device.  proto [n_tot,S,C] (consumed), feats [n_img,N,C]; instances grouped by image.
    -> (proto [n_tot,S,C], sim [n_tot,S,N], trace [n_shift,n_tot,N] int32 or None).
    impl: 'v2' (one persistent cooperative kernel, two CTAs per SM; C % 128 == 0, <= 256 seed columns and <= 16 instances per
    image), 'fused' (the round-1 persistent kernel: C <= 768, <= 64 seed columns, <= 8 instances), 'tc' (batched split-fp16 affinity
    GEMM + small kernels), 'fp32' (CUDA-core kernels, any C); None picks the first that fits."""
    L = _l.load()
    n_tot, S, C = proto.shape
    n_img, N, _ = feats.shape
    dev = proto.device
    proto = proto.contiguous().clone()
    sim = torch.empty(n_tot, S, N, device=dev, dtype=torch.float32)
    trace = torch.empty(n_shift, n_tot, N, device=dev, dtype=torch.int32) if want_trace else None
    if n_per_img is None:
        n_per_img = torch.bincount(obj_img.long(), minlength=n_img).cpu().tolist()      # host sync: pass n_per_img to avoid it
    kmax = max(n_per_img) * S
    max_obj = max(n_per_img)
    if impl is None:
        if not use_tensor_cores or C % 64 != 0:
            impl = 'fp32'
        elif C % 128 == 0 and C <= 768 and kmax <= 64 and max_obj <= 8 and (N + 255) // 256 <= _num_sms(dev):
            impl = 'fused'          # <= 64 seed columns, ViT-S/B width: 256 tokens per CTA amortise the seed tiles best (0.37 vs 0.42 ms at cfg2)
        elif L.as_mean_shift_v2_supported(N, C, kmax, max_obj) and S <= 127 and \
                ((N + 63) // 64 + 1) // 2 <= (2 if kmax <= 64 and max_obj <= 8 else 1) * _num_sms(dev):
            impl = 'v2'             # everything else the persistent design covers: <= 256 seed columns, <= 16 instances, C <= 1024
        elif kmax <= 256 and S * C * 4 <= 200 * 1024:
            impl = 'tc'
        else:
            impl = 'fp32'
    if impl == 'v2':
        d_first, d_nobj = _group_index(n_per_img, dev)
        nbytes = L.as_mean_shift_v2_workspace(n_img, N, C, kmax, max_obj)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _l.check(L.as_mean_shift_v2(ctypes_ptr(feats), _feat_args(feats), n_img, N, C, hp, wp, _l.ptr(d_first), _l.ptr(d_nobj),
                                    kmax, max_obj, _l.ptr(rois), n_tot, S, _l.ptr(proto), _l.ptr(sim), n_shift, float(tau),
                                    float(temp), int(clamp0), _l.ptr(trace), _l.ptr(ws), nbytes, _l.stream_ptr()), 'as_mean_shift_v2')
        return proto, sim, trace
    if impl in ('fused', 'tc'):
        d_first, d_nobj = _group_index(n_per_img, dev)
        if impl == 'fused':
            nbytes = L.as_mean_shift_fused_workspace(n_img, N, C)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            _l.check(L.as_mean_shift_fused(ctypes_ptr(feats), _feat_args(feats), n_img, N, C, hp, wp, _l.ptr(obj_img),
                                           _l.ptr(d_first), _l.ptr(d_nobj), kmax, _l.ptr(rois), n_tot, S, _l.ptr(proto),
                                           _l.ptr(sim), n_shift, float(tau), float(temp), int(clamp0), _l.ptr(trace),
                                           _l.ptr(ws), nbytes, _l.stream_ptr()), 'as_mean_shift_fused')
            return proto, sim, trace
        nbytes = L.as_mean_shift_tc_workspace(n_img, n_tot, S, N, C, kmax)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _l.check(L.as_mean_shift_tc(ctypes_ptr(feats), _feat_args(feats), n_img, N, C, hp, wp, _l.ptr(obj_img), _l.ptr(d_first),
                                    _l.ptr(d_nobj), kmax, _l.ptr(rois), n_tot, S, _l.ptr(proto), _l.ptr(sim), n_shift, float(tau),
                                    float(temp), int(clamp0), _l.ptr(trace), _l.ptr(ws), nbytes, _l.stream_ptr()), 'as_mean_shift_tc')
        return proto, sim, trace
    nbytes = L.as_mean_shift_workspace(n_img, n_tot, S, N, C)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    _l.check(L.as_mean_shift(ctypes_ptr(feats), _feat_args(feats), n_img, N, C, hp, wp, _l.ptr(obj_img), _l.ptr(rois),
                             n_tot, S, _l.ptr(proto), _l.ptr(sim), n_shift, float(tau), float(temp), int(clamp0),
                             _l.ptr(trace), _l.ptr(ws), nbytes, _l.stream_ptr()), 'as_mean_shift')
    return proto, sim, trace

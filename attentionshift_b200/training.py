"""Backward of the ViT block on the device kernels (what DDP training of the backbone needs).

The reference trains ``VisionTransformerDet`` through torch autograd over its unfused block (VT:109-124) under apex AMP O1
(fp16 GEMMs with loss scaling, mmdet/apis/train.py:83, mmdet/utils/optimizer.py:23-38).  Here the two contraction-heavy parts
have hand-written backward passes behind ``torch.autograd.Function``:

  * ``AttentionFn``  qkv projection + scaled-dot-product attention.  Forward = the inference kernels (tcgen05 GEMM with the
                     head-split epilogue, flash attention); backward = ``as_mhsa_bwd`` (tcgen05: dK / dV and dQ kernels that
                     recompute P from the saved (m, l) row statistics -- nothing of size T x T is kept) followed by the
                     projection's dX / dW GEMMs on the same tcgen05 GEMM kernel.
  * ``LinearFn``     proj / fc1 / fc2: forward, dX and dW all on ``as_linear_f16``.

``BlockFn`` (the default of ``block_forward``) makes a whole block one autograd node: LayerNorm backward (with the residual
gradient folded in), GELU backward + bias gradients, casts and the attention backward's operand preparation run on the kernels
of ``csrc/vit_train.cu``; only the forward GELU is a torch op.  The per-op functions below remain for the patch embedding and as
the unfused comparison path (``fused=False``).  Like the reference's AMP path the
gradients that enter a GEMM are rounded to fp16: callers scale the loss (``apex.amp.scale_loss`` in the reference) when their
gradients are small.  Layout shuffles between the head-major attention tensors and the token-major GEMM operands (transposes,
zero padding of the reduction dimension to a multiple of 64) are torch copies.
"""
import os

import torch
from torch.autograd.function import once_differentiable

from . import lib as _l
from . import ops


def transpose_pad(x, Rp):
    """[batch..., R, C] fp16 -> [batch..., C, Rp] fp16, zero padded for rows >= R (``as_transpose_pad_f16``: the K-major operand
    layout of the backward GEMMs and of the attention backward)."""
    L = _l.load()
    x = x.contiguous()
    R, C = x.shape[-2], x.shape[-1]
    batch = x.numel() // (R * C)
    out = torch.empty(*x.shape[:-2], C, Rp, device=x.device, dtype=torch.float16)
    _l.check(L.as_transpose_pad_f16(_l.ptr(x), _l.ptr(out), batch, R, C, Rp, _l.stream_ptr()), 'as_transpose_pad_f16')
    return out


def _pad_t(x, Tpad):
    return transpose_pad(x, Tpad)


def mhsa_bwd(q, k, vt, o, d_o, m, l, T):
    """Backward of ``ops.mhsa_fwd``.  q, k [B,h,T,64] fp16, vt [B,h,64,Tpad] fp16, o / d_o [B,T,h*64], m, l [B,h,T].
    -> dq, dk, dv [B,h,T,64] fp32 (w.r.t. the unscaled q, k: the head_dim^-0.5 of VT:79 is applied inside)."""
    L = _l.load()
    B, h = q.shape[0], q.shape[1]
    Tpad = vt.shape[-1]
    d_oh = d_o.reshape(B, T, h, 64).permute(0, 2, 1, 3).contiguous().half()                   # [B,h,T,64]
    delta = (d_o.float() * o.float()).reshape(B, T, h, 64).sum(-1).permute(0, 2, 1).contiguous()   # [B,h,T]
    v = transpose_pad(vt, 64)[..., :T, :].contiguous()                                       # [B,h,T,64]
    qt, kt, dot = _pad_t(q, Tpad), _pad_t(k, Tpad), _pad_t(d_oh, Tpad)
    dq = torch.empty(B, h, T, 64, device=q.device, dtype=torch.float32)
    dk, dv = torch.empty_like(dq), torch.empty_like(dq)
    _l.check(L.as_mhsa_bwd(_l.ptr(q), _l.ptr(k), _l.ptr(v), _l.ptr(d_oh), _l.ptr(qt), _l.ptr(kt), _l.ptr(dot), _l.ptr(m), _l.ptr(l),
                           _l.ptr(delta), _l.ptr(dq), _l.ptr(dk), _l.ptr(dv), B, T, Tpad, h, _l.stream_ptr()), 'as_mhsa_bwd')
    return dq, dk, dv


def _linear_grads(x16, weight, dy16, need_dx, need_dw):
    """dX = dY W and dW = dY^T X on the tcgen05 GEMM (reduction dimension of dW padded to a multiple of 64 with zeros)."""
    M, N = dy16.shape
    K = x16.shape[1]
    dx = dw = None
    if need_dx:
        dx = ops.linear_f16(dy16, weight.detach().t().contiguous().half(), None, ops.EPI_F16)        # [M,K]
    if need_dw:
        Mp = (M + 63) // 64 * 64
        dw = ops.linear_f16(transpose_pad(dy16, Mp), transpose_pad(x16, Mp), None, ops.EPI_F32)      # [N,K] fp32
    return dx, dw


class LinearFn(torch.autograd.Function):
    """y = x16 @ weight^T + bias.  x16 [M,K] fp16, weight [N,K] (fp32 parameter), out fp32 (``out_f32``) or fp16."""

    @staticmethod
    def forward(ctx, x16, weight, bias, out_f32):
        y = ops.linear_f16(x16, weight.detach().half().contiguous(), None if bias is None else bias.detach().float(),
                           ops.EPI_F32 if out_f32 else ops.EPI_F16)
        ctx.save_for_backward(x16, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x16, weight = ctx.saved_tensors
        dy16 = dy.half().contiguous()
        dx, dw = _linear_grads(x16, weight, dy16, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        db = dy.float().sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None


class LinearResidFn(torch.autograd.Function):
    """y = resid + x16 @ weight^T + bias in one kernel (the GEMM's fp32 residual epilogue): VT:114-115's ``x + f(x)``."""

    @staticmethod
    def forward(ctx, x16, weight, bias, resid):
        y = ops.linear_f16(x16, weight.detach().half().contiguous(), None if bias is None else bias.detach().float(),
                           ops.EPI_RESID_F32, resid=resid.contiguous())
        ctx.save_for_backward(x16, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x16, weight = ctx.saved_tensors
        dy16 = dy.half().contiguous()
        dx, dw = _linear_grads(x16, weight, dy16, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        db = dy.float().sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, dy


class AttentionFn(torch.autograd.Function):
    """VT:74-83 (qkv Linear, head split, softmax(q k^T * scale) v) on [B*T, C] fp16 tokens.
    -> (o [B*T, C] fp16, q, k [B,h,T,64], m, l [B,h,T]); q, k, m, l are not differentiable (they feed the head-mean pass)."""

    @staticmethod
    def forward(ctx, xn16, weight, bias, B, T, heads):
        Tpad = (T + 127) // 128 * 128
        q, k, vt = ops.qkv_proj(xn16, weight.detach().half().contiguous(), None if bias is None else bias.detach().float(), B, T, heads, Tpad)
        o, m, l = ops.mhsa_fwd(q, k, vt, T)
        ctx.save_for_backward(xn16, weight, q, k, vt, o, m, l)
        ctx.dims = (B, T, heads)
        ctx.has_bias = bias is not None
        ctx.mark_non_differentiable(q, k, m, l)
        return o.view(B * T, -1), q, k, m, l

    @staticmethod
    @once_differentiable
    def backward(ctx, d_o, _dq, _dk, _dm, _dl):
        xn16, weight, q, k, vt, o, m, l = ctx.saved_tensors
        B, T, h = ctx.dims
        C = h * 64
        dq, dk, dv = mhsa_bwd(q, k, vt, o, d_o.reshape(B, T, C), m, l, T)
        # back to the token-major [B*T, 3C] layout of the qkv Linear output (VT:76: index = which * C + head * 64 + d)
        dqkv = torch.stack((dq, dk, dv), dim=0).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * C).half()
        dx, dw = _linear_grads(xn16, weight, dqkv, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        db = dqkv.float().sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None, None


_WS = {}
USE_TN_GEMM = os.environ.get('AS_TRAIN_TN_GEMM', '1') != '0'      # dW GEMMs read the row-major activations directly


def _workspace(dev, nbytes):
    """One scratch buffer per device for the column partials of the reductions below (stream-ordered reuse)."""
    ws = _WS.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = _WS[dev] = torch.empty(max(int(nbytes), 1 << 24), dtype=torch.uint8, device=dev)
    return ws


def colsum(x, cast=False):
    """x [M,N] fp32 / fp16 -> column sums [N] fp32 (the bias gradient); ``cast``: also half(x) from the same pass."""
    L = _l.load()
    x = x.contiguous()
    M, N = x.shape
    out = torch.empty(N, device=x.device, dtype=torch.float32)
    x16 = torch.empty(M, N, device=x.device, dtype=torch.float16) if cast else None
    ws = _workspace(x.device, L.as_colsum_workspace(N))
    _l.check(L.as_colsum(_l.ptr(x), int(x.dtype == torch.float16), M, N, _l.ptr(x16), _l.ptr(out), _l.ptr(ws), ws.numel(),
                         _l.stream_ptr()), 'as_colsum')
    return out, x16


def gelu_bwd(d_hid, pre):
    """d_hid, pre [M,N] fp16 -> (d_pre = d_hid * gelu'(pre) fp16, its column sums fp32)."""
    L = _l.load()
    M, N = pre.shape
    d_pre = torch.empty_like(pre)
    db = torch.empty(N, device=pre.device, dtype=torch.float32)
    ws = _workspace(pre.device, L.as_colsum_workspace(N))
    _l.check(L.as_gelu_bwd_f16(_l.ptr(d_hid.contiguous()), _l.ptr(pre), M, N, _l.ptr(d_pre), _l.ptr(db), _l.ptr(ws), ws.numel(),
                               _l.stream_ptr()), 'as_gelu_bwd_f16')
    return d_pre, db


def layernorm_bwd(x, gamma, dy16, resid_grad, eps):
    """Backward of ``ops.layernorm_f16``: -> (dx fp32 [M,C] (+ resid_grad), d gamma, d beta)."""
    L = _l.load()
    M, C = x.shape
    dx = torch.empty_like(x)
    dgb = torch.empty(2, C, device=x.device, dtype=torch.float32)
    ws = _workspace(x.device, L.as_layernorm_bwd_workspace(C))
    _l.check(L.as_layernorm_bwd(_l.ptr(x), _l.ptr(gamma.detach().float().contiguous()), _l.ptr(dy16.contiguous()),
                                _l.ptr(resid_grad), M, C, float(eps), _l.ptr(dx), _l.ptr(dgb), _l.ptr(ws), ws.numel(),
                                _l.stream_ptr()), 'as_layernorm_bwd')
    return dx, dgb[0], dgb[1]


def attn_bwd_prep(d_o16, o16, B, T, heads):
    """d_o, o [B*T, h*64] fp16 -> (d_oh [B,h,T,64] fp16, delta [B,h,T] fp32)."""
    L = _l.load()
    d_oh = torch.empty(B, heads, T, 64, device=o16.device, dtype=torch.float16)
    delta = torch.empty(B, heads, T, device=o16.device, dtype=torch.float32)
    _l.check(L.as_attn_bwd_prep(_l.ptr(d_o16.contiguous()), _l.ptr(o16), B, T, heads, _l.ptr(d_oh), _l.ptr(delta), _l.stream_ptr()),
             'as_attn_bwd_prep')
    return d_oh, delta


def mhsa_bwd_qkv16(q, k, v, d_oh, qt, kt, dot, m, l, delta, B, T, Tpad, heads):
    """``as_mhsa_bwd_ex`` in its fp16 mode: dQ | dK | dV as one [B*T, 3*heads*64] fp16 tensor in the layout of the qkv Linear's
    output (VT:76)."""
    L = _l.load()
    dqkv = torch.empty(B * T, 3 * heads * 64, device=q.device, dtype=torch.float16)
    _l.check(L.as_mhsa_bwd_ex(_l.ptr(q), _l.ptr(k), _l.ptr(v), _l.ptr(d_oh), _l.ptr(qt), _l.ptr(kt), _l.ptr(dot), _l.ptr(m),
                              _l.ptr(l), _l.ptr(delta), None, None, None, _l.ptr(dqkv), B, T, Tpad, heads, _l.stream_ptr()),
             'as_mhsa_bwd_ex')
    return dqkv


def drop_path_scale(B, drop_prob, device):
    """Stochastic depth, VT:21-29: the per-sample factor ``floor(keep + U[0,1)) / keep`` (0 or 1 / keep) of one residual branch."""
    keep = 1.0 - float(drop_prob)
    return (keep + torch.rand(B, device=device)).floor_() / keep


def _drop_path_apply(x, z, s, B):
    """x, z = x + branch(x) [B*T, C] fp32 (the GEMM's residual epilogue), s [B] -> x + s_b * branch(x)  (VT:114-115)."""
    M, C = x.shape
    return torch.addcmul(x.view(B, -1, C), (z - x).view(B, -1, C), s.view(B, 1, 1)).view(M, C)


def _scale_rows(g, s, B):
    """Gradient of the branch under stochastic depth: rows of sample b times s[b]."""
    M, C = g.shape
    return (g.view(B, -1, C) * s.view(B, 1, 1)).view(M, C)


class BlockFn(torch.autograd.Function):
    """One pre-LN block (VT:109-124) with a hand-scheduled backward: every pass over a [tokens, C] tensor that autograd would
    spend on a cast, a bias-gradient reduction, a GELU / LayerNorm backward or a residual-gradient add is folded into one of
    the kernels of ``csrc/vit_train.cu``; the attention backward writes dQ | dK | dV straight into the qkv Linear's layout.
    Saved per block: x, x1 (fp32), LN outputs, o, pre, gelu(pre) (fp16), q, k, v^T, (m, l), fp16 weights.
    -> (x_out [B*T,C] fp32, q, k, m, l); q, k, m, l are not differentiable (they feed the head-mean pass)."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkvw, qkvb, pw, pb, n2w, n2b, f1w, f1b, f2w, f2b, B, T, heads, eps1, eps2, s1=None, s2=None):
        import torch.nn.functional as F
        f32 = lambda t: None if t is None else t.detach().float().contiguous()
        w16 = lambda t: t.detach().half().contiguous()
        x = x.contiguous()
        Tpad = (T + 127) // 128 * 128
        xn1 = ops.layernorm_f16(x, f32(n1w), f32(n1b), eps1)
        wq, wp, w1, w2 = w16(qkvw), w16(pw), w16(f1w), w16(f2w)
        q, k, vt = ops.qkv_proj(xn1, wq, f32(qkvb), B, T, heads, Tpad)
        o, m, l = ops.mhsa_fwd(q, k, vt, T)
        o = o.view(B * T, -1)
        x1 = ops.linear_f16(o, wp, f32(pb), ops.EPI_RESID_F32, resid=x)
        if s1 is not None:                                                # stochastic depth on the attention branch (VT:114)
            x1 = _drop_path_apply(x, x1, s1, B)
        xn2 = ops.layernorm_f16(x1, f32(n2w), f32(n2b), eps2)
        pre = ops.linear_f16(xn2, w1, f32(f1b), ops.EPI_F16)
        hid = F.gelu(pre)
        x2 = ops.linear_f16(hid, w2, f32(f2b), ops.EPI_RESID_F32, resid=x1)
        if s2 is not None:                                                # ... and on the MLP branch (VT:115)
            x2 = _drop_path_apply(x1, x2, s2, B)
        ctx.save_for_backward(x, xn1, q, k, vt, o, m, l, x1, xn2, pre, hid, wq, wp, w1, w2, n1w, n2w, s1, s2)
        ctx.dims = (B, T, heads, eps1, eps2)
        ctx.has_bias = (qkvb is not None, pb is not None, f1b is not None, f2b is not None)
        ctx.mark_non_differentiable(q, k, m, l)
        return x2, q, k, m, l

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _dq, _dk, _dm, _dl):
        x, xn1, q, k, vt, o, m, l, x1, xn2, pre, hid, wq, wp, w1, w2, n1w, n2w, s1, s2 = ctx.saved_tensors
        B, T, heads, eps1, eps2 = ctx.dims
        need = ctx.needs_input_grad
        M, C = x.shape
        Mp = (M + 63) // 64 * 64
        Tpad = vt.shape[-1]
        wt = lambda w: transpose_pad(w, w.shape[0])                       # [N,K] -> [K,N]: operand of dX = dY W
        if USE_TN_GEMM and C % 64 == 0:
            dwf = lambda d16, a16: ops.linear_tn_f16(d16.contiguous(), a16.contiguous())                               # dY^T A
        else:
            dwf = lambda d16, a16: ops.linear_f16(transpose_pad(d16, Mp), transpose_pad(a16, Mp), None, ops.EPI_F32)
        dy = dy.contiguous().float()
        # ---- MLP branch: x2 = x1 + fc2(gelu(fc1(norm2(x1))))
        db2, dy16 = colsum(dy if s2 is None else _scale_rows(dy, s2, B), cast=True)       # the branch sees s2 * dY, the residual dY
        d_hid = ops.linear_f16(dy16, wt(w2), None, ops.EPI_F16)
        dw2 = dwf(dy16, hid) if need[11] else None
        d_pre, db1 = gelu_bwd(d_hid, pre)
        d_xn2 = ops.linear_f16(d_pre, wt(w1), None, ops.EPI_F16)
        dw1 = dwf(d_pre, xn2) if need[9] else None
        dx1, dg2, dbt2 = layernorm_bwd(x1, n2w, d_xn2, dy, eps2)          # + the residual branch's gradient
        # ---- attention branch: x1 = x + proj(attn(norm1(x)))
        dbp, dy16 = colsum(dx1 if s1 is None else _scale_rows(dx1, s1, B), cast=True)
        d_o = ops.linear_f16(dy16, wt(wp), None, ops.EPI_F16)
        dwp = dwf(dy16, o) if need[5] else None
        d_oh, delta = attn_bwd_prep(d_o, o, B, T, heads)
        v = transpose_pad(vt, 64)[..., :T, :].contiguous()
        qt, kt, dot = transpose_pad(q, Tpad), transpose_pad(k, Tpad), transpose_pad(d_oh, Tpad)
        dqkv = mhsa_bwd_qkv16(q, k, v, d_oh, qt, kt, dot, m, l, delta, B, T, Tpad, heads)
        dbq, _ = colsum(dqkv)
        d_xn1 = ops.linear_f16(dqkv, wt(wq), None, ops.EPI_F16)
        dwq = dwf(dqkv, xn1) if need[3] else None
        dx0, dg1, dbt1 = layernorm_bwd(x, n1w, d_xn1, dx1, eps1)
        hb = ctx.has_bias
        return (dx0, dg1, dbt1, dwq, dbq if hb[0] else None, dwp, dbp if hb[1] else None, dg2, dbt2, dw1, db1 if hb[2] else None,
                dw2, db2 if hb[3] else None, None, None, None, None, None, None, None)


def block_forward(blk, x, B, T, heads, want_attn, headmean_kwargs=None, fused=True, drop_path=0.0):
    """VT:109-124 with autograd.  x [B*T, C] fp32 residual stream -> (x, head-mean attention (detached) or None).
    ``fused`` (default): the whole block is one autograd node (``BlockFn``); otherwise per-op nodes with torch LayerNorm / GELU."""
    import torch.nn.functional as F
    C = x.shape[1]
    s1 = s2 = None
    if drop_path > 0.0:                       # stochastic depth of the training forward (VT:97, 114-115; CFG:28 drop_path_rate=0.05)
        s1, s2 = drop_path_scale(B, drop_path, x.device), drop_path_scale(B, drop_path, x.device)
    if fused and C % 128 == 0 and C <= 1024:
        x2, q, k, m, l = BlockFn.apply(x, blk.norm1.weight, blk.norm1.bias, blk.attn.qkv.weight, blk.attn.qkv.bias,
                                       blk.attn.proj.weight, blk.attn.proj.bias, blk.norm2.weight, blk.norm2.bias,
                                       blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight, blk.mlp.fc2.bias, B, T, heads,
                                       blk.norm1.eps, blk.norm2.eps, s1, s2)
        attn = None
        if want_attn:
            with torch.no_grad():
                attn, _ = ops.attn_headmean(q, k, m, l, T, **(headmean_kwargs or {}))
        return x2, attn
    xn = F.layer_norm(x, (C,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps).half()
    o, q, k, m, l = AttentionFn.apply(xn, blk.attn.qkv.weight, blk.attn.qkv.bias, B, T, heads)
    attn = None
    if want_attn:
        with torch.no_grad():
            attn, _ = ops.attn_headmean(q, k, m, l, T, **(headmean_kwargs or {}))
    z = LinearResidFn.apply(o, blk.attn.proj.weight, blk.attn.proj.bias, x)
    x = z if s1 is None else _drop_path_apply(x, z, s1, B)
    xn = F.layer_norm(x, (C,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps).half()
    hid = F.gelu(LinearFn.apply(xn, blk.mlp.fc1.weight, blk.mlp.fc1.bias, False))       # fp16 in / out, fp32 inside
    z = LinearResidFn.apply(hid, blk.mlp.fc2.weight, blk.mlp.fc2.bias, x)
    x = z if s2 is None else _drop_path_apply(x, z, s2, B)
    return x, attn

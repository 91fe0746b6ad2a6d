"""Backward of the ViT block on the device kernels (what DDP training of the backbone needs).

The reference trains ``VisionTransformerDet`` through torch autograd over its unfused block (VT:109-124) under apex AMP O1
(fp16 GEMMs with loss scaling, mmdet/apis/train.py:83, mmdet/utils/optimizer.py:23-38).  Here the two contraction-heavy parts
have hand-written backward passes behind ``torch.autograd.Function``:

  * ``AttentionFn``  qkv projection + scaled-dot-product attention.  Forward = the inference kernels (tcgen05 GEMM with the
                     head-split epilogue, flash attention); backward = ``as_mhsa_bwd`` (tcgen05: dK / dV and dQ kernels that
                     recompute P from the saved (m, l) row statistics -- nothing of size T x T is kept) followed by the
                     projection's dX / dW GEMMs on the same tcgen05 GEMM kernel.
  * ``LinearFn``     proj / fc1 / fc2: forward, dX and dW all on ``as_linear_f16``.

LayerNorm, GELU and the residual adds stay torch ops (element-wise, autograd's own backward).  Like the reference's AMP path the
gradients that enter a GEMM are rounded to fp16: callers scale the loss (``apex.amp.scale_loss`` in the reference) when their
gradients are small.  Layout shuffles between the head-major attention tensors and the token-major GEMM operands (transposes,
zero padding of the reduction dimension to a multiple of 64) are torch copies.
"""
import torch

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


def block_forward(blk, x, B, T, heads, want_attn, headmean_kwargs=None):
    """VT:109-124 with autograd.  x [B*T, C] fp32 residual stream -> (x, head-mean attention (detached) or None)."""
    import torch.nn.functional as F
    C = x.shape[1]
    xn = F.layer_norm(x, (C,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps).half()
    o, q, k, m, l = AttentionFn.apply(xn, blk.attn.qkv.weight, blk.attn.qkv.bias, B, T, heads)
    attn = None
    if want_attn:
        with torch.no_grad():
            attn, _ = ops.attn_headmean(q, k, m, l, T, **(headmean_kwargs or {}))
    x = LinearResidFn.apply(o, blk.attn.proj.weight, blk.attn.proj.bias, x)
    xn = F.layer_norm(x, (C,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps).half()
    hid = F.gelu(LinearFn.apply(xn, blk.mlp.fc1.weight, blk.mlp.fc1.bias, False))       # fp16 in / out, fp32 inside
    x = LinearResidFn.apply(hid, blk.mlp.fc2.weight, blk.mlp.fc2.bias, x)
    return x, attn

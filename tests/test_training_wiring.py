"""Wiring of ``training.BlockFn`` (the hand-scheduled block backward) checked on the CPU: every device kernel it calls is
replaced by a torch restatement of that kernel's contract (same dtypes, same layouts), and the block's output and all
gradients are compared with torch autograd over VT:109-115 written out in fp32 -- with and without stochastic depth
(VT:21-29, CFG:28).  This pins the ORDER of operations, the gradient routing (residual vs branch, which operand feeds which
GEMM) and the arity of the autograd node; the kernels themselves are pinned on the GPU (tests/test_gpu_training.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

from attentionshift_b200 import blocks, ops, training

LOG2E = 1.4426950408889634


def _emulate_kernels(mp):
    def layernorm_f16(x, g, b, eps=1e-6):
        return F.layer_norm(x.float(), (x.shape[1],), g, b, eps).half()

    def qkv_proj(x, w, bias, B, T, heads, Tpad):
        y = x.float() @ w.float().t() + (0 if bias is None else bias)
        y = y.view(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)                      # VT:76
        vt = torch.zeros(B, heads, 64, Tpad, dtype=torch.float16)
        vt[..., :T] = y[2].transpose(-1, -2).half()
        return y[0].half().contiguous(), y[1].half().contiguous(), vt

    def mhsa_fwd(q, k, vt, T):
        s = (q.float() @ k.float().transpose(-1, -2)) * (0.125 * LOG2E)
        m = s.max(-1).values
        p = torch.exp2(s - m[..., None])
        l = p.sum(-1)
        o = (p / l[..., None]) @ vt[..., :T].float().transpose(-1, -2)
        return o.permute(0, 2, 1, 3).reshape(q.shape[0], T, -1).half(), m, l

    def linear_f16(x, w, bias=None, mode=ops.EPI_F16, resid=None):
        assert x.dtype == torch.float16 and w.dtype == torch.float16
        y = x.float() @ w.float().t() + (0 if bias is None else bias)
        if mode == ops.EPI_F16:
            return y.half()
        return y if mode == ops.EPI_F32 else resid + y

    def linear_tn_f16(a, b):
        assert a.dtype == torch.float16 and b.dtype == torch.float16
        return a.float().t() @ b.float()

    def transpose_pad(x, Rp):
        out = torch.zeros(*x.shape[:-2], x.shape[-1], Rp, dtype=torch.float16)
        out[..., :x.shape[-2]] = x.transpose(-1, -2)
        return out

    def colsum(x, cast=False):
        return x.float().sum(0), (x.half() if cast else None)

    def gelu_bwd(d_hid, pre):
        p = pre.float().requires_grad_(True)
        with torch.enable_grad():
            (F.gelu(p) * d_hid.float()).sum().backward()
        d_pre = p.grad.half()
        return d_pre, d_pre.float().sum(0)

    def layernorm_bwd(x, gamma, dy16, resid_grad, eps):
        xx, gg = x.detach().clone().requires_grad_(True), gamma.detach().float().clone().requires_grad_(True)
        bb = torch.zeros_like(gg).requires_grad_(True)
        with torch.enable_grad():
            (F.layer_norm(xx, (x.shape[1],), gg, bb, eps) * dy16.float()).sum().backward()
        return xx.grad + (0 if resid_grad is None else resid_grad), gg.grad, bb.grad

    def attn_bwd_prep(d_o, o, B, T, heads):
        d_oh = d_o.view(B, T, heads, 64).permute(0, 2, 1, 3).contiguous()
        delta = (d_o.float() * o.float()).view(B, T, heads, 64).sum(-1).permute(0, 2, 1).contiguous()
        return d_oh, delta

    def mhsa_bwd_qkv16(q, k, v, d_oh, qt, kt, dot, m, l, delta, B, T, Tpad, heads):
        qq, kk, vv = (t.float().clone().requires_grad_(True) for t in (q, k, v))
        with torch.enable_grad():
            p = ((qq @ kk.transpose(-1, -2)) * 0.125).softmax(-1)                   # VT:79-80
            ((p @ vv) * d_oh.float()).sum().backward()
        g = torch.stack((qq.grad, kk.grad, vv.grad), 0)                             # [3,B,h,T,64]
        return g.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * heads * 64).half()

    for name, fn in (('layernorm_f16', layernorm_f16), ('qkv_proj', qkv_proj), ('mhsa_fwd', mhsa_fwd), ('linear_f16', linear_f16),
                     ('linear_tn_f16', linear_tn_f16)):
        mp.setattr(ops, name, fn)
    for name, fn in (('transpose_pad', transpose_pad), ('colsum', colsum), ('gelu_bwd', gelu_bwd), ('layernorm_bwd', layernorm_bwd),
                     ('attn_bwd_prep', attn_bwd_prep), ('mhsa_bwd_qkv16', mhsa_bwd_qkv16)):
        mp.setattr(training, name, fn)


def _reference_block(blk, x, B, T, heads, s1, s2):
    """VT:109-115 (+ VT:21-29) in fp32 torch ops."""
    C = x.shape[1]
    xn = F.layer_norm(x, (C,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
    qkv = (xn @ blk.attn.qkv.weight.t() + blk.attn.qkv.bias).view(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    p = ((qkv[0] @ qkv[1].transpose(-1, -2)) * 0.125).softmax(-1)
    o = (p @ qkv[2]).permute(0, 2, 1, 3).reshape(B * T, C)
    y = o @ blk.attn.proj.weight.t() + blk.attn.proj.bias
    x = x + y * s1.repeat_interleave(T)[:, None]
    xn = F.layer_norm(x, (C,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
    y = F.gelu(xn @ blk.mlp.fc1.weight.t() + blk.mlp.fc1.bias) @ blk.mlp.fc2.weight.t() + blk.mlp.fc2.bias
    return x + y * s2.repeat_interleave(T)[:, None]


@pytest.mark.parametrize('drop', [0.0, 0.25])
@pytest.mark.parametrize('fused', [True, False])
def test_block_backward_wiring(monkeypatch, drop, fused):
    _emulate_kernels(monkeypatch)
    torch.manual_seed(0)
    B, T, heads, C = 3, 37, 2, 128
    blk = blocks.Block(C, heads, mlp_ratio=4.0, qkv_bias=True)
    for p in blk.parameters():
        p.data.add_(torch.randn_like(p) * 0.05)
    keep = 1.0 - drop
    scales = [torch.tensor([1.0, 0.0, 1.0]) / keep, torch.tensor([0.0, 1.0, 1.0]) / keep] if drop else [torch.ones(B), torch.ones(B)]
    it = iter(scales)
    monkeypatch.setattr(training, 'drop_path_scale', lambda b, p, dev: next(it))
    if not fused:                               # the per-op path calls the device attention backward through mhsa_bwd
        def mhsa_bwd(q, k, vt, o, d_o, m, l, Tq):
            g = training.mhsa_bwd_qkv16(q, k, vt[..., :Tq].transpose(-1, -2), d_o.reshape(B, Tq, heads, 64).permute(0, 2, 1, 3),
                                        None, None, None, m, l, None, B, Tq, vt.shape[-1], heads)
            g = g.float().view(B, Tq, 3, heads, 64).permute(2, 0, 3, 1, 4)
            return g[0].contiguous(), g[1].contiguous(), g[2].contiguous()
        monkeypatch.setattr(training, 'mhsa_bwd', mhsa_bwd)
    x0 = torch.randn(B * T, C)
    w = torch.randn(B * T, C)
    x = x0.clone().requires_grad_(True)
    y, attn = training.block_forward(blk, x, B, T, heads, False, fused=fused, drop_path=drop)
    assert attn is None
    (y * w).sum().backward()
    got = [y.detach(), x.grad] + [p.grad.clone() for p in blk.parameters()]
    blk.zero_grad(set_to_none=True)
    xr = x0.clone().requires_grad_(True)
    yr = _reference_block(blk, xr, B, T, heads, scales[0], scales[1])
    (yr * w).sum().backward()
    ref = [yr.detach(), xr.grad] + [p.grad.clone() for p in blk.parameters()]
    names = ['y', 'dx'] + [n for n, _ in blk.named_parameters()]
    for n, a, b in zip(names, got, ref):
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        assert err < 2e-2, (n, err)


def test_drop_path_scale_is_the_reference_formula():
    """VT:24-28: floor(keep + U) / keep per sample."""
    torch.manual_seed(3)
    s = training.drop_path_scale(10000, 0.2, 'cpu')
    vals = set(torch.unique(s).tolist())
    assert vals <= {0.0, 1.25}
    assert abs(float((s > 0).float().mean()) - 0.8) < 0.02 and math.isclose(float(s.mean()), 1.0, abs_tol=0.03)


@pytest.mark.parametrize('drop_path_rate', [0.0, 0.3])
def test_backbone_training_forward_wiring(monkeypatch, drop_path_rate):
    """``VisionTransformerDet._forward_train`` end to end on the CPU with the emulated kernels: parameter gradients against
    the oracle backbone under autograd (rate 0), and stochastic depth drawn per block with the decay rule of VT:160."""
    from attentionshift_b200.registry import build_backbone
    from attentionshift_b200.synthetic import vit_state_dict
    from oracle import vit as V
    _emulate_kernels(monkeypatch)

    def patch_im2col_f16(img):
        B = img.shape[0]
        return F.unfold(img, kernel_size=16, stride=16).transpose(1, 2).reshape(-1, 768).half()
    monkeypatch.setattr(ops, 'patch_im2col_f16', patch_im2col_f16)
    drawn = []
    real = training.drop_path_scale
    monkeypatch.setattr(training, 'drop_path_scale', lambda b, p, dev: (drawn.append(round(p, 6)), real(b, p, dev))[1])
    embed, heads, depth, img, n_pt = 128, 2, 3, 64, 4
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=8)
    m = build_backbone(dict(type='VisionTransformerDet', img_size=img, patch_size=16, embed_dim=embed, depth=depth, num_heads=heads,
                            mlp_ratio=4, qkv_bias=True, with_fpn=False, last_feat=True, return_attention=False,
                            point_tokens_num=n_pt, with_point_head=False, out_indices=[depth - 1], drop_path_rate=drop_path_rate))
    m.load_state_dict(sd, strict=False)
    m.train()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, img, img, generator=gen)
    N = (img // 16) ** 2
    w1, w2 = torch.randn(2, 1 + N, embed, generator=gen), torch.randn(2, n_pt, embed, generator=gen)
    torch.manual_seed(1)
    out = m._forward_train(x)            # forward() routes here for device tensors in training mode with autograd on
    ((out['last_feat'] * w1).sum() + (out['point_tokens'] * w2).sum()).backward()
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert len(grads) >= 4 + 12 * depth and all(torch.isfinite(g).all() for g in grads.values())
    if drop_path_rate == 0.0:
        assert drawn == []
        ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        ref = V.backbone_forward(x, ref_sd, depth, heads, n_point_tokens=n_pt)
        ((ref['last_feat'] * w1).sum() + (ref['point_tokens'] * w2).sum()).backward()
        for n, g in grads.items():
            if n in ref_sd and ref_sd[n].grad is not None:
                r = ref_sd[n].grad
                assert float((g.reshape(r.shape) - r).abs().max() / r.abs().max().clamp_min(1e-30)) < 2e-2, n
    else:
        # two draws per block with a positive rate (block 0 has rate 0: VT:160 starts the ramp at zero)
        assert drawn == [0.15, 0.15, 0.3, 0.3]

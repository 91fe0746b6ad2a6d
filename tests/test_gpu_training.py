"""Backward of the ViT block on the device kernels against torch autograd in fp32: the attention backward alone
(as_mhsa_bwd vs autograd through VT:79-83 restated on the same fp16 operands) and the parameter gradients of a whole small
backbone against the CPU oracle differentiated by autograd.  Tolerances: P, dS and every GEMM operand are rounded to fp16
(fp32 accumulation) -- the arithmetic of the reference's own apex-O1 training -- so gradients agree to ~1e-2 of their scale."""
import pytest
import torch

from attentionshift_b200.synthetic import vit_state_dict
from oracle import vit as V

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('B,heads,T', [(1, 2, 128), (2, 3, 297), (1, 2, 640), (1, 12, 1125), (1, 1, 4197)])
def test_mhsa_backward_vs_autograd(B, heads, T):
    from attentionshift_b200 import ops, training
    torch.manual_seed(T)
    Tpad = (T + 127) // 128 * 128
    q = (torch.randn(B, heads, T, 64, device='cuda') * 1.2).half()
    k = (torch.randn(B, heads, T, 64, device='cuda') * 1.2).half()
    v = torch.randn(B, heads, T, 64, device='cuda').half()
    vt = torch.zeros(B, heads, 64, Tpad, device='cuda', dtype=torch.float16)
    vt[..., :T] = v.transpose(-1, -2)
    g = torch.randn(B, T, heads * 64, device='cuda')
    o, m, l = ops.mhsa_fwd(q, k, vt, T)
    dq, dk, dv = training.mhsa_bwd(q, k, vt, o, g, m, l, T)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    attn = ((qf @ kf.transpose(-2, -1)) * 0.125).softmax(-1)                      # VT:79-80
    of = (attn @ vf).transpose(1, 2).reshape(B, T, -1)                             # VT:83
    (of * g).sum().backward()
    errs = dict(dq=_rel(dq, qf.grad), dk=_rel(dk, kf.grad), dv=_rel(dv, vf.grad))
    print(f'mhsa backward B={B} h={heads} T={T}: max-norm relative errors {errs}')
    assert max(errs.values()) < 1e-2, errs
    dq2, dk2, dv2 = training.mhsa_bwd(q, k, vt, o, g, m, l, T)                    # deterministic: no atomics
    assert torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)


@pytest.mark.parametrize('embed,heads,depth,img,n_pt', [(128, 2, 2, 224, 12), (384, 6, 3, 160, 100)])
def test_backbone_parameter_gradients_vs_oracle(embed, heads, depth, img, n_pt):
    from attentionshift_b200.registry import build_backbone
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=8)
    m = build_backbone(dict(type='VisionTransformerDet', img_size=img, patch_size=16, embed_dim=embed, depth=depth, num_heads=heads,
                            mlp_ratio=4, qkv_bias=True, with_fpn=False, last_feat=True, return_attention=True, point_tokens_num=n_pt,
                            with_point_head=False, out_indices=[depth - 1]))
    m.load_state_dict(sd, strict=False)
    m = m.cuda().train()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, img, img, generator=gen)
    N = (img // 16) ** 2
    w1 = torch.randn(2, 1 + N, embed, generator=gen)
    w2 = torch.randn(2, n_pt, embed, generator=gen)
    out = m(x.cuda())
    assert out['last_feat'].requires_grad and all(a is not None and not a.requires_grad for a in out['attns'])
    loss = (out['last_feat'] * w1.cuda()).sum() + (out['point_tokens'] * w2.cuda()).sum()
    loss.backward()
    # oracle: the same forward in fp32 on the CPU, differentiated by autograd
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = V.backbone_forward(x, ref_sd, depth, heads, n_point_tokens=n_pt)
    ((ref['last_feat'] * w1).sum() + (ref['point_tokens'] * w2).sum()).backward()
    torch.testing.assert_close(out['last_feat'].detach().cpu(), ref['last_feat'].detach(), rtol=0, atol=2e-3 * float(ref['last_feat'].detach().abs().max()))
    worst = {}
    for name, p in m.named_parameters():
        if name not in ref_sd or ref_sd[name].grad is None:
            continue
        assert p.grad is not None, name
        worst[name] = _rel(p.grad.detach().cpu().reshape(ref_sd[name].grad.shape), ref_sd[name].grad)
    bad = {k: v for k, v in worst.items() if v > 2e-2}
    print('parameter gradients, worst max-norm relative errors:', sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    assert len(worst) >= 4 + 12 * depth and not bad, bad


@pytest.mark.parametrize('batch,R,C,Rp', [(1, 100, 64, 128), (3, 4197, 64, 4224), (1, 33576, 768, 33600), (2, 65, 130, 66)])
def test_transpose_pad(batch, R, C, Rp):
    """as_transpose_pad_f16: [batch, R, C] -> [batch, C, Rp], zero padded rows (the K-major operands of the backward GEMMs and of
    as_mhsa_bwd) -- bit-exact against torch."""
    from attentionshift_b200 import training
    x = torch.randn(batch, R, C, device='cuda').half()
    y = training.transpose_pad(x, Rp)
    ref = torch.zeros(batch, C, Rp, device='cuda', dtype=torch.float16)
    ref[..., :R] = x.transpose(-1, -2)
    assert torch.equal(y, ref)


# ---------------------------------------------------------------- element-wise / reduction kernels of the fused block backward
@pytest.mark.parametrize('M,N', [(1, 4), (37, 128), (4197, 768), (1000, 3072), (333, 2304)])
def test_colsum_and_cast(M, N):
    from attentionshift_b200 import training
    torch.manual_seed(M + N)
    x = torch.randn(M, N, device='cuda') * 3
    s, x16 = training.colsum(x, cast=True)
    assert torch.equal(x16, x.half())
    ref = x.double().sum(0)
    assert float((s.double() - ref).abs().max()) <= 1e-5 * float(x.abs().sum(0).max())
    s2, none = training.colsum(x16)
    assert none is None
    assert float((s2.double() - x16.double().sum(0)).abs().max()) <= 1e-5 * float(x.abs().sum(0).max())
    assert torch.equal(s, training.colsum(x, cast=True)[0])                      # fixed summation order


@pytest.mark.parametrize('M,N', [(5, 8), (777, 3072)])
def test_gelu_backward_vs_autograd(M, N):
    from attentionshift_b200 import training
    torch.manual_seed(N)
    pre = (torch.randn(M, N, device='cuda') * 2).half()
    dh = torch.randn(M, N, device='cuda').half()
    d_pre, db = training.gelu_bwd(dh, pre)
    p = pre.float().requires_grad_(True)
    (torch.nn.functional.gelu(p) * dh.float()).sum().backward()
    assert float((d_pre.float() - p.grad).abs().max()) <= 2e-3 * float(p.grad.abs().max())       # fp16 output
    assert float((db - d_pre.float().sum(0)).abs().max()) <= 1e-4 * float(d_pre.float().abs().sum(0).max())


@pytest.mark.parametrize('M,C,resid', [(3, 128, False), (1000, 768, True), (4197, 1024, True), (50, 384, False)])
def test_layernorm_backward_vs_autograd(M, C, resid):
    from attentionshift_b200 import ops, training
    torch.manual_seed(C + M)
    x = torch.randn(M, C, device='cuda') * 2 + 0.3
    g = torch.randn(C, device='cuda')
    b = torch.randn(C, device='cuda')
    dy = torch.randn(M, C, device='cuda').half()
    rg = torch.randn(M, C, device='cuda') if resid else None
    dx, dg, db = training.layernorm_bwd(x, g, dy, rg, 1e-6)
    xf, gf, bf = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xf, (C,), gf, bf, 1e-6)
    (y * dy.float()).sum().backward()
    ref_dx = xf.grad + (rg if resid else 0)
    assert _rel(dx, ref_dx) < 1e-5
    assert _rel(dg, gf.grad) < 1e-5 and _rel(db, bf.grad) < 1e-5
    # and the forward it differentiates is the inference kernel
    assert _rel(ops.layernorm_f16(x, g, b, 1e-6).float(), y.detach()) < 2e-3


def test_attention_backward_operands_and_layout():
    """as_attn_bwd_prep against its torch restatement, and as_mhsa_bwd_ex's fp16 qkv-layout output against the fp32 outputs
    of as_mhsa_bwd shuffled by torch (VT:76: column = which*C + head*64 + d)."""
    from attentionshift_b200 import lib as _l
    from attentionshift_b200 import ops, training
    torch.manual_seed(0)
    B, heads, T = 2, 3, 297
    C = heads * 64
    Tpad = (T + 127) // 128 * 128
    q = torch.randn(B, heads, T, 64, device='cuda').half()
    k = torch.randn(B, heads, T, 64, device='cuda').half()
    v = torch.randn(B, heads, T, 64, device='cuda').half()
    vt = torch.zeros(B, heads, 64, Tpad, device='cuda', dtype=torch.float16)
    vt[..., :T] = v.transpose(-1, -2)
    o, m, l = ops.mhsa_fwd(q, k, vt, T)
    d_o = torch.randn(B, T, C, device='cuda').half()
    d_oh, delta = training.attn_bwd_prep(d_o.view(B * T, C), o.view(B * T, C), B, T, heads)
    assert torch.equal(d_oh, d_o.view(B, T, heads, 64).permute(0, 2, 1, 3).contiguous())
    ref_delta = (d_o.float() * o.float()).view(B, T, heads, 64).sum(-1).permute(0, 2, 1)
    assert float((delta - ref_delta).abs().max()) <= 1e-5 * float(ref_delta.abs().max()) + 1e-6
    dq, dk, dv = training.mhsa_bwd(q, k, vt, o, d_o, m, l, T)
    L = _l.load()
    qt, kt, dot = (training.transpose_pad(t, Tpad) for t in (q, k, d_oh))
    dqkv = torch.zeros(B * T, 3 * C, device='cuda', dtype=torch.float16)
    _l.check(L.as_mhsa_bwd_ex(_l.ptr(q), _l.ptr(k), _l.ptr(v.contiguous()), _l.ptr(d_oh), _l.ptr(qt), _l.ptr(kt), _l.ptr(dot),
                              _l.ptr(m), _l.ptr(l), _l.ptr(delta), None, None, None, _l.ptr(dqkv), B, T, Tpad, heads,
                              _l.stream_ptr()), 'as_mhsa_bwd_ex')
    ref = torch.stack((dq, dk, dv), dim=0).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * C)
    # delta differs in the last bit (fma order) between the two routes: compare at fp16 resolution of the gradient scale
    assert float((dqkv.float() - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


@pytest.mark.parametrize('embed,heads', [(128, 2), (384, 6)])
def test_fused_block_equals_unfused(embed, heads):
    """BlockFn (one autograd node, csrc/vit_train.cu) against the per-op path (torch LayerNorm / GELU autograd) on one block."""
    from attentionshift_b200 import blocks, training
    torch.manual_seed(embed)
    B, T = 2, 297
    blk = blocks.Block(embed, heads, mlp_ratio=4.0, qkv_bias=True).cuda()
    for p in blk.parameters():
        p.data.add_(torch.randn_like(p) * 0.02)
    x0 = torch.randn(B * T, embed, device='cuda')
    w = torch.randn(B * T, embed, device='cuda')
    grads = []
    for fused in (True, False):
        blk.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        y, _ = training.block_forward(blk, x, B, T, heads, False, fused=fused)
        (y * w).sum().backward()
        grads.append((y.detach(), x.grad) + tuple(p.grad.clone() for p in blk.parameters()))
    names = ['y', 'dx'] + [n for n, _ in blk.named_parameters()]
    for n, a, b in zip(names, *grads):
        assert _rel(a, b) < 1e-2, (n, _rel(a, b))


@pytest.mark.parametrize('R,M,N', [(64, 64, 64), (100, 128, 256), (4197, 768, 768), (2 * 4197, 768, 3072), (1000, 2304, 768),
                                   (333, 192, 64)])
def test_weight_gradient_gemm_without_transposes(R, M, N):
    """as_linear_tn_f16 (MN-major operands) against the K-major GEMM on explicitly transposed copies and fp32 torch."""
    from attentionshift_b200 import ops, training
    torch.manual_seed(R + M)
    a = torch.randn(R, M, device='cuda').half()
    b = torch.randn(R, N, device='cuda').half()
    out = ops.linear_tn_f16(a, b)
    ref = (a.double().t() @ b.double()).float()
    assert _rel(out, ref) < 5e-5                      # fp32 accumulation over R products on the tensor cores
    Rp = (R + 63) // 64 * 64
    alt = ops.linear_f16(training.transpose_pad(a, Rp), training.transpose_pad(b, Rp), None, ops.EPI_F32)
    assert _rel(out, alt) < 5e-5

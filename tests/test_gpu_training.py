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

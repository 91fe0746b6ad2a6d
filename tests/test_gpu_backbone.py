"""VisionTransformerDet on the device kernels vs (a) the golden vectors produced by the UNMODIFIED reference class and
(b) the CPU oracle at ViT-B width.  Tolerances: the reference's own GPU path is fp16 GEMM + fp32 softmax (apex O1);
we use fp16 operands with fp32 accumulation, so last_feat agrees to ~1e-3 of its scale and the head-mean attention
maps to <= 1e-3 relative (north_star)."""
import os

import pytest
import torch

from attentionshift_b200.synthetic import vit_state_dict
from oracle import vit as V

pytestmark = pytest.mark.gpu


def _build(embed, heads, depth, img, n_pt, sd):
    from attentionshift_b200.registry import build_backbone
    m = build_backbone(dict(type='VisionTransformerDet', img_size=img, patch_size=16, embed_dim=embed, depth=depth,
                            num_heads=heads, mlp_ratio=4, qkv_bias=True, with_fpn=False, last_feat=True,
                            return_attention=True, point_tokens_num=n_pt, with_point_head=False,
                            out_indices=[depth - 1]))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m.cuda().eval()


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def test_backbone_vs_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, 'vit_e128_d2.pt'))
    m_ = g['meta']
    sd = vit_state_dict(m_['embed'], m_['depth'], m_['heads'], m_['img'], n_point_tokens=m_['n_pt'], seed=m_['seed'])
    model = _build(m_['embed'], m_['heads'], m_['depth'], m_['img'], m_['n_pt'], sd)
    gen = torch.Generator().manual_seed(m_['seed'] + 1)
    x = torch.randn(2, 3, m_['img'], m_['img'], generator=gen)
    out = model(x.cuda())
    assert set(out) >= {'org_feats', 'feature', 'point_tokens', 'attns', 'last_feat'}
    for a, b in zip(out['attns'], g['attns']):
        torch.testing.assert_close(a.cpu(), b, rtol=1e-3, atol=2e-6)
    assert _rel(out['last_feat'].cpu(), g['last_feat']) < 2e-3
    assert _rel(out['point_tokens'].cpu(), g['point_tokens']) < 2e-3


@pytest.mark.parametrize('embed,heads,depth,img', [(768, 12, 2, 224), (384, 6, 3, 160)])
def test_backbone_vs_oracle(embed, heads, depth, img):
    sd = vit_state_dict(embed, depth, heads, img, seed=3)
    model = _build(embed, heads, depth, img, 100, sd)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(2, 3, img, img, generator=gen)
    out = model(x.cuda())
    ref = V.backbone_forward(x, sd, depth, heads)
    for a, b in zip(out['attns'], ref['attns']):
        torch.testing.assert_close(a.cpu(), b, rtol=1e-3, atol=2e-6)
    assert _rel(out['last_feat'].cpu(), ref['last_feat']) < 2e-3


def test_cuda_graph_replay_equals_eager():
    """cuda_graph=True replays the forward as one CUDA graph: same bits as the eager launches, for every new input, and the
    graph is rebuilt when a weight changes."""
    from attentionshift_b200.registry import build_backbone
    embed, heads, depth, img, n_pt = 128, 2, 2, 224, 12
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=4)
    cfg = dict(type='VisionTransformerDet', img_size=img, patch_size=16, embed_dim=embed, depth=depth, num_heads=heads,
               mlp_ratio=4, qkv_bias=True, with_fpn=False, last_feat=True, return_attention=True, point_tokens_num=n_pt,
               with_point_head=True, out_indices=[depth - 1])
    eager = build_backbone(dict(cfg))
    graph = build_backbone(dict(cfg, cuda_graph=True))
    for m in (eager, graph):
        m.load_state_dict(sd, strict=False)
        m.cuda().eval()
    gen = torch.Generator().manual_seed(2)
    for rep in range(3):
        x = torch.randn(2, 3, img, img, generator=gen).cuda()
        a, b = eager(x), graph(x)
        assert torch.equal(a['last_feat'], b['last_feat'])
        # the point heads are torch nn.Linear (cuBLAS picks its algorithm per context): close, not bit-equal
        torch.testing.assert_close(a['outputs_coord'], b['outputs_coord'], rtol=0, atol=2e-3)
        for u, v in zip(a['attns'], b['attns']):
            assert torch.equal(u, v)
        assert b['attns'][-1]._as_rowsum_part is not None
    assert len(graph._graphs) == 1
    with torch.no_grad():
        for m in (eager, graph):
            m.blocks[0].mlp.fc1.bias.add_(0.5)
    a, b = eager(x), graph(x)
    assert torch.equal(a['last_feat'], b['last_feat'])


@pytest.mark.parametrize('img,n_pt', [(224, 100), (320, 12)])
def test_rollout_format_equals_full_maps(img, n_pt):
    """attn_format='rollout' (transposed operand + row sums for every layer but the last, point-token rows of the last one) must
    give the roll-out slab the SAME bits as the reference-shaped production of full fp32 maps, and the partially written last
    map must hold exactly the full map's rows from its first computed tile on."""
    from attentionshift_b200 import attention_shift as AS
    from attentionshift_b200.registry import build_backbone
    embed, heads, depth = 128, 2, 8
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=6)
    outs = {}
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, img, img, generator=gen).cuda()
    for fmt in ('full', 'rollout'):
        m = build_backbone(dict(type='VisionTransformerDet', img_size=img, patch_size=16, embed_dim=embed, depth=depth, num_heads=heads,
                                mlp_ratio=4, qkv_bias=True, with_fpn=False, last_feat=True, return_attention=True,
                                point_tokens_num=n_pt, with_point_head=False, attn_layers=7, attn_format=fmt, out_indices=[depth - 1]))
        m.load_state_dict(sd, strict=False)
        outs[fmt] = m.cuda().eval()(x)
    full, lean = outs['full']['attns'][-7:], outs['rollout']['attns'][-7:]
    assert all(a.numel() == 0 for a in lean[:-1]) and lean[-1].shape == full[-1].shape
    r_full, r_lean = AS.rollout_rows(full, n_pt), AS.rollout_rows(lean, n_pt)
    assert torch.equal(r_full, r_lean)
    v0 = lean[-1]._as_valid_from
    assert v0 <= full[-1].shape[1] - n_pt and torch.equal(lean[-1][:, v0:], full[-1][:, v0:])
    assert torch.equal(outs['full']['last_feat'], outs['rollout']['last_feat'])

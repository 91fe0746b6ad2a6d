"""MIL layer selection (RH:2953-2972 + mae_bbox_head_mil.py:140-169) on the device: the head against a golden produced by the
UNMODIFIED reference class, and the whole ``mil_select`` stage (RoIAlign 7x7 over the per-layer pseudo boxes + the head) on the
device against the same module on the host."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(cfg):
    from attentionshift_b200 import registry
    import attentionshift_b200.mil  # noqa: F401
    return registry.build_head(dict(type='MAEBoxHeadMIL', pretrained=True, use_checkpoint=False, **cfg)).eval()


def test_mil_head_on_device_vs_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, 'mil_head.pt'))
    m = _build(g['cfg'])
    missing, unexpected = m.load_state_dict(g['state_dict'], strict=True)
    assert not missing and not unexpected
    m = m.cuda()
    with torch.no_grad():
        idx, loss = m(g['x'].cuda(), gt_labels=[l.cuda() for l in g['labels']])
    assert torch.equal(idx.cpu(), g['gt_index'])                                  # the layer choice: exact
    torch.testing.assert_close(loss.cpu(), g['mil_loss'], rtol=1e-4, atol=1e-6)   # cuBLAS vs CPU GEMM summation order


def test_mil_select_device_equals_host(golden_dir):
    from attentionshift_b200 import mil
    g = torch.load(os.path.join(golden_dir, 'mil_head.pt'))
    m = _build(g['cfg'])
    m.load_state_dict(g['state_dict'], strict=True)
    gen = torch.Generator().manual_seed(5)
    B, C, hp = 2, g['cfg']['in_channels'], 28
    fmap = torch.randn(B, C, hp, hp, generator=gen)
    boxes = []
    for n in (3, 3):
        xy = torch.rand(n, 7, 2, generator=gen) * 250
        wh = torch.rand(n, 7, 2, generator=gen) * 150 + 30
        boxes.append(torch.cat([xy, xy + wh], dim=-1))
    with torch.no_grad():
        h_idx, h_loss = mil.mil_select(m, fmap, boxes, g['labels'], stride=16, roi_size=7)
        md = _build(g['cfg'])
        md.load_state_dict(g['state_dict'], strict=True)
        md = md.cuda()
        d_idx, d_loss = mil.mil_select(md, fmap.cuda(), [b.cuda() for b in boxes], [l.cuda() for l in g['labels']], stride=16, roi_size=7)
    assert all(torch.equal(a.cpu(), b) for a, b in zip(d_idx, h_idx))
    torch.testing.assert_close(d_loss['mil_loss'].cpu(), h_loss['mil_loss'], rtol=1e-4, atol=1e-6)

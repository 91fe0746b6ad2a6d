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


def test_mil_select_on_device_kernels(golden_dir):
    """``mil_select_device``: token-major RoIAlign kernel + the head's Linear layers on the tcgen05 GEMM.  RoI features against
    torchvision's roi_align (fp32, 1e-5); layer choice against the fp32 module wherever the fp32 scores separate the top two
    layers by more than the fp16 GEMM noise."""
    from torchvision.ops import roi_align
    from attentionshift_b200 import lib, mil
    g = torch.load(os.path.join(golden_dir, 'mil_head.pt'))
    m = _build(g['cfg'])
    m.load_state_dict(g['state_dict'], strict=True)
    m = m.cuda()
    gen = torch.Generator().manual_seed(8)
    B, C, hp, wp = 2, g['cfg']['in_channels'], 28, 20
    feats = torch.randn(B, hp * wp, C, generator=gen).cuda()
    boxes = []
    for n in (3, 3):
        xy = torch.rand(n, 7, 2, generator=gen) * torch.tensor([wp * 16 - 120., hp * 16 - 120.])
        wh = torch.rand(n, 7, 2, generator=gen) * 150 + 20
        boxes.append(torch.cat([xy, xy + wh], dim=-1).cuda())
    boxes[0][0, 0] = torch.tensor([-30., -20., 60., 50.])            # reaches outside the map: the boundary rules of RoIAlign
    labels = [l.cuda() for l in g['labels']]
    idx, score = mil.mil_select_device(m, feats, boxes, labels, hp, wp, stride=16, roi_size=7)
    # RoI features of the kernel vs torchvision on the [B,C,Hp,Wp] view
    fmap = feats.view(B, hp, wp, C).permute(0, 3, 1, 2).contiguous()
    rois = mil.boxes_to_rois(boxes).float()
    ref = roi_align(fmap, rois, 7, spatial_scale=1.0 / 16, sampling_ratio=0, aligned=True)           # [R,C,7,7]
    x = torch.empty(rois.shape[0], 49, C, device='cuda')
    L = lib.load()
    lib.check(L.as_roi_align_tokens(lib.ptr(feats), feats.stride(0), lib.ptr(rois.contiguous()), rois.shape[0], hp, wp, C, 7, 1.0 / 16,
                                    lib.ptr(x), lib.stream_ptr()), 'as_roi_align_tokens')
    torch.testing.assert_close(x, ref.flatten(2).transpose(1, 2), rtol=1e-5, atol=1e-5)
    with torch.no_grad():
        r_idx, _ = mil.mil_select(m, fmap, boxes, labels, stride=16, roi_size=7)
        feats_r = ref
        t = m.decoder_embed(m.norm(feats_r.flatten(2).transpose(1, 2))) if m.with_decoder_embed else feats_r.flatten(2).transpose(1, 2)
        t = torch.relu(m.fc2(torch.relu(m.fc1(t.reshape(t.shape[0], -1)))))
        bag = m.classification_branch(t).reshape(-1, 7, 20).softmax(-1) * m.proposal_branch(t).reshape(-1, 7, 20).softmax(-2)
        ref_score = torch.gather(bag, -1, torch.cat(labels).reshape(-1, 1, 1).repeat(1, 7, 1))[..., 0]
    torch.testing.assert_close(score, ref_score, rtol=2e-2, atol=1e-5)
    top2 = ref_score.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.05 * top2[:, 0]
    got, want = torch.cat(idx), torch.cat(r_idx)
    assert torch.equal(got[clear], want[clear]) and int(clear.sum()) >= 1

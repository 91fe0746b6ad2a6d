"""The MIL hook of AttnShiftRoIHead (RH:2308-2312) on the CPU: box hand-over layout and the accepted return conventions."""
from types import SimpleNamespace

import pytest
import torch

from attentionshift_b200.head import AttnShiftRoIHead, mil_from_reference


def _head(fn):
    return AttnShiftRoIHead(bbox_head=dict(cam_layer=7), mil_fn=fn)


def test_boxes_arrive_as_the_reference_builds_them():
    L, n_per = 7, [2, 1, 3]
    boxes = torch.arange(L * 6 * 4, dtype=torch.float32).reshape(L, 6, 4)
    seen = {}

    def fn(per_img, labels, roi_feat, metas):
        seen.update(per_img=per_img, labels=labels, roi_feat=roi_feat, metas=metas)
        return [torch.full((b.shape[0],), i) for i, b in enumerate(per_img)]       # per-image list of layer indices

    idx, losses = _head(fn)._mil_select(boxes, n_per, ['l0', 'l1', 'l2'], 'feat', 'metas')
    assert [tuple(b.shape) for b in seen['per_img']] == [(2, L, 4), (1, L, 4), (3, L, 4)]
    assert torch.equal(seen['per_img'][1][0], boxes[:, 2])                     # instance 2 = image 1's only one, all layers
    assert (seen['labels'], seen['roi_feat'], seen['metas']) == (['l0', 'l1', 'l2'], 'feat', 'metas')
    assert idx.tolist() == [0, 0, 1, 2, 2, 2] and idx.dtype == torch.long and losses == {}


def test_return_conventions_and_reference_adaptor():
    boxes = torch.zeros(7, 3, 4)
    flat = torch.tensor([4, 1, 6])
    assert _head(lambda *a: flat)._mil_select(boxes, [3], None, None, None)[0].tolist() == [4, 1, 6]
    idx, losses = _head(lambda *a: (flat, {'mil_loss': 1.5}))._mil_select(boxes, [3], None, None, None)
    assert idx.tolist() == [4, 1, 6] and losses == {'mil_loss': 1.5}
    calls = []

    def ref_mil(x, sampling, gt_bboxes, gt_labels, img_metas, return_index=False):      # signature of RH:2953
        calls.append((x, sampling, len(gt_bboxes), gt_labels, img_metas, return_index))
        return ['boxes'], {'mil_loss': 2.0}, (torch.tensor([5, 5]), torch.tensor([0]))
    ref = SimpleNamespace(_mil_forward_train=ref_mil)
    idx, losses = _head(mil_from_reference(ref))._mil_select(boxes, [2, 1], 'labels', 'roi_feat', 'metas')
    assert idx.tolist() == [5, 5, 0] and losses == {'mil_loss': 2.0}
    assert calls == [('roi_feat', None, 2, 'labels', 'metas', True)]
    with pytest.raises(ValueError):
        _head(lambda *a: torch.tensor([1]))._mil_select(boxes, [3], None, None, None)


def test_builtin_mil_stage_from_the_reference_config_shapes():
    """mil_head=dict(type='MAEBoxHeadMIL') in the config: the head builds the MIL module (reference parameter names) and the
    hook runs RoIAlign 7x7 (stride 16) + the module on whatever device the tensors live on (here: the CPU)."""
    from torchvision.ops import roi_align
    from attentionshift_b200 import mil as M
    torch.manual_seed(0)
    head = AttnShiftRoIHead(bbox_head=dict(cam_layer=7),
                            bbox_roi_extractor=dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=0),
                                                    out_channels=64, featmap_strides=[16]),
                            mil_head=dict(type='MAEBoxHeadMIL', in_channels=64, embed_dim=32, num_classes=20, num_layers_query=7,
                                          hidden_dim=48, roi_size=7, with_cls=False, with_reg=False))
    assert {'mil_head.fc1.weight', 'mil_head.decoder_embed.bias', 'mil_head.classification_branch.weight'} <= set(head.state_dict())
    L, n_per = 7, [2, 1]
    fmap = torch.randn(2, 64, 14, 14)
    g = torch.Generator().manual_seed(3)
    xy = torch.rand(L, 3, 2, generator=g) * 120
    boxes = torch.cat((xy, xy + 20 + 80 * torch.rand(L, 3, 2, generator=g)), dim=-1)              # [L, n_tot, 4]
    labels = [torch.tensor([4, 11]), torch.tensor([0])]
    with torch.no_grad():
        idx, losses = head._mil_select(boxes, n_per, labels, [fmap], None)
    assert idx.shape == (3,) and idx.dtype == torch.long and int(idx.max()) < L and losses['mil_loss'].ndim == 0
    # by hand: rois in (instance, layer) order with the image index in front, as bbox2roi makes them (RH:2955)
    per_img = list(boxes.permute(1, 0, 2).split(n_per, dim=0))
    rois = M.boxes_to_rois(per_img)
    assert rois[:, 0].tolist() == [0.] * 14 + [1.] * 7 and torch.equal(rois[7:14, 1:], boxes[:, 1])
    with torch.no_grad():
        want, want_loss = head.mil_head(roi_align(fmap, rois, 7, 1 / 16, 0, True), gt_labels=labels)
    assert torch.equal(idx, want) and torch.equal(losses['mil_loss'], want_loss)
    # differentiable, like the reference's: the MIL loss reaches the head's parameters and the feature map
    fmap.requires_grad_(True)
    _, losses = head._mil_select(boxes, n_per, labels, [fmap], None)
    losses['mil_loss'].backward()
    assert fmap.grad is not None and head.mil_head.fc1.weight.grad is not None

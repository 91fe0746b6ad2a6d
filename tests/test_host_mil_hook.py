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

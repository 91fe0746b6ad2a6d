"""The CPU oracle (oracle/*.py restatement) against the committed golden vectors, which
were produced by the UNMODIFIED reference (tests/golden/make_golden.py).  Runs anywhere."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from attentionshift_b200.synthetic import structured_scene, vit_state_dict
from oracle import attnshift as O
from oracle import vit as V

FTOL = dict(rtol=1e-5, atol=1e-6)   # same torch build => bit-equal in practice; slack for other host ISAs


def _scene(meta):
    hp, n = meta['hp'], meta['n_obj']
    sc = structured_scene(hp, hp, meta['c'], n, seed=meta['scene_seed'], noise=meta['noise'])
    H = hp * 16
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n, H, H)
    return sc, up


@pytest.mark.parametrize('name', ['attnshift_224_c32.pt', 'attnshift_448_c64.pt'])
def test_attention_shift_chain(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name))
    meta = g['meta']
    sc, up = _scene(meta)
    torch.manual_seed(meta['rng_seed'])
    out = O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'],
                                  sc['gt_labels'], pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10,
                                  corr_size=21, obj_tau=0.85, mean_shift_times=meta['n_shift'])
    assert torch.equal(out['mask_points_coords'], g['mask_points_coords'])
    assert torch.equal(out['mask_points_labels'], g['mask_points_labels'])
    assert torch.equal(out['points_a'], g['points_a']) and torch.equal(out['points_b'], g['points_b'])
    torch.testing.assert_close(out['inst_fg_feat'], g['fg_feat'], **FTOL)
    torch.testing.assert_close(out['inst_bg_feat'], g['bg_feat'], **FTOL)
    if 'map_fg_last' in g:
        torch.testing.assert_close(out['map_cos_fg'], g['map_fg_last'], **FTOL)
        torch.testing.assert_close(out['map_cos_bg'], g['map_bg_last'], **FTOL)
    packed = torch.from_numpy(np.packbits(out['pseudo_gt_masks'].numpy()))
    assert torch.equal(packed, g['pseudo_masks_packed'])
    assert out['num_parts'] == g['num_parts']
    torch.testing.assert_close(out['semantic_centers_org'][0], g['sc_coords_org'], **FTOL)
    assert torch.equal(out['semantic_centers_org'][1], g['sc_labels_org'])
    assert torch.equal(out['corres_gts'], g['corres_gt'])
    assert len(out['sim_fg']) == len(g['sim_fg'])
    for a, b in zip(out['sim_fg'], g['sim_fg']):
        torch.testing.assert_close(a, b, **FTOL)


@pytest.mark.parametrize('name', ['attnshift_224_c32.pt', 'attnshift_448_c64.pt'])
def test_mean_shift_and_boxes(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name))
    meta = g['meta']
    sc, up = _scene(meta)
    prot, sim = O.mean_shift_from_maps(g['seeds_map'], sc['vit_feat'], sc['rois'], n_shift=meta['n_shift'])
    torch.testing.assert_close(prot, g['ms_prot'], **FTOL)
    torch.testing.assert_close(sim, g['ms_sim'], **FTOL)
    H = meta['hp'] * 16
    boxes = [O.bbox_from_cam(up[l, i].clone(), sc['gt_points'][i], 0.2, 0.5, (H, H))[0]
             for l in range(7) for i in range(meta['n_obj'])]
    assert torch.equal(torch.cat(boxes), g['cam_boxes'])


def test_rollout(golden_dir):
    g = torch.load(os.path.join(golden_dir, 'rollout_t61.pt'))
    m = g['meta']
    gen = torch.Generator().manual_seed(m['seed'])
    attns = [torch.softmax(4 * torch.randn(m['b'], m['t'], m['t'], generator=gen), -1) for _ in range(m['layers'])]
    full = O.rollout(attns)
    torch.testing.assert_close(full[:, :, -10:, :], g['out_rows'], **FTOL)
    rows = O.rollout_rows(attns, 10)
    torch.testing.assert_close(rows, g['out_rows'], rtol=1e-5, atol=1e-7)


def test_vit_backbone(golden_dir):
    g = torch.load(os.path.join(golden_dir, 'vit_e128_d2.pt'))
    m = g['meta']
    sd = vit_state_dict(m['embed'], m['depth'], m['heads'], m['img'], n_point_tokens=m['n_pt'], seed=m['seed'])
    gen = torch.Generator().manual_seed(m['seed'] + 1)
    x = torch.randn(2, 3, m['img'], m['img'], generator=gen)
    out = V.backbone_forward(x, sd, m['depth'], m['heads'], n_point_tokens=m['n_pt'])
    for a, b in zip(out['attns'], g['attns']):
        torch.testing.assert_close(a, b, **FTOL)
    torch.testing.assert_close(out['last_feat'], g['last_feat'], **FTOL)
    torch.testing.assert_close(out['point_tokens'], g['point_tokens'], **FTOL)


def test_ccl_connectivity_assumption():
    """cc_torch is absent from the reference tree (parity UNPINNED): the oracle assumes
    8-connectivity.  A diagonal pair must be ONE component (it would be two under
    4-connectivity) -- make the assumption visible."""
    a = np.zeros((4, 4), np.uint8)
    a[0, 0] = a[1, 1] = 1
    a[3, 0] = 1
    lab = O.ccl_label(a)
    assert lab[0, 0] == lab[1, 1] != 0
    assert lab[3, 0] not in (0, lab[0, 0])


def test_update_fg_map(golden_dir):
    """A15 second-round aggregation (RH:2737-2844) against the reference's output (maps stored as fp16)."""
    g = torch.load(os.path.join(golden_dir, 'update_fg_320_c48.pt'))
    m = g['meta']
    hp, c, n = m['hp'], m['c'], m['n_obj']
    sc = structured_scene(hp, hp, c, n, seed=m['scene_seed'], noise=0.4)
    H = hp * 16
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n, H, H)
    torch.manual_seed(m['scene_seed'])
    o = O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'], sc['gt_labels'],
                                mean_shift_times=4)
    vit = torch.cat((torch.zeros(1, 1, c), sc['vit_feat'].flatten(1).t()[None]), dim=1)
    coords = torch.cat(o['semantic_centers_split'])
    num_parts = [int(x.shape[0]) for x in o['semantic_centers_split']]
    torch.manual_seed(m['rng_seed'])
    maps, masks = O.update_fg_map([o['map_cos_fg'].clone()], vit, [coords], [num_parts], [o['inst_fg_feat']],
                                  [o['inst_bg_feat']], [sc['rois']], 0.6)
    torch.testing.assert_close(maps[0], g['maps'].float(), rtol=2e-3, atol=1e-3)
    packed = torch.from_numpy(np.packbits(masks[0].numpy()))
    # the masks threshold the maps at 0.6 x max: bit-equal with the same torch build
    assert (packed != g['masks_packed']).float().mean().item() < 1e-3


def test_point_assigner(golden_dir):
    """Point-token <-> GT matching (RH:2237-2257): attentionshift_b200.assigner against the stored outputs of the reference's
    HungarianPointAssigner + PointPseudoSampler."""
    from attentionshift_b200 import assigner as A
    g = torch.load(os.path.join(golden_dir, 'point_assigner.pt'))
    for c in g['cases']:
        pos, pos_gt = A.hungarian_point_assign(c['pred'], c['cls'], c['gt_points'], c['gt_labels'], c['img_wh'], 1.0, 10.0, 1)
        assert torch.equal(pos, c['pos_inds']) and torch.equal(pos_gt, c['pos_gt'])

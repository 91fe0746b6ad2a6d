"""Attention-shift stages on the device vs the CPU oracle / the reference-produced goldens (same seeded inputs)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from attentionshift_b200.synthetic import structured_scene
from oracle import attnshift as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _scene(hp, c, n_obj, seed, noise):
    sc = structured_scene(hp, hp, c, n_obj, seed=seed, noise=noise)
    H = hp * 16
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n_obj, H, H)
    return sc, up


def test_rollout_rows_vs_oracle():
    from attentionshift_b200 import attention_shift as AS
    torch.manual_seed(0)
    B, T, L, n_rows = 2, 333, 7, 100
    ld = 384
    bufs, views = [], []
    for _ in range(L):
        buf = torch.zeros(B, T, ld)
        buf[:, :, :T] = torch.softmax(3 * torch.randn(B, T, T), -1)
        bufs.append(buf)
        views.append(buf.to(DEV)[:, :, :T])
    ref = O.rollout([b[:, :, :T].contiguous() for b in bufs])[:, :, -n_rows:, :]
    out = AS.rollout_rows(views, n_rows)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-9)


def test_rollout_rows_tensor_core_path():
    """Roll-out fed by the head-mean kernel's own outputs (split-fp16 transposed maps) vs the fp32 oracle."""
    from attentionshift_b200 import attention_shift as AS
    from attentionshift_b200 import ops
    torch.manual_seed(1)
    B, heads, T, L, n_rows = 2, 2, 333, 4, 100
    Tpad = (T + 127) // 128 * 128
    maps = []
    for _ in range(L):
        q = (torch.randn(B, heads, T, 64, device=DEV) * 1.2).half()
        k = (torch.randn(B, heads, T, 64, device=DEV) * 1.2).half()
        vt = torch.zeros(B, heads, 64, Tpad, device=DEV, dtype=torch.float16)
        _, m, l = ops.mhsa_fwd(q, k, vt, T)
        maps.append(ops.attn_headmean(q, k, m, l, T)[0])
    ref = O.rollout([a.cpu().contiguous() for a in maps])[:, :, -n_rows:, :]
    out_tc = AS.rollout_rows(maps, n_rows, use_tensor_cores=True)
    out_cc = AS.rollout_rows(maps, n_rows, use_tensor_cores=False)
    assert out_tc.stride(2) == Tpad                     # really took the tensor-core path (padded row stride)
    torch.testing.assert_close(out_cc.cpu(), ref, rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(out_tc.cpu(), ref, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize('hp,n_obj,seed', [(14, 2, 2), (28, 3, 5), (64, 3, 1)])
def test_cam_boxes_bit_exact(hp, n_obj, seed):
    """A6/A7: integer outputs (kept-component mask, box extent) must be bit-exact on identical CAMs."""
    from attentionshift_b200 import attention_shift as AS
    sc, up = _scene(hp, 16, n_obj, seed, 0.4)
    if seed == 5:   # add a far-away blob so the area filter and the multi-component path are exercised
        sc['cams_low'][:, 0, 1:3, 1:3] += 0.9
        sc['cams_low'][:, 1, -3:-1, -4:-1] += 0.6
        up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (hp * 16, hp * 16), mode='bilinear').reshape(7, n_obj, hp * 16, hp * 16)
    N, T, n_rows = hp * hp, 1 + hp * hp + 10, 10
    rows = torch.zeros(1, 7, n_rows, T)
    for j in range(n_obj):
        rows[0, :, 3 + j, 1:1 + N] = sc['cams_low'][:, j].reshape(7, N)
    obj_img = torch.zeros(n_obj, dtype=torch.int32, device=DEV)
    obj_pt = (torch.arange(n_obj, dtype=torch.int32) + 3).to(DEV)
    cams, mm, boxes, keep = AS.cam_boxes(rows.to(DEV), obj_img, obj_pt, sc['gt_points'].to(DEV), hp, hp, 0.2, 0.5, want_keep_mask=True)
    assert torch.equal(cams.cpu().reshape(7, n_obj, hp, hp), sc['cams_low'])
    assert torch.equal(mm.cpu()[..., 0], up.flatten(2).min(-1)[0]) and torch.equal(mm.cpu()[..., 1], up.flatten(2).max(-1)[0])
    keep = keep.cpu().reshape(7, n_obj, hp * 16, hp * 16).bool()
    for l in range(7):
        for j in range(n_obj):
            ob, om = O.bbox_from_cam(up[l, j].clone(), sc['gt_points'][j], 0.2, 0.5, (hp * 16, hp * 16))
            assert torch.equal(keep[l, j], om), (l, j)
            assert torch.equal(boxes[l, j].cpu(), ob[0]), (l, j, boxes[l, j], ob)


def _run_chain(sc, hp, n_shift, rng, seed_before=None):
    from attentionshift_b200 import attention_shift as AS
    n_obj = sc['rois'].shape[0]
    N = hp * hp
    feats = sc['vit_feat'].permute(1, 2, 0).reshape(1, N, -1).contiguous().to(DEV)
    ar = torch.arange(n_obj)
    cam_sel = sc['cams_low'][sc['gt_index'], ar].reshape(n_obj, N).contiguous().to(DEV)
    mm = AS.cam_minmax(cam_sel, hp, hp)
    rois = sc['rois'].to(DEV)
    if seed_before is not None:
        torch.manual_seed(seed_before)
    rm = AS.refined_maps(cam_sel, mm, feats, [n_obj], rois, sc['gt_points'].to(DEV), hp, hp, rng, refine_times=2, obj_tau=0.85,
                         mask_thr=0.6)
    coords, labels = AS.mask_points(rm['map_fg'], rm['map_bg'], rois, [n_obj], rng, pos_thr=0.6, neg_thr=0.1, num_gt=10, corr_size=21)
    obj_img = torch.zeros(n_obj, dtype=torch.int32, device=DEV)
    parts = AS.semantic_parts(rm['map_fg'], feats, obj_img, rois, hp, hp, pos_thr=0.6, n_shift=n_shift, n_points=20,
                              num_semantic_points=3, want_trace=True)
    asm = AS.assemble_parts(parts, [n_obj], [sc['gt_labels'].to(DEV)], hp, hp)[0]
    torch.cuda.synchronize()
    return rm, coords, labels, parts, asm


def _iou(a, b):
    a, b = a.bool(), b.bool()
    return ((a & b).sum().item() + 1e-9) / ((a | b).sum().item() + 1e-9)


@pytest.mark.parametrize('name', ['attnshift_224_c32.pt', 'attnshift_448_c64.pt'])
def test_chain_vs_reference_golden(golden_dir, name):
    """Whole per-image chain (A8-A13) against vectors produced by the UNMODIFIED reference, same RNG stream."""
    from attentionshift_b200 import attention_shift as AS
    g = torch.load(os.path.join(golden_dir, name))
    meta = g['meta']
    hp, n = meta['hp'], meta['n_obj']
    sc = structured_scene(hp, hp, meta['c'], n, seed=meta['scene_seed'], noise=meta['noise'])
    rm, coords, labels, parts, asm = _run_chain(sc, hp, meta['n_shift'], AS.StreamRng(), seed_before=meta['rng_seed'])
    # seed points (index arithmetic + RNG replay): bit-exact.  golden points_a = fg(+supplement) rows, points_b = bg rows
    pts = rm['pts'].cpu().long()
    assert torch.equal(pts[0, :n + 1], g['points_a'])
    assert torch.equal(pts[0, n + 1:2 * n + 1], g['points_b'])
    torch.testing.assert_close(rm['centroid'][0, :n + 1].cpu(), g['fg_feat'].flatten(1), rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(rm['centroid'][0, n + 1:2 * n + 1].cpu(), g['bg_feat'].flatten(1), rtol=1e-3, atol=1e-5)
    if 'map_fg_last' in g:
        torch.testing.assert_close(rm['map_fg'].cpu(), g['map_fg_last'], rtol=1e-3, atol=1e-5)
        torch.testing.assert_close(rm['map_bg'].cpu(), g['map_bg_last'], rtol=1e-3, atol=1e-5)
    H = hp * 16
    gmask = torch.from_numpy(np.unpackbits(g['pseudo_masks_packed'].numpy())[:n * H * H].reshape(n, H, H))
    for j in range(n):
        assert _iou(rm['mask'][j].cpu(), gmask[j]) >= 0.999
    assert torch.equal(labels.cpu(), g['mask_points_labels'])
    assert torch.equal(coords.cpu(), g['mask_points_coords'])
    # mean-shift seeds / prototypes / maps
    assert torch.equal(parts['seed_map'].cpu().reshape(n, hp, hp), g['seeds_map'])
    torch.testing.assert_close(parts['prot'].cpu().flatten(0, 1), g['ms_prot'], rtol=1e-3, atol=1e-4 * g['ms_prot'].abs().max().item())
    torch.testing.assert_close(parts['sim'].cpu().flatten(0, 1).unflatten(-1, (hp, hp)), g['ms_sim'], rtol=1e-3, atol=1e-4)
    assert asm['num_parts'] == g['num_parts']
    torch.testing.assert_close(asm['semantic_centers_org'][0].cpu(), g['sc_coords_org'], rtol=0, atol=0)
    assert torch.equal(asm['semantic_centers_org'][1].cpu(), g['sc_labels_org'])
    assert torch.equal(asm['corres_gts'].cpu(), g['corres_gt'])
    for a, b in zip(asm['sim_fg'], g['sim_fg']):
        torch.testing.assert_close(a.cpu(), b, rtol=1e-3, atol=1e-4)


def test_batched_keyed_rng_vs_oracle():
    """Two images with different instance counts in ONE device batch, KeyedRng; the oracle is re-seeded per key."""
    from attentionshift_b200 import attention_shift as AS
    hp, c = 28, 64
    scs = [structured_scene(hp, hp, c, n, seed=s, noise=0.4) for n, s in [(2, 13), (3, 3)]]
    N = hp * hp
    rng = AS.KeyedRng(5)
    n_per = [2, 3]
    feats = torch.stack([sc['vit_feat'].permute(1, 2, 0).reshape(N, c) for sc in scs]).contiguous().to(DEV)
    cam_sel = torch.cat([sc['cams_low'][sc['gt_index'], torch.arange(n)].reshape(n, N) for sc, n in zip(scs, n_per)]).contiguous().to(DEV)
    mm = AS.cam_minmax(cam_sel, hp, hp)
    rois = torch.cat([sc['rois'] for sc in scs]).to(DEV)
    gtp = torch.cat([sc['gt_points'] for sc in scs]).to(DEV)
    rm = AS.refined_maps(cam_sel, mm, feats, n_per, rois, gtp, hp, hp, rng, refine_times=2, obj_tau=0.85, mask_thr=0.6)
    coords, labels = AS.mask_points(rm['map_fg'], rm['map_bg'], rois, n_per, rng, pos_thr=0.6, neg_thr=0.1, num_gt=10, corr_size=21)
    obj_img = torch.tensor([0, 0, 1, 1, 1], dtype=torch.int32, device=DEV)
    parts = AS.semantic_parts(rm['map_fg'], feats, obj_img, rois, hp, hp, pos_thr=0.6, n_shift=5, n_points=20, num_semantic_points=3)
    asm = AS.assemble_parts(parts, n_per, [sc['gt_labels'].to(DEV) for sc in scs], hp, hp)
    torch.cuda.synchronize()
    o = 0
    for i, (sc, n) in enumerate(zip(scs, n_per)):
        H = hp * 16
        up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n, H, H)
        ref = O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'], sc['gt_labels'],
                                      pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10, corr_size=21, obj_tau=0.85,
                                      mean_shift_times=5, hook=lambda key: torch.manual_seed(rng.seed_for(key)), img=i)
        assert torch.equal(coords[o:o + n].cpu(), ref['mask_points_coords'])
        assert torch.equal(labels[o:o + n].cpu(), ref['mask_points_labels'])
        torch.testing.assert_close(rm['map_fg'][o:o + n].cpu(), ref['map_cos_fg'], rtol=1e-3, atol=1e-5)
        for j in range(n):
            assert _iou(rm['mask'][o + j].cpu(), ref['pseudo_gt_masks'][j]) >= 0.999
        assert asm[i]['num_parts'] == ref['num_parts']
        torch.testing.assert_close(asm[i]['semantic_centers_org'][0].cpu(), ref['semantic_centers_org'][0], rtol=0, atol=0)
        o += n

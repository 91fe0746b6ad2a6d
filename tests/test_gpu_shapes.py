"""Shape coverage beyond the square ViT-B case: non-square images (BASELINE cfg5 is 1344x800), ViT-L width, and the
128-scene mask IoU target of north_star (mask IoU vs the reference algorithm >= 0.999)."""
import pytest
import torch
import torch.nn.functional as F

from attentionshift_b200.synthetic import structured_scene, vit_state_dict
from oracle import attnshift as O
from oracle import vit as V

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _iou(a, b):
    a, b = a.bool(), b.bool()
    return ((a & b).sum().item() + 1e-9) / ((a | b).sum().item() + 1e-9)


def test_backbone_non_square_vit_l_width():
    """H != W (96 x 160) and C = 1024 / 16 heads (ViT-L block shape, one layer)."""
    from attentionshift_b200.registry import build_backbone
    embed, heads, depth = 1024, 16, 1
    sd = vit_state_dict(embed, depth, heads, 96, n_point_tokens=20, seed=2)        # square 6x6 position table, resized below
    m = build_backbone(dict(type='VisionTransformerDet', img_size=96, patch_size=16, embed_dim=embed, depth=depth, num_heads=heads,
                            mlp_ratio=4, qkv_bias=True, with_fpn=False, last_feat=True, return_attention=True,
                            point_tokens_num=20, with_point_head=False, out_indices=[0]))
    m.load_state_dict(sd, strict=False)
    m = m.cuda().eval()
    x = torch.randn(2, 3, 96, 160, generator=torch.Generator().manual_seed(0))
    out = m(x.cuda())
    ref = V.backbone_forward(x, sd, depth, heads, n_point_tokens=20)
    torch.testing.assert_close(out['attns'][0].cpu(), ref['attns'][0], rtol=1e-3, atol=2e-6)
    rel = ((out['last_feat'].cpu() - ref['last_feat']).abs().max() / ref['last_feat'].abs().max()).item()
    assert rel < 2e-3
    assert out['org_feats'].shape == (2, 1, embed, 6, 10)


def test_attention_shift_non_square():
    """hp != wp through the whole chain (roll-out excluded): 12 x 20 patch grid, KeyedRng, vs the oracle."""
    from attentionshift_b200 import attention_shift as AS
    hp, wp, c, n = 12, 20, 64, 2
    sc = structured_scene(hp, wp, c, n, seed=8, noise=0.4)
    N = hp * wp
    feats = sc['vit_feat'].permute(1, 2, 0).reshape(1, N, c).contiguous().to(DEV)
    cam_sel = sc['cams_low'][sc['gt_index'], torch.arange(n)].reshape(n, N).contiguous().to(DEV)
    mm = AS.cam_minmax(cam_sel, hp, wp)
    rois = sc['rois'].to(DEV)
    rng = AS.KeyedRng(2)
    rm = AS.refined_maps(cam_sel, mm, feats, [n], rois, sc['gt_points'].to(DEV), hp, wp, rng)
    coords, labels = AS.mask_points(rm['map_fg'], rm['map_bg'], rois, [n], rng, pos_thr=0.6, neg_thr=0.1, num_gt=10)
    parts = AS.semantic_parts(rm['map_fg'], feats, torch.zeros(n, dtype=torch.int32, device=DEV), rois, hp, wp, n_shift=4,
                              n_per_img=[n])
    asm = AS.assemble_parts(parts, [n], [sc['gt_labels'].to(DEV)], hp, wp)[0]
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, wp), (hp * 16, wp * 16), mode='bilinear').reshape(7, n, hp * 16, wp * 16)
    o = O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'], sc['gt_labels'],
                                mean_shift_times=4, hook=lambda key: torch.manual_seed(rng.seed_for(key)), img=0)
    torch.testing.assert_close(rm['map_fg'].cpu(), o['map_cos_fg'], rtol=1e-3, atol=1e-5)
    assert torch.equal(coords.cpu(), o['mask_points_coords']) and torch.equal(labels.cpu(), o['mask_points_labels'])
    assert asm['num_parts'] == o['num_parts']
    torch.testing.assert_close(asm['semantic_centers_org'][0].cpu(), o['semantic_centers_org'][0], rtol=0, atol=0)


def test_mask_iou_on_128_synthetic_scenes():
    """north_star: mask IoU vs the reference >= 0.999 on 128 held-out synthetic images (one device batch of 128 images)."""
    from attentionshift_b200 import attention_shift as AS
    hp, c = 14, 32
    n_per = [1 + (i % 3) for i in range(128)]
    scs = [structured_scene(hp, hp, c, n, seed=1000 + i, noise=0.35) for i, n in enumerate(n_per)]
    N = hp * hp
    feats = torch.stack([sc['vit_feat'].permute(1, 2, 0).reshape(N, c) for sc in scs]).contiguous().to(DEV)
    cam_sel = torch.cat([sc['cams_low'][sc['gt_index'], torch.arange(n)].reshape(n, N) for sc, n in zip(scs, n_per)]).contiguous().to(DEV)
    mm = AS.cam_minmax(cam_sel, hp, hp)
    rois = torch.cat([sc['rois'] for sc in scs]).to(DEV)
    gtp = torch.cat([sc['gt_points'] for sc in scs]).to(DEV)
    rng = AS.KeyedRng(77)
    rm = AS.refined_maps(cam_sel, mm, feats, n_per, rois, gtp, hp, hp, rng)
    masks = rm['mask'].cpu()
    ious, o = [], 0
    for i, (sc, n) in enumerate(zip(scs, n_per)):
        H = hp * 16
        up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n, H, H)
        amap = up[sc['gt_index'], torch.arange(n)]
        fg, _, _, _, _, _ = O.refined_maps(amap, sc['vit_feat'], sc['rois'], thr_pos=0.2, thr_neg=0.1, num_points=20, refine_times=2,
                                           obj_tau=0.85, gt_points=sc['gt_points'],
                                           hook=lambda key: torch.manual_seed(rng.seed_for(key)), img=i)
        ref = O.pseudo_masks(fg[-1], 0.6)
        for j in range(n):
            ious.append(_iou(masks[o + j], ref[j]))
        o += n
    assert min(ious) >= 0.999, (min(ious), sum(ious) / len(ious))

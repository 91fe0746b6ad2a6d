"""AttnShiftRoIHead.seed_pseudo_gt end to end (roll-out -> CAM boxes -> refined maps -> mask points -> parts -> masks) on the
device against the CPU oracle chain, fed with the oracle backbone's fp32 outputs so that only the attention-shift half
is under test (the backbone has its own tests)."""
import pytest
import torch
import torch.nn.functional as F

from attentionshift_b200.synthetic import vit_state_dict
from oracle import attnshift as O
from oracle import vit as V

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _iou(a, b):
    a, b = a.bool(), b.bool()
    return ((a & b).sum().item() + 1e-9) / ((a | b).sum().item() + 1e-9)


def test_seed_pseudo_gt_vs_oracle():
    from attentionshift_b200 import attention_shift as AS
    from attentionshift_b200.registry import build_head
    embed, heads, depth, img, n_pt = 64, 1, 7, 224, 12
    hp = img // 16
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=4)
    # sharpen the attention so CAMs have structure: scale the qkv weights up
    for i in range(depth):
        sd[f'blocks.{i}.attn.qkv.weight'] = sd[f'blocks.{i}.attn.qkv.weight'] * 12
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, img, img, generator=g)
    ref = V.backbone_forward(x, sd, depth, heads, n_point_tokens=n_pt)
    n_per = [2, 1]
    pos_inds = [torch.tensor([3, 7]), torch.tensor([5])]
    gt_points = [torch.tensor([[60., 80.], [150., 130.]]), torch.tensor([[100., 100.]])]
    gt_index = [torch.tensor([6, 2]), torch.tensor([4])]
    labels = [torch.tensor([1, 5]), torch.tensor([9])]
    rng = AS.KeyedRng(11)
    head = build_head(dict(type='AttnShiftRoIHead', bbox_head=dict(cam_layer=7, seed_thr=0.2, seed_multiple=0.5),
                           mean_shift_times_local=3, n_seeds=20, num_semantic_points=3, rng=rng))
    attns = [a.to(DEV) for a in ref['attns']]
    vit_feat = ref['last_feat'][:, 1:].permute(0, 2, 1).unflatten(-1, (hp, hp)).to(DEV)        # DET:77 layout [B,C,Hp,Wp]
    out = head.seed_pseudo_gt(None, None, None, None, None, vit_feat=vit_feat, point_cls=torch.zeros(2, n_pt, 20), attns=attns,
                              gt_points=gt_points, gt_points_labels=labels, return_mask=True, pos_mask_thr=0.6, neg_mask_thr=0.1,
                              num_mask_point_gt=10, corr_size=21, obj_tau=0.85, pos_inds=pos_inds, gt_index=gt_index)
    assert set(out) >= {'pseudo_gt_labels', 'pseudo_gt_bboxes', 'mil_losses', 'best_attn_idx', 'map_cos_fg', 'mask_points_coords',
                        'mask_points_labels', 'semantic_centers', 'semantic_centers_split', 'semantic_centers_feat_split',
                        'semantic_centers_feat', 'num_parts', 'semantic_centers_org', 'pseudo_gt_masks', 'corres_gts',
                        'inst_fg_feat', 'inst_bg_feat'}                    # RH:2398-2415
    rows = O.rollout_rows(ref['attns'][-7:], n_pt)
    for i, n in enumerate(n_per):
        low, up = O.cams_from_rollout(rows[i], pos_inds[i], n_pt, hp, hp)
        boxes = torch.stack([torch.cat([O.bbox_from_cam(up[l, j].clone(), gt_points[i][j], 0.2, 0.5, (img, img))[0] for j in range(n)])
                             for l in range(7)])                            # [7, n, 4]
        pb = boxes[gt_index[i], torch.arange(n)]
        torch.testing.assert_close(out['pseudo_gt_bboxes'][i].cpu(), pb, rtol=0, atol=0)
        o = O.attention_shift_image(up, gt_index[i], pb, ref['last_feat'][i, 1:].t().unflatten(-1, (hp, hp)).contiguous(),
                                    gt_points[i], labels[i], pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10,
                                    corr_size=21, obj_tau=0.85, mean_shift_times=3,
                                    hook=lambda key: torch.manual_seed(rng.seed_for(key)), img=i)
        torch.testing.assert_close(out['map_cos_fg'][i].cpu(), o['map_cos_fg'], rtol=1e-3, atol=1e-5)
        assert torch.equal(out['mask_points_coords'][i].cpu(), o['mask_points_coords'])
        assert torch.equal(out['mask_points_labels'][i].cpu(), o['mask_points_labels'])
        for j in range(n):
            assert _iou(torch.from_numpy(out['pseudo_gt_masks'][i][j]), o['pseudo_gt_masks'][j]) >= 0.999
        assert out['num_parts'][i] == o['num_parts']
        torch.testing.assert_close(out['semantic_centers_org'][0][i].cpu(), o['semantic_centers_org'][0], rtol=0, atol=0)


def test_update_fg_map_vs_oracle():
    """A15 (RH:2737-2844): second-round aggregation on the device against the oracle, two images in one batch; the first
    round (maps, instance features, part centres) comes from the oracle so that only this stage is under test."""
    from attentionshift_b200 import attention_shift as AS
    from attentionshift_b200.registry import build_head
    from attentionshift_b200.synthetic import structured_scene
    hp, c, n = 20, 48, 3
    H = hp * 16
    rng = AS.KeyedRng(21)
    head = build_head(dict(type='AttnShiftRoIHead', bbox_head=dict(cam_layer=7), rng=rng))
    first, vit = [], []
    for seed in (9, 17):
        sc = structured_scene(hp, hp, c, n, seed=seed, noise=0.4)
        up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n, H, H)
        torch.manual_seed(seed)
        o = O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'], sc['gt_labels'],
                                    mean_shift_times=4)
        assert min(int(s.shape[0]) for s in o['semantic_centers_split']) > 0
        first.append((sc, o))
        vit.append(torch.cat((torch.zeros(1, c), sc['vit_feat'].flatten(1).t()), dim=0))
    vit = torch.stack(vit)                                                              # [2, 1+N, C]
    fg = [o['map_cos_fg'] for _, o in first]
    coords = [torch.cat(o['semantic_centers_split']) for _, o in first]
    parts = [[int(s.shape[0]) for s in o['semantic_centers_split']] for _, o in first]
    f_fg = [o['inst_fg_feat'] for _, o in first]
    f_bg = [o['inst_bg_feat'] for _, o in first]
    boxes = [sc['rois'] for sc, _ in first]
    maps, masks = head.update_fg_map([m.to(DEV) for m in fg], None, vit.to(DEV), [x.to(DEV) for x in coords], parts,
                                     [x.to(DEV) for x in f_fg], [x.to(DEV) for x in f_bg], [b.to(DEV) for b in boxes], 0.6)
    # the oracle is re-seeded per key from the head's RNG AFTER the call (the head advances its RNG step once per call)
    r_maps, r_masks = O.update_fg_map([m.clone() for m in fg], vit, coords, parts, f_fg, f_bg, boxes, 0.6,
                                      hook=lambda key: torch.manual_seed(rng.seed_for(key)))
    for i in range(2):
        torch.testing.assert_close(maps[i].cpu(), r_maps[i], rtol=1e-3, atol=1e-4)      # north_star fp32 tolerance
        assert masks[i].dtype.name == 'uint8' and masks[i].shape == (n, H, H)
        for j in range(n):
            assert _iou(torch.from_numpy(masks[i][j]), r_masks[i][j]) >= 0.999


def test_seed_pseudo_gt_point_matching_like_reference():
    """pos_inds=None: the head matches point tokens to GT points itself (HungarianPointAssigner, RH:2237-2257).  Must equal the
    call that is handed the same match explicitly, with the GT points / labels permuted into matched-token order."""
    from attentionshift_b200 import assigner as A
    from attentionshift_b200 import attention_shift as AS
    from attentionshift_b200.registry import build_head
    embed, heads, depth, img, n_pt = 64, 1, 7, 224, 12
    hp = img // 16
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=4)
    for i in range(depth):
        sd[f'blocks.{i}.attn.qkv.weight'] = sd[f'blocks.{i}.attn.qkv.weight'] * 12
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, img, img, generator=g)
    ref = V.backbone_forward(x, sd, depth, heads, n_point_tokens=n_pt)
    gt_points = [torch.tensor([[60., 80.], [150., 130.]]), torch.tensor([[100., 100.]])]
    labels = [torch.tensor([1, 5]), torch.tensor([9])]
    gt_index = [torch.tensor([6, 2]), torch.tensor([4])]
    point_reg = torch.rand(2, n_pt, 2, generator=g)
    point_cls = torch.randn(2, n_pt, 20, generator=g)
    attns = [a.to(DEV) for a in ref['attns']]
    vit_feat = ref['last_feat'][:, 1:].permute(0, 2, 1).unflatten(-1, (hp, hp)).to(DEV)
    metas = [dict(img_shape=(img, img, 3))] * 2

    def run(**kw):
        head = build_head(dict(type='AttnShiftRoIHead', bbox_head=dict(cam_layer=7, seed_thr=0.2, seed_multiple=0.5),
                               mean_shift_times_local=3, n_seeds=20, num_semantic_points=3, rng=AS.KeyedRng(11)))
        return head.seed_pseudo_gt(None, metas, None, None, None, vit_feat=vit_feat, point_cls=point_cls.to(DEV),
                                   point_reg=point_reg.to(DEV), attns=attns, return_mask=False, pos_mask_thr=0.6,
                                   neg_mask_thr=0.1, num_mask_point_gt=10, corr_size=21, obj_tau=0.85, gt_index=gt_index, **kw)

    auto = run(gt_points=gt_points, gt_points_labels=labels)
    pos, pgt = zip(*[A.hungarian_point_assign(point_reg[i], point_cls[i], gt_points[i], labels[i], (img, img)) for i in range(2)])
    manual = run(gt_points=[gt_points[i][pgt[i]] for i in range(2)], gt_points_labels=[labels[i][pgt[i]] for i in range(2)],
                 pos_inds=list(pos))
    for i in range(2):
        assert torch.equal(auto['pseudo_gt_bboxes'][i], manual['pseudo_gt_bboxes'][i])
        assert torch.equal(auto['map_cos_fg'][i], manual['map_cos_fg'][i])
        assert torch.equal(auto['pseudo_gt_labels'][i].cpu(), labels[i][pgt[i]])

"""Generate tests/golden/*.pt by running the UNMODIFIED reference (via oracle.ref_loader)
on fixed-seed synthetic inputs.  Only runs where /root/reference exists (the build
container).  Goldens are tied to the torch version that produced them (CPU RNG streams
and kernels): torch %s.

    python tests/golden/make_golden.py
"""
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from attentionshift_b200.synthetic import structured_scene, vit_state_dict  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def attnshift_case(hp, c, n_obj, scene_seed, rng_seed, noise, n_shift, keep_maps):
    rh = ref_loader.load_rh()
    sc = structured_scene(hp, hp, c, n_obj, seed=scene_seed, noise=noise)
    H = hp * 16
    cams_up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n_obj, H, H)
    self_ = SimpleNamespace()
    self_.mean_shift_grid_prototype = lambda *a, **k: rh.methods.mean_shift_grid_prototype(self_, *a, **k)
    torch.manual_seed(rng_seed)
    r = rh.methods.get_mask_sample_points_roi_best_attn_feat_refine(
        self_, cams_up, sc['rois'], sc['gt_index'], sc['vit_feat'].clone(), pos_thr=0.6, neg_thr=0.1,
        num_gt=10, corr_size=21, obj_tau=0.85, gt_points=sc['gt_points'])
    coords, labels, fg, bg, p_a, p_b, f_fg, f_bg = r
    s = rh.methods.get_semantic_centers(self_, fg[-1].clone(), bg[-1].clone(), sc['rois'], sc['vit_feat'].clone(),
                                        pos_thr=0.6, refine_times=n_shift, gt_labels=sc['gt_labels'],
                                        num_semantic_points=3)
    masks = torch.where(fg[-1] > fg[-1].flatten(1).max(1)[0][:, None, None] * 0.6,
                        torch.ones_like(fg[-1]), torch.zeros_like(fg[-1])).to(torch.uint8)
    # bbox-from-CAM for every (layer, instance), cc_torch replaced by the unpinned stand-in
    boxes = []
    for l in range(7):
        for i in range(n_obj):
            b, _ = rh.get_bbox_from_cam_fast(cams_up[l, i].clone(), sc['gt_points'][i].clone(), cam_thr=0.2,
                                             area_ratio=0.5, img_size=(H, H))
            boxes.append(b)
    # direct mean-shift call on the same seeds (prototype + sim goldens)
    hard = torch.where(fg[-1] > 0.6, torch.ones_like(fg[-1]), torch.zeros_like(fg[-1]))
    fg_low = F.interpolate(rh.corrosion_batch(hard[None], corr_size=11)[0].unsqueeze(0), (hp, hp), mode='bilinear')[0]
    seeds_map = torch.where(fg_low > 0.6, torch.ones_like(fg_low), torch.zeros_like(fg_low))
    prot, sim = rh.methods.mean_shift_grid_prototype(self_, seeds_map, sc['vit_feat'], sc['rois'], tau=0.1, temp=0.1,
                                                     n_shift=n_shift)
    g = dict(meta=dict(hp=hp, c=c, n_obj=n_obj, scene_seed=scene_seed, rng_seed=rng_seed, noise=noise,
                       n_shift=n_shift, torch=str(torch.__version__)),
             mask_points_coords=coords, mask_points_labels=labels, points_a=p_a, points_b=p_b,
             fg_feat=f_fg, bg_feat=f_bg,
             sc_coords=s[0][0], sc_labels=s[0][1], sc_split=[x.clone() for x in s[1]],
             sim_fg=[x.clone() for x in s[2]], sc_feat=s[4], num_parts=s[5], sc_coords_org=s[6], sc_labels_org=s[7],
             corres_gt=s[8], pseudo_masks_packed=torch.from_numpy(__import__('numpy').packbits(masks.numpy())),
             cam_boxes=torch.cat(boxes), ms_prot=prot, ms_sim=sim, seeds_map=seeds_map)
    if keep_maps:
        g['map_fg_last'] = fg[-1].clone()
        g['map_bg_last'] = bg[-1].clone()
    return g


def update_fg_case(hp, c, n_obj, scene_seed, rng_seed):
    """A15 (RH:2737-2844): the reference's second-round aggregation on the oracle-produced first round (which the other
    goldens pin against the reference)."""
    from oracle import attnshift as O
    rh = ref_loader.load_rh()
    sc = structured_scene(hp, hp, c, n_obj, seed=scene_seed, noise=0.4)
    H = hp * 16
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n_obj, H, H)
    torch.manual_seed(scene_seed)
    o = O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'], sc['gt_labels'],
                                mean_shift_times=4)
    vit = torch.cat((torch.zeros(1, 1, c), sc['vit_feat'].flatten(1).t()[None]), dim=1)
    coords = torch.cat(o['semantic_centers_split'])
    num_parts = [int(x.shape[0]) for x in o['semantic_centers_split']]
    assert min(num_parts) > 0
    self_ = SimpleNamespace()
    self_.update_fg_map_single_v3 = lambda *a, **k: rh.methods.update_fg_map_single_v3(self_, *a, **k)
    torch.manual_seed(rng_seed)
    maps, masks = rh.methods.update_fg_map(self_, [o['map_cos_fg'].clone()], None, vit.clone(), [coords], [num_parts],
                                           [o['inst_fg_feat']], [o['inst_bg_feat']], [sc['rois']], 0.6)
    return dict(meta=dict(hp=hp, c=c, n_obj=n_obj, scene_seed=scene_seed, rng_seed=rng_seed, torch=str(torch.__version__)),
                maps=maps[0].half(),      # fp16 keeps the fixture small; far finer than the tests' 1e-3
                masks_packed=torch.from_numpy(__import__('numpy').packbits(masks[0])))


def assigner_case(seeds=((0, 100, 3), (1, 100, 7), (2, 12, 12), (3, 50, 1))):
    """Point-token <-> GT matching (RH:2237-2257): the reference's HungarianPointAssigner + PointPseudoSampler with the costs
    of configs/mae/attnshift_voc12aug.py:182-187 on random predictions."""
    Assigner, Sampler = ref_loader.load_point_assigner()
    ref = Assigner(cls_cost=dict(type='FocalLossCost', weight=1.0), reg_cost=dict(type='PointL1Cost', weight=10.0), times=1)
    cases = []
    for seed, n_prop, n_gt in seeds:
        g = torch.Generator().manual_seed(seed)
        pred = torch.rand(n_prop, 2, generator=g)
        cls = torch.randn(n_prop, 20, generator=g) * 2
        gtp = torch.rand(n_gt, 2, generator=g) * torch.tensor([1000., 600.])
        lab = torch.randint(0, 20, (n_gt,), generator=g)
        ar = ref.assign(pred, cls, gtp, lab, dict(img_shape=(600, 1000, 3)))
        sr = Sampler().sample(ar, pred, gtp)
        cases.append(dict(pred=pred, cls=cls, gt_points=gtp, gt_labels=lab, img_wh=(1000, 600), pos_inds=sr.pos_inds.clone(),
                          pos_gt=sr.pos_assigned_gt_inds.clone()))
    return dict(meta=dict(torch=str(torch.__version__)), cases=cases)


def rollout_case(t, layers, b, seed):
    rh = ref_loader.load_rh()
    gen = torch.Generator().manual_seed(seed)
    attns = [torch.softmax(4 * torch.randn(b, t, t, generator=gen), -1) for _ in range(layers)]
    out = rh.attns_project_to_feature(attns)
    return dict(meta=dict(t=t, layers=layers, b=b, seed=seed), out_rows=out[:, :, -10:, :].clone())


def vit_case(embed, heads, depth, img, n_pt, seed):
    VTD = ref_loader.load_vtd()
    m = VTD(img_size=img, patch_size=16, embed_dim=embed, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
            with_fpn=False, last_feat=True, return_attention=True, point_tokens_num=n_pt, with_point_head=False,
            out_indices=[depth - 1])
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    m.eval()
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(2, 3, img, img, generator=gen)
    with torch.no_grad():
        out = m(x)
    return dict(meta=dict(embed=embed, heads=heads, depth=depth, img=img, n_pt=n_pt, seed=seed),
                attns=[a.clone() for a in out['attns']], last_feat=out['last_feat'].clone(),
                point_tokens=out['point_tokens'].clone())


def mil_case(seed=3):
    """MAEBoxHeadMIL (mae_bbox_head_mil.py:140-169), the reference class itself: parameters, RoI features in, layer choice and
    MIL loss out (mmcv's RoIAlign is absent here, so the golden starts at the RoI features)."""
    Ref = ref_loader.load_mil_head()
    cfg = dict(in_channels=48, img_size=224, patch_size=16, embed_dim=32, depth=4, num_heads=8, mlp_ratio=4., num_classes=20,
               num_layers_query=7, loss_mil_factor=1.0, hidden_dim=64, roi_size=7)
    torch.manual_seed(seed)
    ref = Ref(pretrained=True, use_checkpoint=False, **cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    n_inst = 6
    x = torch.randn(n_inst * 7, 48, 7, 7, generator=g)
    labels = [torch.tensor([3, 19, 0]), torch.tensor([7, 7, 11])]
    with torch.no_grad():
        idx, loss = ref(x.clone(), gt_labels=[l.clone() for l in labels])
    return dict(cfg=cfg, state_dict={k: v.clone() for k, v in ref.state_dict().items()}, x=x, labels=labels, gt_index=idx, mil_loss=loss)


if __name__ == '__main__':
    assert ref_loader.available(), 'needs /root/reference'
    torch.set_num_threads(8)
    if '--only-mil' in sys.argv:
        torch.save(mil_case(), os.path.join(OUT, 'mil_head.pt'))
        print('mil_head.pt', os.path.getsize(os.path.join(OUT, 'mil_head.pt')))
        sys.exit(0)
    torch.save(attnshift_case(14, 32, 2, scene_seed=11, rng_seed=5, noise=0.3, n_shift=5, keep_maps=True),
               os.path.join(OUT, 'attnshift_224_c32.pt'))
    torch.save(attnshift_case(28, 64, 3, scene_seed=3, rng_seed=7, noise=0.4, n_shift=10, keep_maps=False),
               os.path.join(OUT, 'attnshift_448_c64.pt'))
    torch.save(update_fg_case(20, 48, 3, scene_seed=9, rng_seed=10), os.path.join(OUT, 'update_fg_320_c48.pt'))
    torch.save(assigner_case(), os.path.join(OUT, 'point_assigner.pt'))
    torch.save(rollout_case(t=61, layers=7, b=2, seed=1), os.path.join(OUT, 'rollout_t61.pt'))
    torch.save(vit_case(embed=128, heads=2, depth=2, img=64, n_pt=12, seed=0), os.path.join(OUT, 'vit_e128_d2.pt'))
    torch.save(mil_case(), os.path.join(OUT, 'mil_head.pt'))
    for f in sorted(os.listdir(OUT)):
        if f.endswith('.pt'):
            print(f, os.path.getsize(os.path.join(OUT, f)))

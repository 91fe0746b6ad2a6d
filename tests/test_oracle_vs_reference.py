"""Pin the oracle restatement bit-for-bit against the UNMODIFIED reference executed on
CPU (oracle.ref_loader).  Skipped where /root/reference does not exist (the GPU box)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from attentionshift_b200.synthetic import structured_scene, vit_state_dict
from oracle import attnshift as O
from oracle import ref_loader
from oracle import vit as V

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')


def _eq(a, b):
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_eq(x, y) for x, y in zip(a, b))
    if torch.is_tensor(a):
        return torch.is_tensor(b) and a.shape == b.shape and torch.equal(a, b)
    return a == b


@pytest.mark.parametrize('hp,c,n_obj,seed,noise', [(14, 32, 2, 21, 0.3), (28, 48, 3, 9, 0.4), (20, 40, 5, 4, 0.4)])
def test_chain_bit_exact(hp, c, n_obj, seed, noise):
    rh = ref_loader.load_rh()
    sc = structured_scene(hp, hp, c, n_obj, seed=seed, noise=noise)
    H = hp * 16
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n_obj, H, H)
    self_ = SimpleNamespace()
    self_.mean_shift_grid_prototype = lambda *a, **k: rh.methods.mean_shift_grid_prototype(self_, *a, **k)
    torch.manual_seed(seed)
    r = rh.methods.get_mask_sample_points_roi_best_attn_feat_refine(
        self_, up, sc['rois'], sc['gt_index'], sc['vit_feat'].clone(), pos_thr=0.6, neg_thr=0.1, num_gt=10,
        gt_points=sc['gt_points'])
    torch.manual_seed(seed)
    o = O.mask_sample_points(up, sc['rois'], sc['gt_index'], sc['vit_feat'].clone(), pos_thr=0.6, neg_thr=0.1,
                             num_gt=10, gt_points=sc['gt_points'])
    assert _eq(list(r), list(o))
    r2 = rh.methods.get_semantic_centers(self_, r[2][-1].clone(), r[3][-1].clone(), sc['rois'], sc['vit_feat'].clone(),
                                         pos_thr=0.6, refine_times=4, gt_labels=sc['gt_labels'], num_semantic_points=3)
    o2 = O.semantic_centers(o[2][-1].clone(), o[3][-1].clone(), sc['rois'], sc['vit_feat'].clone(), pos_thr=0.6,
                            refine_times=4, gt_labels=sc['gt_labels'], num_semantic_points=3)
    assert _eq(list(r2), list(o2))


def test_rollout_and_bbox_bit_exact():
    rh = ref_loader.load_rh()
    gen = torch.Generator().manual_seed(0)
    attns = [torch.softmax(3 * torch.randn(2, 45, 45, generator=gen), -1) for _ in range(7)]
    assert torch.equal(rh.attns_project_to_feature(attns), O.rollout(attns))
    sc = structured_scene(14, 14, 16, 2, seed=2)
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, 14, 14), (224, 224), mode='bilinear').reshape(7, 2, 224, 224)
    for l in range(7):
        for i in range(2):
            rb, rm = rh.get_bbox_from_cam_fast(up[l, i].clone(), sc['gt_points'][i].clone(), cam_thr=0.2,
                                               area_ratio=0.5, img_size=(224, 224))
            ob, om = O.bbox_from_cam(up[l, i].clone(), sc['gt_points'][i], 0.2, 0.5, (224, 224))
            assert torch.equal(rb, ob) and torch.equal(rm, om)


def test_vit_block_bit_exact():
    vt = ref_loader.load_vt()
    torch.manual_seed(0)
    blk = vt.Block(dim=128, num_heads=2, mlp_ratio=4, qkv_bias=True,
                   norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6), return_attention=True).eval()
    sd = {'b.' + k: v for k, v in blk.state_dict().items()}
    x = torch.randn(2, 37, 128)
    with torch.no_grad():
        ry, ra = blk(x)
        oy, oa = V.block(x, sd, 'b.', 2)
    assert torch.equal(ry, oy)
    assert torch.equal(ra.mean(1), oa)


@pytest.mark.parametrize('hp,c,n_obj,seed', [(20, 48, 3, 9), (28, 64, 3, 11)])
def test_update_fg_map_bit_exact(hp, c, n_obj, seed):
    """A15: second-round aggregation (RH:2737-2844) on the outputs of the first round."""
    rh = ref_loader.load_rh()
    sc = structured_scene(hp, hp, c, n_obj, seed=seed, noise=0.4)
    H = hp * 16
    up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (H, H), mode='bilinear').reshape(7, n_obj, H, H)
    torch.manual_seed(seed)
    o = O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'], sc['gt_labels'],
                                mean_shift_times=4)
    vit = torch.cat((torch.zeros(1, 1, c), sc['vit_feat'].flatten(1).t()[None]), dim=1)          # [1, 1+N, C], cls first
    coords = torch.cat(o['semantic_centers_split']) if len(o['semantic_centers_split']) else torch.zeros(0, 2)
    num_parts = [int(s.shape[0]) for s in o['semantic_centers_split']]
    assert min(num_parts) > 0
    args = ([o['map_cos_fg']], None, vit, [coords], [num_parts], [o['inst_fg_feat']], [o['inst_bg_feat']], [sc['rois']], 0.6)
    self_ = SimpleNamespace()
    self_.update_fg_map_single_v3 = lambda *a, **k: rh.methods.update_fg_map_single_v3(self_, *a, **k)
    torch.manual_seed(seed + 1)
    r_maps, r_masks = rh.methods.update_fg_map(self_, *[a.clone() if torch.is_tensor(a) else a for a in args])
    torch.manual_seed(seed + 1)
    o_maps, o_masks = O.update_fg_map(args[0], args[2].clone(), args[3], args[4], args[5], args[6], args[7], args[8])
    assert torch.equal(r_maps[0], o_maps[0])
    assert (torch.from_numpy(r_masks[0]) == o_masks[0]).all()
    assert o_maps[0].abs().sum() > 0


@pytest.mark.parametrize('seed,n_prop,n_gt,times', [(0, 100, 3, 1), (1, 100, 7, 1), (2, 12, 12, 1), (3, 100, 2, 2)])
def test_point_assigner_equals_reference(seed, n_prop, n_gt, times):
    """Point-token <-> GT matching (RH:2237-2257): attentionshift_b200.assigner against the reference's
    HungarianPointAssigner + PointPseudoSampler with the costs of configs/mae/attnshift_voc12aug.py:182-187."""
    from attentionshift_b200 import assigner as A
    Assigner, Sampler = ref_loader.load_point_assigner()
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand(n_prop, 2, generator=g)
    cls = torch.randn(n_prop, 20, generator=g) * 2
    gtp = torch.rand(n_gt, 2, generator=g) * torch.tensor([1000., 600.])
    lab = torch.randint(0, 20, (n_gt,), generator=g)
    ref = Assigner(cls_cost=dict(type='FocalLossCost', weight=1.0), reg_cost=dict(type='PointL1Cost', weight=10.0), times=times)
    ar = ref.assign(pred, cls, gtp, lab, dict(img_shape=(600, 1000, 3)))
    sr = Sampler().sample(ar, pred, gtp)
    pos, pos_gt = A.hungarian_point_assign(pred, cls, gtp, lab, (1000, 600), 1.0, 10.0, times)
    assert torch.equal(pos, sr.pos_inds) and torch.equal(pos_gt, sr.pos_assigned_gt_inds)
    assert torch.equal(lab[pos_gt], ar.labels[sr.pos_inds])


def test_reference_config_builds_unchanged():
    """configs/mae/attnshift_voc12aug.py: the `backbone` and `roi_head` dicts build through this repo's registry as they are
    (mmdet passes train_cfg.rcnn / test_cfg.rcnn to the RoI head), and the values the hot path reads arrive."""
    import os
    from attentionshift_b200 import registry
    ns = {}
    exec(open(os.path.join(ref_loader.REF_ROOT, 'configs/mae/attnshift_voc12aug.py')).read(), ns)
    m = ns['model']
    bb = registry.build_backbone(dict(m['backbone']))
    assert bb.embed_dim == 384 and bb.num_heads == 6 and len(bb.blocks) == 12 and bb.point_tokens_num == 100
    assert bb.return_attention and bb.last_feat and bb.drop_path_rate == 0.05
    hd = registry.build_head(dict(m['roi_head'], train_cfg=m['train_cfg']['rcnn'], test_cfg=m['test_cfg']['rcnn']))
    assert (hd.cam_layer, hd.seed_thr, hd.seed_multiple) == (7, 0.2, 0.5)
    assert (hd.num_semantic_points, hd.mean_shift_times_local) == (5, 10)
    assert (hd.point_cls_weight, hd.point_reg_weight, hd.point_times) == (1.0, 10.0, 1)
    # the second registry name of the same class (the reference's rename is incomplete)
    assert type(registry.build_head(dict(m['roi_head'], type='StandardRoIHeadMaskPointSampleDeformAttnReppoints'))) is type(hd)


def test_mil_head_equals_reference():
    """MIL layer selection (MAEBoxHeadMIL, mae_bbox_head_mil.py:140-169) with the shipped config's sizes: same parameters in,
    same layer choice and loss out."""
    from attentionshift_b200 import registry
    import attentionshift_b200.mil  # noqa: F401  (registers MAEBoxHeadMIL)
    Ref = ref_loader.load_mil_head()
    cfg = dict(in_channels=384, img_size=224, patch_size=16, embed_dim=256, depth=4, num_heads=8, mlp_ratio=4., num_classes=20,
               num_layers_query=7, loss_mil_factor=1.0, hidden_dim=1024, roi_size=7)
    torch.manual_seed(0)
    ref = Ref(pretrained=True, use_checkpoint=False, **cfg).eval()
    ours = registry.build_head(dict(type='MAEBoxHeadMIL', pretrained=True, use_checkpoint=False, with_cls=False, with_reg=False, **cfg)).eval()
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5 * 7, 384, 7, 7, generator=g)
    labels = [torch.tensor([3, 19]), torch.tensor([0, 7, 7])]
    with torch.no_grad():
        r_idx, r_loss = ref(x.clone(), gt_labels=[l.clone() for l in labels])
        o_idx, o_loss = ours(x.clone(), gt_labels=labels)
    assert torch.equal(r_idx, o_idx) and torch.equal(r_loss, o_loss)
    assert o_idx.shape == (5,) and int(o_idx.max()) < 7

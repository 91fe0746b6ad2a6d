"""Host-side contract of the drop-in classes (no GPU): registry policy, the forward-only guard of the backbone, the per-call /
per-rank advance of the default RNG, and the reference-head composition helpers."""
import pytest
import torch
import torch.nn as nn


def test_backbone_refuses_to_be_trained_silently():
    from attentionshift_b200.registry import build_backbone
    cfg = dict(type='VisionTransformerDet', img_size=64, patch_size=16, embed_dim=128, depth=1, num_heads=2, mlp_ratio=4, qkv_bias=True,
               train_backward=False)
    bb = build_backbone(dict(cfg))
    bb.train()
    with pytest.raises(RuntimeError, match='forward-only'):
        bb(torch.zeros(1, 3, 64, 64))
    # frozen, eval, no_grad and the explicit opt-in all get past the guard (and then fail only for want of a CUDA device)
    for prep in (lambda m: m.requires_grad_(False), lambda m: m.eval()):
        m = build_backbone(dict(cfg)).train()
        prep(m)
        with pytest.raises(Exception) as e:
            m(torch.zeros(1, 3, 64, 64))
        assert 'forward-only' not in str(e.value)
    m = build_backbone(dict(cfg, allow_detached_training=True)).train()
    with pytest.raises(Exception) as e:
        m(torch.zeros(1, 3, 64, 64))
    assert 'forward-only' not in str(e.value)


def test_registry_policy():
    from attentionshift_b200 import registry
    registry._register_all()
    from attentionshift_b200.backbone import VisionTransformerDet
    from attentionshift_b200.head import AttnShiftRoIHead
    assert registry.BACKBONES.get('VisionTransformerDetB200') is VisionTransformerDet
    assert registry.HEADS.get('AttnShiftRoIHeadB200') is AttnShiftRoIHead
    if not registry.USING_MMDET:          # shim registry: no reference classes exist, the reference's names build ours
        assert registry.BACKBONES.get('VisionTransformerDet') is VisionTransformerDet
        assert registry.HEADS.get('AttnShiftRoIHead') is AttnShiftRoIHead
        assert registry.HEADS.get('StandardRoIHeadMaskPointSampleDeformAttnReppoints') is AttnShiftRoIHead


def test_keyed_rng_advances_per_call_and_rank():
    from attentionshift_b200.attention_shift import KeyedRng
    a, b = KeyedRng(5, rank=0), KeyedRng(5, rank=1)
    key = (0, 0, 3)
    assert a.seed_for(key) != b.seed_for(key)                   # DDP ranks draw different points
    s0 = a.seed_for(key)
    a.next_step()
    assert a.seed_for(key) != s0                                # ... and so does every new call
    # no collisions between slots, stages and images (the old ``stage * 101 + obj`` formula collided at obj >= 101)
    seeds = {a.seed_for((i, st, o)) for i in range(8) for st in range(4) for o in range(300)}
    assert len(seeds) == 8 * 4 * 300
    from attentionshift_b200.registry import build_head
    h = build_head(dict(type='AttnShiftRoIHead', bbox_head=dict(cam_layer=7)))
    assert h.rng.step == 0 and hasattr(h.rng, 'next_step')


def test_attach_and_dropin_keep_the_reference_head():
    """``attach`` / ``make_dropin``: the reference head keeps its losses / sub-heads, only the two attention-shift methods are
    replaced (tested on a stand-in reference class: mmdet is not installed here)."""
    from attentionshift_b200 import head as H

    class RefHead(nn.Module):                 # stand-in with the reference's attribute names
        def __init__(self, mean_shift_times_local=7, num_semantic_points=3):
            super().__init__()
            self.bbox_head = type('B', (), dict(cam_layer=5, seed_thr=0.3, seed_multiple=0.4))()
            self.mean_shift_times_local = mean_shift_times_local
            self.num_semantic_points = num_semantic_points
            self.train_cfg = None
            self.lin = nn.Linear(2, 2)

        def forward_train(self):
            return 'reference losses'

        def seed_pseudo_gt(self, *a, **k):
            return 'reference'

        def _mil_forward_train(self, *a, **k):
            return 'mil'

    ref = H.attach(RefHead())
    assert ref.forward_train() == 'reference losses'
    assert ref.seed_pseudo_gt.__self__ is ref._as_b200 and ref.update_fg_map.__self__ is ref._as_b200
    fast = ref._as_b200
    assert (fast.cam_layer, fast.seed_thr, fast.seed_multiple, fast.mean_shift_times_local) == (5, 0.3, 0.4, 7)
    assert fast.mil_fn is not None and '_as_b200' not in dict(ref.named_modules())
    D = H.make_dropin(RefHead)
    d = D(mean_shift_times_local=4)
    assert isinstance(d, RefHead) and d._as_b200.mean_shift_times_local == 4 and D.__name__ == 'AttnShiftRoIHead'
    import copy
    import pickle
    h2 = H.AttnShiftRoIHead(mil_head=dict(type='MAEBoxHeadMIL', in_channels=32, embed_dim=32, hidden_dim=16, num_classes=20, roi_size=7,
                                          num_layers_query=7))
    c = copy.deepcopy(h2)                     # no closure over the original module: the copy uses ITS mil_head
    assert c.mil_head is not h2.mil_head and c.mil_fn is None
    pickle.dumps(h2.state_dict())
    h2.mil_head.init_weights()
    assert float(h2.mil_head.fc1.bias.detach().abs().max()) == 0.0


def test_match_points_host_route_for_host_tensors():
    """``match_points`` only takes the device solver for predictions that live on the GPU; host tensors (and
    ``device_matching = False``) go the reference's way -- scipy on the host -- and return what the assigner returns."""
    import torch
    from attentionshift_b200 import assigner as A
    from attentionshift_b200.registry import build_head
    head = build_head(dict(type='AttnShiftRoIHead', bbox_head=dict(cam_layer=7, seed_thr=0.2, seed_multiple=0.5)))
    assert head.device_matching is True
    g = torch.Generator().manual_seed(4)
    reg, cls = torch.rand(2, 30, 2, generator=g), torch.randn(2, 30, 20, generator=g)
    pts = [torch.rand(3, 2, generator=g) * 200, torch.rand(1, 2, generator=g) * 200]
    lab = [torch.randint(0, 20, (3,), generator=g), torch.randint(0, 20, (1,), generator=g)]
    pos, pgt = head.match_points(reg, cls, pts, lab, [(200, 200)] * 2)
    for i in range(2):
        p_ref, g_ref = A.hungarian_point_assign(reg[i], cls[i], pts[i], lab[i], (200, 200))
        assert torch.equal(pos[i], p_ref) and torch.equal(pgt[i], g_ref)
        assert pos[i].numel() == pts[i].shape[0] and bool((pos[i][1:] > pos[i][:-1]).all())

"""``AttnShiftRoIHead`` -- drop-in for the attention-shift part of the reference RoI head
(mmdet/models/roi_heads/stdroi_point_deform_attn_reppoints.py, class at RH:1325, ``seed_pseudo_gt`` RH:2209-2415).

Registered under BOTH names the reference configs use (``AttnShiftRoIHead`` in configs/mae/attnshift_voc12aug.py:60 and
``StandardRoIHeadMaskPointSampleDeformAttnReppoints``, the class's actual name -- the reference's rename is incomplete).
``seed_pseudo_gt`` keeps the reference signature and return keys; its body runs on the sm_100a kernels through
``attention_shift``.  The two learned / host-side selections that sit in the middle of the reference function are
outside this path and enter through hooks:
  * point-token <-> GT matching (HungarianPointAssigner + PointPseudoSampler, scipy on the host in the reference): done
    the same way on the host (``assigner.py``, pinned against the reference classes), or bypassed with the ``pos_inds``
    kwarg (instances then pair ``pos_inds[i][j]`` with ``gt_points[i][j]``);
  * the MIL layer choice (RoIAlign + MAEBoxHeadMIL, RH:2953-2972): ``gt_index`` kwarg, or a ``mil_fn`` callable that gets
    the per-layer pseudo boxes exactly as the reference hands them to ``_mil_forward_train`` (``mil_from_reference`` wraps the
    reference head's own MIL stage; its losses come back under ``mil_losses``), or -- when the config carries a
    ``mil_head=dict(type='MAEBoxHeadMIL', ...)`` -- this repo's restatement of that stage (``mil.py``: torchvision RoIAlign +
    the same four Linear layers, parameter names as in the reference).
The loss-side methods of the reference class (forward_train / simple_test) are not part of the hot path.
"""
import torch
import torch.nn as nn

from . import attention_shift as AS
from .registry import HEADS


class AttnShiftRoIHead(nn.Module):
    """The attention-shift half of the reference RoI head (``seed_pseudo_gt`` RH:2209-2415, ``update_fg_map`` RH:2737-2760).
    It does NOT implement the loss side (``forward_train`` / ``simple_test`` / ``init_weights`` / bbox and mask heads), so it
    never takes the reference class's registry names inside a real mmdet: see ``_register`` at the bottom of this module and
    ``attach`` / ``make_dropin`` for how it is combined with the reference head there."""

    def __init__(self, mil_head=None, bbox_roi_extractor=None, bbox_head=None, mask_roi_extractor=None, mask_head=None,
                 shared_head=None, mae_head=None, bbox_rec_head=None, train_cfg=None, test_cfg=None, visualize=False,
                 epoch=0, epoch_semantic_centers=0, num_semantic_points=3, semantic_to_token=False, pca_dim=128,
                 mean_shift_times_local=10, reppoints_head=None, num_reppoints_head=1, n_seeds=20, mil_fn=None, rng=None):
        super().__init__()
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        # CFG:182-187 point_assigner: FocalLossCost weight 1, PointL1Cost weight 10, times 1
        pa = (train_cfg.get('point_assigner') if hasattr(train_cfg, 'get') else None) or {}
        self.point_cls_weight = float((pa.get('cls_cost') or {}).get('weight', 1.0))
        self.point_reg_weight = float((pa.get('reg_cost') or {}).get('weight', 10.0))
        self.point_times = int(pa.get('times', 1))
        if self.point_times != 1:
            raise ValueError('point_assigner.times > 1 is not part of the hot path (the shipped config uses 1)')
        cfg = bbox_head if isinstance(bbox_head, dict) else {}
        # CFG:102-105 -- the reference reads these off ``self.bbox_head``
        self.cam_layer = int(cfg.get('cam_layer', 7))
        self.seed_thr = float(cfg.get('seed_thr', 0.2))
        self.seed_multiple = float(cfg.get('seed_multiple', 0.5))
        self.visualize = visualize
        self.epoch = epoch
        self.epoch_semantic_centers = epoch_semantic_centers
        self.num_semantic_points = num_semantic_points
        self.semantic_to_token = semantic_to_token
        self.pca_dim = pca_dim
        self.mean_shift_times_local = mean_shift_times_local
        self.n_seeds = n_seeds                  # 20 in the reference (hard-coded at RH:2024)
        self.mil_fn = mil_fn
        if mil_fn is None and isinstance(mil_head, dict) and mil_head.get('type') == 'MAEBoxHeadMIL':
            # the reference builds ``self.mil_head`` from this dict (RH:1352-1356) and RoIAligns with ``bbox_roi_extractor``
            # (CFG:64-68: output 7, stride 16); same attribute name, same parameter names -> checkpoints load
            from . import mil as _mil
            self.mil_head = _mil.MAEBoxHeadMIL(**{k: v for k, v in mil_head.items() if k != 'type'})
            rex = bbox_roi_extractor if isinstance(bbox_roi_extractor, dict) else {}
            stride = int((rex.get('featmap_strides') or [16])[0])
            rsize = int((rex.get('roi_layer') or {}).get('output_size', self.mil_head.roi_size))
            self._mil_stride, self._mil_rsize = stride, rsize          # mil_fn stays None: _mil_select dispatches to self.mil_head
        # default RNG: keyed off torch's seed and the rank, advanced once per call (the reference draws from the advancing global
        # generator: different seed / mask points every iteration and on every rank)
        self.rng = rng if rng is not None else AS.KeyedRng(torch.initial_seed() & 0x7fffffff)
        self.with_mil = mil_head is not None or mil_fn is not None or hasattr(self, 'mil_head')
        self.with_deform_sup = False
        self.device_matching = True             # match_points: device solver when the predictions live on the GPU
        self._mask_bufs = {}

    def _mask_buffer(self, shape):
        """Pinned landing buffer for the uint8 masks (RH:2358 hand-off).  Two buffers per shape alternate, so the numpy views
        handed out by one call stay valid until the call after the next; no cudaHostAlloc in the steady state."""
        key = tuple(shape)
        ent = self._mask_bufs.get(key)
        if ent is None:
            if len(self._mask_bufs) > 8:
                self._mask_bufs.clear()
            ent = self._mask_bufs[key] = [[torch.empty(key, dtype=torch.uint8, pin_memory=True) for _ in range(2)], 0]
        ent[1] ^= 1
        return ent[0][ent[1]]

    # ---- the two selections that sit next to the device path ---------------------------------------------
    def match_points(self, point_reg, point_cls, gt_points, gt_labels, imgs_wh):
        """RH:2237-2257: the reference's HungarianPointAssigner + PointPseudoSampler per image.  -> (pos_inds, pos_gt): per
        image the matched point tokens in ascending order and the GT each one belongs to.  Predictions on the GPU: the whole
        batch is matched on the device (``as_hungarian_points``, the solver scipy runs, without the reference's ``cost.cpu()``
        round trip); host tensors (or ``device_matching = False``) take the reference's route, scipy on the host."""
        from .assigner import hungarian_point_assign, hungarian_point_assign_device
        from .assigner import DEVICE_MATCH_MAX
        fits = point_reg.shape[1] <= DEVICE_MATCH_MAX and max([int(g.shape[0]) for g in gt_points] + [0]) <= DEVICE_MATCH_MAX
        if point_reg.is_cuda and self.device_matching and fits:
            return hungarian_point_assign_device(point_reg, point_cls, gt_points, gt_labels, imgs_wh, self.point_cls_weight,
                                                 self.point_reg_weight)
        reg = point_reg.detach().float().cpu()
        cls = point_cls.detach().float().cpu()
        pos, pgt = [], []
        for i in range(reg.shape[0]):
            p_i, g_i = hungarian_point_assign(reg[i], cls[i], gt_points[i].detach().float().cpu(), gt_labels[i].detach().cpu().long(),
                                              imgs_wh[i], self.point_cls_weight, self.point_reg_weight, self.point_times)
            pos.append(p_i)
            pgt.append(g_i)
        return pos, pgt

    def _mil_select(self, boxes, n_per_img, gt_labels, roi_feature_map, img_metas, feats=None, hp=None, wp=None):
        """RH:2308-2312: hand the per-layer pseudo boxes to the MIL head and take its layer choice.  boxes [L, n_tot, 4].
        ``mil_fn(boxes_per_img, gt_labels, roi_feature_map, img_metas)`` gets the boxes as the reference builds them (per
        image [n_i, L, 4], RH:2296-2306) and may return the layer index per instance (tensor or per-image list), a pair
        ``(index, losses)``, or the reference's ``_mil_forward_train(..., return_index=True)`` triple
        ``(boxes, losses, index)`` -- see ``mil_from_reference``.  -> (gt_index [n_tot] long on the device, losses dict)."""
        per_img = list(boxes.permute(1, 0, 2).split(list(n_per_img), dim=0))
        if self.mil_fn is not None:
            out = self.mil_fn(per_img, gt_labels, roi_feature_map, img_metas)
        else:                                   # own MIL head (a bound dispatch: survives deepcopy / pickling of the module)
            from . import mil as _mil
            if roi_feature_map is None and feats is not None and feats.is_cuda and not torch.is_grad_enabled():
                # no feature map handed over and nothing to train (pseudo-label generation under no_grad): the selection runs on
                # this repo's kernels, on the ViT tokens themselves = the reference's roi_skip_fpn=True feature map (CFG:58, DETB:122-127)
                out = _mil.mil_select_device(self.mil_head, feats, per_img, gt_labels, hp, wp, self._mil_stride, self._mil_rsize)[0]
            else:
                out = _mil.mil_select(self.mil_head, roi_feature_map, per_img, gt_labels, self._mil_stride, self._mil_rsize)
        losses = {}
        if isinstance(out, tuple) and len(out) == 3:
            _, losses, idx = out
        elif isinstance(out, tuple) and len(out) == 2:
            idx, losses = out
        else:
            idx = out
        idx = torch.cat([g.reshape(-1) for g in idx]) if isinstance(idx, (list, tuple)) else idx.reshape(-1)
        if idx.numel() != sum(n_per_img):
            raise ValueError(f'mil_fn returned {idx.numel()} layer indices for {sum(n_per_img)} instances')
        return idx.to(boxes.device).long(), dict(losses)

    @torch.no_grad()
    def update_fg_map(self, map_cos_fg, map_cos_bg, vit_feat, semantic_centers_coords, obj_num_parts, inst_fg_feat, inst_bg_feat,
                      gt_bboxes, pos_mask_thr):
        """RH:2737-2760, same signature and return value: the second-round aggregation of the part centres into refined
        instance maps (``update_fg_map_single_v3``) and their pseudo masks.  map_cos_fg: list of [n_i,H,W]; vit_feat
        [B,1+N,C] (cls token first); semantic_centers_coords: list of [P_i,2]; obj_num_parts: list of lists; inst_fg_feat /
        inst_bg_feat: the ``seed_pseudo_gt`` outputs; gt_bboxes: list of [n_i,4].
        -> (list of [n_i,H,W] maps on the device, list of uint8 numpy masks)."""
        if hasattr(self.rng, 'next_step'):
            self.rng.next_step()
        n_per_img = [int(m.shape[0]) for m in map_cos_fg]
        H, W = map_cos_fg[0].shape[-2:]
        hp, wp = H // 16, W // 16
        feats = AS.token_major(vit_feat[:, 1:])
        dev = feats.device
        maps, masks = AS.update_fg_maps(torch.cat(list(map_cos_fg)).contiguous(), feats, semantic_centers_coords, obj_num_parts,
                                        inst_fg_feat, inst_bg_feat, torch.cat([b.reshape(-1, 4).float() for b in gt_bboxes]).to(dev).contiguous(),
                                        n_per_img, hp, wp, self.rng, pos_mask_thr=pos_mask_thr)
        m_host = self._mask_buffer(masks.shape)
        m_host.copy_(masks, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return list(maps.split(n_per_img, dim=0)), [m.numpy() for m in m_host.split(n_per_img, dim=0)]

    def seed_pseudo_gt(self, x, img_metas, proposal_list, gt_bboxes, gt_labels, gt_bboxes_ignore=None, gt_masks=None,
                       vit_feat=None, img=None, point_init=None, point_cls=None, point_reg=None, imgs_whwh=None,
                       attns=None, gt_points=None, gt_points_labels=None, roi_feature_map=None, return_mask=False,
                       pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10, corr_size=21, point_adjuster=None,
                       edges=None, obj_tau=0.85, pos_inds=None, gt_index=None):
        if hasattr(self.rng, 'next_step'):
            self.rng.next_step()                # a new set of random draws per call (and per rank), like the reference's stream
        feats = AS.token_major(vit_feat)
        if vit_feat.dim() == 4:
            hp, wp = vit_feat.shape[-2:]
        else:
            hp, wp = x[2].shape[-2:]
        dev = feats.device
        B = feats.shape[0]
        n_prop = point_cls.size(1) if point_cls is not None else attns[-1].shape[1] - 1 - hp * wp
        if pos_inds is None:
            # RH:2237-2257: Hungarian match of the point tokens to the GT points; instances then follow the matched TOKEN
            # order and carry the point / label of the GT they were matched to (get_targets + labels[i][pos_inds])
            if img_metas is not None and all('img_shape' in m for m in img_metas):
                wh = [(m['img_shape'][1], m['img_shape'][0]) for m in img_metas]
            elif imgs_whwh is not None:
                wh = [tuple(v) for v in imgs_whwh.reshape(B, -1)[:, :2].tolist()]
            else:
                wh = [(wp * 16., hp * 16.)] * B
            pos_inds, pos_gt = self.match_points(point_reg, point_cls, gt_points, gt_points_labels, wh)
            gt_points = [gt_points[i][pos_gt[i].to(gt_points[i].device)] for i in range(B)]
            gt_points_labels = [gt_points_labels[i][pos_gt[i].to(gt_points_labels[i].device)] for i in range(B)]
        n_per_img = [int(p.shape[0]) for p in pos_inds]
        n_tot = sum(n_per_img)
        obj_img = AS.instance_image_index(n_per_img, dev)
        pts_cat = torch.cat([p.reshape(-1, 2).float() for p in gt_points])
        pos_cat = torch.cat([p.reshape(-1) for p in pos_inds])
        if pts_cat.is_cuda or pos_cat.is_cuda:
            pts, obj_pt = pts_cat.to(dev).contiguous(), pos_cat.to(dev).to(torch.int32)
        else:
            # host inputs: ONE pinned upload carries the matched point-token index and the GT point (fp32 bit pattern)
            obj_pt, pts_bits = AS._upload_i32([pos_cat.numpy(), pts_cat.contiguous().numpy().view('int32')], dev)
            pts = pts_bits.view(torch.float32)
        lab_sizes = [int(l.numel()) for l in gt_points_labels[:B]]
        labels = list(torch.cat([l.reshape(-1) for l in gt_points_labels[:B]]).to(dev, non_blocking=True).split(lab_sizes))
        # ^ RH:2269 labels[i][pos_inds]: one label per matched GT
        mil_losses = {}
        # A5-A7
        rows = AS.rollout_rows(list(attns[-self.cam_layer:]), n_prop)
        cams, mm = AS.cam_maps(rows, obj_img, obj_pt, hp, wp)
        ar = torch.arange(n_tot, device=dev)
        if gt_index is not None:
            gt_index = torch.cat([g.reshape(-1) for g in gt_index]) if isinstance(gt_index, (list, tuple)) else gt_index
            gt_index = gt_index.to(dev).long()
            # seed-candidate counts start their trip to the host now and overlap the connected-components stage
            begun = AS.refined_maps_begin(cams[gt_index, ar].contiguous(), mm[gt_index, ar].contiguous(), n_per_img, hp, wp)
        boxes, _ = AS.cam_bbox(cams, mm, pts, hp, wp, self.seed_thr, self.seed_multiple)
        if gt_index is None:
            if self.mil_fn is None and not hasattr(self, 'mil_head'):
                raise ValueError('seed_pseudo_gt needs gt_index= or a mil_fn (MIL head is outside the hot path)')
            gt_index, mil_losses = self._mil_select(boxes, n_per_img, labels, roi_feature_map, img_metas, feats, hp, wp)
            begun = AS.refined_maps_begin(cams[gt_index, ar].contiguous(), mm[gt_index, ar].contiguous(), n_per_img, hp, wp)
        pseudo_boxes = boxes[gt_index, ar].contiguous()                # RH:2965-2967 gather of the chosen layer's box
        # A8, A13
        rm = AS.refined_maps(begun['cam_low'], begun['cam_mm'], feats, n_per_img, pseudo_boxes, pts, hp, wp, self.rng,
                             refine_times=2, obj_tau=obj_tau, mask_thr=pos_mask_thr, begun=begun)
        # A12 candidates are counted on the device while A9-A11 are enqueued; the host only then waits for the counts
        mp = AS.mask_points_begin(rm['map_fg'], rm['map_bg'], pseudo_boxes, pos_thr=pos_mask_thr, neg_thr=neg_mask_thr,
                                  corr_size=corr_size)
        parts = AS.semantic_parts(rm['map_fg'], feats, obj_img, pseudo_boxes, hp, wp, pos_thr=pos_mask_thr,
                                  n_shift=self.mean_shift_times_local, n_points=self.n_seeds,
                                  num_semantic_points=self.num_semantic_points, n_per_img=n_per_img)
        coords, plabels = AS.mask_points(rm['map_fg'], rm['map_bg'], pseudo_boxes, n_per_img, self.rng, pos_thr=pos_mask_thr,
                                         neg_thr=neg_mask_thr, num_gt=num_mask_point_gt, corr_size=corr_size, begun=mp)
        per_img = AS.assemble_parts(parts, n_per_img, labels, hp, wp)
        split = lambda t: list(t.split(n_per_img, dim=0))
        if return_mask:                                         # RH:2358 hand-off: uint8 numpy masks on the host
            m_dev = rm['mask']
            # one transfer into a persistent pinned buffer (two alternate, see _mask_buffer) instead of one pageable copy per
            # image; the numpy arrays are views of it
            m_host = self._mask_buffer(m_dev.shape)
            m_host.copy_(m_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            masks = [m.numpy() for m in m_host.split(n_per_img, dim=0)]
        else:
            masks = split(rm['mask'])
        grp = rm['groups']
        fg_feat = [rm['centroid'][g, :n + 1].reshape(n + 1, -1, 1, 1) for g, n in enumerate(n_per_img)]
        bg_feat = [rm['centroid'][g, n + 1:2 * n + 1].reshape(n, -1, 1, 1) for g, n in enumerate(n_per_img)]
        out = dict(pseudo_gt_labels=labels, pseudo_gt_bboxes=split(pseudo_boxes), mil_losses=mil_losses,
                   best_attn_idx=split(gt_index), map_cos_fg=split(rm['map_fg']),
                   mask_points_coords=split(coords), mask_points_labels=split(plabels),
                   semantic_centers=[p['semantic_centers'] for p in per_img],
                   semantic_centers_split=[p['semantic_centers_split'] for p in per_img],
                   semantic_centers_feat_split=[p['semantic_centers_feat_split'] for p in per_img],
                   semantic_centers_feat=[p['semantic_centers_feat'] for p in per_img],
                   num_parts=[p['num_parts'] for p in per_img],
                   semantic_centers_org=([p['semantic_centers_org'][0] for p in per_img],
                                         [p['semantic_centers_org'][1] for p in per_img]),
                   pseudo_gt_masks=masks, corres_gts=[p['corres_gts'] for p in per_img],
                   inst_fg_feat=fg_feat, inst_bg_feat=bg_feat)
        if self.visualize:
            out.update(map_cos_bg=split(rm['map_bg']), sim_fg=[p['sim_fg'] for p in per_img], attns=cams,
                       points_bg=rm['pts'], points_fg=rm['pts'])
        self._last = dict(rows=rows, cams=cams, boxes=boxes, refined=rm, parts=parts)
        return out


def mil_from_reference(ref_head):
    """``mil_fn`` that defers to the REFERENCE RoI head's MIL stage (``_mil_forward_train``, RH:2953-2972: RoIAlign over the
    7 x n_gt pseudo boxes + MAEBoxHeadMIL), for running this repo's ``seed_pseudo_gt`` inside the reference detector:

        fast = build_head(dict(cfg.model.roi_head, train_cfg=cfg.model.train_cfg.rcnn, mil_fn=mil_from_reference(ref_head)))
        ref_head.seed_pseudo_gt = fast.seed_pseudo_gt          # losses / test-time methods stay the reference's
    """
    def fn(boxes_per_img, gt_labels, roi_feature_map, img_metas):
        return ref_head._mil_forward_train(roi_feature_map, None, boxes_per_img, gt_labels, img_metas, return_index=True)
    return fn


REFERENCE_NAMES = ('AttnShiftRoIHead', 'StandardRoIHeadMaskPointSampleDeformAttnReppoints')


def attach(ref_head, **fast_kwargs):
    """Give an INSTANCE of the reference RoI head this repo's device path: ``seed_pseudo_gt`` and ``update_fg_map`` are replaced,
    everything else (losses, test-time methods, sub-heads, ``init_weights``) stays the reference's.  The fast head reads its
    knobs off the reference instance and defers the MIL stage to the reference's own ``_mil_forward_train``."""
    kw = dict(bbox_head=dict(cam_layer=getattr(getattr(ref_head, 'bbox_head', None), 'cam_layer', 7),
                             seed_thr=getattr(getattr(ref_head, 'bbox_head', None), 'seed_thr', 0.2),
                             seed_multiple=getattr(getattr(ref_head, 'bbox_head', None), 'seed_multiple', 0.5)),
              train_cfg=getattr(ref_head, 'train_cfg', None), num_semantic_points=getattr(ref_head, 'num_semantic_points', 3),
              mean_shift_times_local=getattr(ref_head, 'mean_shift_times_local', 10),
              mil_fn=mil_from_reference(ref_head) if hasattr(ref_head, '_mil_forward_train') else None)
    kw.update(fast_kwargs)
    fast = AttnShiftRoIHead(**kw)
    object.__setattr__(ref_head, '_as_b200', fast)          # not a registered submodule: it owns no parameters
    ref_head.seed_pseudo_gt = fast.seed_pseudo_gt
    ref_head.update_fg_map = fast.update_fg_map
    return ref_head


def make_dropin(ref_cls):
    """Class-level version of ``attach`` for the registry: a subclass of the reference head whose constructor runs the
    reference's and then swaps in the device path.  ``HEADS.register_module(name='AttnShiftRoIHead', module=make_dropin(Ref))``
    makes configs/mae build the full head (losses included) with the fast ``seed_pseudo_gt``."""
    class AttnShiftRoIHeadDropIn(ref_cls):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            attach(self)
    AttnShiftRoIHeadDropIn.__name__ = 'AttnShiftRoIHead'
    return AttnShiftRoIHeadDropIn


def _register():
    """Registry policy (mmdet/models/builder.py:6-12).  Always: ``AttnShiftRoIHeadB200`` = this module's class.  Without mmdet
    (the shim registry of this repo) there is no reference head, so the reference's two names build it as well.  Inside a
    real mmdet the reference class keeps its name; the config's ``AttnShiftRoIHead`` (which the reference itself never
    registers -- its rename is incomplete, roi_heads/__init__.py:24) becomes the reference class + device path when the
    reference class is there, and is left alone otherwise."""
    from .registry import USING_MMDET
    HEADS.register_module(name='AttnShiftRoIHeadB200', force=True, module=AttnShiftRoIHead)
    if not USING_MMDET:
        HEADS.register_module(name=list(REFERENCE_NAMES), force=True, module=AttnShiftRoIHead)
        return
    ref = HEADS.get(REFERENCE_NAMES[1])                     # pragma: no cover - needs mmdet
    if ref is not None and HEADS.get(REFERENCE_NAMES[0]) is None:
        HEADS.register_module(name=REFERENCE_NAMES[0], module=make_dropin(ref))


_register()

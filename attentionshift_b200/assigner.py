"""Point-token <-> GT matching of ``seed_pseudo_gt`` (RH:2237-2257): the reference's ``HungarianPointAssigner``
(mmdet/core/bbox/assigners/hungarian_point_assigner.py:53-109) with its configured costs
(configs/mae/attnshift_voc12aug.py:182-187: ``FocalLossCost`` weight 1 + ``PointL1Cost`` weight 10,
mmdet/core/bbox/match_costs/match_cost.py:52-106) followed by ``PointPseudoSampler``
(mmdet/core/bbox/samplers/point_pseudo_sampler.py:34-37).

Host logic like the reference (which copies the cost matrix to the CPU and calls scipy's ``linear_sum_assignment``); it is
adjacent to the device hot path and decides which point-token rows of the roll-out become CAMs.
"""
import torch


def focal_loss_cost(cls_pred, gt_labels, weight=1.0, alpha=0.25, gamma=2, eps=1e-12):
    """match_cost.py:90-106.  cls_pred [P, n_cls] logits, gt_labels [G] -> [P, G]."""
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    return (pos[:, gt_labels] - neg[:, gt_labels]) * weight


def point_l1_cost(point_pred, gt_points_norm, weight=1.0):
    """match_cost.py:56-58."""
    return torch.cdist(point_pred, gt_points_norm, p=1) * weight


def hungarian_point_assign(point_pred, cls_pred, gt_points, gt_labels, img_wh, cls_weight=1.0, reg_weight=10.0, times=1):
    """-> (pos_inds [k] ascending proposal indices, pos_gt [k] index of the GT each one is matched to).
    point_pred [P,2] predicted points in [0,1]; cls_pred [P,n_cls] logits; gt_points [G,2] pixels; gt_labels [G];
    img_wh = (w, h) of img_meta['img_shape'].  ``times`` > 1 repeats the matching on the not yet taken proposals
    (hungarian_point_assigner.py:111-138)."""
    from scipy.optimize import linear_sum_assignment
    P, G = point_pred.shape[0], gt_points.shape[0]
    assigned = torch.full((P,), -1, dtype=torch.long)
    if G == 0 or P == 0:
        if G == 0:
            assigned[:] = 0
        return torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)
    factor = gt_points.new_tensor([float(img_wh[0]), float(img_wh[1])]).unsqueeze(0)
    cost = focal_loss_cost(cls_pred, gt_labels, cls_weight) + point_l1_cost(point_pred, gt_points / factor, reg_weight)
    cost = cost.detach().cpu()
    assigned[:] = 0
    for _ in range(max(int(times), 1)):
        rows, cols = linear_sum_assignment(cost)
        rows, cols = torch.from_numpy(rows), torch.from_numpy(cols)
        if times > 1:
            cost[rows] += 1000
        assigned[rows] = cols + 1
    pos_inds = torch.nonzero(assigned > 0, as_tuple=False).squeeze(-1).unique()
    return pos_inds, assigned[pos_inds] - 1


DEVICE_MATCH_MAX = 512          # as_hungarian_points: proposals and GTs per image (shared-memory tables of the solver)


def hungarian_point_assign_device(point_pred, cls_pred, gt_points, gt_labels, imgs_wh, cls_weight=1.0, reg_weight=10.0,
                                  want_status=False):
    """The same matching for a whole batch without leaving the device (``as_hungarian_points``: the shortest-augmenting-path
    solver scipy uses, fp64, one warp per image) -- no ``cost.cpu()`` sync in the middle of ``seed_pseudo_gt``.
    point_pred [B,P,2], cls_pred [B,P,n_cls] on the GPU; gt_points / gt_labels: per-image lists; imgs_wh: per-image (w, h).
    -> (pos_inds, pos_gt): per-image int64 device tensors of min(P, G_i) entries, as ``hungarian_point_assign`` returns them.
    The costs are the reference's element-wise formulas (match_cost.py:56-58, 90-106) evaluated for each GT's own image only."""
    from . import lib as _l
    dev = point_pred.device
    B, P = point_pred.shape[:2]
    n_g = [int(g.shape[0]) for g in gt_points[:B]]
    first = [0]
    for n in n_g[:-1]:
        first.append(first[-1] + n)
    tot = sum(n_g)
    if tot == 0 or P == 0:
        empty = torch.zeros(0, dtype=torch.long, device=dev)
        r = ([empty] * B, [empty] * B)
        return r + (torch.zeros(B, dtype=torch.int32, device=dev),) if want_status else r
    img_of = torch.repeat_interleave(torch.arange(B), torch.tensor(n_g)).to(dev, non_blocking=True)
    pts = torch.cat([g.reshape(-1, 2).float() for g in gt_points[:B]]).to(dev, non_blocking=True)
    lab = torch.cat([l.reshape(-1).long() for l in gt_labels[:B]]).to(dev, non_blocking=True)
    wh = torch.tensor([[float(w), float(h)] for w, h in imgs_wh[:B]]).to(dev, non_blocking=True)
    p = cls_pred.detach().float()[img_of, :, lab].sigmoid()                              # [tot, P]
    eps, alpha, gamma = 1e-12, 0.25, 2
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    reg = (point_pred.detach().float()[img_of] - (pts / wh[img_of]).unsqueeze(1)).abs().sum(-1)      # [tot, P]
    cost = ((pos - neg) * cls_weight + reg * reg_weight).contiguous()
    meta = torch.tensor([first, n_g], dtype=torch.int32).to(dev, non_blocking=True)
    out = torch.empty(3, max(tot, B), dtype=torch.int32, device=dev)
    L = _l.load()
    _l.check(L.as_hungarian_points(_l.ptr(cost), _l.ptr(meta[0]), _l.ptr(meta[1]), B, P, max(n_g), _l.ptr(out[0]), _l.ptr(out[1]),
                                   _l.ptr(out[2]), _l.stream_ptr()), 'as_hungarian_points')
    pos_inds, pos_gt = [], []
    for i in range(B):
        k = min(P, n_g[i])
        pos_inds.append(out[0, first[i]:first[i] + k].long())
        pos_gt.append(out[1, first[i]:first[i] + k].long())
    return (pos_inds, pos_gt, out[2, :B]) if want_status else (pos_inds, pos_gt)

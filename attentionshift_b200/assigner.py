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

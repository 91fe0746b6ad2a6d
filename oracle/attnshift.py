"""CPU oracle (torch fp32) for the attention-shift half of the hot path.

TEST INFRASTRUCTURE ONLY -- never imported by ``attentionshift_b200``.

A from-scratch restatement of the algorithms in the reference file
``mmdet/models/roi_heads/stdroi_point_deform_attn_reppoints.py`` (abbreviated RH
below, line numbers as of reference commit dc3b87d).  It deliberately keeps the
same torch primitive per arithmetic step (``F.cosine_similarity``, ``softmax``,
``matmul`` ...) so that on CPU it is bit-identical to the reference; that equality
is asserted by ``tests/test_oracle_vs_reference.py`` wherever ``/root/reference``
exists and frozen into ``tests/golden/*.pt`` for everywhere else.

Parity status per stage:
  rollout, refined maps, grid seeds, mean-shift, part filtering/merging, part
  centres, mask points, pseudo masks : pinned against the reference run in the
  build container.
  connected components (``ccl_label``): PARITY UNPINNED -- ``cc_torch`` is an
  un-vendored third-party CUDA extension (no source, no pinned version, no tests
  in the reference).  Only the partition of foreground pixels matters to the caller
  (RH:69-83); we assume 8-connectivity (the upstream project is a block-based
  union-find after YACCLAB's BUF, which is 8-connected).

Layout conventions follow the reference: ``vit_feat`` is ``[C, Hp, Wp]``, maps are
``[n, H, W]``, points are ``(x, y)`` pixel coordinates, boxes are ``x1,y1,x2,y2``.
"""
import numpy as np
import torch
import torch.nn.functional as F

PATCH = 16


# --------------------------------------------------------------------------- A7
def ccl_label(binary):
    """8-connected component labels of a 2-D 0/1 array (stand-in for cc_torch,
    call site RH:68).  Returns int32 labels, 0 = background; label VALUES are not
    part of the contract, only the partition."""
    from scipy import ndimage
    lab, _ = ndimage.label(np.asarray(binary) != 0, structure=np.ones((3, 3), dtype=np.int32))
    return lab.astype(np.int32)


def bbox_from_cam(cam, point, cam_thr=0.2, area_ratio=0.5, img_size=None):
    """RH:60-116 ``get_bbox_from_cam_fast`` with box_method='expand'.

    cam [H,W] fp32, point (x, y).  -> (bbox [1,4] fp32, kept-component mask [H,W] bool).
    Min-max normalise, binarise at cam_thr, label components, keep components whose
    area >= area_ratio * largest area, take their joint extent, then mirror the far
    side of that extent around the GT point (clipped to the image)."""
    img_h, img_w = img_size
    cam = (cam - cam.min()) / (cam.max() - cam.min()).clamp(1e-6)
    binary = (cam >= cam_thr)
    labels = torch.from_numpy(ccl_label(binary.numpy().astype(np.uint8)))
    ids = labels.unique()
    ids = ids[ids != 0]
    if len(ids) == 0:
        # the reference would raise on torch.stack([]); a CAM always has its own
        # maximum >= thr so this cannot happen for cam_thr <= 1.
        raise ValueError("empty CAM")
    areas = torch.stack([(labels == i).sum() for i in ids])
    keep_ids = ids[areas >= area_ratio * areas.max()]
    keep = torch.zeros_like(binary)
    for i in keep_ids:
        keep |= (labels == i)
    ys, xs = torch.nonzero(keep, as_tuple=True)
    xs = xs.to(torch.float32)
    ys = ys.to(torch.float32)
    px0, py0, px1, py1 = xs.min(), ys.min(), xs.max(), ys.max()
    xc, yc = point
    if abs(xc - px0) > abs(xc - px1):
        x0 = px0
        x1 = xc * 2 - x0
        x1 = x1 if x1 < img_w else float(img_w)
    else:
        x1 = px1
        x0 = xc * 2 - x1
        x0 = x0 if x0 > 0 else 0.0
    if abs(yc - py0) > abs(yc - py1):
        y0 = py0
        y1 = yc * 2 - y0
        y1 = y1 if y1 < img_h else float(img_h)
    else:
        y1 = py1
        y0 = yc * 2 - y1
        y0 = y0 if y0 > 0 else 0.0
    return cam.new_tensor([[x0, y0, x1, y1]]), keep


# --------------------------------------------------------------------------- A5 / A6
def rollout(attns):
    """RH:1257-1272 ``attns_project_to_feature``.  attns: list of L tensors [B,T,T]
    (head-mean attention of the last L layers, oldest first).  -> [B, L, T, T] where
    index 0 is the last layer alone and index L-1 the product over all L layers."""
    a = torch.stack(attns)                                   # [L,B,T,T]
    eye = torch.eye(a.size(2), dtype=a.dtype)
    aug = a + eye
    aug = aug / aug.sum(-1).unsqueeze(-1)
    joint = torch.zeros_like(aug)
    joint[-1] = aug[-1]
    for i in range(2, len(attns) + 1):
        joint[-i] = torch.matmul(joint[-(i - 1)], aug[-i])
    return joint.flip(0).permute(1, 0, 2, 3)


def rollout_rows(attns, n_rows):
    """Same arithmetic restricted to the last ``n_rows`` rows -- the only rows the
    callers read (RH:2272).  Row i of a matrix product depends only on row i of the
    left factor, so this equals ``rollout(attns)[:, :, -n_rows:, :]`` up to the
    GEMM's internal summation order.  -> [B, L, n_rows, T]."""
    a = torch.stack(attns)
    eye = torch.eye(a.size(2), dtype=a.dtype)
    aug = a + eye
    aug = aug / aug.sum(-1).unsqueeze(-1)
    out = [aug[-1][:, -n_rows:, :]]
    for i in range(2, len(attns) + 1):
        out.append(torch.matmul(out[-1], aug[-i]))
    return torch.stack(out, dim=1)


def cams_from_rollout(rows, pos_inds, n_point_tokens, hp, wp):
    """RH:2272-2275.  rows [L, n_point_tokens, T] of one image -> (low-res CAMs
    [L, n_gt, hp, wp], upsampled CAMs [L, n_gt, 16hp, 16wp] bilinear, align_corners=False)."""
    cams = rows[:, :, 1:-n_point_tokens].permute(1, 0, 2)[pos_inds].permute(1, 0, 2)
    n_gt = cams.shape[1]
    low = cams.reshape(-1, 1, hp, wp)
    up = F.interpolate(low, (hp * PATCH, wp * PATCH), mode='bilinear')
    return low.reshape(-1, n_gt, hp, wp), up.reshape(-1, n_gt, hp * PATCH, wp * PATCH)


# --------------------------------------------------------------------------- A8
def box_to_mask(bboxes, size, default=0.5):
    """RH:303-309 ``box2mask`` -- inclusive integer box painted with 1 over ``default``."""
    n = bboxes.shape[0]
    m = torch.zeros(n, size[0], size[1], dtype=bboxes.dtype) + default
    for i in range(n):
        b = bboxes[i]
        m[i, int(b[1]):int(b[3] + 1), int(b[0]):int(b[2] + 1)] = 1.0
    return m


def norm_maps(maps):
    """RH:329-333 per-map min-max normalisation (no epsilon)."""
    n = maps.shape[0]
    mx = maps.view(n, -1, 1).max(dim=1, keepdim=True)[0]
    mn = maps.view(n, -1, 1).min(dim=1, keepdim=True)[0]
    return (maps - mn) / (mx - mn)


def sample_points(maps, num_points=10, thr=0.2, is_pos=False, gt_points=None, hook=None, keys=None):
    """RH:343-371 ``sample_point_grid``: per map, ``num_points`` random pixels (x, y)
    among those >= thr (is_pos) or < thr.  Consumes the torch CPU default generator
    exactly like the reference: one ``torch.randint(num, (len(arange(0,num,step)),))``
    per map that has enough candidates."""
    out = []
    for i, m in enumerate(maps):
        factor = 1.0
        cand = ((m >= thr * factor) if is_pos else (m < thr * factor)).nonzero(as_tuple=False).view(-1, 2)
        num = cand.shape[0]
        if num < num_points:
            if is_pos:
                out.append(torch.cat((cand, gt_points[i].repeat(num_points - num, 1)), dim=0))
                continue
            while num < num_points:
                factor *= 2
                cand = (m < thr * factor).nonzero(as_tuple=False)
                num = cand.shape[0]
        step = num // num_points
        if hook is not None:
            hook(keys[i])                 # test harness only: re-seed per (image, stage, slot) -- see KeyedRng
        n_draw = torch.arange(0, num, step=step).shape
        idx = torch.randint(num, n_draw) % num
        out.append(cand[idx][:num_points])
    return torch.stack(out).flip(-1)


def seed_prototype_map(point_coords, vit_feat):
    """RH:335-341 + RH:312-320: mean feature of the sampled pixels' patches, then
    cosine similarity of that prototype with every patch.  point_coords [K,P,2] (x,y)
    pixels; vit_feat [C,Hp,Wp].  -> [K,Hp,Wp]."""
    c, hp, wp = vit_feat.shape
    tok = vit_feat.permute(1, 2, 0)                           # [Hp,Wp,C]
    iy = (point_coords[..., 1].long() // PATCH).clamp(0, hp)
    ix = (point_coords[..., 0].long() // PATCH).clamp(0, wp)
    pf = tok[iy.flatten(), ix.flatten()].unflatten(0, iy.shape)  # [K,P,C]
    proto = pf.mean(dim=1, keepdim=True)                      # [K,1,C]
    tokens = tok.reshape(1, hp * wp, c).expand(point_coords.shape[0], -1, -1)
    sim = F.cosine_similarity(tokens, proto, dim=2)
    return sim.unflatten(1, (hp, wp))


def refined_similarity(point_coords, vit_feat, bboxes, refine_times=1, tau=0.85, is_select=False):
    """RH:668-707 ``get_refined_similarity``.

    -> (maps [refine_times+1, K, Hp, Wp], centroid feature [K, C, 1, 1]).
    Each refinement: zero the affinities below tau*rowmax, affinity-weighted mean of
    the patch features, cosine map of that centroid.  With ``is_select`` the first
    n_obj maps are multiplied by their (patch-grid) box mask and every emitted map
    keeps a patch only in the row that attains the cross-seed argmax."""
    feats = vit_feat[None]
    cos0 = seed_prototype_map(point_coords, vit_feat)
    cur = cos0.clone()
    n_obj = bboxes.shape[0]
    bmask = box_to_mask(bboxes // PATCH, cos0.shape[-2:], default=0)
    outs = []

    def emit(m):
        if is_select:
            m[:n_obj] = m[:n_obj] * bmask
            win = m.argmax(0, keepdim=True).expand_as(m)
            rows = torch.arange(m.shape[0])
            outs.append(torch.where(win == rows[:, None, None], m.clone(), torch.zeros_like(m)))
        else:
            outs.append(m.clone())

    emit(cos0)
    centroid = None
    for _ in range(refine_times):
        mx = cur.flatten(1).max(1, keepdim=True)[0].unsqueeze(-1)
        cur[cur < mx * tau] *= 0
        w = cur.unsqueeze(1)
        centroid = (feats * w).sum([2, 3], keepdim=True) / w.sum([2, 3], keepdim=True).clamp(1e-8)
        cur = F.cosine_similarity(feats, centroid, dim=1)
        emit(cur)
    return torch.stack(outs), centroid


def _normalize_map(m):
    """RH:1037-1040."""
    return m / (m.flatten(-2).max(-1, keepdim=True)[0].unsqueeze(-1) + 1e-8)


def refined_maps(attn_maps, vit_feat, bboxes, thr_pos=0.2, thr_neg=0.1, num_points=20,
                 refine_times=1, obj_tau=0.85, gt_points=None, hook=None, img=0):
    """RH:1000-1019 ``get_cosine_similarity_refined_map`` (+ RH:1042-1046).

    attn_maps [n_obj,H,W]; -> (fg [R+1,n_obj,H,W], bg [R+1,n_obj,H,W], points_fg
    [n_obj+1,P,2], points_bg [n_obj,P,2], fg_feat, bg_feat).  RNG draw order is
    bg, fg, bg_supp -- identical to the reference."""
    an = norm_maps(attn_maps)
    n = an.shape[0]
    pts_bg = sample_points(an, thr=thr_neg, num_points=num_points, hook=hook, keys=[(img, 0, n + 1 + j) for j in range(n)])
    pts_fg = sample_points(an, thr=thr_pos, num_points=num_points, is_pos=True, gt_points=gt_points, hook=hook,
                           keys=[(img, 0, j) for j in range(n)])
    pts_supp = sample_points(an.mean(0, keepdim=True), thr=thr_neg, num_points=num_points, hook=hook, keys=[(img, 0, n)])
    pts_fg = torch.cat((pts_fg, pts_supp), dim=0)
    sim_fg, fg_feat = refined_similarity(pts_fg, vit_feat, bboxes, refine_times, obj_tau, is_select=True)
    sim_bg, bg_feat = refined_similarity(pts_bg, vit_feat, bboxes, refine_times, obj_tau)
    size = attn_maps.shape[-2:]
    sim_fg = F.interpolate(sim_fg, size, mode='bilinear')[:, :an.shape[0]]
    sim_bg = F.interpolate(sim_bg, size, mode='bilinear')
    fused = (1 - sim_bg) * sim_fg
    fmax = fused.flatten(-2, -1).max(-1, keepdim=True)[0].unsqueeze(-1).clamp(1e-8)
    bg = _normalize_map(sim_bg.clone())
    fg = _normalize_map(fused.clone())
    bg = bg + (1 - (fg * 0.5 + bg * 0.5))
    bmax = bg.flatten(-2, -1).max(-1, keepdim=True)[0].unsqueeze(-1).clamp(1e-8)
    return fused / fmax, bg / bmax, pts_fg, pts_bg, fg_feat, bg_feat


# --------------------------------------------------------------------------- A12
def erode(m, k):
    """RH:1182-1187 / RH:145-146: min-pool k x k, stride 1, pad k//2 (pad value -inf
    for the negated max-pool, i.e. padding never wins)."""
    shape = m.shape
    x = m.reshape(1, -1, shape[-2], shape[-1]) if m.ndim >= 3 else m.reshape(1, 1, *shape)
    return (-F.max_pool2d(-x, k, 1, k // 2)).reshape(shape)


def fill_index(idx, n):
    """RH:1147-1155 ``fill_in_idx``: repeat a short index list up to length n."""
    assert idx.shape[0] != 0
    if idx.shape[0] >= n / 2:
        return torch.cat((idx, idx[:n - idx.shape[0]]), dim=0)
    idx = idx.repeat(n // idx.shape[0], 1)
    return fill_index(idx, n)


def mask_points_in_box(map_fg, map_bg, pos_thr, neg_thr, num_gt, corr_size, hook=None, key=None):
    """RH:433-461 ``get_mask_points_single_box_cos_map_fg_bg`` on a box crop:
    candidates = eroded(fg > max*pos_thr) pixels (label 1) followed by
    (bg > max*neg_thr) pixels (label 0); pick ``torch.randperm(n)[:num_gt]``."""
    pos = erode((map_fg > map_fg.max() * pos_thr).float(), corr_size).nonzero(as_tuple=False)
    neg = (map_bg > map_bg.max() * neg_thr).nonzero(as_tuple=False)
    both = torch.cat((pos, neg), dim=0)
    if hook is not None:
        hook(key)
    chosen = torch.randperm(both.shape[0])[:num_gt]
    lab = torch.cat((torch.ones(pos.shape[0], dtype=torch.bool), torch.zeros(neg.shape[0], dtype=torch.bool)))
    if chosen.shape[0] < num_gt:
        if chosen.shape[0] == 0:
            return -torch.ones(num_gt, 2), torch.zeros(num_gt, dtype=torch.bool)
        chosen = fill_index(chosen, num_gt)
    return both[chosen], lab[chosen]


def mask_sample_points(attn, rois, attn_idx, vit_feat, pos_thr=0.6, neg_thr=0.6, num_gt=20,
                       corr_size=21, refine_times=2, obj_tau=0.85, gt_points=None, hook=None, img=0):
    """RH:1966-1993.  attn [L,n_obj,H,W] upsampled CAMs; attn_idx [n_obj] chosen layer
    per instance.  -> (coords [n_obj,num_gt,2] (x,y), labels [n_obj,num_gt] bool,
    map_fg, map_bg, points_bg(sic: fg), points_fg(sic: bg), feats_fg, feats_bg)
    -- the reference unpacks the two point tensors in swapped order and we keep that."""
    n = attn.shape[1]
    amap = attn.detach().clone()[attn_idx, torch.arange(n)]
    fg, bg, p_fg, p_bg, f_fg, f_bg = refined_maps(amap, vit_feat, rois, thr_pos=0.2, thr_neg=0.1,
                                                  num_points=20, refine_times=refine_times,
                                                  obj_tau=obj_tau, gt_points=gt_points, hook=hook, img=img)
    coords, labels = [], []
    for i in range(fg[0].shape[0]):
        x0, y0, x1, y1 = rois[i].int().tolist()
        c, l = mask_points_in_box(fg[-1][i][y0:y1, x0:x1], bg[-1][i][y0:y1, x0:x1],
                                  pos_thr, neg_thr, num_gt, corr_size, hook=hook, key=(img, 1, i))
        c[:, 0] += y0
        c[:, 1] += x0
        coords.append(c.flip(1))
        labels.append(l)
    return torch.stack(coords).float(), torch.stack(labels), fg, bg, p_fg, p_bg, f_fg, f_bg


# --------------------------------------------------------------------------- A9 / A10
def grid_seed_coords(maps, rois=None, thr=0.35, n_points=20):
    """RH:1786-1808 deterministic seed grid: every (num_pos // n_points)-th positive
    patch in row-major order; fewer than n_points -> repeat-fill; none -> box centre.
    maps [n_obj,Hp,Wp].  -> long [n_obj,n_points,2] as (row, col)."""
    sel = []
    for i, m in enumerate(maps):
        pos = m >= thr
        num = pos.sum()
        idx = pos.nonzero()
        if num >= n_points:
            c = idx[torch.arange(0, num, step=num // n_points)[:n_points]]
        elif num > 0:
            c = fill_index(idx, n_points)
        elif rois is not None:
            c = ((rois[i][:2] + rois[i][2:]) // (2 * PATCH)).long().view(1, 2).flip(1).repeat(n_points, 1)
        else:
            pos = m >= 0
            num = pos.sum()
            idx = pos.nonzero()
            c = idx[torch.arange(0, num, step=num // n_points)[:n_points]]
        sel.append(c)
    return torch.stack(sel)


def seed_density(prot, feats, onehot):
    """RH:882-908 ``update_density_batch``: per seed 1 - mean cosine to its assigned
    tokens (no assigned token -> 1), clamped at 1e-10.  -> [n_obj,S,1]."""
    sim = F.cosine_similarity(prot[:, :, None], feats[:, None], dim=-1)
    tot = (sim * onehot).sum(-1)
    cnt = onehot.sum(-1)
    d = 1 - torch.where(cnt >= 1, tot / cnt, torch.zeros_like(tot))
    return d.clamp(1e-10).unsqueeze(-1)


def mean_shift(prot, feats, feats_org, tau=0.1, temp=0.1, n_shift=5, trace=None):
    """RH:830-854 ``cosine_shift_batch`` -- the attention-shift loop.

    prot [n_obj,S,C]; feats [n_obj,N,C] (box-masked token copies); feats_org [N,C].
    Per iteration: cosine affinity, softmax over tokens with temperature temp*tau,
    hard-assign each token to the seed with the largest weight (first wins ties),
    new seed = sum of its assigned tokens weighted by its own softmax weights
    (un-normalised), tau <- per-seed density.  -> (prot [n_obj*S,C], sim [n_obj*S,N])
    where sim is against the UNMASKED tokens.  ``trace`` (a list) receives the
    per-iteration assignment [n_obj,N] for index-parity checks."""
    for _ in range(n_shift):
        sim = F.cosine_similarity(prot[:, :, None], feats[:, None], dim=-1)
        w = F.softmax(sim / (temp * tau), dim=-1)
        assign = w.argmax(1, keepdim=True)
        rows = torch.arange(prot.shape[1], dtype=assign.dtype)[None, :, None].expand(prot.shape[0], prot.shape[1], -1)
        onehot = torch.where(rows == assign, torch.ones_like(w), torch.zeros_like(w))
        prot = torch.matmul(w * onehot, feats)
        tau = seed_density(prot, feats, onehot)
        if trace is not None:
            trace.append(assign[:, 0].clone())
    sim = F.cosine_similarity(prot[:, :, None, :], feats_org[None, None, :, :], dim=-1)
    return prot.flatten(0, 1), sim.flatten(0, 1)


def mean_shift_from_maps(maps, vit_feat, rois, thr=0.35, n_shift=5, tau=0.1, temp=0.1, n_points=20, trace=None):
    """RH:1778-1840 ``mean_shift_grid_prototype`` (rois given).  -> (prot [n_obj*S,C],
    sim [n_obj*S,Hp,Wp] clamped at 0)."""
    c, hp, wp = vit_feat.shape
    sel = grid_seed_coords(maps, rois, thr, n_points)
    tok = vit_feat.permute(1, 2, 0)
    prot = tok[sel[..., 0].flatten(), sel[..., 1].flatten()].unflatten(0, sel.shape[:2]).clone()
    bm = box_to_mask(rois // PATCH, (hp, wp), default=0)
    feats = (vit_feat[None] * bm[:, None]).flatten(-2).transpose(1, 2).clone()
    feats_org = vit_feat.flatten(-2).transpose(0, 1).clone()
    prot, sim = mean_shift(prot.clone(), feats, feats_org, tau=tau, temp=temp, n_shift=n_shift, trace=trace)
    return prot, sim.unflatten(-1, (hp, wp)).clamp(0)


# --------------------------------------------------------------------------- A11
def filter_seed_maps(maps, pos_maps, pos_thr=0.85):
    """RH:265-275: keep a seed when the mean of the fg map over its (sim > 0.8) region
    is >= pos_thr.  maps [n_obj,S,Hp,Wp], pos_maps [n_obj,Hp,Wp] -> keep [n_obj,S] bool."""
    fore = torch.where(maps > 0.8, torch.ones_like(maps), torch.zeros_like(maps))
    score = (pos_maps[:, None] * fore).sum(dim=[-2, -1]) / fore.sum(dim=[-2, -1]).clamp(1e-6)
    return score >= pos_thr


def merge_prototypes(protos, thr=0.95):
    """RH:278-294 greedy, order-dependent merge: walk seeds in order; seed i absorbs
    every not-yet-absorbed seed j >= i with cos >= thr (mean of the absorbed set)."""
    out = []
    for p in protos:
        if p.shape[0] == 0:
            out.append([])
            continue
        sim = F.cosine_similarity(p[None], p[:, None], dim=-1)
        live = torch.where(torch.triu(sim, diagonal=0) >= thr, torch.ones_like(sim), torch.zeros_like(sim))
        merged = []
        for i in range(live.shape[0]):
            w = live[i]
            if w.sum() > 0:
                merged.append(torch.matmul(w, p) / (w.sum() + 1e-8))
            live[w > 0] *= 0
        out.append(torch.stack(merged))
    return out


def part_maps(proto, tokens_hwc):
    """RH:297-301 ``cal_similarity``."""
    if isinstance(proto, list):
        return torch.zeros(0, 0)
    return F.cosine_similarity(proto[:, None, None, :], tokens_hwc[None], dim=-1)


def part_centers(maps, rois, obj_label, vit_feat, num_max_keep=50, num_max_obj=3):
    """RH:222-262 ``get_center_coord_with_feat``: per instance, parts ordered by
    area(sim > 0.9) descending, at most num_max_obj+1 of them; centre = mean of the
    arg-max patch coordinates (ties averaged), mapped to pixels as (c + 0.5) * 16; kept
    when inside the box.  Returns the reference's 8-tuple."""
    coords, labels, feats, corr = [], [], [], []
    split = [0 for _ in range(len(maps))]
    for i, m in enumerate(maps):
        if m.shape[0] == 0:
            continue
        top = m.flatten(1).topk(dim=1, k=1)[0][:, -1, None, None]
        peak = (m >= top).nonzero().float()
        x0, y0, x1, y1 = rois[i]
        order = (m > 0.9).sum(dim=[-2, -1]).argsort(descending=True, dim=0)
        for k in range(m.shape[0]):
            if k > num_max_obj:
                break
            cm = peak[peak[:, 0] == order[k]].mean(dim=0)[1:].flip(0)
            xy = (cm + 0.5) * PATCH
            if (xy[0] >= x0) & (xy[0] <= x1) & (xy[1] >= y0) & (xy[1] <= y1):
                coords.append(xy)
                labels.append(obj_label[i])
                corr.append(i)
                feats.append(vit_feat[:, cm[1].long(), cm[0].long()])
                split[i] += 1
    if len(coords) == 0:
        z2 = torch.zeros(0, 2, dtype=rois[0].dtype)
        zl = torch.zeros(0, dtype=obj_label[0].dtype)
        return [z2, zl], [], [], [], split, z2.clone(), zl.clone(), torch.zeros(0, dtype=torch.long)
    coords = torch.stack(coords)
    labels = torch.stack(labels)
    c_org, l_org = coords.clone(), labels.clone()
    feats = torch.stack(feats)
    c_split = list(coords.split(split, dim=0))
    f_split = list(feats.split(split, dim=0))
    if coords.shape[0] > num_max_keep:
        pick = torch.randperm(coords.shape[0])[:num_max_keep]
        coords, labels = coords[pick], labels[pick]
    return [coords, labels], c_split, f_split, feats, split, c_org, l_org, torch.tensor(corr, dtype=torch.long)


def semantic_centers(map_fg, map_bg, rois, vit_feat, pos_thr=0.35, refine_times=5, gt_labels=None,
                     merge_thr=0.85, num_semantic_points=3, n_points=20, trace=None):
    """RH:1995-2031 ``get_semantic_centers``.  map_fg/map_bg [n_obj,H,W] (last
    refinement step).  Returns the reference's 9-tuple.  ``n_points`` is 20 in the
    reference (hard-coded at RH:2024); exposed here so BASELINE's 4/16/32/64-seed
    configs stay self-consistent."""
    hp, wp = vit_feat.shape[-2:]
    hard = torch.where(map_fg > pos_thr, torch.ones_like(map_fg), torch.zeros_like(map_fg))
    fg_low = F.interpolate(erode(hard, 11).unsqueeze(0), (hp, wp), mode='bilinear')[0]
    bg_low = F.interpolate(map_bg.unsqueeze(0).max(dim=1, keepdim=True)[0], (hp, wp), mode='bilinear')[0]
    seeds_map = torch.where(fg_low > pos_thr, torch.ones_like(fg_low), torch.zeros_like(fg_low))
    prot, sim = mean_shift_from_maps(seeds_map, vit_feat, rois, tau=0.1, temp=0.1, n_shift=refine_times,
                                     n_points=n_points, trace=trace)
    keep = filter_seed_maps(sim.unflatten(0, (sim.shape[0] // n_points, n_points)), fg_low)
    split = keep.sum(dim=-1).tolist()
    merged = merge_prototypes(prot[keep.flatten()].split(split, dim=0), thr=merge_thr)
    sims = [part_maps(p, vit_feat.permute(1, 2, 0)) for p in merged]
    (cc, c_split, f_split, f_all, nparts, c_org, l_org, corr) = part_centers(
        sims, rois, gt_labels, vit_feat, num_max_obj=num_semantic_points)
    return cc, c_split, sims, f_split, f_all, nparts, c_org, l_org, corr


# --------------------------------------------------------------------------- A13
def pseudo_masks(map_fg_last, pos_mask_thr):
    """RH:2356-2358: uint8 [n_obj,H,W] = fg map > rowmax * thr."""
    mx = map_fg_last.flatten(1).max(1)[0][:, None, None]
    return (map_fg_last > mx * pos_mask_thr).to(torch.uint8)


# --------------------------------------------------------------------------- A14 (per image chain)
def attention_shift_image(cams_up, gt_index, pseudo_boxes, vit_feat, gt_points, gt_labels,
                          pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10, corr_size=21,
                          obj_tau=0.85, mean_shift_times=10, num_semantic_points=3, n_points=20, hook=None, img=0,
                          trace=None):
    """The per-image body of ``seed_pseudo_gt`` after the MIL selection (RH:2332-2361):
    refined fg/bg maps -> mask-head point labels -> semantic (part) centres -> pseudo
    instance masks.  Returns a dict with the keys of RH:2398-2415 that this path owns."""
    coords, labels, fg, bg, p_a, p_b, f_fg, f_bg = mask_sample_points(
        cams_up, pseudo_boxes, gt_index, vit_feat, pos_thr=pos_mask_thr, neg_thr=neg_mask_thr,
        num_gt=num_mask_point_gt, corr_size=corr_size, obj_tau=obj_tau, gt_points=gt_points, hook=hook, img=img)
    sc = semantic_centers(fg[-1].clone(), bg[-1].clone(), pseudo_boxes, vit_feat, pos_thr=pos_mask_thr,
                          refine_times=mean_shift_times, gt_labels=gt_labels,
                          num_semantic_points=num_semantic_points, n_points=n_points, trace=trace)
    return dict(mask_points_coords=coords, mask_points_labels=labels, map_cos_fg=fg[-1], map_cos_bg=bg[-1],
                semantic_centers=sc[0], semantic_centers_split=sc[1], sim_fg=sc[2],
                semantic_centers_feat_split=sc[3], semantic_centers_feat=sc[4], num_parts=sc[5],
                semantic_centers_org=(sc[6], sc[7]), corres_gts=sc[8],
                pseudo_gt_masks=pseudo_masks(fg[-1], pos_mask_thr), inst_fg_feat=f_fg, inst_bg_feat=f_bg,
                points_a=p_a, points_b=p_b)


# --------------------------------------------------------------------------- A15 (second-round aggregation)
def point_sample(feats, points, **kw):
    """mmcv.ops.point_sample (mmcv-full 1.3.8, ``mmcv/ops/point_sample.py``; third-party, absent from the reference tree --
    restated from its published source, PARITY UNPINNED for this one call): bilinear ``grid_sample`` at points given in
    [0,1] x [0,1] (x, y), ``align_corners=False``.  feats [N,C,H,W]; points [N,P,2] -> [N,C,P], or [N,Hg,Wg,2] -> [N,C,Hg,Wg]."""
    add_dim = points.dim() == 3
    if add_dim:
        points = points.unsqueeze(2)
    out = F.grid_sample(feats, points * 2.0 - 1.0, align_corners=False, **kw)
    return out.squeeze(3) if add_dim else out


def extract_bg_coords(bg_map, num_groups=3, num_points_per_group=5, hook=None, key=None):
    """RH:28-49: up to ``num_groups * num_points_per_group`` random non-zero pixels of ``bg_map`` [H,W] (``torch.randperm``
    on the CPU generator), repeated to the full count, as ``(index + 0.5) / size`` -- in (row, col) order and with the POINT
    order reversed (the reference's ``.flip(0)`` flips dim 0 of the [P,2] array).  -> [num_groups, P, 2]."""
    nz = torch.nonzero(bg_map)
    max_points = num_groups * num_points_per_group
    if nz.size(0) == 0:
        idx = torch.ones(max_points, 2, dtype=nz.dtype)
    else:
        k = min(max_points, nz.size(0))
        if hook is not None:
            hook(key)                       # test hook: re-seed the CPU generator per (image, stage, instance) key
        perm = torch.randperm(nz.size(0))
        idx = nz[perm[:k]]
        while idx.size(0) < max_points:
            idx = torch.cat((idx, nz[perm[:max_points - idx.size(0)]]))
    coords = idx.float() + 0.5
    coords = (coords / coords.new_tensor(bg_map.shape[:2])).flip(0)
    return coords.reshape(num_groups, -1, 2)


def refined_similarity_input_map(cos_map, feats, bboxes, refine_times=1, tau=0.85, is_select=False):
    """RH:710-748 ``get_refined_similarity_input_map``: the refinement loop of RH:668-707 started from given maps.
    cos_map [K,Hp,Wp] (modified in place like the reference), feats [1,C,Hp,Wp] -> (maps [refine_times+1,K,Hp,Wp], centroids)."""
    cur = cos_map.clone()
    n_obj = bboxes.shape[0]
    bmask = box_to_mask(bboxes // PATCH, cos_map.shape[-2:], default=0)
    outs = []

    def emit(m):
        if is_select:
            m[:n_obj] = m[:n_obj] * bmask
            win = m.argmax(0, keepdim=True).expand_as(m)
            rows = torch.arange(m.shape[0])
            outs.append(torch.where(win == rows[:, None, None], m.clone(), torch.zeros_like(m)))
        else:
            outs.append(m.clone())

    emit(cos_map)
    centroid = None
    for _ in range(refine_times):
        mx = cur.flatten(1).max(1, keepdim=True)[0].unsqueeze(-1)
        cur[cur < mx * tau] *= 0
        w = cur.unsqueeze(1)
        centroid = (feats * w).sum([2, 3], keepdim=True) / w.sum([2, 3], keepdim=True).clamp(1e-8)
        cur = F.cosine_similarity(feats, centroid, dim=1)
        emit(cur)
    return torch.stack(outs), centroid


def update_fg_map_single(map_cos_fg, feats, coords, num_parts, inst_fg_feats, inst_bg_feats, bboxes, img_size, hook=None,
                         key=None):
    """RH:2812-2844 ``update_fg_map_single_v3``.  map_cos_fg [n,H,W]; feats [1,C,Hp,Wp]; coords [P,2] part centres (x,y)
    pixels; num_parts: parts per instance; inst_fg_feats [1,n+1,C]; inst_bg_feats [1,n,C]; bboxes [n,4]; img_size (W,H).
    -> refined instance maps [n,H,W].  Kept quirks of the reference: the part features enter as the SCALAR mean of all
    their elements (``torch.mean`` without a dim), and the background supplement is sampled at (row, col) fed as (x, y)."""
    sc_feat = point_sample(feats, (coords / img_size)[None]).permute(0, 2, 1)
    split = sc_feat.split(num_parts, dim=1)
    n = bboxes.shape[0]
    inst = []
    for i in range(n):
        inst.append(torch.mean(split[i]) * 0.5 + inst_fg_feats[:, i] * 0.5 if split[i].shape[0] != 0 else inst_fg_feats[:, i])
    inst.append(inst_fg_feats[:, -1])
    inst = torch.cat(inst, dim=0)
    bg_map = map_cos_fg.sum(0) == 0
    bg_coords = extract_bg_coords(bg_map, num_groups=1, hook=hook, key=key)
    bg_supp = point_sample(feats, bg_coords[None], mode='bilinear').mean(-1).permute(0, 2, 1)[0]
    inst = torch.cat((inst, bg_supp), dim=0)
    fn = feats / torch.norm(feats, p=2, keepdim=True, dim=1)
    attn = torch.einsum('nchw, nmc -> nmhw', fn, inst[None] / torch.norm(inst[None], p=2, keepdim=True, dim=2))
    bg_attn = torch.einsum('nchw, nmc -> nmhw', fn, inst_bg_feats / torch.norm(inst_bg_feats, p=2, keepdim=True, dim=2))
    attn, _ = refined_similarity_input_map(attn[0], feats, bboxes, 3, is_select=True)
    size = map_cos_fg.shape[-2:]
    attn = F.interpolate(attn[-1, :n][None], size, mode='bilinear')[0]
    bg_attn = F.interpolate(bg_attn, size, mode='bilinear')[0]
    attn = (1 - bg_attn) * attn
    attn /= attn.flatten(-2, -1).max(-1, keepdim=True)[0].unsqueeze(-1).clamp(1e-8)
    return attn.detach().clone()


def update_fg_map(map_cos_fg, vit_feat, coords, num_parts, inst_fg_feat, inst_bg_feat, gt_bboxes, pos_mask_thr, hook=None):
    """RH:2737-2760 ``update_fg_map`` for a batch.  map_cos_fg: list of [n_i,H,W]; vit_feat [B,1+N,C] (cls token first);
    coords / num_parts / gt_bboxes: per image; inst_fg_feat [n_i+1,C,1,1], inst_bg_feat [n_i,C,1,1] (seed_pseudo_gt outputs).
    -> (list of refined maps, list of uint8 masks)."""
    img_size = vit_feat.new_tensor(map_cos_fg[0].shape[-2:][::-1])
    patch = (img_size.flip(0) // 16).long().tolist()
    maps, masks = [], []
    for i in range(vit_feat.shape[0]):
        feats = vit_feat[[i]][:, 1:].permute(0, 2, 1).unflatten(-1, patch)
        attn = update_fg_map_single(map_cos_fg[i], feats, coords[i], num_parts[i], inst_fg_feat[i][:, :, 0, 0][None],
                                    inst_bg_feat[i][:, :, 0, 0][None], gt_bboxes[i], img_size, hook=hook, key=(i, 2, 0))
        drop = attn.sum(dim=[1, 2]) == 0
        attn[drop] = map_cos_fg[i][drop]
        maps.append(attn)
        masks.append(torch.where(attn > attn.flatten(1).max(1)[0][:, None, None] * pos_mask_thr, torch.ones_like(attn),
                                 torch.zeros_like(attn)).to(torch.uint8))
    return maps, masks

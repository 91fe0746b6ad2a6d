"""Attention shift on the device: roll-out slab -> CAM boxes -> refined instance maps -> mask points ->
mean-shift part discovery -> pseudo masks.  Host orchestration of the C-ABI kernels for the body of
``seed_pseudo_gt`` (RH = mmdet/models/roi_heads/stdroi_point_deform_attn_reppoints.py, RH:2261-2361).

Instances of the whole batch are a flat list (``obj_img`` maps instance -> image), every stage is batched over it.
The reference synchronises with the host for every instance / component / seed map (``unique()``, ``.tolist()``,
``nonzero()``, python loops); here the host is consulted exactly three times per batch:
  (1) candidate counts for the random seed points   (the RNG is torch's CPU generator, like the reference),
  (2) candidate counts for the mask-head points,
  (3) the final part counts / validity flags needed to build the ragged python lists of the return dict.

RNG discipline: kernels take the drawn indices as inputs.  ``StreamRng`` replays the reference's
call sequence on one generator (exact stream parity; needs one image per call and a full ``randperm``);
``KeyedRng`` keys one generator per (image, stage, instance) so a batch needs no per-image ordering and
``randperm(n)[:k]`` is evaluated from its first k draws only.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import lib as _l
from . import ops

PATCH = 16
CAM_BBOX_SCRATCH_BYTES = 1 << 30     # transient connected-components scratch per call (see cam_bbox)
SEED_LEVELS = 4          # threshold-doubling rounds of the seed sampling answered by one counting pass


# --------------------------------------------------------------------------------------------------- RNG front ends
class StreamRng:
    """The reference's own call sequence on a single torch CPU generator (default: the global one)."""

    def __init__(self, generator=None):
        self.g = generator

    def randint(self, key, high, n):
        return torch.randint(high, (n,), generator=self.g)

    def randperm_head(self, key, n, k):
        return torch.randperm(n, generator=self.g)[:k]


_M64 = (1 << 64) - 1


def _mix64(x):
    """splitmix64 finaliser: a bijection of 64-bit integers with full avalanche."""
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


class KeyedRng:
    """One torch CPU generator per (rank, call, image, stage, instance) key.  Same algorithms as torch.randint /
    torch.randperm (mt19937, ``random() % range``, forward Fisher-Yates) so a reference run seeded with ``seed_for(key)``
    right before the corresponding call yields identical indices.  The raw mt19937 outputs of many keys come from one call
    into the C ABI (``as_mt19937_draws``), which is what lets the batched host code draw for every instance at once.

    ``step`` advances once per ``seed_pseudo_gt`` / ``update_fg_map`` call (``next_step``, called by the head) and ``rank``
    defaults to the process's RANK, so -- like the reference's advancing global generator -- no two training iterations and
    no two DDP ranks draw the same seed points / mask points.  The key is hashed through a 64-bit mixer: slots cannot collide
    the way a linear ``stage * 101 + obj`` formula does once an image has more than 100 instances."""

    def __init__(self, base_seed=0, rank=None):
        self.base = int(base_seed)
        self.rank = int(os.environ.get('RANK', 0) or 0) if rank is None else int(rank)
        self.step = 0

    def next_step(self):
        self.step += 1
        return self.step

    def seed_for(self, key):
        img, stage, obj = key
        h = _mix64(self.base & _M64)
        for v in (self.rank, self.step, img, stage, obj):
            h = _mix64(h ^ (int(v) & _M64))
        return h & 0x7fffffff

    def draws(self, keys, k):
        """Raw 32-bit generator outputs: uint32 array [len(keys), k] (k <= 624)."""
        L = _l.load()
        seeds = np.array([self.seed_for(key) for key in keys], dtype=np.uint32)
        out = np.empty((len(keys), k), dtype=np.uint32)
        if len(keys) and k:
            rc = L.as_mt19937_draws(seeds.ctypes.data_as(ctypes.c_void_p), len(keys), k, out.ctypes.data_as(ctypes.c_void_p))
            if rc != 0:
                raise ValueError('as_mt19937_draws: bad argument')
        return out

    def randint(self, key, high, n):
        if n > 624:
            return torch.randint(high, (n,), generator=torch.Generator().manual_seed(self.seed_for(key)))
        return torch.from_numpy((self.draws([key], n)[0].astype(np.int64) % int(high)))

    @staticmethod
    def perm_head(raw, n, k):
        """First k entries of torch.randperm(n) on the CPU (forward Fisher-Yates, TensorFactories.cpp): they depend only on
        the first k draws z_i = random() % (n - i).  raw: the generator's raw outputs (>= k of them)."""
        perm, out = {}, []
        for i in range(min(k, n)):
            if i < n - 1:
                j = i + int(raw[i]) % (n - i)
                vi, vj = perm.get(i, i), perm.get(j, j)
                perm[i], perm[j] = vj, vi
                out.append(vj)
            else:
                out.append(perm.get(i, i))
        return out

    def randperm_head(self, key, n, k):
        return torch.tensor(self.perm_head(self.draws([key], min(k, 624))[0], n, k), dtype=torch.int64)


def _fill_index(idx, n):
    """RH:1147-1155 ``fill_in_idx`` on a 1-D index tensor."""
    assert idx.shape[0] != 0
    if idx.shape[0] >= n / 2:
        return torch.cat((idx, idx[:n - idx.shape[0]]), dim=0)
    return _fill_index(idx.repeat(n // idx.shape[0]), n)


def _i32(x, dev):
    return torch.as_tensor(x, dtype=torch.int32).to(dev, non_blocking=True)


def _f32(x, dev):
    return torch.as_tensor(x, dtype=torch.float32).to(dev, non_blocking=True)


class _PinnedRing:
    """A fixed arena of pinned host memory handed out as a ring: the small host<->device transfers of a step (index
    uploads, candidate counts) never call cudaHostAlloc again after start-up -- a pinned allocation costs milliseconds and
    showed up as occasional 50 ms steps when torch's caching host allocator had to grow.  A region is reused only after
    the event recorded behind its transfer has completed (normally hundreds of steps earlier)."""

    def __init__(self, nbytes=8 << 20):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        self.size = nbytes
        self.head = 0
        self.inflight = []          # (start, end, event), in allocation order

    def take(self, nbytes):
        """-> (uint8 view of ``nbytes`` pinned bytes, token for ``done``)."""
        nbytes = max((int(nbytes) + 255) & ~255, 256)
        if nbytes > self.size:
            return self._fresh(nbytes), None
        for _ in range(4):
            if self.head + nbytes > self.size:
                self.head = 0
            start, end = self.head, self.head + nbytes
            busy = [e for e in self.inflight if e[0] < end and e[1] > start and e[2] is None]
            if busy:                                      # still being filled / read by the host: step over it
                self.head = max(e[1] for e in busy)
                continue
            keep = []
            for ent in self.inflight:
                if ent[0] < end and ent[1] > start:       # the ring has come round: its transfer must have completed
                    ent[2].synchronize()
                else:
                    keep.append(ent)
            self.inflight = keep
            self.head = end
            token = [start, end, None]
            self.inflight.append(token)
            return self.buf[start:end], token
        return self._fresh(nbytes), None

    @staticmethod
    def _fresh(nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)

    def done(self, token):
        """The host is finished with the region and the transfer that uses it has been enqueued on the current stream."""
        if token is not None:
            token[2] = torch.cuda.Event()
            token[2].record()


_RINGS = {}


def _ring(dev):
    key = str(dev)
    r = _RINGS.get(key)
    if r is None:
        r = _RINGS[key] = _PinnedRing()
    return r


def _upload_i32(arrays, dev):
    """Several small host integer arrays -> the device in ONE pinned, asynchronous copy.  Returns int32 device views (in
    order, shaped like the inputs).  Pageable uploads stall the host for a driver round trip each; this path has dozens."""
    arrays = [np.ascontiguousarray(a, dtype=np.int32) for a in arrays]
    sizes = [a.size for a in arrays]
    total = max(sum(sizes), 1)
    ring = _ring(dev)
    raw, token = ring.take(total * 4)
    host = raw[:total * 4].view(torch.int32)
    hv = host.numpy()
    o = 0
    for a, n in zip(arrays, sizes):
        hv[o:o + n] = a.reshape(-1)
        o += n
    d = host.to(dev, non_blocking=True)
    ring.done(token)
    out, o = [], 0
    for a, n in zip(arrays, sizes):
        out.append(d[o:o + n].view(a.shape))
        o += n
    return out


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 1), device=dev, dtype=torch.uint8)


def _sp():
    return _l.stream_ptr()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _Pending:
    """Device -> pinned-host copy enqueued now, awaited later: the host keeps launching kernels while the counts travel."""

    def __init__(self, t):
        ring = _ring(t.device)
        raw, token = ring.take(t.numel() * t.element_size())
        self.host = raw[:t.numel() * t.element_size()].view(t.dtype).view(t.shape)
        self.host.copy_(t, non_blocking=True)
        self.ring, self.token = ring, token
        self.ev = torch.cuda.Event()
        self.ev.record()

    def get(self):
        """Host tensor (a view of the pinned ring: read it before the ring comes round, i.e. right away)."""
        self.ev.synchronize()
        if self.token is not None:
            self.ring.done(self.token)
            self.token = None
        return self.host

    def __del__(self):
        # never read: hand the ring region back (behind an event, the copy may still be in flight) instead of leaving it
        # marked busy for good
        try:
            if self.token is not None:
                self.ring.done(self.token)
                self.token = None
        except Exception:
            pass


_OBJ_IMG = {}


def instance_image_index(n_per_img, dev):
    """obj_img [n_tot] int32 on the device (instance -> image), cached per batch composition (read-only)."""
    key = (tuple(n_per_img), str(dev))
    t = _OBJ_IMG.get(key)
    if t is None:
        if len(_OBJ_IMG) > 64:
            _OBJ_IMG.clear()
        t = torch.repeat_interleave(torch.arange(len(n_per_img), dtype=torch.int32), torch.tensor(list(n_per_img))).to(dev)
        _OBJ_IMG[key] = t
    return t


def token_major(vit_feat):
    """[B,C,Hp,Wp] (the reference hands over a permuted view of last_feat, DET:77) or [B,N,C] -> [B,N,C] fp32 with
    unit channel stride, without a copy when the memory already is token-major."""
    if vit_feat.dim() == 4:
        b, c, hp, wp = vit_feat.shape
        t = vit_feat.permute(0, 2, 3, 1).reshape(b, hp * wp, c)
    else:
        t = vit_feat
    if t.dtype != torch.float32 or t.stride(2) != 1 or t.stride(1) != t.shape[2]:
        t = t.float().contiguous()
    return t


# --------------------------------------------------------------------------------------------------- A5: roll-out
def rollout_rows(attns, n_rows, use_tensor_cores=True):
    """RH:1257-1272 restricted to the last ``n_rows`` rows.  attns: list of L tensors [B,T,T] (row stride may be padded).
    -> [B, L, n_rows, T] view (index 0 = last layer alone).  When the maps carry the split-fp16 transposed copies written
    by the head-mean kernel (``_as_t16``) the products run on the tensor cores; otherwise on the fp32 CUDA-core slab GEMM."""
    L = _l.load()
    nl = len(attns)
    # maps produced for the roll-out only (backbone attn_format='rollout') are empty placeholders that carry the operands as
    # attributes; the LAST layer's map is always a real tensor (its last n_rows rows are read)
    B, T, _ = attns[-1].shape
    ld = attns[-1].stride(1)
    dev = attns[-1].device
    if getattr(attns[-1], '_as_valid_from', 0) > T - n_rows:
        raise ValueError('the last layer\'s head-mean map does not hold the rows the roll-out reads')
    parts = []
    for a in attns:
        p = getattr(a, '_as_rowsum_part', None)
        if a.numel() == 0:
            assert p is not None and getattr(a, '_as_t16', None) is not None and a._as_shape[:2] == (B, T)
        else:
            assert a.dtype == torch.float32 and a.stride(2) == 1 and a.stride(1) == ld and a.stride(0) == T * ld
        parts.append(p if p is not None else a.sum(-1, keepdim=True).contiguous())
    ntile = parts[0].shape[2]
    a_ptrs = (ctypes.c_void_p * nl)(*[(a if a.numel() else attns[-1]).data_ptr() for a in attns])
    p_ptrs = (ctypes.c_void_p * nl)(*[p.data_ptr() for p in parts])
    t16 = [getattr(a, '_as_t16', None) for a in attns]
    if use_tensor_cores and all(t is not None for t in t16[:-1]) and n_rows <= 128:
        ldt = (T + 127) // 128 * 128
        dummy = t16[0][0] if t16[0] is not None else attns[0]
        h_ptrs = (ctypes.c_void_p * nl)(*[(t[0] if t is not None else dummy).data_ptr() for t in t16])
        l_ptrs = (ctypes.c_void_p * nl)(*[(t[1] if t is not None else dummy).data_ptr() for t in t16])
        out = torch.empty(B, nl, n_rows, ldt, device=dev, dtype=torch.float32)
        nbytes = L.as_rollout_tc_workspace(B, T, ldt)
        ws = _ws(nbytes, dev)
        _l.check(L.as_rollout_rows_tc(a_ptrs, h_ptrs, l_ptrs, p_ptrs, nl, B, T, ld, ldt, ops.T_SCALE, ntile, n_rows, _p(out),
                                      _p(ws), nbytes, _sp()), 'as_rollout_rows_tc')
        return out[..., :T]
    if any(a.numel() == 0 for a in attns):
        raise ValueError('roll-out-only head-mean maps need the tensor-core roll-out (n_rows <= 128)')
    out = torch.empty(B, nl, n_rows, T, device=dev, dtype=torch.float32)
    nbytes = L.as_rollout_workspace(B, T, n_rows)
    ws = _ws(nbytes, dev)
    _l.check(L.as_rollout_rows(a_ptrs, p_ptrs, nl, B, T, ld, ntile, n_rows, _p(out), _p(ws), nbytes, _sp()), 'as_rollout_rows')
    return out


# --------------------------------------------------------------------------------------------------- A6 / A7
def cam_maps(rows, obj_img, obj_pt, hp, wp):
    """RH:2272-2275 without the up-sampling: slice the matched point-token rows -> low-res CAMs [L,n_tot,N] and the
    min / max [L,n_tot,2] their x16 bilinear up-sampling would have."""
    L = _l.load()
    B, nl, n_rows, T = rows.shape
    n_tot = obj_img.shape[0]
    N = hp * wp
    dev = rows.device
    ldr = rows.stride(2)
    assert rows.stride(3) == 1 and rows.stride(1) == n_rows * ldr and rows.stride(0) == nl * n_rows * ldr
    cams = torch.empty(nl, n_tot, N, device=dev, dtype=torch.float32)
    _l.check(L.as_cam_gather(_p(rows), _p(obj_img), _p(obj_pt), nl, n_rows, ldr, N, n_tot, _p(cams), _sp()), 'as_cam_gather')
    n_maps = nl * n_tot
    mm = torch.empty(n_maps, 2, device=dev, dtype=torch.float32)
    scratch = torch.empty(n_maps * 2, device=dev, dtype=torch.int32)
    _l.check(L.as_cam_minmax(_p(cams), n_maps, hp, wp, _p(mm), _p(scratch), _sp()), 'as_cam_minmax')
    return cams, mm.view(nl, n_tot, 2)


def cam_bbox(cams, mm, gt_points, hp, wp, cam_thr=0.2, area_ratio=0.5, want_keep_mask=False):
    """RH:60-116 get_bbox_from_cam_fast for every (layer, gt) map.  -> (boxes [L,n_tot,4], keep_mask or None)."""
    L = _l.load()
    nl, n_tot, N = cams.shape
    dev = cams.device
    n_maps = nl * n_tot
    boxes = torch.empty(n_maps, 4, device=dev, dtype=torch.float32)
    keep = torch.empty(n_maps, hp * 16, wp * 16, device=dev, dtype=torch.uint8) if want_keep_mask else None
    # scratch is sized for the worst case of H * ceil(W / 2) runs per map (7.3 MB per 1024^2 map): bound it by walking the maps in
    # groups of whole layers (the kernels are per-map, the [layer][instance] order makes every group a contiguous slice), so a
    # crowded batch costs a few more launches instead of gigabytes of transient memory
    per_layer = L.as_cam_bbox_workspace(n_tot, hp * 16, wp * 16)
    step = max(1, min(nl, CAM_BBOX_SCRATCH_BYTES // max(per_layer, 1)))
    nbytes = L.as_cam_bbox_workspace(step * n_tot, hp * 16, wp * 16)
    ws = _ws(nbytes, dev)
    mm2 = mm.reshape(n_maps, 2)
    for l0 in range(0, nl, step):
        l1 = min(nl, l0 + step)
        m0, m1 = l0 * n_tot, l1 * n_tot
        _l.check(L.as_cam_bbox(_p(cams[l0:l1]), _p(mm2[m0:m1]), _p(gt_points), m1 - m0, n_tot, hp, wp, float(cam_thr),
                               float(area_ratio), _p(boxes[m0:m1]), _p(keep[m0:m1]) if keep is not None else None, _p(ws), nbytes,
                               _sp()), 'as_cam_bbox')
    return boxes.view(nl, n_tot, 4), keep


def cam_boxes(rows, obj_img, obj_pt, gt_points, hp, wp, cam_thr=0.2, area_ratio=0.5, want_keep_mask=False):
    """RH:2272-2290: cam_maps + cam_bbox.  -> (cams [L,n_tot,N], minmax [L,n_tot,2], boxes [L,n_tot,4], keep_mask or None)."""
    cams, mm = cam_maps(rows, obj_img, obj_pt, hp, wp)
    boxes, keep = cam_bbox(cams, mm, gt_points, hp, wp, cam_thr, area_ratio, want_keep_mask)
    return cams, mm, boxes, keep


def cosine_maps(feats, grp_img, protos, clamp0=False):
    """sim[g,s,n] = cos(protos[g,s], feats[grp_img[g],n]).  feats [n_img,N,C], protos [G,S,C] -> [G,S,N]."""
    L = _l.load()
    n_img, N, C = feats.shape
    G, S, _ = protos.shape
    sim = torch.empty(G, S, N, device=feats.device, dtype=torch.float32)
    nbytes = L.as_cosine_maps_workspace(n_img, G, S, N, C)
    ws = _ws(nbytes, feats.device)
    _l.check(L.as_cosine_maps(_p(feats), feats.stride(0), n_img, N, C, _p(grp_img), _p(protos), G, S, _p(sim), int(clamp0),
                              _p(ws), nbytes, _sp()), 'as_cosine_maps')
    return sim


# --------------------------------------------------------------------------------------------------- A8
class _Groups:
    """Row layout of the refinement state: per image g, rows [0,n) fg instances, row n the image-level bg supplement,
    rows (n, 2n] bg instances; padded to S = max(2n+1)."""

    def __init__(self, n_per_img, dev):
        self.n = list(n_per_img)
        self.first = [0]
        for k in self.n[:-1]:
            self.first.append(self.first[-1] + k)
        self.S = max(2 * k + 1 for k in self.n)
        self.G = len(self.n)
        self.d_first, self.d_n = _upload_i32([self.first, self.n], dev)
        self.d_img = torch.arange(self.G, dtype=torch.int32, device=dev)
        self.d_row_img = self.d_img.repeat_interleave(self.S)


_PLANS = {}


def _seed_plan(n_per_img, dev, thr_pos, thr_neg):
    """Everything about the seed-sampling items that only depends on the instance counts per image: built (and uploaded)
    once, reused by every later batch of the same composition.  Read-only afterwards."""
    key = (tuple(n_per_img), str(dev), float(thr_pos), float(thr_neg))
    plan = _PLANS.get(key)
    if plan is not None:
        return plan
    grp = _Groups(n_per_img, dev)
    # items ordered per image as the reference draws them: bg instances, fg instances, supplement
    kinds, ia, ib, thr, item_img, item_slot = [], [], [], [], [], []
    for g, n in enumerate(grp.n):
        o0 = grp.first[g]
        for j in range(n):
            kinds.append(0); ia.append(o0 + j); ib.append(0); thr.append(thr_neg); item_img.append(g); item_slot.append(n + 1 + j)
        for j in range(n):
            kinds.append(1); ia.append(o0 + j); ib.append(0); thr.append(thr_pos); item_img.append(g); item_slot.append(j)
        kinds.append(2); ia.append(o0); ib.append(o0 + n); thr.append(thr_neg); item_img.append(g); item_slot.append(n)
    d_kind, d_a, d_b = _upload_i32([kinds, ia, ib], dev)
    plan = dict(grp=grp, kinds=kinds, ia=ia, ib=ib, thr=thr, item_img=item_img, item_slot=item_slot, n_items=len(kinds),
                d_kind=d_kind, d_a=d_a, d_b=d_b, d_thr=_f32(thr, dev), np_kind=np.array(kinds), np_a=np.array(ia),
                np_row=np.array(item_img) * grp.S + np.array(item_slot),
                keys=[(g, 0, sl) for g, sl in zip(item_img, item_slot)])
    if len(_PLANS) > 64:
        _PLANS.clear()
    _PLANS[key] = plan
    return plan


def refined_maps_begin(cam_low, cam_mm, n_per_img, hp, wp, thr_pos=0.2, thr_neg=0.1):
    """Candidate counting for the seed sampling (RH:343-352) + asynchronous copy of the counts to the host.  Issue this as
    early as the selected CAMs exist: the counts cross PCIe while the GPU runs the connected-components stage."""
    L = _l.load()
    dev = cam_low.device
    H = hp * PATCH
    st = dict(_seed_plan(n_per_img, dev, thr_pos, thr_neg))
    st.update(cam_low=cam_low, cam_mm=cam_mm, hp=hp, wp=wp)
    n_items = st['n_items']

    d_kind, d_a, d_b = st['d_kind'], st['d_a'], st['d_b']

    def count(d_thr, levels=1):
        # NB: closes over the tensors it needs, NOT over ``st`` -- ``st['count'] = count`` would otherwise make a reference cycle
        # that keeps ~1 MB of device tensors per call alive until the cyclic GC runs; the caching allocator then keeps growing
        # its pool and every growth is a cudaMalloc (seen as one 40-100 ms step in ten, profiles/step_outliers_r2.txt)
        rc = torch.empty(levels, n_items, H, device=dev, dtype=torch.int32)
        _l.check(L.as_norm_rowcount(_p(cam_low), _p(cam_mm), _p(d_kind), _p(d_a), _p(d_b), _p(d_thr), n_items,
                                    hp, wp, levels, _p(rc), _sp()), 'as_norm_rowcount')
        return rc, d_thr, _Pending(rc.sum(2, dtype=torch.int32))

    st['count'] = count
    # one pass counts for thr, 2 thr, 4 thr, 8 thr: the first rounds of the reference's threshold-doubling loop need no
    # second kernel (and no second trip to the host)
    st['rowcnt'], st['d_thr'], st['pending'] = count(st['d_thr'], SEED_LEVELS)
    return st


def select_seed_candidates(kind, inst, row, totals, ks, P, n_rows, gt_points_fn):
    """Host half of RH:343-371 once the candidate counts are known (pure numpy, CPU-tested against the oracle).
    kind [n_items] (1 = foreground item), inst [n_items] instance whose GT point pads a short foreground item, row [n_items]
    destination row of the [n_rows, P, 2] point table, totals [n_items] candidate counts, ks [n_items, P] drawn candidate
    ranks (``randint(total)[:P]``).  -> (pts_host [n_rows,P,2] int32: the GT fill, zeros elsewhere; sel_item, sel_k: which
    candidate (row-major rank) of which item the device has to look up; sel_dst: its flat slot in the point table).
    A foreground item with fewer than P candidates takes all of them and repeats the GT point (RH:354-358) -- in (y, x)
    order: the reference appends the (x, y) point to (row, col) candidates and flips the last dimension of everything."""
    n_items = len(kind)
    short = (kind == 1) & (totals < P)
    slot_j = np.arange(P)[None, :]
    ks = np.where(short[:, None], slot_j, ks)
    use = ~short[:, None] | (slot_j < totals[:, None])       # which of the P slots are real candidates
    pts_host = np.zeros((n_rows, P, 2), dtype=np.int32)
    if short.any():
        gtp = gt_points_fn().astype(np.int32)                # int(float) truncation, as the reference's .long()
        for i in np.nonzero(short)[0]:
            pts_host[row[i], int(totals[i]):] = gtp[inst[i]][::-1]
    sel_item = np.broadcast_to(np.arange(n_items)[:, None], (n_items, P))[use]
    sel_k = ks[use]
    sel_dst = (row[:, None] * P + slot_j)[use]
    return pts_host, sel_item, sel_k, sel_dst


def refined_maps(cam_low, cam_mm, feats, n_per_img, rois, gt_points, hp, wp, rng, thr_pos=0.2, thr_neg=0.1,
                 num_points=20, refine_times=2, obj_tau=0.85, mask_thr=0.6, want_bg=True, want_mask=True, begun=None):
    """RH:1000-1019 for every instance of the batch.
    cam_low [n_tot,N] / cam_mm [n_tot,2]: selected-layer CAM (low-res) and min/max of its up-sampling; feats [n_img,N,C];
    rois [n_tot,4]; gt_points [n_tot,2].  -> dict(map_fg, map_bg [n_tot,H,W], mask uint8, fg_low, bg_low [n_tot,N],
    fg_feat [G,S,C] (rows 0..n of each group), pts [G,S,P,2], groups).  ``begun``: state of refined_maps_begin."""
    L = _l.load()
    dev = feats.device
    n_img, N, C = feats.shape
    n_tot = cam_low.shape[0]
    H, W = hp * PATCH, wp * PATCH
    st = begun if begun is not None else refined_maps_begin(cam_low, cam_mm, n_per_img, hp, wp, thr_pos, thr_neg)
    grp, kinds, ia, thr, item_img, item_slot, n_items = (st['grp'], st['kinds'], st['ia'], st['thr'], st['item_img'],
                                                          st['item_slot'], st['n_items'])
    d_kind, d_a, d_b = st['d_kind'], st['d_a'], st['d_b']
    P = num_points
    rowcnt, d_thr = st['rowcnt'], st['d_thr']
    lv_tot = st['pending'].get().numpy().astype(np.int64)      # host sync (1) -- normally already satisfied; [levels, n_items]
    np_kind, np_a, np_row = st['np_kind'], st['np_a'], st['np_row']
    # bg candidates too few -> the reference doubles the threshold until there are enough (RH:360-364): take the first
    # counted level that has enough; only if even the last one falls short do the doubling rounds continue on the device
    n_lv = lv_tot.shape[0]
    enough = (lv_tot >= P) | (np_kind == 1)[None, :]
    level = np.where(enough.any(0), enough.argmax(0), n_lv - 1)
    totals = lv_tot[level, np.arange(n_items)]
    factor = np.exp2(level).astype(np.float64)
    if n_lv == 1 or not (level > 0).any():
        rowcnt = rowcnt[0]
    else:
        d_level, = _upload_i32([level], dev)
        rowcnt = rowcnt[d_level.long(), torch.arange(n_items, device=dev)]       # [n_items, H] rows of the chosen levels
        d_thr = _f32((np.array(thr) * factor).tolist(), dev)
    while ((np_kind != 1) & (totals < P)).any():
        factor[(np_kind != 1) & (totals < P)] *= 2
        rowcnt, d_thr, pend = st['count'](_f32((np.array(thr) * factor).tolist(), dev))
        rowcnt = rowcnt[0]
        totals = pend.get().numpy().astype(np.int64)[0]
    short = (np_kind == 1) & (totals < P)                    # RH:354-358: all candidates, then the GT point repeated
    if hasattr(rng, 'draws'):
        ks = rng.draws(st['keys'], P).astype(np.int64) % np.maximum(totals, 1)[:, None]       # randint(num)[:P]
    else:                                                    # a single reference-ordered stream: item by item
        ks = np.zeros((n_items, P), dtype=np.int64)
        for i in np.nonzero(~short)[0]:
            num = int(totals[i])
            n_draw = len(range(0, num, num // P))
            ks[i] = (rng.randint(st['keys'][i], num, n_draw) % num)[:P].numpy()
    pts_host, sel_item, sel_k, sel_dst = select_seed_candidates(np_kind, np_a, np_row, totals, ks, P, grp.G * grp.S,
                                                                lambda: gt_points.detach().cpu().numpy())
    n_sel = int(sel_item.size)
    pts, d_item, d_k, d_dst = _upload_i32([pts_host, sel_item, sel_k, sel_dst], dev)
    pts = pts.view(grp.G, grp.S, P, 2)
    if n_sel:
        xy = torch.empty(n_sel, 2, device=dev, dtype=torch.int32)
        _l.check(L.as_norm_select(_p(cam_low), _p(cam_mm), _p(d_kind), _p(d_a), _p(d_b), _p(d_thr), hp, wp, _p(rowcnt),
                                  _p(d_item), _p(d_k), n_sel, _p(xy), _sp()), 'as_norm_select')
        pts.view(-1, 2)[d_dst.long()] = xy
    # ---- prototypes and refinement loop
    GS = grp.G * grp.S
    row_img = grp.d_row_img
    proto = torch.empty(grp.G, grp.S, C, device=dev, dtype=torch.float32)
    _l.check(L.as_seed_proto(_p(feats), feats.stride(0), _p(row_img), _p(pts), GS, P, C, hp, wp, _p(proto), _sp()), 'as_seed_proto')
    cur = cosine_maps(feats, grp.d_img, proto)
    fg_low = torch.empty(n_tot, N, device=dev, dtype=torch.float32)
    bg_low = torch.empty(n_tot, N, device=dev, dtype=torch.float32)
    wsum = torch.empty(GS, device=dev, dtype=torch.float32)
    nb = L.as_weighted_centroid_workspace(grp.G, grp.S, N, C)
    ws = _ws(nb, dev)
    centroid = proto
    assert refine_times >= 1
    for r in range(refine_times):
        _l.check(L.as_refine_threshold(_p(cur), GS, N, float(obj_tau), _p(wsum), _sp()), 'as_refine_threshold')
        centroid = torch.empty_like(proto)
        _l.check(L.as_weighted_centroid(_p(feats), feats.stride(0), _p(grp.d_img), _p(cur), _p(wsum), grp.G, grp.S, N, C,
                                        _p(centroid), _p(ws), nb, _sp()), 'as_weighted_centroid')
        cur = cosine_maps(feats, grp.d_img, centroid)
        _l.check(L.as_refine_select(_p(cur), grp.G, grp.S, N, wp, _p(grp.d_first), _p(grp.d_n), _p(rois),
                                    int(r == refine_times - 1), 1, _p(fg_low), _p(bg_low), _sp()), 'as_refine_select')
    # ---- full resolution
    map_fg = torch.empty(n_tot, H, W, device=dev, dtype=torch.float32)
    map_bg = torch.empty(n_tot, H, W, device=dev, dtype=torch.float32) if want_bg else None
    mask = torch.empty(n_tot, H, W, device=dev, dtype=torch.uint8) if want_mask else None
    stats = torch.empty(n_tot * 3, device=dev, dtype=torch.int32)
    _l.check(L.as_fuse_instance_maps(_p(fg_low), _p(bg_low), n_tot, hp, wp, float(mask_thr), _p(map_fg), _p(map_bg), _p(mask),
                                     _p(stats), _sp()), 'as_fuse_instance_maps')
    return dict(map_fg=map_fg, map_bg=map_bg, mask=mask, fg_low=fg_low, bg_low=bg_low, centroid=centroid, pts=pts, groups=grp)


# --------------------------------------------------------------------------------------------------- A12
def mask_points_begin(map_fg, map_bg, rois, pos_thr=0.6, neg_thr=0.6, corr_size=21):
    """Candidate maps + per-row counts of RH:433-443 and the asynchronous copy of the totals (and the boxes) to the host."""
    L = _l.load()
    dev = map_fg.device
    n_tot, H, W = map_fg.shape
    pos = torch.empty(n_tot, H, W, device=dev, dtype=torch.uint8)
    rowcnt = torch.empty(n_tot, H, 2, device=dev, dtype=torch.int32)
    nbytes = L.as_mask_candidates_workspace(n_tot, H, W)
    ws = _ws(nbytes, dev)
    _l.check(L.as_mask_candidates(_p(map_fg), _p(map_bg), _p(rois), n_tot, H, W, float(pos_thr), float(neg_thr), int(corr_size),
                                  _p(pos), _p(rowcnt), _p(ws), nbytes, _sp()), 'as_mask_candidates')
    return dict(pos=pos, rowcnt=rowcnt, ws=ws, map_bg=map_bg, rois=rois, neg_thr=neg_thr, pending=_Pending(rowcnt.sum(1, dtype=torch.int32)),
                pending_rois=_Pending(rois.detach().int()), shape=(n_tot, H, W))


def mask_points(map_fg, map_bg, rois, n_per_img, rng, pos_thr=0.6, neg_thr=0.6, num_gt=20, corr_size=21, begun=None):
    """RH:1980-1990 + RH:433-461.  -> (coords [n_tot,num_gt,2] fp32 (x,y), labels [n_tot,num_gt] bool)."""
    L = _l.load()
    dev = map_fg.device
    st = begun if begun is not None else mask_points_begin(map_fg, map_bg, rois, pos_thr, neg_thr, corr_size)
    n_tot, H, W = st['shape']
    pos, rowcnt, ws = st['pos'], st['rowcnt'], st['ws']
    totals = st['pending'].get().numpy().astype(np.int64)    # host sync (2): [n_tot, 2] = (#pos, #neg) candidates
    keys = [(g, 1, j) for g, n in enumerate(n_per_img) for j in range(n)]
    raw = rng.draws(keys, min(num_gt, 624)) if hasattr(rng, 'draws') else None
    sel_obj, sel_kind, sel_k, sel_dst = [], [], [], []
    coords = np.zeros((n_tot, num_gt, 2), dtype=np.float32)
    labels = np.zeros((n_tot, num_gt), dtype=np.bool_)
    rois_i = st['pending_rois'].get()                         # enqueued right behind the counts: on the host by now
    for o, key in enumerate(keys):
        n_pos, n_neg = int(totals[o, 0]), int(totals[o, 1])
        tot = n_pos + n_neg
        if raw is not None and num_gt <= 624:
            chosen = KeyedRng.perm_head(raw[o], tot, num_gt)
        else:
            chosen = rng.randperm_head(key, tot, num_gt).tolist()
        if len(chosen) < num_gt:
            if len(chosen) == 0:                             # RH:452-455 sentinel (-1,-1), then the crop offset is added
                coords[o, :, 0] = -1.0 + float(rois_i[o, 0])
                coords[o, :, 1] = -1.0 + float(rois_i[o, 1])
                continue
            chosen = _fill_index(torch.tensor(chosen), num_gt).tolist()
        for t, k in enumerate(chosen):
            is_pos = k < n_pos
            sel_obj.append(o); sel_kind.append(0 if is_pos else 1); sel_k.append(k if is_pos else k - n_pos)
            sel_dst.append(o * num_gt + t)
            labels[o, t] = is_pos
    n_sel = len(sel_obj)
    d_coords, d_labels, d_obj, d_kind, d_k, d_dst = _upload_i32([coords.view(np.int32), labels.astype(np.int32), sel_obj, sel_kind,
                                                                 sel_k, sel_dst], dev)
    coords = d_coords.view(torch.float32)                    # bit pattern of the fp32 sentinels travels inside the int32 pack
    labels = d_labels.bool()
    if n_sel:
        xy = torch.empty(n_sel, 2, device=dev, dtype=torch.int32)
        _l.check(L.as_mask_select(_p(pos), _p(map_bg), _p(rois), _p(ws), float(neg_thr), _p(rowcnt), _p(d_obj),
                                  _p(d_kind), _p(d_k), n_sel, H, W, _p(xy), _sp()), 'as_mask_select')
        coords.view(-1, 2)[d_dst.long()] = xy.float()
    return coords, labels


# --------------------------------------------------------------------------------------------------- A9 - A11
def semantic_parts(map_fg, feats, obj_img, rois, hp, wp, pos_thr=0.6, n_shift=10, n_points=20, merge_thr=0.85,
                   num_semantic_points=3, want_trace=False, n_per_img=None):
    """RH:1995-2031 up to (not including) the ragged list assembly.  Everything stays on the device.
    -> dict(fg_low, seed_map, seed_tok, prot, sim, keep, merged, n_merged, part_maps, centers, valid, part_id, cfeat, trace)."""
    L = _l.load()
    dev = map_fg.device
    n_tot, H, W = map_fg.shape
    N = hp * wp
    C = feats.shape[2]
    S = n_points
    fg_low = torch.empty(n_tot, N, device=dev, dtype=torch.float32)
    seed_map = torch.empty(n_tot, N, device=dev, dtype=torch.float32)
    _l.check(L.as_erode_downsample(_p(map_fg), n_tot, H, W, float(pos_thr), 11, _p(fg_low), _p(seed_map), _sp()), 'as_erode_downsample')
    seed_tok, proto0 = ops.grid_seeds(seed_map, feats, obj_img, rois, wp, S, thr=0.35)
    prot, sim, trace = ops.mean_shift(proto0, feats, obj_img, rois, hp, wp, n_shift, tau=0.1, temp=0.1, clamp0=True,
                                      want_trace=want_trace, n_per_img=n_per_img)
    keep = torch.empty(n_tot, S, device=dev, dtype=torch.int32)
    _l.check(L.as_filter_seeds(_p(sim), _p(fg_low), n_tot, S, N, 0.85, _p(keep), None, _sp()), 'as_filter_seeds')
    merged = torch.empty(n_tot, S, C, device=dev, dtype=torch.float32)
    n_merged = torch.empty(n_tot, device=dev, dtype=torch.int32)
    _l.check(L.as_merge_prototypes(_p(prot), _p(keep), n_tot, S, C, float(merge_thr), _p(merged), _p(n_merged), _sp()), 'as_merge_prototypes')
    pmaps = cosine_maps(feats, obj_img, merged)
    KP = num_semantic_points + 1
    centers = torch.empty(n_tot, KP, 2, device=dev, dtype=torch.float32)
    valid = torch.empty(n_tot, KP, device=dev, dtype=torch.int32)
    part_id = torch.empty(n_tot, KP, device=dev, dtype=torch.int32)
    cfeat = torch.empty(n_tot, KP, C, device=dev, dtype=torch.float32)
    stat = torch.empty(n_tot, S, 4, device=dev, dtype=torch.float32)
    _l.check(L.as_part_centers(_p(pmaps), _p(n_merged), _p(rois), _p(feats), feats.stride(0), _p(obj_img), n_tot, S, N, C, wp, KP,
                               _p(centers), _p(valid), _p(part_id), _p(cfeat), _p(stat), _sp()), 'as_part_centers')
    return dict(fg_low=fg_low, seed_map=seed_map, seed_tok=seed_tok, prot=prot, sim=sim, keep=keep, merged=merged,
                n_merged=n_merged, part_maps=pmaps, centers=centers, valid=valid, part_id=part_id, cfeat=cfeat, trace=trace,
                pending=(_Pending(n_merged), _Pending(valid)))


def assemble_parts(parts, n_per_img, gt_labels, hp, wp, num_max_keep=50):
    """Build the reference's ragged python structures (RH:244-262, RH:2024-2031) per image.  One host sync (3); the valid
    part centres of the whole batch are gathered by ONE device index_select each (coordinates, features, labels) and
    handed out as views."""
    n_merged = parts['pending'][0].get().tolist()            # host sync (3)
    valid = parts['pending'][1].get().numpy().astype(bool)   # [n_tot, KP]
    dev = parts['centers'].device
    n_tot, KP = valid.shape
    flat = np.nonzero(valid.reshape(-1))[0]                  # row-major: instance, then part -- the reference's order
    owner_glob = flat // KP
    per_obj = valid.sum(1)
    G = len(n_per_img)
    firsts = np.concatenate(([0], np.cumsum(n_per_img))).astype(np.int64)
    lab_firsts = np.concatenate(([0], np.cumsum([int(l.numel()) for l in gt_labels]))).astype(np.int64)
    lab_all = torch.cat([l.reshape(-1) for l in gt_labels]) if G else torch.zeros(0, dtype=torch.long)
    owner_img = np.repeat(np.arange(G), n_per_img)[owner_glob]
    owner_local = owner_glob - firsts[owner_img]             # index of the owning instance inside its image
    d_flat, d_lab, d_local = _upload_i32([flat, lab_firsts[owner_img] + owner_local, owner_local], dev)
    d_flat, d_lab, d_local = d_flat.long(), d_lab.long(), d_local.long()
    C = parts['cfeat'].shape[-1]
    coords_all = parts['centers'].reshape(-1, 2).index_select(0, d_flat)
    feats_all = parts['cfeat'].reshape(-1, C).index_select(0, d_flat)
    labels_all = lab_all.to(dev).index_select(0, d_lab)       # RH:2269-style gather: the label of the owning instance
    out = []
    o = c0 = 0
    for g, n in enumerate(n_per_img):
        split = per_obj[o:o + n].tolist()
        cnt = int(sum(split))
        coords, feats, labels = coords_all[c0:c0 + cnt], feats_all[c0:c0 + cnt], labels_all[c0:c0 + cnt]
        owner = d_local[c0:c0 + cnt]
        sim_fg = [parts['part_maps'][o + j, :n_merged[o + j]].unflatten(-1, (hp, wp)) if n_merged[o + j] else torch.zeros(0, 0)
                  for j in range(n)]
        if cnt == 0:
            z2 = torch.zeros(0, 2, device=dev)
            out.append(dict(semantic_centers=[z2, labels], semantic_centers_split=[], sim_fg=sim_fg,
                            semantic_centers_feat_split=[], semantic_centers_feat=[], num_parts=split,
                            semantic_centers_org=(z2.clone(), labels.clone()), corres_gts=torch.zeros(0, dtype=torch.long, device=dev)))
        else:
            c_org, l_org = coords, labels
            if cnt > num_max_keep:
                pick = torch.randperm(cnt)[:num_max_keep].to(dev)
                coords, labels = coords[pick], labels[pick]
            out.append(dict(semantic_centers=[coords, labels], semantic_centers_split=list(c_org.split(split, dim=0)),
                            sim_fg=sim_fg, semantic_centers_feat_split=list(feats.split(split, dim=0)),
                            semantic_centers_feat=feats, num_parts=split, semantic_centers_org=(c_org, l_org),
                            corres_gts=owner))
        o += n
        c0 += cnt
    return out


# --------------------------------------------------------------------------------------------------- A15
def update_fg_maps(map_cos_fg, feats, coords, num_parts, inst_fg_feat, inst_bg_feat, rois, n_per_img, hp, wp, rng,
                   pos_mask_thr=0.6, refine_times=3, tau=0.85):
    """Second-round aggregation, RH:2737-2844 (``update_fg_map`` + ``update_fg_map_single_v3`` + ``extract_bg_coords`` +
    ``get_refined_similarity_input_map``) for the whole batch.
    map_cos_fg [n_tot,H,W] first-round maps; feats [n_img,N,C] token-major; coords: per image [P_i,2] part centres (x, y)
    pixels; num_parts: per image, parts per instance; inst_fg_feat / inst_bg_feat: per image [n_i+1,C,1,1] / [n_i,C,1,1]
    (``seed_pseudo_gt`` outputs); rois [n_tot,4].  -> (maps [n_tot,H,W] fp32, masks [n_tot,H,W] uint8), instances flat.
    The prototypes are a few dozen vectors assembled with torch glue (the reference samples them with mmcv's point_sample =
    ``F.grid_sample``, kept as is, quirks included); the affinity / refinement / fusion passes are the first round's kernels."""
    L = _l.load()
    dev = feats.device
    n_img, N, C = feats.shape
    n_tot, H, W = map_cos_fg.shape
    plan = _seed_plan(n_per_img, dev, 0.2, 0.1)
    grp = plan['grp']
    S = max(n_per_img) + 2
    obj_img = instance_image_index(n_per_img, dev)
    fmap = feats.view(n_img, hp, wp, C).permute(0, 3, 1, 2)             # [n_img,C,hp,wp] view for grid_sample
    img_size = torch.tensor([float(W), float(H)], device=dev)
    # background supplement: 5 random pixels where no first-round map responds (RH:2826-2829)
    resp = torch.zeros(n_img, H, W, device=dev, dtype=map_cos_fg.dtype).index_add_(0, obj_img.long(), map_cos_fg)
    nz = torch.nonzero(resp == 0)                                         # host sync: [K,3] (image, row, col), row-major
    per_img = torch.bincount(nz[:, 0], minlength=n_img).tolist()
    starts = np.concatenate(([0], np.cumsum(per_img)))
    pick = []
    for g in range(n_img):
        cnt = per_img[g]
        if cnt == 0:
            pick.append(None)
            continue
        head = rng.randperm_head((g, 2, 0), cnt, 5).tolist()
        sel = list(head)
        while len(sel) < 5:
            sel += head[:5 - len(sel)]
        pick.append([int(starts[g]) + k for k in sel])
    flat = [k for p_ in pick if p_ is not None for k in p_]
    d_pick, = _upload_i32([flat], dev)
    chosen = nz.index_select(0, d_pick.long())[:, 1:].float() if flat else nz[:0, 1:].float()
    protos = torch.zeros(grp.G, S, C, device=dev, dtype=torch.float32)
    bg_protos = torch.empty(n_tot, 1, C, device=dev, dtype=torch.float32)
    o = c0 = 0
    for g, n in enumerate(n_per_img):
        fg_f = inst_fg_feat[g].reshape(n + 1, C)
        bg_protos[o:o + n, 0] = inst_bg_feat[g].reshape(n, C)
        pts = (coords[g].to(dev).float() / img_size)[None, :, None, :]     # [1,P,1,2] in [0,1]: mmcv point_sample
        sc = F.grid_sample(fmap[g:g + 1], pts * 2.0 - 1.0, align_corners=False)[0, :, :, 0].t()      # [P,C]
        for j, s_ in enumerate(sc.split(list(num_parts[g]), dim=0)):
            protos[g, j] = torch.mean(s_) * 0.5 + fg_f[j] * 0.5             # scalar mean of ALL elements, as the reference
        protos[g, n] = fg_f[n]
        if pick[g] is None:
            idx = torch.ones(5, 2, device=dev)
        else:
            idx = chosen[c0:c0 + 5]
            c0 += 5
        bc = ((idx + 0.5) / torch.tensor([float(H), float(W)], device=dev)).flip(0)[None, None]        # [1,1,5,2], (row, col) fed as (x, y)
        protos[g, n + 1] = F.grid_sample(fmap[g:g + 1], bc * 2.0 - 1.0, mode='bilinear', align_corners=False)[0, :, 0].mean(-1)
        o += n
    # affinity, refinement loop (threshold -> weighted centroid -> affinity -> box mask / winner-take-all), fusion
    bg_low = cosine_maps(feats, obj_img, bg_protos).reshape(n_tot, N)
    cur = cosine_maps(feats, grp.d_img, protos)
    GS = grp.G * S
    wsum = torch.empty(GS, device=dev, dtype=torch.float32)
    nb = L.as_weighted_centroid_workspace(grp.G, S, N, C)
    ws = _ws(nb, dev)
    fg_low = torch.empty(n_tot, N, device=dev, dtype=torch.float32)
    for r in range(refine_times):
        _l.check(L.as_refine_threshold(_p(cur), GS, N, float(tau), _p(wsum), _sp()), 'as_refine_threshold')
        centroid = torch.empty(grp.G, S, C, device=dev, dtype=torch.float32)
        _l.check(L.as_weighted_centroid(_p(feats), feats.stride(0), _p(grp.d_img), _p(cur), _p(wsum), grp.G, S, N, C,
                                        _p(centroid), _p(ws), nb, _sp()), 'as_weighted_centroid')
        cur = cosine_maps(feats, grp.d_img, centroid)
        _l.check(L.as_refine_select(_p(cur), grp.G, S, N, wp, _p(grp.d_first), _p(grp.d_n), _p(rois),
                                    int(r == refine_times - 1), 2, _p(fg_low), None, _sp()), 'as_refine_select')
    maps = torch.empty(n_tot, H, W, device=dev, dtype=torch.float32)
    mask = torch.empty(n_tot, H, W, device=dev, dtype=torch.uint8)
    stats = torch.empty(n_tot * 3, device=dev, dtype=torch.int32)
    _l.check(L.as_fuse_instance_maps(_p(fg_low), _p(bg_low), n_tot, hp, wp, float(pos_mask_thr), _p(maps), None, _p(mask),
                                     _p(stats), _sp()), 'as_fuse_instance_maps')
    # an instance whose refined map vanished keeps its first-round map (RH:2754-2756)
    drop = (maps.flatten(1).sum(1) == 0)[:, None, None]
    old_mask = (map_cos_fg > map_cos_fg.flatten(1).max(1)[0][:, None, None] * pos_mask_thr).to(torch.uint8)
    return torch.where(drop, map_cos_fg, maps), torch.where(drop, old_mask, mask)


def cam_minmax(lows, hp, wp):
    """min / max of the x16 bilinear up-sampling of every [hp,wp] map.  lows [n_maps, N] -> [n_maps, 2]."""
    L = _l.load()
    n_maps = lows.shape[0]
    mm = torch.empty(n_maps, 2, device=lows.device, dtype=torch.float32)
    scratch = torch.empty(n_maps * 2, device=lows.device, dtype=torch.int32)
    _l.check(L.as_cam_minmax(_p(lows), n_maps, hp, wp, _p(mm), _p(scratch), _sp()), 'as_cam_minmax')
    return mm

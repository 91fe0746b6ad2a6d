"""Host half of the seed sampling (RH:343-371) on the CPU: ``select_seed_candidates`` + ``KeyedRng`` against the oracle's
``sample_points`` (itself pinned to the reference).  The device's part -- the k-th candidate of an item in row-major order
(``as_norm_select``) -- is emulated with ``nonzero``."""
import numpy as np
import torch

from attentionshift_b200 import attention_shift as AS
from oracle import attnshift as O


def _case(seed, n, H, W, P, force_short):
    g = torch.Generator().manual_seed(seed)
    maps = torch.rand(n, H, W, generator=g)
    if force_short:                      # instance 1: a handful of pixels above the foreground threshold
        maps[1] = torch.rand(H, W, generator=g) * 0.1
        maps[1, 3, 5] = 1.0
        maps[1, 3, 6] = 0.9
    an = O.norm_maps(maps)
    gt_points = torch.tensor([[5.7, 2.2], [6.9, 3.1], [1.0, 8.5]][:n])
    return an, gt_points


def _device_lookup(an, kinds, inst, thr, item, k):
    n = an.shape[0]
    if kinds[item] == 1:
        cand = (an[inst[item]] >= thr[item]).nonzero()
    elif kinds[item] == 0:
        cand = (an[inst[item]] < thr[item]).nonzero()
    else:
        cand = (an.mean(0) < thr[item]).nonzero()
    return cand[k].flip(-1)              # (row, col) -> (x, y)


def _ours(an, gt_points, rng, P, thr_pos=0.2, thr_neg=0.1):
    n = an.shape[0]
    # one image: items in the reference's draw order (bg instances, fg instances, supplement), rows as in _Groups
    kinds = np.array([0] * n + [1] * n + [2])
    inst = np.array(list(range(n)) * 2 + [0])
    row = np.array([n + 1 + j for j in range(n)] + list(range(n)) + [n])
    thr = np.array([thr_neg] * n + [thr_pos] * n + [thr_neg])
    keys = [(0, 0, int(r)) for r in row]

    def total(i, t):
        if kinds[i] == 1:
            return int((an[inst[i]] >= t).sum())
        return int(((an[inst[i]] if kinds[i] == 0 else an.mean(0)) < t).sum())
    totals = np.array([total(i, thr[i]) for i in range(len(kinds))])
    for i in range(len(kinds)):          # threshold doubling for background items (RH:360-364)
        while kinds[i] != 1 and totals[i] < P:
            thr[i] *= 2
            totals[i] = total(i, thr[i])
    ks = rng.draws(keys, P).astype(np.int64) % np.maximum(totals, 1)[:, None]
    pts, sel_item, sel_k, sel_dst = AS.select_seed_candidates(kinds, inst, row, totals, ks, P, 2 * n + 1,
                                                             lambda: gt_points.numpy())
    flat = torch.from_numpy(pts.copy()).view(-1, 2)
    for it, k, d in zip(sel_item, sel_k, sel_dst):
        flat[d] = _device_lookup(an, kinds, inst, thr, it, int(k)).int()
    return flat.view(2 * n + 1, P, 2)


def _oracle(an, gt_points, rng, P, thr_pos=0.2, thr_neg=0.1):
    n = an.shape[0]
    hook = lambda key: torch.manual_seed(rng.seed_for(key))
    bg = O.sample_points(an, thr=thr_neg, num_points=P, hook=hook, keys=[(0, 0, n + 1 + j) for j in range(n)])
    fg = O.sample_points(an, thr=thr_pos, num_points=P, is_pos=True, gt_points=gt_points, hook=hook, keys=[(0, 0, j) for j in range(n)])
    sp = O.sample_points(an.mean(0, keepdim=True), thr=thr_neg, num_points=P, hook=hook, keys=[(0, 0, n)])
    return torch.cat((fg, sp, bg), dim=0)          # rows: fg 0..n-1, supplement n, bg n+1..2n


def test_selection_equals_oracle_sampling():
    rng = AS.KeyedRng(7)
    for seed, force_short in ((0, False), (1, True), (2, False), (3, True)):
        an, gtp = _case(seed, 3, 24, 40, 20, force_short)
        ours = _ours(an, gtp, rng, 20)
        ref = _oracle(an, gtp, rng, 20)
        assert torch.equal(ours.long(), ref.long()), (seed, force_short)
    an, gtp = _case(1, 3, 24, 40, 20, True)
    assert int((an[1] >= 0.2).sum()) < 20          # the short-foreground branch (GT point repeated, in (y, x) order) ran

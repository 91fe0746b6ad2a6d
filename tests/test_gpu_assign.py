"""Point-token <-> GT matching on the device (``as_hungarian_points``) against scipy's
``linear_sum_assignment`` -- the solver the reference calls on the host (hungarian_point_assigner.py:95-99) -- on the same
cost matrices, and against the stored outputs of the reference's HungarianPointAssigner + PointPseudoSampler."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _solve_device(costs, P):
    """costs: list of [P, G_i] float32 arrays (assign()'s layout) -> per image (pos_inds, pos_gt, status)."""
    from attentionshift_b200 import lib as _l
    n_g = [c.shape[1] for c in costs]
    first = np.concatenate([[0], np.cumsum(n_g)[:-1]]).astype(np.int32)
    tot = int(sum(n_g))
    flat = np.concatenate([np.ascontiguousarray(c.T).reshape(-1) for c in costs]) if tot else np.zeros(0, np.float32)
    d_cost = torch.from_numpy(flat.astype(np.float32)).to(DEV)
    d_first = torch.from_numpy(first).to(DEV)
    d_n = torch.tensor(n_g, dtype=torch.int32, device=DEV)
    out = torch.full((3, max(tot, len(costs))), -7, dtype=torch.int32, device=DEV)
    L = _l.load()
    _l.check(L.as_hungarian_points(_l.ptr(d_cost), _l.ptr(d_first), _l.ptr(d_n), len(costs), P, max(n_g), _l.ptr(out[0]),
                                   _l.ptr(out[1]), _l.ptr(out[2]), _l.stream_ptr()), 'as_hungarian_points')
    out = out.cpu().numpy()
    res = []
    for i, g in enumerate(n_g):
        k = min(P, g)
        res.append((out[0, first[i]:first[i] + k], out[1, first[i]:first[i] + k], out[2, i] if g else 0))
    return res


def _scipy(c):
    from scipy.optimize import linear_sum_assignment
    rows, cols = linear_sum_assignment(c.astype(np.float64))
    o = np.argsort(rows)
    return rows[o], cols[o]


@pytest.mark.parametrize('kind', ['normal', 'small_ints', 'constant', 'focal_like'])
def test_matches_scipy(kind):
    """Same matching (not just the same cost): continuous costs, tie-heavy integer costs, a constant matrix (scipy returns the
    identity there by construction) -- proposals x GTs both wider and taller than square, 1 x 1, ragged batch with an empty
    image."""
    rng = np.random.default_rng(5)
    for P in (1, 7, 100, 130):
        gs = [1, 3, 7, 20, 0, 100, 150, P]
        costs = []
        for g in gs:
            if kind == 'normal':
                c = rng.standard_normal((P, g))
            elif kind == 'small_ints':
                c = rng.integers(0, 4, size=(P, g)).astype(np.float64)
            elif kind == 'constant':
                c = np.full((P, g), 1.5)
            else:
                c = -np.log(rng.random((P, g)) + 1e-12) * 0.25 + 10 * rng.random((P, g))
            costs.append(c.astype(np.float32))
        got = _solve_device(costs, P)
        for c, (pi, pg, st) in zip(costs, got):
            if c.shape[1] == 0:
                continue
            rows, cols = _scipy(c)
            assert st == 0
            assert np.array_equal(pi, rows) and np.array_equal(pg, cols), (kind, P, c.shape)


def test_invalid_costs_flagged():
    """NaN / -inf entries (scipy raises ValueError) and matrices without a finite matching: status 1, indices stay in range."""
    P = 10
    a = np.random.default_rng(0).random((P, 3)).astype(np.float32)
    nan = a.copy(); nan[2, 1] = np.nan
    ninf = a.copy(); ninf[0, 0] = -np.inf
    infeasible = a.copy(); infeasible[:, 2] = np.inf
    got = _solve_device([a, nan, ninf, infeasible], P)
    assert [int(g[2]) for g in got] == [0, 1, 1, 1]
    for pi, pg, _ in got:
        assert pi.min() >= 0 and pi.max() < P and pg.min() >= 0 and pg.max() < 3


def test_reference_golden(golden_dir):
    """The device route of ``AttnShiftRoIHead.match_points`` against the outputs of the unmodified reference classes
    (tests/golden/point_assigner.pt, made by tests/golden/make_golden.py)."""
    from attentionshift_b200 import assigner as A
    g = torch.load(os.path.join(golden_dir, 'point_assigner.pt'))
    for c in g['cases']:
        pos, pos_gt, st = A.hungarian_point_assign_device(c['pred'][None].to(DEV), c['cls'][None].to(DEV), [c['gt_points']],
                                                          [c['gt_labels']], [c['img_wh']], 1.0, 10.0, want_status=True)
        assert int(st[0]) == 0
        assert torch.equal(pos[0].cpu(), c['pos_inds']) and torch.equal(pos_gt[0].cpu(), c['pos_gt'])


def test_batched_equals_host_route():
    """A ragged batch (0 .. 25 GTs per image, 100 point tokens) through both routes of the head."""
    from attentionshift_b200 import assigner as A
    g = torch.Generator().manual_seed(2)
    B, P = 6, 100
    n_g = [3, 0, 1, 25, 7, 2]
    reg = torch.rand(B, P, 2, generator=g)
    cls = torch.randn(B, P, 20, generator=g)
    pts = [torch.rand(n, 2, generator=g) * torch.tensor([640., 480.]) for n in n_g]
    lab = [torch.randint(0, 20, (n,), generator=g) for n in n_g]
    wh = [(640, 480)] * B
    pos, pgt = A.hungarian_point_assign_device(reg.to(DEV), cls.to(DEV), pts, lab, wh)
    for i in range(B):
        p_h, g_h = A.hungarian_point_assign(reg[i], cls[i], pts[i], lab[i], wh[i])
        assert torch.equal(pos[i].cpu(), p_h) and torch.equal(pgt[i].cpu(), g_h), i

"""Device mean-shift (as_grid_seeds + as_mean_shift) against the CPU oracle on identical seeded inputs."""
import pytest
import torch

from attentionshift_b200.synthetic import structured_scene
from oracle import attnshift as O

pytestmark = pytest.mark.gpu


def _f64_margin_trace(prot, feats, n_shift, tau=0.1, temp=0.1):
    """float64 re-run of RH:830-854 that also returns the relative top-2 weight margin per token."""
    prot, feats = prot.double(), feats.double()
    out = []
    F = torch.nn.functional
    for _ in range(n_shift):
        sim = F.cosine_similarity(prot[:, :, None], feats[:, None], dim=-1)
        w = F.softmax(sim / (temp * tau), dim=-1)
        top, topi = w.topk(2, dim=1)
        # fp32 evaluation noise of a weight is ~ eps32 * |logit| (the softmax temperature amplifies the rounding of
        # the cosine), so a top-2 gap only counts as "clear" when it exceeds that noise of both contenders
        lg = (sim / (temp * tau)).abs().expand_as(w)
        noise = 4e-6 * (lg.gather(1, topi[:, :1]) + lg.gather(1, topi[:, 1:2]))[:, 0] + 1e-4
        margin = (top[:, 0] - top[:, 1]) / top[:, 0].clamp_min(1e-300) - noise
        assign = w.argmax(1, keepdim=True)
        onehot = torch.zeros_like(w).scatter_(1, assign, 1.0)
        prot = torch.matmul(w * onehot, feats)
        s2 = F.cosine_similarity(prot[:, :, None], feats[:, None], dim=-1)
        cnt = onehot.sum(-1)
        tau = (1 - torch.where(cnt >= 1, (s2 * onehot).sum(-1) / cnt.clamp_min(1), torch.zeros_like(cnt))).clamp(1e-10).unsqueeze(-1)
        out.append((assign[:, 0], margin, top[:, 0]))
    return out


@pytest.mark.parametrize('impl', ['fp32', 'tc', 'fused'])
@pytest.mark.parametrize('hp,c,n_obj,S,n_shift,seed', [(14, 32, 2, 20, 5, 11), (28, 64, 3, 20, 10, 3), (28, 64, 3, 16, 5, 5),
                                                       (20, 48, 5, 4, 2, 7), (20, 64, 5, 4, 2, 7), (20, 128, 5, 4, 2, 7),
                                                       (64, 768, 3, 16, 5, 1)])
def test_mean_shift_vs_oracle(hp, c, n_obj, S, n_shift, seed, impl):
    if impl != 'fp32' and c % 64:
        pytest.skip('tensor-core path needs C % 64 == 0')
    if impl == 'fused' and c % 128:
        pytest.skip('persistent kernel needs C % 128 == 0 (128-channel accumulator blocks in TMEM)')
    from attentionshift_b200 import ops
    sc = structured_scene(hp, hp, c, n_obj, seed=seed, noise=0.4)
    # foreground seed maps: the instance disks (owner labels) on the patch grid
    maps = torch.stack([((sc['labels'] == 2 * i + 1) | (sc['labels'] == 2 * i + 2)).float() for i in range(n_obj)])
    if seed == 7:  # degenerate seed sets
        maps[1] = 0            # exercise the "no positive patch -> box centre" branch (RH:1799)
        maps[2, :] = 0
        maps[2, 3, 4] = 1      # and the repeat-fill branch (RH:1795)
    trace = []
    o_prot, o_sim = O.mean_shift_from_maps(maps, sc['vit_feat'], sc['rois'], n_shift=n_shift, n_points=S, trace=trace)

    dev = 'cuda'
    feats = sc['vit_feat'].permute(1, 2, 0).reshape(1, hp * hp, c).contiguous().to(dev)
    obj_img = torch.zeros(n_obj, dtype=torch.int32, device=dev)
    rois = sc['rois'].to(dev)
    tok, proto0 = ops.grid_seeds(maps.reshape(n_obj, -1).to(dev), feats, obj_img, rois, hp, S)
    sel = O.grid_seed_coords(maps, sc['rois'], 0.35, S)
    assert torch.equal(tok.cpu().long(), sel[..., 0] * hp + sel[..., 1])          # seed indices: bit-exact
    prot, sim, tr = ops.mean_shift(proto0, feats, obj_img, rois, hp, hp, n_shift, want_trace=True, n_per_img=[n_obj],
                                   impl=impl)
    torch.cuda.synchronize()

    # hard assignments: exact wherever the float64 margin is not at rounding level
    bm = O.box_to_mask(sc['rois'] // 16, (hp, hp), default=0)
    feats_m = (sc['vit_feat'][None] * bm[:, None]).flatten(-2).transpose(1, 2)
    tokf = sc['vit_feat'].permute(1, 2, 0)
    p0 = tokf[sel[..., 0].flatten(), sel[..., 1].flatten()].unflatten(0, sel.shape[:2])
    ref64 = _f64_margin_trace(p0, feats_m, n_shift)
    total, bad_clear = 0, 0
    for it in range(n_shift):
        got = tr[it].cpu().long()
        want = trace[it]
        mism = got != want
        clear = (ref64[it][1] > 0) & (ref64[it][2] > 1e-30)
        bad_clear += int((mism & clear & (ref64[it][0] == want)).sum())
        total += int(mism.sum())
    assert bad_clear == 0, f'{bad_clear} assignment flips with a clear float64 margin'
    assert total <= 0.002 * n_shift * n_obj * hp * hp + 2, f'{total} rounding-level flips'
    # prototypes / similarity maps: 1e-3 relative (north_star tolerance) -- measured far tighter
    # (atol is tied to the tensor's scale: near-zero prototype components carry the summation-order noise of N terms)
    torch.testing.assert_close(prot.cpu().flatten(0, 1), o_prot, rtol=1e-3, atol=1e-4 * o_prot.abs().max().item())
    torch.testing.assert_close(sim.cpu().unflatten(-1, (hp, hp)).flatten(0, 1), o_sim, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize('n_img,hp,c,S', [(3, 32, 128, 16), (11, 64, 128, 12), (2, 24, 384, 20)])
def test_fused_matches_tc_multi_image(n_img, hp, c, S):
    """The persistent kernel against the multi-launch tensor-core variant on a ragged batch (1-3 instances per image; with
    11 images of 4096 tokens the 16-CTA groups take more than one round over the 148 SMs)."""
    from attentionshift_b200 import ops
    dev = 'cuda'
    g = torch.Generator().manual_seed(100 + n_img)
    N = hp * hp
    n_per_img = [1 + (i % 3) for i in range(n_img)]
    scenes = [structured_scene(hp, hp, c, n_per_img[i], seed=50 + i, noise=0.4) for i in range(n_img)]
    feats = torch.stack([s['vit_feat'].permute(1, 2, 0).reshape(N, c) for s in scenes]).contiguous().to(dev)
    obj_img = torch.cat([torch.full((n,), i, dtype=torch.int32) for i, n in enumerate(n_per_img)]).to(dev)
    rois = torch.cat([s['rois'] for s in scenes]).to(dev)
    maps = torch.cat([torch.stack([((s['labels'] == 2 * j + 1) | (s['labels'] == 2 * j + 2)).float() for j in range(n)])
                      for s, n in zip(scenes, n_per_img)]).reshape(-1, N).to(dev)
    _, proto0 = ops.grid_seeds(maps, feats, obj_img, rois, hp, S)
    out = {}
    for impl in ('tc', 'fused'):
        out[impl] = ops.mean_shift(proto0, feats, obj_img, rois, hp, hp, 4, want_trace=True, n_per_img=n_per_img, impl=impl)
    torch.cuda.synchronize()
    (p_a, s_a, t_a), (p_b, s_b, t_b) = out['tc'], out['fused']
    agree = (t_a == t_b).float().mean().item()
    assert agree > 0.999, agree
    scale = p_a.abs().max().item()
    assert (p_a - p_b).abs().max().item() <= 2e-3 * scale
    assert (s_a - s_b).abs().max().item() <= 2e-3
    # run-to-run determinism of the ordered reductions
    again = ops.mean_shift(proto0, feats, obj_img, rois, hp, hp, 4, want_trace=True, n_per_img=n_per_img, impl='fused')
    assert torch.equal(again[0], p_b) and torch.equal(again[1], s_b) and torch.equal(again[2], t_b)

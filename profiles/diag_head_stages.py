"""Stage-by-stage comparison of the device head against the oracle on the DEVICE backbone's outputs (debugging aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import torch
spec = importlib.util.spec_from_file_location('tp', os.path.join(os.path.dirname(__file__), '..', 'tests', 'test_gpu_pipeline.py'))
tp = importlib.util.module_from_spec(spec); spec.loader.exec_module(tp)
from oracle import attnshift as O

embed, heads, depth, img, n_pt, n_obj, S, iters, B = 384, 6, 7, 448, 100, 2, 20, 3, 4
if len(sys.argv) > 1 and sys.argv[1] == 'cfg2':
    embed, heads, depth, img, n_obj, S, iters, B = 768, 12, 12, 1024, 3, 16, 5, 1
scale = float(os.environ.get('QKV_SCALE', 4.0))
hp = img // 16
sd, bb, head, rng, x, gt_points, pos_inds, gt_index, labels = tp._setup(embed, heads, depth, img, n_pt, B, n_obj, 21, scale, S, iters)
out_b, res = tp._device_pass(bb, head, x, gt_points, pos_inds, gt_index, labels, hp)
last = head._last
rm = last['refined']
attns7 = [a.cpu() for a in out_b['attns'][-7:]]
lf = out_b['last_feat'].cpu()
print('pts', tuple(rm['pts'].shape), 'centroid', tuple(rm['centroid'].shape), 'fg_low', tuple(rm['fg_low'].shape))
o0 = 0
for i in range(B):
    n = len(pos_inds[i])
    rows = O.rollout_rows([a[i:i + 1] for a in attns7], n_pt)[0]
    low, up = O.cams_from_rollout(rows, pos_inds[i], n_pt, hp, hp)
    d_rows = last['rows'][i].cpu()
    print(f'img {i}: rollout rows max rel diff {float(((d_rows[..., :rows.shape[-1]] - rows).abs() / rows.abs().clamp_min(1e-12)).max()):.2e}')
    boxes = torch.stack([torch.cat([O.bbox_from_cam(up[l, j].clone(), gt_points[i][j], 0.2, 0.5, (img, img))[0] for j in range(n)]) for l in range(7)])
    pb = boxes[gt_index[i], torch.arange(n)]
    print('   boxes equal', torch.equal(last['boxes'][:, o0:o0 + n].cpu(), boxes))
    vit = lf[i, 1:].t().unflatten(-1, (hp, hp)).contiguous()
    attn_sel = up[gt_index[i], torch.arange(n)]
    hook = lambda key: torch.manual_seed(rng.seed_for(key))
    fg, bg, pts_fg, pts_bg, f_fg, f_bg = O.refined_maps(attn_sel, vit, pb, num_points=20, refine_times=2, obj_tau=0.85, gt_points=gt_points[i], hook=hook, img=i)
    d_pts = rm['pts'][i].cpu()
    print('   fg seed points equal', torch.equal(d_pts[:n + 1].long(), pts_fg.long()), ' bg seed points equal', torch.equal(d_pts[n + 1:2 * n + 1].long(), pts_bg.long()))
    if not torch.equal(d_pts[:n + 1].long(), pts_fg.long()):
        ne = (d_pts[:n + 1].long() != pts_fg.long()).any(-1)
        print('     differing fg points per row', ne.sum(-1).tolist(), 'first', d_pts[:n + 1][ne][:3].tolist(), pts_fg[ne][:3].tolist())
    cen = rm['centroid'][i].cpu()
    print('   fg centroid rel diff', float((cen[:n + 1] - f_fg.flatten(1)).abs().max() / f_fg.abs().max()), ' bg centroid rel diff',
          float((cen[n + 1:2 * n + 1] - f_bg.flatten(1)).abs().max() / f_bg.abs().max()))
    dm = res['map_cos_fg'][i].cpu()
    print('   map_cos_fg max abs diff', float((dm - fg[-1]).abs().max()), ' frac > 1e-3', float(((dm - fg[-1]).abs() > 1e-3 * fg[-1].abs() + 1e-5).float().mean()))
    # refinement on the oracle side from the DEVICE's seed points (isolates everything after the sampling)
    sim_fg, _ = O.refined_similarity(d_pts[:n + 1].float(), vit, pb, 2, 0.85, is_select=True)
    sim_bg, _ = O.refined_similarity(d_pts[n + 1:2 * n + 1].float(), vit, pb, 2, 0.85)
    print('   with device seed points: fg_low max abs diff', float((rm['fg_low'][o0:o0 + n].cpu().unflatten(-1, (hp, hp)) - sim_fg[-1][:n]).abs().max()),
          ' bg_low', float((rm['bg_low'][o0:o0 + n].cpu().unflatten(-1, (hp, hp)) - sim_bg[-1]).abs().max()))
    o0 += n

"""Seed-candidate counts (RH:343-352) on the device vs the oracle, per item, on the device backbone's outputs (debugging aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import torch
spec = importlib.util.spec_from_file_location('tp', os.path.join(os.path.dirname(__file__), '..', 'tests', 'test_gpu_pipeline.py'))
tp = importlib.util.module_from_spec(spec); spec.loader.exec_module(tp)
from oracle import attnshift as O
from attentionshift_b200 import attention_shift as AS

embed, heads, depth, img, n_pt, n_obj, S, iters, B = 384, 6, 7, 448, 100, 2, 20, 3, 4
scale = float(os.environ.get('QKV_SCALE', 4.0))
hp = img // 16
sd, bb, head, rng, x, gt_points, pos_inds, gt_index, labels = tp._setup(embed, heads, depth, img, n_pt, B, n_obj, 21, scale, S, iters)
out_b, res = tp._device_pass(bb, head, x, gt_points, pos_inds, gt_index, labels, hp)
last = head._last
attns7 = [a.cpu() for a in out_b['attns'][-7:]]
n_per_img = [n_obj] * B
dev = 'cuda'
obj_img = AS.instance_image_index(n_per_img, dev)
gi = torch.cat(gt_index).to(dev)
ar = torch.arange(B * n_obj, device=dev)
cams = last['cams']
mm = AS.cam_maps(last['rows'], obj_img, torch.cat(pos_inds).to(dev).int(), hp, hp)[1]
begun = AS.refined_maps_begin(cams[gi, ar].contiguous(), mm[gi, ar].contiguous(), n_per_img, hp, hp)
lv = begun['pending'].get().numpy()
print('device counts [levels, items]:\n', lv)
print('kinds', begun['kinds'])
it = 0
for i in range(B):
    rows = O.rollout_rows([a[i:i + 1] for a in attns7], n_pt)[0]
    low, up = O.cams_from_rollout(rows, pos_inds[i], n_pt, hp, hp)
    sel = up[gt_index[i], torch.arange(n_obj)]
    an = O.norm_maps(sel)
    d_low = cams[gi, ar][i * n_obj:(i + 1) * n_obj].cpu().unflatten(-1, (hp, hp))
    print(f'img {i}: low-res CAM equal to oracle: {torch.equal(d_low, low[gt_index[i], torch.arange(n_obj)])}, max rel diff '
          f'{float(((d_low - low[gt_index[i], torch.arange(n_obj)]).abs() / low[gt_index[i], torch.arange(n_obj)].abs()).max()):.2e}')
    d_mm = mm[gi, ar][i * n_obj:(i + 1) * n_obj].cpu()
    print('   minmax device', d_mm.flatten().tolist(), ' oracle', [(float(s.min()), float(s.max())) for s in sel])
    for j in range(n_obj):
        print(f'   bg item {it}: oracle', [int((an[j] < 0.1 * f).sum()) for f in (1, 2, 4, 8)], 'device', lv[:, it].tolist()); it += 1
    for j in range(n_obj):
        print(f'   fg item {it}: oracle', int((an[j] >= 0.2).sum()), 'device', lv[:, it].tolist()); it += 1
    m = an.mean(0)
    print(f'   supp item {it}: oracle', [int((m < 0.1 * f).sum()) for f in (1, 2, 4, 8)], 'device', lv[:, it].tolist()); it += 1

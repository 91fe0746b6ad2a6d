"""A/B of the fused mean-shift kernel's group barrier: global counter + polling under a cooperative launch (AS_MS_CLUSTER=0)
vs one thread-block cluster per image with barrier.cluster (AS_MS_CLUSTER=1).  cfg2 shapes, structured scene.
Run once per setting (the switch is read once per process)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import lib, ops
from attentionshift_b200.synthetic import structured_scene

dev = 'cuda'
n_img, hp, C, n_obj, S, iters = 8, 64, 768, 3, 16, 5
N = hp * hp
scenes = [structured_scene(hp, hp, C, n_obj, seed=10 + i, noise=0.4) for i in range(n_img)]
feats = torch.stack([s['vit_feat'].permute(1, 2, 0).reshape(N, C) for s in scenes]).contiguous().to(dev)
obj_img = torch.arange(n_img, dtype=torch.int32).repeat_interleave(n_obj).to(dev)
rois = torch.cat([s['rois'] for s in scenes]).to(dev)
maps = torch.cat([torch.stack([((s['labels'] == 2 * j + 1) | (s['labels'] == 2 * j + 2)).float() for j in range(n_obj)])
                  for s in scenes]).reshape(-1, N).to(dev)
_, proto0 = ops.grid_seeds(maps, feats, obj_img, rois, hp, S)
npi = [n_obj] * n_img
full = rois.clone(); full[:, 0] = 0; full[:, 1] = 0; full[:, 2] = hp * 16; full[:, 3] = hp * 16      # boxes = whole image (the bench's case)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for name, r in (('disk boxes', rois), ('full-image boxes', full)):
    ms = timeit(lambda: ops.mean_shift(proto0, feats, obj_img, r, hp, hp, iters, n_per_img=npi, impl='fused'))
    print(f'AS_MS_CLUSTER={os.environ.get("AS_MS_CLUSTER", "default")}  {name:18s} {ms:7.4f} ms   co-resident clusters '
          f'{lib.load().as_mean_shift_fused_occupancy()}')

"""Target of the ncu capture of the mean-shift kernels: the attention-shift loop alone at the cfg2 shapes (8 images x 4096 tokens
x 768 channels, 3 instances x 16 seeds, 5 iterations), launched three times.  usage: python profiles/ncu_ms_target.py [v2|fused]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import ops
from attentionshift_b200.synthetic import structured_scene

impl = sys.argv[1] if len(sys.argv) > 1 else 'v2'
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
for _ in range(3):
    ops.mean_shift(proto0, feats, obj_img, rois, hp, hp, iters, n_per_img=[n_obj] * n_img, impl=impl)
torch.cuda.synchronize()
print('done')

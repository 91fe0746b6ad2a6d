"""Slow-step forensics: allocator counters and host time per phase for every step of a bench-like loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from attentionshift_b200 import ops, attention_shift as AS

cfg = dict(bench.WORKLOAD)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
bb, head = bench.build_models(cfg, dev)
inputs = bench.make_inputs(cfg, 0)
img_dev = inputs[0].to(dev)
marks = []
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t = time.perf_counter()
        r = f(*a, **k)
        marks.append((name, time.perf_counter() - t))
        return r
    setattr(mod, name, g)
for nm in ('rollout_rows', 'cam_maps', 'refined_maps_begin', 'cam_bbox', 'refined_maps', 'mask_points_begin', 'semantic_parts', 'mask_points', 'assemble_parts'):
    wrap(AS, nm)
def stats():
    s = torch.cuda.memory_stats()
    h = torch.cuda.host_memory_stats() if hasattr(torch.cuda, 'host_memory_stats') else {}
    return (s['segment.all.allocated'], s['segment.all.freed'], s['num_alloc_retries'], s.get('num_device_alloc', -1), s.get('num_device_free', -1),
            h.get('num_host_alloc', -1), h.get('num_host_free', -1))
for _ in range(3):
    bench.one_step(bb, head, img_dev, inputs, False)
torch.cuda.synchronize()
prev = stats()
for i in range(120):
    marks.clear()
    t0 = time.perf_counter()
    out = bb(img_dev)
    t1 = time.perf_counter()
    bench_res = None
    hp = 64
    _, gt_points, pos_inds, gt_index, labels = inputs
    res = head.seed_pseudo_gt(out['feature'], None, None, None, None, vit_feat=out['last_feat'][:, 1:].unflatten(1, (hp, hp)).permute(0, 3, 1, 2),
                              point_cls=out['outputs_class'], point_reg=out['outputs_coord'], attns=out['attns'], gt_points=gt_points,
                              gt_points_labels=labels, return_mask=False, pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10,
                              corr_size=21, obj_tau=0.85, pos_inds=pos_inds, gt_index=gt_index)
    t2 = time.perf_counter()
    cur = stats()
    if t2 - t0 > 0.035:
        print(f'step {i}: {1e3 * (t2 - t0):.1f} ms  bb host {1e3 * (t1 - t0):.1f}  stats delta {tuple(c - p for c, p in zip(cur, prev))}')
        print('   ', ' '.join(f'{n}={1e3 * t:.1f}' for n, t in marks))
    prev = cur
torch.cuda.synchronize()
print('done', stats())

"""Why is one bench step in ~10 twice as long as the others?  Runs 150 steps of the cfg2 workload and, per step, records the host
time spent in the backbone call and in seed_pseudo_gt, the device time of the step (CUDA events) and the SM clock (NVML, every
10th step); prints the distribution and the slow steps with their neighbours.  usage: python profiles/step_outliers.py [steps]"""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
cfg = dict(bench.CONFIGS['cfg2'])
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
bb, head = bench.build_models(cfg, dev)
inputs = bench.make_inputs(cfg, 0)
img = inputs[0].to(dev)
for _ in range(5):
    bench.one_step(bb, head, img, inputs, False)
torch.cuda.synchronize()
if os.environ.get('AS_NO_GC'):
    gc.collect(); gc.disable()
rows = []
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
_, gt_points, pos_inds, gt_index, labels = inputs
hp = wp = 64
for i in range(steps):
    t0 = time.perf_counter()
    out = bb(img)
    t1 = time.perf_counter()
    vit_feat = out['last_feat'][:, 1:]
    res = head.seed_pseudo_gt(None, None, None, None, None, vit_feat=vit_feat.unflatten(1, (hp, wp)).permute(0, 3, 1, 2),
                              point_cls=out['outputs_class'], point_reg=out['outputs_coord'], attns=out['attns'], gt_points=gt_points,
                              gt_points_labels=labels, return_mask=False, pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10,
                              corr_size=21, obj_tau=0.85, pos_inds=pos_inds, gt_index=gt_index)
    t2 = time.perf_counter()
    ev[i + 1].record()
    rows.append([t1 - t0, t2 - t1, 0.0, torch.cuda.memory_reserved() / 2**20])
torch.cuda.synchronize()
for i in range(steps):
    rows[i][2] = ev[i].elapsed_time(ev[i + 1])
t = torch.tensor([r[:3] for r in rows]) * torch.tensor([1e3, 1e3, 1.0])
med = t.median(0).values
print('median ms: host backbone %.2f, host head %.2f, device step %.2f;  mean device step %.2f, max %.2f' % (med[0], med[1], med[2], t[:, 2].mean(), t[:, 2].max()))
slow = (t[:, 2] > 1.4 * med[2]).nonzero().flatten().tolist()
print('slow steps (> 1.4 x median device time):', slow)
for i in slow[:12]:
    for j in range(max(0, i - 1), min(steps, i + 2)):
        print('   step %3d: host backbone %6.2f ms, host head %6.2f ms, device %6.2f ms, reserved %.0f MiB%s' % (j, t[j, 0], t[j, 1], t[j, 2], rows[j][3], '   <-- slow' if j == i else ''))

# ---- which objects of a step live in reference cycles (only the cyclic GC can free them -> delayed frees -> pool growth)?
gc.enable(); gc.collect()
gc.disable()
gc.set_debug(gc.DEBUG_SAVEALL)
bench.one_step(bb, head, img, inputs, False)
torch.cuda.synchronize()
n = gc.collect()
import collections
kinds = collections.Counter(type(o).__name__ for o in gc.garbage)
tens = [o for o in gc.garbage if isinstance(o, torch.Tensor)]
print('objects in reference cycles after ONE step:', n, dict(kinds.most_common(8)), '; tensors among them:', len(tens),
      'device MiB held: %.2f' % (sum(t.numel() * t.element_size() for t in tens if t.is_cuda) / 2**20))
for o in gc.garbage:
    if isinstance(o, dict) and any(isinstance(v, torch.Tensor) for v in o.values()):
        print('   dict in a cycle with tensor values, keys:', list(o.keys())[:12])
gc.set_debug(0); gc.garbage.clear(); gc.enable()

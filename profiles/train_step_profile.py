"""Kernel table of ONE training step (bench.py --mode train, cfg2, one GPU) under torch.profiler: where the 150 ms go.
usage: python profiles/train_step_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from torch.profiler import ProfilerActivity, profile
import bench

cfg = dict(bench.CONFIGS['cfg2']); cfg['cuda_graph'] = False
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
bb, head = bench.build_models(cfg, dev)
bb.train()
opt = torch.optim.AdamW(bb.parameters(), lr=1e-5, fused=True)
img = bench.make_inputs(cfg, 0)[0].to(dev)

def step():
    out = bb(img)
    loss = (out['last_feat'].float().pow(2).mean() + out['outputs_class'].float().pow(2).mean()) * 1024.0
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); out = bb(img); e1.record()
loss = (out['last_feat'].float().pow(2).mean() + out['outputs_class'].float().pow(2).mean()) * 1024.0
opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); e2.record()
torch.cuda.synchronize()
print('forward (autograd, head-mean maps incl.) %.1f ms, backward + AdamW %.1f ms' % (e0.elapsed_time(e1), e1.elapsed_time(e2)))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if 'cuda' in str(e.device_type).lower() and e.name:
        n = e.name.replace('(anonymous namespace)::', '').replace('void ', '')
        n = n.split('<')[0].split('(')[0][:60]
        agg[n][0] += 1; agg[n][1] += e.device_time / 1e3 if hasattr(e, 'device_time') else e.cuda_time / 1e3
tot = sum(v[1] for v in agg.values())
print('%d kernels / copies, %.1f ms of device time' % (sum(v[0] for v in agg.values()), tot))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print('%9.3f ms %5d  %5.1f%%  %s' % (t, n, 100 * t / tot, k))

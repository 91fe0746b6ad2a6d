"""One bench step under torch.profiler (CUPTI): GPU time by kernel name, idle gaps between kernels (host syncs / launch
latency) with the kernels on both sides.  Run on the GPU box: python profiles/step_timeline.py [e2e]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import ProfilerActivity, profile

cfg = dict(bench.WORKLOAD)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
bb, head = bench.build_models(cfg, dev)
inputs = bench.make_inputs(cfg, 0)
img_dev = inputs[0].to(dev)
mask = len(sys.argv) > 1 and sys.argv[1] == 'e2e'
for _ in range(3):
    bench.one_step(bb, head, img_dev, inputs, mask)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        bench.one_step(bb, head, img_dev, inputs, mask)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if 'cuda' in str(e.device_type).lower()]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
print(f'2 steps: span {(t1 - t0) / 1e3:.2f} ms, {len(ev)} device activities')
agg = {}
busy = 0.0
for e in ev:
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
    busy += e.time_range.end - e.time_range.start
print(f'sum of activity durations {busy / 1e3:.2f} ms')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{us / 2e3:8.3f} ms/step {n // 2:5d}  {k}')
gaps = []
end = ev[0].time_range.end
prev = ev[0]
for e in ev[1:]:
    g = e.time_range.start - end
    if g > 0:
        gaps.append((g, prev.name[:50], e.name[:50]))
    if e.time_range.end > end:
        end = e.time_range.end
        prev = e
tot_gap = sum(g for g, _, _ in gaps)
print(f'idle between activities: {tot_gap / 2e3:.3f} ms/step in {len(gaps) // 2} gaps/step')
for g, a, b in sorted(gaps, key=lambda x: -x[0])[:40]:
    print(f'{g:8.1f} us   after {a:50s} before {b}')

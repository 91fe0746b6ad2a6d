"""Host-side activity inside the largest GPU idle gap of a step (torch.profiler, CPU + CUDA).  python profiles/step_gap_detail.py"""
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
for _ in range(3):
    bench.one_step(bb, head, img_dev, inputs, False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        bench.one_step(bb, head, img_dev, inputs, False)
    torch.cuda.synchronize()
evs = list(prof.events())
gpu = sorted([e for e in evs if 'cuda' in str(e.device_type).lower()], key=lambda e: e.time_range.start)
cpu = sorted([e for e in evs if 'cuda' not in str(e.device_type).lower()], key=lambda e: e.time_range.start)
gaps = []
end, prev = gpu[0].time_range.end, gpu[0]
for e in gpu[1:]:
    if e.time_range.start - end > 300:
        gaps.append((end, e.time_range.start, prev.name[:40], e.name[:40]))
    if e.time_range.end > end:
        end, prev = e.time_range.end, e
for a, b, pn, nn in gaps:
    print(f'=== GPU idle {b - a:.0f} us after {pn} before {nn}')
    inside = [e for e in cpu if e.time_range.end > a - 200 and e.time_range.start < b and (e.time_range.end - e.time_range.start) > 30]
    for e in inside[:60]:
        print(f'   +{e.time_range.start - a:9.0f} us  dur {e.time_range.end - e.time_range.start:8.0f}  {e.name[:90]}')

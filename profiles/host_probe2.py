"""Per-step wall time with the backbone as a CUDA graph, timers on / off.  python profiles/host_probe2.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from attentionshift_b200 import ops

cfg = dict(bench.WORKLOAD)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
bb, head = bench.build_models(cfg, dev)
inputs = bench.make_inputs(cfg, 0)
img_dev = inputs[0].to(dev)

def seg():
    s = torch.cuda.memory_stats()
    return s['segment.all.allocated'], s['segment.all.freed'], round(s['reserved_bytes.all.current'] / 2**30, 2)

for timers in (True, False):
    if timers:
        ops.TIMERS.enable()
    else:
        ops.TIMERS.disable()
    for i in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops.TIMERS.begin_step() if timers else None
        t_a = time.perf_counter()
        out = bb(img_dev)
        t_b = time.perf_counter()
        bench.one_step  # noqa
        res = bench.one_step(bb, head, img_dev, inputs, False) if False else None
        torch.cuda.synchronize()
        t_c = time.perf_counter()
        r = bench.one_step(bb, head, img_dev, inputs, False)
        torch.cuda.synchronize()
        t_d = time.perf_counter()
        print(f'timers={timers} step {i}: backbone host {1e3*(t_b-t_a):6.2f} ms, backbone total {1e3*(t_c-t_a):6.2f} ms, full step {1e3*(t_d-t_c):6.2f} ms  {seg()}', flush=True)

"""Bench-like loop (no per-step sync) with host timestamps and per-step CUDA events: where do slow steps come from?"""
import os, sys, time, gc
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
ops.TIMERS.enable()
for _ in range(3):
    ops.TIMERS.begin_step()
    bench.one_step(bb, head, img_dev, inputs, False)
torch.cuda.synchronize()
for trial in range(3):
    if trial == 2:
        gc.disable()
    n = 30
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ts = []
    gcs = []
    ev[0].record()
    for i in range(n):
        t0 = time.perf_counter()
        g0 = gc.get_stats()[2]['collections']
        ops.TIMERS.begin_step()
        bench.one_step(bb, head, img_dev, inputs, False)
        ev[i + 1].record()
        ts.append(time.perf_counter() - t0)
        gcs.append(gc.get_stats()[2]['collections'] - g0)
    torch.cuda.synchronize()
    gpu = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print(f'trial {trial} (gc {"off" if trial == 2 else "on"}): mean host {1e3 * sum(ts) / n:.2f} ms, mean gpu interval {sum(gpu) / n:.2f} ms')
    print('  host ms :', ' '.join(f'{1e3 * t:.0f}' for t in ts))
    print('  gpu  ms :', ' '.join(f'{g:.0f}' for g in gpu))
    print('  gen2 gc :', ' '.join(str(g) for g in gcs))

"""Host (python + driver) time to ENQUEUE one step vs the GPU time of the step.  python profiles/host_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

cfg = dict(bench.WORKLOAD)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
bb, head = bench.build_models(cfg, dev)
inputs = bench.make_inputs(cfg, 0)
img_dev = inputs[0].to(dev)
for _ in range(3):
    bench.one_step(bb, head, img_dev, inputs, False)
torch.cuda.synchronize()
print('cpu count', os.cpu_count())
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = bb(img_dev)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'backbone: host enqueue {1e3 * (t1 - t0):.2f} ms, until GPU done {1e3 * (t2 - t0):.2f} ms')
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    bench.one_step(bb, head, img_dev, inputs, False)
torch.cuda.synchronize()
print(f'full step wall {1e3 * (time.perf_counter() - t0) / 5:.2f} ms')
def seg():
    s = torch.cuda.memory_stats()
    return s['segment.all.allocated'], s['segment.all.freed'], s['reserved_bytes.all.current'] / 2**30, s['num_alloc_retries']
print('segments allocated/freed, reserved GiB, retries:', seg())
for i in range(12):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bench.one_step(bb, head, img_dev, inputs, False)
    torch.cuda.synchronize()
    print(f'step {i}: {1e3 * (time.perf_counter() - t0):7.2f} ms   {seg()}')

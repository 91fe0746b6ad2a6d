"""Where does the end-to-end step spend its time?  python profiles/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

cfg = dict(bench.WORKLOAD)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
bb, head = bench.build_models(cfg, dev)
inputs = bench.make_inputs(cfg, 0)
img_host = inputs[0]
img_dev = img_host.to(dev)
for _ in range(3):
    bench.one_step(bb, head, img_dev, inputs, True)
torch.cuda.synchronize()


def wall(fn, n=5):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


print('step, resident image, no mask D2H : %.2f ms' % wall(lambda: bench.one_step(bb, head, img_dev, inputs, False)))
print('step, resident image, mask D2H    : %.2f ms' % wall(lambda: bench.one_step(bb, head, img_dev, inputs, True)))
buf = torch.empty_like(img_dev)
print('image H2D (100 MB pinned) alone   : %.2f ms' % wall(lambda: buf.copy_(img_host, non_blocking=True)))
m = torch.empty(24, 1024, 1024, dtype=torch.uint8, device=dev)
def d2h():
    h = torch.empty(m.shape, dtype=torch.uint8, pin_memory=True)
    h.copy_(m, non_blocking=True)
    torch.cuda.current_stream().synchronize()
print('mask D2H (25 MB, pinned alloc)    : %.2f ms' % wall(d2h))
t = time.perf_counter(); h = torch.empty(m.shape, dtype=torch.uint8, pin_memory=True); print('pinned alloc 25 MB first: %.2f ms' % ((time.perf_counter() - t) * 1e3))
def both():
    buf.copy_(img_host, non_blocking=True)
    bench.one_step(bb, head, buf, inputs, True)
print('H2D on the same stream + step     : %.2f ms' % wall(both))
cs = torch.cuda.Stream()
def overl():
    with torch.cuda.stream(cs):
        buf.copy_(img_host, non_blocking=True)
    bench.one_step(bb, head, img_dev, inputs, True)
print('H2D on a side stream || step      : %.2f ms' % wall(overl))
print('step on buf (no copy)             : %.2f ms' % wall(lambda: bench.one_step(bb, head, buf, inputs, True)))
def both_sync():
    buf.copy_(img_host, non_blocking=True)
    torch.cuda.synchronize()
    bench.one_step(bb, head, buf, inputs, True)
print('H2D, sync, step on buf            : %.2f ms' % wall(both_sync))
def both_nomask():
    buf.copy_(img_host, non_blocking=True)
    bench.one_step(bb, head, buf, inputs, False)
print('H2D same stream + step (no mask)  : %.2f ms' % wall(both_nomask))
img_host2 = torch.randn(8, 3, 1024, 1024).pin_memory()
def both2():
    buf.copy_(img_host2, non_blocking=True)
    bench.one_step(bb, head, buf, inputs, True)
print('H2D of OTHER random image + step  : %.2f ms' % wall(both2))
print('step on img_dev again             : %.2f ms' % wall(lambda: bench.one_step(bb, head, img_dev, inputs, True)))
from attentionshift_b200 import ops
ops.TIMERS.enable()
both2(); torch.cuda.synchronize()
fam = ops.TIMERS.summary(); ops.TIMERS.disable()
print({k: round(v['ms'], 2) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])[:12]})

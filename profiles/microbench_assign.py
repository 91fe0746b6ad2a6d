"""Point matching of one cfg2 batch (8 images x 100 point tokens x 3 GTs; and 8 x 100 x 20): the reference's route (D2H of
the predictions + scipy per image on the host) against the device solver (as_hungarian_points)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from attentionshift_b200 import assigner as A

dev = 'cuda:0'
g = torch.Generator().manual_seed(0)
for n_gt in (3, 20):
    B, P = 8, 100
    reg = torch.rand(B, P, 2, generator=g).to(dev)
    cls = torch.randn(B, P, 20, generator=g).to(dev)
    pts = [torch.rand(n_gt, 2, generator=g) * 1024 for _ in range(B)]
    lab = [torch.randint(0, 20, (n_gt,), generator=g) for _ in range(B)]
    wh = [(1024, 1024)] * B

    def host():
        r, c = reg.float().cpu(), cls.float().cpu()
        return [A.hungarian_point_assign(r[i], c[i], pts[i], lab[i], wh[i]) for i in range(B)]

    def device():
        return A.hungarian_point_assign_device(reg, cls, pts, lab, wh)

    for name, fn in (('host (scipy)', host), ('device', device)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        t_issue = (time.perf_counter() - t0) / 20
        torch.cuda.synchronize()
        print(f'n_gt {n_gt:2d}  {name:13s} host time / call {t_issue * 1e3:7.3f} ms   device span / call {e0.elapsed_time(e1) / 20:7.3f} ms')

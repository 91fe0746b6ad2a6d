"""Head-mean pass at the cfg2 shapes: CUDA-event time of the three production modes (env AS_HEADMEAN_ORDER=image selects the
per-image tile order).  usage: python profiles/microbench_headmean.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import ops
dev = 'cuda'
B, T, C, H = 8, 4197, 768, 12
Tpad = (T + 127) // 128 * 128
x = (torch.randn(B * T, C, device=dev) * 0.5).half()
w = (torch.randn(3 * C, C, device=dev) * 0.05).half()
q, k, vt = ops.qkv_proj(x, w, torch.zeros(3 * C, device=dev), B, T, H, Tpad)
o, m, l = ops.mhsa_fwd(q, k, vt, T)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

def t(fn, n=8):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()                                   # L2 flush between timed launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n

print('order', os.environ.get('AS_HEADMEAN_ORDER', 'contiguous'))
print('full maps          %.3f ms' % t(lambda: ops.attn_headmean(q, k, m, l, T)))
print('roll-out operands  %.3f ms' % t(lambda: ops.attn_headmean(q, k, m, l, T, want_map=False)))
print('last-layer rows    %.3f ms' % t(lambda: ops.attn_headmean(q, k, m, l, T, want_transposed=False, row0=T - 100)))

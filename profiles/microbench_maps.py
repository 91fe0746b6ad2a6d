"""Map-stage kernels at the cfg2 shapes: cosine maps (refinement: 8 groups x 7 rows; parts: 24 x 16).  python profiles/microbench_maps.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import attention_shift as AS
dev = 'cuda'
feats = torch.randn(8, 4096, 768, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
g8 = torch.arange(8, dtype=torch.int32, device=dev)
p7 = torch.randn(8, 7, 768, device=dev)
o24 = torch.arange(8, dtype=torch.int32, device=dev).repeat_interleave(3)
p16 = torch.randn(24, 16, 768, device=dev)
print('cosine_maps refine (8 x 7)  : %.3f ms' % timeit(lambda: AS.cosine_maps(feats, g8, p7)))
print('cosine_maps parts (24 x 16) : %.3f ms' % timeit(lambda: AS.cosine_maps(feats, o24, p16)))
ref = torch.nn.functional.cosine_similarity(p7[:, :, None], feats[:, None], dim=-1)
print('max abs err vs torch        : %.2e' % (AS.cosine_maps(feats, g8, p7) - ref).abs().max().item())
w = torch.rand(8, 7, 4096, device=dev)
w[w < 0.9] = 0
wsum = w.sum(-1).reshape(-1).contiguous()
L = AS._l.load()
nb = L.as_weighted_centroid_workspace(8, 7, 4096, 768)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
out = torch.empty(8, 7, 768, device=dev)
def wc():
    AS._l.check(L.as_weighted_centroid(AS._p(feats), feats.stride(0), AS._p(g8), AS._p(w), AS._p(wsum), 8, 7, 4096, 768, AS._p(out), AS._p(ws), nb, AS._sp()), 'wc')
print('weighted_centroid (8 x 7)   : %.3f ms' % timeit(wc))
refc = torch.einsum('gsn,gnc->gsc', w, feats) / wsum.view(8, 7, 1)
print('max rel err vs torch        : %.2e' % ((out - refc).abs().max() / refc.abs().max()).item())

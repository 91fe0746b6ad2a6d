"""Attention-shift loop alone at the cfg2 shapes (8 images x 4096 tokens x 768 channels, 3 instances x 16 seeds, 5 iterations):
CUDA-event time per variant, and the per-phase nanoseconds of the persistent kernel (as_mean_shift_fused_debug).
Run on the GPU box: python profiles/microbench_meanshift.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import ops, lib
from attentionshift_b200.synthetic import structured_scene

dev = 'cuda'
n_img, hp, C, n_obj, S, iters = 8, 64, 768, 3, 16, 5
N = hp * hp
scenes = [structured_scene(hp, hp, C, n_obj, seed=10 + i, noise=0.4) for i in range(n_img)]
feats = torch.stack([s['vit_feat'].permute(1, 2, 0).reshape(N, C) for s in scenes]).contiguous().to(dev)
obj_img = torch.arange(n_img, dtype=torch.int32).repeat_interleave(n_obj).to(dev)
rois = torch.cat([s['rois'] for s in scenes]).to(dev)
maps = torch.cat([torch.stack([((s['labels'] == 2 * j + 1) | (s['labels'] == 2 * j + 2)).float() for j in range(n_obj)])
                  for s in scenes]).reshape(-1, N).to(dev)
_, proto0 = ops.grid_seeds(maps, feats, obj_img, rois, hp, S)
npi = [n_obj] * n_img


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


b_alg = ((iters + 1) * N * C * 4 + n_obj * S * N * 4 + 2 * n_obj * S * C * 4) * n_img
for impl in ('v2', 'fused', 'tc', 'fp32'):
    ms = timeit(lambda: ops.mean_shift(proto0, feats, obj_img, rois, hp, hp, iters, n_per_img=npi, impl=impl))
    print(f'{impl:6s} {ms:8.3f} ms   B_alg {b_alg / 1e6:.1f} MB -> {b_alg / ms / 1e6:8.1f} GB/s')

L = lib.load()
# ---- per-phase nanoseconds of the second-generation persistent kernel (256 CTAs of 128 tokens, two per SM)
G2 = ((N + 63) // 64 + 1) // 2
dbg2 = torch.zeros(n_img * G2 * 16, dtype=torch.int64, device=dev)
L.as_mean_shift_v2_debug(lib.ptr(dbg2))
ops.mean_shift(proto0, feats, obj_img, rois, hp, hp, iters, n_per_img=npi, impl='v2')
torch.cuda.synchronize()
L.as_mean_shift_v2_debug(None)
d2 = dbg2.view(n_img * G2, 16).double() / 1e3
names2 = ['0 seeds p^ + barrier', '1 affinity epilogue + column stats', '2 barrier 1', '3 statistics / partial Z', '4 barrier 2',
          '5 assign + weight tiles', '6 update (tcgen05) + epilogue', '7 barrier 3', '8 reduce', '9 barrier 4', '10 affinity main loop (TMA + tcgen05)']
print('v2 phase                           mean us   max us   (worker thread 0 of each CTA, summed over the call)')
for k, nme in enumerate(names2):
    print(f'{nme:34s} {d2[:, k].mean().item():8.1f} {d2[:, k].max().item():8.1f}')
print('v2 total', d2.sum(1).mean().item())

dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
L.as_mean_shift_fused_debug(lib.ptr(dbg))
ops.mean_shift(proto0, feats, obj_img, rois, hp, hp, iters, n_per_img=npi, impl='fused')
torch.cuda.synchronize()
L.as_mean_shift_fused_debug(None)
d = dbg.view(148, 16)[:128].double() / 1e3
names = ['0 seeds p^ + barrier', '1 affinity epilogue + column stats', '2 barrier 1', '3 statistics / partial Z', '4 barrier 2', '5 assign + weight tiles',
         '6 update (tcgen05) + epilogue', '7 barrier 3', '8 reduce', '9 barrier 4', '10 affinity main loop (TMA + tcgen05)']
print('phase                              mean us   max us   (worker thread 0 of each CTA, summed over the call)')
for k, nme in enumerate(names):
    print(f'{nme:34s} {d[:, k].mean().item():8.1f} {d[:, k].max().item():8.1f}')
print('total', d.sum(1).mean().item())

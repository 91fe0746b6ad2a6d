"""Per-kernel CUDA-event timings at the cfg2 shapes (B=8, T=4197, ViT-B).  Run on the GPU box: python profiles/microbench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import ops

dev = 'cuda'
B, T, C, H = 8, 4197, 768, 12
M = B * T
Tpad = (T + 127) // 128 * 128


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


x768 = (torch.randn(M, C, device=dev) * 0.5).half()
x3072 = (torch.randn(M, 4 * C, device=dev) * 0.5).half()
w_qkv = (torch.randn(3 * C, C, device=dev) * 0.05).half()
w_proj = (torch.randn(C, C, device=dev) * 0.05).half()
w_fc1 = (torch.randn(4 * C, C, device=dev) * 0.05).half()
w_fc2 = (torch.randn(C, 4 * C, device=dev) * 0.05).half()
b768, b2304, b3072 = torch.zeros(C, device=dev), torch.zeros(3 * C, device=dev), torch.zeros(4 * C, device=dev)
resid = torch.randn(M, C, device=dev)
xf = torch.randn(M, C, device=dev)
g = torch.ones(C, device=dev)

res = {}
res['qkv  (N=2304,K=768, head-split f16)'] = (timeit(lambda: ops.qkv_proj(x768, w_qkv, b2304, B, T, H, Tpad)), 2 * M * C * 3 * C)
res['proj (N=768,K=768, f32 resid)'] = (timeit(lambda: ops.linear_f16(x768, w_proj, b768, ops.EPI_RESID_F32, resid=resid)), 2 * M * C * C)
res['fc1  (N=3072,K=768, GELU f16)'] = (timeit(lambda: ops.linear_f16(x768, w_fc1, b3072, ops.EPI_GELU_F16)), 2 * M * C * 4 * C)
res['fc1* (N=3072,K=768, plain f16)'] = (timeit(lambda: ops.linear_f16(x768, w_fc1, b3072, ops.EPI_F16)), 2 * M * C * 4 * C)
res['fc2  (N=768,K=3072, f32 resid)'] = (timeit(lambda: ops.linear_f16(x3072, w_fc2, b768, ops.EPI_RESID_F32, resid=resid)), 2 * M * C * 4 * C)
res['fc2* (N=768,K=3072, f32 plain)'] = (timeit(lambda: ops.linear_f16(x3072, w_fc2, b768, ops.EPI_F32)), 2 * M * C * 4 * C)
q, k, vt = ops.qkv_proj(x768, w_qkv, b2304, B, T, H, Tpad)
res['mhsa_fwd'] = (timeit(lambda: ops.mhsa_fwd(q, k, vt, T)), 4 * T * T * C * B)
o, m, l = ops.mhsa_fwd(q, k, vt, T)
res['headmean (+transposed)'] = (timeit(lambda: ops.attn_headmean(q, k, m, l, T)), 2 * T * T * C * B)
res['headmean (no transposed)'] = (timeit(lambda: ops.attn_headmean(q, k, m, l, T, want_transposed=False)), 2 * T * T * C * B)
res['headmean v1 (+transposed)'] = (timeit(lambda: ops.attn_headmean(q, k, m, l, T, slices=1)), 2 * T * T * C * B)
res['layernorm'] = (timeit(lambda: ops.layernorm_f16(xf, g, b768)), 0)
mm = torch.matmul
res['torch.matmul f16 qkv-shape (cuBLAS)'] = (timeit(lambda: mm(x768, w_qkv.t())), 2 * M * C * 3 * C)
res['torch.matmul f16 fc2-shape (cuBLAS)'] = (timeit(lambda: mm(x3072, w_fc2.t())), 2 * M * C * 4 * C)
for k_, (ms, fl) in res.items():
    print(f'{k_:42s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s')

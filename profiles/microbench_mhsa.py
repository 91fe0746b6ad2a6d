"""as_mhsa_fwd schedules side by side at the cfg2 shape (B=8, T=4197, 12 heads): parity against a torch fp32 reference of
VT:79-83 (output, and the (m, l) statistics through the head-mean pass) and CUDA-event time per variant.
Run on the GPU box: python profiles/microbench_mhsa.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import ops, lib

dev = 'cuda'
L = lib.load()


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


def ref(q, k, vt, T):
    qf, kf, vf = q.float(), k.float(), vt[..., :T].float().transpose(-1, -2)
    attn = ((qf @ kf.transpose(-2, -1)) * 0.125).softmax(-1)
    return (attn @ vf).transpose(1, 2).reshape(q.shape[0], T, -1), attn.mean(1)


def make(B, H, T, std, seed=0):
    torch.manual_seed(seed)
    Tpad = (T + 127) // 128 * 128
    q = (torch.randn(B, H, T, 64, device=dev) * std).half()
    k = (torch.randn(B, H, T, 64, device=dev) * std).half()
    vt = torch.zeros(B, H, 64, Tpad, device=dev, dtype=torch.float16)
    vt[..., :T] = torch.randn(B, H, 64, T, device=dev).half()
    return q, k, vt


variants = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
for std, T in ((0.6, 4197), (1.5, 4197), (2.5, 1125), (1.0, 4096)):
    q, k, vt = make(1, 3, T, std)
    ro, rmean = ref(q, k, vt, T)
    for v in variants:
        assert L.as_mhsa_set_variant(v) == 0
        o, m, l = ops.mhsa_fwd(q, k, vt, T)
        mean, _ = ops.attn_headmean(q, k, m, l, T, want_transposed=False)
        torch.cuda.synchronize()
        eo = (o.float() - ro).abs().max().item() / ro.abs().max().item()
        em = ((mean - rmean).abs() / (rmean.abs() + 1e-6 * rmean.abs().max())).max().item()
        print(f'std {std} T {T} variant {v}: max |o - ref| / max|ref| = {eo:.2e}   head-mean max rel err = {em:.2e}', flush=True)

B, T, H = 8, 4197, 12
for std in (0.6, 1.5):
    q, k, vt = make(B, H, T, std, seed=1)
    for v in variants:
        L.as_mhsa_set_variant(v)
        ms = timeit(lambda: ops.mhsa_fwd(q, k, vt, T))
        print(f'std {std} variant {v}: {ms:7.3f} ms   {4 * T * T * 768 * B / ms / 1e9:7.1f} TFLOP/s', flush=True)
L.as_mhsa_set_variant(4)

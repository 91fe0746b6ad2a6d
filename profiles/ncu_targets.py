"""Tiny driver for ncu: runs each hot kernel a couple of times at the cfg2 shapes (see profiles/run_ncu2.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200 import ops
dev = 'cuda'
B, T, C, H = 8, 4197, 768, 12
M = B * T
Tpad = (T + 127) // 128 * 128
x768 = (torch.randn(M, C, device=dev) * 0.5).half()
w_qkv = (torch.randn(3 * C, C, device=dev) * 0.05).half()
w_proj = (torch.randn(C, C, device=dev) * 0.05).half()
w_fc1 = (torch.randn(4 * C, C, device=dev) * 0.05).half()
b768, b2304, b3072 = torch.zeros(C, device=dev), torch.zeros(3 * C, device=dev), torch.zeros(4 * C, device=dev)
resid = torch.randn(M, C, device=dev)
which = sys.argv[1]
for _ in range(2):
    if which == 'proj':
        ops.linear_f16(x768, w_proj, b768, ops.EPI_RESID_F32, resid=resid)
    elif which == 'fc1':
        ops.linear_f16(x768, w_fc1, b3072, ops.EPI_GELU_F16)
    elif which == 'qkv':
        ops.qkv_proj(x768, w_qkv, b2304, B, T, H, Tpad)
    elif which in ('mhsa', 'headmean', 'headmean_lean', 'headmean_rows'):
        q, k, vt = ops.qkv_proj(x768, w_qkv, b2304, B, T, H, Tpad)
        o, m, l = ops.mhsa_fwd(q, k, vt, T)
        if which == 'headmean':            # reference-shaped production: fp32 map + transposed pair + row sums
            ops.attn_headmean(q, k, m, l, T)
        elif which == 'headmean_lean':     # roll-out operand only (every layer but the last, attn_format='rollout')
            ops.attn_headmean(q, k, m, l, T, want_map=False)
        elif which == 'headmean_rows':     # the last layer: point-token rows only
            ops.attn_headmean(q, k, m, l, T, want_transposed=False, row0=T - 100)
torch.cuda.synchronize()
if which == 'mhsa_bwd':        # attention backward at the cfg2 shapes (training path)
    from attentionshift_b200 import training
    q, k, vt = ops.qkv_proj(x768, w_qkv, b2304, B, T, H, Tpad)
    o, m, l = ops.mhsa_fwd(q, k, vt, T)
    g = torch.randn(B, T, C, device=dev)
    for _ in range(2):
        training.mhsa_bwd(q, k, vt, o, g, m, l, T)
    torch.cuda.synchronize()

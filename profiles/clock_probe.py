"""SM clock and power while one kernel runs back to back (NVML, 5 ms period): is a kernel running into the power cap?
python profiles/clock_probe.py"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch
from attentionshift_b200 import ops

pynvml.nvmlInit()
hd = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = 'cuda'
B, T, H = 8, 4197, 12
Tpad = (T + 127) // 128 * 128
torch.manual_seed(0)
q = (torch.randn(B, H, T, 64, device=dev) * 0.6).half()
k = (torch.randn(B, H, T, 64, device=dev) * 0.6).half()
vt = torch.zeros(B, H, 64, Tpad, device=dev, dtype=torch.float16)
vt[..., :T] = torch.randn(B, H, 64, T, device=dev).half()
o, m, l = ops.mhsa_fwd(q, k, vt, T)
x = (torch.randn(B * T, 3072, device=dev) * 0.5).half()
w = (torch.randn(768, 3072, device=dev) * 0.05).half()


def run(name, fn, flops, n):
    rows, stop = [], [False]

    def sample():
        while not stop[0]:
            rows.append((pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1e3))
            time.sleep(0.005)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    th = threading.Thread(target=sample)
    th.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    stop[0] = True
    th.join()
    ms = s.elapsed_time(e) / n
    clk = sorted(r[0] for r in rows)
    pw = sorted(r[1] for r in rows)
    print(f'{name:28s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TF/s   SM clock min/median/max {clk[0]}/{clk[len(clk) // 2]}/{clk[-1]} MHz   power median/max {pw[len(pw) // 2]:.0f}/{pw[-1]:.0f} W   ({len(rows)} samples)')


run('mhsa_fwd x10', lambda: ops.mhsa_fwd(q, k, vt, T), 4 * T * T * 768 * B, 10)
run('mhsa_fwd x400', lambda: ops.mhsa_fwd(q, k, vt, T), 4 * T * T * 768 * B, 400)
run('headmean x300', lambda: ops.attn_headmean(q, k, m, l, T), 2 * T * T * 768 * B, 300)
run('fc2-shape GEMM x1000', lambda: ops.linear_f16(x, w, None, ops.EPI_F32), 2 * B * T * 768 * 3072, 1000)
run('cuBLAS fc2-shape x1000', lambda: torch.matmul(x, w.t()), 2 * B * T * 768 * 3072, 1000)

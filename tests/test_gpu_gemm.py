"""tcgen05 linear kernel vs a plain torch fp32 reference of the same op (fp16 operands, fp32 accumulate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b):
    return x.float() @ w.float().t() + b


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (300, 768, 768), (1000, 2304, 768), (4197, 3072, 768), (257, 768, 3072), (200, 384, 384)])
def test_linear_modes(M, N, K):
    from attentionshift_b200 import ops
    torch.manual_seed(0)
    x = (torch.randn(M, K, device='cuda') * 0.5).half()
    w = (torch.randn(N, K, device='cuda') * 0.05).half()
    b = torch.randn(N, device='cuda')
    ref = _ref(x, w, b)
    tol = dict(rtol=2e-3, atol=2e-3)      # fp16 output rounding dominates
    y = ops.linear_f16(x, w, b, ops.EPI_F32)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)   # fp32 out: only accumulation-order noise
    y = ops.linear_f16(x, w, b, ops.EPI_F16)
    torch.testing.assert_close(y.float(), ref, **tol)
    y = ops.linear_f16(x, w, b, ops.EPI_GELU_F16)
    torch.testing.assert_close(y.float(), torch.nn.functional.gelu(ref), **tol)
    r = torch.randn(M, N, device='cuda')
    y = ops.linear_f16(x, w, b, ops.EPI_RESID_F32, resid=r)
    torch.testing.assert_close(y, ref + r, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('B,T,heads', [(2, 297, 12), (1, 4197, 12), (3, 117, 2)])
def test_qkv_split(B, T, heads):
    from attentionshift_b200 import ops
    torch.manual_seed(1)
    C = heads * 64
    Tpad = (T + 127) // 128 * 128
    x = (torch.randn(B * T, C, device='cuda') * 0.5).half()
    w = (torch.randn(3 * C, C, device='cuda') * 0.05).half()
    b = torch.randn(3 * C, device='cuda') * 0.1
    q, k, vt = ops.qkv_proj(x, w, b, B, T, heads, Tpad)
    ref = _ref(x, w, b).reshape(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)   # VT:76
    tol = dict(rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(q.float(), ref[0], **tol)
    torch.testing.assert_close(k.float(), ref[1], **tol)
    torch.testing.assert_close(vt[..., :T].float(), ref[2].transpose(-1, -2), **tol)
    assert vt[..., T:].abs().max().item() == 0

"""tcgen05 flash attention + head-mean pass vs a torch fp32 reference of VT:74-86 on the same fp16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, vt, T):
    qf, kf, vf = q.float(), k.float(), vt[..., :T].float().transpose(-1, -2)
    attn = ((qf @ kf.transpose(-2, -1)) * 0.125).softmax(-1)          # VT:79-80
    o = (attn @ vf).transpose(1, 2).reshape(q.shape[0], T, -1)         # VT:83
    return o, attn


@pytest.mark.parametrize('variant', [4, 3, 2, 1])          # attention schedules, see as_mhsa_set_variant
@pytest.mark.parametrize('slices', [4, 1])              # head-mean schedules: persistent / one CTA per tile
@pytest.mark.parametrize('B,heads,T', [(1, 2, 128), (2, 3, 297), (1, 12, 1125), (1, 2, 4197), (3, 16, 640)])
def test_mhsa_and_headmean(B, heads, T, slices, variant):
    from attentionshift_b200 import lib, ops
    assert lib.load().as_mhsa_set_variant(variant) == 0
    torch.manual_seed(0)
    Tpad = (T + 127) // 128 * 128
    q = (torch.randn(B, heads, T, 64, device='cuda') * 1.5).half()
    k = (torch.randn(B, heads, T, 64, device='cuda') * 1.5).half()
    vt = torch.zeros(B, heads, 64, Tpad, device='cuda', dtype=torch.float16)
    vt[..., :T] = torch.randn(B, heads, 64, T, device='cuda').half()
    o, m, l = ops.mhsa_fwd(q, k, vt, T)
    lib.load().as_mhsa_set_variant(4)
    ro, rattn = _ref(q, k, vt, T)
    # P is rounded to fp16 before P@V and O is stored in fp16: ~1e-3 of the output scale
    err = (o.float() - ro).abs().max().item()
    assert err <= 4e-3 * ro.abs().max().item() + 1e-3, err
    mean, part = ops.attn_headmean(q, k, m, l, T, slices=slices)
    rmean = rattn.mean(1)                                               # VTD:236
    torch.testing.assert_close(mean, rmean, rtol=1e-3, atol=1e-6)       # north_star fp32 tolerance
    torch.testing.assert_close(part.sum(-1), rmean.sum(-1), rtol=1e-4, atol=1e-5)
    # transposed split-fp16 copy (roll-out operand): hi + lo reproduces the fp32 map to ~2^-22 of T_SCALE
    hi, lo = mean._as_t16
    rec = (hi.float() + lo.float())[:, :T, :T].transpose(1, 2) / ops.T_SCALE
    torch.testing.assert_close(rec, mean, rtol=2e-6, atol=1e-9)
    assert hi[:, T:].abs().sum().item() == 0 and hi[:, :, T:].abs().sum().item() == 0

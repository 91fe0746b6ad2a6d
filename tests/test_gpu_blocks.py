"""The stand-alone ``Block`` (VT:88-124) at the RoI decoders' shapes (dim 256, 8 heads = head_dim 32, 50 / 197 tokens per RoI:
mae_bbox_head_rec.py, mae_mask_head_pointSup.py) and ``as_mhsa_small`` against torch fp32."""
import pytest
import torch

from oracle import vit as V

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,T,heads,d', [(3, 50, 8, 32), (5, 197, 8, 32), (2, 256, 4, 64), (1, 1, 2, 32), (64, 33, 8, 32)])
def test_mhsa_small_vs_torch(B, T, heads, d):
    from attentionshift_b200 import blocks
    torch.manual_seed(T)
    qkv = torch.randn(B * T, 3 * heads * d, device='cuda').half()
    o = blocks.mhsa_small(qkv, B, T, heads, d)
    q, k, v = qkv.float().view(B, T, 3, heads, d).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1) @ v
    ref = ref.transpose(1, 2).reshape(B * T, heads * d)
    torch.testing.assert_close(o.float(), ref, rtol=2e-3, atol=2e-3)       # fp16 output rounding


@pytest.mark.parametrize('dim,heads,T,B', [(256, 8, 50, 16), (256, 8, 197, 4), (384, 6, 120, 2)])
def test_block_vs_oracle(dim, heads, T, B):
    from attentionshift_b200.blocks import Block
    torch.manual_seed(dim + T)
    blk = Block(dim, heads, mlp_ratio=4., qkv_bias=True, norm_layer=lambda n: torch.nn.LayerNorm(n, eps=1e-6)).cuda().eval()
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.05)
    torch.nn.init.ones_(blk.norm1.weight); torch.nn.init.ones_(blk.norm2.weight)
    x = torch.randn(B, T, dim)
    out = blk(x.cuda())
    sd = {'b.' + k: v.detach().cpu() for k, v in blk.state_dict().items()}
    ref, _ = V.block(x, sd, 'b.', heads)                                   # oracle restatement of VT:109-124 (fp32, CPU)
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < 2e-3

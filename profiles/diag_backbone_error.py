"""Where does the device backbone's deviation from the fp32 oracle come from when the attention is sharpened?
Per layer: (a) accumulated error of the residual stream, (b) error of ONE block fed the oracle's own input (no accumulation),
(c) the same for a torch fp32 block whose GEMM operands are rounded to fp16 (emulation of apex O1 arithmetic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentionshift_b200.registry import build_backbone
from attentionshift_b200.synthetic import vit_state_dict
from oracle import vit as V

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())

def emu_block(x, sd, p, heads):
    """fp32 block with every GEMM operand rounded to fp16 (fp32 accumulate), like the device path / apex O1."""
    import torch.nn.functional as F
    h16 = lambda t: t.half().float()
    B, T, C = x.shape
    xn = h16(F.layer_norm(x, (C,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], 1e-6))
    qkv = h16(xn @ h16(sd[p + 'attn.qkv.weight']).t() + sd[p + 'attn.qkv.bias'])
    qkv = qkv.reshape(B, T, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * (C // heads) ** -0.5).softmax(-1)
    o = h16((h16(attn) @ v).transpose(1, 2).reshape(B, T, C))
    x = x + o @ h16(sd[p + 'attn.proj.weight']).t() + sd[p + 'attn.proj.bias']
    xn = h16(F.layer_norm(x, (C,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-6))
    hid = h16(F.gelu(xn @ h16(sd[p + 'mlp.fc1.weight']).t() + sd[p + 'mlp.fc1.bias']))
    x = x + hid @ h16(sd[p + 'mlp.fc2.weight']).t() + sd[p + 'mlp.fc2.bias']
    return x, attn.mean(1)

embed, heads, depth, img, n_pt = 768, 12, 12, 224, 100
for scale in (1.0, 2.0, 4.0):
    sd = vit_state_dict(embed, depth, heads, img, n_point_tokens=n_pt, seed=5)
    for i in range(depth):
        sd[f'blocks.{i}.attn.qkv.weight'] = sd[f'blocks.{i}.attn.qkv.weight'] * scale
    bb = build_backbone(dict(type='VisionTransformerDet', img_size=img, patch_size=16, embed_dim=embed, depth=depth, num_heads=heads,
                             mlp_ratio=4, qkv_bias=True, with_fpn=False, last_feat=True, return_attention=True, point_tokens_num=n_pt,
                             with_point_head=False, out_indices=[depth - 1]))
    bb.load_state_dict(sd, strict=False)
    bb = bb.cuda().eval()
    x = torch.randn(1, 3, img, img, generator=torch.Generator().manual_seed(105))
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        xo = V.prepare_tokens(x, sd, 16).cuda()               # oracle stream (fp32, on the GPU for speed)
        xd = bb.prepare_tokens(x.cuda())
        B, T, C = xd.shape
        xs = xd.view(B * T, C)
        xe = xo.clone()
        print(f'--- qkv_scale {scale}: tokens rel err {rel(xd, xo):.2e}')
        for i in range(depth):
            p = f'blocks.{i}.'
            xo_in = xo
            xo, ao = V.block(xo, sdc, p, heads)
            xs, ad = bb._block(i, xs, B, T, True)
            x1, a1 = bb._block(i, xo_in.reshape(B * T, C).contiguous(), B, T, True)     # one device block on the oracle's input
            xe_acc, _ = emu_block(xe, sdc, p, heads)
            xe1, ae1 = emu_block(xo_in, sdc, p, heads)
            xe = xe_acc
            amax = float(ao.max())
            print(f'layer {i:2d}: accumulated x {rel(xs.view(B, T, C), xo):.2e} (fp16-emulation {rel(xe, xo):.2e}) | single block x {rel(x1.view(B, T, C), xo):.2e} '
                  f'(emu {rel(xe1, xo):.2e}) attn abs {float((a1 - ao).abs().max()):.2e} (emu {float((ae1 - ao).abs().max()):.2e}) attn max {amax:.3f} |x| {float(xo.abs().max()):.1f}')

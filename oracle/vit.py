"""CPU oracle (torch fp32) for the ViT half of the hot path.

TEST INFRASTRUCTURE ONLY -- never imported by ``attentionshift_b200``.

Functional restatement (parameters come from a state_dict that uses the
reference's names, e.g. ``blocks.3.attn.qkv.weight``) of:
  VT  = models/vision_transformer.py            (Attention VT:62-86, Block VT:88-124,
                                                 pos-embed resize VT:187-207)
  VTD = mmdet/models/backbones/visual_transformer_det.py
                                                (prepare_tokens VTD:192-214, forward VTD:221-275)
Pinned against the reference classes by tests/test_oracle_vs_reference.py and the
``vit_*.pt`` goldens.
"""
import math

import torch
import torch.nn.functional as F


def attention(x, qkv_w, qkv_b, proj_w, proj_b, num_heads):
    """VT:74-86.  x [B,T,C] -> (y [B,T,C], attn [B,h,T,T]); scale = head_dim**-0.5 (VT:67)."""
    b, t, c = x.shape
    d = c // num_heads
    qkv = F.linear(x, qkv_w, qkv_b).reshape(b, t, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(b, t, c)
    return F.linear(y, proj_w, proj_b), attn


def block(x, sd, prefix, num_heads, eps=1e-6):
    """VT:109-124 pre-LN block with init_values=0 (no gamma), GELU MLP (VT:40-59).
    -> (x, head-mean attention [B,T,T]) -- the mean over heads is VTD:236/242."""
    c = x.shape[-1]
    p = lambda n: sd[prefix + n]
    y, attn = attention(F.layer_norm(x, (c,), p('norm1.weight'), p('norm1.bias'), eps),
                        p('attn.qkv.weight'), p('attn.qkv.bias'), p('attn.proj.weight'), p('attn.proj.bias'),
                        num_heads)
    x = x + y
    h = F.layer_norm(x, (c,), p('norm2.weight'), p('norm2.bias'), eps)
    h = F.linear(F.gelu(F.linear(h, p('mlp.fc1.weight'), p('mlp.fc1.bias'))), p('mlp.fc2.weight'), p('mlp.fc2.bias'))
    return x + h, attn.mean(1)


def resized_pos_embed(pos_embed, n_patch_tokens, w, h, patch):
    """VT:187-207 bicubic resize of the patch position table (note the reference's
    +0.1 fudge and its (w, h) naming: w is tensor dim 2, h is dim 3)."""
    n0 = pos_embed.shape[1] - 1
    if n_patch_tokens == n0 and w == h:
        return pos_embed
    dim = pos_embed.shape[-1]
    w0, h0 = w // patch + 0.1, h // patch + 0.1
    g = int(math.sqrt(n0))
    pp = F.interpolate(pos_embed[:, 1:].reshape(1, g, g, dim).permute(0, 3, 1, 2),
                       scale_factor=(w0 / math.sqrt(n0), h0 / math.sqrt(n0)), mode='bicubic')
    assert int(w0) == pp.shape[-2] and int(h0) == pp.shape[-1]
    pp = pp.permute(0, 2, 3, 1).view(1, -1, dim)
    return torch.cat((pos_embed[:, 0].unsqueeze(0), pp), dim=1)


def prepare_tokens(img, sd, patch):
    """VTD:192-214: conv patch embed, cls token, position table, then the point
    tokens (+ their own position table) appended at the END.  -> [B, 1+N+Tp, C]."""
    b, _, w, h = img.shape
    x = F.conv2d(img, sd['patch_embed.proj.weight'], sd['patch_embed.proj.bias'], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd['cls_token'].expand(b, -1, -1), x), dim=1)
    x = x + resized_pos_embed(sd['pos_embed'], x.shape[1] - 1, w, h, patch)
    pt = sd['point_token'].expand(b, -1, -1) + sd['point_pos_embed'].expand(b, -1, -1)
    return torch.cat((x, pt), dim=1)


def backbone_forward(img, sd, depth, num_heads, patch=16, n_point_tokens=100):
    """VTD:221-275 restricted to what the attention-shift path consumes:
    -> dict(attns = list depth x [B,T,T], last_feat [B,1+N,C], point_tokens [B,Tp,C], x [B,T,C])."""
    x = prepare_tokens(img, sd, patch)
    attns = []
    for i in range(depth):
        x, a = block(x, sd, f'blocks.{i}.', num_heads)
        attns.append(a)
    return dict(attns=attns, last_feat=x[:, :-n_point_tokens], point_tokens=x[:, -n_point_tokens:], x=x)

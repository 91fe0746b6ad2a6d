"""``Block`` -- the reference's pre-LN transformer block (models/vision_transformer.py:88-124, ``Attention`` VT:62-86, ``Mlp``
VT:40-59) as a stand-alone module on the device kernels, for the RoI decoders that reuse it on short sequences (SURVEY.md 8f-3:
``MAEBoxHeadRec`` -- 4 blocks of dim 256 / 8 heads on 50 tokens per RoI, mae_bbox_head_rec.py:45-52,148-167;
``MAEMaskHeadPointSup`` -- the same on 197 tokens, mae_mask_head_pointSup.py:83-90,172-190).

Same constructor signature and parameter names as the reference class (``norm1``, ``attn.qkv``, ``attn.proj``, ``norm2``,
``mlp.fc1``, ``mlp.fc2``), so a decoder's checkpoint loads.  Forward-only (inference / no-grad): LayerNorm -> fp16 kernel,
qkv / proj / fc1(+GELU) / fc2 on the tcgen05 GEMM, attention on ``as_mhsa_small`` (head_dim 32 or 64, T <= 256) -- or on the
flash kernel of the backbone for head_dim 64 and longer sequences.  ``drop`` / ``drop_path`` are the identity in inference mode;
``init_values`` (layer scale) is supported as in the reference.
"""
import torch
import torch.nn as nn

from . import lib as _l
from . import ops


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


def mhsa_small(qkv, B, T, heads, head_dim):
    """VT:79-83 on the qkv Linear's output [B*T, 3*heads*head_dim] fp16 -> [B*T, heads*head_dim] fp16 (T <= 256)."""
    L = _l.load()
    o = torch.empty(B * T, heads * head_dim, device=qkv.device, dtype=torch.float16)
    _l.check(L.as_mhsa_small(_l.ptr(qkv), _l.ptr(o), B, T, heads, head_dim, _l.stream_ptr()), 'as_mhsa_small')
    return o


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, init_values=0, return_attention=False):
        super().__init__()
        if dim % num_heads or dim // num_heads not in (32, 64):
            raise ValueError('head_dim must be 32 or 64')
        if qk_scale:
            raise ValueError('qk_scale is not supported (the shipped configs leave it None: head_dim^-0.5)')
        if act_layer is not nn.GELU:
            raise ValueError('the fused fc1 epilogue implements GELU (the reference default)')
        if return_attention:
            raise ValueError('return_attention is the backbone\'s business (VisionTransformerDet): the RoI decoders never ask for it')
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        if init_values > 0:                                         # VT:103-107
            self.gamma_1 = nn.Parameter(init_values * torch.ones(dim))
            self.gamma_2 = nn.Parameter(init_values * torch.ones(dim))
        else:
            self.gamma_1 = self.gamma_2 = None
        self._w16 = {}

    def _half(self, name, p):
        key = (p.data_ptr(), p._version, p.device)
        hit = self._w16.get(name)
        if hit is None or hit[0] != key:
            hit = (key, p.detach().to(torch.float16).contiguous())
            self._w16[name] = hit
        return hit[1]

    @torch.no_grad()
    def forward(self, x, return_attention=False):
        B, T, C = x.shape
        h, d = self.num_heads, self.head_dim
        xs = x.reshape(B * T, C).float().contiguous()
        xn = ops.layernorm_f16(xs, self.norm1.weight.detach(), self.norm1.bias.detach(), self.norm1.eps)
        qb = self.attn.qkv.bias
        if T <= 256:
            qkv = ops.linear_f16(xn, self._half('qkv', self.attn.qkv.weight), None if qb is None else qb.detach().float(), ops.EPI_F16)
            o = mhsa_small(qkv, B, T, h, d)
        else:
            if d != 64:
                raise ValueError('sequences longer than 256 tokens need head_dim 64 (the backbone\'s flash kernel)')
            Tpad = (T + 127) // 128 * 128
            q, k, vt = ops.qkv_proj(xn, self._half('qkv', self.attn.qkv.weight), None if qb is None else qb.detach(), B, T, h, Tpad)
            o = ops.mhsa_fwd(q, k, vt, T)[0].view(B * T, C)
        pw, pb = self._half('proj', self.attn.proj.weight), self.attn.proj.bias.detach()
        if self.gamma_1 is None:
            xs = ops.linear_f16(o, pw, pb, ops.EPI_RESID_F32, resid=xs)
        else:
            xs = xs + self.gamma_1.detach() * ops.linear_f16(o, pw, pb, ops.EPI_F32)
        xn = ops.layernorm_f16(xs, self.norm2.weight.detach(), self.norm2.bias.detach(), self.norm2.eps)
        hid = ops.linear_f16(xn, self._half('fc1', self.mlp.fc1.weight), self.mlp.fc1.bias.detach(), ops.EPI_GELU_F16)
        fw, fb = self._half('fc2', self.mlp.fc2.weight), self.mlp.fc2.bias.detach()
        if self.gamma_2 is None:
            xs = ops.linear_f16(hid, fw, fb, ops.EPI_RESID_F32, resid=xs)
        else:
            xs = xs + self.gamma_2.detach() * ops.linear_f16(hid, fw, fb, ops.EPI_F32)
        return xs.view(B, T, C)

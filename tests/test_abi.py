"""The C-ABI library loads without a GPU and exports every symbol declared in include/attnshift_b200.h; the ctypes
binding table mirrors the header; the host-side helpers (registry shim, RNG front ends) behave like the reference's."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'attnshift_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(as_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from attentionshift_b200 import build, lib
    path = build.build()
    cdll = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(cdll, s), f'{s} declared in the header but not exported'
    assert sorted(lib.SIGNATURES) == syms, set(lib.SIGNATURES) ^ set(syms)
    lib.load()


def test_header_argument_counts_match_binding():
    from attentionshift_b200 import lib
    src = open(os.path.join(ROOT, 'include', 'attnshift_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    for name, args in re.findall(r'\b(as_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        n = len([a for a in args.split(',') if a.strip() and a.strip() != 'void'])
        assert n == len(lib.SIGNATURES[name][1]), name


def test_no_product_import_of_the_oracle():
    pkg = os.path.join(ROOT, 'attentionshift_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith('.py'):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_registry_names_and_ctor_surface():
    from attentionshift_b200 import registry
    registry._register_all()
    for n in ('AttnShiftRoIHead', 'StandardRoIHeadMaskPointSampleDeformAttnReppoints'):
        assert n in registry.HEADS.module_dict if hasattr(registry.HEADS, 'module_dict') else True
    assert registry.HEADS.get('AttnShiftRoIHead') is registry.HEADS.get('StandardRoIHeadMaskPointSampleDeformAttnReppoints')
    bb = registry.build_backbone(dict(type='VisionTransformerDet', img_size=64, patch_size=16, embed_dim=128, depth=1, num_heads=2,
                                      mlp_ratio=4, qkv_bias=True, last_feat=True, return_attention=True, point_tokens_num=8))
    names = set(dict(bb.named_parameters()))
    # MAE checkpoints must load: reference parameter names (VTD:184)
    for k in ('cls_token', 'pos_embed', 'point_token', 'point_pos_embed', 'patch_embed.proj.weight', 'blocks.0.norm1.weight',
              'blocks.0.attn.qkv.weight', 'blocks.0.attn.qkv.bias', 'blocks.0.attn.proj.weight', 'blocks.0.mlp.fc1.weight',
              'blocks.0.mlp.fc2.bias', 'fpn1.0.weight', 'class_embed.layers.0.weight', 'bbox_embed.layers.2.bias'):
        assert k in names, k
    with pytest.raises(KeyError):
        registry.build_backbone(dict(type='NoSuchBackbone'))
    with pytest.raises(TypeError):
        bb.init_weights(pretrained=3)
    assert bb.train(False) is bb


def test_keyed_rng_equals_torch_generators():
    from attentionshift_b200.attention_shift import KeyedRng, StreamRng, _fill_index
    r = KeyedRng(3)
    for n in (7, 4096, 5000, 250001):
        key = (1, 1, 2)
        torch.manual_seed(r.seed_for(key))
        assert torch.equal(r.randperm_head(key, n, 10), torch.randperm(n)[:10])
        torch.manual_seed(r.seed_for(key))
        assert torch.equal(r.randint(key, n, 21), torch.randint(n, (21,)))
    torch.manual_seed(9)
    a = StreamRng().randint(None, 100, 5)
    torch.manual_seed(9)
    assert torch.equal(a, torch.randint(100, (5,)))
    assert _fill_index(torch.tensor([4, 9]), 10).tolist() == [4, 9] * 5
    assert _fill_index(torch.tensor([1, 2, 3, 4, 5, 6]), 10).tolist() == [1, 2, 3, 4, 5, 6, 1, 2, 3, 4]


def test_hot_path_raises_without_cuda():
    """No CPU fallback: the product path must fail loudly when it cannot run on the CUDA extension."""
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from attentionshift_b200 import ops
    with pytest.raises(Exception):
        ops.layernorm_f16(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))

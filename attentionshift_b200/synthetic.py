"""Synthetic inputs for the hot path.

Random-init ViT weights give near-uniform attention and unstructured features, so the
attention-shift stage gets a *structured* generator: ``n_obj`` disks on the patch
grid, token feature = per-region base vector + noise, CAMs = disk indicator + noise.
Used by the tests, the golden-vector script and bench.py (CPU tensors; callers move
them to the device).
"""
import torch

PATCH = 16
CAM_LAYERS = 7


def structured_scene(hp, wp, c, n_obj, seed, noise=0.5, cam_layers=CAM_LAYERS, core=0.75, part_scale=0.6):
    """-> dict(vit_feat [C,hp,wp], cams_low [L,n_obj,hp,wp], gt_points [n_obj,2] (x,y px),
    rois [n_obj,4] px, gt_index [n_obj] long, gt_labels [n_obj] long, labels [hp,wp] long).

    Every instance is a disk of radius 0.18*min(hp,wp) patches made of two *parts* (an inner
    core of radius ``core``*r and the surrounding ring): token = instance vector + part vector
    + noise, so the mean-shift stage finds real part structure (interior parts survive the
    foreground filter of RH:265-275, boundary parts do not)."""
    g = torch.Generator().manual_seed(seed)
    r = 0.18 * min(hp, wp)
    yy, xx = torch.meshgrid(torch.arange(hp, dtype=torch.float32), torch.arange(wp, dtype=torch.float32), indexing='ij')
    centers = torch.empty(n_obj, 2)
    labels = torch.zeros(hp, wp, dtype=torch.long)          # 0 bg, 2i+1 core of i, 2i+2 ring of i
    owner = torch.zeros(hp, wp, dtype=torch.long)           # 0 bg, i+1 instance i
    for i in range(n_obj):
        cy = r + (hp - 1 - 2 * r) * torch.rand((), generator=g)
        cx = r + (wp - 1 - 2 * r) * torch.rand((), generator=g)
        centers[i, 0], centers[i, 1] = cx, cy
        d2 = (yy - cy) ** 2 + (xx - cx) ** 2
        labels[d2 <= r * r] = 2 * i + 2
        labels[d2 <= (core * r) ** 2] = 2 * i + 1
        owner[d2 <= r * r] = i + 1
    obj_base = torch.randn(n_obj + 1, c, generator=g)
    part_base = torch.randn(2 * n_obj + 1, c, generator=g)
    feat = obj_base[owner] + part_scale * part_base[labels] + noise * torch.randn(hp, wp, c, generator=g)   # [hp,wp,C]
    cams = torch.empty(cam_layers, n_obj, hp, wp)
    for i in range(n_obj):
        d = (owner == i + 1)
        for l in range(cam_layers):
            cams[l, i] = d.float() + 0.05 * torch.rand(hp, wp, generator=g)
    gt_points = (centers * PATCH).floor()
    rois = torch.stack([(centers[:, 0] - r).clamp(0) * PATCH, (centers[:, 1] - r).clamp(0) * PATCH,
                        (centers[:, 0] + r + 1).clamp(max=wp - 1) * PATCH, (centers[:, 1] + r + 1).clamp(max=hp - 1) * PATCH], dim=1).floor()
    gt_index = torch.randint(0, cam_layers, (n_obj,), generator=g)
    gt_labels = torch.randint(0, 20, (n_obj,), generator=g)
    return dict(vit_feat=feat.permute(2, 0, 1).contiguous(), cams_low=cams, gt_points=gt_points, rois=rois,
                gt_index=gt_index, gt_labels=gt_labels, labels=labels)


def vit_state_dict(embed_dim, depth, num_heads, img_size, patch=PATCH, n_point_tokens=100, seed=0,
                   mlp_ratio=4, in_chans=3, std=0.02):
    """Random-init backbone weights with the reference's parameter names and init
    (trunc_normal std .02 for Linear / tokens / tables, LayerNorm = (1, 0): VT:173-185,
    VTD:149-150).  Conv patch-embed keeps torch's default init like the reference."""
    g = torch.Generator().manual_seed(seed)

    def tn(*shape):
        return torch.nn.init.trunc_normal_(torch.empty(*shape), std=std, generator=g)

    sd = {}
    n = (img_size // patch) ** 2
    fan_in = in_chans * patch * patch
    bound = 1.0 / fan_in ** 0.5
    sd['patch_embed.proj.weight'] = (torch.rand(embed_dim, in_chans, patch, patch, generator=g) * 2 - 1) * bound
    sd['patch_embed.proj.bias'] = (torch.rand(embed_dim, generator=g) * 2 - 1) * bound
    sd['cls_token'] = tn(1, 1, embed_dim)
    sd['pos_embed'] = tn(1, n + 1, embed_dim)
    sd['point_token'] = tn(1, n_point_tokens, embed_dim)
    sd['point_pos_embed'] = tn(1, n_point_tokens, embed_dim)
    hid = int(embed_dim * mlp_ratio)
    for i in range(depth):
        p = f'blocks.{i}.'
        sd[p + 'norm1.weight'] = torch.ones(embed_dim)
        sd[p + 'norm1.bias'] = torch.zeros(embed_dim)
        sd[p + 'attn.qkv.weight'] = tn(3 * embed_dim, embed_dim)
        sd[p + 'attn.qkv.bias'] = torch.zeros(3 * embed_dim)
        sd[p + 'attn.proj.weight'] = tn(embed_dim, embed_dim)
        sd[p + 'attn.proj.bias'] = torch.zeros(embed_dim)
        sd[p + 'norm2.weight'] = torch.ones(embed_dim)
        sd[p + 'norm2.bias'] = torch.zeros(embed_dim)
        sd[p + 'mlp.fc1.weight'] = tn(hid, embed_dim)
        sd[p + 'mlp.fc1.bias'] = torch.zeros(hid)
        sd[p + 'mlp.fc2.weight'] = tn(embed_dim, hid)
        sd[p + 'mlp.fc2.bias'] = torch.zeros(embed_dim)
    return sd

"""MIL layer selection of ``seed_pseudo_gt`` (RH:2308-2312, ``_mil_forward_train`` RH:2953-2972): RoIAlign 7x7 over the
7 x n_gt per-layer pseudo boxes on the stride-16 ViT feature map, then ``MAEBoxHeadMIL``
(mmdet/models/roi_heads/bbox_heads/mae_bbox_head_mil.py:140-169), which scores every (instance, layer) box and picks, per
instance, the layer whose box explains the instance's class best.

Adjacent to the device hot path: a few hundred RoIs and four small Linear layers -- left to library
ops (``torchvision.ops.roi_align`` = mmcv's RoIAlign with ``aligned=True``, cuBLAS Linear).  It is differentiable like the
reference's, so ``mil_loss`` trains the MIL head when the caller runs it with gradients enabled.  Parameter names equal the
reference's (``norm``, ``decoder_embed``, ``fc1``, ``fc2``, ``proposal_branch``, ``classification_branch``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .registry import HEADS


class MAEBoxHeadMIL(nn.Module):
    def __init__(self, in_channels, img_size=224, patch_size=16, embed_dim=256, depth=4, num_heads=8, mlp_ratio=4.,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., pretrained=False,
                 use_checkpoint=False, num_layers_query=12, loss_mil_factor=1.0, hidden_dim=1024, roi_size=7, num_classes=80,
                 **kwargs):
        super().__init__()
        self.pretrained = pretrained
        self.num_classes = num_classes
        self.num_layers_query = num_layers_query
        self.loss_mil_factor = loss_mil_factor
        self.hidden_dim = hidden_dim
        self.roi_size = roi_size
        self.with_decoder_embed = in_channels != embed_dim                       # MIL:50-54
        if self.with_decoder_embed:
            self.norm = nn.LayerNorm(in_channels, eps=1e-6)
            self.decoder_embed = nn.Linear(in_channels, embed_dim, bias=True)
        self.fc1 = nn.Linear(embed_dim * roi_size ** 2, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.proposal_branch = nn.Linear(hidden_dim, num_classes)
        self.classification_branch = nn.Linear(hidden_dim, num_classes)

    def _init_weights(self, m):                                                   # MIL:63-70
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02, a=-2., b=2.)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def init_weights(self, pretrained=None):                                      # MIL:72-101
        """Reference initialisation: trunc_normal(0.02) Linear weights, zero biases, unit LayerNorm -- or, when the head was
        built with ``pretrained=True`` and a checkpoint path is given, that checkpoint minus the encoder's entries.  (The
        reference's ``pretrained is None`` branch also touches a ``det_token`` that its constructor never creates.)"""
        if self.pretrained and isinstance(pretrained, str):
            import os
            if not os.path.isfile(pretrained):
                raise ValueError(f'checkpoint path {pretrained} is invalid')
            ckpt = torch.load(pretrained, map_location='cpu')
            sd = ckpt.get('state_dict', ckpt.get('model', ckpt))
            sd = {k: v for k, v in sd.items() if not (k.startswith('patch_embed') or k.startswith('blocks') or k == 'pos_embed')}
            self.load_state_dict(sd, strict=False)
        elif pretrained is None or isinstance(pretrained, str):
            self.apply(self._init_weights)
        else:
            raise TypeError('pretrained must be a str or None')

    def mil_losses(self, cls_score, labels):                                      # MIL:134-138
        cls_score = cls_score.clamp(1e-6, 1 - 1e-6)
        labels = labels.clamp(0, 1)
        return (-labels * torch.log(cls_score) - (1 - labels) * torch.log(1 - cls_score)).mean()

    def forward(self, x, gt_labels=None):
        """x [n_inst * L, C, r, r] RoI features (layers fastest), gt_labels [n_inst] (or a per-image list).
        -> (layer index per instance [n_inst], mil_loss)."""
        if isinstance(gt_labels, (list, tuple)):
            gt_labels = torch.cat(list(gt_labels))
        n = x.shape[0]
        t = x.flatten(2).transpose(1, 2)                                         # [n, r*r, C]
        if self.with_decoder_embed:
            t = self.decoder_embed(self.norm(t))
        t = F.relu(self.fc1(t.reshape(n, -1)))
        t = F.relu(self.fc2(t))
        L, K = self.num_layers_query, self.num_classes
        cls = self.classification_branch(t).reshape(-1, L, K).softmax(-1)         # which class, per box
        prop = self.proposal_branch(t).reshape(-1, L, K).softmax(-2)              # which layer, per class
        bag = cls * prop
        score = torch.gather(bag, dim=-1, index=gt_labels.reshape(-1, 1, 1).repeat(1, L, 1))[..., 0]
        gt_index = score.max(-1)[1]
        binary = torch.zeros((len(gt_labels), K)).type_as(gt_labels)
        binary[torch.arange(len(gt_labels)).type_as(gt_labels), gt_labels] = 1
        return gt_index, self.loss_mil_factor * self.mil_losses(bag.sum(1), binary)


if 'MAEBoxHeadMIL' not in HEADS.module_dict:      # inside a real mmdet the reference's own class keeps the registry name
    HEADS.register_module(module=MAEBoxHeadMIL)


def boxes_to_rois(boxes_per_img):
    """mmdet bbox2roi (core/bbox/transforms.py:58-77): [batch index, x1, y1, x2, y2] rows, images in order."""
    rois = []
    for i, b in enumerate(boxes_per_img):
        b = b.reshape(-1, 4)
        rois.append(torch.cat([b.new_full((b.shape[0], 1), i), b], dim=-1) if b.shape[0] else b.new_zeros((0, 5)))
    return torch.cat(rois, 0)


def mil_select(mil_head, feature_map, boxes_per_img, gt_labels, stride=16, roi_size=7):
    """``_mil_forward_train`` (RH:2953-2972) without the box gather the caller does itself.
    feature_map [B,C,Hp,Wp] (or the reference's one-element list); boxes_per_img: per image [n_i, L, 4].
    -> (per-image list of layer indices, {'mil_loss': loss})."""
    from torchvision.ops import roi_align
    fmap = feature_map[0] if isinstance(feature_map, (list, tuple)) else feature_map
    if fmap is None:
        raise ValueError('the MIL stage needs roi_feature_map= (the stride-16 ViT feature map, DET:85)')
    rois = boxes_to_rois(boxes_per_img).to(fmap.dtype)
    feats = roi_align(fmap, rois, roi_size, spatial_scale=1.0 / stride, sampling_ratio=0, aligned=True)
    idx, loss = mil_head(feats, gt_labels=gt_labels)
    return list(idx.split([int(b.shape[0]) for b in boxes_per_img], dim=0)), {'mil_loss': loss}


def _pad_cols(t, mult=64):
    k = t.shape[1]
    kp = (k + mult - 1) // mult * mult
    if kp == k:
        return t.contiguous()
    out = t.new_zeros(t.shape[0], kp)
    out[:, :k] = t
    return out


@torch.no_grad()
def mil_select_device(mil_head, feats, boxes_per_img, gt_labels, hp, wp, stride=16, roi_size=7):
    """The selection half of ``_mil_forward_train`` (RH:2953-2972) on this repo's kernels, for calls that do not train the MIL
    head (inference, frozen head, ``torch.no_grad()``): RoIAlign on the token-major feature map (``as_roi_align_tokens`` -- no
    [B,C,Hp,Wp] transpose), LayerNorm -> fp16, ``decoder_embed`` / ``fc1`` / ``fc2`` / both branches on the tcgen05 GEMM (fp16
    operands, fp32 accumulate; the two 20-class branches share one GEMM), the softmaxes / gather / arg-max of MIL:155-161 on the
    [n_inst, L, classes] scores in torch.  feats [B, hp*wp, C] fp32; boxes_per_img: per image [n_i, L, 4].
    -> per-image list of layer indices.  (The training call keeps ``mil_select``: its loss needs autograd.)"""
    from . import lib as _l
    from . import ops
    L = _l.load()
    dev = feats.device
    rois = boxes_to_rois(boxes_per_img).float().to(dev).contiguous()
    R, C = rois.shape[0], feats.shape[2]
    x = torch.empty(R, roi_size * roi_size, C, device=dev, dtype=torch.float32)
    assert feats.dtype == torch.float32 and feats.stride(2) == 1 and feats.stride(1) == C
    _l.check(L.as_roi_align_tokens(_l.ptr(feats) if feats.is_contiguous() else __import__('ctypes').c_void_p(feats.data_ptr()),
                                   feats.stride(0), _l.ptr(rois), R, hp, wp, C, roi_size, 1.0 / stride, _l.ptr(x), _l.stream_ptr()),
             'as_roi_align_tokens')
    h = lambda p: p.detach().half().contiguous()
    t = x.view(R * roi_size * roi_size, C)
    if mil_head.with_decoder_embed:
        t = ops.layernorm_f16(t, mil_head.norm.weight.detach(), mil_head.norm.bias.detach(), mil_head.norm.eps)
        t = ops.linear_f16(_pad_cols(t), _pad_cols(h(mil_head.decoder_embed.weight)), mil_head.decoder_embed.bias.detach().float(), ops.EPI_F16)
    else:
        t = t.half()
    t = t.reshape(R, -1)
    t = torch.relu(ops.linear_f16(_pad_cols(t), _pad_cols(h(mil_head.fc1.weight)), mil_head.fc1.bias.detach().float(), ops.EPI_F32)).half()
    t = torch.relu(ops.linear_f16(_pad_cols(t), _pad_cols(h(mil_head.fc2.weight)), mil_head.fc2.bias.detach().float(), ops.EPI_F32)).half()
    K = mil_head.num_classes
    w = torch.cat((mil_head.classification_branch.weight, mil_head.proposal_branch.weight)).detach()
    b = torch.cat((mil_head.classification_branch.bias, mil_head.proposal_branch.bias)).detach().float()
    n_out = (2 * K + 31) // 32 * 32                                  # the GEMM writes whole 32-column groups
    wp_ = w.new_zeros(n_out, w.shape[1]); wp_[:2 * K] = w
    bp_ = b.new_zeros(n_out); bp_[:2 * K] = b
    s = ops.linear_f16(_pad_cols(t), _pad_cols(wp_.half()), bp_, ops.EPI_F32)
    Lq = mil_head.num_layers_query
    cls = s[:, :K].reshape(-1, Lq, K).softmax(-1)                     # MIL:155-156
    prop = s[:, K:2 * K].reshape(-1, Lq, K).softmax(-2)
    bag = cls * prop
    labels = torch.cat(list(gt_labels)) if isinstance(gt_labels, (list, tuple)) else gt_labels
    score = torch.gather(bag, dim=-1, index=labels.to(dev).reshape(-1, 1, 1).repeat(1, Lq, 1))[..., 0]
    idx = score.max(-1)[1]
    return list(idx.split([int(b_.shape[0]) for b_ in boxes_per_img], dim=0)), score

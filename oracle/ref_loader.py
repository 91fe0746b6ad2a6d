"""Run the UNMODIFIED reference code on CPU (build container only).

TEST INFRASTRUCTURE.  Mechanism:

* ``models/vision_transformer.py`` imports cleanly once ``timm.models.registry`` is
  stubbed (imported at VT:19, never used) and ``/root/reference`` is on sys.path.
* ``mmdet/models/roi_heads/stdroi_point_deform_attn_reppoints.py`` (RH) cannot be
  imported (needs mmcv / mmdet.core / cc_torch / matplotlib and a broken package
  __init__), but every hot-path function in it is pure torch: we ``ast.parse`` the
  file and ``exec`` each top-level FunctionDef in file order (so the LAST definition
  of a name wins, exactly as Python binds them), and lift the few needed methods
  out of the ClassDef body to be called with a SimpleNamespace as ``self``.
* ``mmdet/models/backbones/visual_transformer_det.py`` (VTD) is extracted the same
  way with stubs for BACKBONES / load_checkpoint / get_root_logger.

``cc_torch.connected_components_labeling`` is absent from the tree; the loader
injects the oracle's scipy 8-connectivity stand-in (documented as unpinned).  ``mmcv.ops.point_sample`` (only used by
the second-round aggregation, RH:2737-2844) is injected the same way: a restatement of mmcv-full 1.3.8's published
``F.grid_sample`` wrapper.
"""
import ast
import math
import os
import random
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("ATTNSHIFT_REFERENCE", "/root/reference")
RH_PATH = "mmdet/models/roi_heads/stdroi_point_deform_attn_reppoints.py"
VTD_PATH = "mmdet/models/backbones/visual_transformer_det.py"


def available():
    return os.path.isfile(os.path.join(REF_ROOT, RH_PATH))


_cache = {}


def _stub_timm():
    if "timm.models.registry" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    registry = types.ModuleType("timm.models.registry")
    registry.register_model = lambda f: f
    timm.models = models
    models.registry = registry
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", models)
    sys.modules.setdefault("timm.models.registry", registry)


def load_vt():
    """-> module object of reference models/vision_transformer.py"""
    if "vt" in _cache:
        return _cache["vt"]
    _stub_timm()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_vision_transformer", os.path.join(REF_ROOT, "models/vision_transformer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache["vt"] = mod
    return mod


def load_vtd():
    """-> the reference class VisionTransformerDet (VTD:60-275)."""
    if "vtd" in _cache:
        return _cache["vtd"]
    vt = load_vt()
    src = open(os.path.join(REF_ROOT, VTD_PATH)).read()
    tree = ast.parse(src)

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c

    ns = dict(torch=torch, nn=nn, F=F, os=os, math=math,
              checkpoint=torch.utils.checkpoint,
              BACKBONES=_Reg(), VisionTransformer=vt.VisionTransformer,
              trunc_normal_=vt.trunc_normal_,
              load_checkpoint=lambda *a, **k: None,
              get_root_logger=lambda *a, **k: types.SimpleNamespace(info=lambda *a, **k: None))
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            code = compile(ast.Module(body=[node], type_ignores=[]), VTD_PATH, "exec")
            exec(code, ns)
    _cache["vtd"] = ns["VisionTransformerDet"]
    return _cache["vtd"]


def load_rh():
    """-> namespace (SimpleNamespace) with the reference RH module functions plus
    the class methods we need as plain functions taking ``self`` first."""
    if "rh" in _cache:
        return _cache["rh"]
    from oracle.attnshift import ccl_label, point_sample  # unpinned stand-ins for cc_torch / mmcv.ops.point_sample
    src = open(os.path.join(REF_ROOT, RH_PATH)).read()
    tree = ast.parse(src)
    ns = dict(torch=torch, nn=nn, F=F, math=math, random=random, np=np, os=os, point_sample=point_sample,
              connected_components_labeling=lambda m: torch.from_numpy(
                  ccl_label(m.cpu().numpy())).to(m.device))
    methods = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            code = compile(ast.Module(body=[node], type_ignores=[]), RH_PATH, "exec")
            exec(code, ns)
        elif isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef):
                    sub.decorator_list = []
                    code = compile(ast.Module(body=[sub], type_ignores=[]), RH_PATH, "exec")
                    exec(code, ns, methods)
    out = types.SimpleNamespace(**{k: v for k, v in ns.items() if callable(v)})
    out.methods = types.SimpleNamespace(**methods)
    _cache["rh"] = out
    return out


def load_point_assigner():
    """-> (HungarianPointAssigner class, PointPseudoSampler class) of the reference, AST-extracted with stubs for the mmdet
    registries (``BBOX_ASSIGNERS``, ``MATCH_COST``, ``BBOX_SAMPLERS``), ``AssignResult`` and the sampling result."""
    if "assigner" in _cache:
        return _cache["assigner"]

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c

    class AssignResult:
        def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
            self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels

    class PointSamplingResult:
        def __init__(self, pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags):
            self.pos_inds, self.neg_inds = pos_inds, neg_inds
            self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1

    ns = dict(torch=torch, MATCH_COST=_Reg(), BBOX_ASSIGNERS=_Reg(), BBOX_SAMPLERS=_Reg(), AssignResult=AssignResult,
              BaseAssigner=object, BaseSampler=object, PointSamplingResult=PointSamplingResult, SamplingResult=object,
              bbox_overlaps=None, bbox_cxcywh_to_xyxy=None, bbox_xyxy_to_cxcywh=None)
    try:
        from scipy.optimize import linear_sum_assignment
    except ImportError:
        linear_sum_assignment = None
    ns["linear_sum_assignment"] = linear_sum_assignment
    for rel in ("mmdet/core/bbox/match_costs/match_cost.py", "mmdet/core/bbox/assigners/hungarian_point_assigner.py",
                "mmdet/core/bbox/samplers/point_pseudo_sampler.py"):
        tree = ast.parse(open(os.path.join(REF_ROOT, rel)).read())
        for node in tree.body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                exec(compile(ast.Module(body=[node], type_ignores=[]), rel, "exec"), ns)
    ns["build_match_cost"] = lambda cfg: ns[cfg["type"]](**{k: v for k, v in cfg.items() if k != "type"})
    _cache["assigner"] = (ns["HungarianPointAssigner"], ns["PointPseudoSampler"])
    return _cache["assigner"]


def load_mil_head():
    """-> the reference class MAEBoxHeadMIL (mae_bbox_head_mil.py:18-169), AST-extracted; its mmdet base class ``BBoxHead`` is
    replaced by a stub that only keeps ``num_classes`` (the shipped config builds it with ``with_cls=False, with_reg=False``,
    so the base contributes no parameters)."""
    if "mil" in _cache:
        return _cache["mil"]
    vt = load_vt()
    from functools import partial
    from collections import OrderedDict

    class BBoxHead(nn.Module):
        def __init__(self, num_classes=80, **kwargs):
            super().__init__()
            self.num_classes = num_classes

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c

    ns = dict(torch=torch, nn=nn, F=F, os=os, math=math, partial=partial, OrderedDict=OrderedDict, HEADS=_Reg(), BBoxHead=BBoxHead,
              Block=vt.Block, trunc_normal_=vt.trunc_normal_, get_root_logger=lambda *a, **k: None,
              _load_checkpoint=None, load_state_dict=None, checkpoint=None)
    rel = "mmdet/models/roi_heads/bbox_heads/mae_bbox_head_mil.py"
    tree = ast.parse(open(os.path.join(REF_ROOT, rel)).read())
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            exec(compile(ast.Module(body=[node], type_ignores=[]), rel, "exec"), ns)
    _cache["mil"] = ns["MAEBoxHeadMIL"]
    return _cache["mil"]

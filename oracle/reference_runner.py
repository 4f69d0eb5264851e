"""Runs the UNMODIFIED reference head from /root/reference on the CPU -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``bench.py --impl reference`` / its ``cpu_baseline`` leg use this, and only where /root/reference exists (the build
container; the GPU box has no copy, there the oracle port in ``checkerpose_oracle.py`` is timed instead and the JSON line
says ``kind: "port"``).  Nothing is copied from the reference: its modules are imported from where they lie, with the
absent third-party packages that are off the path (timm = backbone, ...) stubbed in ``sys.modules`` exactly as
``tests/golden/make_golden.py`` does.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference/checkerpose"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        return None


class _FeatureBackbone(nn.Module):
    """Stands in for timm's HRNet-W18 features_only model: returns the maps it is given (the backbone is off the path)."""

    def forward(self, feats):
        return list(feats)


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("timm", "pytz", "mmcv", "imgaug", "imgaug.augmenters", "imageio", "pyprogressivex", "plyfile"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    import model.backbone as ref_backbone
    ref_backbone.get_timm_backbone = lambda **kw: _FeatureBackbone()
    import model.init as ref_init
    import model.pipeline as ref_pipe
    ref_init.get_timm_backbone = ref_backbone.get_timm_backbone
    return ref_init, ref_pipe


def build_reference_head(N, K, p3d_normed, state_dict, max_batch):
    """-> run(feats) calling the reference's PoseNet_GNNskip.forward (pipeline.py:351-384) on CPU tensors."""
    ref_init, ref_pipe = _import_reference()
    inet = ref_init.InitNet_GNN(npoint=N, p3d_normed=p3d_normed, res_log2=3, backbone_name="hrnet_w18", pretrain_backbone=False,
                                max_batch_size=max_batch, num_graph_module=2, graph_k=K)
    net = ref_pipe.PoseNet_GNNskip(inet, npoint=N, p3d_normed=p3d_normed, res_log2=6, num_filters=256, max_batch_size=max_batch,
                                   local_k=2, leaky_slope=0.01, num_graph_module=3, graph_k=K)
    net.load_state_dict(state_dict, strict=True)
    net.eval()

    def run(feats):
        with torch.no_grad():
            return net(feats, p3d_normed.expand(feats[0].shape[0], -1, -1))
    return run

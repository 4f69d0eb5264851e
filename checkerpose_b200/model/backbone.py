"""Backbone loading -- stays on the reference's path (checkerpose/model/backbone.py:39-50).

The HRNet-W18 backbone is outside the accelerated hot path (BASELINE.json north_star): it is created
by timm exactly as the reference does.  timm is not installed in the build/bench image, so synthetic
runs pass ``img_backbone=FeatureListBackbone()`` to the nets and feed the four feature maps directly.
"""
import torch.nn as nn


class FeatureListBackbone(nn.Module):
    """Identity 'backbone' for synthetic runs: the input already is the list of backbone feature maps
    [(B,128,64,64), (B,256,32,32), (B,512,16,16), (B,1024,8,8)] (HRNet-W18, pipeline.py:12)."""

    def forward(self, feats):
        return list(feats)


def get_timm_backbone(model_name="resnet34", concat_decoder=True, pretrained=True):
    """Same contract as the reference: a timm ``features_only`` model returning all stage outputs."""
    try:
        import timm
    except ImportError as e:  # pragma: no cover - timm is absent in the build image
        raise RuntimeError("timm is required for the real backbone (reference path); for synthetic feature "
                           "maps construct the net with img_backbone=FeatureListBackbone()") from e
    if model_name in ["convnext_tiny", "convnext_small", "convnext_base"]:
        out_indices = (1, 2, 3) if concat_decoder else (3,)
    elif model_name in ["resnet34", "hrnet_w18", "hrnet_w18_small", "hrnet_w30"]:
        out_indices = (1, 2, 3, 4) if concat_decoder else (4,)
    elif model_name in ["darknet53"]:
        out_indices = (1, 2, 3, 4, 5) if concat_decoder else (5,)
    else:
        raise ValueError("timm_backbone {} not supported yet".format(model_name))
    return timm.create_model(model_name=model_name, pretrained=pretrained, in_chans=3, features_only=True,
                             out_indices=out_indices)

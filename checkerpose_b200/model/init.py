"""Drop-in for checkerpose/model/init.py: InitNet_GNN (low-level bits) on the sm_100a kernels."""
import torch
import torch.nn as nn

from .. import head, ops
from .backbone import get_timm_backbone
from .pipeline import LazyKnnGraph, StaticGraph_module, get_graph_feature, knn  # noqa: F401  (re-exported like the reference)

CONV1X1_IN_CHANS = {
    "resnet34": 512,
    "convnext_tiny": 768,
    "convnext_small": 768,
    "convnext_base": 1024,
    "darknet53": 1024,
    "hrnet_w18": 1024,
    "hrnet_w18_small": 1024,
    "hrnet_w30": 1024,
}


class InitNet_GNN(head.DeviceScopedModule):
    """init.py:71-128.  Extra keyword ``img_backbone`` (default None = timm, as the reference) lets
    synthetic runs inject ``FeatureListBackbone``; it adds no parameters."""

    _graph_module_cls = StaticGraph_module

    def __init__(self, npoint, p3d_normed, res_log2=3, backbone_name="resnet34", pretrain_backbone=True,
                 num_conv1x1=1, max_batch_size=64, num_graph_module=2, graph_k=20, graph_leaky_slope=0.2,
                 img_backbone=None):
        super(InitNet_GNN, self).__init__()
        self.num_out_bits = 1 + 2 * res_log2
        self.npoint = npoint
        self.backbone_name = backbone_name
        self.max_batch_size = max_batch_size
        self.img_backbone = img_backbone if img_backbone is not None else get_timm_backbone(
            model_name=backbone_name, concat_decoder=True, pretrained=pretrain_backbone)
        if num_conv1x1 == 1:
            self.conv1x1 = nn.Conv2d(in_channels=CONV1X1_IN_CHANS[backbone_name], out_channels=npoint,
                                     kernel_size=1, stride=1, padding=0)
        else:
            conv1x1 = [nn.Conv2d(in_channels=CONV1X1_IN_CHANS[backbone_name], out_channels=npoint,
                                 kernel_size=1, stride=1, padding=0)]
            for i in range(num_conv1x1 - 1):
                conv1x1.append(nn.LeakyReLU(negative_slope=0.01))
                conv1x1.append(nn.Conv2d(in_channels=npoint, out_channels=npoint, kernel_size=1, stride=1, padding=0))
            self.conv1x1 = nn.Sequential(*conv1x1)
        self.pre_query_block = nn.ModuleList()
        knn_idx = LazyKnnGraph(p3d_normed, graph_k)
        for i in range(num_graph_module):
            self.pre_query_block.append(self._graph_module_cls(input_dim=64, output_dim=64, knn_idx=knn_idx,
                                                               leaky_slope=graph_leaky_slope))
        self.mlp = nn.Linear(in_features=64, out_features=self.num_out_bits)

    def _forward_impl(self, img, obj_ids, return_img_feats, return_graph_feats):
        dtype = head.get_compute_dtype()
        img_feats = self.img_backbone(img)
        logits, gfeat, ctx = head.init_head_node_major(self, img_feats[-1], obj_ids, dtype)
        logits = logits[:, :, :self.num_out_bits].contiguous()
        if ctx is not None:   # plan order -> the reference's keypoint order
            logits = ctx.to_keypoints(logits)
        out = logits.permute(0, 2, 1)  # (B, #bits, N)
        if return_img_feats:
            return out, img_feats
        elif return_graph_feats:
            if ctx is not None:
                gfeat = ctx.to_keypoints(gfeat)
            return out, img_feats, ops.convert(gfeat, img_feats[-1].dtype).permute(0, 2, 1)
        return out

    def forward(self, img, return_img_feats=False, return_graph_feats=False):
        return self._forward_impl(img, None, return_img_feats, return_graph_feats)

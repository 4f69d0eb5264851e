"""Drop-in for checkerpose/model/pipeline_lm.py (LM-13 multi-object nets): identical maths, but
``knn_idx`` is a (num_obj, N, K) table and every EdgeConv selects its graph per sample with the
1-based ``obj_ids`` (pipeline_lm.py:55-57) -- the kernels take that as a per-RoI graph selector."""
import torch
import torch.nn as nn

from .. import head, ops
from .pipeline import (IMG_FEATS_DIMS, Index2Feat_module, MLP_QueryNet, from_bit_prob_to_id,  # noqa: F401
                       from_code_prob_to_id, from_code_to_id, from_gt_bit_to_id, from_gt_code_to_id,
                       from_mask_prob_to_mask, get_gdrn_upsample_module, get_graph_feature,
                       get_MLP_leakyReLU_layers, knn, _to_io)
from . import pipeline as _single


class StaticGraph_module(_single.StaticGraph_module):
    def forward(self, x, batch_indices, obj_ids):
        return self._forward_impl(x, obj_ids)


class Refine_moduleGNN(_single.Refine_moduleGNN):
    _graph_module_cls = StaticGraph_module

    def forward(self, img_feat, graph_feat, p3d_normed, roi_mask_bit, prev_x_id, prev_y_id, obj_ids):
        return self._forward_impl(img_feat, graph_feat, roi_mask_bit, prev_x_id, prev_y_id, obj_ids)


class PoseNet_GNNskip(_single.PoseNet_GNNskip):
    _refine_cls = Refine_moduleGNN

    def forward(self, img, p3d_normed, obj_ids, stage=None):
        img_feats = self.init_net.img_backbone(img)
        out, _ = head.pose_head_forward(self, img_feats, obj_ids=obj_ids, stage=stage)
        return out

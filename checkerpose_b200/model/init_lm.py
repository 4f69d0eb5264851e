"""Drop-in for checkerpose/model/init_lm.py: InitNet_GNN.forward(img, obj_ids, ...) (init_lm.py:110)."""
from . import init as _single
from .init import CONV1X1_IN_CHANS, get_graph_feature, knn  # noqa: F401
from .pipeline_lm import StaticGraph_module


class InitNet_GNN(_single.InitNet_GNN):
    _graph_module_cls = StaticGraph_module

    def forward(self, img, obj_ids, return_img_feats=False, return_graph_feats=False):
        return self._forward_impl(img, obj_ids, return_img_feats, return_graph_feats)

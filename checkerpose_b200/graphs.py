"""CUDA-graph capture of the head for a fixed batch shape.

One step of the head is ~150 kernel launches (34 of ours plus the library convolutions and torch glue).  At 256 RoIs per
GPU the host enqueues them faster than the GPU drains them, but at small shards (128 RoIs per GPU in the strong-scaled
YCB-V configuration, 64 RoIs in the init-net configuration) the launch stream becomes visible.  ``CapturedHead`` records
the whole forward once -- every kernel of the C ABI takes its stream from ``torch.cuda.current_stream()`` and never
synchronises, so the path is capturable as it stands -- and replays it with one ``cudaGraphLaunch``.

The capture owns static input buffers; ``__call__`` copies the caller's tensors into them on the current stream and
returns the static outputs (valid until the next call).
"""
from __future__ import annotations

import torch


class CapturedHead:
    def __init__(self, net, feats, p3d_normed, bbox=None, obj_ids=None, packed=True, stage=None, warmup=3):
        """feats: list of CUDA feature maps (the shapes / dtypes the graph is specialised for); bbox (B,4) or None (then the
        reference's 6-tuple only); obj_ids (B) for the LM nets."""
        self.net, self.packed, self.stage = net, packed, stage
        self.feats = [torch.empty_like(f) for f in feats]
        self.p3d = p3d_normed
        self.bbox = None if bbox is None else torch.empty_like(bbox)
        self.obj_ids = None if obj_ids is None else torch.empty_like(obj_ids)
        self._load(feats, bbox, obj_ids)
        side = torch.cuda.Stream(device=feats[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):      # builds the kNN graphs / plans / packed weights and lets the libraries pick algorithms
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self._forward()

    def _forward(self):
        if self.bbox is None:
            args = (self.feats, self.p3d) + (() if self.obj_ids is None else (self.obj_ids,))
            return self.net(*args, stage=self.stage), None
        return self.net.forward_with_correspondences(self.feats, self.p3d, self.bbox, stage=self.stage, obj_ids=self.obj_ids,
                                                     packed=self.packed)

    def _load(self, feats, bbox, obj_ids):
        for d, s in zip(self.feats, feats):
            if d.data_ptr() != s.data_ptr():
                d.copy_(s, non_blocking=True)
        if self.bbox is not None and bbox is not None and bbox.data_ptr() != self.bbox.data_ptr():
            self.bbox.copy_(bbox, non_blocking=True)
        if self.obj_ids is not None and obj_ids is not None and obj_ids.data_ptr() != self.obj_ids.data_ptr():
            self.obj_ids.copy_(obj_ids, non_blocking=True)

    @property
    def inputs(self):
        """The static input buffers: fill them directly (e.g. as the destination of the H2D copy) to skip the staging copy."""
        return self.feats

    def __call__(self, feats=None, bbox=None, obj_ids=None):
        if feats is not None:
            self._load(feats, bbox, obj_ids)
        self.graph.replay()
        return self.out

"""Synthetic inputs and weights for the GNN keypoint head.

There is no network in the build/bench environment, so trained checkpoints and BOP images are
not available.  Everything the parity tests and ``bench.py`` feed to the head is generated here
from a seed, on the CPU, with plain torch RNG calls (identical in this container and on the GPU
box because both run the same image):

* keypoint clouds: the reference's own shipped FPS files, packed once into
  ``checkerpose_b200/data/fps_202212.npz`` by ``checkerpose_b200/data/make_fps_fixture.py`` (input data of the
  path, not test infrastructure: nothing under ``checkerpose_b200/`` reads ``tests/``)
  (reference loader: ``checkerpose/test.py:145-148``);
* ``pc_normalize``: ``checkerpose/aux_utils/pointnet2_utils.py:11-20``;
* HRNet-W18 shaped feature maps (channel table ``checkerpose/model/pipeline.py:12``);
* a variance-preserving ``state_dict`` with exactly the reference's key names and shapes
  (``checkerpose/model/init.py:71-107``, ``checkerpose/model/pipeline.py:214-349``).  PyTorch's
  default init collapses the refine-stage logits, which would make a
  parity test on the decoded codes degenerate, hence He-normal weights and randomised BatchNorm
  statistics including negative gammas.

Nothing in this file touches ``oracle/`` or the CUDA library.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np
import torch

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FPS_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "fps_202212.npz")

HRNET_W18_DIMS = (128, 256, 512, 1024)   # checkerpose/model/pipeline.py:12
HRNET_W18_SIZES = (64, 32, 16, 8)        # for a 256x256 RoI crop
FPS_OBJECTS = {
    "lm": tuple(range(1, 16)),
    "lmo": (1, 5, 6, 8, 9, 10, 11, 12),
    "ycbv": tuple(range(1, 22)),
}

_fps_cache = None


def load_fps_xyz(dataset: str, obj_id: int, num_p3d: int = 4096) -> np.ndarray:
    """First ``num_p3d`` FPS keypoints (mm, float32->float64) of a BOP object (1-based id)."""
    global _fps_cache
    if _fps_cache is None:
        _fps_cache = np.load(FPS_FIXTURE)
    xyz = _fps_cache[f"{dataset}/{obj_id}"]
    return np.asarray(xyz[:num_p3d], dtype=np.float64)


def pc_normalize(pc: np.ndarray, return_stat: bool = False):
    """Centre and scale a cloud into the unit sphere (pointnet2_utils.py:11-20)."""
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    pc = pc / m
    if return_stat:
        return pc, centroid, m
    return pc


def p3d_normed_tensor(xyz: np.ndarray) -> torch.Tensor:
    """(N,3) float64 mm -> (1,3,N) float32 normalised, as ``test.py:151-155`` does."""
    pn = pc_normalize(np.array(xyz, dtype=np.float64, copy=True))
    return torch.as_tensor(pn, dtype=torch.float32).transpose(1, 0).unsqueeze(0).contiguous()


def synthetic_features(batch: int, gen: torch.Generator) -> list:
    """Four post-ReLU-like HRNet-W18 feature maps, NCHW float32 on the CPU."""
    feats = []
    for c, s in zip(HRNET_W18_DIMS, HRNET_W18_SIZES):
        feats.append(torch.relu(torch.randn(batch, c, s, s, generator=gen)))
    return feats


def synthetic_bboxes(batch: int, gen: torch.Generator) -> torch.Tensor:
    """(B,4) float32 [x, y, w, h]: x,y ~ U[0,400) ints, w = h ~ U[64,256] ints."""
    xy = torch.randint(0, 400, (batch, 2), generator=gen).float()
    wh = torch.randint(64, 257, (batch, 1), generator=gen).float().expand(-1, 2)
    return torch.cat([xy, wh], dim=1).contiguous()


# --------------------------------------------------------------------------------------------
# state_dict layout
# --------------------------------------------------------------------------------------------
def _bn_keys(prefix, c):
    return [(prefix + ".weight", (c,), "bn_gamma"), (prefix + ".bias", (c,), "bn_beta"),
            (prefix + ".running_mean", (c,), "bn_mean"), (prefix + ".running_var", (c,), "bn_var"),
            (prefix + ".num_batches_tracked", (), "bn_count")]


def head_param_spec(npoint: int, res_log2: int = 6, num_filters: int = 256,
                    init_num_graph_module: int = 2, num_graph_module: int = 3,
                    local_k: int = 2, query_dims=None, seg_output_dim: int = 2,
                    include_refine: bool = True, prefix_init: str = "init_net."):
    """[(key, shape, kind, fan_in)] for ``PoseNet_GNNskip(InitNet_GNN(...))`` without a backbone.

    With ``include_refine=False`` and ``prefix_init=""`` this is a bare ``InitNet_GNN``.
    """
    qd = (num_filters, 256, 64) if query_dims is None else tuple(query_dims)
    spec = []
    p = prefix_init
    spec.append((p + "conv1x1.weight", (npoint, 1024, 1, 1), "w", 1024))
    spec.append((p + "conv1x1.bias", (npoint,), "b", None))
    for i in range(init_num_graph_module):
        spec.append((p + f"pre_query_block.{i}.conv.0.weight", (64, 128, 1, 1), "w", 128))
        spec += [(k, s, kind, None) for k, s, kind in _bn_keys(p + f"pre_query_block.{i}.conv.1", 64)]
    spec.append((p + "mlp.weight", (7, 64), "w", 64))
    spec.append((p + "mlp.bias", (7,), "b", None))
    if not include_refine:
        return spec
    nref = res_log2 - 3
    for i in range(nref):
        u = f"up_net.{i}."
        if i == 0:
            spec.append((u + "0.weight", (1024, num_filters, 3, 3), "w", 1024 * 9 // 4))
            spec += [(k, s, kind, None) for k, s, kind in _bn_keys(u + "1", num_filters)]
            spec.append((u + "3.weight", (num_filters, num_filters, 3, 3), "w", num_filters * 9))
            spec += [(k, s, kind, None) for k, s, kind in _bn_keys(u + "4", num_filters)]
            spec.append((u + "6.weight", (num_filters, num_filters, 3, 3), "w", num_filters * 9))
            spec += [(k, s, kind, None) for k, s, kind in _bn_keys(u + "7", num_filters)]
        else:
            cin = num_filters + HRNET_W18_DIMS[-i - 1]
            spec.append((u + "1.weight", (num_filters, cin, 3, 3), "w", cin * 9))
            spec += [(k, s, kind, None) for k, s, kind in _bn_keys(u + "2", num_filters)]
            spec.append((u + "4.weight", (num_filters, num_filters, 3, 3), "w", num_filters * 9))
            spec += [(k, s, kind, None) for k, s, kind in _bn_keys(u + "5", num_filters)]
    for i in range(nref):
        r = f"refine_net.{i}."
        gdim = 64 if i == 0 else qd[0]
        emb = qd[0] // 4
        spec.append((r + "local_feat_ext_block.patch_generator.weight",
                     (emb, num_filters, local_k, local_k), "w", num_filters * local_k * local_k))
        spec.append((r + "local_feat_ext_block.patch_generator.bias", (emb,), "b", None))
        spec.append((r + "pre_graph_module.0.weight", (qd[0], qd[0] + gdim), "w", qd[0] + gdim))
        spec.append((r + "pre_graph_module.0.bias", (qd[0],), "b", None))
        spec.append((r + "pre_graph_module.2.weight", (qd[0], qd[0]), "w", qd[0]))
        spec.append((r + "pre_graph_module.2.bias", (qd[0],), "b", None))
        ngm = num_graph_module if isinstance(num_graph_module, int) else num_graph_module[i]
        for j in range(ngm):
            spec.append((r + f"pre_query_block.{j}.conv.0.weight", (qd[0], 2 * qd[0], 1, 1), "w", 2 * qd[0]))
            spec += [(k, s, kind, None) for k, s, kind in _bn_keys(r + f"pre_query_block.{j}.conv.1", qd[0])]
        dims = qd + (2,)
        for j in range(1, len(dims)):
            spec.append((r + f"query_block.mlps.{2 * (j - 1)}.weight", (dims[j], dims[j - 1]), "w", dims[j - 1]))
            spec.append((r + f"query_block.mlps.{2 * (j - 1)}.bias", (dims[j],), "b", None))
    spec.append(("seg_block.weight", (seg_output_dim, num_filters, 1, 1), "w", num_filters))
    spec.append(("seg_block.bias", (seg_output_dim,), "b", None))
    return spec


def abwoprog_param_spec(npoint: int, res_log2: int = 6, num_filters: int = 256, init_num_graph_module: int = 2,
                        num_graph_module: int = 3, query_dims=None, seg_output_dim: int = 2):
    """Spec of ``PoseNet_GNNskip_ABwoProg(InitNet_GNN(...))`` (pipeline_lm.py:430-517): the refine stages only refine the
    graph feature (no image sampling, no per-stage query); ONE query MLP emits all 2*res_log2+1 bits at the end."""
    qd = (num_filters, 256, 64) if query_dims is None else tuple(query_dims)
    full = head_param_spec(npoint, res_log2, num_filters, init_num_graph_module, num_graph_module, 2, query_dims, seg_output_dim)
    spec = [e for e in full if not e[0].startswith("refine_net.") and not e[0].startswith("seg_block.")]
    for i in range(res_log2 - 3):
        r = f"refine_net.{i}."
        gdim = 64 if i == 0 else qd[0]
        spec.append((r + "pre_graph_module.0.weight", (qd[0], gdim), "w", gdim))
        spec.append((r + "pre_graph_module.0.bias", (qd[0],), "b", None))
        spec.append((r + "pre_graph_module.2.weight", (qd[0], qd[0]), "w", qd[0]))
        spec.append((r + "pre_graph_module.2.bias", (qd[0],), "b", None))
        ngm = num_graph_module if isinstance(num_graph_module, int) else num_graph_module[i]
        for j in range(ngm):
            spec.append((r + f"pre_query_block.{j}.conv.0.weight", (qd[0], 2 * qd[0], 1, 1), "w", 2 * qd[0]))
            spec += [(k, sh, kind, None) for k, sh, kind in _bn_keys(r + f"pre_query_block.{j}.conv.1", qd[0])]
    spec.append(("seg_block.weight", (seg_output_dim, num_filters, 1, 1), "w", num_filters))
    spec.append(("seg_block.bias", (seg_output_dim,), "b", None))
    dims = qd + (2 * res_log2 + 1,)
    for j in range(1, len(dims)):
        spec.append((f"query_block.mlps.{2 * (j - 1)}.weight", (dims[j], dims[j - 1]), "w", dims[j - 1]))
        spec.append((f"query_block.mlps.{2 * (j - 1)}.bias", (dims[j],), "b", None))
    return spec


def synthetic_state_dict(spec, gen: torch.Generator) -> "OrderedDict[str, torch.Tensor]":
    """He-normal weights, small biases, randomised BN stats with 25 % negative gammas."""
    sd = OrderedDict()
    for key, shape, kind, fan_in in spec:
        if kind == "w":
            t = torch.randn(*shape, generator=gen) * float(np.sqrt(2.0 / fan_in))
        elif kind == "b":
            t = torch.randn(*shape, generator=gen) * 0.1
        elif kind == "bn_gamma":
            t = 0.5 + torch.rand(*shape, generator=gen)
            flip = torch.rand(*shape, generator=gen) < 0.25
            t = torch.where(flip, -t, t)
        elif kind in ("bn_beta", "bn_mean"):
            t = torch.randn(*shape, generator=gen) * 0.1
        elif kind == "bn_var":
            t = 0.5 + torch.rand(*shape, generator=gen)
        elif kind == "bn_count":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise ValueError(kind)
        sd[key] = t
    return sd


def tensor_checksum(t: torch.Tensor) -> float:
    """Cheap content fingerprint used by the golden files to detect RNG drift."""
    x = t.detach().double().flatten()
    w = torch.arange(1, x.numel() + 1, dtype=torch.float64) % 9973
    return float((x * w).sum())

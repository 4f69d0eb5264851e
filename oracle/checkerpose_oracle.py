"""CPU oracle for the CheckerPose GNN keypoint head  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file restates, as plain functions over CPU tensors, the algorithm of the reference's
post-backbone head (``/root/reference/checkerpose/model/{init,init_lm,pipeline,pipeline_lm}.py``
plus the vectorisable first half of ``from_id_to_pose``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it, and only as the checker / CPU baseline.  The product path
(``checkerpose_b200/``) never imports it and has no CPU fallback.

Parity is PINNED: ``tests/golden/make_golden.py`` imports the unmodified reference modules in the
build container (timm stubbed -- the backbone is outside the path) and stores their outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below against those
vectors.

Why torch on the CPU rather than numpy/C: the reference *is* PyTorch, its arithmetic on this path
is float32 conv / matmul / index / max, and the oracle doubles as the reference-arm CPU baseline,
which should use the same BLAS-backed primitives the reference would.  The formulation is the
reference's own per-edge one (gather -> [x_j - x_i ; x_i] -> 1x1 conv -> BN -> LeakyReLU -> max),
NOT the factored form the CUDA kernels use, so agreement between the two is a real check.

Every function takes ``dtype`` from its inputs: pass float64 tensors to get a higher-precision
reference for margin analysis.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, used by every BN on the path


# ------------------------------------------------------------------------------------------
# K1: kNN graph  (pipeline.py:18-23 == init.py:27-32 == pipeline_lm.py:18-23 == init_lm.py:27-32)
# ------------------------------------------------------------------------------------------
def pairwise_neg_sqdist(x: torch.Tensor) -> torch.Tensor:
    """x (B,C,N) -> (B,N,N) negative squared distances, reference operation order."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)                 # pipeline.py:19
    xx = torch.sum(x ** 2, dim=1, keepdim=True)                      # pipeline.py:20
    return -xx - inner - xx.transpose(2, 1)                          # pipeline.py:21


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """x (B,C,N) -> idx (B,N,k) int64 (pipeline.py:22)."""
    return pairwise_neg_sqdist(x).topk(k=k, dim=-1)[1]


# ------------------------------------------------------------------------------------------
# K2: EdgeConv   (pipeline.py:27-40, 45-59; LM per-sample graph: pipeline_lm.py:55-60)
# ------------------------------------------------------------------------------------------
def get_graph_feature(x: torch.Tensor, knn_idx: torch.Tensor) -> torch.Tensor:
    """x (B,C,N), knn_idx (1|B,N,K) -> (B,2C,N,K) = cat([x_j - x_i, x_i]) (pipeline.py:27-40)."""
    B, C, N = x.shape
    K = knn_idx.shape[2]
    idx = knn_idx.expand(B, -1, -1) if knn_idx.shape[0] == 1 else knn_idx
    flat = idx.reshape(B, 1, N * K).expand(-1, C, -1)
    nbr = torch.gather(x, 2, flat).view(B, C, N, K)                   # x[b, :, idx[b,n,k]]
    ctr = x.unsqueeze(3).expand(-1, -1, -1, K)
    return torch.cat([nbr - ctr, ctr], dim=1)


def static_graph_module(x, knn_idx, conv_w, bn_w, bn_b, bn_mean, bn_var, leaky_slope=0.2):
    """StaticGraph_module.forward in eval mode (pipeline.py:55-59)."""
    f = get_graph_feature(x, knn_idx)
    y = F.conv2d(f, conv_w)                                           # 1x1, bias=False
    y = F.batch_norm(y, bn_mean, bn_var, bn_w, bn_b, training=False, eps=BN_EPS)
    y = F.leaky_relu(y, leaky_slope)
    return y.max(dim=-1)[0]


def _sg_from_sd(x, knn_idx, sd, prefix, slope):
    return static_graph_module(x, knn_idx, sd[prefix + "conv.0.weight"], sd[prefix + "conv.1.weight"],
                               sd[prefix + "conv.1.bias"], sd[prefix + "conv.1.running_mean"],
                               sd[prefix + "conv.1.running_var"], slope)


def mlp_leaky(x, sd, prefix, num_layers, slope, last_act):
    """get_MLP_leakyReLU_layers (pipeline.py:61-69): Linear at indices 0,2,4,... of the Sequential."""
    for j in range(num_layers):
        x = F.linear(x, sd[f"{prefix}{2 * j}.weight"], sd[f"{prefix}{2 * j}.bias"])
        if j < num_layers - 1 or last_act:
            x = F.leaky_relu(x, slope)
    return x


# ------------------------------------------------------------------------------------------
# K4: sign-bit decode   (pipeline.py:72-127)
# ------------------------------------------------------------------------------------------
def from_code_to_id(code, class_base=2):
    """(B,L,N) integer bits -> (B,N), MSB first (pipeline.py:72-82)."""
    L = code.shape[1]
    ids = code[:, 0, :] * (class_base ** (L - 1))
    for i in range(1, L):
        ids = ids + code[:, i, :] * (class_base ** (L - 1 - i))
    return ids


def from_code_prob_to_id(code_prob, class_base=2):
    code = torch.where(torch.sigmoid(code_prob) > 0.5, 1, 0)         # pipeline.py:89-90
    return from_code_to_id(code, class_base)


def from_gt_code_to_id(gt_code, class_base=2):
    return from_code_to_id(torch.where(gt_code > 0.5, 1, 0), class_base)   # pipeline.py:94-101


def from_bit_prob_to_id(bit_prob):
    return torch.where(torch.sigmoid(bit_prob[:, 0, :]) > 0.5, 1, 0)  # pipeline.py:103-110


def from_gt_bit_to_id(gt_bit):
    return torch.where(gt_bit[:, 0, :] > 0.5, 1, 0)                   # pipeline.py:112-118


def from_mask_prob_to_mask(mask_prob):
    return torch.where(torch.sigmoid(mask_prob) > 0.5, 1.0, 0.0).to(mask_prob.dtype)  # :120-127


# ------------------------------------------------------------------------------------------
# K3: Index2Feat 4-tap integer gather   (pipeline.py:130-164)
# ------------------------------------------------------------------------------------------
def index2feat(img_feat_highres, pg_w, pg_b, pixel_x_id, pixel_y_id, kernel_size):
    """-> (B, 4*embed, N); tap order (2y,2x),(2y+k,2x),(2y,2x+k),(2y+k,2x+k) (pipeline.py:156-163)."""
    patches = F.conv2d(img_feat_highres, pg_w, pg_b, stride=1, padding=kernel_size - 1)
    B = patches.shape[0]
    bi = torch.arange(B).view(B, 1).expand(-1, pixel_x_id.shape[1])
    k = kernel_size
    sf1 = patches[bi, :, 2 * pixel_y_id, 2 * pixel_x_id]
    sf2 = patches[bi, :, 2 * pixel_y_id + k, 2 * pixel_x_id]
    sf3 = patches[bi, :, 2 * pixel_y_id, 2 * pixel_x_id + k]
    sf4 = patches[bi, :, 2 * pixel_y_id + k, 2 * pixel_x_id + k]
    return torch.cat([sf1, sf2, sf3, sf4], dim=2).permute(0, 2, 1)


# ------------------------------------------------------------------------------------------
# image branch (library part of the path; pipeline.py:183-211)
# ------------------------------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=BN_EPS)


def upsample_module(x, sd, prefix, is_convtrans):
    if is_convtrans:
        x = F.conv_transpose2d(x, sd[prefix + "0.weight"], stride=2, padding=1, output_padding=1)
        x = F.relu(_bn(x, sd, prefix + "1"))
        x = F.conv2d(x, sd[prefix + "3.weight"], padding=1)
        x = F.relu(_bn(x, sd, prefix + "4"))
        x = F.conv2d(x, sd[prefix + "6.weight"], padding=1)
        x = F.relu(_bn(x, sd, prefix + "7"))
    else:
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)  # UpsamplingBilinear2d
        x = F.conv2d(x, sd[prefix + "1.weight"], padding=1)
        x = F.relu(_bn(x, sd, prefix + "2"))
        x = F.conv2d(x, sd[prefix + "4.weight"], padding=1)
        x = F.relu(_bn(x, sd, prefix + "5"))
    return x


# ------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------
def _select_graph(knn_idx, obj_ids):
    """LM variant indexes a (num_obj,N,K) table with 1-based ids (pipeline_lm.py:56-57)."""
    return knn_idx if obj_ids is None else knn_idx[obj_ids - 1]


def init_head(feat_last, sd, knn_idx, npoint, num_graph_module=2, graph_slope=0.2, obj_ids=None,
              prefix="init_net."):
    """InitNet_GNN.forward after the backbone (init.py:112-122).  -> logits (B,7,N), graph (B,64,N)."""
    out = F.conv2d(feat_last, sd[prefix + "conv1x1.weight"], sd[prefix + "conv1x1.bias"])
    g = out.reshape(-1, npoint, 64).permute(0, 2, 1)                  # init.py:114
    idx = _select_graph(knn_idx, obj_ids)
    for i in range(num_graph_module):
        g = _sg_from_sd(g, idx, sd, f"{prefix}pre_query_block.{i}.", graph_slope)
    o = F.linear(g.permute(0, 2, 1), sd[prefix + "mlp.weight"], sd[prefix + "mlp.bias"]).permute(0, 2, 1)
    return o, g


def refine_module(img_feat, graph_feat, roi_mask_bit, prev_x_id, prev_y_id, sd, prefix, knn_idx,
                  local_k=2, leaky_slope=0.01, num_graph_module=3, graph_slope=0.2, obj_ids=None):
    """Refine_moduleGNN.forward (pipeline.py:262-298).  -> new bits (B,2,N), graph feat (B,C,N)."""
    lf = index2feat(img_feat, sd[prefix + "local_feat_ext_block.patch_generator.weight"],
                    sd[prefix + "local_feat_ext_block.patch_generator.bias"], prev_x_id, prev_y_id, local_k)
    lf = lf * roi_mask_bit                                            # pipeline.py:280
    lf = torch.cat([lf, graph_feat], dim=1).permute(0, 2, 1)          # pipeline.py:283-284
    lf = mlp_leaky(lf, sd, prefix + "pre_graph_module.", 2, leaky_slope, last_act=True).permute(0, 2, 1)
    idx = _select_graph(knn_idx, obj_ids)
    for j in range(num_graph_module):
        lf = _sg_from_sd(lf, idx, sd, f"{prefix}pre_query_block.{j}.", graph_slope)
    bits = mlp_leaky(lf.permute(0, 2, 1), sd, prefix + "query_block.mlps.", 3, leaky_slope,
                     last_act=False).permute(0, 2, 1)                 # MLP_QueryNet ignores pts (:174-180)
    return bits, lf


def pose_head(img_feats, sd, init_knn_idx, refine_knn_idx, npoint, res_log2=6, local_k=2,
              leaky_slope=0.01, init_num_graph_module=2, num_graph_module=3, init_graph_slope=0.2,
              graph_slope=0.2, obj_ids=None, stage=None, return_intermediates=False):
    """PoseNet_GNNskip.forward after the backbone (pipeline.py:351-384; LM: pipeline_lm.py:392-425).

    img_feats: the four backbone maps [(B,128,64,64),(B,256,32,32),(B,512,16,16),(B,1024,8,8)].
    refine_knn_idx: one index table per refine stage (they are equal when graph_k matches).
    Returns (roi_bit (B,1,N), x_bits (B,L,N), y_bits (B,L,N), seg (B,2,H,W), x_id (B,N), y_id (B,N)).
    """
    nref = res_log2 - 3
    nact = nref if stage is None else stage
    bits, graph_feat = init_head(img_feats[-1], sd, init_knn_idx, npoint, init_num_graph_module,
                                 init_graph_slope, obj_ids)
    img_feat = img_feats[-1]
    roi_bit, x_bits, y_bits = bits[:, 0:1], bits[:, 1:4], bits[:, 4:]
    roi_mask = from_mask_prob_to_mask(roi_bit)
    x_id = from_code_prob_to_id(x_bits)
    y_id = from_code_prob_to_id(y_bits)
    inter = {"graph_feat": [graph_feat], "img_feat": []}
    for i in range(nact):
        if i > 0:
            img_feat = torch.cat([img_feat, img_feats[-i - 1]], dim=1)
        img_feat = upsample_module(img_feat, sd, f"up_net.{i}.", is_convtrans=(i == 0))
        ngm = num_graph_module if isinstance(num_graph_module, int) else num_graph_module[i]
        new_bits, graph_feat = refine_module(img_feat, graph_feat, roi_mask, x_id, y_id, sd,
                                             f"refine_net.{i}.", refine_knn_idx[i], local_k, leaky_slope,
                                             ngm, graph_slope, obj_ids)
        inter["graph_feat"].append(graph_feat)
        inter["img_feat"].append(img_feat)
        x_bits = torch.cat([x_bits, new_bits[:, 0:1]], dim=1)
        y_bits = torch.cat([y_bits, new_bits[:, 1:2]], dim=1)
        x_id = x_id * 2 + from_bit_prob_to_id(new_bits[:, 0:1])       # pipeline.py:380-381
        y_id = y_id * 2 + from_bit_prob_to_id(new_bits[:, 1:2])
    seg = F.conv2d(img_feat, sd["seg_block.weight"], sd["seg_block.bias"])
    out = (roi_bit, x_bits, y_bits, seg, x_id, y_id)
    return (out, inter) if return_intermediates else out


# ------------------------------------------------------------------------------------------
# correspondences: first half of from_id_to_pose (test_network_with_test_data.py:50-66) with the
# RoI grid of bop_dataset_pytorch.py:266-269,223-235,359.
# ------------------------------------------------------------------------------------------
def pose_head_abwoprog(img_feats, sd, init_knn_idx, refine_knn_idx, npoint, res_log2=6, leaky_slope=0.01,
                       init_num_graph_module=2, num_graph_module=3, init_graph_slope=0.2, graph_slope=0.2,
                       obj_ids=None, stage=None):
    """PoseNet_GNNskip_ABwoProg.forward after the backbone (pipeline_lm.py:484-517) with
    Refine_moduleGNN_ABwoProg.forward (pipeline_lm.py:324-339): the stages refine the graph feature only, one
    MLP_QueryNet emits all 2*res_log2+1 logits at the end, ids are the MSB-first decode of the thresholded bits."""
    nact = (res_log2 - 3) if stage is None else stage
    _, graph_feat = init_head(img_feats[-1], sd, init_knn_idx, npoint, init_num_graph_module, init_graph_slope, obj_ids)
    img_feat = img_feats[-1]
    for i in range(nact):
        if i > 0:
            img_feat = torch.cat([img_feat, img_feats[-i - 1]], dim=1)                      # :498
        img_feat = upsample_module(img_feat, sd, f"up_net.{i}.", is_convtrans=(i == 0))     # :499
        prefix = f"refine_net.{i}."
        lf = mlp_leaky(graph_feat.permute(0, 2, 1), sd, prefix + "pre_graph_module.", 2, leaky_slope,
                       last_act=True).permute(0, 2, 1)                                      # :332-334
        idx = _select_graph(refine_knn_idx[i], obj_ids)
        ngm = num_graph_module if isinstance(num_graph_module, int) else num_graph_module[i]
        for j in range(ngm):
            lf = _sg_from_sd(lf, idx, sd, f"{prefix}pre_query_block.{j}.", graph_slope)     # :337-338
        graph_feat = lf
    seg = F.conv2d(img_feat, sd["seg_block.weight"], sd["seg_block.bias"])                  # :502
    bits = mlp_leaky(graph_feat.permute(0, 2, 1), sd, "query_block.mlps.", 3, leaky_slope, last_act=False).permute(0, 2, 1)
    roi_bit, x_bits, y_bits = bits[:, 0:1], bits[:, 1:res_log2 + 1], bits[:, res_log2 + 1:]  # :507-509
    return roi_bit, x_bits, y_bits, seg, from_code_prob_to_id(x_bits), from_code_prob_to_id(y_bits)


def farthest_point_sample_init_center(xyz, npoint):
    """preprocess_data/get_fps_points.py:65-90, restated line by line (NumPy float64)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    num_xyz = xyz.shape[0]
    xyz_max = xyz.max(axis=0)
    xyz_min = xyz.min(axis=0)
    farthest_xyz = (xyz_max + xyz_min) / 2
    xyz_extent = np.linalg.norm(xyz_max - xyz_min)
    fps_xyz = np.zeros((npoint, 3))
    fps_ids = []
    distances_to_set = np.ones(num_xyz) * xyz_extent * 10
    for sample_id in range(npoint):
        distances = np.linalg.norm((xyz - farthest_xyz), axis=1)
        mask = distances < distances_to_set
        distances_to_set[mask] = distances[mask]
        farthest_id = int(np.argmax(distances_to_set))
        farthest_xyz = xyz[farthest_id, :]
        fps_ids.append(farthest_id)
        fps_xyz[sample_id, :] = farthest_xyz
    return fps_ids, fps_xyz


def roi_xy_ori(bbox, size):
    """bbox (4,) [x,y,w,h] -> (size,size,2) grid: (x + u*w/size, y + v*h/size)."""
    bbox = np.asarray(bbox, dtype=np.float64)
    u = np.linspace(0, size - 1, size)
    xy = np.asarray(np.meshgrid(u, u)).transpose(1, 2, 0)             # (h,w,2): [...,0]=u, [...,1]=v
    out = np.zeros_like(xy)
    out[:, :, 0] = (bbox[2] / size) * xy[:, :, 0] + bbox[0]
    out[:, :, 1] = (bbox[3] / size) * xy[:, :, 1] + bbox[1]
    return out


def id_to_correspondences(roi_logit, seg_logit, x_id, y_id, bbox):
    """One RoI.  roi_logit (N,), seg_logit (2,H,W) [visib, full] (test.py:313-314), ids (N,) int.

    Returns p2d (N,2) float64 and three boolean masks: all / full-mask / visib-mask, i.e. the
    ``valid_mask`` of the three ``from_id_to_pose`` calls in test.py:335-368.
    """
    roi_logit = np.asarray(roi_logit, dtype=np.float64)
    seg_logit = np.asarray(seg_logit, dtype=np.float64)
    size = seg_logit.shape[-1]
    grid = roi_xy_ori(bbox, size)
    p2d = grid[y_id, x_id]
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    valid_all = sig(roi_logit) > 0.5
    seg = sig(seg_logit) > 0.5
    valid_visib = np.logical_and(valid_all, seg[0][y_id, x_id])
    valid_full = np.logical_and(valid_all, seg[1][y_id, x_id])
    return p2d, valid_all, valid_full, valid_visib


# ------------------------------------------------------------------------------------------
# helpers named in north_star: binary_code_helper/class_id_encoder_decoder.py, common_ops.py
# ------------------------------------------------------------------------------------------
def class_code_vecs_to_class_id_vec(class_code_vecs, class_base=2):           # :30-38
    out = np.zeros(class_code_vecs.shape[0])
    L = class_code_vecs.shape[1]
    for i in range(L):
        out = out + class_code_vecs[:, i] * (class_base ** (L - 1 - i))
    return out


def class_code_images_to_class_id_image(class_code_images, class_base=2):     # :17-28  (H,W,C)
    out = np.zeros(class_code_images.shape[:2])
    L = class_code_images.shape[2]
    for i in range(L):
        out = out + class_code_images[:, :, i] * (class_base ** (L - 1 - i))
    return out


def class_code_images_to_class_id_image_torch(class_code_images, class_base=2):  # :40-52 (C,H,W)
    L = class_code_images.shape[0]
    out = torch.zeros(class_code_images.shape[1:], dtype=torch.float32)
    for i in range(L):
        out = out + class_code_images[i] * (class_base ** (L - 1 - i))
    return out


def class_code_images_to_class_id_image_torch_batch(class_code_images, class_base=2):  # :54-63
    B, L, H, W = class_code_images.shape
    out = torch.zeros((B, H, W), dtype=torch.float32)
    for i in range(L):
        out = out + class_code_images[:, i] * (class_base ** (L - 1 - i))
    return out.long()


def class_id_vec_to_class_code_vecs(class_id_vec, class_base=2, iteration=8):  # :88-101
    iteration = int(iteration)
    out = np.zeros((len(class_id_vec), iteration))
    v = class_id_vec.astype(int)
    step = math.log2(class_base)
    for i in range(iteration):
        s1 = np.right_shift(v, int(step * (iteration - i - 1)))
        s2 = np.right_shift(v, int(step * (iteration - i)))
        out[:, i] = s1 - s2 * (2 ** step)
    return out


def code_to_id(class_code, class_base=2):                                      # :104-114
    n = len(class_code)
    return sum(class_code[i] * (class_base ** (n - 1 - i)) for i in range(n))


def str_code_to_id(s, class_base=2):                                           # :116-127
    n = len(s)
    return sum(int(s[i]) * (class_base ** (n - 1 - i)) for i in range(n))


def from_output_to_class_mask(pred_mask_prob, thershold=0.5):                  # common_ops.py:5-11
    p = torch.sigmoid(pred_mask_prob).detach().cpu().numpy()
    m = np.zeros(p.shape)
    m[p > thershold] = 1.0
    return m


def from_output_to_class_mask_torch(pred_mask_prob, thershold=0.5):            # common_ops.py:14-18
    return torch.where(torch.sigmoid(pred_mask_prob.detach()) > thershold, 1.0, 0.0)


def from_output_to_class_binary_code(pred_code_prob, loss_type, thershold=0.5,
                                     divided_num_each_interation=2, binary_code_length=16):
    """common_ops.py:21-40 (BCE-family branch and CE branch)."""
    if loss_type in ["BCE", "L1", "SSIM", "L1_SSIM"]:
        return from_output_to_class_mask(pred_code_prob, thershold)
    if loss_type == "CE":
        p = pred_code_prob.reshape(-1, divided_num_each_interation, pred_code_prob.shape[2],
                                   pred_code_prob.shape[3])
        p = torch.softmax(p, dim=1).detach().cpu().numpy()
        code = np.expand_dims(np.argmax(p, axis=1), axis=1)
        return code.reshape(-1, binary_code_length, code.shape[2], code.shape[3])
    raise ValueError(loss_type)

"""Generate golden vectors by running the UNMODIFIED reference modules (CPU, this container only).

Run:  python tests/golden/make_golden.py           (needs /root/reference; writes tests/golden/*.npz)

The reference cannot travel to the GPU box, so its outputs on seeded synthetic inputs are frozen
here.  Inputs and weights are NOT stored (the N=512 head has 10.4 M parameters); they are
regenerated from the seed by ``checkerpose_b200.synthetic`` and guarded by checksums stored
beside the outputs, so an RNG drift shows up as a checksum failure rather than a parity failure.

Third-party modules that the reference imports but that are absent here and off the path
(timm = backbone, pytz/mmcv/imgaug/... = data loading / evaluation) are stubbed in ``sys.modules``.
``cv2.solvePnPRansac`` is intercepted to capture the correspondences ``from_id_to_pose`` builds
(test_network_with_test_data.py:50-66) without running PnP, which is outside the path.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REF = "/root/reference/checkerpose"
sys.path.insert(0, REF)
sys.path.insert(0, "/root/reference/bop_toolkit")

from checkerpose_b200 import synthetic as syn  # noqa: E402


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        return None


for _name in ("timm", "pytz", "mmcv", "imgaug", "imgaug.augmenters", "imageio", "pyprogressivex",
              "plyfile", "png", "vispy", "glumpy"):
    try:
        __import__(_name)
    except Exception:
        sys.modules[_name] = _Stub(_name)


class FeatureBackbone(nn.Module):
    """Stands in for timm's HRNet-W18 features_only model: returns the maps it is given."""

    def forward(self, feats):
        return list(feats)


import model.backbone as ref_backbone  # noqa: E402

ref_backbone.get_timm_backbone = lambda **kw: FeatureBackbone()
import model.init as ref_init  # noqa: E402
import model.init_lm as ref_init_lm  # noqa: E402
import model.pipeline as ref_pipe  # noqa: E402
import model.pipeline_lm as ref_pipe_lm  # noqa: E402

ref_init.get_timm_backbone = ref_backbone.get_timm_backbone
ref_init_lm.get_timm_backbone = ref_backbone.get_timm_backbone
import common_ops as ref_common  # noqa: E402
from binary_code_helper import class_id_encoder_decoder as ref_codec  # noqa: E402

torch.set_grad_enabled(False)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ---------------------------------------------------------------------------------------------
def golden_knn():
    out = {}
    for ds, objs in syn.FPS_OBJECTS.items():
        for oid in objs:
            p = syn.p3d_normed_tensor(syn.load_fps_xyz(ds, oid, 512))
            idx = ref_pipe.knn(p, 20)[0]
            out[f"{ds}_{oid}_n512_k20"] = np.sort(idx.numpy(), axis=1).astype(np.int16)
    for ds, oid in (("lmo", 1), ("ycbv", 21), ("lm", 3)):
        for n in (1024, 2048, 4096):
            p = syn.p3d_normed_tensor(syn.load_fps_xyz(ds, oid, n))
            idx = ref_pipe.knn(p, 20)[0]
            out[f"{ds}_{oid}_n{n}_k20"] = np.sort(idx.numpy(), axis=1).astype(np.int16)
    for k in (8, 16, 32, 40):
        p = syn.p3d_normed_tensor(syn.load_fps_xyz("lm", 9, 512))
        out[f"lm_9_n512_k{k}"] = np.sort(ref_pipe.knn(p, k)[0].numpy(), axis=1).astype(np.int16)
    # generic C (not 3), batched, unsorted order kept: first neighbour is self
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 16, 200, generator=g)
    out["rand_c16_n200_k12_x"] = x.numpy()
    out["rand_c16_n200_k12"] = ref_pipe.knn(x, 12).numpy().astype(np.int16)
    save("knn", **out)


def _load_sg(mod, sd, prefix):
    own = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    mod.load_state_dict(own, strict=True)
    mod.eval()


def golden_modules():
    g = torch.Generator().manual_seed(21)
    out = {}
    # StaticGraph_module, shared graph (pipeline.py:45) and per-sample graph (pipeline_lm.py:45)
    for tag, B, C, Co, N, K in (("a", 2, 64, 64, 128, 20), ("b", 3, 256, 256, 64, 8), ("c", 2, 32, 48, 100, 5)):
        p = torch.randn(1, 3, N, generator=g)
        idx = ref_pipe.knn(p, K)
        spec = [("conv.0.weight", (Co, 2 * C, 1, 1), "w", 2 * C)] + \
               [(k, s, kind, None) for k, s, kind in syn._bn_keys("conv.1", Co)]
        sd = syn.synthetic_state_dict(spec, g)
        m = ref_pipe.StaticGraph_module(C, Co, idx, leaky_slope=0.2)
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = torch.randn(B, C, N, generator=g)
        bi = torch.arange(B).view(B, 1).repeat(1, N * K)
        y = m(x, bi)
        for k, v in sd.items():
            out[f"sg_{tag}_{k}"] = v
        out[f"sg_{tag}_idx"] = idx.numpy().astype(np.int32)
        out[f"sg_{tag}_x"] = x
        out[f"sg_{tag}_y"] = y
        # LM flavour: table of 4 graphs, per-sample selection with 1-based ids
        p4 = torch.randn(4, 3, N, generator=g)
        idx4 = ref_pipe_lm.knn(p4, K)
        mlm = ref_pipe_lm.StaticGraph_module(C, Co, idx4, leaky_slope=0.2)
        mlm.load_state_dict(sd, strict=True)
        mlm.eval()
        obj_ids = torch.tensor([(i * 3) % 4 + 1 for i in range(B)])
        out[f"sg_{tag}_lm_idx"] = idx4.numpy().astype(np.int32)
        out[f"sg_{tag}_lm_obj"] = obj_ids
        out[f"sg_{tag}_lm_y"] = mlm(x, bi, obj_ids)
    # get_graph_feature
    x = torch.randn(2, 5, 30, generator=g)
    idx = ref_pipe.knn(torch.randn(1, 3, 30, generator=g), 4)
    out["ggf_x"], out["ggf_idx"] = x, idx.numpy().astype(np.int32)
    out["ggf_y"] = ref_pipe.get_graph_feature(x, idx, torch.arange(2).view(2, 1).repeat(1, 30 * 4))
    # Index2Feat_module
    for tag, H, k, fd, ed, B in (("a", 16, 2, 32, 16, 2), ("b", 16, 2, 256, 64, 1), ("c", 16, 4, 32, 8, 2)):
        m = ref_pipe.Index2Feat_module(feat_dim=fd, embed_dim=ed, kernel_size=k)
        spec = [("patch_generator.weight", (ed, fd, k, k), "w", fd * k * k), ("patch_generator.bias", (ed,), "b", None)]
        sd = syn.synthetic_state_dict(spec, g)
        m.load_state_dict(sd, strict=True)
        N = 50
        feat = torch.relu(torch.randn(B, fd, H, H, generator=g))
        xid = torch.randint(0, H // 2, (B, N), generator=g)
        yid = torch.randint(0, H // 2, (B, N), generator=g)
        bi = torch.arange(B).view(B, 1).repeat(1, N)
        for kk, v in sd.items():
            out[f"i2f_{tag}_{kk}"] = v
        out[f"i2f_{tag}_feat"], out[f"i2f_{tag}_xid"], out[f"i2f_{tag}_yid"] = feat, xid, yid
        out[f"i2f_{tag}_y"] = m(feat, bi, xid, yid)
    # MLP_QueryNet
    m = ref_pipe.MLP_QueryNet(feat_dims=(256, 256, 64), pt_dim=3, out_dim=2, leaky_slope=0.01)
    spec = [("mlps.0.weight", (256, 256), "w", 256), ("mlps.0.bias", (256,), "b", None),
            ("mlps.2.weight", (64, 256), "w", 256), ("mlps.2.bias", (64,), "b", None),
            ("mlps.4.weight", (2, 64), "w", 64), ("mlps.4.bias", (2,), "b", None)]
    sd = syn.synthetic_state_dict(spec, g)
    m.load_state_dict(sd, strict=True)
    x = torch.randn(2, 40, 256, generator=g)
    for kk, v in sd.items():
        out[f"mq_{kk}"] = v
    out["mq_x"] = x
    out["mq_y"] = m(x, torch.randn(2, 40, 3, generator=g))
    save("modules", **out)


def golden_decode():
    g = torch.Generator().manual_seed(31)
    out = {}
    cp = torch.randn(3, 6, 77, generator=g) * 2
    cp[0, 0, :5] = torch.tensor([0.0, 1e-9, -1e-9, 3e-8, 1e-7])   # sigmoid(x)>0.5 edge cases in fp32
    out["code_prob"] = cp
    out["code_prob_id"] = ref_pipe.from_code_prob_to_id(cp)
    out["gt_code_id"] = ref_pipe.from_gt_code_to_id(torch.sigmoid(cp))
    out["bit_prob_id"] = ref_pipe.from_bit_prob_to_id(cp[:, 0:1])
    out["gt_bit_id"] = ref_pipe.from_gt_bit_to_id(torch.sigmoid(cp[:, 0:1]))
    out["mask"] = ref_pipe.from_mask_prob_to_mask(cp)
    code = torch.randint(0, 2, (3, 6, 77), generator=g)
    out["code"] = code
    out["code_id"] = ref_pipe.from_code_to_id(code)
    # common_ops.py
    for thr in (0.5, 0.3, 0.9):
        out[f"co_mask_{thr}"] = ref_common.from_output_to_class_mask(cp, thershold=thr)
        out[f"co_mask_torch_{thr}"] = ref_common.from_output_to_class_mask_torch(cp, thershold=thr)
        out[f"co_code_bce_{thr}"] = ref_common.from_output_to_class_binary_code(cp, "BCE", thershold=thr)
    ce_in = torch.randn(2, 32, 8, 8, generator=g)
    out["co_ce_in"] = ce_in
    out["co_code_ce"] = ref_common.from_output_to_class_binary_code(ce_in, "CE", divided_num_each_interation=2,
                                                                    binary_code_length=16)
    out["co_batch_size"] = np.array(ref_common.get_batch_size(0.75, 32))
    out["co_dim_tuple"] = np.array(ref_common.from_dim_str_to_tuple("256_256_64"))
    # class_id_encoder_decoder.py
    vecs = torch.randint(0, 2, (100, 6), generator=g).numpy().astype(np.float64)
    out["cc_vecs"] = vecs
    out["cc_vecs_id"] = ref_codec.class_code_vecs_to_class_id_vec(vecs)
    imgs = torch.randint(0, 2, (9, 11, 16), generator=g).numpy().astype(np.float64)
    out["cc_hwc"] = imgs
    out["cc_hwc_id"] = ref_codec.class_code_images_to_class_id_image(imgs)
    chw = torch.randint(0, 2, (16, 9, 11), generator=g).float()
    out["cc_chw"] = chw
    out["cc_chw_id"] = ref_codec.class_code_images_to_class_id_image_torch(chw)
    bchw = torch.randint(0, 2, (2, 16, 9, 11), generator=g).float()
    out["cc_bchw"] = bchw
    out["cc_bchw_id"] = ref_codec.class_code_images_to_class_id_image_torch_batch(bchw)
    ids = torch.randint(0, 64, (100,), generator=g).numpy()
    out["cc_ids"] = ids
    out["cc_ids_code"] = ref_codec.class_id_vec_to_class_code_vecs(ids, class_base=2, iteration=6)
    idimg = torch.randint(0, 256, (7, 5), generator=g).numpy()
    out["cc_idimg"] = idimg
    out["cc_idimg_code"] = ref_codec.class_id_image_to_class_code_images(idimg, class_base=2, iteration=8,
                                                                         number_of_class=256)
    out["cc_code_to_id"] = np.array(ref_codec.code_to_id([1, 0, 1, 1, 0]))
    out["cc_str_code_to_id"] = np.array(ref_codec.str_code_to_id("10110"))
    save("decode", **out)


def golden_correspondences():
    import cv2
    import test_network_with_test_data as ref_t
    from bop_dataset_pytorch import mapping_pixel_position_to_original_position_2d as ref_map

    captured = {}

    def fake_pnp(p3d, p2d, K, **kw):
        captured["p3d"], captured["p2d"] = np.array(p3d), np.array(p2d)
        return True, np.zeros((3, 1)), np.zeros((3, 1)), None

    ref_t.cv2.solvePnPRansac = fake_pnp
    g = torch.Generator().manual_seed(41)
    out = {}
    N, S = 300, 64
    xyz = syn.load_fps_xyz("lmo", 1, N)
    roi_x = np.linspace(0, S - 1, S)
    roi_xy = np.asarray(np.meshgrid(roi_x, roi_x)).transpose((1, 2, 0))      # bop_dataset_pytorch.py:266-269
    for c in range(3):
        bbox = syn.synthetic_bboxes(1, g)[0].numpy().astype(np.float64)
        grid = ref_map(roi_xy, bbox, S)
        roi_logit = torch.randn(N, generator=g).numpy()
        seg_logit = torch.randn(2, S, S, generator=g).numpy()
        xid = torch.randint(0, S, (N,), generator=g).numpy()
        yid = torch.randint(0, S, (N,), generator=g).numpy()
        roi_bit = (1 / (1 + np.exp(-roi_logit)) > 0.5).astype(np.float64)[:, None]
        seg = (1 / (1 + np.exp(-seg_logit)) > 0.5).astype(np.float64)
        out[f"c{c}_bbox"], out[f"c{c}_roi_logit"], out[f"c{c}_seg_logit"] = bbox, roi_logit, seg_logit
        out[f"c{c}_xid"], out[f"c{c}_yid"], out[f"c{c}_grid"] = xid, yid, grid
        for tag, kw in (("all", dict(check_seg=False)), ("full", dict(check_seg=True, seg_mask=seg[1])),
                        ("visib", dict(check_seg=True, seg_mask=seg[0]))):
            captured.clear()
            ref_t.from_id_to_pose(p3d_xyz=xyz, roi_xy_ori=grid, cam_K=np.eye(3), roi_mask_bit=roi_bit,
                                  pixel_x_id=xid, pixel_y_id=yid, use_progressivex=False, **kw)
            out[f"c{c}_{tag}_p3d"], out[f"c{c}_{tag}_p2d"] = captured["p3d"], captured["p2d"]
    save("correspondences", **out)


def _build_ref_head(npoint, p3d, lm, max_b=8):
    if lm:
        init = ref_init_lm.InitNet_GNN(npoint=npoint, p3d_normed=p3d, res_log2=3, backbone_name="hrnet_w18",
                                       pretrain_backbone=False, num_conv1x1=1, max_batch_size=max_b,
                                       num_graph_module=2, graph_k=20, graph_leaky_slope=0.2)
        net = ref_pipe_lm.PoseNet_GNNskip(init, npoint=npoint, p3d_normed=p3d, res_log2=6, num_filters=256,
                                          max_batch_size=max_b, query_dims=None, local_k=2, leaky_slope=0.01,
                                          num_graph_module=3, graph_k=20, graph_leaky_slope=0.2, query_type="mlp")
    else:
        init = ref_init.InitNet_GNN(npoint=npoint, p3d_normed=p3d, res_log2=3, backbone_name="hrnet_w18",
                                    pretrain_backbone=False, num_conv1x1=1, max_batch_size=max_b,
                                    num_graph_module=2, graph_k=20, graph_leaky_slope=0.2)
        net = ref_pipe.PoseNet_GNNskip(init, npoint=npoint, p3d_normed=p3d, res_log2=6, num_filters=256,
                                       max_batch_size=max_b, query_dims=None, local_k=2, leaky_slope=0.01,
                                       num_graph_module=3, graph_k=20, graph_leaky_slope=0.2, query_type="mlp")
    return net


HEAD_CASES = {
    # name: (dataset, objects, N, B, seed, lm)
    "head_lmo_ape_n512_b1": ("lmo", (1,), 512, 1, 1234 + 0, False),      # BASELINE.json configs[0]
    "head_ycbv21_n128_b2": ("ycbv", (21,), 128, 2, 1234 + 5, False),
    "head_lm15_n128_b3": ("lm", tuple(range(1, 16)), 128, 3, 1234 + 4, True),
}


def head_case_inputs(name):
    ds, objs, N, B, seed, lm = HEAD_CASES[name]
    g = torch.Generator().manual_seed(seed)
    p3d = torch.cat([syn.p3d_normed_tensor(syn.load_fps_xyz(ds, o, N)) for o in objs], dim=0)
    sd = syn.synthetic_state_dict(syn.head_param_spec(N), g)
    feats = syn.synthetic_features(B, g)
    obj_ids = torch.tensor([objs[(i * 7) % len(objs)] for i in range(B)]) if lm else None
    return p3d, sd, feats, obj_ids


def golden_heads():
    for name, (ds, objs, N, B, seed, lm) in HEAD_CASES.items():
        p3d, sd, feats, obj_ids = head_case_inputs(name)
        net = _build_ref_head(N, p3d, lm)
        missing, unexpected = net.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        net.eval()
        if lm:
            outs = net(feats, p3d[obj_ids - 1], obj_ids)
            init_bits, _, init_g = net.init_net(feats, obj_ids, return_graph_feats=True)
        else:
            outs = net(feats, p3d.expand(B, -1, -1))
            init_bits, _, init_g = net.init_net(feats, return_graph_feats=True)
        roi, xb, yb, seg, xid, yid = outs
        save(name, roi_bit=roi, x_bits=xb, y_bits=yb, seg=seg, x_id=xid, y_id=yid,
             init_bits=init_bits, init_graph_feat=init_g,
             obj_ids=(obj_ids if obj_ids is not None else np.zeros(0)),
             checksum_sd=np.array([syn.tensor_checksum(v) for v in sd.values() if v.dtype.is_floating_point]).sum(),
             checksum_feat=np.array([syn.tensor_checksum(f) for f in feats]).sum())


def fps_case_cloud(case):
    """Synthetic CAD-like vertex clouds (float64, mm), regenerated from the seed by the tests (must mirror tests/helpers.py)."""
    g = torch.Generator().manual_seed(7000 + case)
    V = (3000, 1531, 20000)[case]
    d = torch.randn(V, 3, generator=g, dtype=torch.float64)
    d = d / d.norm(dim=1, keepdim=True)
    r = 40.0 + 15.0 * torch.sin(3.0 * d[:, 0]) * torch.cos(2.0 * d[:, 1]) + torch.rand(V, generator=g, dtype=torch.float64)
    xyz = d * r[:, None] * torch.tensor([1.0, 0.6, 1.4], dtype=torch.float64)
    if case == 1:     # duplicated vertices: argmax ties resolve to the first index
        xyz = torch.cat([xyz, xyz[:300]], dim=0)
    return xyz.numpy()


def golden_fps():
    """farthest_point_sample_init_center of the unmodified reference (preprocess_data/get_fps_points.py:65-90)."""
    sys.path.insert(0, os.path.join(REF, "preprocess_data"))
    import get_fps_points as ref_fps
    out = {}
    for case, npoint in ((0, 256), (1, 128), (2, 512)):
        xyz = fps_case_cloud(case)
        ids, fxyz = ref_fps.farthest_point_sample_init_center(xyz, npoint)
        out[f"c{case}_ids"] = np.asarray(ids, dtype=np.int64)
        out[f"c{case}_xyz"] = fxyz
        out[f"c{case}_checksum"] = np.float64(syn.tensor_checksum(torch.from_numpy(xyz)))
    save("fps", **out)


ABWOPROG_CASE = ("lm", tuple(range(1, 16)), 128, 3, 1234 + 7)   # must mirror tests/helpers.py::ABWOPROG_CASE


def abwoprog_case_inputs():
    ds, objs, N, B, seed = ABWOPROG_CASE
    g = torch.Generator().manual_seed(seed)
    p3d = torch.cat([syn.p3d_normed_tensor(syn.load_fps_xyz(ds, o, N)) for o in objs], dim=0)
    sd = syn.synthetic_state_dict(syn.abwoprog_param_spec(N), g)
    feats = syn.synthetic_features(B, g)
    obj_ids = torch.tensor([objs[(i * 5 + 2) % len(objs)] for i in range(B)])
    return p3d, sd, feats, obj_ids


def golden_abwoprog():
    """PoseNet_GNNskip_ABwoProg (pipeline_lm.py:430-517), the ablation net test_lm.py:173-176 builds."""
    ds, objs, N, B, seed = ABWOPROG_CASE
    p3d, sd, feats, obj_ids = abwoprog_case_inputs()
    init = ref_init_lm.InitNet_GNN(npoint=N, p3d_normed=p3d, res_log2=3, backbone_name="hrnet_w18", pretrain_backbone=False,
                                   num_conv1x1=1, max_batch_size=8, num_graph_module=2, graph_k=20, graph_leaky_slope=0.2)
    net = ref_pipe_lm.PoseNet_GNNskip_ABwoProg(init, npoint=N, p3d_normed=p3d, res_log2=6, num_filters=256, max_batch_size=8,
                                               query_dims=None, local_k=2, leaky_slope=0.01, num_graph_module=3, graph_k=20,
                                               graph_leaky_slope=0.2, query_type="mlp")
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    net.eval()
    roi, xb, yb, seg, xid, yid = net(feats, p3d[obj_ids - 1], obj_ids)
    save("head_abwoprog_lm15_n128_b3", roi_bit=roi, x_bits=xb, y_bits=yb, seg=seg, x_id=xid, y_id=yid, obj_ids=obj_ids,
         checksum_sd=np.array([syn.tensor_checksum(v) for v in sd.values() if v.dtype.is_floating_point]).sum(),
         checksum_feat=np.array([syn.tensor_checksum(f) for f in feats]).sum())


if __name__ == "__main__":
    which = sys.argv[1:] or ["knn", "modules", "decode", "corr", "heads", "abwoprog", "fps"]
    if "abwoprog" in which:
        golden_abwoprog()
    if "fps" in which:
        golden_fps()
    if "knn" in which:
        golden_knn()
    if "modules" in which:
        golden_modules()
    if "decode" in which:
        golden_decode()
    if "corr" in which:
        golden_correspondences()
    if "heads" in which:
        golden_heads()

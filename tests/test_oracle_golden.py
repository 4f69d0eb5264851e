"""CPU: pin oracle/ against the golden vectors produced by the unmodified reference modules."""
import numpy as np
import pytest
import torch

from helpers import HEAD_CASES, check_head_checksums, head_case_inputs, knn_sets_equal, syn
from oracle import checkerpose_oracle as orc

torch.set_grad_enabled(False)


def test_knn_fps_fixtures(golden):
    g = golden("knn")
    n = 0
    for key, ref in g.items():
        if key.startswith("rand"):
            continue
        ds, oid, ns, ks = key.split("_")
        N, K = int(ns[1:]), int(ks[1:])
        p = syn.p3d_normed_tensor(syn.load_fps_xyz(ds, int(oid), N))
        idx = orc.knn(p, K)[0].numpy()
        assert idx.dtype == np.int64
        assert len(knn_sets_equal(idx, ref)) == 0, key
        assert (idx[:, 0] == np.arange(N)).all(), "self must be neighbour 0"
        n += 1
    assert n == 44 + 9 + 4


def test_knn_generic_channels(golden):
    g = golden("knn")
    x = torch.from_numpy(g["rand_c16_n200_k12_x"])
    assert np.array_equal(orc.knn(x, 12).numpy(), g["rand_c16_n200_k12"].astype(np.int64))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_static_graph_module(golden, tag):
    g = golden("modules")
    t = lambda k: torch.from_numpy(g[f"sg_{tag}_{k}"])
    args = (t("conv.0.weight"), t("conv.1.weight"), t("conv.1.bias"), t("conv.1.running_mean"), t("conv.1.running_var"))
    y = orc.static_graph_module(t("x"), t("idx").long(), *args, leaky_slope=0.2)
    assert torch.allclose(y, t("y"), rtol=1e-5, atol=1e-5)
    idx_lm = t("lm_idx").long()[t("lm_obj") - 1]
    y = orc.static_graph_module(t("x"), idx_lm, *args, leaky_slope=0.2)
    assert torch.allclose(y, t("lm_y"), rtol=1e-5, atol=1e-5)


def test_get_graph_feature(golden):
    g = golden("modules")
    y = orc.get_graph_feature(torch.from_numpy(g["ggf_x"]), torch.from_numpy(g["ggf_idx"]).long())
    assert np.array_equal(y.numpy(), g["ggf_y"])


@pytest.mark.parametrize("tag,k", [("a", 2), ("b", 2), ("c", 4)])
def test_index2feat(golden, tag, k):
    g = golden("modules")
    t = lambda n: torch.from_numpy(g[f"i2f_{tag}_{n}"])
    y = orc.index2feat(t("feat"), t("patch_generator.weight"), t("patch_generator.bias"), t("xid"), t("yid"), k)
    assert torch.allclose(y, t("y"), rtol=1e-6, atol=1e-6)


def test_mlp_query(golden):
    g = golden("modules")
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("mq_mlps")}
    y = orc.mlp_leaky(torch.from_numpy(g["mq_x"]), sd, "mlps.", 3, 0.01, last_act=False)
    assert torch.allclose(y, torch.from_numpy(g["mq_y"]), rtol=1e-5, atol=1e-5)


def test_decode_functions(golden):
    g = golden("decode")
    cp = torch.from_numpy(g["code_prob"])
    assert np.array_equal(orc.from_code_prob_to_id(cp).numpy(), g["code_prob_id"])
    assert np.array_equal(orc.from_gt_code_to_id(torch.sigmoid(cp)).numpy(), g["gt_code_id"])
    assert np.array_equal(orc.from_bit_prob_to_id(cp[:, 0:1]).numpy(), g["bit_prob_id"])
    assert np.array_equal(orc.from_gt_bit_to_id(torch.sigmoid(cp[:, 0:1])).numpy(), g["gt_bit_id"])
    assert np.array_equal(orc.from_mask_prob_to_mask(cp).numpy(), g["mask"])
    assert np.array_equal(orc.from_code_to_id(torch.from_numpy(g["code"])).numpy(), g["code_id"])
    for thr in (0.5, 0.3, 0.9):
        assert np.array_equal(orc.from_output_to_class_mask(cp, thr), g[f"co_mask_{thr}"])
        assert np.array_equal(orc.from_output_to_class_mask_torch(cp, thr).numpy(), g[f"co_mask_torch_{thr}"])
        assert np.array_equal(orc.from_output_to_class_binary_code(cp, "BCE", thr), g[f"co_code_bce_{thr}"])
    ce = orc.from_output_to_class_binary_code(torch.from_numpy(g["co_ce_in"]), "CE", 0.5, 2, 16)
    assert np.array_equal(ce, g["co_code_ce"])


def test_codec_functions(golden):
    g = golden("decode")
    assert np.array_equal(orc.class_code_vecs_to_class_id_vec(g["cc_vecs"]), g["cc_vecs_id"])
    assert np.array_equal(orc.class_code_images_to_class_id_image(g["cc_hwc"]), g["cc_hwc_id"])
    assert np.array_equal(orc.class_code_images_to_class_id_image_torch(torch.from_numpy(g["cc_chw"])).numpy(), g["cc_chw_id"])
    assert np.array_equal(orc.class_code_images_to_class_id_image_torch_batch(torch.from_numpy(g["cc_bchw"])).numpy(), g["cc_bchw_id"])
    assert np.array_equal(orc.class_id_vec_to_class_code_vecs(g["cc_ids"], 2, 6), g["cc_ids_code"])
    assert orc.code_to_id([1, 0, 1, 1, 0]) == int(g["cc_code_to_id"])
    assert orc.str_code_to_id("10110") == int(g["cc_str_code_to_id"])


def test_correspondences(golden):
    g = golden("correspondences")
    xyz = syn.load_fps_xyz("lmo", 1, 300)
    for c in range(3):
        grid = orc.roi_xy_ori(g[f"c{c}_bbox"], 64)
        assert np.array_equal(grid, g[f"c{c}_grid"])
        p2d, v_all, v_full, v_vis = orc.id_to_correspondences(
            g[f"c{c}_roi_logit"], g[f"c{c}_seg_logit"], g[f"c{c}_xid"], g[f"c{c}_yid"], g[f"c{c}_bbox"])
        for tag, m in (("all", v_all), ("full", v_full), ("visib", v_vis)):
            assert np.array_equal(xyz[m], g[f"c{c}_{tag}_p3d"])
            assert np.array_equal(p2d[m], g[f"c{c}_{tag}_p2d"])


@pytest.mark.parametrize("name", list(HEAD_CASES))
def test_full_head(golden, name):
    g = golden(name)
    ds, objs, N, B, seed, lm = HEAD_CASES[name]
    p3d, sd, feats, obj_ids = head_case_inputs(name)
    check_head_checksums(g, sd, feats)
    idx = orc.knn(p3d, 20)
    out, inter = orc.pose_head(feats, sd, idx, [idx] * 3, N, obj_ids=obj_ids, return_intermediates=True)
    roi, xb, yb, seg, xid, yid = out
    assert np.array_equal(xid.numpy(), g["x_id"]) and np.array_equal(yid.numpy(), g["y_id"])
    for a, k in ((roi, "roi_bit"), (xb, "x_bits"), (yb, "y_bits"), (seg, "seg"), (inter["graph_feat"][0], "init_graph_feat")):
        assert torch.allclose(a, torch.from_numpy(g[k]), rtol=1e-4, atol=1e-4), k
    assert xid.dtype == torch.int64 and xb.shape == (B, 6, N) and seg.shape == (B, 2, 64, 64)


def test_abwoprog_head(golden):
    """oracle.pose_head_abwoprog vs the unmodified PoseNet_GNNskip_ABwoProg (pipeline_lm.py:430-517)."""
    from helpers import ABWOPROG_CASE, abwoprog_case_inputs
    g = golden("head_abwoprog_lm15_n128_b3")
    ds, objs, N, B, seed = ABWOPROG_CASE
    p3d, sd, feats, obj_ids = abwoprog_case_inputs()
    check_head_checksums(g, sd, feats)
    assert np.array_equal(obj_ids.numpy(), g["obj_ids"])
    idx = orc.knn(p3d, 20)
    roi, xb, yb, seg, xid, yid = orc.pose_head_abwoprog(feats, sd, idx, [idx] * 3, N, obj_ids=obj_ids)
    assert np.array_equal(xid.numpy(), g["x_id"]) and np.array_equal(yid.numpy(), g["y_id"])
    for a, k in ((roi, "roi_bit"), (xb, "x_bits"), (yb, "y_bits"), (seg, "seg")):
        assert torch.allclose(a, torch.from_numpy(g[k]), rtol=1e-4, atol=1e-4), k
    assert xb.shape == (B, 6, N) and yb.shape == (B, 6, N) and xid.dtype == torch.int64


def test_fps(golden):
    """oracle.farthest_point_sample_init_center vs the unmodified reference (get_fps_points.py:65-90)."""
    from helpers import FPS_CASES, fps_case_cloud, syn
    g = golden("fps")
    for case, npoint in FPS_CASES:
        xyz = fps_case_cloud(case)
        assert np.isclose(syn.tensor_checksum(torch.from_numpy(xyz)), float(g[f"c{case}_checksum"]), rtol=1e-12)
        ids, fxyz = orc.farthest_point_sample_init_center(xyz, npoint)
        assert np.array_equal(np.asarray(ids), g[f"c{case}_ids"]) and np.array_equal(fxyz, g[f"c{case}_xyz"])

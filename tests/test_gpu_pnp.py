"""the batched GPU RANSAC PnP front-end (cp_pnp_ransac) against the call the reference makes,
cv2.solvePnPRansac(..., reprojectionError=2, iterationsCount=150, flags=cv2.SOLVEPNP_EPNP)
(checkerpose/test_network_with_test_data.py:103-106; OpenCV is a third-party dependency of the reference, opencv-python
4.13.0 in this image), on the same correspondence records.

The two solvers draw different RANSAC samples and refit differently (EPnP vs Gauss-Newton on the reprojection error), so the
comparison is on the ESTIMATE: rotation within 0.2 degrees and translation within 2 mm (|t| ~ 0.6-1.2 m) of OpenCV's, both
within the same distance of the ground truth; our inlier count (of the refitted pose) is at least OpenCV's (of its best
RANSAC hypothesis) up to 3 %.  Scenes: the shipped FPS keypoint clouds under a
random pose, projected with the LM camera, quantised to the 64 x 64 RoI grid exactly like the head's records, with 20 % of
the keypoints sent to random cells and 10 % marked outside the RoI.
"""
import numpy as np
import pytest
import torch

from helpers import syn

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

K_LM = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1.0]])
TOL_R_DEG, TOL_T_MM = 0.2, 2.0


def scene(ds, obj, N, seed, outlier=0.2, invalid=0.1):
    rng = np.random.default_rng(seed)
    X = syn.load_fps_xyz(ds, obj, N).astype(np.float64)
    R, _ = cv2.Rodrigues(rng.normal(size=3) * 1.2)
    t = np.array([rng.uniform(-150, 150), rng.uniform(-100, 100), rng.uniform(600, 1200)])
    Y = X @ R.T + t
    uv = np.stack([K_LM[0, 0] * Y[:, 0] / Y[:, 2] + K_LM[0, 2], K_LM[1, 1] * Y[:, 1] / Y[:, 2] + K_LM[1, 2]], 1)
    lo, hi = uv.min(0), uv.max(0)
    side = float(np.ceil((hi - lo).max() * 1.2))
    c = (lo + hi) / 2
    bbox = np.array([np.floor(c[0] - side / 2), np.floor(c[1] - side / 2), side, side], dtype=np.float32)
    xid = np.clip(np.floor((uv[:, 0] - bbox[0]) / (side / 64)), 0, 63).astype(np.int64)
    yid = np.clip(np.floor((uv[:, 1] - bbox[1]) / (side / 64)), 0, 63).astype(np.int64)
    bad = rng.random(N) < outlier
    xid[bad] = rng.integers(0, 64, int(bad.sum()))
    yid[bad] = rng.integers(0, 64, int(bad.sum()))
    roi_logit = np.where(rng.random(N) >= invalid, 3.0, -3.0).astype(np.float32)
    return X, R, t, bbox, xid, yid, roi_logit


def rot_err_deg(Ra, Rb):
    return float(np.degrees(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1))))


def cv2_pose(X, p2d, valid):
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X[valid], p2d[valid].astype(np.float64), K_LM, None, reprojectionError=2,
                                             iterationsCount=150, flags=cv2.SOLVEPNP_EPNP)
    R, _ = cv2.Rodrigues(rvec)
    return R, tvec.ravel(), 0 if inl is None else len(inl)


@pytest.mark.parametrize("ds,objs,N", [("lmo", (1, 5, 6, 8), 512), ("ycbv", (2, 5, 15, 21), 4096), ("lm", (3, 9), 1024)])
def test_pnp_ransac_matches_opencv(ds, objs, N):
    from checkerpose_b200 import ops
    B = len(objs)
    scenes = [scene(ds, o, N, 100 * o + N) for o in objs]
    p3d = torch.tensor(np.stack([s[0] for s in scenes]), dtype=torch.float32).cuda()                    # (G = B, N, 3)
    roi = torch.tensor(np.stack([s[6] for s in scenes])).view(B, 1, N).cuda()
    g = torch.Generator().manual_seed(N)
    seg = torch.randn(B, 2, 64, 64, generator=g).cuda()
    bbox = torch.tensor(np.stack([s[3] for s in scenes])).cuda()
    xid = torch.tensor(np.stack([s[4] for s in scenes])).cuda()
    yid = torch.tensor(np.stack([s[5] for s in scenes])).cuda()
    packed = ops.correspondences_packed(roi, seg, bbox, xid, yid)
    sel = torch.arange(B, dtype=torch.int32).cuda()
    K = torch.tensor(K_LM, dtype=torch.float32).cuda()
    R, t, ninl, mask = ops.pnp_ransac(packed, p3d, K, graph_sel=sel, flag=ops.FLAG_ALL, reproj_thresh=2.0, iterations=150, seed=7,
                                      return_inlier_mask=True)
    R2, t2, ninl2 = ops.pnp_ransac(packed, p3d, K, graph_sel=sel, flag=ops.FLAG_ALL, reproj_thresh=2.0, iterations=150, seed=7)
    assert torch.equal(R, R2) and torch.equal(t, t2) and torch.equal(ninl, ninl2), "deterministic for a given seed"
    assert torch.equal(mask.sum(1).to(torch.int32), ninl)
    # the 12-byte records give the same answer as the packed rows
    R3, t3, ninl3 = ops.pnp_ransac(ops.correspondences(roi, seg, bbox, xid, yid), p3d, K, graph_sel=sel, seed=7)
    assert torch.equal(R, R3) and torch.equal(t, t3) and torch.equal(ninl, ninl3)
    uv, flags, _, _, _ = ops.unpack_correspondences_host(packed)
    for b, (X, Rgt, tgt, _, _, _, _) in enumerate(scenes):
        valid = (flags[b] & 1) != 0
        Rcv, tcv, ncv = cv2_pose(X, uv[b], valid)
        Ro, to = R[b].double().cpu().numpy(), t[b].double().cpu().numpy()
        assert abs(np.linalg.det(Ro) - 1) < 1e-4 and np.abs(Ro @ Ro.T - np.eye(3)).max() < 1e-4
        dR, dt = rot_err_deg(Ro, Rcv), float(np.linalg.norm(to - tcv))
        print(f"{ds}/{objs[b]} N={N}: valid {int(valid.sum())}, inliers ours {int(ninl[b])} / OpenCV {ncv}; ours vs OpenCV {dR:.3f} deg, {dt:.2f} mm; "
              f"vs ground truth: ours {rot_err_deg(Ro, Rgt):.3f} deg {np.linalg.norm(to - tgt):.2f} mm, OpenCV {rot_err_deg(Rcv, Rgt):.3f} deg "
              f"{np.linalg.norm(tcv - tgt):.2f} mm")
        assert dR < TOL_R_DEG and dt < TOL_T_MM, (dR, dt)
        # OpenCV reports the inliers of its best RANSAC hypothesis; ours are those of the refitted pose (never fewer, up to noise)
        assert 0.97 * ncv - 2 <= int(ninl[b]) <= int(valid.sum())
        assert rot_err_deg(Ro, Rgt) < rot_err_deg(Rcv, Rgt) + TOL_R_DEG


def test_pnp_ransac_flag_sets_and_degenerate_rois():
    """The three validity sets of test.py:335-368 (all / full mask / visible mask) select different correspondences; a RoI
    with fewer than 4 valid ones returns R = I, t = 0 like the reference (test_network_with_test_data.py:111-114)."""
    from checkerpose_b200 import ops
    N = 512
    X, Rgt, tgt, bbox, xid, yid, roi_logit = scene("lmo", 9, N, 5)
    B = 3
    roi = torch.tensor(roi_logit).view(1, 1, N).repeat(B, 1, 1).cuda()
    roi[2] = -1.0                                   # RoI 2: nothing valid
    roi[2, 0, :3] = 1.0                             # ... except 3 keypoints (< 4)
    seg = torch.full((B, 2, 64, 64), 2.0).cuda()
    seg[1, 1, :, :32] = -2.0                        # RoI 1: the full mask rejects the left half of the grid
    bb = torch.tensor(bbox).view(1, 4).repeat(B, 1).cuda()
    xi = torch.tensor(xid).view(1, N).repeat(B, 1).cuda()
    yi = torch.tensor(yid).view(1, N).repeat(B, 1).cuda()
    packed = ops.correspondences_packed(roi, seg, bb, xi, yi)
    p3d = torch.tensor(X, dtype=torch.float32).view(1, N, 3).cuda()
    K = torch.tensor(K_LM, dtype=torch.float32).cuda()
    R_all, t_all, n_all = ops.pnp_ransac(packed, p3d, K, flag=ops.FLAG_ALL)
    R_full, t_full, n_full, m_full = ops.pnp_ransac(packed, p3d, K, flag=ops.FLAG_FULL, return_inlier_mask=True)
    assert int(n_all[2]) == 0 and torch.equal(R_all[2].cpu(), torch.eye(3)) and float(t_all[2].abs().max()) == 0.0
    assert int(n_full[1]) < int(n_all[1]) and int(n_full[0]) == int(n_all[0])
    assert not bool(m_full[1][(xi[1] < 32)].any()), "keypoints outside the full mask can not be inliers of that set"
    for b in (0, 1):
        assert rot_err_deg(R_full[b].double().cpu().numpy(), Rgt) < 0.6
        assert np.linalg.norm(t_full[b].double().cpu().numpy() - tgt) < 6.0


def test_from_id_to_pose_dropin_matches_the_reference_call():
    """checkerpose_b200.test_network_with_test_data.from_id_to_pose (numpy in / out, the reference's signature) against the
    reference's own steps re-stated with OpenCV (test_network_with_test_data.py:50-115), incl. check_seg and return_inliers."""
    from checkerpose_b200.test_network_with_test_data import from_id_to_pose
    N = 1024
    X, Rgt, tgt, bbox, xid, yid, roi_logit = scene("ycbv", 12, N, 77)
    gy, gx = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    roi_xy_ori = np.stack([bbox[0] + gx * (bbox[2] / 64), bbox[1] + gy * (bbox[3] / 64)], -1).astype(np.float32)   # bop_dataset_pytorch.py:223-235
    roi_mask_bit = (roi_logit > 0).astype(np.float32).reshape(N, 1)
    seg_mask = np.ones((64, 64), dtype=np.float32)
    seg_mask[:8] = 0.0
    for check_seg in (False, True):
        R, t, inl = from_id_to_pose(X, roi_xy_ori, K_LM, roi_mask_bit, xid, yid, check_seg=check_seg, seg_mask=seg_mask, return_inliers=True)
        p2d = roi_xy_ori[yid, xid]
        valid = roi_mask_bit[:, 0] > 0.5
        if check_seg:
            valid &= seg_mask[yid, xid] > 0.5
        Rcv, tcv, ncv = cv2_pose(X, p2d, valid)
        assert R.shape == (3, 3) and t.shape == (3, 1) and inl.ndim == 1
        assert rot_err_deg(R, Rcv) < TOL_R_DEG and np.linalg.norm(t.ravel() - tcv) < TOL_T_MM
        assert valid[inl].all() and 0.97 * ncv - 2 <= len(inl) <= int(valid.sum())
    with pytest.raises(RuntimeError):
        from_id_to_pose(X, roi_xy_ori, K_LM, roi_mask_bit, xid, yid, use_progressivex=True)
    R, t = from_id_to_pose(X, roi_xy_ori, K_LM, np.zeros((N, 1), dtype=np.float32), xid, yid)
    assert np.array_equal(R, np.eye(3)) and np.array_equal(t, np.zeros((3, 1)))


def test_forward_with_pose_runs_the_whole_chain():
    """PoseNet_GNNskip.forward_with_pose = head -> packed records -> cp_pnp_ransac on the device; with random-init weights the
    correspondences carry no pose, so this checks the plumbing: the records it solves are the ones forward_with_correspondences
    returns, and the poses equal a separate pnp_ransac call on them."""
    from checkerpose_b200 import head, ops
    from checkerpose_b200.model import init, pipeline
    from checkerpose_b200.model.backbone import FeatureListBackbone
    N, B = 256, 3
    g = torch.Generator().manual_seed(3)
    xyz = syn.load_fps_xyz("lmo", 5, N)
    p3d = syn.p3d_normed_tensor(xyz).cuda()
    sd = syn.synthetic_state_dict(syn.head_param_spec(N), g)
    inet = init.InitNet_GNN(npoint=N, p3d_normed=p3d, res_log2=3, backbone_name="hrnet_w18", pretrain_backbone=False, img_backbone=FeatureListBackbone())
    net = pipeline.PoseNet_GNNskip(inet, npoint=N, p3d_normed=p3d, res_log2=6, local_k=2, num_graph_module=3)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    feats = [f.cuda() for f in syn.synthetic_features(B, g)]
    bbox = syn.synthetic_bboxes(B, g).cuda()
    K = torch.tensor(K_LM, dtype=torch.float32).cuda()
    X = torch.tensor(xyz, dtype=torch.float32).view(1, N, 3).cuda()
    head.set_compute_dtype(torch.float32)
    out, packed, R, t, ninl = net.forward_with_pose(feats, p3d.expand(B, -1, -1), bbox, X, K, seed=5)
    out2, packed2 = net.forward_with_correspondences(feats, p3d.expand(B, -1, -1), bbox, packed=True)
    assert torch.equal(packed, packed2) and R.shape == (B, 3, 3) and t.shape == (B, 3) and ninl.shape == (B,)
    R2, t2, n2 = ops.pnp_ransac(packed2, X, K, seed=5)
    assert torch.equal(R, R2) and torch.equal(t, t2) and torch.equal(ninl, n2)

"""Drop-in for ``from_id_to_pose`` of checkerpose/test_network_with_test_data.py:32-119 (SURVEY.md section 8a row 12 and
section 8f rank 2): decoded ids -> 2D-3D correspondences -> pose.

The correspondence half is ``cp_correspondences``' arithmetic; the OpenCV branch (``cv2.solvePnPRansac`` with EPnP,
reference :103-106) is replaced by the batched GPU RANSAC-P3P + Gauss-Newton solver ``cp_pnp_ransac``.  The
Progressive-X branch stays on the reference's path (north_star) and raises here.  numpy in, numpy out, like the reference;
for whole batches use ``ops.pnp_ransac`` on the records the head already produced on the device.
"""
import numpy as np
import torch

from . import ops


def from_id_to_pose(p3d_xyz, roi_xy_ori, cam_K, roi_mask_bit, pixel_x_id, pixel_y_id, check_seg=False,
                    seg_mask=None, use_progressivex=False, neighborhood_ball_radius=20, spatial_coherence_weight=0.1,
                    prog_max_iters=400, discard_bd_pixel=0, return_inliers=False,
                    reprojErr_thresh=2, cv_max_iters=150, seed=0):
    """Same arguments and return values as the reference (R (3,3), t (3,1)[, inlier keypoint ids or None])."""
    if use_progressivex:
        raise RuntimeError("checkerpose_b200: the Progressive-X solver stays on the reference's path; call the reference's "
                           "from_id_to_pose for use_progressivex=True")
    if not torch.cuda.is_available():
        raise RuntimeError("checkerpose_b200: a CUDA device is required (there is no CPU fallback)")
    p3d_xyz = np.asarray(p3d_xyz)
    N = p3d_xyz.shape[0]
    roi_h, roi_w, _ = roi_xy_ori.shape
    pixel_x_id, pixel_y_id = np.asarray(pixel_x_id), np.asarray(pixel_y_id)
    disc_p2d = np.asarray(roi_xy_ori)[pixel_y_id, pixel_x_id]                                   # reference :54
    valid = np.asarray(roi_mask_bit)[:, 0] > 0.5                                                # :56
    if check_seg:
        valid = np.logical_and(valid, np.asarray(seg_mask)[pixel_y_id, pixel_x_id] > 0.5)       # :58
    if discard_bd_pixel > 0:                                                                    # :59-62
        inside = ((pixel_y_id >= discard_bd_pixel) & (pixel_y_id < roi_h - discard_bd_pixel) &
                  (pixel_x_id >= discard_bd_pixel) & (pixel_x_id < roi_w - discard_bd_pixel))
        valid = np.logical_and(valid, inside)
    R_predict, t_predict, inliers = np.eye(3), np.zeros((3, 1)), None
    if int(valid.sum()) >= 4:                                                                   # :98
        npad = N + (N & 1)                                                                      # the kernel wants an even N
        rec = np.zeros((1, npad, 3), dtype=np.int32)
        rec[0, :N, :2] = disc_p2d.astype(np.float32).view(np.int32)
        rec[0, :N, 2] = valid.astype(np.int32)
        xyz = np.zeros((1, npad, 3), dtype=np.float32)
        xyz[0, :N] = p3d_xyz
        dev = torch.device("cuda", torch.cuda.current_device())
        R, t, ninl, mask = ops.pnp_ransac(torch.from_numpy(rec).to(dev), torch.from_numpy(xyz).to(dev),
                                          torch.as_tensor(np.asarray(cam_K), dtype=torch.float32, device=dev), flag=1,
                                          reproj_thresh=float(reprojErr_thresh), iterations=int(cv_max_iters), seed=seed,
                                          return_inlier_mask=True)
        if int(ninl[0]) > 0:
            R_predict = R[0].double().cpu().numpy()
            t_predict = t[0].double().cpu().numpy().reshape(3, 1)
            inliers = np.nonzero(mask[0, :N].cpu().numpy())[0]
    if return_inliers:
        return R_predict, t_predict, inliers
    return R_predict, t_predict

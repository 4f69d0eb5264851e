"""Throughput of the batched GPU RANSAC PnP (cp_pnp_ransac) next to cv2.solvePnPRansac on the host, on the same synthetic
scenes as tests/test_gpu_pnp.py: python scripts/kbench_pnp.py [B] [N]"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from checkerpose_b200 import ops  # noqa: E402
from test_gpu_pnp import K_LM, rot_err_deg, scene  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    objs = [1 + (i % 21) for i in range(B)]
    scenes = [scene("ycbv", o, N, 1000 + i) for i, o in enumerate(objs)]
    p3d = torch.tensor(np.stack([s[0] for s in scenes]), dtype=torch.float32).cuda()
    roi = torch.tensor(np.stack([s[6] for s in scenes])).view(B, 1, N).cuda()
    seg = torch.ones(B, 2, 64, 64).cuda()
    bbox = torch.tensor(np.stack([s[3] for s in scenes])).cuda()
    xid = torch.tensor(np.stack([s[4] for s in scenes])).cuda()
    yid = torch.tensor(np.stack([s[5] for s in scenes])).cuda()
    packed = ops.correspondences_packed(roi, seg, bbox, xid, yid)
    sel = torch.arange(B, dtype=torch.int32).cuda()
    K = torch.tensor(K_LM, dtype=torch.float32).cuda()
    for it in (150, 256, 1024):
        for _ in range(2):
            R, t, n = ops.pnp_ransac(packed, p3d, K, graph_sel=sel, iterations=it, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            R, t, n = ops.pnp_ransac(packed, p3d, K, graph_sel=sel, iterations=it, seed=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        errs = [rot_err_deg(R[b].double().cpu().numpy(), scenes[b][1]) for b in range(B)]
        print(f"cp_pnp_ransac B={B} N={N} iterations={it}: {ms:.3f} ms per batch = {B / ms * 1e3:.0f} RoIs/s; "
              f"median rotation error vs ground truth {np.median(errs):.3f} deg, max {np.max(errs):.3f} deg")
    uv, flags, _, _, _ = ops.unpack_correspondences_host(packed)
    nb = min(B, 16)
    t0 = time.perf_counter()
    errs = []
    for b in range(nb):
        valid = (flags[b] & 1) != 0
        ok, rvec, tvec, inl = cv2.solvePnPRansac(scenes[b][0][valid], uv[b][valid].astype(np.float64), K_LM, None, reprojectionError=2,
                                                 iterationsCount=150, flags=cv2.SOLVEPNP_EPNP)
        errs.append(rot_err_deg(cv2.Rodrigues(rvec)[0], scenes[b][1]))
    dt = (time.perf_counter() - t0) / nb
    print(f"cv2.solvePnPRansac (EPnP, 150 iterations) on one host core: {dt * 1e3:.2f} ms per RoI = {1 / dt:.0f} RoIs/s; "
          f"median rotation error {np.median(errs):.3f} deg")


if __name__ == "__main__":
    main()

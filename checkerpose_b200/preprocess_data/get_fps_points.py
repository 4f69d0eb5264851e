"""Drop-in for checkerpose/preprocess_data/get_fps_points.py:65-90: same function name, NumPy in / NumPy out, ids
bit-exact with the reference's float64 loop -- computed by cp_fps on the GPU (no CPU fallback)."""
import ctypes as C

import numpy as np
import torch

from .. import ops
from .._lib import check, lib


def farthest_point_sample_init_center(xyz, npoint):
    ''' compute the FPS points of the given 3D points
    Args:
        xyz: point cloud data, shape (N, 3)
        npoint: number of FPS points
    Returns (fps_ids: list of int, fps_xyz: (npoint, 3) float64), as the reference.
    '''
    if not torch.cuda.is_available():
        raise RuntimeError("checkerpose_b200: farthest_point_sample_init_center needs a CUDA device (there is no CPU fallback)")
    xyz = np.ascontiguousarray(np.asarray(xyz, dtype=np.float64))
    if xyz.ndim != 2 or xyz.shape[1] != 3:
        raise ValueError("xyz must have shape (N, 3)")
    # get_fps_points.py:74-80, on the host in NumPy as the reference does (O(N), once)
    xyz_max = xyz.max(axis=0)
    xyz_min = xyz.min(axis=0)
    xyz_center = (xyz_max + xyz_min) / 2
    xyz_extent = np.linalg.norm(xyz_max - xyz_min)
    V = xyz.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.from_numpy(xyz).to(dev)
    dist = torch.empty(V, dtype=torch.float64, device=dev)
    ids = torch.empty(int(npoint), dtype=torch.int64, device=dev)
    out = torch.empty((int(npoint), 3), dtype=torch.float64, device=dev)
    center = (C.c_double * 3)(*[float(v) for v in xyz_center])
    check(lib.cp_fps(x.data_ptr(), V, int(npoint), center, float(np.ones(1)[0] * xyz_extent * 10), dist.data_ptr(),
                     ids.data_ptr(), out.data_ptr(), ops._stream()), "cp_fps")
    ops._count()
    return [int(i) for i in ids.cpu().tolist()], out.cpu().numpy()

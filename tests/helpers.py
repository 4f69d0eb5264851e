"""Shared helpers for the parity tests (inputs regenerated from seeds, comparators)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from checkerpose_b200 import synthetic as syn  # noqa: E402

# must mirror tests/golden/make_golden.py::HEAD_CASES
HEAD_CASES = {
    "head_lmo_ape_n512_b1": ("lmo", (1,), 512, 1, 1234 + 0, False),
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


def fps_case_cloud(case):
    """must mirror tests/golden/make_golden.py::fps_case_cloud"""
    g = torch.Generator().manual_seed(7000 + case)
    V = (3000, 1531, 20000)[case]
    d = torch.randn(V, 3, generator=g, dtype=torch.float64)
    d = d / d.norm(dim=1, keepdim=True)
    r = 40.0 + 15.0 * torch.sin(3.0 * d[:, 0]) * torch.cos(2.0 * d[:, 1]) + torch.rand(V, generator=g, dtype=torch.float64)
    xyz = d * r[:, None] * torch.tensor([1.0, 0.6, 1.4], dtype=torch.float64)
    if case == 1:     # duplicated vertices: argmax ties resolve to the first index
        xyz = torch.cat([xyz, xyz[:300]], dim=0)
    return xyz.numpy()


FPS_CASES = ((0, 256), (1, 128), (2, 512))


# must mirror tests/golden/make_golden.py::ABWOPROG_CASE
ABWOPROG_CASE = ("lm", tuple(range(1, 16)), 128, 3, 1234 + 7)


def abwoprog_case_inputs():
    ds, objs, N, B, seed = ABWOPROG_CASE
    g = torch.Generator().manual_seed(seed)
    p3d = torch.cat([syn.p3d_normed_tensor(syn.load_fps_xyz(ds, o, N)) for o in objs], dim=0)
    sd = syn.synthetic_state_dict(syn.abwoprog_param_spec(N), g)
    feats = syn.synthetic_features(B, g)
    obj_ids = torch.tensor([objs[(i * 5 + 2) % len(objs)] for i in range(B)])
    return p3d, sd, feats, obj_ids


def check_head_checksums(gold, sd, feats):
    cs_sd = np.array([syn.tensor_checksum(v) for v in sd.values() if v.dtype.is_floating_point]).sum()
    cs_f = np.array([syn.tensor_checksum(f) for f in feats]).sum()
    assert np.isclose(cs_sd, float(gold["checksum_sd"]), rtol=1e-12), "synthetic weights drifted from the golden run"
    assert np.isclose(cs_f, float(gold["checksum_feat"]), rtol=1e-12), "synthetic features drifted from the golden run"


def knn_sets_equal(idx, ref_sorted, tie_rows=None):
    """Rows of idx (N,K) equal rows of ref_sorted as SETS, except rows listed in tie_rows."""
    got = np.sort(np.asarray(idx), axis=1)
    bad = np.nonzero((got != np.asarray(ref_sorted)).any(axis=1))[0]
    if tie_rows is not None:
        bad = np.setdiff1d(bad, tie_rows)
    return bad


def knn_tie_rows(p3d_1cn: torch.Tensor, k: int, tol: float = 1e-6):
    """Rows whose k-th and (k+1)-th nearest distances differ by <= tol (fp64 distances)."""
    x = p3d_1cn[0].double().t()                       # (N,3)
    d = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    s = torch.sort(d, dim=1)[0]
    if k >= s.shape[1]:
        return np.zeros(0, dtype=np.int64)
    gap = s[:, k] - s[:, k - 1]
    return np.nonzero((gap <= tol).numpy())[0]


def rel_err(a, b):
    """max |a-b| / max(|b|, rms(b)) -- elementwise relative error with an rms floor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    floor = np.sqrt((b ** 2).mean()) + 1e-30
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())

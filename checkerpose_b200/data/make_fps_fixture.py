"""Pack the reference's shipped FPS keypoint clouds into one small fixture.

Source (read-only, this container only):
  /root/reference/checkerpose/datasets/BOP_DATASETS/{lm,lmo,ycbv}/fps_202212/obj_*.pkl
  (written by preprocess_data/get_fps_points.py:95,118-122; loaded by test.py:145-148).
Each pickle holds {'npoint': 4096, 'id': [...], 'xyz': (4096,3) float64 in mm}; every value is
exactly float32-representable (checked below), so the fixture stores float32.

Output: checkerpose_b200/data/fps_202212.npz with one array per object, key "<dataset>/<obj_id>".
These are DATA fixtures (keypoint clouds), not reference source code.
"""
import glob, os, pickle
import numpy as np

ROOT = "/root/reference/checkerpose/datasets/BOP_DATASETS"
out = {}
for ds in ("lm", "lmo", "ycbv"):
    for p in sorted(glob.glob(os.path.join(ROOT, ds, "fps_202212", "obj_*.pkl"))):
        d = pickle.load(open(p, "rb"))
        xyz = np.asarray(d["xyz"])
        assert xyz.shape == (4096, 3) and xyz.dtype == np.float64
        x32 = xyz.astype(np.float32)
        assert np.array_equal(x32.astype(np.float64), xyz), p
        oid = int(os.path.basename(p)[4:10])
        out[f"{ds}/{oid}"] = x32
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fps_202212.npz")
np.savez_compressed(dst, **out)
print(len(out), "objects ->", dst, os.path.getsize(dst), "bytes")

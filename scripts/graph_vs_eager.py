"""Eager step vs CUDA-graph replay of the same step (default bench workload): python scripts/graph_vs_eager.py [batch]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import argparse  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from checkerpose_b200 import head  # noqa: E402
from checkerpose_b200.graphs import CapturedHead  # noqa: E402

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
args = argparse.Namespace(config="full4096", npoint=0, graph_k=0, batch=B)
dev = torch.device("cuda", 0)
head.set_compute_dtype(torch.bfloat16)
wl = bench.workload(args, 1)
case = bench.build_case(wl, dev, 0)
feats, bbox = bench.make_inputs(case, dev, torch.bfloat16, 0)
net = case["net"]
pexp = case["p3d"].expand(B, -1, -1)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_eager = timed(lambda: net.forward_with_correspondences(feats, pexp, bbox, packed=True))
cap = CapturedHead(net, feats, pexp, bbox, packed=True)
t_graph = timed(lambda: cap())
print(f"B={B}: eager {t_eager:.3f} ms/step, CUDA graph replay {t_graph:.3f} ms/step ({100 * (t_eager - t_graph) / t_eager:.1f} % faster)")

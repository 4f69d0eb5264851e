"""K1 (cp_knn) standalone: keypoint-count x graph_k sweep at C = 3 (the shipped graphs) and a generic-C case.
Prints ms per graph and distance pairs per second; python scripts/kbench_knn.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from checkerpose_b200 import ops, synthetic as syn  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda", 0)
    print("K1 cp_knn, 15 LM graphs per call (B' = 15, C = 3), fp32 direct-difference distances + warp top-k")
    print(f"{'N':>6} {'K':>4} {'ms/call':>9} {'ms/graph':>9} {'Gpairs/s':>9}")
    for N in (512, 1024, 2048, 4096):
        p = torch.cat([syn.p3d_normed_tensor(syn.load_fps_xyz("lm", o, N)) for o in range(1, 16)], dim=0).to(dev)
        for K in (8, 16, 20, 32, 40):
            t = timed(lambda: ops.knn(p, K))
            print(f"{N:6d} {K:4d} {t:9.4f} {t / 15:9.4f} {15 * N * N / t / 1e6:9.1f}")
    g = torch.Generator(device=dev).manual_seed(1)
    for C, N, K in ((64, 4096, 20), (16, 2048, 12)):
        x = torch.randn(4, C, N, generator=g, device=dev)
        t = timed(lambda: ops.knn(x, K))
        print(f"generic C={C} N={N} K={K} B'=4: {t:.4f} ms/call, {4 * N * N * C * 2 / t / 1e9:.2f} TFLOP/s-equivalent of distance FMAs")


if __name__ == "__main__":
    main()

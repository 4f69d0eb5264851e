"""Micro-benchmark of the split-precision GEMM / implicit-GEMM conv (cp_gemm_x3) at the float32-mode shapes of the
benchmark configuration (256 RoIs, N=4096): python scripts/kbench_x3.py [lin] [conv] [c1x1]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from checkerpose_b200 import ops  # noqa: E402


def timed(fn, n=5, warm=2):
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
    what = set(sys.argv[1:]) or {"lin", "conv", "c1x1"}
    B = int(os.environ.get("KB_B", 256))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    if "lin" in what:
        for K, Nout in ((256, 256), (256, 512), (256, 64), (64, 128)):
            M = B * 4096
            a = torch.randn(M, K, generator=g, device=dev)
            w = torch.randn(Nout, K, generator=g, device=dev) / K ** 0.5
            ws = ops.pack_weight_split(w)
            bias = torch.randn(Nout, generator=g, device=dev)
            out = torch.empty((M, Nout), device=dev)
            t = timed(lambda: ops.gemm_x3_linear(a, ws, Nout, bias, True, 0.01, out=out))
            fl = 2.0 * M * K * Nout
            print(f"x3 linear M={M} K={K} N={Nout}: {t:.4f} ms  {fl / t / 1e9:.0f} TFLOP/s useful ({3 * fl / t / 1e9:.0f} bf16-equivalent)  "
                  f"in+out {(M * K * 4 + M * Nout * 4) / t / 1e6:.0f} GB/s")
    if "conv" in what:
        for H, Cin in ((64, 256), (32, 256)):
            x = torch.randn(B, H, H, Cin, generator=g, device=dev)
            w = torch.randn(256, 9 * Cin, generator=g, device=dev) / (9 * Cin) ** 0.5
            ws = ops.pack_weight_split(w)
            bias = torch.randn(256, generator=g, device=dev)
            t = timed(lambda: ops.gemm_x3_conv(x, ws, 256, 3, 3, 1, H, H, bias, True, 0.0), n=3, warm=1)
            fl = 2.0 * B * H * H * 9 * Cin * 256
            print(f"x3 conv3x3 B={B} H={H} {Cin}->256: {t:.4f} ms  {fl / t / 1e9:.0f} TFLOP/s useful ({3 * fl / t / 1e9:.0f} bf16-equivalent)")
    if "c1x1" in what:
        M, K, Nout = B * 64, 1024, 4096
        a = torch.randn(M, K, generator=g, device=dev)
        w = torch.randn(Nout, K, generator=g, device=dev) / K ** 0.5
        ws = ops.pack_weight_split(w)
        t = timed(lambda: ops.gemm_x3_linear(a, ws, Nout, None, False, 0.0))
        fl = 2.0 * M * K * Nout
        print(f"x3 conv1x1 M={M} K={K} N={Nout}: {t:.4f} ms  {fl / t / 1e9:.0f} TFLOP/s useful ({3 * fl / t / 1e9:.0f} bf16-equivalent)")


if __name__ == "__main__":
    main()

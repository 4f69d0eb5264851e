"""bf16 implicit-GEMM convolution (cp_conv_bf16) at the image-branch shapes of the benchmark configuration, next to cuDNN
through torch on the same tensors: python scripts/kbench_conv.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

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
    B = int(os.environ.get("KB_B", 256))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    for H, Cin, Cout, k in ((64, 512, 256, 3), (64, 256, 256, 3), (32, 768, 256, 3), (32, 256, 256, 3), (16, 256, 256, 3), (64, 256, 64, 2)):
        pad = 1
        x = torch.randn(B, H, H, Cin, generator=g, device=dev).to(torch.bfloat16)
        w = (torch.randn(Cout, Cin, k, k, generator=g, device=dev) / (k * k * Cin) ** 0.5)
        bias = torch.randn(Cout, generator=g, device=dev)
        wp = ops.pack_weight(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous())
        Ho = H + 2 * pad - k + 1
        t = timed(lambda: ops.conv_bf16(x, wp, Cout, k, k, pad, Ho, Ho, bias, True, 0.0))
        xc = x.permute(0, 3, 1, 2)          # channels_last NCHW view
        wc = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        bc = bias.to(torch.bfloat16)
        tc = timed(lambda: torch.relu_(F.conv2d(xc, wc, bc, padding=pad)))
        fl = 2.0 * B * Ho * Ho * k * k * Cin * Cout
        print(f"conv{k}x{k} B={B} H={H} {Cin}->{Cout}: tcgen05 {t:.3f} ms ({fl / t / 1e9:.0f} TFLOP/s)   cuDNN {tc:.3f} ms ({fl / tc / 1e9:.0f} TFLOP/s)")


if __name__ == "__main__":
    main()

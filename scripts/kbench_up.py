"""x2 bilinear upsampling + concat (cp_upsample2x_cat_nhwc_to) at the two shapes of the benchmark step:
python scripts/kbench_up.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "scripts"))
import torch  # noqa: E402

from checkerpose_b200 import ops  # noqa: E402
from kbench_conv import timed  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    B = int(os.environ.get("KB_B", 256))
    for H, Ca, Cb in ((16, 256, 512), (32, 256, 256)):
        a = torch.randn(B, H, H, Ca, generator=g, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
        b = torch.randn(B, H, H, Cb, generator=g, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
        t = timed(lambda: ops.upsample2x_cat_padded(a, b), n=20)
        out_b = B * (2 * H + 1) ** 2 * (Ca + Cb) * 2
        in_b = B * H * H * (Ca + Cb) * 2
        print(f"upsample2x_cat {H}x{H} -> {2 * H}x{2 * H}, {Ca}+{Cb} ch: {t:.4f} ms  ({(out_b + in_b) / t / 1e6:.0f} GB/s incl. zero_border launch)")


if __name__ == "__main__":
    main()

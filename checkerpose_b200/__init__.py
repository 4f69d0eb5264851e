"""checkerpose_b200: B200-native (sm_100a) implementation of CheckerPose's GNN keypoint head.

Public surface = the reference's own module API, re-exported from ``checkerpose_b200.model``,
``checkerpose_b200.common_ops`` and ``checkerpose_b200.binary_code_helper``; plus
``set_compute_dtype`` ("fp32": split-precision tensor-core mode that reproduces the reference's decode / "bf16": fastest mode)
and ``set_image_branch`` (bf16 mode: "tcgen05" = our implicit-GEMM convolutions, "cudnn" = library convolutions).
Importing the kernels requires the in-tree CUDA library (``python -m checkerpose_b200.build``).
"""
__version__ = "0.1.0"


def set_compute_dtype(dtype):
    from . import head
    head.set_compute_dtype(dtype)


def get_compute_dtype():
    from . import head
    return head.get_compute_dtype()


def set_image_branch(kind):
    from . import head
    head.set_image_branch(kind)


def get_image_branch():
    from . import head
    return head.get_image_branch()

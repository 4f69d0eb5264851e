"""Drop-in for checkerpose/common_ops.py: same names, argument meaning (incl. the ``thershold``
spelling) and return types, with the sigmoid/threshold/argmax work done by the CUDA kernels.
Inputs must be CUDA tensors (no CPU fallback)."""
import numpy as np
import torch

from . import ops


def from_output_to_class_mask(pred_mask_prob, thershold=0.5):
    """logits -> numpy float64 {0,1} array where sigmoid(x) > thershold (common_ops.py:5-11)."""
    return ops.threshold(pred_mask_prob.detach(), thr=thershold, apply_sigmoid=True).cpu().numpy().astype(np.float64)


def from_output_to_class_mask_torch(pred_mask_prob, thershold=0.5):
    """logits -> float32 tensor {0,1} on the same device (common_ops.py:14-18)."""
    return ops.threshold(pred_mask_prob.detach(), thr=thershold, apply_sigmoid=True)


def from_output_to_class_binary_code(pred_code_prob, BinaryCode_Loss_Type, thershold=0.5,
                                     divided_num_each_interation=2, binary_code_length=16):
    """common_ops.py:21-40.  BCE-family: thresholded sigmoid; CE: argmax over groups of
    ``divided_num_each_interation`` channels."""
    if BinaryCode_Loss_Type in ["BCE", "L1", "SSIM", "L1_SSIM"]:
        return from_output_to_class_mask(pred_code_prob, thershold)
    if BinaryCode_Loss_Type == "CE":
        h, w = pred_code_prob.shape[2], pred_code_prob.shape[3]
        code = ops.group_argmax(pred_code_prob.detach().reshape(-1, divided_num_each_interation, h, w),
                                divided_num_each_interation)
        return code.cpu().numpy().reshape(-1, binary_code_length, h, w)
    raise UnboundLocalError("pred_code: unknown BinaryCode_Loss_Type {}".format(BinaryCode_Loss_Type))


def get_batch_size(second_dataset_ratio, batch_size):
    """Split a batch between the two training datasets (common_ops.py:43-46)."""
    second = int(batch_size * second_dataset_ratio)
    return batch_size - second, second


def from_dim_str_to_tuple(src_str):
    """"256_256_64" -> (256, 256, 64); None stays None (common_ops.py:49-56)."""
    if src_str is None:
        return None
    return tuple(int(tok) for tok in src_str.split("_"))

"""Drop-in for checkerpose/binary_code_helper/class_id_encoder_decoder.py.

MSB-first code <-> integer id conversions on the CUDA kernels ``cp_bits_to_id`` / ``cp_id_to_bits``.
The numpy-facing functions keep their numpy-in / numpy-out (float64) contract by staging through the
current CUDA device; the torch-facing ones take and return CUDA tensors.  ``code_to_id`` and
``str_code_to_id`` convert one host-side code word and stay scalar Python.
"""
import numpy as np
import torch

from .. import ops


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("checkerpose_b200: a CUDA device is required (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def RGB_image_to_class_id_image(RGB_image):
    """(H,W,3) integer image in OpenCV's B,G,R channel order -> (H,W) int ids = B << 16 | G << 8 | R (reference :6-15).
    Host-side dataset helper on numpy arrays (the reference requires numpy here too); no kernel involved."""
    img = np.asarray(RGB_image).astype(int)
    return (img[:, :, 0] << 16) + (img[:, :, 1] << 8) + img[:, :, 2]


def class_code_images_to_class_id_image(class_code_images, class_base=2):
    """(H,W,C) numpy code images -> (H,W) float64 ids (reference :17-28)."""
    t = torch.as_tensor(np.ascontiguousarray(class_code_images), dtype=torch.float32, device=_dev())
    return ops.bits_to_id(t, 2, binarize=False, base=class_base, as_long=False).cpu().numpy().astype(np.float64)


def class_code_vecs_to_class_id_vec(class_code_vecs, class_base=2):
    """(N,C) numpy code vectors -> (N,) float64 ids (reference :30-38)."""
    t = torch.as_tensor(np.ascontiguousarray(class_code_vecs), dtype=torch.float32, device=_dev())
    return ops.bits_to_id(t, 1, binarize=False, base=class_base, as_long=False).cpu().numpy().astype(np.float64)


def class_code_images_to_class_id_image_torch(class_code_images, class_base=2):
    """(C,H,W) tensor -> (H,W) float32 ids on the same device (reference :40-52)."""
    return ops.bits_to_id(class_code_images, 0, binarize=False, base=class_base, as_long=False)


def class_code_images_to_class_id_image_torch_batch(class_code_images, class_base=2):
    """(B,C,H,W) tensor -> (B,H,W) int64 ids (reference :54-63)."""
    return ops.bits_to_id(class_code_images, 1, binarize=False, base=class_base, as_long=True)


def class_id_image_to_class_code_images(class_id_image, class_base=2, iteration=8, number_of_class=256):
    """(H,W) numpy ids -> (H,W,iteration) float64 digits (reference :65-85)."""
    if class_base ** iteration != number_of_class:
        raise ValueError('this combination of base and itration is not possible')
    ids = torch.as_tensor(np.ascontiguousarray(class_id_image).astype(np.int64), device=_dev())
    return ops.id_to_bits(ids, int(iteration), class_base).cpu().numpy().astype(np.float64)


def class_id_vec_to_class_code_vecs(class_id_vec, class_base=2, iteration=8):
    """(N,) numpy ids -> (N,iteration) float64 digits (reference :88-101)."""
    ids = torch.as_tensor(np.ascontiguousarray(class_id_vec).astype(np.int64), device=_dev())
    return ops.id_to_bits(ids, int(iteration), class_base).cpu().numpy().astype(np.float64)


def code_to_id(class_code, class_base=2):
    """One host-side code word (sequence of digits) -> id (reference :104-114)."""
    value = 0
    for digit in class_code:
        value = value * class_base + digit
    return value


def str_code_to_id(str_class_code, class_base=2):
    """One host-side code string, e.g. "10110" -> 22 (reference :116-127)."""
    return code_to_id([int(ch) for ch in str_class_code], class_base)

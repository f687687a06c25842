"""GPU-side paired target augmentation (extension beyond the reference's module API).

The reference builds the two target views on CPU data-loader workers: ``utils/folder.py:127-147`` applies
``transform`` and ``transform_aug`` (``resnet50_dwt_mec_officehome.py:526-542``) to the same PIL image; the augmented
pipeline is RandomCrop -> RandomHorizontalFlip -> ToTensor -> ``_random_affine_augmentation`` (cv2.warpAffine,
``:481-487``) -> ``_gaussian_blur`` (kernel size 1 = identity, ``:489-492``) -> Normalize.  At the step rates of the
CUDA layers (thousands of images/s per GPU) those workers cannot keep up, and they ship 602 KB of float32 per view.

``PairedAugment`` takes the *resized uint8* images (196 KB each at 256x256) that the workers would have cropped,
draws the same random quantities (crop corners, flip, the four N(0, 0.1) entries of the affine matrix) and produces
both normalised float views in one kernel launch, optionally straight into slices of the model's input batch.
The warp reproduces cv2.warpAffine bit for bit (fixed-point coordinates and all).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native as nv

MEAN = (0.485, 0.456, 0.406)      # resnet50_dwt_mec_officehome.py:530
STD = (0.229, 0.224, 0.225)


def draw_params(batch, src_size, crop, rng, affine_sigma=0.1):
    """The random draws of one batch, same distributions as the reference pipelines: RandomCrop corners uniform in
    [0, src - crop], flip with p = 0.5, affine [[1+n0, n1, 0], [n2, 1+n3, 0]] with n ~ N(0, sigma).
    rng: numpy Generator.  -> dict of CPU tensors (crop_plain, crop_aug int32 [B,2]; flip uint8 [B]; affine f32 [B,2,3])."""
    span = src_size - crop + 1
    n = rng.normal(0.0, affine_sigma, size=(batch, 4)).astype(np.float64)
    affine = np.zeros((batch, 2, 3), dtype=np.float32)
    affine[:, 0, 0] = np.float32(1 + n[:, 0]); affine[:, 0, 1] = np.float32(n[:, 1])
    affine[:, 1, 0] = np.float32(n[:, 2]); affine[:, 1, 1] = np.float32(1 + n[:, 3])
    return dict(crop_plain=torch.from_numpy(rng.integers(0, span, (batch, 2)).astype(np.int32)),
                crop_aug=torch.from_numpy(rng.integers(0, span, (batch, 2)).astype(np.int32)),
                flip=torch.from_numpy((rng.random(batch) < 0.5).astype(np.uint8)),
                affine=torch.from_numpy(affine))


class PairedAugment:
    def __init__(self, crop=224, mean=MEAN, std=STD):
        self.crop = int(crop)
        self._mean = (ctypes.c_float * 3)(*mean)
        self._std = (ctypes.c_float * 3)(*std)

    def __call__(self, images, crop_plain=None, crop_aug=None, flip=None, affine=None, out_plain=None, out_aug=None,
                 want_aug=True, channels_last=False):
        """images: [B, H, W, 3] uint8 CUDA tensor (resized, HWC as decoded).  Parameters: CUDA tensors shaped as
        draw_params() returns them.  out_plain / out_aug: optional preallocated float32 [B,3,crop,crop] destinations
        (contiguous, or channels_last when channels_last=True) -- e.g. slices of the [3B,3,crop,crop] model input.
        -> (plain, aug)  (aug is None when want_aug=False: the source domain has no augmented view)."""
        if images.dim() != 4 or images.shape[3] != 3 or images.dtype != torch.uint8:
            raise ValueError(f"expected a [B, H, W, 3] uint8 batch, got {tuple(images.shape)} {images.dtype}")
        dev = nv.require_cuda(images, any_dtype=True)
        images = images.contiguous()
        b, sh, sw = images.shape[0], images.shape[1], images.shape[2]
        fmt = torch.channels_last if channels_last else torch.contiguous_format

        def dest(t):
            if t is None:
                return torch.empty((b, 3, self.crop, self.crop), dtype=torch.float32, device=dev, memory_format=fmt)
            if t.shape != (b, 3, self.crop, self.crop) or t.dtype != torch.float32 or not t.is_contiguous(memory_format=fmt):
                raise ValueError("output must be a float32 [B, 3, crop, crop] tensor dense in the requested memory format")
            nv.require_cuda(t)
            return t

        def param(t, shape, dtype, name):
            if t is None:
                raise ValueError(f"{name} is required")
            if tuple(t.shape) != shape or t.dtype != dtype:
                raise ValueError(f"{name} must be {dtype} of shape {shape}, got {t.dtype} {tuple(t.shape)}")
            if nv.require_cuda(t, any_dtype=True) != dev:
                raise nv.NativeError(f"{name} is on {t.device}, the images on {dev}")
            return t.contiguous()

        crop_plain = param(crop_plain, (b, 2), torch.int32, "crop_plain")
        plain = dest(out_plain)
        aug = None
        if want_aug:
            crop_aug = param(crop_aug, (b, 2), torch.int32, "crop_aug")
            flip = param(flip, (b,), torch.uint8, "flip")
            affine = param(affine, (b, 2, 3), torch.float32, "affine")
            aug = dest(out_aug)
        with torch.cuda.device(dev):
            rc = nv.lib().dwt_augment_pair(nv.ptr(images), b, sh, sw, self.crop, nv.ptr(crop_plain),
                                           nv.ptr(crop_aug) if want_aug else None, nv.ptr(flip) if want_aug else None,
                                           nv.ptr(affine) if want_aug else None, self._mean, self._std, nv.ptr(plain),
                                           nv.ptr(aug) if want_aug else None, nv.LAYOUT_NHWC if channels_last else 0,
                                           nv.stream_ptr(dev))
        nv.check(rc)
        return plain, aug

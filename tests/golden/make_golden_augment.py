"""Golden fixture for the paired target augmentation, produced by the UNMODIFIED reference functions.

    python tests/golden/make_golden_augment.py        # needs /root/reference, cv2, torchvision, PIL

Runs the reference's own `_random_affine_augmentation` / `_gaussian_blur`
(resnet50_dwt_mec_officehome.py:481-492) inside the torchvision pipeline of :526-542, with the random draws made
explicit: the four np.random.normal draws of the affine matrix are replayed from the same seed, crop corners and
flips are chosen here and applied through torchvision's functional API.  Writes tests/golden/augment.npz.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
warnings.filterwarnings("ignore")
os.chdir(REF)
sys.path.insert(0, REF)
sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
import resnet50_dwt_mec_officehome as ref     # noqa: E402  (the experiment script, unmodified)
from PIL import Image                         # noqa: E402
import torchvision.transforms as T            # noqa: E402
import torchvision.transforms.functional as TF  # noqa: E402

assert ref.__file__.startswith(REF)
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def smooth_image(rng, h, w):
    """Natural-image-like content: low-frequency colour gradients + texture, uint8."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 100 * np.sin(xx / (5 + 3 * k) + rng.random() * 6) * np.cos(yy / (7 + 2 * k) + rng.random() * 6)
                    for k in range(3)], axis=-1)
    img += rng.normal(0, 12, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(11)
    cases = {}
    # (name, resized size, crop, batch, affine scale): the reference's sigma 0.1, a stronger warp, a tiny image
    for name, size, crop, B, scale in [("ref", 40, 32, 4, 0.1), ("strong", 36, 32, 3, 0.35), ("tiny", 9, 7, 3, 0.1)]:
        imgs = np.stack([smooth_image(rng, size, size) for _ in range(B)])
        crop_plain = rng.integers(0, size - crop + 1, (B, 2)).astype(np.int32)
        crop_aug = rng.integers(0, size - crop + 1, (B, 2)).astype(np.int32)
        flip = (rng.random(B) < 0.5).astype(np.uint8)
        flip[0], flip[-1] = 1, 0
        mats, plains, augs = [], [], []
        for b in range(B):
            pil = Image.fromarray(imgs[b])
            # plain view: :526-531
            p = TF.normalize(TF.to_tensor(TF.crop(pil, int(crop_plain[b, 0]), int(crop_plain[b, 1]), crop, crop)), MEAN, STD)
            # augmented view: :534-542 with the draws pinned
            a = TF.crop(pil, int(crop_aug[b, 0]), int(crop_aug[b, 1]), crop, crop)
            if flip[b]:
                a = TF.hflip(a)
            a = TF.to_tensor(a)
            seed = 1000 + 17 * b
            np.random.seed(seed)
            if scale == 0.1:
                n = [np.random.normal(0.0, 0.1) for _ in range(4)]      # the draws the reference is about to make
                np.random.seed(seed)
                a = ref._random_affine_augmentation(a)                   # unmodified reference function
            else:                                                         # same function body, wider draws
                n = [np.random.normal(0.0, scale) for _ in range(4)]
                import cv2
                M = np.float32([[1 + n[0], n[1], 0], [n[2], 1 + n[3], 0]])
                a = torch.from_numpy(np.transpose(cv2.warpAffine(np.transpose(a.numpy(), [1, 2, 0]), M, (crop, crop)), [2, 0, 1]))
            mats.append(np.float32([[1 + n[0], n[1], 0], [n[2], 1 + n[3], 0]]))
            a = ref._gaussian_blur(a)                                    # unmodified reference function
            a = TF.normalize(a, MEAN, STD)
            plains.append(p.numpy()); augs.append(a.numpy())
        for k, v in dict(images=imgs, crop_plain=crop_plain, crop_aug=crop_aug, flip=flip, affine=np.stack(mats),
                         plain=np.stack(plains), aug=np.stack(augs), crop=np.int32(crop)).items():
            cases[f"{name}/{k}"] = v
    import cv2
    cases["versions"] = np.array([f"cv2 {cv2.__version__}", f"torch {torch.__version__}"])
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **cases)
    print("wrote", os.path.join(HERE, "augment.npz"), os.path.getsize(os.path.join(HERE, "augment.npz")), "bytes")


if __name__ == "__main__":
    main()

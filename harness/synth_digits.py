"""Synthetic USPS / MNIST files in the exact formats usps_mnist.py reads: there is
no network, so BASELINE.json configs[0] ("USPS->MNIST, 1 epoch, plumbing") runs on these."""
from __future__ import annotations

import gzip
import os
import pickle

import numpy as np
import torch


def write_digit_files(data_root: str, seed: int = 0, usps=(256, 64), mnist=(512, 128)) -> None:
    """data_root/usps/usps_28x28.pkl  : gzip pickle [[train_imgs (n,1,28,28) float, labels], [test ...]]
    (usps_mnist.py:106-120);  data_root/mnist/processed/{training,test}.pt : (uint8 [n,28,28], int64 [n])
    (usps_mnist.py:153,163-167)."""
    rng = np.random.default_rng(seed)

    def digits(n):
        labels = rng.integers(0, 10, n)
        imgs = rng.random((n, 28, 28)) * 0.2
        for i, l in enumerate(labels):                     # a class-dependent blob so that 1 epoch can learn something
            r, c = 4 + 2 * (l // 5) * 5, 3 + (l % 5) * 4
            imgs[i, r:r + 8, c:c + 6] += 0.8
        return np.clip(imgs, 0, 1), labels

    os.makedirs(os.path.join(data_root, "usps"), exist_ok=True)
    os.makedirs(os.path.join(data_root, "mnist", "processed"), exist_ok=True)
    sets = []
    for n in usps:
        imgs, labels = digits(n)
        sets.append([imgs.reshape(n, 1, 28, 28).astype(np.float32), labels.astype(np.int64)])
    with gzip.open(os.path.join(data_root, "usps", "usps_28x28.pkl"), "wb") as f:
        pickle.dump(sets, f)
    for n, name in zip(mnist, ("training.pt", "test.pt")):
        imgs, labels = digits(n)
        torch.save((torch.from_numpy((imgs * 255).astype(np.uint8)), torch.from_numpy(labels.astype(np.int64))),
                   os.path.join(data_root, "mnist", "processed", name))


def digit_batches(seed: int, steps: int, batch: int = 32):
    """In-memory (source images, source labels, target images) batches of the same synthetic digits."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(steps):
        lab = rng.integers(0, 10, batch)
        def make(labels):
            x = rng.random((len(labels), 1, 28, 28)) * 0.2
            for i, l in enumerate(labels):
                r, c = 4 + 2 * (l // 5) * 5, 3 + (l % 5) * 4
                x[i, 0, r:r + 8, c:c + 6] += 0.8
            return torch.tensor((x - 0.5) / 0.5, dtype=torch.float32)
        out.append((make(lab), torch.tensor(lab), make(rng.integers(0, 10, batch))))
    return out

"""LeNet-DWT measurement harness for BASELINE.json configs[0] (USPS -> MNIST plumbing run).

Topology of the reference's LeNet (/root/reference/usps_mnist.py:196-278) re-stated over a `layers`
namespace (the product package on a GPU, the CPU port in CPU tests): two domains (source | target),
whitening after conv1 (32 ch, 28x28) and conv2 (48 ch, 14x14), stock nn.BatchNorm1d on the FC layers,
one shared gamma/beta per site; eval uses the target branch.  Parameter names match the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class LeNetDWT(nn.Module):
    def __init__(self, layers, group_size=4):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 32, kernel_size=5, padding=2)
        self.conv2 = nn.Conv2d(32, 48, kernel_size=5, padding=2)
        self.fc3, self.fc4, self.fc5 = nn.Linear(2352, 100), nn.Linear(100, 100), nn.Linear(100, 10)
        for i, c in ((1, 32), (2, 48)):
            setattr(self, f"ws{i}", layers.WTransform2d(num_features=c, group_size=group_size))
            setattr(self, f"wt{i}", layers.WTransform2d(num_features=c, group_size=group_size))
            setattr(self, f"gamma{i}", nn.Parameter(torch.ones(c, 1, 1)))
            setattr(self, f"beta{i}", nn.Parameter(torch.zeros(c, 1, 1)))
        for i, c in ((3, 100), (4, 100), (5, 10)):
            setattr(self, f"bns{i}", nn.BatchNorm1d(c, affine=False))
            setattr(self, f"bnt{i}", nn.BatchNorm1d(c, affine=False))
            setattr(self, f"gamma{i}", nn.Parameter(torch.ones(1, c)))
            setattr(self, f"beta{i}", nn.Parameter(torch.zeros(1, c)))

    def _site(self, i, x, prefix):
        src, tgt = getattr(self, f"{prefix}s{i}"), getattr(self, f"{prefix}t{i}")
        if self.training:
            a, b = torch.split(x, x.shape[0] // 2, dim=0)
            y = torch.cat((src(a), tgt(b)), dim=0)
        else:
            y = tgt(x)
        return y * getattr(self, f"gamma{i}") + getattr(self, f"beta{i}")

    def forward(self, x):
        x = F.max_pool2d(F.relu(self._site(1, self.conv1(x), "w")), 2, 2)
        x = F.max_pool2d(F.relu(self._site(2, self.conv2(x), "w")), 2, 2)
        x = torch.flatten(x, 1)
        x = F.relu(self._site(3, self.fc3(x), "bn"))
        x = F.relu(self._site(4, self.fc4(x), "bn"))
        return self._site(5, self.fc5(x), "bn")


def entropy_loss(logits):
    """usps_mnist.py:183-194."""
    return -(F.softmax(logits, dim=1) * F.log_softmax(logits, dim=1)).sum(-1).mean()


def train_epoch(model, optimizer, batches, device, lambda_entropy=0.1):
    """usps_mnist.py:281-308 on in-memory batches; returns the list of (cls_loss, entropy_loss)."""
    model.train()
    log = []
    for src, labels, tgt in batches:
        data = torch.cat((src, tgt), dim=0).to(device)
        optimizer.zero_grad()
        out = model(data)
        so, to = torch.split(out, out.shape[0] // 2, dim=0)
        cls = F.nll_loss(F.log_softmax(so, dim=1), labels.to(device))
        ent = lambda_entropy * entropy_loss(to)
        (cls + ent).backward()
        optimizer.step()
        log.append((cls.item(), ent.item()))
    return log

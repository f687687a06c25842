"""Drop-in for the reference's ``utils/consensus_loss.py``: one fused kernel launch computes
both log-softmaxes, the per-sample min over classes, the batch mean and both gradients.

Reference: /root/reference/utils/consensus_loss.py:5-24.
"""
from __future__ import annotations

import torch.nn as nn

from . import functional as F


class MinEntropyConsensusLoss(nn.Module):
    def __init__(self, num_classes, device):
        super().__init__()
        self.num_classes = num_classes
        self.device = device

    def forward(self, x, y):
        return F.mec_loss(x, y)

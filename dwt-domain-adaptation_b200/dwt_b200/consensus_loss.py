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


class HeadLoss(nn.Module):
    """Extension: the whole head of the Office-Home training step,
    ``nll_loss(log_softmax(source), y) + lambda * MEC(target, target_aug)``
    (resnet50_dwt_mec_officehome.py:421-428), as ONE kernel launch producing the loss, its two parts
    and the gradient of all 3B logit rows.  ``forward(logits [3B,K], labels [B])`` returns the total
    loss; ``.parts`` holds the detached tensor [total, classification, lambda*MEC] of the last call."""

    def __init__(self, num_classes, lambda_mec=0.1):
        super().__init__()
        self.num_classes, self.lambda_mec = num_classes, lambda_mec
        self.parts = None

    def forward(self, logits, labels):
        total, self.parts = F.head_loss(logits, labels, self.lambda_mec)
        return total

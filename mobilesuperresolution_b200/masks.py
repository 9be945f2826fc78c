"""Learnable binary channel masks of the NAS supernet (host-side mirror of models/ops.py).

In the reference a mask is a depthwise 1x1 conv multiplying every channel by {0,1} on every forward
(models/ops.py:18-26).  Here masks never touch the activations: ``keep_indices`` resolves them once, at
prepare time, into the pruned (IN, M1, M2) filter slices the CUDA kernels run.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.init as init


def rounding(weight: torch.Tensor, least_channel: int = 8) -> torch.Tensor:
    """Keep-mask of a mask weight: ``w >= 0.5`` unless fewer than ``least_channel`` survive, then the
    ``least_channel`` largest.  Same contract as models/ops.py:33-43."""
    keep = (weight >= 0.5).float()
    if least_channel <= 0 or float(keep.sum()) >= least_channel:
        return keep
    kth = torch.topk(weight, least_channel, dim=0).values[-1]
    return (weight >= kth).float()


class BinaryConv2d(nn.Conv2d):
    """Parameter container with the reference's constructor and state_dict entry (``weight`` (C,1,1,1),
    U(0.5,1) init; models/ops.py:7-16).  Its forward is never part of the CUDA path."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, groups=1,
                 bias=False, least_channel=8):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        init.uniform_(self.weight, 0.5, 1)
        self.least_channel = least_channel

    def init(self, value=0.5):
        init.constant_(self.weight, value)

    def keep_indices(self) -> torch.Tensor:
        keep = rounding(self.weight.detach().float().cpu(), self.least_channel).view(-1)
        return torch.nonzero(keep > 0, as_tuple=False).view(-1)

    def forward(self, x, y=None):
        raise RuntimeError("BinaryConv2d masks are folded into filter slices at prepare time; "
                           "the B200 path never applies them to activations")

"""Drop-in for the reference's Metrics/losses.py (CrossEntropyLoss, SoftDiceLoss,
hierarchical_consistency_loss) over the fused sm_100a loss kernels.

Conventions kept from the reference (Metrics/losses.py:16-177): `class_weight` is a Python
list of K floats; targets are ternary {1, 0, -1} with -1 = ignore and may be a channel slice
of a wider tensor; CE maps a NaN sample (any class without valid pixels) to the constant 1.0;
Dice drops NaN samples and returns None when none is left; both return 0-dim differentiable
fp32 tensors."""
import torch.nn as nn

from rhseg_b200 import loss_ops


class SoftDiceLoss(nn.Module):
    def __init__(self, smooth=0, num_classes=3):
        super().__init__()
        self.smooth = smooth

    def forward(self, outs, targets, logits_input=False, class_weight=None):
        res = loss_ops.level_loss(outs, targets, class_weight, float(self.smooth), logits_input)
        return res.dice_or_none()


class CrossEntropyLoss(nn.Module):
    def __init__(self, smooth=0.0):
        super().__init__()
        self.smooth = smooth

    def forward(self, outs, targets, logits_input=False, class_weight=None):
        # smooth=None: share the statistics pass with whichever dice evaluation sees the same tensors
        return loss_ops.level_loss(outs, targets, class_weight, None, logits_input).ce


def hierarchical_consistency_loss(probs_per_level, levels, parent_of, reduction="mean"):
    """mean_{b,h,w} |sum_children P_c - P_parent| averaged over (level, parent) pairs."""
    return loss_ops.consistency_loss(probs_per_level, levels, parent_of, reduction)

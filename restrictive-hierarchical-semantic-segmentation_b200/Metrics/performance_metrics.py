"""Drop-in for the reference's Metrics/performance_metrics.py.

The five wrappers share ONE integer confusion-matrix pass per (prediction, target) pair
(memoised on tensor identity) instead of five argmax + torchmetrics evaluations
(performance_metrics.py:52-141).  Semantics restated from torchmetrics' multiclass metrics
with average=None: F1, Jaccard, Accuracy (= per-class recall), Precision, Recall; child levels
get a prepended background class that is then ignored (ignore_index=0) and sliced off."""
import torch

from rhseg_b200 import metric_ops


class ProcessClasses(torch.nn.Module):
    """(probs, targets) -> two float [B,H,W] class-index maps (performance_metrics.py:31-47).
    Kept for API compatibility; the metric wrappers below do not materialise these maps."""

    def forward(self, probs, targets, child_classes=False):
        if child_classes:
            probs = torch.cat([(probs.sum(dim=1, keepdim=True) == 0).float(), probs], dim=1)
            targets = torch.cat([(targets.sum(dim=1, keepdim=True) == 0).float(), targets], dim=1)
        return probs.argmax(dim=1).float(), targets.argmax(dim=1).float()


class _ConfusionMetric(torch.nn.Module):
    row = None

    def __init__(self, smooth=1):
        super().__init__()
        self.process_classes = ProcessClasses()

    def forward(self, probs, targets, device, num_classes, child_classes=False):
        if probs.shape[1] != num_classes:
            raise ValueError("num_classes=%d does not match the %d prediction channels" % (num_classes, probs.shape[1]))
        r = metric_ops.level_ratios(probs, targets, bool(child_classes))[metric_ops.ROW[self.row]]
        r = r[1:] if child_classes else r
        return r.to(device)


class DiceScore(_ConfusionMetric):
    row = "dice"


class Jaccardindex(_ConfusionMetric):
    row = "iou"


class Accuracy(_ConfusionMetric):
    row = "accuracy"


class Precision(_ConfusionMetric):
    row = "precision"


class Recall(_ConfusionMetric):
    row = "recall"

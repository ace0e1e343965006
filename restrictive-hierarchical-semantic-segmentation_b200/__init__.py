"""rhseg_b200 — B200-native (sm_100a) restrictive-hierarchy head, loss and metric path.

Importable as `rhseg_b200` (alias module at the repository root).  The sub-packages
`Models/`, `Metrics/` and `tree_util.py` mirror the reference's module layout so that putting
this directory on sys.path ahead of the reference makes train.py / predictEval.py pick them up
unchanged."""
from . import native  # noqa: F401
from .tree_tables import ClassTree, build_hierarchy_indices, compiled_tree, get_level_classes  # noqa: F401
from .head import hier_head_forward  # noqa: F401
from .loss_ops import clear_memo, consistency_loss, level_loss  # noqa: F401
from .fused import FusedFlatStep, FusedHierStep, StepOutput  # noqa: F401
from .metric_ops import (concat_image_logits, confusion_from_logits, confusion_matrix, level_ratios,  # noqa: F401
                         predict_onehot, ratios, stitch_flat_to_levels)

__version__ = "0.1.0"

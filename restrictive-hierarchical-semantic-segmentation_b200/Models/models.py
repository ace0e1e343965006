"""Drop-in for the reference's Models/models.py hierarchical wrappers.

Same public names, constructor / forward signatures, return conventions and parameter names
(`heads.{L}.conv.*`, `classifiers.{L}.*`, `films.{i}.mlp.1.*`) as the reference
(Models/models.py:189-306 UNet, :554-832 HighResolutionNet), so train.py / predictEval.py and
published checkpoints keep working.  The donor backbones run on stock PyTorch; everything
after the last feature map (FiLM, 1x1 heads, upsample, restrictive softmax, composition, and
their backward) is one autograd node over the sm_100a kernels in csrc/.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from rhseg_b200 import donors, native
from rhseg_b200.donors import (BasicBlock, Bottleneck, HighResolutionModule, blocks_dict, conv3x3,  # noqa: F401
                               double_conv, down, inconv, outconv, up)
from rhseg_b200.head import hier_head_forward
from rhseg_b200.tree_tables import ClassTree, build_hierarchy_indices, get_level_classes  # noqa: F401

BN_MOMENTUM = donors.BN_MOMENTUM
ALIGN_CORNERS = None


class FiLM(nn.Module):
    """Parameter container (and stand-alone op) for the FiLM conditioner.  Inside the
    hierarchical models it is never run on its own: gamma/beta are folded into the following
    1x1 head conv by rhseg_film_fold.  Layout matches the reference: mlp = [Flatten, Linear]."""

    def __init__(self, feat_ch: int, cond_ch: int):
        super().__init__()
        self.cond_pool = nn.AdaptiveAvgPool2d(1)
        self.mlp = nn.Sequential(nn.Flatten(), nn.Linear(cond_ch, 2 * feat_ch))

    def forward(self, feats, cond_map):
        vec = self.cond_pool(cond_map).flatten(1) if cond_map.dim() == 4 else cond_map
        gamma, beta = self.mlp(vec).chunk(2, dim=1)
        return torch.addcmul(beta[:, :, None, None], feats, gamma[:, :, None, None])


def _leaf_count(hierarchy):
    return sum(len(v) for v in get_level_classes(hierarchy, inc_parent=False).values())


class _HierarchyMixin:
    """Shared construction / forward of the multi-level head for both donors."""

    # the reference re-runs the donor on the SAME input once per level (models.py:267, :277,
    # :757, :773).  In train() mode every pass updates the BatchNorm running statistics and owns its autograd
    # branch, so all passes are kept.  In eval() mode the passes are bit-identical (frozen statistics, no
    # dropout), so one pass is run and its feature map feeds every level: same outputs, 1/levels of the donor
    # cost.  Set to False to replay the reference's pass count exactly.
    reuse_backbone_in_eval = True

    def _level_features(self, run_backbone, x):
        n = self._tree.num_levels
        if not self.training and self.reuse_backbone_in_eval:
            f = run_backbone(x)
            return [f] * n
        return [run_backbone(x) for _ in range(n)]

    def _init_hierarchy(self, hierarchy, feat_ch, make_head, head_list_name):
        tree = ClassTree(hierarchy)
        self._tree = tree
        self.levels, self.parent_of, self.children_of = tree.levels, tree.parent_of, tree.children_of
        self.child_groups = tree.child_groups
        heads = nn.ModuleList(make_head(feat_ch, k) for k in tree.head_channels)
        setattr(self, head_list_name, heads)
        self.films = nn.ModuleList(FiLM(feat_ch=feat_ch, cond_ch=len(tree.levels[L - 1]))
                                   for L in range(1, tree.num_levels))

    def _head_forward(self, feats, head_convs, out_size=None):
        if not feats[0].is_cuda:
            raise native.NativeError("the hierarchical head runs on sm_100a kernels only: move the model "
                                     "and its input to a CUDA device (there is no CPU fallback)")
        lin = [f.mlp[1] for f in self.films]
        return hier_head_forward(self._tree, feats, [c.weight for c in head_convs], [c.bias for c in head_convs],
                                 [l.weight for l in lin], [l.bias for l in lin], out_size)


class UNet(nn.Module, _HierarchyMixin):
    """Flat (model_type == 0 or type == 0): returns ([], logits).
    Hierarchical: returns (probs_per_level, logits_per_level), lists of [B,K_L,H,W]."""

    def __init__(self, size=620, n_channels=1, hierarchy={}, model_type=0):
        super().__init__()
        self.model_type = model_type
        self.hierarchy = hierarchy
        donors.attach_unet_backbone(self, n_channels)
        if model_type == 0:
            self.out_flat = outconv(donors.UNET_FEATURES, _leaf_count(hierarchy))
        else:
            self._init_hierarchy(hierarchy, donors.UNET_FEATURES, outconv, "heads")

    def _run_unet(self, x):
        return donors.run_unet_backbone(self, x)

    def forward(self, x, type=0, hierarchy={}, threshold=0.5):
        if self.model_type == 0 or type == 0:
            return [], self.out_flat(self._run_unet(x))
        # one donor pass per level on the same input, exactly like the reference (:267, :277):
        # every level gets its own feature tensor, autograd branch and BN running-stat update
        feats = self._level_features(self._run_unet, x)
        return self._head_forward(feats, [h.conv for h in self.heads])


class HighResolutionNet(nn.Module, _HierarchyMixin):
    """HRNetV2-W48 donor with the same hierarchical head (logits bilinearly upsampled to the
    input size, align_corners from the config, before the activations)."""

    def __init__(self, config, hierarchy={}, model_type=0, **kwargs):
        super().__init__()
        global ALIGN_CORNERS
        extra = config.MODEL.EXTRA
        ALIGN_CORNERS = config.MODEL.ALIGN_CORNERS
        self.model_type = model_type
        self.hierarchy = hierarchy
        self.align_corners = bool(ALIGN_CORNERS)
        feat_ch = donors.attach_hrnet_backbone(self, extra, self.align_corners)
        k = extra["FINAL_CONV_KERNEL"]

        def classifier(cin, cout):
            return nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=1 if k == 3 else 0)

        if model_type == 0:
            self.classifier = classifier(feat_ch, _leaf_count(hierarchy))
        else:
            if k != 1 or not self.align_corners:
                raise native.NativeError("the fused hierarchical head covers the reference configuration "
                                         "(FINAL_CONV_KERNEL=1, ALIGN_CORNERS=True) only")
            self._init_hierarchy(hierarchy, feat_ch, classifier, "classifiers")

    def _forward_backbone(self, x):
        return donors.run_hrnet_backbone(self, x)

    def forward(self, x):
        size = (x.shape[-2], x.shape[-1])
        if self.model_type == 0:
            z = self.classifier(self._forward_backbone(x))
            return [], F.interpolate(z, size=size, mode="bilinear", align_corners=self.align_corners)
        feats = self._level_features(self._forward_backbone, x)
        return self._head_forward(feats, list(self.classifiers), out_size=size)

    def init_weights(self, pretrained="", device="cpu"):
        """Loads a checkpoint by exact or suffix key match with equal shapes (reference :804-832)."""
        ckpt = torch.load(pretrained, map_location=device)
        ckpt = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
        cleaned = {}
        for key, val in ckpt.items():
            for prefix in ("model.", "module.", "net.", "network."):
                if key.startswith(prefix):
                    key = key[len(prefix):]
            cleaned[key] = val
        own = self.state_dict()
        picked = {}
        for name, cur in own.items():
            hit = cleaned.get(name)
            if hit is not None and hit.size() == cur.size():
                picked[name] = hit
                continue
            for cname, cval in cleaned.items():
                if (name.endswith(cname) or cname.endswith(name)) and cval.size() == cur.size():
                    picked[name] = cval
                    break
        missing = [n for n in own if n not in picked]
        print(f"Loaded {len(picked)} / {len(own)} layers.")
        if missing:
            print(f"Missing {len(missing)} layers (first 10): {missing[:10]}")
        own.update(picked)
        self.load_state_dict(own)
        return self

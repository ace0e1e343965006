"""ORACLE — test infrastructure only, never product code.

CPU restatement (plain PyTorch fp32 ops, autograd for the gradients) of the
reference's restrictive-hierarchy head + hierarchical loss + confusion-matrix
metric path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
file; nothing under ``restrictive-hierarchical-semantic-segmentation_b200/``
does (tests/test_no_oracle_in_product.py enforces it).

Every function names the reference lines it follows (paths relative to the
reference root).  The op sequence mirrors the reference op for op so that the
timing of this port is representative of the reference's own CPU path.

Pinning status
  * head / loss / consistency / train-path prediction: PINNED against outputs
    of the unmodified reference imported in the build container
    (tests/golden/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).
  * metrics: PARITY UNPINNED.  The reference delegates to the un-vendored,
    unpinned third-party package ``torchmetrics`` (Metrics/performance_metrics.py:62
    etc.), which is not installed offline and for which the reference holds no
    tests.  ``multiclass_confusion`` / ``ratios_from_confusion`` restate its
    published multiclass semantics (int64 bincount confusion matrix, rows with
    target == ignore_index dropped, zero-division -> 0, Accuracy(average=None)
    == per-class recall) and are cross-checked against scikit-learn.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

GATE_EPS = 1e-6  # Models/models.py:272, :761


# --------------------------------------------------------------------------
# (a1) class tree -> per-level name tables
# --------------------------------------------------------------------------
def names_per_depth(tree: dict, parents_too: bool) -> Dict[int, List[str]]:
    """Models/models.py:82-98 (get_level_classes): pre-order walk, names bucketed by depth.
    parents_too=False keeps only leaves (flat mode)."""
    out: Dict[int, List[str]] = {}

    def walk(sub, depth):
        if not isinstance(sub, dict) or not sub:
            return
        for name, kids in sub.items():
            out.setdefault(depth, [])
            if parents_too or not kids:
                out[depth].append(name)
            if isinstance(kids, dict):
                walk(kids, depth + 1)

    walk(tree, 0)
    return out


def hierarchy_tables(tree: dict):
    """Models/models.py:38-54 (build_hierarchy_indices) + :229-238 (child_groups)."""
    by_depth = names_per_depth(tree, parents_too=True)
    levels = [by_depth[d] for d in sorted(by_depth)]
    parent_of, children_of = {}, {}

    def walk(sub, parent):
        for name, kids in sub.items():
            parent_of[name] = parent
            if isinstance(kids, dict) and len(kids) > 0:
                children_of[name] = list(kids.keys())
                walk(kids, name)
            else:
                children_of.setdefault(name, [])

    walk(tree, None)
    groups = []
    for L in range(1, len(levels)):
        groups.append([(p, children_of.get(p, [])) for p in levels[L - 1] if len(children_of.get(p, [])) > 0])
    return levels, parent_of, children_of, groups


# --------------------------------------------------------------------------
# (a2)-(a4) head forward
# --------------------------------------------------------------------------
def film_apply(feats: torch.Tensor, cond_map: torch.Tensor, film_w: torch.Tensor, film_b: torch.Tensor):
    """Models/models.py:67-77 (FiLM.forward): global-average-pool the previous level's
    probability map, one Linear, per-(sample, channel) affine."""
    cond = F.adaptive_avg_pool2d(cond_map, 1).flatten(1) if cond_map.dim() == 4 else cond_map
    gb = F.linear(cond, film_w, film_b)
    C = feats.size(1)
    return feats * gb[:, :C, None, None] + gb[:, C:, None, None]


def head_forward(feats_per_level: Sequence[torch.Tensor],
                 head_w: Sequence[torch.Tensor], head_b: Sequence[torch.Tensor],
                 film_w: Sequence[torch.Tensor], film_b: Sequence[torch.Tensor],
                 levels, groups, out_size: Optional[Tuple[int, int]] = None):
    """Models/models.py:263-306 (UNet) and :757-802 (HRNet; out_size => bilinear
    align_corners=True upsample of the logits, :766/:776).  ``feats_per_level[L]`` stands in
    for the L-th backbone pass (the backbone is re-run on the same x, :277/:773)."""

    def conv1x1(f, L):
        z = F.conv2d(f, head_w[L], head_b[L])
        if out_size is not None:
            z = F.interpolate(z, size=out_size, mode="bilinear", align_corners=True)
        return z

    probs, logits = [], []
    z0 = conv1x1(feats_per_level[0], 0)
    probs.append(torch.sigmoid(z0))
    logits.append(z0)
    for L in range(1, len(levels)):
        f = film_apply(feats_per_level[L], probs[L - 1], film_w[L - 1], film_b[L - 1])
        z = conv1x1(f, L)
        grp = groups[L - 1]
        if len(grp) == 0:
            probs.append(torch.zeros_like(z))
            logits.append(z)
            continue
        pieces, start = [], 0
        for pname, kids in grp:
            g = len(kids)
            zg = z[:, start:start + g]
            pi = levels[L - 1].index(pname)
            Pp = probs[L - 1][:, pi:pi + 1]
            Q = torch.softmax(zg + torch.log(Pp + GATE_EPS), dim=1)
            pieces.append(Pp * Q)
            start += g
        probs.append(torch.cat(pieces, dim=1))
        logits.append(z)
    return probs, logits


def flat_forward(feats, w, b, out_size=None):
    """Models/models.py:258-261 / :754-758: flat model = one 1x1 conv (+ upsample for HRNet)."""
    z = F.conv2d(feats, w, b)
    if out_size is not None:
        z = F.interpolate(z, size=out_size, mode="bilinear", align_corners=True)
    return z


# --------------------------------------------------------------------------
# (a5) losses
# --------------------------------------------------------------------------
def ce_loss(outs, targets, class_weight, logits_input=True, ignore=-1):
    """Metrics/losses.py:95-134.  Per sample, per class: boolean-mask gather, mean of
    -(t * logp * w); class mean; NaN sample -> 1.0; batch mean."""
    if logits_input:
        outs = F.log_softmax(outs, dim=1)
    B, K = outs.size(0), outs.size(1)
    o = outs.contiguous().view(B, K, -1)
    t = targets.contiguous().view(B, K, -1)
    keep = t != ignore
    w = torch.tensor(class_weight).unsqueeze(1).to(o.device)
    per_sample = []
    for b in range(B):
        acc = 0.0
        for c in range(K):
            sel = keep[b][c]
            acc = acc + (-(t[b][c][sel] * o[b][c][sel] * w[c]).mean())
        per_sample.append(torch.nan_to_num(acc / K, nan=1.0))
    return torch.stack(per_sample).mean()


def dice_loss(outs, targets, class_weight, logits_input=True, smooth=0.0, ignore=-1.0):
    """Metrics/losses.py:23-86.  Per sample ONE weighted dice over all classes (masked
    gather per class); NaN samples dropped; None when nothing is left."""
    if logits_input:
        outs = F.softmax(outs, dim=1)
    B, K = outs.size(0), outs.size(1)
    o = outs.contiguous().view(B, K, -1)
    t = targets.contiguous().view(B, K, -1)
    keep = t != ignore
    w = torch.tensor(class_weight).unsqueeze(1).to(o.device)
    per_sample = []
    for b in range(B):
        inter, union = 0.0, 0.0
        for c in range(K):
            sel = keep[b][c]
            oc, tc = o[b][c][sel], t[b][c][sel]
            inter = inter + (oc * tc * w[c]).sum()
            union = union + (oc * w[c]).sum() + (tc * w[c]).sum()
        score = (2 * inter + smooth) / (union + smooth)
        per_sample.append(1.0 - score.mean())
    kept = [l for l in per_sample if not torch.isnan(l)]
    return torch.stack(kept).mean() if kept else None


def consistency_loss(probs_per_level, levels, parent_of, reduction="mean"):
    """Metrics/losses.py:150-177: mean |sum_children P_c - P_p| per (level, parent), averaged."""
    total, count = 0.0, 0
    for L in range(1, len(levels)):
        prev, cur = probs_per_level[L - 1], probs_per_level[L]
        for pi, pname in enumerate(levels[L - 1]):
            idx = [i for i, c in enumerate(levels[L]) if parent_of.get(c, None) == pname]
            if not idx:
                continue
            diff = (cur[:, idx].sum(dim=1, keepdim=True) - prev[:, pi:pi + 1]).abs()
            total = total + (diff.mean() if reduction == "mean" else diff.sum())
            count += 1
    if count == 0:
        return probs_per_level[0].sum() * 0
    return total / count


def total_loss(logits_per_level, targets_per_level, level_weights, probs_per_level=None,
               levels=None, parent_of=None, cur_epoch=None, pretrain_epoch=None):
    """train.py:111-152 (get_loss) without the bookkeeping lists: sum_L (CE_L + Dice_L)
    [+ consistency], with the level-pretrain cap of :125-126/:133-134."""
    n = len(logits_per_level)
    cap = None
    if pretrain_epoch is not None:
        cap = int(min(n - 1, cur_epoch // pretrain_epoch))
    loss, per_level = 0.0, [0.0] * n
    for L in range(n):
        if cap is not None and L > cap:
            continue
        w = None if level_weights is None else level_weights[L]
        lce = ce_loss(logits_per_level[L], targets_per_level[L], w, True)
        ldi = dice_loss(logits_per_level[L], targets_per_level[L], w, True)
        if lce is not None:
            loss = loss + lce
            per_level[L] += lce.item()
        if ldi is not None:
            loss = loss + ldi
            per_level[L] += ldi.item()
    if probs_per_level is not None and levels is not None and parent_of is not None:
        loss = loss + consistency_loss(probs_per_level, levels, parent_of, "mean")
    return loss, per_level


# --------------------------------------------------------------------------
# (a7) train-path prediction glue
# --------------------------------------------------------------------------
def predict_onehot_masked(logits_per_level, targets_per_level):
    """train.py:206-231: one_hot(argmax(softmax(z))) as float, zeroed where the target is -1;
    eval targets are the targets with -1 replaced by 0."""
    preds, eval_t = [], []
    for z, t in zip(logits_per_level, targets_per_level):
        idx = torch.argmax(F.softmax(z, dim=1), dim=1)
        oh = F.one_hot(idx, num_classes=z.size(1)).permute(0, 3, 1, 2).float()
        preds.append(torch.where(t == -1, 0, oh))
        eval_t.append(torch.where(t == -1, 0, t))
    return preds, eval_t


# --------------------------------------------------------------------------
# (a8) metrics
# --------------------------------------------------------------------------
def process_classes(probs, targets, child_classes):
    """Metrics/performance_metrics.py:31-47: argmax, with a prepended 'nothing positive'
    background channel for child levels."""
    if child_classes:
        probs = torch.cat([(probs.sum(dim=1, keepdim=True) == 0).float(), probs], dim=1)
        targets = torch.cat([(targets.sum(dim=1, keepdim=True) == 0).float(), targets], dim=1)
    return torch.argmax(probs, dim=1).float(), torch.argmax(targets, dim=1).float()


def multiclass_confusion(pred_idx, tgt_idx, num_classes, ignore_index):
    """torchmetrics (un-vendored) multiclass stat-scores update, average=None path:
    drop rows whose target == ignore_index, bincount(target * nc + pred) -> [nc, nc] int64
    (row = target, column = prediction)."""
    p = pred_idx.flatten().to(torch.long)
    t = tgt_idx.flatten().to(torch.long)
    if ignore_index is not None:
        sel = t != ignore_index
        p, t = p[sel], t[sel]
    return torch.bincount(t * num_classes + p, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def _safe_div(num, den):
    num = num if num.is_floating_point() else num.float()
    den = den if den.is_floating_point() else den.float()
    den = torch.where(den == 0, torch.ones_like(den), den)
    return num / den


def ratios_from_confusion(conf):
    """torchmetrics reductions with average=None, zero_division=0:
    F1 = 2tp/(2tp+fn+fp); Jaccard = tp/(rowsum+colsum-tp); Accuracy = Recall = tp/(tp+fn);
    Precision = tp/(tp+fp).  Arithmetic order kept (int64 -> f32, then the ratio)."""
    tp = conf.diag()
    fp = conf.sum(0) - tp
    fn = conf.sum(1) - tp
    return {
        "dice": _safe_div(2.0 * tp, 2.0 * tp + 1.0 * fn + fp),
        "iou": _safe_div(tp, conf.sum(0) + conf.sum(1) - tp),
        "accuracy": _safe_div(tp, tp + fn),
        "precision": _safe_div(tp, tp + fp),
        "recall": _safe_div(tp, tp + fn),
    }


def level_confusion(probs, targets, num_classes, child_classes):
    """performance_metrics.py:59-66 etc.: the (K or K+1)^2 confusion matrix every one of the
    five wrappers rebuilds."""
    p, t = process_classes(probs, targets, child_classes)
    if child_classes:
        return multiclass_confusion(p, t, num_classes + 1, 0)
    return multiclass_confusion(p, t, num_classes, -1)


def level_metrics(probs, targets, num_classes, child_classes):
    """One dict of the five [num_classes] f32 vectors the wrappers return."""
    r = ratios_from_confusion(level_confusion(probs, targets, num_classes, child_classes))
    if child_classes:
        r = {k: v[1:] for k, v in r.items()}
    return r


def all_level_metrics(outputs, targets):
    """train.py:38-51 (get_metrics): per-level vectors concatenated in level order."""
    cat = {k: [] for k in ("iou", "accuracy", "dice", "precision", "recall")}
    for L, (o, t) in enumerate(zip(outputs, targets)):
        r = level_metrics(o, t, t.shape[1], L != 0)
        for k in cat:
            cat[k].append(r[k])
    return {k: torch.cat(v) for k, v in cat.items()}


# --------------------------------------------------------------------------
# (f3) flat -> hierarchy stitching
# --------------------------------------------------------------------------
def stitch_flat_to_levels(flat, tree: dict):
    """predictEval.py:85-185 (get_parent_masks + combine_levels) as predict() calls them (:381-388):
    flat leaf channels (breadth-first leaf order) -> per-level tensors; a parent is the union
    (any > 0) of its descendant leaves."""
    from collections import deque
    children = {}
    stack = [tree]
    while stack:
        t = stack.pop()
        for k, v in t.items():
            children[k] = list(v.keys()) if isinstance(v, dict) and len(v) > 0 else []
            if children[k]:
                stack.append(v)
    q, bfs, levels = deque((n, s, 0) for n, s in tree.items()), [], []
    while q:
        name, sub, d = q.popleft()
        bfs.append(name)
        if len(levels) <= d:
            levels.append([])
        levels[d].append(name)
        if isinstance(sub, dict) and len(sub) > 0:
            q.extend((cn, cs, d + 1) for cn, cs in sub.items())
    leaf_idx = {n: i for i, n in enumerate(n for n in bfs if not children[n])}

    def leaves_under(n):
        return [n] if not children[n] else [l for c in children[n] for l in leaves_under(c)]

    out = []
    for names in levels:
        chans = []
        for n in names:
            if not children[n]:
                chans.append(flat[:, leaf_idx[n]:leaf_idx[n] + 1])
            else:
                idx = sorted(set(leaf_idx[l] for l in leaves_under(n)))
                chans.append((flat[:, idx] > 0).any(dim=1, keepdim=True).to(flat.dtype))
        out.append(torch.cat(chans, dim=1))
    return out


# --------------------------------------------------------------------------
# synthetic inputs — shared by tests, smoke and bench
# --------------------------------------------------------------------------
def synth_targets(levels, groups, B, H, W, gen: torch.Generator, blobs: bool = False, device="cpu"):
    """Ternary {1,0,-1} targets per level following Data/dataset.py:227-265 (process_ignore_values).  The generator lives
    in the neutral module tools/synth.py (bench.py's product arm uses it without importing the oracle)."""
    from tools import synth
    return synth.synth_targets(levels, groups, B, H, W, gen, blobs=blobs, device=device)

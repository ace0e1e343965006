"""Confusion-matrix metric operators over the C-ABI kernels (memoised per tensor pair so the
five wrappers of Metrics/performance_metrics.py share one pass)."""
import weakref

import torch

from . import native
from .native import call, ptr, stream_of

ROW = {"dice": 0, "iou": 1, "accuracy": 2, "precision": 3, "recall": 4}


def _strided(t):
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 4:
        raise native.NativeError("metrics expect [B,K,H,W] tensors")
    if not (t.stride(3) == 1 and t.stride(2) == t.size(3)):
        t = t.contiguous()
    return t


def confusion_matrix(probs, targets, child_classes: bool) -> torch.Tensor:
    """int64 [nc,nc] (row = target, col = prediction); nc = K (+1 with the background class)."""
    native.require_cuda(probs, targets)
    p, t = _strided(probs), _strided(targets)
    B, K = p.shape[0], p.shape[1]
    n_pix = p.shape[2] * p.shape[3]
    if tuple(t.shape) != tuple(p.shape):
        raise native.NativeError("prediction / target shape mismatch: %s vs %s" % (tuple(p.shape), tuple(t.shape)))
    nc = K + 1 if child_classes else K
    conf = torch.empty((nc, nc), dtype=torch.int64, device=p.device)
    with native.device_guard(p):
        call("rhseg_confusion_matrix", ptr(p), p.stride(0), p.stride(1), ptr(t), t.stride(0), t.stride(1),
             B, K, n_pix, 1 if child_classes else 0, ptr(conf), stream_of(p))
    return conf


def confusion_from_logits(logits, targets, child_classes: bool) -> torch.Tensor:
    """Same matrix as predict_onehot + confusion_matrix, straight from logits and ternary
    targets (train.py:206-232 fused; no one-hot tensors)."""
    native.require_cuda(logits, targets)
    z = logits if logits.is_contiguous() else logits.contiguous()
    t = _strided(targets)
    B, K = z.shape[0], z.shape[1]
    n_pix = z.shape[2] * z.shape[3]
    nc = K + 1 if child_classes else K
    conf = torch.empty((nc, nc), dtype=torch.int64, device=z.device)
    with native.device_guard(z):
        call("rhseg_confusion_from_logits", ptr(z), ptr(t), t.stride(0), t.stride(1), B, K, n_pix,
             1 if child_classes else 0, ptr(conf), stream_of(z))
    return conf


def ratios(conf: torch.Tensor) -> torch.Tensor:
    """fp32 [5,nc]: rows dice/F1, IoU, accuracy(=recall), precision, recall."""
    nc = conf.shape[0]
    out = torch.empty((5, nc), dtype=torch.float32, device=conf.device)
    with native.device_guard(conf):
        call("rhseg_metric_ratios", ptr(conf), nc, ptr(out), stream_of(conf))
    return out


def predict_onehot(logits, targets, want_index=False):
    """train.py:206-231: (one_hot(argmax(softmax(z))) zeroed where t == -1, targets with -1 -> 0)."""
    native.require_cuda(logits, targets)
    z = logits if logits.is_contiguous() else logits.contiguous()
    t = _strided(targets)
    B, K = z.shape[0], z.shape[1]
    n_pix = z.shape[2] * z.shape[3]
    onehot = torch.empty_like(z)
    eval_t = torch.empty_like(z)
    idx = torch.empty((B,) + tuple(z.shape[2:]), dtype=torch.int32, device=z.device) if want_index else None
    with native.device_guard(z):
        call("rhseg_predict_onehot", ptr(z), ptr(t), t.stride(0), t.stride(1), B, K, n_pix, ptr(onehot), ptr(eval_t),
             ptr(idx), stream_of(z))
    return (onehot, eval_t, idx) if want_index else (onehot, eval_t)


_MEMO = []
_MEMO_MAX = 16


def level_ratios(probs, targets, child_classes: bool) -> torch.Tensor:
    """Memoised ratios(confusion_matrix(...)) keyed on tensor identity + version."""
    for i in range(len(_MEMO) - 1, -1, -1):
        p_ref, p_ver, t_ref, t_ver, child, res = _MEMO[i]
        p, t = p_ref(), t_ref()
        if p is None or t is None:
            del _MEMO[i]
            continue
        if p is probs and t is targets and p_ver == probs._version and t_ver == targets._version and child == bool(child_classes):
            return res
    res = ratios(confusion_matrix(probs, targets, child_classes))
    _MEMO.append((weakref.ref(probs), probs._version, weakref.ref(targets), targets._version, bool(child_classes), res))
    if len(_MEMO) > _MEMO_MAX:
        del _MEMO[0]
    return res


def stitch_flat_to_levels(flat, tree):
    """Flat model output / target [B, n_leaves, H, W] (leaf channels in breadth-first order) -> list of
    per-level tensors [B, K_L, H, W] with every parent filled as the union of its descendant leaves
    (predictEval.get_parent_masks + combine_levels, predictEval.py:85-185)."""
    import ctypes
    native.require_cuda(flat)
    x = flat if flat.dtype == torch.float32 else flat.float()
    x = x if x.is_contiguous() else x.contiguous()
    B, nl, H, W = x.shape
    masks, counts = tree.stitch_masks()
    if nl != len(tree.leaf_order()):
        raise native.NativeError("flat tensor has %d channels, the tree has %d leaves" % (nl, len(tree.leaf_order())))
    out = torch.empty((B, len(masks), H, W), dtype=torch.float32, device=x.device)
    arr = (ctypes.c_uint32 * len(masks))(*masks)
    call("rhseg_stitch_levels", ptr(x), B, nl, H * W, arr, len(masks), ptr(out), stream_of(x))
    levels, s = [], 0
    for k in counts:
        levels.append(out[:, s:s + k])
        s += k
    return levels


def concat_image_logits(image, logits):
    """cat([image, logits], dim=1) as one kernel (stand-alone utility; not part of the reference forward)."""
    native.require_cuda(image, logits)
    a = image.float().contiguous()
    b = logits.float().contiguous()
    if a.shape[0] != b.shape[0] or tuple(a.shape[2:]) != tuple(b.shape[2:]):
        raise native.NativeError("image %s and logits %s do not share batch / spatial size" % (tuple(a.shape), tuple(b.shape)))
    B, ca, H, W = a.shape
    out = torch.empty((B, ca + b.shape[1], H, W), dtype=torch.float32, device=a.device)
    call("rhseg_concat_image_logits", ptr(a), ca, ptr(b), b.shape[1], B, H * W, ptr(out), stream_of(a))
    return out

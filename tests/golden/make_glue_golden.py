"""Generates tests/golden/glue_*.npz: the UNMODIFIED reference's train.train_epoch (train.py:161-279) run on CPU over a
tiny synthetic loader, with the reference's own Models / Metrics.losses.  The GPU test (tests/test_gpu_reference_glue.py)
runs the SAME train.train_epoch on top of the drop-in Models / Metrics on CUDA and compares everything it returns and
the parameters after the optimiser steps.

    python tests/golden/make_glue_golden.py

Two deviations from "unmodified", both forced, identical in the GPU test:
  * train.get_loss is called through a wrapper that drops the two keyword arguments its signature does not have
    (train.py:239 passes lambda_cons / lambda_kl; the shipped script raises TypeError there);
  * torchmetrics is not installable offline: the five metric objects handed to train_epoch are the restated ones of
    oracle/hier_oracle.py (same call signature as Metrics/performance_metrics.py:52-141).
The donor backbone is replaced by one seeded 3x3 convolution registered on the model (its weights travel in the
fixture; a full UNet state dict would be 54 MB): gradients still flow through the head into donor parameters.
"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import hier_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

W_TL = [[0.0297, 1.577, 0.9619, 0.1770], [1.5432, 0.2638, 1.0413, 3.9722]]  # README.md:71


class OracleMetric:
    """Metrics/performance_metrics.py wrapper signature over the restated torchmetrics slice."""

    def __init__(self, key):
        self.key = key

    def __call__(self, probs, targets, device, num_classes, child_classes=False):
        conf = O.level_confusion(probs.detach().cpu(), targets.detach().cpu(), num_classes, bool(child_classes))
        r = O.ratios_from_confusion(conf)[self.key]
        return (r[1:] if child_classes else r).to(device)


class Loader(list):
    """train_epoch only needs len(), iteration and .dataset (for the progress print)."""

    @property
    def dataset(self):
        return [None] * sum(len(d) for d, _ in self)


def case_inputs(tree, n_batches, B, H, W, seed, notooth_sample=None):
    levels, parent_of, _, groups = O.hierarchy_tables(tree)
    gen = torch.Generator().manual_seed(seed)
    batches = []
    for i in range(n_batches):
        data = torch.randn(B, 3, H, W, generator=gen)
        target = torch.cat(O.synth_targets(levels, groups, B, H, W, gen, blobs=(i % 2 == 1)), dim=1)
        batches.append((data, target))
    return batches


def attach_tiny_donor(model, state=None):
    """One 3x3 convolution as the donor (registered: the optimiser sees it, dfeats flow into it)."""
    model.tiny_donor = nn.Conv2d(3, 64, 3, padding=1)
    if state is not None:
        model.tiny_donor.load_state_dict(state)
    model._run_unet = lambda x: torch.tanh(model.tiny_donor(x))
    return model


def run_epoch(train, model, losses_mod, metric_objs, batches, tree, device, pretrain, epoch_num, lr):
    orig = train.get_loss

    def get_loss_no_bad_kwargs(*a, lambda_cons=None, lambda_kl=None, **k):
        return orig(*a, **k)

    train.get_loss = get_loss_no_bad_kwargs
    try:
        args = types.SimpleNamespace(num_classes=[len(l) for l in model.levels], model_type=1, model_select=0, level_weights=W_TL,
                                     level0_pretrain_epochs=pretrain, batch_size=len(batches[0][0]))
        nL = len(model.levels)
        loss_fns = [[losses_mod.CrossEntropyLoss(), losses_mod.SoftDiceLoss()] for _ in range(nL)]
        params = [p for n, p in model.named_parameters() if n.startswith(("tiny_donor", "heads", "films"))]
        opt = torch.optim.SGD(params, lr=lr)
        acc, iou, dice, prec, rec = metric_objs
        out = train.train_epoch(model, device, Loader(batches), opt, 1, loss_fns, args, tree, None, acc, iou, dice, prec, rec, epoch_num)
    finally:
        train.get_loss = orig
    return out


def pack_result(out):
    loss, clss, acc, iou, dice, prec, rec, level_loss = out
    return dict(loss=np.float64(loss), accuracy=np.float64(acc), iou=np.float64(iou), dice=np.float64(dice), precision=np.float64(prec),
                recall=np.float64(rec), level_loss=np.asarray(level_loss, dtype=np.float64),
                class_metrics=np.asarray([[c[k] for k in ("accuracy", "iou", "dice", "precision", "recall")] for c in clss], dtype=np.float64))


def main():
    ap = argparse.ArgumentParser()
    ap.parse_args()
    rmodels, rlosses, rtrain = ref_loader.load_reference()
    tree = json.load(open(os.path.join(ref_loader.reference_root(), "class_tree_tl.json")))
    for name, pretrain, epoch_num in (("glue_unet_tl", None, 0), ("glue_unet_tl_curriculum", 2, 1)):
        torch.manual_seed(3)
        model = rmodels.UNet(size=48, n_channels=3, hierarchy=tree, model_type=1)
        attach_tiny_donor(model)
        with torch.no_grad():  # make FiLM matter
            for f in model.films:
                f.mlp[1].weight.mul_(3.0)
                f.mlp[1].bias.add_(torch.randn_like(f.mlp[1].bias) * 0.5 + 1.0)
        keep = {k: v.detach().clone() for k, v in model.state_dict().items() if k.startswith(("tiny_donor", "heads", "films"))}
        batches = case_inputs(tree, n_batches=3, B=2, H=40, W=48, seed=17)
        metric_objs = [OracleMetric(k) for k in ("accuracy", "iou", "dice", "precision", "recall")]
        out = run_epoch(rtrain, model, rlosses, metric_objs, batches, tree, "cpu", pretrain, epoch_num, lr=0.05)
        res = pack_result(out)
        after = {k: v.detach().clone() for k, v in model.state_dict().items() if k in keep}
        arrays = {"init." + k: v.numpy() for k, v in keep.items()}
        arrays.update({"final." + k: v.numpy() for k, v in after.items()})
        for i, (d, t) in enumerate(batches):
            arrays["data%d" % i] = d.numpy()
            arrays["target%d" % i] = t.numpy().astype(np.int8)
        arrays.update(res)
        arrays["meta"] = np.asarray(json.dumps(dict(tree=tree, pretrain=pretrain, epoch_num=epoch_num, lr=0.05, n_batches=len(batches))))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, "loss %.6f" % res["loss"], "dice %.4f" % res["dice"], "level_loss", res["level_loss"])


if __name__ == "__main__":
    main()

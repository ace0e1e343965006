"""Worker of tests/test_gpu_shapes.py::test_conv_backward_tile_groups: the conv backward's tile-group unit order
(RHSEG_TUNE_BWD_GROUP, read once per process -> set by the parent test) against the fp64 oracle on planes that span
several pixel tiles, with a partial last group, for 16-byte-aligned (UNet-like) and 4-byte-aligned (HRNet-like) planes."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from helpers import close  # noqa: E402
from oracle import hier_oracle as O  # noqa: E402
import rhseg_b200  # noqa: E402
from test_gpu_shapes import TL, _inputs  # noqa: E402


def main():
    assert int(os.environ.get("RHSEG_TUNE_BWD_GROUP", "0")) > 0
    for name, B, C, h, w in (("aligned_5tiles", 3, 24, 70, 60), ("odd_4tiles", 2, 40, 65, 61)):
        levels, parent_of, groups, chans, tensors, targets, weights = _inputs(TL, B, C, h, w, None, seed=5)
        ref = [[t.double().requires_grad_(True) for t in grp] for grp in tensors]
        _, logits_r = O.head_forward(*ref, levels, groups, None)
        onehots, _ = O.predict_onehot_masked([z.detach().float() for z in logits_r], targets)
        loss_r, _ = O.total_loss(logits_r, targets, weights, onehots, levels, parent_of)
        loss_r.backward()
        step = rhseg_b200.FusedHierStep(TL, weights)
        leaves = [[t.clone().cuda().requires_grad_(True) for t in grp] for grp in tensors]
        out = step(*leaves, torch.cat(targets, dim=1).cuda(), None)
        out.loss.backward()
        assert abs(out.loss.item() - loss_r.item()) <= 1e-5 * abs(loss_r.item())
        for grp, rgrp, nm in zip(leaves, ref, ("dfeats", "dhead_w", "dhead_b", "dfilm_w", "dfilm_b")):
            for i, (a, b) in enumerate(zip(grp, rgrp)):
                close(a.grad, b.grad, what="%s %s%d" % (name, nm, i))
    print("CONV_GROUP_OK")


if __name__ == "__main__":
    main()

"""Checks that need the reference tree (/root/reference): run in the build container, skipped on
the GPU box.  They pin the donor backbones and the module drop-in mechanics against the unmodified
reference."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

# /root/reference in the build container, the staged byte copies (oracle/_ref, tools/stage_reference.py) elsewhere
REF = ref_loader.reference_root() or "/root/reference"
pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _ref():
    import ref_shim
    return ref_shim, ref_shim.load_reference()


def test_unet_donor_matches_reference_backbone():
    ref_shim, (rmodels, _, _) = _ref()
    from rhseg_b200.Models import models
    tree = json.load(open(os.path.join(REF, "class_tree_tl_extended.json")))
    torch.manual_seed(0)
    theirs = rmodels.UNet(size=64, n_channels=3, hierarchy=tree, model_type=1).eval()
    ours = models.UNet(size=64, n_channels=3, hierarchy=tree, model_type=1).eval()
    assert set(ours.state_dict()) == set(theirs.state_dict())
    assert all(ours.state_dict()[k].shape == v.shape for k, v in theirs.state_dict().items())
    ours.load_state_dict(theirs.state_dict())
    x = torch.randn(2, 3, 40, 56)
    with torch.no_grad():
        assert torch.equal(ours._run_unet(x), theirs._run_unet(x))
    assert ours.levels == theirs.levels and ours.parent_of == theirs.parent_of
    assert ours.child_groups == theirs.child_groups and ours.children_of == theirs.children_of


def test_hrnet_donor_matches_reference_backbone():
    ref_shim, (rmodels, _, _) = _ref()
    from rhseg_b200.Models import models
    cfg = ref_shim.hrnet_config()
    tree = json.load(open(os.path.join(REF, "class_tree_tl.json")))
    torch.manual_seed(0)
    theirs = rmodels.HighResolutionNet(cfg, hierarchy=tree, model_type=1).eval()
    ours = models.HighResolutionNet(cfg, hierarchy=tree, model_type=1).eval()
    sd = theirs.state_dict()
    assert set(ours.state_dict()) == set(sd)
    assert all(ours.state_dict()[k].shape == v.shape for k, v in sd.items())
    ours.load_state_dict(sd)
    x = torch.randn(1, 3, 64, 96)
    with torch.no_grad():
        a, b = ours._forward_backbone(x), theirs._forward_backbone(x)
    assert a.shape == b.shape == (1, 720, 16, 24)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), float((a - b).abs().max())
    flat_o = models.HighResolutionNet(cfg, hierarchy=tree, model_type=0).eval()
    flat_t = rmodels.HighResolutionNet(cfg, hierarchy=tree, model_type=0).eval()
    flat_o.load_state_dict(flat_t.state_dict())
    with torch.no_grad():
        assert torch.allclose(flat_o(x)[1], flat_t(x)[1], rtol=1e-5, atol=1e-6)


def test_module_dropin_shadows_reference_modules():
    """with the package directory ahead of the reference root, the reference's
    train.py imports OUR Models / Metrics / tree_util."""
    import subprocess
    code = r"""
import sys
sys.path.insert(0, %r); import ref_shim
for n in ("timm","timm.models","timm.models.vision_transformer","segmentation_models_pytorch","torchmetrics","yacs","yacs.config",
          "matplotlib","matplotlib.pyplot","skimage","skimage.io","skimage.transform"):
    ref_shim._stub(n)
sys.modules["yacs.config"].CfgNode = ref_shim.CfgNode
sys.path[:0] = [%r, %r, %r]
import train
pkg = %r
assert train.models.__file__.startswith(pkg), train.models.__file__
assert train.losses.__file__.startswith(pkg), train.losses.__file__
assert train.performance_metrics.__file__.startswith(pkg)
import tree_util; assert tree_util.__file__.startswith(pkg)
import Data.dataset as d; assert d.__file__.startswith(%r)
fns = [[train.losses.CrossEntropyLoss(), train.losses.SoftDiceLoss(num_classes=4)]]
print("ok")
""" % (os.path.join(ROOT, "tests", "golden"), ROOT, os.path.join(ROOT, "restrictive-hierarchical-semantic-segmentation_b200"), REF,
       os.path.join(ROOT, "restrictive-hierarchical-semantic-segmentation_b200"), REF)
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REF, env=env, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_oracle_stitching_matches_reference_predicteval():
    """oracle.stitch_flat_to_levels == predictEval.get_parent_masks + combine_levels on both shipped trees."""
    import importlib
    ref_shim, _ = _ref()
    for n in ("matplotlib.colors", "skimage.measure", "skimage.color", "scipy", "sklearn"):
        ref_shim._stub(n)
    pe = importlib.import_module("predictEval")
    from oracle import hier_oracle as O
    for fname in ("class_tree_tl.json", "class_tree_tl_extended.json"):
        tree = json.load(open(os.path.join(REF, fname)))
        names = [n for n in pe.bfs_order(tree) if not pe.children_map(tree).get(n)]
        idx = {n: i for i, n in enumerate(names)}
        g = torch.Generator().manual_seed(3)
        lab = torch.randint(0, len(names), (2, 9, 11), generator=g)
        flat = torch.nn.functional.one_hot(lab, len(names)).permute(0, 3, 1, 2).float()
        tgt = torch.nn.functional.one_hot(torch.randint(0, len(names), (2, 9, 11), generator=g), len(names)).permute(0, 3, 1, 2).float()
        par, par_t, _ = pe.get_parent_masks([flat], [tgt], tree, idx)
        parent_order = [n for n in pe.bfs_order(tree) if pe.children_map(tree).get(n)]
        want = pe.combine_levels([flat], par, tree, names, parent_order)
        got = O.stitch_flat_to_levels(flat, tree)
        assert len(want) == len(got) and all(torch.equal(a, b) for a, b in zip(want, got))


def test_oracle_process_classes_pinned_to_reference():
    """ProcessClasses (Metrics/performance_metrics.py:27-47) is torch-only: the reference's own class, imported with the
    torchmetrics import stubbed, pins the oracle's restatement (and through it the CUDA confusion kernels) on the three
    kinds of input that reach it: masked one-hots, composed probabilities, ties / all-zero pixels."""
    ref_shim, _ = _ref()
    rpm = ref_loader.load_reference_module("Metrics.performance_metrics")
    from oracle import hier_oracle as O
    theirs = rpm.ProcessClasses()
    g = torch.Generator().manual_seed(9)
    for K in (1, 2, 3, 4, 7):
        B, H, W = 2, 13, 17
        lab = torch.randint(0, K, (B, H, W), generator=g)
        onehot = torch.nn.functional.one_hot(lab, K).permute(0, 3, 1, 2).float()
        ignore = torch.rand(B, 1, H, W, generator=g) < 0.3
        masked = torch.where(ignore, torch.zeros_like(onehot), onehot)            # train.py:230
        probs = torch.rand(B, K, H, W, generator=g)
        probs[:, :, :2] = 0.0                                                     # nothing positive
        probs[:, :, 2:4] = probs[:, :1, 2:4]                                      # exact ties
        tgt = torch.nn.functional.one_hot(torch.randint(0, K, (B, H, W), generator=g), K).permute(0, 3, 1, 2).float()
        tgt = torch.where(torch.rand(B, 1, H, W, generator=g) < 0.4, torch.zeros_like(tgt), tgt)
        for p in (masked, probs):
            for child in (False, True):
                a, b = theirs(p, tgt, child)
                c, d = O.process_classes(p, tgt, child)
                assert torch.equal(a, c.to(a.dtype)) and torch.equal(b, d.to(b.dtype)), (K, child)

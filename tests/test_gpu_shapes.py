"""GPU parity over ragged shapes: odd pixel counts (scalar TMA rows), channel counts that are not
a multiple of the pipeline stage, batch sizes beyond the per-pass limits of the small kernels,
non-integer upsampling factors (tiled adjoint fallback, scalar hi-res path), deep / wide trees.
Reference = the oracle with autograd on CPU (itself pinned to the reference by the fixtures)."""
import pytest
import torch

from helpers import close
from oracle import hier_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

TL = {"background": {}, "upper": {}, "lower": {}, "tooth": {"pulp": {}, "dentin": {}, "enamel": {}, "composite": {}}}
EXT = {"background": {}, "tooth+alveolar": {"alveolar": {"upper": {}, "lower": {}},
                                            "tooth": {"composite": {}, "healthy": {"pulp": {}, "dentin": {}, "enamel": {}}}}}
ADV = {"a": {}, "b": {"b0": {}, "b1": {"b1x": {}}}, "c": {"c0": {}, "c1": {}, "c2": {}}}
WIDE = {"r%d" % i: ({"r%d_%d" % (i, j): {} for j in range(2)} if i % 2 == 0 else {}) for i in range(8)}  # K = [8, 8]

CASES = [
    ("adv_oddN_B5", ADV, 5, 20, 7, 9, None),
    ("ext_up_noninteger", EXT, 3, 40, 6, 5, (18, 22)),
    ("tl_up4", TL, 2, 16, 9, 8, (36, 32)),
    ("wide8_C33", WIDE, 2, 33, 8, 8, None),
    ("ext_fullres_B9", EXT, 9, 24, 6, 10, None),
    ("tl_up_odd_out", TL, 2, 18, 5, 7, (19, 27)),
]


def _inputs(tree_dict, B, C, h, w, out_size, seed):
    levels, parent_of, _, groups = O.hierarchy_tables(tree_dict)
    chans = [len(levels[0])] + [sum(len(k) for _, k in g) for g in groups]
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(B, C, h, w, generator=g) for _ in chans]
    hw = [torch.randn(k, C, 1, 1, generator=g) * 0.3 for k in chans]
    hb = [torch.randn(k, generator=g) * 0.1 for k in chans]
    fw = [torch.randn(2 * C, kp, generator=g) * 0.7 for kp in chans[:-1]]
    fb = [torch.randn(2 * C, generator=g) * 0.3 + 1.0 for _ in chans[:-1]]
    H, W = out_size if out_size else (h, w)
    targets = O.synth_targets(levels, groups, B, H, W, g, blobs=True)
    weights = [[0.4 + 0.3 * ((i * 7 + L) % 5) for i in range(k)] for L, k in enumerate(chans)]
    return levels, parent_of, groups, chans, (feats, hw, hb, fw, fb), targets, weights


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_step_and_dropin_against_oracle(case):
    import rhseg_b200
    from rhseg_b200 import metric_ops
    from rhseg_b200.Metrics import losses
    name, tree_dict, B, C, h, w, out_size = case
    levels, parent_of, groups, chans, tensors, targets, weights = _inputs(tree_dict, B, C, h, w, out_size, seed=len(name))
    n = len(chans)
    # ---- oracle (CPU autograd, in float64: its own fp32 summation error would otherwise be the largest term of the
    # comparison and force tolerances beyond north_star's 1e-5) ----
    ref_leaves = [[t.double().requires_grad_(True) for t in grp] for grp in tensors]
    probs_r, logits_r = O.head_forward(*ref_leaves, levels, groups, out_size)
    onehots_r, evals_r = O.predict_onehot_masked([z.detach().float() for z in logits_r], targets)
    loss_r, per_level_r = O.total_loss(logits_r, targets, weights, onehots_r, levels, parent_of)
    loss_r.backward()
    # ---- fused step ----
    step = rhseg_b200.FusedHierStep(tree_dict, weights)
    leaves = [[t.clone().to(DEV).requires_grad_(True) for t in grp] for grp in tensors]
    target = torch.cat(targets, dim=1).to(DEV)
    out = step(*leaves, target, out_size)
    out.loss.backward()
    assert abs(out.loss.item() - loss_r.item()) <= 1e-5 * abs(loss_r.item()), (out.loss.item(), loss_r.item())
    for L in range(n):
        close(out.logits[L], logits_r[L], what=f"{name} logits{L}")
        close(out.probs[L], probs_r[L], what=f"{name} probs{L}")
        # integers from the device's own logits (argmax may legitimately differ on ~1e-7 near-ties)
        oh_d, ev_d = O.predict_onehot_masked([out.logits[L].cpu()], [targets[L]])
        assert torch.equal(out.confusion[L].cpu(), O.level_confusion(oh_d[0], ev_d[0], chans[L], L != 0)), f"{name} conf{L}"
    for grp, ref_grp, nm in zip(leaves, ref_leaves, ("dfeats", "dhead_w", "dhead_b", "dfilm_w", "dfilm_b")):
        for i, (a, b) in enumerate(zip(grp, ref_grp)):
            close(a.grad, b.grad, what=f"{name} {nm}{i}")
    # ---- drop-in modules on the same inputs ----
    tree = rhseg_b200.ClassTree(tree_dict)
    leaves2 = [[t.clone().to(DEV).requires_grad_(True) for t in grp] for grp in tensors]
    probs, logits = rhseg_b200.hier_head_forward(tree, *leaves2, out_size)
    tg = [t.to(DEV) for t in targets]
    total, onehots = 0.0, []
    for L in range(n):
        onehots.append(metric_ops.predict_onehot(logits[L].detach(), tg[L])[0])
        total = total + losses.CrossEntropyLoss()(logits[L], tg[L], True, weights[L])
        d = losses.SoftDiceLoss()(logits[L], tg[L], True, weights[L])
        total = total + (d if d is not None else 0.0)
    total = total + losses.hierarchical_consistency_loss(onehots, tree.levels, tree.parent_of)
    total.backward()
    assert abs(total.item() - loss_r.item()) <= 1e-5 * abs(loss_r.item())
    for grp, ref_grp, nm in zip(leaves2, ref_leaves, ("dfeats", "dhead_w", "dhead_b", "dfilm_w", "dfilm_b")):
        for i, (a, b) in enumerate(zip(grp, ref_grp)):
            close(a.grad, b.grad, what=f"{name} dropin {nm}{i}")


def test_eval_mode_consistency_and_metrics_on_composed_probs():
    """test()-style use (train.py:324-340): metrics on composed float probabilities with raw ternary
    targets, consistency on real probabilities (~0 by construction)."""
    import rhseg_b200
    from rhseg_b200.Metrics import losses, performance_metrics as pm
    levels, parent_of, groups, chans, tensors, targets, weights = _inputs(EXT, 2, 16, 12, 16, None, seed=5)
    tree = rhseg_b200.ClassTree(EXT)
    with torch.no_grad():
        probs, logits = rhseg_b200.hier_head_forward(tree, *[[t.to(DEV) for t in g] for g in tensors])
        cons = losses.hierarchical_consistency_loss(probs, tree.levels, tree.parent_of)
        ref_cons = O.consistency_loss([p.cpu() for p in probs], levels, parent_of)
    assert abs(cons.item() - ref_cons.item()) <= 1e-7 and cons.item() < 1e-6
    for L in range(len(chans)):
        t = targets[L].to(DEV)
        want = O.level_metrics(probs[L].cpu(), targets[L], chans[L], L != 0)
        for key, mod in (("iou", pm.Jaccardindex()), ("accuracy", pm.Accuracy()), ("precision", pm.Precision())):
            assert torch.equal(mod(probs[L], t, DEV, chans[L], L != 0).cpu(), want[key]), (L, key)


@pytest.mark.parametrize("tree_dict", [TL, EXT, ADV], ids=["tl", "ext", "adv"])
def test_flat_to_hierarchy_stitching(tree_dict):
    """predictEval's flat -> hierarchy stitching as one table-driven kernel, bit-exact."""
    import rhseg_b200
    from rhseg_b200 import metric_ops
    tree = rhseg_b200.ClassTree(tree_dict)
    nl = len(tree.leaf_order())
    g = torch.Generator().manual_seed(nl)
    for shape in ((2, 9, 11), (3, 16, 24)):
        lab = torch.randint(0, nl + 1, shape, generator=g)  # class nl = "no leaf predicted"
        flat = torch.nn.functional.one_hot(lab, nl + 1).permute(0, 3, 1, 2).float()[:, :nl].contiguous()
        want = O.stitch_flat_to_levels(flat, tree_dict)
        got = metric_ops.stitch_flat_to_levels(flat.to(DEV), tree)
        assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in want]
        for a, b in zip(got, want):
            assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("shape", [(2, 3, 4, 9, 11), (1, 1, 8, 16, 24)])
def test_concat_image_logits_utility(shape):
    import rhseg_b200
    B, ci, K, H, W = shape
    g = torch.Generator().manual_seed(1)
    x, z = torch.randn(B, ci, H, W, generator=g), torch.randn(B, K, H, W, generator=g)
    got = rhseg_b200.concat_image_logits(x.to(DEV), z.to(DEV))
    assert torch.equal(got.cpu(), torch.cat([x, z], dim=1))


@pytest.mark.parametrize("shape", [(2, 3, 10, 12, 40, 48), (1, 4, 155, 155, 620, 620), (3, 1, 7, 9, 21, 28), (2, 8, 6, 8, 24, 32),
                                   (1, 2, 5, 5, 5, 8), (1, 2, 6, 8, 50, 16), (5, 4, 9, 7, 33, 20), (1, 3, 2, 3, 7, 12)],
                         ids=lambda s: "B%d_K%d_%dx%d_to_%dx%d" % s)
def test_upsample_adjoint_kernels_agree_with_autograd(shape):
    """The three adjoint kernels (band kernel into a pre-zeroed buffer, row kernel + y-reduction, tiled kernel)
    against autograd through F.interpolate(bilinear, align_corners=True) in fp64."""
    from rhseg_b200 import native
    from rhseg_b200.native import call, ptr
    B, K, Hf, Wf, H, W = shape
    g = torch.Generator().manual_seed(7 + sum(shape))
    dz_hi = torch.randn(B, K, H, W, generator=g)
    x = torch.zeros(B, K, Hf, Wf, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=True).backward(dz_hi.double())
    want = x.grad.float()
    d = dz_hi.to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    tmp = torch.empty(B, K, H, Wf, device=DEV)
    outs = {}
    outs["band"] = torch.zeros(B, K, Hf, Wf, device=DEV)
    call("rhseg_upsample_adjoint", ptr(d), B, K, Hf, Wf, H, W, ptr(outs["band"]), None, native.DZ_PREZEROED, st)
    outs["rows"] = torch.full((B, K, Hf, Wf), float("nan"), device=DEV)
    call("rhseg_upsample_adjoint", ptr(d), B, K, Hf, Wf, H, W, ptr(outs["rows"]), ptr(tmp), 0, st)
    outs["tiled"] = torch.full((B, K, Hf, Wf), float("nan"), device=DEV)
    call("rhseg_upsample_adjoint", ptr(d), B, K, Hf, Wf, H, W, ptr(outs["tiled"]), None, 0, st)
    for name, got in outs.items():
        close(got, want, what="adjoint[%s]" % name)
    # the band kernel is deterministic (at most two CTAs add into a row)
    again = torch.zeros(B, K, Hf, Wf, device=DEV)
    call("rhseg_upsample_adjoint", ptr(d), B, K, Hf, Wf, H, W, ptr(again), None, native.DZ_PREZEROED, st)
    assert torch.equal(again, outs["band"])


@pytest.mark.parametrize("group", [1, 2, 3])
def test_conv_backward_tile_groups(group):
    """The conv backward walks (sample, tile group, channel stage, tile): small groups keep dz in L2 between channel
    stages for large planes x batches (BASELINE configs[4]).  The group size is a per-process setting, so the check runs
    in a worker process: planes of 4-5 tiles with a partial last group, 16-byte- and 4-byte-aligned."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, RHSEG_TUNE_BWD_GROUP=str(group))
    r = subprocess.run([sys.executable, os.path.join(here, "conv_group_worker.py")], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "CONV_GROUP_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]

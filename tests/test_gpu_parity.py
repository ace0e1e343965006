"""GPU parity: the sm_100a kernels (through the C-ABI) against the oracle and the reference-
generated golden fixtures, on the same seeded inputs.

Tolerances (north_star): confusion matrices and argmax masks bit-exact; probabilities, loss
and gradients within 1e-5 relative fp32 (helpers.close adds the scale-relative floor that
SURVEY.md section 7 calls for)."""
import json
import math

import numpy as np
import pytest
import torch

from helpers import HIER_CASES, Fixture, close
from oracle import hier_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rh():
    import rhseg_b200
    return rhseg_b200


def _run_head(fx, requires_grad=True):
    rh = _rh()
    tree = rh.ClassTree(fx.tree)
    mk = lambda ts: [t.to(DEV).requires_grad_(requires_grad) for t in ts]
    feats = mk(fx.per_level("feats"))
    hw, hb = mk(fx.per_level("head_w")), mk(fx.per_level("head_b"))
    fw, fb = mk(fx.per_level("film_w", n=fx.nL - 1)), mk(fx.per_level("film_b", n=fx.nL - 1))
    probs, logits = rh.hier_head_forward(tree, feats, hw, hb, fw, fb, fx.out_size)
    return tree, feats, hw, hb, fw, fb, probs, logits


@pytest.mark.parametrize("name", HIER_CASES)
def test_head_forward_matches_reference(name):
    fx = Fixture(name)
    *_, probs, logits = _run_head(fx, requires_grad=False)
    for L in range(fx.nL):
        close(logits[L], fx.t(f"logits{L}"), what=f"{name} logits{L}")
        close(probs[L], fx.t(f"probs{L}"), what=f"{name} probs{L}")


@pytest.mark.parametrize("name", HIER_CASES)
def test_losses_and_gradients_match_reference(name):
    """Drop-in modules driven the way train.get_loss drives them (train.py:132-149)."""
    from rhseg_b200.Metrics import losses
    fx = Fixture(name)
    tree, feats, hw, hb, fw, fb, probs, logits = _run_head(fx)
    full_target = torch.cat([t.to(DEV) for t in fx.per_level("target")], dim=1)
    targets, s = [], 0
    for L in range(fx.nL):                      # channel slices of one wide tensor, like train.py:185-193
        k = logits[L].shape[1]
        targets.append(full_target[:, s:s + k])
        s += k
    total = 0.0
    for L in range(fx.nL):
        ce = losses.CrossEntropyLoss()(logits[L], targets[L], class_weight=fx.level_weights[L], logits_input=True)
        di = losses.SoftDiceLoss(num_classes=4)(logits[L], targets[L], class_weight=fx.level_weights[L], logits_input=True)
        assert abs(ce.item() - fx.f(f"ce{L}")) <= 1e-5 * max(1.0, abs(fx.f(f"ce{L}"))), (L, ce.item(), fx.f(f"ce{L}"))
        total = total + ce
        if math.isnan(fx.f(f"dice{L}")):
            assert di is None
        else:
            assert abs(di.item() - fx.f(f"dice{L}")) <= 1e-5, (L, di.item(), fx.f(f"dice{L}"))
            total = total + di
    from rhseg_b200 import metric_ops
    onehots = [metric_ops.predict_onehot(logits[L].detach(), targets[L])[0] for L in range(fx.nL)]
    cons = losses.hierarchical_consistency_loss(onehots, tree.levels, tree.parent_of, reduction="mean")
    assert abs(cons.item() - fx.f("consistency_train")) <= 1e-6
    total = total + cons
    assert abs(total.item() - fx.f("total_loss")) <= 1e-5 * abs(fx.f("total_loss"))
    with torch.no_grad():
        cons_eval = losses.hierarchical_consistency_loss([p.detach() for p in probs], tree.levels, tree.parent_of)
    assert abs(cons_eval.item() - fx.f("consistency_eval")) <= 1e-6
    total.backward()
    for L in range(fx.nL):
        close(feats[L].grad, fx.t(f"dfeats{L}"), what=f"{name} dfeats{L}")
        close(hw[L].grad, fx.t(f"dhead_w{L}"), what=f"{name} dhead_w{L}")
        close(hb[L].grad, fx.t(f"dhead_b{L}"), what=f"{name} dhead_b{L}")
    for i in range(fx.nL - 1):
        close(fw[i].grad, fx.t(f"dfilm_w{i}"), what=f"{name} dfilm_w{i}")
        close(fb[i].grad, fx.t(f"dfilm_b{i}"), what=f"{name} dfilm_b{i}")


@pytest.mark.parametrize("name", HIER_CASES)
def test_train_path_prediction_bit_exact(name):
    from rhseg_b200 import metric_ops
    fx = Fixture(name)
    for L in range(fx.nL):
        z = fx.t(f"logits{L}", DEV)
        t = fx.t(f"target{L}", DEV)
        onehot, eval_t, idx = metric_ops.predict_onehot(z, t, want_index=True)
        assert torch.equal(onehot.cpu(), fx.t(f"onehot{L}")), f"{name} onehot{L}"
        assert torch.equal(eval_t.cpu(), torch.where(fx.t(f"target{L}") == -1, 0, fx.t(f"target{L}")))
        ref_idx = torch.argmax(torch.softmax(z, dim=1), dim=1)          # same device, ATen
        assert torch.equal(idx.long(), ref_idx)


def _metric_inputs(K, B, H, W, seed, child, mode):
    g = torch.Generator().manual_seed(seed)
    lab_p = torch.randint(0, K, (B, H, W), generator=g)
    lab_t = torch.randint(0, K, (B, H, W), generator=g)
    tgt = torch.nn.functional.one_hot(lab_t, K).permute(0, 3, 1, 2).float()
    if mode == "onehot":
        probs = torch.nn.functional.one_hot(lab_p, K).permute(0, 3, 1, 2).float()
        if child:
            hole = torch.rand(B, 1, H, W, generator=g) < 0.5
            probs = torch.where(hole, torch.zeros_like(probs), probs)
            tgt = torch.where(hole, torch.zeros_like(tgt), tgt)
    else:  # composed float probabilities + raw ternary targets (test() path, train.py:324-337)
        probs = torch.rand(B, K, H, W, generator=g)
        if child:
            hole = torch.rand(B, 1, H, W, generator=g) < 0.5
            tgt = torch.where(hole, torch.full_like(tgt, -1.0), tgt)
    return probs, tgt


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 7, 8])
@pytest.mark.parametrize("child", [False, True])
@pytest.mark.parametrize("mode", ["onehot", "float"])
@pytest.mark.parametrize("shape", [(3, 17, 19), (2, 24, 32)])
def test_confusion_matrix_bit_exact(K, child, mode, shape):
    from rhseg_b200 import metric_ops
    B, H, W = shape
    probs, tgt = _metric_inputs(K, B, H, W, 100 + K, child, mode)
    want = O.level_confusion(probs, tgt, K, child)
    got = metric_ops.confusion_matrix(probs.to(DEV), tgt.to(DEV), child)
    assert torch.equal(got.cpu(), want)
    r_want = O.ratios_from_confusion(want)
    r_got = metric_ops.ratios(got).cpu()
    for key, row in metric_ops.ROW.items():
        assert torch.equal(r_got[row], r_want[key]), key


@pytest.mark.parametrize("name", ["unet_tl", "unet_ext", "hrnet_ext"])
def test_metric_wrappers_and_fused_logit_metrics(name):
    """The five drop-in wrappers vs the oracle's restated torchmetrics slice; and the fused
    logits->confusion kernel vs predict_onehot + confusion."""
    from rhseg_b200 import metric_ops
    from rhseg_b200.Metrics import performance_metrics as pm
    fx = Fixture(name)
    wrappers = {"iou": pm.Jaccardindex(), "accuracy": pm.Accuracy(), "dice": pm.DiceScore(),
                "precision": pm.Precision(), "recall": pm.Recall()}
    for L in range(fx.nL):
        z, t = fx.t(f"logits{L}", DEV), fx.t(f"target{L}", DEV)
        onehot, eval_t = metric_ops.predict_onehot(z, t)
        want = O.level_metrics(onehot.cpu(), eval_t.cpu(), z.shape[1], L != 0)
        for key, mod in wrappers.items():
            got = mod(onehot, eval_t, DEV, z.shape[1], L != 0)
            assert got.shape == (z.shape[1],) and torch.equal(got.cpu(), want[key]), (name, L, key)
        fused = metric_ops.confusion_from_logits(z, t, L != 0)
        assert torch.equal(fused, metric_ops.confusion_matrix(onehot, eval_t, L != 0))


@pytest.mark.parametrize("K", [1, 2, 3, 5, 6, 7, 8])
def test_loss_kernels_other_channel_counts(K):
    """CE/Dice value + gradient for every compiled K, odd pixel counts, strided targets, raw
    (logits_input=False) mode; oracle = reference restatement with autograd."""
    from rhseg_b200.Metrics import losses
    g = torch.Generator().manual_seed(K)
    B, H, W = 3, 9, 11
    w = [0.3 + 0.4 * i for i in range(K)]
    for logits_input in (True, False):
        z = torch.randn(B, K, H, W, generator=g)
        if not logits_input:
            z = torch.softmax(z, 1)
        lab = torch.randint(0, K, (B, H, W), generator=g)
        t = torch.nn.functional.one_hot(lab, K).permute(0, 3, 1, 2).float()
        t = torch.where(torch.rand(B, 1, H, W, generator=g) < 0.3, torch.full_like(t, -1.0), t)
        wide = torch.cat([torch.zeros(B, 2, H, W), t, torch.zeros(B, 1, H, W)], 1)
        zc = z.clone().requires_grad_(True)
        ref = O.ce_loss(zc, t, w, logits_input) + O.dice_loss(zc, t, w, logits_input)
        ref.backward()
        zg = z.to(DEV).requires_grad_(True)
        tg = wide.to(DEV)[:, 2:2 + K]
        got = losses.CrossEntropyLoss()(zg, tg, logits_input, w) + losses.SoftDiceLoss()(zg, tg, logits_input, w)
        got.backward()
        assert abs(got.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
        close(zg.grad, zc.grad, what=f"K={K} logits_input={logits_input}")


def test_dice_none_and_ce_nan_conventions():
    from rhseg_b200.Metrics import losses
    z = torch.randn(2, 4, 8, 8, device=DEV, requires_grad=True)
    t = torch.full((2, 4, 8, 8), -1.0, device=DEV)
    w = [1.0, 2.0, 3.0, 4.0]
    ce = losses.CrossEntropyLoss()(z, t, True, w)
    assert ce.item() == 1.0                      # every sample NaN -> nan_to_num(nan=1.0)
    assert losses.SoftDiceLoss()(z, t, True, w) is None
    ce.backward()
    assert float(z.grad.abs().max()) == 0.0
    with pytest.raises(ValueError):
        losses.CrossEntropyLoss()(z, t, True, None)


def test_head_other_channel_counts_and_user_prob_grads():
    """Random trees covering K in 1..8, gradients placed directly on the probabilities (not just
    the reference's loss) — oracle autograd on CPU."""
    rh = _rh()
    tree_dict = {"r0": {"a": {}, "b": {}, "c": {"c0": {}, "c1": {}}}, "r1": {"d": {"d0": {}}},
                 "r2": {}, "r3": {"e": {}, "f": {}, "g": {}, "h": {}}, "r4": {}}
    tree = rh.ClassTree(tree_dict)
    levels, parent_of, children_of, groups = O.hierarchy_tables(tree_dict)
    assert tree.head_channels == [5, 8, 3]
    g = torch.Generator().manual_seed(3)
    B, C, H, W = 2, 24, 10, 14
    feats = [torch.randn(B, C, H, W, generator=g) for _ in range(3)]
    hw = [torch.randn(k, C, 1, 1, generator=g) * 0.3 for k in tree.head_channels]
    hb = [torch.randn(k, generator=g) * 0.1 for k in tree.head_channels]
    fw = [torch.randn(2 * C, kp, generator=g) for kp in tree.head_channels[:-1]]
    fb = [torch.randn(2 * C, generator=g) * 0.3 + 1 for _ in range(2)]
    gp = [torch.randn(B, k, H, W, generator=g) for k in tree.head_channels]
    gz = [torch.randn(B, k, H, W, generator=g) for k in tree.head_channels]

    def run(dev, fn):
        # the CPU oracle runs in float64 (the comparison is then at north_star's 1e-5, not at the oracle's fp32 error)
        dt = torch.float64 if dev == "cpu" else torch.float32
        leaves = [[t.clone().to(dev, dt).requires_grad_(True) for t in grp] for grp in (feats, hw, hb, fw, fb)]
        probs, logits = fn(*leaves)
        loss = sum((p * a.to(dev, dt)).sum() for p, a in zip(probs, gp)) + sum((z * a.to(dev, dt)).sum() for z, a in zip(logits, gz))
        loss.backward()
        return probs, logits, leaves

    p_ref, z_ref, l_ref = run("cpu", lambda f, a, b, c, d: O.head_forward(f, a, b, c, d, levels, groups))
    p_got, z_got, l_got = run(DEV, lambda f, a, b, c, d: rh.hier_head_forward(tree, f, a, b, c, d))
    for L in range(3):
        close(z_got[L], z_ref[L], what=f"logits{L}")
        close(p_got[L], p_ref[L], what=f"probs{L}")
    for grp_got, grp_ref, nm in zip(l_got, l_ref, ("dfeats", "dhead_w", "dhead_b", "dfilm_w", "dfilm_b")):
        for i, (a, b) in enumerate(zip(grp_got, grp_ref)):
            close(a.grad, b.grad, what=f"{nm}{i}")


def test_children_sum_to_parent_property_full_size():
    """Size-independent properties at BASELINE.json's full size (620x620, B=4, UNet tl):
    sum_children P_c == P_parent, confusion row sums == pixel counts, loss invariant under a
    batch permutation."""
    rh = _rh()
    from rhseg_b200 import metric_ops
    from rhseg_b200.Metrics import losses
    tl = Fixture("unet_tl").tree
    tree = rh.ClassTree(tl)
    levels, parent_of, _, groups = O.hierarchy_tables(tl)
    torch.manual_seed(0)
    B, C, H, W = 4, 64, 620, 620
    feats = [torch.randn(B, C, H, W, device=DEV) for _ in range(2)]
    hw = [torch.randn(4, C, 1, 1, device=DEV) * 0.2 for _ in range(2)]
    hb = [torch.zeros(4, device=DEV) for _ in range(2)]
    fw = [torch.randn(2 * C, 4, device=DEV)]
    fb = [torch.ones(2 * C, device=DEV)]
    probs, logits = rh.hier_head_forward(tree, feats, hw, hb, fw, fb)
    tooth = levels[0].index("tooth")
    diff = (probs[1].sum(1) - probs[0][:, tooth]).abs().max().item()
    assert diff <= 4 * 1.2e-7 * 4, diff
    targets = [t.to(DEV) for t in O.synth_targets(levels, groups, B, H, W, torch.Generator().manual_seed(1))]
    conf0 = metric_ops.confusion_from_logits(logits[0], targets[0], False)
    assert int(conf0.sum()) == B * H * W
    assert torch.equal(conf0.sum(1).cpu(), targets[0].sum(dim=(0, 2, 3)).long().cpu())
    conf1 = metric_ops.confusion_from_logits(logits[1], targets[1], True)
    assert int(conf1.sum()) == int((targets[0][:, tooth] == 1).sum())
    w = Fixture("unet_tl").level_weights
    perm = torch.tensor([2, 0, 3, 1], device=DEV)
    for L in range(2):
        a = losses.CrossEntropyLoss()(logits[L], targets[L], True, w[L]) + losses.SoftDiceLoss()(logits[L], targets[L], True, w[L])
        zp, tp = logits[L][perm].contiguous(), targets[L][perm].contiguous()
        b = losses.CrossEntropyLoss()(zp, tp, True, w[L]) + losses.SoftDiceLoss()(zp, tp, True, w[L])
        assert abs(a.item() - b.item()) <= 1e-6 * abs(a.item())


def test_dropin_unet_module_matches_oracle():
    """models.UNet end to end (stock donor + fused head) against the oracle head fed with the
    same donor features; state-dict keys are the reference's."""
    from rhseg_b200.Models import models
    fx = Fixture("unet_ext")
    torch.manual_seed(4)
    m = models.UNet(size=32, n_channels=3, hierarchy=fx.tree, model_type=1).to(DEV)
    m.eval()  # frozen BN statistics -> the donor passes are identical
    x = torch.randn(2, 3, 32, 48, device=DEV)
    with torch.no_grad():
        probs, logits = m(x, type=1, hierarchy=fx.tree)
        f = m._run_unet(x).cpu().double()   # fp64 oracle head on the device's own donor features
    levels, parent_of, _, groups = O.hierarchy_tables(fx.tree)
    sd = {k: v.cpu().double() for k, v in m.state_dict().items()}
    n = len(levels)
    p_ref, z_ref = O.head_forward([f] * n, [sd[f"heads.{L}.conv.weight"] for L in range(n)],
                                  [sd[f"heads.{L}.conv.bias"] for L in range(n)],
                                  [sd[f"films.{i}.mlp.1.weight"] for i in range(n - 1)],
                                  [sd[f"films.{i}.mlp.1.bias"] for i in range(n - 1)], levels, groups)
    for L in range(n):
        close(logits[L], z_ref[L], what=f"logits{L}")
        close(probs[L], p_ref[L], what=f"probs{L}")
    assert m.levels == levels and m.parent_of == parent_of


def test_flat_seven_class_losses_and_metrics_match_reference():
    """BASELINE.json configs[3]: flat 7-class weighted Dice+CE (reference-generated fixture) and the
    confusion-matrix metrics of its argmax prediction (oracle's restated torchmetrics slice)."""
    import os
    import numpy as np
    from helpers import GOLDEN
    from rhseg_b200 import metric_ops
    from rhseg_b200.Metrics import losses, performance_metrics as pm
    z = np.load(os.path.join(GOLDEN, "flat7.npz"))
    logits = torch.from_numpy(z["logits"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(z["target"]).float().to(DEV)
    w = [float(v) for v in z["weights"]]
    ce = losses.CrossEntropyLoss()(logits, t, class_weight=w, logits_input=True)
    di = losses.SoftDiceLoss(num_classes=7)(logits, t, class_weight=w, logits_input=True)
    assert abs(ce.item() - float(z["ce"])) <= 1e-5 and abs(di.item() - float(z["dice"])) <= 1e-5
    (ce + di).backward()
    close(logits.grad, z["dlogits"], what="flat dlogits")
    # train.py:206-216 for model_type 0: one-hot of argmax(softmax), no -1 in flat targets
    onehot, eval_t = metric_ops.predict_onehot(logits.detach(), t)
    ref_oh, ref_et = O.predict_onehot_masked([logits.detach().cpu()], [t.cpu()])
    assert torch.equal(onehot.cpu(), ref_oh[0])
    want = O.level_metrics(ref_oh[0], ref_et[0], 7, False)
    for key, mod in (("iou", pm.Jaccardindex()), ("dice", pm.DiceScore()), ("recall", pm.Recall())):
        assert torch.equal(mod(onehot, eval_t, DEV, 7, False).cpu(), want[key])


def test_eval_mode_backbone_reuse_is_output_identical():
    """in eval() the per-level donor passes are identical, so one pass feeds all levels;
    outputs equal the pass-per-level replay bit for bit, and train() mode still runs one pass per level."""
    from rhseg_b200.Models import models
    fx = Fixture("unet_ext")
    torch.manual_seed(9)
    m = models.UNet(size=32, n_channels=3, hierarchy=fx.tree, model_type=1).to(DEV)
    x = torch.randn(2, 3, 32, 48, device=DEV)
    calls = {"n": 0}
    orig = m._run_unet

    def counted(inp):
        calls["n"] += 1
        return orig(inp)

    m._run_unet = counted
    m.eval()
    with torch.no_grad():
        p1, z1 = m(x, type=1, hierarchy=fx.tree)
        assert calls["n"] == 1
        m.reuse_backbone_in_eval = False
        p2, z2 = m(x, type=1, hierarchy=fx.tree)
        assert calls["n"] == 1 + len(m.levels)
    for a, b in zip(p1 + z1, p2 + z2):
        assert torch.equal(a, b)
    m.reuse_backbone_in_eval = True
    m.train()
    calls["n"] = 0
    p3, z3 = m(x, type=1, hierarchy=fx.tree)
    assert calls["n"] == len(m.levels)
    sum(z.sum() for z in z3).backward()
    assert m.heads[0].conv.weight.grad is not None and m.films[0].mlp[1].weight.grad is not None
    assert m.inc0.conv.conv[0].weight.grad is not None


def test_one_backward_per_loss_term_on_the_same_tensors():
    """CE back-propagated on its own, then Dice requested for the same (logits, targets) objects.  The memoised
    node has run backward by then (its graph is freed): the second module call must get a fresh node, as the reference's
    two independent modules do, and the accumulated gradient equals d(CE + Dice)."""
    import os
    import numpy as np
    from helpers import GOLDEN
    from rhseg_b200.Metrics import losses
    z = np.load(os.path.join(GOLDEN, "flat7.npz"))
    logits = torch.from_numpy(z["logits"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(z["target"]).float().to(DEV)
    w = [float(v) for v in z["weights"]]
    losses.CrossEntropyLoss()(logits, t, class_weight=w, logits_input=True).backward()
    di = losses.SoftDiceLoss(num_classes=7)(logits, t, class_weight=w, logits_input=True)
    di.backward()
    close(logits.grad, z["dlogits"], what="dlogits after two separate backward passes")
    # the usual sequence (both terms, one backward) still shares one node
    import rhseg_b200
    rhseg_b200.clear_memo()
    logits.grad = None
    ce = losses.CrossEntropyLoss()(logits, t, class_weight=w, logits_input=True)
    di = losses.SoftDiceLoss(num_classes=7)(logits, t, class_weight=w, logits_input=True)
    assert ce.grad_fn is di.grad_fn
    (ce + di).backward()
    close(logits.grad, z["dlogits"], what="dlogits, shared node")


def test_fused_flat_step_matches_the_reference_fixture():
    """rhseg_b200.FusedFlatStep (flat model, BASELINE.json configs[3]) on the reference-generated fixture: CE + Dice and
    d(CE + Dice)/d logits against the reference's recorded values, confusion matrix and ratios against the oracle."""
    import os
    import numpy as np
    from helpers import GOLDEN
    import rhseg_b200
    z = np.load(os.path.join(GOLDEN, "flat7.npz"))
    logits = torch.from_numpy(z["logits"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(z["target"]).float().to(DEV)
    w = [float(v) for v in z["weights"]]
    step = rhseg_b200.FusedFlatStep(7, w)
    out = step(logits, t)
    ref_total = float(z["ce"]) + float(z["dice"])
    assert abs(out.loss.item() - ref_total) <= 1e-5 * abs(ref_total)
    assert abs(out.level_ce[0].item() - float(z["ce"])) <= 1e-5 and abs(out.level_dice[0].item() - float(z["dice"])) <= 1e-5
    out.loss.backward()
    close(logits.grad, z["dlogits"], what="flat dlogits")
    oh, et = O.predict_onehot_masked([logits.detach().cpu()], [t.cpu()])
    want = O.level_confusion(oh[0], et[0], 7, False)
    assert torch.equal(out.confusion[0].cpu(), want)
    r = O.ratios_from_confusion(want)
    for row, key in enumerate(("dice", "iou", "accuracy", "precision", "recall")):
        assert torch.equal(out.ratios[0][row].cpu(), r[key]), key
    # int8 targets and a channels-last target layout (what a one-hot permute produces) give the same result
    out2 = step(logits.detach().requires_grad_(True), t.to(torch.int8))
    out3 = step(logits.detach().requires_grad_(True), t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    assert torch.equal(out2.scalars, out.scalars) and torch.equal(out3.scalars, out.scalars)

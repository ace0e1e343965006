"""GPU parity of the fused training step (rhseg_b200.FusedHierStep): same inputs as the
reference-generated fixtures, compared with the reference's recorded outputs (loss, gradients)
and with the oracle (metrics, consistency)."""
import math

import pytest
import torch

from helpers import HIER_CASES, Fixture, close
from oracle import hier_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _step(fx, requires_grad=True):
    import rhseg_b200
    step = rhseg_b200.FusedHierStep(fx.tree, fx.level_weights)
    mk = lambda ts: [t.to(DEV).requires_grad_(requires_grad) for t in ts]
    feats = mk(fx.per_level("feats"))
    hw, hb = mk(fx.per_level("head_w")), mk(fx.per_level("head_b"))
    fw, fb = mk(fx.per_level("film_w", n=fx.nL - 1)), mk(fx.per_level("film_b", n=fx.nL - 1))
    # the wide target tensor sits inside a larger buffer so that level slices are strided views
    tcat = torch.cat(fx.per_level("target"), dim=1)
    wide = torch.zeros(tcat.shape[0], tcat.shape[1] + 3, *tcat.shape[2:])
    wide[:, 1:1 + tcat.shape[1]] = tcat
    target = wide.to(DEV)[:, 1:1 + tcat.shape[1]]
    out = step(feats, hw, hb, fw, fb, target, fx.out_size)
    return step, out, (feats, hw, hb, fw, fb)


@pytest.mark.parametrize("name", HIER_CASES)
def test_fused_step_matches_reference(name):
    fx = Fixture(name)
    step, out, (feats, hw, hb, fw, fb) = _step(fx)
    ref_total = fx.f("total_loss")
    assert abs(out.loss.item() - ref_total) <= 1e-5 * abs(ref_total), (out.loss.item(), ref_total)
    assert abs(out.consistency.item() - fx.f("consistency_train")) <= 1e-6
    levels, parent_of, _, groups = O.hierarchy_tables(fx.tree)
    for L in range(fx.nL):
        assert abs(out.level_ce[L].item() - fx.f(f"ce{L}")) <= 1e-5 * max(1.0, abs(fx.f(f"ce{L}")))
        if math.isnan(fx.f(f"dice{L}")):
            assert out.level_dice[L].item() == 0.0 and out.scalars[4 + 4 * L].item() == 0.0
        else:
            assert abs(out.level_dice[L].item() - fx.f(f"dice{L}")) <= 1e-5
        close(out.logits[L], fx.t(f"logits{L}"), what=f"{name} logits{L}")
        close(out.probs[L], fx.t(f"probs{L}"), what=f"{name} probs{L}")
        # metrics: the reference's glue (masked one-hot vs eval target) through the oracle's restated torchmetrics
        onehot = fx.t(f"onehot{L}")
        eval_t = torch.where(fx.t(f"target{L}") == -1, 0, fx.t(f"target{L}"))
        want = O.level_confusion(onehot, eval_t, onehot.shape[1], L != 0)
        assert torch.equal(out.confusion[L].cpu(), want), f"{name} confusion{L}"
        r = O.ratios_from_confusion(want)
        for row, key in enumerate(("dice", "iou", "accuracy", "precision", "recall")):
            assert torch.equal(out.ratios[L][row].cpu(), r[key]), (name, L, key)
    out.loss.backward()
    for L in range(fx.nL):
        close(feats[L].grad, fx.t(f"dfeats{L}"), what=f"{name} dfeats{L}")
        close(hw[L].grad, fx.t(f"dhead_w{L}"), what=f"{name} dhead_w{L}")
        close(hb[L].grad, fx.t(f"dhead_b{L}"), what=f"{name} dhead_b{L}")
    for i in range(fx.nL - 1):
        close(fw[i].grad, fx.t(f"dfilm_w{i}"), what=f"{name} dfilm_w{i}")
        close(fb[i].grad, fx.t(f"dfilm_b{i}"), what=f"{name} dfilm_b{i}")


@pytest.mark.parametrize("name", ["unet_ext", "hrnet_ext", "hrnet_tl"])
def test_fused_step_equals_dropin_modules(name):
    """The fused node and the drop-in module sequence are two routes to the same numbers."""
    import rhseg_b200
    from rhseg_b200 import metric_ops
    from rhseg_b200.Metrics import losses
    fx = Fixture(name)
    _, out, leaves_a = _step(fx)
    out.loss.backward()
    tree = rhseg_b200.ClassTree(fx.tree)
    mk = lambda ts: [t.to(DEV).requires_grad_(True) for t in ts]
    feats = mk(fx.per_level("feats"))
    hw, hb = mk(fx.per_level("head_w")), mk(fx.per_level("head_b"))
    fw, fb = mk(fx.per_level("film_w", n=fx.nL - 1)), mk(fx.per_level("film_b", n=fx.nL - 1))
    probs, logits = rhseg_b200.hier_head_forward(tree, feats, hw, hb, fw, fb, fx.out_size)
    targets = [t.to(DEV) for t in fx.per_level("target")]
    total, onehots = 0.0, []
    for L in range(fx.nL):
        onehot, eval_t = metric_ops.predict_onehot(logits[L].detach(), targets[L])
        onehots.append(onehot)
        assert torch.equal(metric_ops.confusion_matrix(onehot, eval_t, L != 0), out.confusion[L])
        total = total + losses.CrossEntropyLoss()(logits[L], targets[L], True, fx.level_weights[L])
        d = losses.SoftDiceLoss()(logits[L], targets[L], True, fx.level_weights[L])
        total = total + (d if d is not None else 0.0)
    total = total + losses.hierarchical_consistency_loss(onehots, tree.levels, tree.parent_of)
    total.backward()
    assert abs(total.item() - out.loss.item()) <= 2e-6 * abs(total.item())
    for a, b in zip(leaves_a, (feats, hw, hb, fw, fb)):
        for x, y in zip(a, b):
            close(x.grad, y.grad, rtol=2e-6, what=name)


def test_fused_step_summary_matches_pack_layout():
    """StepOutput.summary (written by rhseg_step_finalize) == dist.pack_step_summary of the same step."""
    from rhseg_b200 import dist as rdist
    fx = Fixture("unet_tl_odd_notooth")
    _, out, _ = _step(fx, requires_grad=False)
    want = rdist.pack_step_summary(out.scalars, fx.B, out.confusion)
    assert out.summary.shape == want.shape
    close(out.summary, want, rtol=1e-6, what="summary")
    glob, extras = rdist.all_reduce_summary(out.summary, fx.nL, [tuple(c.shape) for c in out.confusion], extra=[out.scalars])
    assert abs(float(glob["total"]) - out.loss.item()) < 1e-5
    assert torch.equal(glob["confusion"][1], out.confusion[1])
    close(extras[0], out.scalars, rtol=1e-6, what="extra")


def test_exchange_buffer_pack_and_unpack():
    """Data-parallel exchange buffer: StepOutput.exchange = [summary | tail]; rhseg_pack_f64 writes the parameter
    gradients behind the summary in one launch (== torch.cat of the fp64 casts), rhseg_unpack_f32 scales them back."""
    import rhseg_b200
    from rhseg_b200 import dist as rdist
    fx = Fixture("hrnet_tl")
    step = rhseg_b200.FusedHierStep(fx.tree, fx.level_weights)
    mk = lambda ts: [t.to(DEV).requires_grad_(True) for t in ts]
    feats = mk(fx.per_level("feats"))
    params = [mk(fx.per_level("head_w")), mk(fx.per_level("head_b")), mk(fx.per_level("film_w", n=fx.nL - 1)),
              mk(fx.per_level("film_b", n=fx.nL - 1))]
    flat = [p for grp in params for p in grp]
    step.exchange_tail = sum(p.numel() for p in flat)
    target = torch.cat(fx.per_level("target"), dim=1).to(DEV)
    out = step(feats, *params, target, fx.out_size)
    grads = torch.autograd.grad(out.loss, flat)
    assert out.exchange.data_ptr() == out.summary.data_ptr()
    assert out.exchange.numel() == out.summary.numel() + step.exchange_tail
    want = torch.cat([out.summary] + [g.reshape(-1).double() for g in grads])
    buf = rdist.pack_exchange(out.summary, grads, out=out.exchange)
    assert buf.data_ptr() == out.exchange.data_ptr()          # packed in place, the summary was not copied
    assert torch.equal(buf, want)
    assert torch.equal(rdist.pack_exchange(out.summary, grads), want)   # without the preallocated tail
    # a non-contiguous / empty part and a lone summary
    odd = [grads[0].transpose(0, 1), torch.empty(0, device=DEV), grads[1]]
    assert torch.equal(rdist.pack_exchange(out.summary, odd),
                       torch.cat([out.summary] + [g.reshape(-1).double() for g in odd]))
    assert torch.equal(rdist.pack_exchange(out.summary), out.summary)
    back = [torch.full_like(g, float("nan")) for g in grads]
    rdist.unpack_exchange(buf, out.summary.numel(), back, scale=0.25)
    for b, g in zip(back, grads):
        assert torch.equal(b, (g.double() * 0.25).float())
    # results with a tail are the results without one
    step.exchange_tail = 0
    out0 = step(feats, *params, target, fx.out_size)
    assert torch.equal(out0.summary, out.summary) and out0.exchange.numel() == out0.summary.numel()


def test_fused_step_replays_from_a_cuda_graph():
    """The whole step (forward, evaluation, loss, backward; side-stream fork/join included) is capturable: a replay
    with new inputs in the same buffers gives what an eager step on those inputs gives."""
    import rhseg_b200
    fx = Fixture("hrnet_tl")
    step = rhseg_b200.FusedHierStep(fx.tree, fx.level_weights)
    mk = lambda ts: [t.to(DEV).requires_grad_(True) for t in ts]
    feats = mk(fx.per_level("feats"))
    params = [mk(fx.per_level("head_w")), mk(fx.per_level("head_b")), mk(fx.per_level("film_w", n=fx.nL - 1)),
              mk(fx.per_level("film_b", n=fx.nL - 1))]
    leaves = feats + [p for grp in params for p in grp]
    target = torch.cat(fx.per_level("target"), dim=1).to(DEV)
    one = torch.ones((), device=DEV)

    def run():
        out = step(feats, *params, target, fx.out_size)
        grads = torch.autograd.grad(out.loss, leaves, grad_outputs=one)
        return out, grads

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_out, g_grads = run()
    gen = torch.Generator().manual_seed(5)
    for it in range(2):
        with torch.no_grad():
            for f in feats:
                f.copy_(torch.randn(f.shape, generator=gen))
            params[0][0].mul_(0.9)
        graph.replay()
        torch.cuda.synchronize()
        got = (g_out.loss.clone(), [c.clone() for c in g_out.confusion], [g.clone() for g in g_grads])
        e_out, e_grads = run()
        torch.cuda.synchronize()
        close(got[0], e_out.loss, what="graph loss %d" % it)
        for a, b in zip(got[1], e_out.confusion):
            assert torch.equal(a, b)
        for i, (a, b) in enumerate(zip(got[2], e_grads)):
            close(a, b, rtol=1e-6, what="graph grad %d/%d" % (it, i))  # same kernels: only the order of fp64 atomics differs


@pytest.mark.parametrize("name", ["unet_tl_odd_notooth", "hrnet_ext"])
def test_int8_targets_give_identical_results(name):
    """The ternary targets may arrive as int8 (a quarter of the bytes over PCIe): rhseg_targets_i8_to_f32 widens them on
    the device and the step's results are bit-identical to the fp32-target step."""
    import rhseg_b200
    fx = Fixture(name)
    step = rhseg_b200.FusedHierStep(fx.tree, fx.level_weights)
    mk = lambda ts: [t.to(DEV).requires_grad_(True) for t in ts]
    target = torch.cat(fx.per_level("target"), dim=1).to(DEV)
    outs = []
    for tg in (target, target.to(torch.int8)):
        feats = mk(fx.per_level("feats"))
        hw, hb = mk(fx.per_level("head_w")), mk(fx.per_level("head_b"))
        fw, fb = mk(fx.per_level("film_w", n=fx.nL - 1)), mk(fx.per_level("film_b", n=fx.nL - 1))
        out = step(feats, hw, hb, fw, fb, tg, fx.out_size)
        out.loss.backward()
        outs.append((out, feats, hw))
    a, b = outs
    assert torch.equal(a[0].scalars, b[0].scalars)
    for L in range(fx.nL):
        assert torch.equal(a[0].confusion[L], b[0].confusion[L])
        assert torch.equal(a[1][L].grad, b[1][L].grad)
    # odd element counts and the 16-element main loop
    from rhseg_b200 import native
    for n in (1, 15, 16, 17, 4099):
        src = torch.randint(-1, 2, (n + 16,), dtype=torch.int8, device=DEV)[:n]
        dst = torch.empty(n, dtype=torch.float32, device=DEV)
        native.call("rhseg_targets_i8_to_f32", src.data_ptr(), n, dst.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert torch.equal(dst, src.float())

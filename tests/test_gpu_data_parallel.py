"""Batch-sharded (data-parallel) exactness and the level-pretrain curriculum of the fused step, on ONE GPU.

SURVEY.md 8(e): the single-process reference averages CE over all samples of the global batch and Dice over the GLOBAL
number of dice-valid samples (Metrics/losses.py:64-66, :117-119).  The ranks of a sharded job are played one after the
other here (the cross-rank sum of the step summary is formed on the host side of the test); the multi-GPU form of the
same check runs inside bench.py at N > 1 (`dp_check`) and in tests/test_gpu_peer_exchange.py.
"""
import ctypes

import pytest
import torch

from helpers import Fixture, close
from oracle import hier_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _leaves(fx, sl=slice(None)):
    mk = lambda ts: [t.to(DEV).requires_grad_(True) for t in ts]
    feats = [t[sl].contiguous().to(DEV).requires_grad_(True) for t in fx.per_level("feats")]
    hw, hb = mk(fx.per_level("head_w")), mk(fx.per_level("head_b"))
    fw, fb = mk(fx.per_level("film_w", n=fx.nL - 1)), mk(fx.per_level("film_b", n=fx.nL - 1))
    target = torch.cat(fx.per_level("target"), dim=1)[sl].contiguous().to(DEV)
    return feats, hw, hb, fw, fb, target


@pytest.mark.parametrize("name,shards", [("unet_tl_odd_notooth", [slice(0, 1), slice(1, 3)]),   # unequal shards; sample 1 has no tooth
                                         ("unet_tl_odd_notooth", [slice(0, 2), slice(2, 3)]),
                                         ("hrnet_ext", [slice(0, 1), slice(1, 2)]),             # rank 0 holds no dice-valid sample of the deep levels
                                         ("unet_ext", [slice(0, 1), slice(1, 2)])])
def test_sharded_gradients_equal_the_global_batch(name, shards):
    import rhseg_b200
    from rhseg_b200 import dist as rdist
    fx = Fixture(name)
    world = len(shards)
    step = rhseg_b200.FusedHierStep(fx.tree, fx.level_weights)
    # the single-process reference: the whole batch in one step
    feats, hw, hb, fw, fb, target = _leaves(fx)
    whole = step(feats, hw, hb, fw, fb, target, fx.out_size)
    whole.loss.backward()
    want_params = [p.grad.clone() for grp in (hw, hb, fw, fb) for p in grp]
    want_dfeats = [f.grad.clone() for f in feats]
    # pass 1 per rank: the summaries that the forward-side exchange would sum
    summaries = []
    for sl in shards:
        f, a, b, c, d, t = _leaves(fx, sl)
        summaries.append(step(f, a, b, c, d, t, fx.out_size).summary.clone())
    gsum = torch.stack(summaries).sum(0)
    glob = rdist.unpack_global(gsum, fx.nL, [tuple(c.shape) for c in whole.confusion])
    assert abs(glob["total"].item() - whole.loss.item()) <= 1e-6 * abs(whole.loss.item())
    for L in range(fx.nL):
        assert torch.equal(glob["confusion"][L], whole.confusion[L])
    # pass 2 per rank: data-parallel step (summary reduction between forward and backward), DDP-style mean of the gradients
    step.data_parallel(lambda s: gsum.clone(), world)
    acc = [torch.zeros_like(g) for g in want_params]
    for sl in shards:
        f, a, b, c, d, t = _leaves(fx, sl)
        out = step(f, a, b, c, d, t, fx.out_size)
        assert torch.equal(out.global_summary, gsum)
        out.loss.backward()
        for dst, p in zip(acc, [p for grp in (a, b, c, d) for p in grp]):
            dst += p.grad / world
        for L in range(fx.nL):  # what the donor's DDP would see: d(global loss)/d f_b = local gradient / world
            close(f[L].grad / world, want_dfeats[L][sl], what=f"{name} dfeats{L} shard {sl}")
    for got, want in zip(acc, want_params):
        close(got, want, what=f"{name} parameter gradient")
    step.data_parallel(None, 1)


@pytest.mark.parametrize("name,cap", [("unet_tl", 0), ("hrnet_ext", 0), ("hrnet_ext", 2), ("unet_ext", 1), ("unet_ext", 3)])
def test_level_cap_follows_the_reference_curriculum(name, cap):
    """train.get_loss skips levels L > cur_epoch // pretrain_epoch (train.py:121-134): loss and gradients of the fused
    step with level_cap against the oracle's autograd over the same cap (oracle pinned by tests/golden/glue_*_curriculum)."""
    import rhseg_b200
    fx = Fixture(name)
    step = rhseg_b200.FusedHierStep(fx.tree, fx.level_weights)
    feats, hw, hb, fw, fb, target = _leaves(fx)
    out = step(feats, hw, hb, fw, fb, target, fx.out_size, level_cap=cap)
    out.loss.backward()
    levels, parent_of, _, groups = O.hierarchy_tables(fx.tree)
    mk = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    rf = mk(fx.per_level("feats"))
    rhw, rhb = mk(fx.per_level("head_w")), mk(fx.per_level("head_b"))
    rfw, rfb = mk(fx.per_level("film_w", n=fx.nL - 1)), mk(fx.per_level("film_b", n=fx.nL - 1))
    probs, logits = O.head_forward(rf, rhw, rhb, rfw, rfb, levels, groups, fx.out_size)
    targets = fx.per_level("target")
    onehots, _ = O.predict_onehot_masked([z.detach() for z in logits], targets)
    total, per_level = O.total_loss(logits, targets, fx.level_weights, onehots, levels, parent_of, cur_epoch=cap, pretrain_epoch=1)
    total.backward()
    assert abs(out.loss.item() - total.item()) <= 1e-5 * abs(total.item())
    n_active = min(fx.nL, cap + 1)
    for L in range(fx.nL):
        if L < n_active:
            close(feats[L].grad, rf[L].grad, what=f"{name} dfeats{L}")
            close(hw[L].grad, rhw[L].grad, what=f"{name} dhead_w{L}")
            close(hb[L].grad, rhb[L].grad, what=f"{name} dhead_b{L}")
        else:  # the reference never touches these (their logits are not part of the loss)
            assert feats[L].grad is None and hw[L].grad is None and rf[L].grad is None
    for i in range(fx.nL - 1):
        if i + 1 < n_active:
            close(fw[i].grad, rfw[i].grad, what=f"{name} dfilm_w{i}")
            close(fb[i].grad, rfb[i].grad, what=f"{name} dfilm_b{i}")
        else:
            assert fw[i].grad is None
    # metrics of every level are still reported (train.py:232 runs get_metrics on all levels)
    for L in range(fx.nL):
        onehot = fx.t(f"onehot{L}")
        eval_t = torch.where(fx.t(f"target{L}") == -1, 0, fx.t(f"target{L}"))
        assert torch.equal(out.confusion[L].cpu(), O.level_confusion(onehot, eval_t, onehot.shape[1], L != 0))


def test_peer_exchange_timeout_is_a_hard_failure():
    """a peer that never arrives must not leave stale sums behind.  World of two on one GPU whose
    'peer' is the rank's own area (nobody ever writes the peer's records): the exchange times out, the WHOLE result is
    NaN, the status is sticky, later exchanges stay NaN and PeerExchange.check() raises."""
    from rhseg_b200 import native
    lib = native.lib()
    ctx = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * native.XCHG_HANDLE_BYTES)()
    assert lib.rhseg_xchg_create(4096, 2, ctypes.byref(ctx), handle) == 0
    assert lib.rhseg_xchg_connect(ctx, 0, bytes(handle) * 2) == 0
    assert lib.rhseg_xchg_set_timeout_ms(ctx, -1) < 0
    assert lib.rhseg_xchg_set_timeout_ms(ctx, 50) == 0
    src = torch.arange(600, dtype=torch.float64, device=DEV) + 1.0
    part = torch.ones(1500, device=DEV)
    res = torch.zeros(2100, dtype=torch.float64, device=DEV)
    ptrs = (ctypes.c_void_p * 1)(part.data_ptr())
    cnts = (ctypes.c_long * 1)(1500)
    st = torch.cuda.current_stream().cuda_stream
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    assert lib.rhseg_xchg_all_reduce(ctx, src.data_ptr(), 600, ptrs, cnts, 1, res.data_ptr(), st) == 0
    t1.record()
    torch.cuda.synchronize()
    assert 40.0 <= t0.elapsed_time(t1) < 2000.0          # waited for the limit, not for ever
    assert bool(torch.isnan(res).all())                   # nothing stale or partial survives
    s = ctypes.c_int(0)
    assert lib.rhseg_xchg_status(ctx, ctypes.byref(s)) == 0 and s.value == 1
    res.zero_()
    t0.record()
    assert lib.rhseg_xchg_all_reduce(ctx, src.data_ptr(), 600, ptrs, cnts, 1, res.data_ptr(), st) == 0
    t1.record()
    torch.cuda.synchronize()
    assert t0.elapsed_time(t1) < 40.0 and bool(torch.isnan(res).all())   # sticky: no second wait, still poisoned
    assert lib.rhseg_xchg_destroy(ctx) == 0


def test_dp_grad_scales_kernel():
    from rhseg_b200 import native
    n = 3
    local = torch.zeros(2 + 4 * n + 5, dtype=torch.float64, device=DEV)
    glob = torch.zeros_like(local)
    local[0], glob[0] = 3.0, 8.0
    for L, (nl, ng) in enumerate([(3.0, 8.0), (0.0, 5.0), (2.0, 0.0)]):
        local[4 + 4 * L], glob[4 + 4 * L] = nl, ng
    g = torch.tensor([0.5], device=DEV)
    out = torch.empty(2 * n, device=DEV)
    native.call("rhseg_dp_grad_scales", local.data_ptr(), glob.data_ptr(), n, 4, g.data_ptr(), out.data_ptr(),
                torch.cuda.current_stream().cuda_stream)
    want = torch.tensor([0.5 * 4 * 3 / 8, 0.5 * 4 * 3 / 8, 0.75, 0.0, 0.75, 0.0])
    assert torch.allclose(out.cpu(), want, rtol=1e-6)

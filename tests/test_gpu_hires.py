"""GPU tests of the output-resolution fast paths (bulk-copy pipelines with the level kind and the group layout
fixed at compile time) against the generic kernels of the same entry points, which the fixture tests pin to the
reference.  Integer results (prediction maps, confusion counts, consistency counts) must be identical; sums differ
only by their summation order."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tree(groups, n_roots):
    """Two-level tree: n_roots root classes, the last len(groups) of them have groups[i] children each."""
    roots = {}
    first_parent = n_roots - len(groups)
    for r in range(n_roots):
        kids = groups[r - first_parent] if r >= first_parent else 0
        roots["r%d" % r] = {"r%d_c%d" % (r, j): {} for j in range(kids)}
    return roots


def _targets(tree, B, H, W, gen):
    from oracle import hier_oracle as O
    levels, parent_of, _, groups = O.hierarchy_tables(tree)
    return torch.cat(O.synth_targets(levels, groups, B, H, W, gen), dim=1)


class _Env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _level_eval(tree, L, z, target, prev_idx, want_idx):
    import rhseg_b200
    from rhseg_b200 import native
    t = rhseg_b200.ClassTree(tree)
    K = t.head_channels[L]
    B, _, H, W = z.shape
    off = sum(t.head_channels[:L])
    tgt = target[:, off:off + K]
    ptg = target[:, off - t.head_channels[L - 1]:off] if L > 0 else None
    nc = K + 1 if L > 0 else K
    words = torch.full((B * K * native.NSTAT + native.MAX_K + nc * nc,), 7.0, dtype=torch.float64, device=DEV)
    idx = torch.full((B, H, W), 99, dtype=torch.uint8, device=DEV) if want_idx else None
    native.call("rhseg_level_eval", z.data_ptr(), tgt.data_ptr(), target.stride(0), target.stride(1),
                ptg.data_ptr() if (ptg is not None and prev_idx is not None) else None, target.stride(0), target.stride(1),
                native.ptr(prev_idx), t.device_tables(DEV)[L].data_ptr(), B, K, H * W,
                (1 if L > 0 else 0) | t.group_hint(L), words.data_ptr(), native.ptr(idx), 0,
                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ns = B * K * native.NSTAT
    return words[:ns].clone(), words[ns:ns + native.MAX_K].clone(), words[ns + native.MAX_K:].view(torch.int64).clone(), idx


EVAL_CASES = [([4], 4), ([2], 2), ([3], 4), ([2, 2], 3), ([1, 3], 2), ([5], 3), ([7], 2), ([8], 1), ([3, 4], 4), ([1], 1), ([2, 2, 2], 3)]


@pytest.mark.parametrize("groups,n_roots", EVAL_CASES)
@pytest.mark.parametrize("hw", [(48, 64), (41, 52)])
def test_level_eval_pipeline_equals_generic_kernel(groups, n_roots, hw):
    H, W = hw  # 48x64: everything 16-byte aligned (pipeline); 41x52: N % 16 != 0 -> both runs take the generic kernel
    tree = _tree(groups, n_roots)
    gen = torch.Generator().manual_seed(11)
    B = 3
    target = _targets(tree, B, H, W, gen).to(DEV)
    K0, K1 = n_roots, sum(groups)
    z0 = (torch.randn(B, K0, H, W, generator=gen) * 3).to(DEV)
    z1 = (torch.randn(B, K1, H, W, generator=gen) * 3).to(DEV)
    # exact ties and near ties exercise the bit-exact argmax path
    z1[:, :, :4, :] = z1[:, :1, :4, :]
    if K1 > 1:
        z1[:, 1, 4:8, :] = z1[:, 0, 4:8, :] + 1e-7
    res = {}
    for name, env in (("pipe", None), ("generic", "1")):
        with _Env(RHSEG_NO_EVAL_PIPE=env):
            s0, c0, f0, i0 = _level_eval(tree, 0, z0, target, None, True)
            s1, c1, f1, i1 = _level_eval(tree, 1, z1, target, i0, True)
            s1n, c1n, f1n, _ = _level_eval(tree, 1, z1, target, None, False)  # child level without consistency inputs
        res[name] = (s0, f0, i0, s1, c1, f1, i1, s1n, f1n)
    a, b = res["pipe"], res["generic"]
    for j in (1, 2, 4, 5, 6, 8):  # confusion matrices, index maps, consistency counts: identical
        assert torch.equal(a[j], b[j]), (groups, j)
    for j in (0, 3, 7):  # statistics: same numbers up to the summation order
        assert torch.allclose(a[j], b[j], rtol=2e-6, atol=1e-6), (groups, j, (a[j] - b[j]).abs().max())
    # and against torch on the same device: prediction map and mask counts
    pred = torch.argmax(torch.softmax(z1, 1), 1)
    assert torch.equal(a[6].long(), pred)
    off = K0
    cnt = (target[:, off:off + K1] != -1).sum((2, 3)).double()
    assert torch.equal(a[3].view(B, K1, 5)[:, :, 1], cnt)


FWD_CASES = [([4], 4, 4), ([2], 2, 4), ([3], 4, 4), ([2, 2], 3, 4), ([1, 3], 2, 3), ([7], 2, 4), ([4], 4, 3)]


@pytest.mark.parametrize("groups,n_roots,scale", FWD_CASES)
@pytest.mark.parametrize("vec", ["0", "2"])
def test_band_forward_equals_generic_kernel(groups, n_roots, scale, vec):
    """rhseg_head_level_fwd / _fwd_eval for an upsampled head: band kernel vs the generic kernel (+ separate eval)."""
    import rhseg_b200
    from rhseg_b200 import native
    tree = _tree(groups, n_roots)
    t = rhseg_b200.ClassTree(tree)
    gen = torch.Generator().manual_seed(5)
    B, C, Hf, Wf = 2, 24, 17, 21
    H, W = Hf * scale, (Wf * scale) // 16 * 16
    target = _targets(tree, B, H, W, gen).to(DEV)
    feats = [torch.randn(B, C, Hf, Wf, generator=gen).to(DEV) for _ in range(2)]
    st = torch.cuda.current_stream().cuda_stream
    tabs = t.device_tables(DEV)

    def run():
        outs = []
        prev_p, prev_idx = None, None
        for L in range(2):
            K = t.head_channels[L]
            K_prev = t.head_channels[L - 1] if L else 0
            eff_w = (torch.randn(B, K, C, generator=torch.Generator().manual_seed(L)) * 0.3).to(DEV)
            eff_b = (torch.randn(B, K, generator=torch.Generator().manual_seed(10 + L)) * 0.3).to(DEV)
            z_lo = torch.zeros(B, K, Hf, Wf, device=DEV)
            z = torch.full((B, K, H, W), 5.0, device=DEV)
            p = torch.full((B, K, H, W), 5.0, device=DEV)
            psum = torch.zeros(B, K, dtype=torch.float64, device=DEV)
            nc = K + 1 if L else K
            words = torch.zeros(B * K * native.NSTAT + native.MAX_K + nc * nc, dtype=torch.float64, device=DEV)
            idx = torch.full((B, H, W), 77, dtype=torch.uint8, device=DEV)
            off = sum(t.head_channels[:L])
            tg = target[:, off:off + K]
            ptg = target[:, off - K_prev:off] if L else None
            native.call("rhseg_head_level_fwd_eval", feats[L].data_ptr(), eff_w.data_ptr(), eff_b.data_ptr(), native.ptr(prev_p),
                        tabs[L].data_ptr(), B, C, Hf, Wf, H, W, K, K_prev, t.act_mode[L] | t.group_hint(L), z_lo.data_ptr(),
                        z.data_ptr(), p.data_ptr(), psum.data_ptr(), tg.data_ptr(), target.stride(0), target.stride(1),
                        native.ptr(ptg), target.stride(0), target.stride(1), native.ptr(prev_idx), words.data_ptr(), idx.data_ptr(),
                        1 | 2, st)
            # the same level without the fused evaluation
            z2, p2 = torch.empty_like(z), torch.empty_like(p)
            psum2 = torch.zeros_like(psum)
            z_lo2 = torch.zeros_like(z_lo)
            native.call("rhseg_head_level_fwd", feats[L].data_ptr(), eff_w.data_ptr(), eff_b.data_ptr(), native.ptr(prev_p),
                        tabs[L].data_ptr(), B, C, Hf, Wf, H, W, K, K_prev, t.act_mode[L] | t.group_hint(L), z_lo2.data_ptr(),
                        z2.data_ptr(), p2.data_ptr(), psum2.data_ptr(), 2, st)
            torch.cuda.synchronize()
            outs.append((z, p, psum, words, idx, z2, p2, psum2, z_lo))
            prev_p, prev_idx = p, idx
        return outs

    with _Env(RHSEG_NO_BAND_FWD=None, RHSEG_TUNE_UP_VEC=vec):
        band = run()
    with _Env(RHSEG_NO_BAND_FWD="1"):
        gen_ = run()
    for L in range(2):
        K = t.head_channels[L]
        zb, pb, sb, wb, ib, z2b, p2b, s2b, zlo = band[L]
        zg, pg, sg, wg, ig, z2g, p2g, s2g, _ = gen_[L]
        ref = torch.nn.functional.interpolate(zlo, size=(H, W), mode="bilinear", align_corners=True)
        assert torch.allclose(zb, ref, rtol=1e-5, atol=2e-6), (L, (zb - ref).abs().max())
        assert torch.allclose(zb, zg, rtol=1e-5, atol=2e-6)
        assert torch.allclose(pb, pg, rtol=1e-5, atol=1e-6), (L, (pb - pg).abs().max())
        assert torch.equal(zb, z2b) and torch.equal(pb, p2b)  # with and without the fused evaluation: same outputs
        assert torch.allclose(sb, sg, rtol=1e-6) and torch.allclose(s2b, sg, rtol=1e-6)
        ns = B * K * native.NSTAT
        # evaluation of the band kernel == stand-alone evaluation of ITS logits (identical inputs -> identical integers)
        tree_eval = _level_eval(tree, L, zb, target, band[L - 1][4] if L else None, True)
        assert torch.equal(ib, tree_eval[3]), L
        assert torch.equal(wb[ns + native.MAX_K:].view(torch.int64), tree_eval[2]), L
        assert torch.equal(wb[ns:ns + native.MAX_K], tree_eval[1]), L
        assert torch.allclose(wb[:ns], tree_eval[0], rtol=2e-6, atol=1e-6), L

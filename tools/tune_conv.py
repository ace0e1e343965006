"""Dev tool: times the two feature-streaming kernels (level forward conv, conv backward)
through the C-ABI for the pipeline shapes selectable with RHSEG_TUNE_* (K=4 only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rhseg_b200  # noqa: E402
from rhseg_b200 import native  # noqa: E402
from rhseg_b200.native import call, ptr  # noqa: E402

PEAK = 6552.3


def time_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(kind):
    dev = "cuda"
    B, K = 4, 4
    if kind == "unet":
        C, Hf, Wf, H, W = 64, 620, 620, 620, 620
    else:
        C, Hf, Wf, H, W = 720, 155, 155, 620, 620
    Nf = Hf * Wf
    tree = rhseg_b200.ClassTree({"a": {}, "b": {}, "c": {}, "d": {}})
    tab = tree.device_tables(dev)
    st = torch.cuda.current_stream().cuda_stream
    # two feature sets alternate so that no launch finds its input in L2
    feats = [torch.randn(B, C, Hf, Wf, device=dev) for _ in range(2)]
    dfe = torch.empty_like(feats[0])
    eff_w = torch.randn(B, K, C, device=dev) * 0.1
    eff_b = torch.randn(B, K, device=dev)
    z_lo = torch.empty(B, K, Hf, Wf, device=dev)
    z = torch.empty(B, K, H, W, device=dev)
    p = torch.empty(B, K, H, W, device=dev)
    psum = torch.zeros(B, K, dtype=torch.float64, device=dev)
    dz = torch.randn(B, K, Hf, Wf, device=dev)
    S = torch.zeros(B, K, C, dtype=torch.float64, device=dev)
    s = torch.zeros(B, K, dtype=torch.float64, device=dev)
    cnt = [0]

    def fwd():
        cnt[0] += 1
        call("rhseg_head_level_fwd", ptr(feats[cnt[0] & 1]), ptr(eff_w), ptr(eff_b), None, ptr(tab[0]), B, C, Hf, Wf, H, W, K, 0, 0,
             ptr(z_lo), ptr(z), ptr(p), ptr(psum), 0, st)

    def bwd():
        cnt[0] += 1
        call("rhseg_head_conv_bwd", ptr(feats[cnt[0] & 1]), ptr(dz), ptr(eff_w), B, C, K, Nf, ptr(dfe), ptr(S), ptr(s), 0, st)

    fwd_bytes = B * (4 * C * Nf + 4 * K * Nf) + (0 if kind == "unet" else 0)
    bwd_bytes = B * (2 * 4 * C * Nf + 4 * K * Nf)
    v = "V4" if kind == "unet" else "S1"
    for t in range(4):
        os.environ["RHSEG_TUNE_FWD_" + v] = str(t)
        ms = time_ms(fwd)
        extra = "" if kind == "unet" else " (incl. hi-res upsample+act pass)"
        print(f"{kind} fwd tune={t}: {ms*1e3:8.1f} us  {fwd_bytes/ms/1e6:7.0f} GB/s ({fwd_bytes/ms/1e6/PEAK*100:4.1f}% of {PEAK}){extra}", flush=True)
    os.environ["RHSEG_TUNE_FWD_" + v] = "0"
    for t in range(4):
        os.environ["RHSEG_TUNE_BWD_" + v] = str(t)
        ms = time_ms(bwd)
        print(f"{kind} bwd tune={t}: {ms*1e3:8.1f} us  {bwd_bytes/ms/1e6:7.0f} GB/s ({bwd_bytes/ms/1e6/PEAK*100:4.1f}% of {PEAK})", flush=True)
    os.environ["RHSEG_TUNE_BWD_" + v] = "0"


if __name__ == "__main__":
    for kind in sys.argv[1:] or ["unet", "hrnet"]:
        run(kind)

#!/usr/bin/env python
"""Benchmark of the restrictive-hierarchy head + loss + metrics path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = the body of the reference's train_epoch (train.py:201-241) minus the donor
backbone and the optimiser, on synthetic per-level donor features:
    head forward (per-level feats -> probs, logits)  ->  train-path prediction + confusion-
    matrix metrics  ->  CE + Dice per level + consistency  ->  backward to per-level dfeats and
    head / FiLM parameter gradients.
`value` is whole-job Mpixel/s (B*H*W pixels per step per GPU) with inputs resident in HBM;
`e2e` is the same step driven from pinned HOST buffers (H2D of the step's features and targets
and D2H of the loss + metrics inside the timed region).

Prints ONE JSON line on rank 0.  `--impl reference` times the oracle port of the reference's
CPU path (oracle/hier_oracle.py; the Python reference itself cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

TL = {"background": {}, "upper": {}, "lower": {}, "tooth": {"pulp": {}, "dentin": {}, "enamel": {}, "composite": {}}}
EXT = {"background": {}, "tooth+alveolar": {"alveolar": {"upper": {}, "lower": {}},
                                            "tooth": {"composite": {}, "healthy": {"pulp": {}, "dentin": {}, "enamel": {}}}}}
W_TL = [[0.0297, 1.577, 0.9619, 0.1770], [1.5432, 0.2638, 1.0413, 3.9722]]  # reference README.md:71

# BASELINE.json configs -> workloads (per-GPU batch; weak scaling)
WORKLOADS = {
    "hrnet_w48_tl_620_b4": dict(tree=TL, kind="hrnet", C=720, H=620, W=620, scale=4, B=4, weights=W_TL),   # configs[1]
    "unet_tl_620_b4": dict(tree=TL, kind="unet", C=64, H=620, W=620, scale=1, B=4, weights=W_TL),          # configs[0]
    "hrnet_w48_ext_620_b4": dict(tree=EXT, kind="hrnet", C=720, H=620, W=620, scale=4, B=4, weights=None),  # configs[2]
    "unet_ext_620_b4": dict(tree=EXT, kind="unet", C=64, H=620, W=620, scale=1, B=4, weights=None),
    "unet_tl_1024_b8": dict(tree=TL, kind="unet", C=64, H=1024, W=1024, scale=1, B=8, weights=W_TL),       # configs[4] at 8 GPUs
}
DEFAULT_WORKLOAD = "hrnet_w48_tl_620_b4"


def feat_hw(wl):
    if wl["scale"] == 1:
        return wl["H"], wl["W"]
    return (wl["H"] + wl["scale"] - 1) // wl["scale"], (wl["W"] + wl["scale"] - 1) // wl["scale"]


def algorithmic_bytes(wl, tree_channels, groups_per_level, B):
    """SURVEY.md 8(d) formulas, fp32, every API-visible tensor moved once per pass."""
    h, w = feat_hw(wl)
    N, Nf, C = wl["H"] * wl["W"], h * w, wl["C"]
    out = dict(fwd=0, loss=0, bwd=0, metrics=0, conv_bwd=[], fwd_conv=[])
    for K, g in zip(tree_channels, groups_per_level):
        out["fwd"] += B * (4 * C * Nf + 2 * 4 * K * N + 4 * g * N)
        out["loss"] += B * (4 * K + 4 * K) * N
        out["bwd"] += B * (2 * 4 * C * Nf + (4 * K + 4 * K + 4 * g) * N)
        out["metrics"] += B * (4 * K + 4 * K) * N
        # dominant kernel (1x1-conv backward): read feats + write dfeats + read dz at feature res
        out["conv_bwd"].append(B * (2 * 4 * C * Nf + 4 * K * Nf))
        out["fwd_conv"].append(B * (4 * C * Nf + 4 * K * Nf))
    out["step"] = out["fwd"] + out["loss"] + out["bwd"]
    return out


def synth_inputs(wl, B, seed, device, pin=False):
    """SURVEY.md 8(d) synthetic inputs: N(0,1) features per level, ternary targets built with the
    dataset's ignore rule, default-style head / FiLM parameter init."""
    from oracle import hier_oracle as O  # input generator only (shared with the tests)
    levels, parent_of, _, groups = O.hierarchy_tables(wl["tree"])
    chans = [len(levels[0])] + [sum(len(k) for _, k in g) for g in groups]
    h, w = feat_hw(wl)
    g = torch.Generator().manual_seed(seed)
    C = wl["C"]
    feats = [torch.randn(B, C, h, w, generator=g) for _ in chans]
    bound = 1.0 / (C ** 0.5)
    hw = [(torch.rand(k, C, 1, 1, generator=g) * 2 - 1) * bound for k in chans]
    hb = [(torch.rand(k, generator=g) * 2 - 1) * bound for k in chans]
    fw = [(torch.rand(2 * C, kp, generator=g) * 2 - 1) / (kp ** 0.5) for kp in chans[:-1]]
    fb = [(torch.rand(2 * C, generator=g) * 2 - 1) / (kp ** 0.5) + 1.0 for kp in chans[:-1]]
    targets = torch.cat(O.synth_targets(levels, groups, B, wl["H"], wl["W"], g), dim=1)
    weights = wl["weights"] or [[1.0] * k for k in chans]
    host = dict(feats=feats, hw=hw, hb=hb, fw=fw, fb=fb, target=targets)
    if pin:
        host["feats"] = [f.pin_memory() for f in feats]
        host["target"] = targets.pin_memory()
    return dict(levels=levels, parent_of=parent_of, groups=groups, chans=chans, weights=weights, host=host)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class GpuStep:
    """Device-resident state + the step through the package's public API."""

    def __init__(self, wl, data, device):
        import rhseg_b200
        from rhseg_b200 import metric_ops
        from rhseg_b200.Metrics import losses
        self.rh, self.metric_ops, self.losses = rhseg_b200, metric_ops, losses
        self.wl, self.data, self.dev = wl, data, device
        self.tree = rhseg_b200.ClassTree(wl["tree"])
        h = data["host"]
        self.feats = [f.to(device).requires_grad_(True) for f in h["feats"]]
        self.params = [[p.to(device).requires_grad_(True) for p in h[k]] for k in ("hw", "hb", "fw", "fb")]
        self.target = h["target"].to(device)
        self.out_size = None if wl["scale"] == 1 else (wl["H"], wl["W"])
        self.ce, self.dice = losses.CrossEntropyLoss(), losses.SoftDiceLoss()
        self.fused = rhseg_b200.FusedHierStep(self.tree, data["weights"])
        # room behind the step summary for the head / FiLM parameter gradients (the data-parallel exchange buffer)
        self.fused.exchange_tail = sum(p.numel() for grp in self.params for p in grp)
        self.result = None
        self.one = torch.ones((), device=device)  # d(loss)/d(loss), allocated once instead of a fill per step

    def targets(self):
        out, s = [], 0
        for k in self.data["chans"]:  # channel slices of the wide target tensor (train.py:185-193)
            out.append(self.target[:, s:s + k])
            s += k
        return out

    def step(self):
        """Fused step (rhseg_b200.FusedHierStep): the whole train_epoch body in one autograd node."""
        hw, hb, fw, fb = self.params
        out = self.fused(self.feats, hw, hb, fw, fb, self.target, self.out_size)
        leaves = self.feats + [p for grp in self.params for p in grp]
        self.grads = torch.autograd.grad(out.loss, leaves, grad_outputs=self.one)  # dfeats per level + head / FiLM parameter grads
        self.result = (out.scalars, out.ratios)
        self.confusion, self.summary, self.exchange = out.confusion, out.summary, out.exchange
        return self.result

    def step_dropin(self):
        """Same work through the drop-in modules, called the way train.py calls them."""
        self.rh.clear_memo()
        hw, hb, fw, fb = self.params
        probs, logits = self.rh.hier_head_forward(self.tree, self.feats, hw, hb, fw, fb, self.out_size)
        targets = self.targets()
        onehots, ratios = [], []
        for L, z in enumerate(logits):  # train.py:206-232: prediction glue + the five metrics
            onehot, eval_t = self.metric_ops.predict_onehot(z.detach(), targets[L])
            ratios.append(self.metric_ops.level_ratios(onehot, eval_t, L != 0))
            onehots.append(onehot)
        loss = None
        for L, z in enumerate(logits):  # train.py:132-143
            w = self.data["weights"][L]
            ce = self.ce(z, targets[L], class_weight=w, logits_input=True)
            di = self.rh.level_loss(z, targets[L], w, 0.0, True).dice  # tensor form of SoftDiceLoss (no host sync)
            loss = ce + di if loss is None else loss + ce + di
        loss = loss + self.losses.hierarchical_consistency_loss(onehots, self.tree.levels, self.tree.parent_of)
        leaves = self.feats + [p for grp in self.params for p in grp]
        grads = torch.autograd.grad(loss, leaves)
        return loss.detach(), ratios, grads


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                out["sm_max_mhz"] = float(f[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"], out["samples"] = sorted(reasons), len(sm)
        return out


def kernels_per_call(name, upsampled):
    if name == "rhseg_head_level_fwd":
        return 2 if upsampled else 1
    return 1


def run_ours(args, rank, world, local_rank):
    from rhseg_b200 import native
    wl = WORKLOADS[args.workload]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = wl["B"]
    data = synth_inputs(wl, B, seed=1000 + rank, device=dev, pin=True)
    st = GpuStep(wl, data, dev)
    upsampled = wl["scale"] != 1
    alg = algorithmic_bytes(wl, data["chans"], [0] + [len(g) for g in data["groups"]], B)

    # kernel-launch accounting + live timing of the dominant kernel (1x1-conv backward)
    counted = {"n": 0}
    conv_events = []
    timing_on = {"v": False}
    raw_call = native.call

    def counting_call(name, *a):
        counted["n"] += kernels_per_call(name, upsampled)
        if timing_on["v"] and name in ("rhseg_head_conv_bwd", "rhseg_head_conv_bwd_params"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            raw_call(name, *a)
            e1.record()
            conv_events.append((e0, e1))
        else:
            raw_call(name, *a)

    import rhseg_b200.fused as fused_mod
    import rhseg_b200.head as head_mod
    import rhseg_b200.loss_ops as loss_mod
    import rhseg_b200.metric_ops as met_mod
    for mod in (native, head_mod, loss_mod, met_mod, fused_mod):
        mod.call = counting_call

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    from rhseg_b200 import dist as rdist

    peer = {"px": None, "kind": "none" if world == 1 else "nccl"}

    def exchange(result):
        """The path's only cross-rank step: ONE all-reduce of the packed step summary (loss terms,
        valid-sample counts, confusion matrices; rhseg_b200.dist) + the head/FiLM parameter gradients
        (a few thousand floats).  Pixel data never leaves its GPU.  Single node: one kernel over NVLink
        peer memory (rhseg_xchg_all_reduce); RHSEG_EXCHANGE=nccl (or no P2P) -> rhseg_pack_f64 + NCCL."""
        if world == 1:
            return
        grads = st.grads[len(st.feats):]
        if peer["px"] is not None:
            st.global_summary = peer["px"].all_reduce(st.summary, grads, out=st.exchange)
            return
        buf = rdist.pack_exchange(st.summary, grads, out=st.exchange)  # one rhseg_pack_f64 launch
        torch.distributed.all_reduce(buf)
        st.global_summary = buf

    if world > 1 and os.environ.get("RHSEG_EXCHANGE", "p2p") == "p2p":
        st.step()  # sizes the exchange buffer
        try:
            peer["px"] = rdist.PeerExchange(st.exchange.numel())
            peer["kind"] = "p2p"
        except Exception as e:
            sys.stderr.write("peer-memory exchange unavailable (%r); using NCCL\n" % (e,))

    # ---- device-resident timing (`value`) ----
    for _ in range(max(args.warmup, 3)):
        exchange(st.step())
    graph = None
    if args.graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    st.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                exchange(st.step())  # NCCL all-reduce is captured with the step (no host work per replay)
            g.replay()
            torch.cuda.synchronize()
            graph = g
        except Exception as e:  # capture is an optimisation, never a requirement
            sys.stderr.write("cuda graph capture unavailable (%r); timing eagerly\n" % (e,))
            graph = None
            torch.cuda.synchronize()

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            exchange(st.step())

    for _ in range(3):
        one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    counted["n"] = 0
    st.step() if graph is None else None
    launches_per_step = counted["n"] if graph is None else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if launches_per_step is None:  # graph mode: count one eager step after timing
        counted["n"] = 0
        st.step()
        launches_per_step = counted["n"]
        torch.cuda.synchronize()

    # ---- the same step through the drop-in modules (reference call sequence), eager ----
    dropin_ms = float("nan")
    if not args.no_dropin:
        for _ in range(3):
            st.step_dropin()
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(min(args.steps, 20)):
            st.step_dropin()
        d1.record()
        torch.cuda.synchronize()
        dropin_ms = d0.elapsed_time(d1) / min(args.steps, 20)

    # ---- dominant-kernel timing, live, on the launching stream (eager steps, inputs > L2) ----
    timing_on["v"] = True
    for _ in range(min(args.steps, 20)):
        st.step()
    torch.cuda.synchronize()
    timing_on["v"] = False
    nL = len(data["chans"])
    per_level = [[] for _ in range(nL)]
    for i, (a, b) in enumerate(conv_events):
        per_level[nL - 1 - (i % nL)].append(a.elapsed_time(b))  # backward visits the last level first
    conv_ms = [statistics.mean(v) for v in per_level]

    # the same kernel on its own: back-to-back launches between ONE event pair (no event / launch gap per launch;
    # 0.55-0.8 GB of features + gradients per launch, far beyond the 126 MB L2).  Reported beside the in-step figure.
    dom_l = max(range(nL), key=lambda L: conv_ms[L])
    K_dom, (hf, wf) = data["chans"][dom_l], feat_hw(wl)
    iso = dict(dz=torch.randn(B, K_dom, hf, wf, device=dev), w=torch.randn(B, K_dom, wl["C"], device=dev),
               df=torch.empty_like(st.feats[dom_l]), S=torch.zeros(B, K_dom, wl["C"], dtype=torch.float64, device=dev),
               s=torch.zeros(B, K_dom, dtype=torch.float64, device=dev))
    cur = torch.cuda.current_stream().cuda_stream

    def conv_alone():
        raw_call("rhseg_head_conv_bwd", st.feats[dom_l].data_ptr(), iso["dz"].data_ptr(), iso["w"].data_ptr(), B, wl["C"], K_dom,
                 hf * wf, iso["df"].data_ptr(), iso["S"].data_ptr(), iso["s"].data_ptr(), 0, cur)
    conv_alone_ms = float("nan")
    if not args.no_alone:
        for _ in range(3):
            conv_alone()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(20):
            conv_alone()
        i1.record()
        torch.cuda.synchronize()
        conv_alone_ms = i0.elapsed_time(i1) / 20
    del iso
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end from pinned host buffers (`e2e`) ----
    host = data["host"]
    h2d = sum(f.numel() * 4 for f in host["feats"]) + host["target"].numel() * 4
    out_host = torch.empty(2 + 4 * len(data["chans"]) + sum(5 * (k + (1 if L else 0)) for L, k in enumerate(data["chans"])),
                           dtype=torch.float32).pin_memory()
    d2h = out_host.numel() * 4

    # two device-side input sets: the H2D copy of step i+1 runs on a copy stream while step i computes
    sets = [(st.feats, st.target),
            ([torch.empty_like(f).requires_grad_(True) for f in st.feats], torch.empty_like(st.target))]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        feats_i, target_i = sets[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])  # the step that last read this set has finished
            with torch.no_grad():
                for dst, src in zip(feats_i, host["feats"]):
                    dst.copy_(src, non_blocking=True)
                target_i.copy_(host["target"], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_step(i):
        upload(i + 1)                                   # prefetch the next step's inputs
        torch.cuda.current_stream().wait_event(ready[i % 2])
        st.feats, st.target = sets[i % 2]
        scal, ratios = st.step()
        consumed[i % 2].record()
        exchange((scal, ratios))
        out_host.copy_(torch.cat([scal] + [r.flatten() for r in ratios]), non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads loss / metrics on the host every step

    for ev in consumed:
        ev.record()
    upload(0)
    for i in range(2):
        e2e_step(i)
    barrier()
    e2e_steps = max(3, min(args.steps, 30))
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(2, 2 + e2e_steps):
        e2e_step(i)
    t1.record()
    barrier()
    copy_stream.synchronize()
    e2e_ms = t0.elapsed_time(t1) / e2e_steps
    st.feats, st.target = sets[0]

    times = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    if rank != 0:
        return None

    px = B * wl["H"] * wl["W"]
    peaks, peak_src = load_peaks()
    dom = max(range(nL), key=lambda L: conv_ms[L])
    achieved = alg["conv_bwd"][dom] / (conv_ms[dom] * 1e-3) / 1e9
    line = {
        "metric": "hier head+loss fwd+bwd Mpixel/s", "value": world * px / (ms * 1e-3) / 1e6, "unit": "Mpixel/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "tree_levels": data["chans"], "feat_channels": wl["C"],
                   "batch_per_gpu": B, "image": [wl["H"], wl["W"]], "feat_hw": list(feat_hw(wl)),
                   "step": "head fwd + train-path prediction + 5 confusion metrics + CE/Dice/consistency + bwd (dfeats, head+FiLM grads); fused step API",
                   "l2_policy": "inputs larger than L2 (%.0f MB of features per step vs 126 MB L2)" % (sum(f.numel() * 4 for f in st.feats) / 1e6),
                   "cuda_graph": graph is not None, "collective": ("1 all-reduce/step (loss+metrics+head grads), %s" % {"p2p": "one peer-memory kernel over NVLink (rhseg_xchg_all_reduce)", "nccl": "rhseg_pack_f64 + NCCL"}[peer["kind"]]) if world > 1 else "none"},
        "step_bytes": {"algorithmic_head_loss_fwd_bwd": alg["step"], "metrics": alg["metrics"],
                       "frac_of_hbm_peak_whole_step": (alg["step"] + alg["metrics"]) / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "roofline": {"kernel": "conv_bwd_kernel (1x1-conv backward, level %d)" % dom, "bound": "hbm", "achieved": achieved,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": load_traffic(args.workload), "peak_source": peak_src,
                     "bytes_per_launch": alg["conv_bwd"][dom], "ms_per_launch": conv_ms[dom],
                     "per_level_ms": conv_ms,
                     "timing": "in-step: one CUDA-event pair around each launch inside eager steps (includes the launch gap the event pair opens)",
                     "alone": {"ms_per_launch": conv_alone_ms, "achieved": alg["conv_bwd"][dom] / (conv_alone_ms * 1e-3) / 1e9,
                               "frac": alg["conv_bwd"][dom] / (conv_alone_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                               "timing": "20 back-to-back launches of the same kernel between one event pair"}},
        "clocks": clocks,
        "e2e": {"value": world * px / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "note": "features+targets copied from pinned host memory every step (copy of step i+1 overlaps step i); PCIe-bound"},
        "dropin_modules": {"ms_per_step": dropin_ms, "value": px / (dropin_ms * 1e-3) / 1e6, "unit": "Mpixel/s",
                           "note": "same step through Models/Metrics drop-in modules in the reference's call order, eager, this rank"},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(wl, steps=20, warmup=2, sample_b=1)  # ~10 s of host work
    return line


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "MEASURED_PEAKS.json (measured copy bandwidth on this pool's B200)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback 6.65 TB/s (B200_PROFILING.md)"


def load_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# reference arm: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------
def cpu_reference(wl, steps, warmup, sample_b):
    from oracle import hier_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    data = synth_inputs(wl, sample_b, seed=7, device="cpu")
    h = data["host"]
    levels, parent_of, groups = data["levels"], data["parent_of"], data["groups"]
    out_size = None if wl["scale"] == 1 else (wl["H"], wl["W"])
    leaves = [[t.clone().requires_grad_(True) for t in h[k]] for k in ("feats", "hw", "hb", "fw", "fb")]
    targets, s = [], 0
    for k in data["chans"]:
        targets.append(h["target"][:, s:s + k])
        s += k

    def step():
        for grp in leaves:
            for t in grp:
                t.grad = None
        probs, logits = O.head_forward(*leaves, levels, groups, out_size)
        onehots, eval_t = O.predict_onehot_masked([z.detach() for z in logits], targets)
        O.all_level_metrics(onehots, eval_t)
        loss, _ = O.total_loss(logits, targets, data["weights"], onehots, levels, parent_of)
        loss.backward()
        return loss.item()

    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = statistics.mean(ts)
    px = sample_b * wl["H"] * wl["W"]
    return {"value": px / sec / 1e6, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d image(s) of the workload per step (%dx%d, %d levels), %d timed steps, torch %s CPU fp32"
                      % (sample_b, wl["H"], wl["W"], len(data["chans"]), steps, torch.__version__),
            "ms_per_step": sec * 1e3}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    wl = WORKLOADS[args.workload]
    base = cpu_reference(wl, steps=args.steps, warmup=args.warmup, sample_b=1)
    return {"impl": "reference", "metric": "hier head+loss fwd+bwd Mpixel/s", "value": base["value"], "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": "oracle port of the reference CPU path on rank 0's host cores"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip timing the drop-in module path (profiling runs)")
    ap.add_argument("--no-alone", action="store_true", help="skip the back-to-back timing of the dominant kernel (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    if world > 1:
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        line = run_ours(args, rank, world, local_rank)
        if line is not None:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

"""ORACLE / baseline infrastructure — never product code.  The reference's OWN CPU implementation of the hot path, timed.

Used by bench.py's `cpu_baseline` leg and by `bench.py --impl reference` (the only places allowed to execute oracle/).
The modules are the UNMODIFIED reference files (Models/models.py, Metrics/losses.py, Metrics/performance_metrics.py,
train.py) imported from /root/reference in the build container or from the byte copies under oracle/_ref on the GPU
box (tools/stage_reference.py), through the stub modules of oracle/ref_loader.py.  One step is the body of
train.train_epoch (train.py:201-241) between the donor backbone and the optimiser, exactly as SURVEY.md 8(d) states it:

    model(x)  with _run_unet / _forward_backbone patched to return the step's synthetic donor features (the heads,
              FiLMs, grouped softmax, composition and - HRNet - the bilinear upsample are the reference's)
    -> the prediction glue of train.py:206-231  -> train.get_metrics (:38-81)  -> train.get_loss (:111-152)
    -> loss.backward()  (gradients reach the features and every head / FiLM parameter)

Two forced deviations: get_loss is called without the two keyword arguments its signature lacks (train.py:239, SURVEY
F6), and `import torchmetrics` resolves to the stand-in of oracle/ref_loader.py (the package is not installable offline;
it implements the five multiclass constructors the reference calls over the restated semantics of hier_oracle.py).
The metric objects are the reference's own Metrics.performance_metrics classes (ProcessClasses + five wrappers, each
rebuilding argmax + confusion matrix as upstream).  If the reference tree is not available the caller falls back to
the oracle port (kind "port").
"""
import os
import statistics
import time
import types

import torch

from oracle import ref_loader


def available():
    return ref_loader.available()


class ReferenceStep:
    """The reference modules set up for one workload of bench.py (`wl`) and one batch of synthetic inputs (`host`:
    feats / hw / hb / fw / fb / target as bench.synth_inputs makes them)."""

    def __init__(self, wl, data):
        self.models, self.losses, self.train = ref_loader.load_reference()
        if self.train is None:
            raise RuntimeError("reference train.py is not importable")
        self.wl, self.data = wl, data
        h = data["host"]
        self.kind = wl["kind"]
        tree = wl["tree"]
        torch.manual_seed(0)
        if self.kind == "flat":
            self.model = None
            self.logits = h["logits"].clone().requires_grad_(True)
            self.loss_fns = [[self.losses.CrossEntropyLoss(), self.losses.SoftDiceLoss(num_classes=wl["K"])]]
        else:
            if self.kind == "unet":
                m = self.models.UNet(size=wl["H"], n_channels=3, hierarchy=tree, model_type=1)
                heads = [hd.conv for hd in m.heads]
            else:
                m = self.models.HighResolutionNet(ref_loader.hrnet_config(), hierarchy=tree, model_type=1)
                heads = list(m.classifiers)
            films = [f.mlp[1] for f in m.films]
            with torch.no_grad():  # the step's synthetic head / FiLM parameters (same values as our arm's)
                for L, hd in enumerate(heads):
                    hd.weight.copy_(h["hw"][L].view_as(hd.weight))
                    hd.bias.copy_(h["hb"][L])
                for i, f in enumerate(films):
                    f.weight.copy_(h["fw"][i])
                    f.bias.copy_(h["fb"][i])
            m.train()
            self.model, self.heads, self.films = m, heads, films
            self.feats = [f.clone().requires_grad_(True) for f in h["feats"]]
            self.loss_fns = [[self.losses.CrossEntropyLoss(), self.losses.SoftDiceLoss(num_classes=k)] for k in data["chans"]]
        self.targets, s = [], 0
        for k in data["chans"]:  # train.py:185-193
            self.targets.append(h["target"][:, s:s + k])
            s += k
        pm = ref_loader.load_reference_module("Metrics.performance_metrics")
        self.metric_objs = [pm.Accuracy(), pm.Jaccardindex(), pm.DiceScore(), pm.Precision(), pm.Recall()]
        self.x = torch.zeros(h["target"].shape[0], 3, wl["H"], wl["W"])
        self.args = types.SimpleNamespace()

    def step(self):
        train, wl = self.train, self.wl
        nK = sum(self.data["chans"])
        if self.kind == "flat":
            self.logits.grad = None
            logits = [self.logits]
        else:
            for t in self.feats + [p for m in self.heads + self.films for p in m.parameters()]:
                t.grad = None
            it = iter(self.feats)
            if self.kind == "unet":
                self.model._run_unet = lambda x: next(it)
                _, logits = self.model(self.x, type=1, hierarchy=wl["tree"])
            else:
                self.model._forward_backbone = lambda x: next(it)
                _, logits = self.model(self.x)
        targets = self.targets
        # train.py:206-231, the reference's own lines (they live inside train_epoch's loop body)
        output_class = [torch.nn.functional.one_hot(torch.argmax(torch.softmax(z, dim=1), dim=1), num_classes=z.shape[1])
                        .permute(0, 3, 1, 2).float() for z in logits]
        eval_targets = list(targets)
        for i in range(len(targets)):
            output_class[i] = torch.where(targets[i] == -1, 0, output_class[i])
            eval_targets[i] = torch.where(targets[i] == -1, 0, targets[i])
        clss = [{k: [] for k in ("accuracy", "iou", "dice", "precision", "recall")} for _ in range(nK)]
        acc, iou, dice, prec, rec = [], [], [], [], []
        a, j, d, p, r = self.metric_objs
        train.get_metrics(output_class, eval_targets, acc, iou, dice, prec, rec, a, j, d, p, r, "cpu", clss, self.args)
        flat = self.kind == "flat"  # train.py:236-239: probs_per_level / model only for hierarchical models
        loss, _, _ = train.get_loss(logits, targets, self.loss_fns, [], self.data["weights"], 0.0, [], cur_epoch=0,
                                    pretrain_epoch=None, probs_per_level=None if flat else output_class,
                                    model=None if flat else self.model)
        loss.backward()
        return float(loss.item())


def time_reference(wl, data, steps, warmup):
    """Seconds per step (mean and all samples) of the reference's CPU path on every host core torch can use."""
    torch.set_num_threads(os.cpu_count() or 1)
    st = ReferenceStep(wl, data)
    loss = None
    for _ in range(warmup):
        loss = st.step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        loss = st.step()
        ts.append(time.perf_counter() - t0)
    ref_loader.restore_package_imports()
    return statistics.mean(ts), ts, loss

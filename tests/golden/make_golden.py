"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through ref_shim.py) on small seeded inputs.  Build-container only: the
reference tree does not exist on the GPU box, the committed .npz files travel instead.

    python tests/golden/make_golden.py

Each fixture stores the inputs (per-level donor features, head/FiLM parameters, ternary
targets, class weights) and the reference's outputs: per-level probabilities and logits,
CE / Dice / consistency / total loss (train.get_loss), the train-path masked one-hot
predictions, and the autograd gradients of the total loss w.r.t. every feature tensor and
every head/FiLM parameter.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

models, losses, train = ref_shim.load_reference()
from oracle import hier_oracle as O  # noqa: E402  (only for the synthetic target generator)

TL = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "class_tree_tl.json")))
EXT = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "class_tree_tl_extended.json")))
# adversarial trees: leaf at depth 0 + multi-group level + single-child group
ADV = {"a": {}, "b": {"b0": {}, "b1": {"b1x": {}}}, "c": {"c0": {}, "c1": {}, "c2": {}}}
W_TL = [[0.0297, 1.577, 0.9619, 0.1770], [1.5432, 0.2638, 1.0413, 3.9722]]  # README.md:71
W_FLAT = [0.0285, 1.5159, 0.9227, 1.4842, 0.2532, 1.0, 3.8021]  # README.md:79


def weights_for(tree_name, levels):
    if tree_name == "tl":
        return W_TL
    return [[1.0 + 0.25 * ((i + L) % 3) for i in range(len(lv))] for L, lv in enumerate(levels)]


def build_model(kind, tree):
    if kind == "unet":
        m = models.UNet(size=64, n_channels=3, hierarchy=tree, model_type=1)
        heads = [h.conv for h in m.heads]
    else:
        m = models.HighResolutionNet(ref_shim.hrnet_config(), hierarchy=tree, model_type=1)
        heads = list(m.classifiers)
    films = [f.mlp[1] for f in m.films]
    # default init keeps |gamma|,|beta| tiny; widen so FiLM actually matters in the fixtures
    with torch.no_grad():
        for f in films:
            f.weight.mul_(3.0)
            f.bias.add_(torch.randn_like(f.bias) * 0.5 + 1.0)
    return m, heads, films


def run_case(name, kind, tree_name, tree, B, h, w, seed, drop_parent_in_sample=None, blobs=False):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    m, heads, films = build_model(kind, tree)
    m.train()
    C = 64 if kind == "unet" else 720
    scale = 1 if kind == "unet" else 4
    H, W = h * scale, w * scale
    nL = len(m.levels)
    feats = [torch.randn(B, C, h, w, generator=gen).mul_(0.7).requires_grad_(True) for _ in range(nL)]
    targets = O.synth_targets(m.levels, m.child_groups, B, H, W, gen, blobs=blobs)
    if drop_parent_in_sample is not None:
        # one sample where the (first) parent class never occurs -> all-ignored child masks
        b = drop_parent_in_sample
        pi = m.levels[0].index(m.child_groups[0][0][0])
        t0 = targets[0][b]
        moved = t0[pi] == 1
        t0[pi][moved] = 0
        t0[(pi + 1) % t0.shape[0]][moved] = 1
        gen2 = torch.Generator().manual_seed(seed + 2)
        rest = O.synth_targets(m.levels, m.child_groups, B, H, W, gen2)
        # rebuild deeper levels consistently for that sample
        fixed = [targets[0]]
        for L in range(1, nL):
            t = rest[L].clone()
            start = 0
            for pname, kids in m.child_groups[L - 1]:
                g = len(kids)
                pidx = m.levels[L - 1].index(pname)
                inside = (fixed[L - 1][:, pidx] == 1).unsqueeze(1)
                lab = torch.randint(0, g, (B, H, W), generator=gen2)
                oh = torch.nn.functional.one_hot(lab, g).permute(0, 3, 1, 2).float()
                t[:, start:start + g] = torch.where(inside, oh, torch.full_like(oh, -1.0))
                start += g
            fixed.append(t)
        targets = fixed
    lw = weights_for(tree_name, m.levels)

    it = iter(feats)
    if kind == "unet":
        m._run_unet = lambda x: next(it)
        x = torch.zeros(B, 3, H, W)
        probs, logits = m(x, type=1, hierarchy=tree)
    else:
        m._forward_backbone = lambda x: next(it)
        x = torch.zeros(B, 3, H, W)
        probs, logits = m(x)

    # train.py:206-231
    out_class = []
    for L in range(nL):
        idx = torch.argmax(torch.nn.functional.softmax(logits[L], dim=1), dim=1)
        out_class.append(torch.nn.functional.one_hot(idx, num_classes=len(m.levels[L])).permute(0, 3, 1, 2).float())
    eval_targets = list(targets)
    for L in range(nL):
        out_class[L] = torch.where(targets[L] == -1, 0, out_class[L])
        eval_targets[L] = torch.where(targets[L] == -1, 0, targets[L])

    loss_fns = [[losses.CrossEntropyLoss(), losses.SoftDiceLoss(num_classes=len(lv))] for lv in m.levels]
    # train.py:239 passes lambda_cons/lambda_kl which get_loss does not accept
    total, _, level_loss = train.get_loss(logits, targets, loss_fns, [], lw, 0.0, [],
                                          probs_per_level=out_class, model=m)
    total.backward()

    rec = {"kind": kind, "tree": json.dumps(tree), "scale": scale, "B": B, "h": h, "w": w,
           "level_weights": json.dumps(lw), "total_loss": total.item(),
           "level_loss": np.array(level_loss, dtype=np.float64)}
    with torch.no_grad():
        rec["consistency_train"] = losses.hierarchical_consistency_loss(out_class, m.levels, m.parent_of).item()
        rec["consistency_eval"] = losses.hierarchical_consistency_loss(probs, m.levels, m.parent_of).item()
    for L in range(nL):
        ce = loss_fns[L][0](logits[L], targets[L], class_weight=lw[L], logits_input=True)
        di = loss_fns[L][1](logits[L], targets[L], class_weight=lw[L], logits_input=True)
        rec[f"ce{L}"] = ce.item()
        rec[f"dice{L}"] = float("nan") if di is None else di.item()
        rec[f"feats{L}"] = feats[L].detach().numpy()
        rec[f"dfeats{L}"] = feats[L].grad.numpy()
        rec[f"target{L}"] = targets[L].numpy().astype(np.int8)
        rec[f"probs{L}"] = probs[L].detach().numpy()
        rec[f"logits{L}"] = logits[L].detach().numpy()
        rec[f"onehot{L}"] = out_class[L].numpy().astype(np.int8)
        rec[f"head_w{L}"] = heads[L].weight.detach().numpy()
        rec[f"head_b{L}"] = heads[L].bias.detach().numpy()
        rec[f"dhead_w{L}"] = heads[L].weight.grad.numpy()
        rec[f"dhead_b{L}"] = heads[L].bias.grad.numpy()
        if L >= 1:
            rec[f"film_w{L-1}"] = films[L - 1].weight.detach().numpy()
            rec[f"film_b{L-1}"] = films[L - 1].bias.detach().numpy()
            rec[f"dfilm_w{L-1}"] = films[L - 1].weight.grad.numpy()
            rec[f"dfilm_b{L-1}"] = films[L - 1].bias.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "total", total.item(), "levels", level_loss)


def run_flat(name, B, H, W, seed):
    """Config 4: flat 7-class weighted Dice+CE on logits; targets are {0,1} one-hots (no -1)."""
    torch.manual_seed(seed)
    z = (torch.randn(B, 7, H, W) * 2).requires_grad_(True)
    lab = torch.randint(0, 7, (B, H, W))
    t = torch.nn.functional.one_hot(lab, 7).permute(0, 3, 1, 2).float()
    ce = losses.CrossEntropyLoss()(z, t, class_weight=W_FLAT, logits_input=True)
    di = losses.SoftDiceLoss(num_classes=7)(z, t, class_weight=W_FLAT, logits_input=True)
    (ce + di).backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), logits=z.detach().numpy(), target=t.numpy().astype(np.int8),
                        weights=np.array(W_FLAT), ce=ce.item(), dice=di.item(), dlogits=z.grad.numpy())
    print(name, ce.item(), di.item())


def survey_anchor():
    """SURVEY.md 8(c) anchor: seed-0 recipe at 620x620 whose total loss is 2.813019."""
    torch.manual_seed(0)
    m = models.UNet(size=620, n_channels=3, hierarchy=TL, model_type=1)
    feats = [torch.randn(4, 64, 620, 620) for _ in range(2)]
    l0 = torch.randint(0, 4, (4, 620, 620))
    l1 = torch.randint(0, 4, (4, 620, 620))
    t0 = torch.nn.functional.one_hot(l0, 4).permute(0, 3, 1, 2).float()
    t1 = torch.where((l0 == 3).unsqueeze(1), torch.nn.functional.one_hot(l1, 4).permute(0, 3, 1, 2).float(),
                     torch.tensor(-1.0))
    it = iter(feats)
    m._run_unet = lambda x: next(it)
    with torch.no_grad():
        _, logits = m(torch.zeros(4, 3, 620, 620), type=1, hierarchy=TL)
        oc = []
        for L, t in enumerate((t0, t1)):
            idx = torch.argmax(torch.softmax(logits[L], 1), 1)
            oc.append(torch.where(t == -1, 0, torch.nn.functional.one_hot(idx, 4).permute(0, 3, 1, 2).float()))
        fns = [[losses.CrossEntropyLoss(), losses.SoftDiceLoss(num_classes=4)] for _ in range(2)]
        total, _, _ = train.get_loss(logits, [t0, t1], fns, [], W_TL, 0.0, [], probs_per_level=oc, model=m)
    print("survey anchor total loss:", total.item())
    return total.item()


if __name__ == "__main__":
    run_case("unet_tl", "unet", "tl", TL, B=3, h=16, w=20, seed=11)
    run_case("unet_tl_odd_notooth", "unet", "tl", TL, B=3, h=13, w=11, seed=12, drop_parent_in_sample=1)
    run_case("unet_ext", "unet", "ext", EXT, B=2, h=12, w=16, seed=13, blobs=True)
    run_case("unet_adv", "unet", "adv", ADV, B=2, h=9, w=12, seed=14)
    run_case("hrnet_tl", "hrnet", "tl", TL, B=2, h=6, w=5, seed=15)
    run_case("hrnet_ext", "hrnet", "ext", EXT, B=2, h=5, w=7, seed=16, drop_parent_in_sample=0)
    run_flat("flat7", B=2, H=10, W=14, seed=17)
    if "--anchor" in sys.argv:
        with open(os.path.join(HERE, "survey_anchor.json"), "w") as f:
            json.dump({"total_loss": survey_anchor(), "recipe": "SURVEY.md 8(c)"}, f)

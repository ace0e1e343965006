"""Synthetic inputs of the hot path: shared by bench.py (both arms), the tests and the oracle.
Neutral module: it imports neither the product package nor the oracle.

Target layout = what the reference's dataset produces (Data/dataset.py:227-265, process_ignore_values; channel order =
breadth-first over the class tree): roots in {0,1}; non-roots 1 on the class, 0 inside the direct parent, -1 outside it."""
import torch
import torch.nn.functional as F


def tree_levels_groups(tree: dict):
    """(levels, groups): class names per depth in breadth-first order, and per level L >= 1 the list of
    (parent name, [child names]) for the parents of level L-1 that have children (Models/models.py:82-98, :229-238)."""
    levels, groups = [], []
    frontier = [(name, sub) for name, sub in tree.items()]
    while frontier:
        levels.append([n for n, _ in frontier])
        grp = [(n, list(sub.keys())) for n, sub in frontier if isinstance(sub, dict) and sub]
        nxt = [(c, sub[c]) for n, sub in frontier if isinstance(sub, dict) for c in sub]
        if not nxt:
            break
        groups.append(grp)
        frontier = nxt
    return levels, groups


def synth_targets(levels, groups, B, H, W, gen: torch.Generator, blobs: bool = False, device="cpu"):
    """Ternary {1,0,-1} fp32 targets per level.  The random draws happen on gen's device (CPU generator -> CPU draws, moved
    to `device` at the end; a CUDA generator draws on the GPU)."""
    gdev = gen.device
    K0 = len(levels[0])
    if blobs:
        coarse = torch.randint(0, K0, (B, 1, max(H // 4, 1), max(W // 4, 1)), generator=gen, device=gdev).float()
        lab = F.interpolate(coarse, size=(H, W), mode="nearest").long().squeeze(1)
    else:
        lab = torch.randint(0, K0, (B, H, W), generator=gen, device=gdev)
    out = [F.one_hot(lab, K0).permute(0, 3, 1, 2).float()]
    for L in range(1, len(levels)):
        t = torch.full((B, len(levels[L]), H, W), -1.0, device=gdev)
        start = 0
        for pname, kids in groups[L - 1]:
            g = len(kids)
            pi = levels[L - 1].index(pname)
            inside = out[L - 1][:, pi] == 1
            lab = torch.randint(0, g, (B, H, W), generator=gen, device=gdev)
            oh = F.one_hot(lab, g).permute(0, 3, 1, 2).float()
            t[:, start:start + g] = torch.where(inside.unsqueeze(1), oh, torch.full_like(oh, -1.0))
            start += g
        out.append(t)
    return [o.to(device) for o in out]


def drop_class_in_sample(targets, levels, groups, b, pname):
    """In sample b, class `pname` of level 0 never occurs (its pixels go to the next class) and everything below it is
    ignored (-1): the 'image without a tooth' case that makes the sample's Dice NaN (Metrics/losses.py:64-66)."""
    pi = levels[0].index(pname)
    t0 = targets[0]
    moved = t0[b, pi] == 1
    t0[b, pi][moved] = 0
    t0[b, (pi + 1) % t0.shape[1]][moved] = 1
    live = {pname}
    for L in range(1, len(levels)):
        start, nxt = 0, set()
        for parent, kids in groups[L - 1]:
            if parent in live:
                targets[L][b, start:start + len(kids)] = -1.0
                nxt.update(kids)
            start += len(kids)
        live = nxt
    return targets

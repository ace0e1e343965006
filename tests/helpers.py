"""Shared test helpers: golden-fixture loading and tolerance checks."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HIER_CASES = ["unet_tl", "unet_tl_odd_notooth", "unet_ext", "unet_adv", "hrnet_tl", "hrnet_ext"]

# north_star tolerance: probabilities / loss / gradients within 1e-5 relative (fp32).  Pure
# relative error is ill-posed for entries that are ~0 next to O(1) neighbours (SURVEY.md 7
# "hard parts"), so the check is |a-b| <= RTOL*|ref| + RTOL*max|ref|.
RTOL = 1e-5


def close(a, ref, rtol=RTOL, what=""):
    a = torch.as_tensor(a).detach().double().cpu()
    ref = torch.as_tensor(ref).detach().double().cpu()
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    if ref.numel() == 0:
        return
    scale = ref.abs().max().item()
    err = (a - ref).abs()
    bound = rtol * ref.abs() + rtol * scale + 1e-12
    bad = err > bound
    assert not bad.any(), "%s: %d/%d out of tolerance, max err %.3e (scale %.3e)" % (
        what, int(bad.sum()), ref.numel(), err.max().item(), scale)


class Fixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.z = z
        self.name = name
        self.kind = str(z["kind"])
        self.tree = json.loads(str(z["tree"]))
        self.scale = int(z["scale"])
        self.B, self.h, self.w = int(z["B"]), int(z["h"]), int(z["w"])
        self.level_weights = json.loads(str(z["level_weights"]))
        self.nL = len(self.level_weights)
        self.out_size = None if self.scale == 1 else (self.h * self.scale, self.w * self.scale)

    def t(self, key, device="cpu", dtype=torch.float32):
        return torch.from_numpy(np.asarray(self.z[key])).to(dtype).to(device)

    def f(self, key):
        return float(self.z[key])

    def per_level(self, prefix, device="cpu", n=None, start=0, dtype=torch.float32):
        n = self.nL if n is None else n
        return [self.t(f"{prefix}{i}", device, dtype) for i in range(start, n)]

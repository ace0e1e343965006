"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

The reference imports several third-party packages that are absent offline
(timm, segmentation_models_pytorch, torchmetrics, yacs, matplotlib, skimage).
None of them is on the head/loss code path, so they are replaced with empty
stand-in modules.  torchmetrics IS live on the metric path; the metric fixtures
therefore come from an independent restatement cross-checked with sklearn
(see make_golden.py), and metric parity is declared "unpinned" in DESIGN.md.

Only tests/golden/make_golden.py uses this file; nothing on the GPU box does.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RHSEG_REFERENCE_ROOT", "/root/reference")


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        sys.modules[sub.__name__] = sub
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def _stub(name):
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        full = ".".join(parts[:i])
        if full not in sys.modules:
            sys.modules[full] = _Anything(full)
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[full])


class CfgNode(dict):
    """Dict-backed stand-in for yacs.config.CfgNode (attribute access + yaml merge)."""

    def __init__(self, init=None, new_allowed=False):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def defrost(self):
        pass

    def freeze(self):
        pass

    def merge_from_list(self, opts):
        pass

    def merge_from_file(self, path):
        import yaml
        def merge(dst, src):
            for k, v in src.items():
                if isinstance(v, dict):
                    if not isinstance(dst.get(k), CfgNode):
                        dst[k] = CfgNode()
                    merge(dst[k], v)
                else:
                    dst[k] = v
        with open(path) as f:
            merge(self, yaml.safe_load(f) or {})


def hrnet_config():
    """The reference's HRNet-W48 config (config/default.py + the shipped yaml)."""
    cfgmod = importlib.import_module("config")
    cfg = cfgmod.config
    ymls = [f for f in os.listdir(os.path.join(REFERENCE_ROOT, "config")) if f.endswith(".yaml")]
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, "config", ymls[0]))
    return cfg


def load_reference():
    """Returns (models, losses, train) modules of the reference."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("timm", "timm.models", "timm.models.vision_transformer",
                 "segmentation_models_pytorch", "torchmetrics", "yacs", "yacs.config",
                 "matplotlib", "matplotlib.pyplot", "skimage", "skimage.io",
                 "skimage.transform", "skimage.color", "skimage.morphology"):
        _stub(name)
    sys.modules["yacs.config"].CfgNode = CfgNode
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own drop-in package must not shadow the reference here
    for k in [k for k in sys.modules if k.split(".")[0] in ("Models", "Metrics", "tree_util", "train", "config", "Data")]:
        del sys.modules[k]
    models = importlib.import_module("Models.models")
    losses = importlib.import_module("Metrics.losses")
    try:
        train = importlib.import_module("train")
    except Exception as e:  # train.py pulls Data/ + config/ (yacs); fall back to None
        train = None
        sys.stderr.write("reference train.py not importable: %r\n" % (e,))
    return models, losses, train

"""ORACLE — test infrastructure only, never product code.

Imports the UNMODIFIED reference (Banksylel/Restrictive-Hierarchical-Semantic-Segmentation) for parity tests and for
the CPU baseline of bench.py.  The tree is looked up at, in this order: $RHSEG_REFERENCE_ROOT, /root/reference (the
build container), oracle/_ref (byte copies staged by tools/stage_reference.py; git-ignored, travels to the GPU box).

The reference imports several third-party packages that are absent offline (timm, segmentation_models_pytorch,
torchmetrics, yacs, matplotlib, skimage).  None of them is on the head / loss path, so they are replaced with empty
stand-in modules.  torchmetrics IS live on the metric path (Metrics/performance_metrics.py:62 ...): metric objects
handed to train.get_metrics by the tests are therefore the restated ones of oracle/hier_oracle.py (parity of that
slice stays "unpinned"), while ProcessClasses itself (torch only) is used as is.

    models, losses, train = load_reference()             # the reference's own Models / Metrics / train
    train = load_train_with_dropin()                      # the reference's train.py on top of OUR Models / Metrics
"""
import importlib
import os
import sys
import types

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(REPO_ROOT, "restrictive-hierarchical-semantic-segmentation_b200")
STAGED = os.path.join(REPO_ROOT, "oracle", "_ref")
_REF_MODULE_ROOTS = ("Models", "Metrics", "tree_util", "train", "predictEval", "config", "Data")


def reference_root():
    env = os.environ.get("RHSEG_REFERENCE_ROOT")
    for cand in (env, "/root/reference", STAGED):
        if cand and os.path.isfile(os.path.join(cand, "train.py")):
            return cand
    return None


def available():
    return reference_root() is not None


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


def install_torchmetrics_standin():
    """torchmetrics is live on the metric path (Metrics/performance_metrics.py:62 ... :140) but not installable offline.
    This stand-in implements exactly the five constructors the reference calls -- F1Score / JaccardIndex / Accuracy /
    Precision / Recall(task='multiclass', num_classes=n, average=None, ignore_index=i) -- over the restated multiclass
    semantics of oracle/hier_oracle.py (parity of that slice stays 'unpinned'), so that the reference's OWN
    ProcessClasses and wrapper classes run unchanged around it (five argmax + confusion passes per level, as upstream)."""
    try:
        import torchmetrics  # noqa: F401  (a real installation wins)
        if not isinstance(sys.modules["torchmetrics"], _Anything):
            return
    except Exception:
        pass
    from oracle import hier_oracle as O
    mod = types.ModuleType("torchmetrics")

    def make(key):
        class _Multiclass:
            def __init__(self, task="multiclass", num_classes=None, average=None, ignore_index=None, **kw):
                if task != "multiclass" or average is not None:
                    raise NotImplementedError("stand-in covers the reference's calls only: task='multiclass', average=None")
                self.num_classes, self.ignore_index = int(num_classes), ignore_index

            def to(self, device):
                return self

            def __call__(self, preds, target):
                conf = O.multiclass_confusion(preds, target, self.num_classes, self.ignore_index)
                return O.ratios_from_confusion(conf)[key]
        return _Multiclass

    mod.F1Score, mod.JaccardIndex, mod.Accuracy = make("dice"), make("iou"), make("accuracy")
    mod.Precision, mod.Recall = make("precision"), make("recall")
    mod.__standin__ = True
    sys.modules["torchmetrics"] = mod


def _prepare(first_paths):
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (looked at $RHSEG_REFERENCE_ROOT, /root/reference, oracle/_ref); "
                           "run tools/stage_reference.py in the build container")
    for name in ("timm", "timm.models", "timm.models.vision_transformer", "segmentation_models_pytorch", "torchmetrics",
                 "yacs", "yacs.config", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.io", "skimage.transform",
                 "skimage.color", "skimage.morphology", "cv2"):
        _stub(name)
    sys.modules["yacs.config"].CfgNode = CfgNode
    install_torchmetrics_standin()
    for p in list(first_paths) + [root]:
        while p in sys.path:
            sys.path.remove(p)
    for p in reversed(list(first_paths) + [root]):
        sys.path.insert(0, p)
    for k in [k for k in sys.modules if k.split(".")[0] in _REF_MODULE_ROOTS]:
        del sys.modules[k]
    return root


def hrnet_config():
    """The reference's HRNet-W48 config (config/default.py + the shipped yaml)."""
    root = reference_root()
    cfg = importlib.import_module("config").config
    ymls = sorted(f for f in os.listdir(os.path.join(root, "config")) if f.endswith(".yaml"))
    cfg.merge_from_file(os.path.join(root, "config", ymls[0]))
    return cfg


def load_reference():
    """(Models.models, Metrics.losses, train) of the reference itself."""
    _prepare([])
    models = importlib.import_module("Models.models")
    losses = importlib.import_module("Metrics.losses")
    try:
        train = importlib.import_module("train")
    except Exception as e:  # train.py pulls Data/ + config/
        train = None
        sys.stderr.write("reference train.py not importable: %r\n" % (e,))
    return models, losses, train


def load_reference_module(name):
    """Any other module of the reference (after load_reference()): 'Metrics.performance_metrics', 'predictEval' ..."""
    return importlib.import_module(name)


def load_train_with_dropin():
    """The reference's train.py importing OUR Models / Metrics / tree_util: the package directory precedes the
    reference root on sys.path, exactly the drop-in mechanism INTEGRATION.md describes.  Returns the train module."""
    _prepare([PKG_DIR])
    train = importlib.import_module("train")
    mods = sys.modules["Models.models"].__file__, sys.modules["Metrics.losses"].__file__
    for m in mods:
        if not os.path.abspath(m).startswith(PKG_DIR):
            raise RuntimeError("train.py did not pick up the drop-in modules: %s" % (m,))
    return train


def restore_package_imports():
    """Forget the Models / Metrics / train modules imported above (tests that switch between the two flavours)."""
    for k in [k for k in sys.modules if k.split(".")[0] in _REF_MODULE_ROOTS]:
        del sys.modules[k]

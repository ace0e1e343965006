"""Stages the files of the reference that the hot path's parity tests and the CPU baseline import into the git-ignored
directory oracle/_ref/ (it travels to the GPU box with the repository snapshot, like a built .so, and never enters
history).  Nothing is edited: the files are byte copies, imported there with the stub modules of oracle/ref_loader.py.

    python tools/stage_reference.py [--reference /root/reference]

What is staged and why (paths relative to the reference root):
    Models/           the hierarchical wrappers + donor backbones (the parity target of Models/models.py)
    Metrics/          losses.py, performance_metrics.py (ProcessClasses is torch-only; the torchmetrics calls are stubbed)
    train.py          get_loss / get_metrics / train_epoch: the glue "drops into train.py unchanged" is measured against
    predictEval.py    get_parent_masks / combine_levels (flat -> hierarchy stitching)
    tree_util.py, config/, Data/            imported by train.py at module load
    class_tree_*.json, class_map*.csv       the two class trees
"""
import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "oracle", "_ref")
ITEMS = ["Models", "Metrics", "config", "Data", "train.py", "predictEval.py", "tree_util.py",
         "class_tree_tl.json", "class_tree_tl_extended.json", "class_map.csv", "class_map_extended.csv"]


def stage(reference="/root/reference", quiet=False):
    if not os.path.isdir(reference):
        raise SystemExit("reference tree not found at %s" % reference)
    os.makedirs(DEST, exist_ok=True)
    n = 0
    for item in ITEMS:
        src, dst = os.path.join(reference, item), os.path.join(DEST, item)
        if os.path.isdir(src):
            for dirpath, dirnames, filenames in os.walk(src):
                dirnames[:] = [d for d in dirnames if d != "__pycache__"]
                rel = os.path.relpath(dirpath, src)
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                for f in filenames:
                    if f.endswith(".pyc"):
                        continue
                    a, b = os.path.join(dirpath, f), os.path.join(dst, rel, f)
                    if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                        shutil.copyfile(a, b)
                        n += 1
        elif os.path.exists(src):
            if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
                n += 1
    if not quiet:
        print("staged %s -> %s (%d file(s) updated)" % (reference, DEST, n))
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    stage(ap.parse_args().reference)

"""Class tree (nested dict, class_tree_tl.json style) -> per-level integer tables.

Reference counterparts: get_level_classes / build_hierarchy_indices
(Models/models.py:82-98, :38-54) and the child_groups construction (:229-238, :636-645).
The name-keyed views (`levels`, `parent_of`, `children_of`, `child_groups`) are kept because
train.py:146-149 reads them off the model; the kernels only ever see the int32 tables."""
from typing import Dict, List, Optional, Tuple

import torch

from . import native


def get_level_classes(hierarchy, depth=0, result=None, inc_parent=False):
    """Names per depth in pre-order; inc_parent=False keeps leaves only (flat models)."""
    if result is None:
        result = {}
    if not isinstance(hierarchy, dict) or not hierarchy:
        return result
    for name in hierarchy:
        sub = hierarchy[name]
        bucket = result.setdefault(depth, [])
        if inc_parent or not sub:
            bucket.append(name)
        if isinstance(sub, dict):
            get_level_classes(sub, depth + 1, result, inc_parent)
    return result


def build_hierarchy_indices(hierarchy):
    """(levels, parent_of, children_of) with the reference's ordering conventions."""
    per_depth = get_level_classes(hierarchy, inc_parent=True)
    levels = [per_depth[d] for d in sorted(per_depth)]
    parent_of: Dict[str, Optional[str]] = {}
    children_of: Dict[str, List[str]] = {}
    stack = [(hierarchy, None)]
    while stack:
        node, parent = stack.pop()
        pending = []
        for name, sub in node.items():
            parent_of[name] = parent
            if isinstance(sub, dict) and len(sub) > 0:
                children_of[name] = list(sub.keys())
                pending.append((sub, name))
            else:
                children_of.setdefault(name, [])
        stack.extend(reversed(pending))
    return levels, parent_of, children_of


class ClassTree:
    """Compiled class tree: name views + per-level int32 tables (host lists, device tensors
    on demand).  One instance is shared by the head, the consistency loss and the metrics."""

    def __init__(self, hierarchy: dict):
        self.hierarchy = hierarchy
        self.levels, self.parent_of, self.children_of = build_hierarchy_indices(hierarchy)
        self.child_groups: List[List[Tuple[str, List[str]]]] = []
        for L in range(1, len(self.levels)):
            self.child_groups.append([(p, self.children_of[p]) for p in self.levels[L - 1]
                                      if len(self.children_of.get(p, [])) > 0])
        self.num_levels = len(self.levels)
        # channels per level as the HEAD sees them (a level without groups still has 1 channel)
        self.head_channels = [len(self.levels[0])] + [max(1, sum(len(c) for _, c in g)) for g in self.child_groups]
        self.parent_channel: List[List[int]] = [[-1] * len(self.levels[0])]
        self.act_mode = [native.ACT_SIGMOID]
        for L in range(1, self.num_levels):
            groups = self.child_groups[L - 1]
            if not groups:
                self.parent_channel.append([0])
                self.act_mode.append(native.ACT_ZEROS)
                continue
            pc = []
            for pname, kids in groups:
                pc += [self.levels[L - 1].index(pname)] * len(kids)
            self.parent_channel.append(pc)
            self.act_mode.append(native.ACT_GROUPED)
        self._host_tables = None
        self._device_tables = {}

    @property
    def host_tables(self):
        if self._host_tables is None:
            tabs = []
            for L in range(self.num_levels):
                k_prev = 0 if L == 0 else self.head_channels[L - 1]
                tabs.append(native.compile_level_table(self.parent_channel[L], k_prev))
            self._host_tables = tabs
        return self._host_tables

    def device_tables(self, device):
        """int32 [num_levels, TABLE_INTS] on `device`, uploaded once per device."""
        key = str(device)
        if key not in self._device_tables:
            self._device_tables[key] = torch.tensor(self.host_tables, dtype=torch.int32, device=device)
        return self._device_tables[key]

    def group_count(self, L):
        return 0 if L == 0 else len(self.child_groups[L - 1])


_TREE_CACHE = {}


def compiled_tree(hierarchy: dict) -> ClassTree:
    """Memoised ClassTree for a hierarchy dict (keyed by its structure)."""
    import json
    key = json.dumps(hierarchy)
    if key not in _TREE_CACHE:
        _TREE_CACHE[key] = ClassTree(hierarchy)
    return _TREE_CACHE[key]


def tree_from_levels(levels, parent_of) -> ClassTree:
    """Rebuilds a ClassTree from the (levels, parent_of) pair the consistency loss receives."""
    def sub(name):
        kids = [c for lv in levels for c in lv if parent_of.get(c, None) == name]
        return {k: sub(k) for k in kids}
    return compiled_tree({r: sub(r) for r in levels[0]})

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

    def group_hint(self, L):
        """RHSEG_GROUP_HINT bits of level L: the common size of its parent groups (the reference's trees: one group of
        4 / 2 / 3 channels, or two groups of 2), 0 when the sizes differ or the level has no group.  OR-ed into the
        act_mode / child arguments, it selects kernels with the group layout fixed at compile time."""
        if L == 0 or not self.child_groups[L - 1]:
            return 0
        sizes = {len(kids) for _, kids in self.child_groups[L - 1]}
        return (sizes.pop() << 8) if len(sizes) == 1 else 0

    # ---- flat -> hierarchy stitching tables (predictEval.py:36-83) ----
    def bfs_names(self):
        """Breadth-first node order (predictEval.bfs_order): the channel order of class_map.csv."""
        from collections import deque
        q, order = deque(self.hierarchy.items()), []
        while q:
            name, sub = q.popleft()
            order.append(name)
            if isinstance(sub, dict) and sub:
                q.extend(sub.items())
        return order

    def bfs_levels(self):
        """predictEval.levels_bfs: names per depth in breadth-first order."""
        from collections import deque
        q, out = deque((n, s, 0) for n, s in self.hierarchy.items()), []
        while q:
            name, sub, d = q.popleft()
            if len(out) <= d:
                out.append([])
            out[d].append(name)
            if isinstance(sub, dict) and sub:
                q.extend((cn, cs, d + 1) for cn, cs in sub.items())
        return out

    def leaf_order(self):
        """Flat-model channel order: leaves in breadth-first order (predictEval.py:381)."""
        return [n for n in self.bfs_names() if not self.children_of.get(n)]

    def stitch_masks(self):
        """(masks, per-level channel counts) for rhseg_stitch_levels: one uint32 per node in bfs_levels order."""
        leaf_idx = {n: i for i, n in enumerate(self.leaf_order())}

        def leaves_under(n):
            kids = self.children_of.get(n, [])
            return [n] if not kids else [l for k in kids for l in leaves_under(k)]

        masks, counts = [], []
        for names in self.bfs_levels():
            counts.append(len(names))
            for n in names:
                if not self.children_of.get(n):
                    masks.append((1 << 31) | (1 << leaf_idx[n]))
                else:
                    m = 0
                    for l in leaves_under(n):
                        m |= 1 << leaf_idx[l]
                    masks.append(m)
        return masks, counts


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

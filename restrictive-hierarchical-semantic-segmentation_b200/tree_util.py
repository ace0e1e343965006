"""Tab-indented text-tree helpers with the reference's names (tree_util.py).

The reference imports these in train.py:9 and Metrics/losses.py:4 but never calls them (its
trees are JSON dicts); they are provided so those imports resolve.  The live
tree code of this package is tree_tables.ClassTree.  All walks below are iterative (explicit
stacks) over one helper, `_leaves_in_order`."""


class node:
    __slots__ = ("name", "children", "channel", "level")

    def __init__(self, name):
        self.name, self.children, self.channel, self.level = name, [], None, None


def _is_leaf(n):
    return len(n.children) == 0


def _leaves_in_order(top):
    """Leaves below (or equal to) `top`, left to right."""
    found, stack = [], [top]
    while stack:
        cur = stack.pop()
        if _is_leaf(cur):
            found.append(cur)
        else:
            stack.extend(reversed(cur.children))
    return found


def create_tree_from_textfile(filename):
    """One class per line, depth = number of tab characters; depth may grow by one per line."""
    root = node("Universal class")
    trail = [root]  # trail[d] = latest node seen at depth d - 1 (trail[0] is the root)
    with open(filename, "r") as fh:
        for raw in fh:
            depth = raw.count("\t")
            if depth >= len(trail):
                raise RuntimeError("Indentation can only increase by one")
            born = node(raw.strip())
            trail[depth].children.append(born)
            trail[depth + 1:] = [born]
    return root


def add_channels(node, channel):
    """Numbers the leaves left to right starting at `channel`; returns the next free number."""
    for offset, leaf in enumerate(_leaves_in_order(node)):
        leaf.channel = channel + offset
    return channel + len(_leaves_in_order(node))


def update_channels(node, class_lookup):
    for leaf in _leaves_in_order(node):
        leaf.channel = class_lookup[leaf.channel]


def add_levels(node, depth):
    """Children of a node visited at `depth` get level depth - 1; a childless start node gets it itself."""
    if _is_leaf(node):
        node.level = depth - 1
        return
    todo = [(node, depth)]
    while todo:
        cur, d = todo.pop()
        for kid in cur.children:
            kid.level = d - 1
            if not _is_leaf(kid):
                todo.append((kid, d - 1))


def getLeafClasses(node, my_list):
    my_list.extend(leaf.channel for leaf in _leaves_in_order(node))
    return my_list


def find_depth(node):
    deepest, frontier = 0, [(node, 0)]
    while frontier:
        cur, d = frontier.pop()
        deepest = max(deepest, d)
        frontier.extend((kid, d + 1) for kid in cur.children)
    return deepest


def getLossLevelList(root, level, myList):
    """Appends, left to right, the leaf-channel list of every subtree cut at `level` (or at a leaf above it)."""
    stack = list(reversed(root.children))
    while stack:
        cur = stack.pop()
        if _is_leaf(cur) or cur.level == level:
            myList.append(getLeafClasses(cur, []))
        else:
            stack.extend(reversed(cur.children))


def getTreeList(node):
    rows = []
    for level in range(find_depth(node)):
        rows.append([])
        getLossLevelList(node, level, rows[-1])
    return rows

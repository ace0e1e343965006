"""Tab-indented text-tree helpers with the reference's names (tree_util.py).

The reference imports these in train.py:9 and Metrics/losses.py:4 but never calls them (its
trees are JSON dicts); they are provided so those imports resolve.  The live
tree code of this package is tree_tables.ClassTree."""


class node:
    def __init__(self, name):
        self.name = name
        self.children = []
        self.channel = None
        self.level = None


def create_tree_from_textfile(filename):
    """One class per line, depth = number of tab characters; depth may grow by one per line."""
    root = node("Universal class")
    path = [root]  # path[d] = most recent node at depth d-1 (path[0] is the root)
    with open(filename, "r") as fh:
        for line in fh:
            depth = line.count("\t")
            if depth > len(path) - 1:
                raise RuntimeError("Indentation can only increase by one")
            fresh = node(line.strip())
            del path[depth + 1:]
            path[depth].children.append(fresh)
            path.append(fresh)
    return root


def add_channels(node, channel):
    if not node.children:
        node.channel = channel
        return channel + 1
    for child in node.children:
        channel = add_channels(child, channel)
    return channel


def update_channels(node, class_lookup):
    if not node.children:
        node.channel = class_lookup[node.channel]
        return
    for child in node.children:
        update_channels(child, class_lookup)


def add_levels(node, depth):
    if not node.children:
        node.level = depth - 1
        return
    for child in node.children:
        child.level = depth - 1
        if child.children:
            add_levels(child, depth - 1)


def getLeafClasses(node, my_list):
    if not node.children:
        my_list.append(node.channel)
        return my_list
    for child in node.children:
        getLeafClasses(child, my_list)
    return my_list


def find_depth(node):
    return 0 if not node.children else 1 + max(find_depth(c) for c in node.children)


def getLossLevelList(root, level, myList):
    for child in root.children:
        if not child.children or child.level == level:
            myList.append(getLeafClasses(child, []))
        else:
            getLossLevelList(child, level, myList)


def getTreeList(node):
    out = []
    for level in range(find_depth(node)):
        row = []
        getLossLevelList(node, level, row)
        out.append(row)
    return out

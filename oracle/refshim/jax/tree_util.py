from ._core import tree_map, tree_leaves  # noqa: F401


def tree_flatten(tree):
    return tree_leaves(tree), None

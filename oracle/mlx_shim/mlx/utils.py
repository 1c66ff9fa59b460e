"""mlx.utils stand-in (only what the reference's loaders import at call time)."""


def tree_flatten(tree, prefix="", is_leaf=None, destination=None):
    """Flat (dotted name, leaf) pairs; with ``destination`` (a dict) the pairs are stored there and it is returned."""
    out = []
    if isinstance(tree, dict):
        for k, v in tree.items():
            out += tree_flatten(v, f"{prefix}.{k}" if prefix else str(k))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out += tree_flatten(v, f"{prefix}.{i}" if prefix else str(i))
    elif tree is not None:
        out.append((prefix, tree))
    if destination is not None:
        destination.update(out)
        return destination
    return out

"""mlx.utils stand-in (only what the reference's loaders import at call time)."""


def tree_flatten(tree, prefix=""):
    out = []
    if isinstance(tree, dict):
        for k, v in tree.items():
            out += tree_flatten(v, f"{prefix}.{k}" if prefix else str(k))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out += tree_flatten(v, f"{prefix}.{i}" if prefix else str(i))
    else:
        out.append((prefix, tree))
    return out

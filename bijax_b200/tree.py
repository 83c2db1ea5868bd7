"""Minimal pytree helpers with JAX's dict semantics (keys visited in sorted order)."""
from __future__ import annotations

from typing import Any, Callable, List, Tuple

import torch


def _is_leaf(x) -> bool:
    return not isinstance(x, (dict, list, tuple))


def tree_flatten(tree) -> Tuple[List[Any], Any]:
    leaves: List[Any] = []

    def rec(node):
        if isinstance(node, dict):
            return ("dict", [(k, rec(node[k])) for k in sorted(node)])
        if isinstance(node, (list, tuple)):
            return (type(node).__name__, [rec(v) for v in node])
        leaves.append(node)
        return ("leaf", None)

    treedef = rec(tree)
    return leaves, treedef


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(td):
        kind, payload = td
        if kind == "dict":
            return {k: rec(v) for k, v in payload}
        if kind == "list":
            return [rec(v) for v in payload]
        if kind == "tuple":
            return tuple(rec(v) for v in payload)
        return next(it)

    return rec(treedef)


def tree_map(fn: Callable, tree, *rest):
    leaves, td = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(td, [fn(*xs) for xs in zip(leaves, *others)])


def tree_leaves(tree):
    return tree_flatten(tree)[0]


def ravel_pytree(tree):
    """jax.flatten_util.ravel_pytree: (flat vector, unravel_fn)."""
    leaves, td = tree_flatten(tree)
    shapes = [tuple(l.shape) for l in leaves]
    sizes = [int(l.numel()) for l in leaves]
    flat = torch.cat([l.reshape(-1) for l in leaves]) if leaves else torch.zeros(0)

    def unravel(vec):
        out, o = [], 0
        for shp, n in zip(shapes, sizes):
            out.append(vec[o : o + n].reshape(shp))
            o += n
        return tree_unflatten(td, out)

    return flat, unravel

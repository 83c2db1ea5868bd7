"""Turn bijax's model description (a pytree of prior distributions + a pytree of bijectors) into the flat per-coordinate
table the CUDA kernels read.

Layout contract (core.py:168-176): the flat latent vector z is the concatenation, in sorted-key order (JAX flattens dicts
by sorted key), of each prior leaf's sample raveled row-major; ``unravel`` is its inverse (advi.py:87).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, List, Tuple

import numpy as np

from . import _native as nv
from . import bijectors as tfb

COORD_DTYPE = np.dtype([("prior", "<u4"), ("chain", "<u4"), ("p0", "<f4"), ("p1", "<f4"), ("lognorm", "<f4")])
assert COORD_DTYPE.itemsize == 20

_BIJ_OPS = {"Identity": nv.BIJ_IDENTITY, "Exp": nv.BIJ_EXP, "Softplus": nv.BIJ_SOFTPLUS, "Sigmoid": nv.BIJ_SIGMOID}


def _np(v) -> np.ndarray:
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=np.float64)


def _param(dist, *names):
    for n in names:
        v = getattr(dist, n, None)
        if v is not None and not callable(v):
            return v
        params = getattr(dist, "parameters", None)
        if isinstance(params, dict) and params.get(n) is not None:
            return params[n]
    raise TypeError(f"{type(dist).__name__} has none of the attributes {names}")


def is_distribution(x) -> bool:
    return type(x).__name__ in ("Normal", "MultivariateNormalDiag", "Beta", "Gamma") or hasattr(x, "log_prob")


def bijector_ops(b) -> List[int]:
    """Ops in APPLICATION order (first applied first)."""
    name = type(b).__name__
    if name == "Chain":
        ops: List[int] = []
        for sub in reversed(list(b.bijectors)):
            ops.extend(bijector_ops(sub))
        return ops
    if name not in _BIJ_OPS:
        raise NotImplementedError(
            f"bijector {name} is outside the fused set (Identity, Exp, Softplus, Sigmoid and chains of them)"
        )
    op = _BIJ_OPS[name]
    return [] if op == nv.BIJ_IDENTITY else [op]


def pack_chain(ops: List[int]) -> int:
    if len(ops) > 8:
        raise NotImplementedError("bijector chains longer than 8 ops are not supported")
    packed = 0
    for i, op in enumerate(ops):
        packed |= (op & 0xF) << (4 * i)
    return packed


def _leaf_rows(dist) -> Tuple[Tuple[int, ...], int, np.ndarray, np.ndarray, np.ndarray]:
    """shape, prior id, p0, p1, lognorm (all broadcast to shape, float64)."""
    name = type(dist).__name__
    if name in ("Normal", "MultivariateNormalDiag"):
        loc = _param(dist, "loc")
        scale = _param(dist, "scale_diag", "scale") if name == "MultivariateNormalDiag" else _param(dist, "scale")
        loc, scale = np.broadcast_arrays(_np(loc), _np(scale))
        lognorm = -0.5 * math.log(2.0 * math.pi) - np.log(scale)
        return loc.shape, nv.PRIOR_NORMAL, loc, scale, lognorm
    if name == "Beta":
        a, b = np.broadcast_arrays(_np(_param(dist, "concentration1")), _np(_param(dist, "concentration0")))
        lg = np.vectorize(math.lgamma)
        lognorm = -(lg(a) + lg(b) - lg(a + b))
        return a.shape, nv.PRIOR_BETA, a, b, lognorm
    if name == "Gamma":
        a, r = np.broadcast_arrays(_np(_param(dist, "concentration")), _np(_param(dist, "rate")))
        lg = np.vectorize(math.lgamma)
        lognorm = a * np.log(r) - lg(a)
        return a.shape, nv.PRIOR_GAMMA, a, r, lognorm
    raise NotImplementedError(
        f"prior {name} is outside the fused set (Normal, MultivariateNormalDiag, Beta, Gamma); callable / hierarchical "
        "priors are out of scope (the reference cannot construct them either: core.py:117)"
    )


def flatten_tree(tree, is_leaf) -> List[Tuple[Tuple[str, ...], Any]]:
    """(path, leaf) pairs in JAX dict order: keys sorted at every level."""
    out: List[Tuple[Tuple[str, ...], Any]] = []

    def rec(node, path):
        if is_leaf(node):
            out.append((path, node))
        elif isinstance(node, dict):
            for k in sorted(node):
                rec(node[k], path + (k,))
        else:
            raise TypeError(f"unsupported pytree node {type(node).__name__} at {'/'.join(path)}")

    rec(tree, ())
    return out


def fill_in_bijector(bijector: Dict, prior: Dict) -> Dict:
    """core.py:107-111 -- default Identity for unlisted keys; mutates and returns the caller's dict like the reference."""
    for key in prior:
        if key not in bijector:
            bijector[key] = tfb.Identity() if is_distribution(prior[key]) else fill_in_bijector({}, prior[key])
        elif isinstance(prior[key], dict) and isinstance(bijector[key], dict):
            fill_in_bijector(bijector[key], prior[key])
    return bijector


def _get_path(tree, path):
    for k in path:
        tree = tree[k]
    return tree


@dataclass
class LatentSpec:
    paths: List[Tuple[str, ...]]
    shapes: List[Tuple[int, ...]]
    offsets: List[int]
    sizes: List[int]
    size: int
    coords: np.ndarray  # structured COORD_DTYPE [size]

    def offset_of(self, name) -> Tuple[int, int]:
        path = (name,) if isinstance(name, str) else tuple(name)
        for p, o, n in zip(self.paths, self.offsets, self.sizes):
            if p == path or "/".join(p) == "/".join(path):
                return o, n
        raise KeyError(f"no latent named {name!r}; have {['/'.join(p) for p in self.paths]}")

    def unravel(self, flat):
        """flat[..., D] -> nested dict of [..., *shape] views (advi.py:87 unravel_fn, batched)."""
        out: Dict[str, Any] = {}
        for p, o, n, shp in zip(self.paths, self.offsets, self.sizes, self.shapes):
            node = out
            for k in p[:-1]:
                node = node.setdefault(k, {})
            node[p[-1]] = flat[..., o : o + n].reshape(tuple(flat.shape[:-1]) + tuple(shp))
        return out


def compile_latents(prior: Dict, bijector: Dict) -> LatentSpec:
    leaves = flatten_tree(prior, is_distribution)
    paths, shapes, offsets, sizes, rows = [], [], [], [], []
    off = 0
    for path, dist in leaves:
        shape, pid, p0, p1, lognorm = _leaf_rows(dist)
        chain = pack_chain(bijector_ops(_get_path(bijector, path)))
        n = int(np.prod(shape)) if shape else 1
        block = np.zeros(n, dtype=COORD_DTYPE)
        block["prior"] = pid
        block["chain"] = chain
        block["p0"] = np.ravel(p0)
        block["p1"] = np.ravel(p1)
        block["lognorm"] = np.ravel(lognorm)
        rows.append(block)
        paths.append(path)
        shapes.append(tuple(int(s) for s in shape))
        offsets.append(off)
        sizes.append(n)
        off += n
    return LatentSpec(paths, shapes, offsets, sizes, off, np.concatenate(rows) if rows else np.zeros(0, COORD_DTYPE))

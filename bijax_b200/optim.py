"""optax-shaped optimisers for ``train_fn`` (utils.py:8,26-27: ``optimizer.init``, ``optimizer.update``,
``optax.apply_updates``).  optax itself is not installable here; any object with the same two methods works.
``adam`` runs its update on the device through ``bjx_adam_update`` (one fused kernel per leaf)."""
from __future__ import annotations

from typing import Any, NamedTuple

import torch

from . import _native as nv
from .tree import tree_flatten, tree_map, tree_unflatten


class GradientTransformation(NamedTuple):
    init: Any
    update: Any


def apply_updates(params, updates):
    return tree_map(lambda p, u: p + u, params, updates)


def sgd(learning_rate: float) -> GradientTransformation:
    def init(params):
        return ()

    def update(grads, state, params=None):
        return tree_map(lambda g: -learning_rate * g, grads), state

    return GradientTransformation(init, update)


class AdamTransformation(NamedTuple):
    """GradientTransformation that also names its hyper-parameters, so that train_fn can run the whole loop on the device."""

    init: Any
    update: Any
    hyper: Any  # (learning_rate, b1, b2, eps)


class AdamState(NamedTuple):
    count: int
    mu: Any
    nu: Any


def adam(learning_rate, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> GradientTransformation:
    """optax.adam (eps outside the square root, bias correction on both moments).  ``learning_rate`` may be an
    optax-style schedule ``count -> lr`` (evaluated on the host at the 0-based step count, like optax's
    ``scale_by_learning_rate``); a schedule keeps train_fn on its host-driven loop."""
    schedule = learning_rate if callable(learning_rate) else None

    def init(params):
        return AdamState(0, tree_map(torch.zeros_like, params), tree_map(torch.zeros_like, params))

    def update(grads, state, params=None):
        count = state.count + 1
        lr = float(schedule(state.count)) if schedule is not None else float(learning_rate)
        g_leaves, td = tree_flatten(grads)
        m_leaves, _ = tree_flatten(state.mu)
        v_leaves, _ = tree_flatten(state.nu)
        lib = nv.load()
        ups = []
        for g, m, v in zip(g_leaves, m_leaves, v_leaves):
            if not (g.is_cuda and g.dtype == torch.float32):
                raise TypeError("optim.adam updates CUDA float32 leaves on the device (bjx_adam_update); there is no CPU path")
            g = g.contiguous()
            # the kernel applies p -= step in place; run it on a zero buffer to obtain the additive update
            u = torch.zeros_like(g)
            with torch.cuda.device(g.device):
                nv.check(lib.bjx_adam_update(u.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), g.numel(), count,
                                             lr, b1, b2, eps, nv.current_stream_ptr()))
            ups.append(u)
        return tree_unflatten(td, ups), AdamState(count, state.mu, state.nu)

    hyper = None if schedule is not None else (float(learning_rate), float(b1), float(b2), float(eps))
    return AdamTransformation(init, update, hyper)

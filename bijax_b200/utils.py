"""bijax/utils.py for the B200 engine: ``train`` == ``train_fn`` (README.md:119 names the former, utils.py:6 defines the
latter), ``initialize_params``, ``seeds_like``.
"""
from __future__ import annotations

from typing import Callable, Set

import torch

from . import optim as _optim
from . import random as brandom
from .tree import ravel_pytree, tree_flatten, tree_map, tree_unflatten


def value_and_grad(loss_fn: Callable) -> Callable:
    """jax.value_and_grad for loss functions of a params pytree whose leaves are torch tensors (utils.py:7)."""

    def fn(params, *args, **kwargs):
        leaves, td = tree_flatten(params)
        leaves = [l.detach().requires_grad_(True) for l in leaves]
        loss = loss_fn(tree_unflatten(td, leaves), *args, **kwargs)
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(l) for g, l in zip(grads, leaves)]
        return loss.detach(), tree_unflatten(td, grads)

    return fn


def train_fn(loss_fn, params, optimizer, n_epochs, seed=None, return_args: Set[str] = frozenset()):
    """utils.py:6-36.  ``optimizer`` is optax-shaped (init / update).  Per-step keys are split(seed, n_epochs)[i] exactly
    as the reference's lax.scan consumes them.  Losses stay on the device; nothing synchronises inside the loop."""
    value_and_grad_fn = value_and_grad(loss_fn)
    state = optimizer.init(params)
    seeds = brandom.split(seed, n_epochs) if seed is not None else None
    losses = []
    params_list = [] if "params_list" in return_args else None
    for i in range(int(n_epochs)):
        if seeds is None:
            loss, grads = value_and_grad_fn(params)
        else:
            loss, grads = value_and_grad_fn(params, seed=seeds[i])
        updates, state = optimizer.update(grads, state)
        params = _optim.apply_updates(params, updates)
        losses.append(loss)
        if params_list is not None:
            params_list.append(params)
    losses = torch.stack(losses) if losses else torch.zeros(0)
    states = state
    return_dict = {"params": params, "losses": losses}
    local = {"params_list": params_list, "states": states, "seeds": seeds}
    for key in return_args:
        return_dict[key] = local[key]
    return return_dict


train = train_fn


def initialize_params(params, seed, initializer=None):
    """utils.py:39-55: a flat initializer(seed, (P,)) draw unraveled into the structure of ``params``."""
    values, unravel = ravel_pytree(params)
    n = values.numel()
    random_values = brandom.normal(seed, (n,)) if initializer is None else initializer(seed, (n,))
    return unravel(random_values)


def seeds_like(params, seed, is_leaf=None):
    """utils.py:58-70: one key per leaf, split in flattening order."""
    from .spec import flatten_tree

    if is_leaf is None:
        leaves, td = tree_flatten(params)
        keys = brandom.split(seed, len(leaves))
        return tree_unflatten(td, [keys[i] for i in range(len(leaves))])
    flat = flatten_tree(params, is_leaf)
    keys = brandom.split(seed, len(flat))
    out = {}
    for i, (path, _) in enumerate(flat):
        node = out
        for k in path[:-1]:
            node = node.setdefault(k, {})
        node[path[-1]] = keys[i]
    return out

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


GRAPH_STEPS = 16  # training steps per captured CUDA graph in the device-resident loop


def _fused_loop_spec(loss_fn, optimizer, seed, return_args):
    """(model, call kwargs) if this train_fn call can run as the device-resident loop: loss_fn is
    functools.partial(model.loss_fn, outputs=..., inputs=..., full_data_size=..., n_samples=...), the optimiser is
    optim.adam, a seed is given, nothing per-step is asked back, and the model is either on one device or row-sharded with
    the peer-memory exchange (model.distributed(shard="rows", p2p=True): the all-reduce then runs on the device too)."""
    import functools

    if seed is None or return_args or getattr(optimizer, "hyper", None) is None:
        return None
    from .minibatch import MinibatchLoss

    if isinstance(loss_fn, MinibatchLoss):  # loss(params, seed) with the batch drawn on the device from the seed
        model = loss_fn.model
        if getattr(model, "_group", None) is not None or getattr(model, "packed_tril", False):
            return None
        return model, {"minibatch": loss_fn, "full_data_size": loss_fn.full_data_size, "n_samples": loss_fn.S}
    if not isinstance(loss_fn, functools.partial) or loss_fn.args:
        return None
    model = getattr(loss_fn.func, "__self__", None)
    if model is None or getattr(loss_fn.func, "__name__", "") != "loss_fn" or not hasattr(model, "_engine_for"):
        return None
    kw = dict(loss_fn.keywords)
    if not {"outputs", "inputs", "full_data_size"} <= set(kw) or set(kw) - {"outputs", "inputs", "full_data_size", "n_samples"}:
        return None
    if getattr(model, "packed_tril", False):
        return None
    if getattr(model, "_group", None) is not None and not (getattr(model, "_shard", "rows") == "rows" and getattr(model, "_p2p", False)):
        return None  # NCCL exchange or sample sharding: host-driven step + all-reduce
    return model, kw


def _train_fused(model, kw, params, optimizer, n_epochs, seed):
    """train_fn's scan (utils.py:22-31) resident on the device: bjx_train_steps captured in CUDA graphs of GRAPH_STEPS
    steps; no per-step host work, no per-replay graph update (the step number lives on the device)."""
    loc, raw, kwargs = model._split_params(params)
    engine, kw_leaves = model._engine_for(kwargs)
    dev = engine.device
    f32 = lambda t: t.detach().to(dev, torch.float32).reshape(-1)
    flat = torch.cat([f32(loc), f32(raw)] + [f32(l) for l in kw_leaves]).contiguous()
    n_epochs = int(n_epochs)
    keys = brandom.split(seed, max(n_epochs, 1)).contiguous()
    mb_state = None
    if "minibatch" in kw:  # batch buffers of the on-device sampler; every step gathers a fresh batch into them
        mb_state = kw["minibatch"]._setup(engine)
        outputs, X = mb_state["yb"], (mb_state["X"] if mb_state["indexed"] else mb_state["Xb"])
    else:
        outputs, inputs = kw["outputs"], kw["inputs"]
        X = None
        if engine.x_name is not None and inputs is not None:
            X = inputs[engine.x_name] if isinstance(inputs, dict) else inputs
    ll_scale = float(kw["full_data_size"]) / float(len(outputs))  # advi.py:92
    S = int(kw.get("n_samples", 1))
    prior_weight, exchange = 1.0, None
    if getattr(model, "_group", None) is not None:
        # row-sharded: len(outputs) of advi.py:92 is the GLOBAL length, the q / prior terms are counted on one rank, and the
        # packed [loss | grads] buffer is summed over the ranks by the peer-memory kernel inside every device-resident step
        ll_scale, prior_weight, exchange = model._shard_context(engine, len(outputs), float(kw["full_data_size"]))
    if mb_state is None:
        # normalise the dataset ONCE (device, fp32, dense columns): the warm-up step, the captured steps and the trailing
        # eager steps must all see the same persistent tensors -- a conversion inside capture would allocate from the
        # graph's private pool, miss the planes cache and re-split the dataset on every replay
        X, outputs, _, _ = engine._data(X, outputs)
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    done = torch.zeros(1, dtype=torch.int64, device=dev)
    losses = torch.empty(max(n_epochs, 1), dtype=torch.float32, device=dev)
    out = torch.empty(1 + flat.numel(), dtype=torch.float32, device=dev)

    def enqueue(n, flat_, m_, v_, done_, losses_):
        engine.train_steps(keys, n, flat_, m_, v_, optimizer.hyper, done_, losses_, out, X, outputs, ll_scale, S, minibatch=mb_state,
                           prior_weight=prior_weight, exchange=exchange)

    n_graphs, rest = divmod(n_epochs, GRAPH_STEPS)
    if n_graphs >= 2:
        # warm up on scratch state (first-use attribute calls, workspace and dataset preparation), then capture once
        enqueue(1, flat.clone(), m.clone(), v.clone(), done.clone(), None)
        if exchange is not None:
            exchange.epoch += 1  # the warm-up step used one epoch of the exchange (its scratch counter read 0)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                enqueue(GRAPH_STEPS, flat, m, v, done, losses)
        torch.cuda.current_stream(dev).wait_stream(side)
        # capture does not execute: state is untouched, the counter is still 0
        for _ in range(n_graphs):
            graph.replay()
    else:
        rest = n_epochs
    if rest:
        enqueue(rest, flat, m, v, done, losses)
    if exchange is not None:
        # the device used epochs epoch_base + 1 .. epoch_base + n_epochs (epoch_base is a launch constant, the step counter is
        # on the device); keep the host-side counter of the exchange in step for whatever uses it next
        exchange.epoch += n_epochs
    # back to the caller's pytree
    D, P = engine.D, engine.P
    key = next(k for k in ("posterior", "approx_posterior") if k in params)
    post = params[key]
    from .advi import _scale_key

    if model.vi_type == "low_rank":
        new_post = {"loc": flat[:D].reshape(loc.shape), "scale_diag": flat[D : 2 * D],
                    "scale_perturb_factor": flat[2 * D : D + P].reshape(D, int(model.rank))}
    else:
        new_post = {"loc": flat[:D].reshape(loc.shape), _scale_key(model.vi_type): flat[D : D + P].reshape(raw.shape)}
    new_params = {key: new_post}
    o = D + P
    for name in sorted(kwargs):
        ls, td = tree_flatten(kwargs[name])
        nl = []
        for l in ls:
            nl.append(flat[o : o + l.numel()].reshape(l.shape))
            o += l.numel()
        new_params[name] = tree_unflatten(td, nl)
    del post
    return {"params": new_params, "losses": losses[:n_epochs]}


def train_fn(loss_fn, params, optimizer, n_epochs, seed=None, return_args: Set[str] = frozenset(), fused: bool = True):
    """utils.py:6-36.  ``optimizer`` is optax-shaped (init / update).  Per-step keys are split(seed, n_epochs)[i] exactly
    as the reference's lax.scan consumes them.  Losses stay on the device; nothing synchronises inside the loop.
    When the call has the canonical shape (partial(model.loss_fn, ...) + optim.adam + seed) the whole loop runs on the
    device (see _train_fused); ``fused=False`` forces the generic host-driven loop."""
    spec = _fused_loop_spec(loss_fn, optimizer, seed, return_args) if fused else None
    if spec is not None and isinstance(params, dict) and any(k in params for k in ("posterior", "approx_posterior")):
        post = params[next(k for k in ("posterior", "approx_posterior") if k in params)]
        if isinstance(post, dict):
            return _train_fused(spec[0], spec[1], params, optimizer, n_epochs, seed)
    value_and_grad_fn = value_and_grad(loss_fn)
    state = state0 = optimizer.init(params)
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
    # the reference hands back any of its locals by name (utils.py:33-35, ``locals()[key]``): ``state`` is the optimiser
    # state BEFORE the scan, ``states`` the one after it
    local = {"params_list": params_list, "states": states, "seeds": seeds, "state": state0, "losses": losses, "params": params,
             "loss_fn": loss_fn, "optimizer": optimizer, "n_epochs": n_epochs, "seed": seed, "return_args": return_args,
             "value_and_grad_fn": value_and_grad_fn}
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

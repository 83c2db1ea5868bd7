"""ORACLE (test infrastructure, never shipped, never on the product path).

Float64 restatement of the optimiser side of the reference's training loop:

  * ``optax.adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0)`` as ``train_fn`` uses it --
    ``state = optimizer.init(params)``, ``updates, state = optimizer.update(grads, state)``,
    ``params = optax.apply_updates(params, updates)``    /root/reference/bijax/utils.py:8,26-27
  * the scan of ``train_fn``: per-step key ``jax.random.split(seed, n_epochs)[i]``, ``value_and_grad`` of the loss, the
    update, the loss record                              /root/reference/bijax/utils.py:22-31

optax is a third-party dependency of the reference (imported at utils.py:3, not even listed in requirements.txt, absent
from this image); the arithmetic below restates its published definition of Adam (``scale_by_adam`` followed by
``scale(-learning_rate)``):

    mu_t  = b1 * mu_{t-1} + (1 - b1) * g_t
    nu_t  = b2 * nu_{t-1} + (1 - b2) * g_t^2
    mu^_t = mu_t / (1 - b1^t),   nu^_t = nu_t / (1 - b2^t)          (t counts from 1)
    u_t   = -learning_rate * mu^_t / (sqrt(nu^_t + eps_root) + eps)   (eps OUTSIDE the root, eps_root = 0)

PARITY STATUS: "parity unpinned" against a live optax (not installable here); pinned by hand-computed steps and the
closed form of the first update (u_1 = -lr * g / (|g| + eps)) in tests/test_oracle.py.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np


class AdamState:
    def __init__(self, n: int):
        self.count = 0
        self.mu = np.zeros(n, dtype=np.float64)
        self.nu = np.zeros(n, dtype=np.float64)


def adam_init(n: int) -> AdamState:
    """optimizer.init(params) for a flat parameter vector of n entries (utils.py:8)."""
    return AdamState(int(n))


def adam_update(grads, state: AdamState, learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
                eps_root: float = 0.0) -> np.ndarray:
    """optimizer.update(grads, state) (utils.py:26): returns the additive updates, advances ``state`` in place."""
    g = np.asarray(grads, dtype=np.float64).reshape(-1)
    state.count += 1
    state.mu = b1 * state.mu + (1.0 - b1) * g
    state.nu = b2 * state.nu + (1.0 - b2) * g * g
    mu_hat = state.mu / (1.0 - b1 ** state.count)
    nu_hat = state.nu / (1.0 - b2 ** state.count)
    return -learning_rate * mu_hat / (np.sqrt(nu_hat + eps_root) + eps)


def train(value_and_grad: Callable[[np.ndarray, np.ndarray], Tuple[float, np.ndarray]], params0, keys, learning_rate: float,
          b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> Dict[str, np.ndarray]:
    """train_fn's scan (utils.py:22-31) on a flat float64 parameter vector: for every key of ``keys`` (= split(seed,
    n_epochs)) take ``value_and_grad(params, key) -> (loss, flat grads)``, apply Adam, record the loss."""
    params = np.asarray(params0, dtype=np.float64).reshape(-1).copy()
    state = adam_init(params.size)
    losses = []
    for key in keys:
        loss, grads = value_and_grad(params, key)
        params = params + adam_update(grads, state, learning_rate, b1, b2, eps)  # optax.apply_updates (utils.py:27)
        losses.append(float(loss))
    return {"params": params, "losses": np.asarray(losses), "state": state}

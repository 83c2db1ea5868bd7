"""Posterior access (bijax/core.py:39-75): what ``ADVI.apply(params)`` returns.

``sample`` runs on the fused CUDA prologue (threefry -> z -> bijector); ``log_prob`` is one fused kernel
(bjx_posterior_log_prob: inverse bijector chains, log-det-Jacobians, diagonal or triangular-solve Gaussian log density).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import _native as nv
from .spec import flatten_tree

class Posterior:
    def __init__(self, model, engine, loc: torch.Tensor, raw_scale: torch.Tensor):
        self.model, self.engine = model, engine
        self.loc, self.raw_scale = loc, raw_scale

    def sample(self, seed, sample_shape=()):
        """core.py:46-55: constrained samples as a pytree with leading ``sample_shape``."""
        shape = tuple(int(s) for s in ((sample_shape,) if isinstance(sample_shape, int) else sample_shape))
        n = math.prod(shape) if shape else 1
        theta = self.engine.posterior_sample(seed, self.loc, self.raw_scale, n)
        return self.model.latents.unravel(theta.reshape(*shape, self.model.size))

    def log_prob(self, sample: Dict, sample_shape=()):
        """core.py:57-71: q(z = b^{-1}(theta)) + sum inverse_log_det_jacobian(theta), one fused kernel over the flattened
        samples (bjx_posterior_log_prob: inverse bijector chains, triangular solve for the full-rank family)."""
        lat = self.model.latents
        dev = self.engine.device
        leaves = dict(flatten_tree(sample, lambda x: isinstance(x, torch.Tensor)))
        nb = len(tuple(sample_shape) if not isinstance(sample_shape, int) else (sample_shape,))
        cols, batch = [], ()
        for path in lat.paths:
            y = leaves[path].to(dev, torch.float32)
            batch = tuple(y.shape[:nb])
            cols.append(y.reshape(*batch, -1))
        theta = torch.cat(cols, dim=-1).reshape(-1, self.model.size)
        out = self.engine.posterior_log_prob(theta, self.loc, self.raw_scale)
        return out.reshape(batch)

    def prob(self, sample, sample_shape=()):
        return torch.exp(self.log_prob(sample, sample_shape))

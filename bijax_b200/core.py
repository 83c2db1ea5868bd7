"""Posterior access (bijax/core.py:39-75): what ``ADVI.apply(params)`` returns.

``sample`` runs on the fused CUDA prologue (threefry -> z -> bijector); ``log_prob`` is post-training, off the per-step
path, and is expressed with torch ops on the device.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import _native as nv
from .spec import bijector_ops, flatten_tree, is_distribution

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def _inverse_and_fldj(y: torch.Tensor, ops):
    """z = b^{-1}(y) and fldj_b(z) for ops given in application order."""
    sp = torch.nn.functional.softplus
    xs = y
    for op in reversed(ops):
        if op == nv.BIJ_EXP:
            xs = torch.log(xs)
        elif op == nv.BIJ_SOFTPLUS:
            xs = xs + torch.log(-torch.expm1(-xs))
        elif op == nv.BIJ_SIGMOID:
            xs = torch.log(xs) - torch.log1p(-xs)
    z = xs
    fldj = torch.zeros_like(z)
    x = z
    for op in ops:
        if op == nv.BIJ_EXP:
            fldj = fldj + x
            x = torch.exp(x)
        elif op == nv.BIJ_SOFTPLUS:
            fldj = fldj - sp(-x)
            x = sp(x)
        elif op == nv.BIJ_SIGMOID:
            fldj = fldj - sp(-x) - sp(x)
            x = torch.sigmoid(x)
    return z, fldj


class Posterior:
    def __init__(self, model, engine, loc: torch.Tensor, raw_scale: torch.Tensor):
        self.model, self.engine = model, engine
        self.loc, self.raw_scale = loc, raw_scale

    def _scale(self):
        name = type(self.model.posterior_params_bijector[1]).__name__
        raw = self.raw_scale.to(self.engine.device, torch.float32)
        s = torch.exp(raw) if name == "Exp" else torch.nn.functional.softplus(raw) if name == "Softplus" else raw
        return s if self.model.vi_type == "mean_field" else torch.tril(s)

    def sample(self, seed, sample_shape=()):
        """core.py:46-55: constrained samples as a pytree with leading ``sample_shape``."""
        shape = tuple(int(s) for s in ((sample_shape,) if isinstance(sample_shape, int) else sample_shape))
        n = math.prod(shape) if shape else 1
        theta = self.engine.posterior_sample(seed, self.loc, self.raw_scale, n)
        return self.model.latents.unravel(theta.reshape(*shape, self.model.size))

    def log_prob(self, sample: Dict, sample_shape=()):
        """core.py:57-71: q(z = b^{-1}(theta)) + sum inverse_log_det_jacobian(theta)."""
        lat = self.model.latents
        dev = self.engine.device
        zs, ildj = [], 0.0
        leaves = dict(flatten_tree(sample, lambda x: isinstance(x, torch.Tensor)))
        bij = dict(flatten_tree(self.model.prior_constraints, lambda x: not isinstance(x, dict)))
        nb = len(tuple(sample_shape) if not isinstance(sample_shape, int) else (sample_shape,))
        for path, shp in zip(lat.paths, lat.shapes):
            y = leaves[path].to(dev, torch.float32)
            batch = y.shape[:nb]
            z, fldj = _inverse_and_fldj(y.reshape(*batch, -1), bijector_ops(bij[path]))
            zs.append(z)
            ildj = ildj - fldj.sum(-1)
        z = torch.cat(zs, dim=-1)
        loc = self.loc.to(dev, torch.float32)
        sc = self._scale()
        if self.model.vi_type == "mean_field":
            zz = (z - loc) / sc
            q = (-0.5 * zz * zz - torch.log(sc) - _HALF_LOG_2PI).sum(-1)
        else:
            diff = (z - loc).reshape(-1, z.shape[-1]).T
            zz = torch.linalg.solve_triangular(sc, diff, upper=False).T.reshape(z.shape)
            q = -0.5 * (zz * zz).sum(-1) - torch.log(torch.abs(torch.diagonal(sc))).sum() - z.shape[-1] * _HALF_LOG_2PI
        return q + ildj

    def prob(self, sample, sample_shape=()):
        return torch.exp(self.log_prob(sample, sample_shape))

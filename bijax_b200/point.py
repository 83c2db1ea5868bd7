"""MAP (bijax/map.py:12-37) and the log-joint of MCMC (bijax/mcmc.py:16-34) on the engine's likelihood kernels.

Both are POINT evaluations of the model at an unconstrained value z (no sampling, no variational term): the same fused
step with S = 1, eps = 0 and ``point_mode`` set (include/bjx.h), so the S x N x D contraction, its transpose and the
prior / bijector terms run on exactly the kernels ADVI uses.  What is NOT here: tfp.mcmc.NoUTurnSampler / sample_chain
(mcmc.py:36-59) -- a different algorithm, outside the hot path.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import _native as nv
from . import likelihoods as lk
from . import random as brandom
from .engine import ElboEngine
from .spec import compile_latents, fill_in_bijector, flatten_tree


class _PointFunction(torch.autograd.Function):
    """custom_vjp for a point evaluation: forward computes the value and d/dz in one launch sequence."""

    @staticmethod
    def forward(ctx, z, run):
        out = run(z.detach())
        ctx.save_for_backward(out)
        ctx.n = z.numel()
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        return g * out[1 : 1 + ctx.n], None


class _PointModel:
    def __init__(self, prior: Dict[str, Any], likelihood, constraints: Optional[Dict], device=None, precision="auto", kernel=None):
        if not isinstance(likelihood, lk.Likelihood):
            raise NotImplementedError(
                "arbitrary python likelihood callables are not on the fused CUDA path; pass one of "
                "bijax_b200.likelihoods.{BernoulliLogits, GaussianLinear, BernoulliProbs}"
            )
        self.prior = prior
        self.prior_constraints = fill_in_bijector({} if constraints is None else constraints, self.prior)
        self.bijector = self.prior_constraints
        assert self.prior_constraints.keys() == self.prior.keys(), "The keys in `prior` and `prior_constraints` must be the same."
        self.likelihood_fn = self.log_likelihood_fn = likelihood
        self.latents = compile_latents(self.prior, self.prior_constraints)
        self.size = self.latents.size
        self.unravel_fn = self.latents.unravel
        self._engine = ElboEngine(self.latents, likelihood, "mean_field", None, (), device=device, precision=precision, kernel=kernel)

    # -- tree <-> flat (sorted-key order, row-major ravel: core.py:168-176)
    def _ravel(self, tree) -> torch.Tensor:
        leaves = dict(flatten_tree(tree, lambda x: isinstance(x, torch.Tensor)))
        dev = self._engine.device
        return torch.cat([leaves[p].to(dev, torch.float32).reshape(-1) for p in self.latents.paths])

    def _evaluate(self, tree, outputs, inputs, mode: int) -> torch.Tensor:
        eng = self._engine
        X = None
        if eng.lik_id != nv.LIK_BERNOULLI_PROBS:
            X = inputs[eng.x_name] if isinstance(inputs, dict) else inputs
        z = self._ravel(tree)

        def run(zd):
            return eng.point_step(zd, X, outputs, 1.0, mode)

        return _PointFunction.apply(z, run)

    def _transform(self, tree):
        """transform_tree(tree, bijector) (core.py:154-155) through the fused prologue."""
        z = self._ravel(tree)
        theta = self._engine.point_transform(z)
        return self.latents.unravel(theta)

    def _init_from_normal_prior(self, seed):
        """sample_dist (core.py:124-130) for priors whose sampler is jax.random.normal (tfd.Normal / MultivariateNormalDiag):
        one key per leaf in flattening order (seeds_like, utils.py:58-70), sample = loc + scale * normal(key, shape)."""
        import numpy as np

        from .spec import is_distribution

        leaves = flatten_tree(self.prior, is_distribution)
        keys = brandom.split(seed, len(leaves))
        dev = self._engine.device
        out: Dict[str, Any] = {}
        for i, (path, d) in enumerate(leaves):
            name = type(d).__name__
            if name == "Normal":
                loc, scale = np.asarray(d.loc, np.float32), np.asarray(d.scale, np.float32)
            elif name == "MultivariateNormalDiag":
                loc, scale = np.asarray(d.loc, np.float32), np.asarray(d.scale_diag, np.float32)
            else:
                raise NotImplementedError(
                    f"init(): sampling a {name} prior needs TFP's rejection samplers, which are outside the fused path; "
                    "pass your own initial value"
                )
            shape = np.broadcast(loc, scale).shape
            eps = brandom.normal(keys[i], tuple(shape))
            val = torch.as_tensor(np.broadcast_to(loc, shape).copy(), device=dev) + torch.as_tensor(np.broadcast_to(scale, shape).copy(), device=dev) * eps
            node = out
            for k in path[:-1]:
                node = node.setdefault(k, {})
            node[path[-1]] = val
        return out


class MAP(_PointModel):
    """bijax/map.py:12-37.  ``params = {"params": unconstrained tree}``."""

    def __init__(self, prior, likelihood_fn, prior_constraints=None, **kw):
        super().__init__(prior, likelihood_fn, prior_constraints, **kw)

    def init(self, seed):
        """map.py:22-25: a prior sample pulled back through the bijectors (Normal-family priors; see _init_from_normal_prior)."""
        theta = self._init_from_normal_prior(seed)
        for path, _ in flatten_tree(self.prior_constraints, lambda x: not isinstance(x, dict)):
            b = self.prior_constraints
            for k in path:
                b = b[k]
            if type(b).__name__ != "Identity":
                raise NotImplementedError("init(): only Identity-constrained Normal priors are sampled here; pass your own initial value")
        return {"params": theta}

    def neg_log_joint(self, params, outputs, inputs=None):
        """map.py:27-33: -(log p(theta) + log_likelihood(theta)), theta = bijector(params["params"]); differentiable."""
        return self._evaluate(params["params"], outputs, inputs, nv.POINT_MAP)

    def apply(self, params):
        return self._transform(params["params"])


class MCMC(_PointModel):
    """The model side of bijax/mcmc.py:16-34 (log_joint over the unconstrained tree); the NUTS driver is out of scope."""

    def __init__(self, prior, bijector=None, log_likelihood_fn=None, **kw):
        super().__init__(prior, log_likelihood_fn, bijector, **kw)

    def log_joint(self, params, outputs, inputs):
        """mcmc.py:30-34: log p~(z) + log_likelihood(bijector(z)); differentiable."""
        return -self._evaluate(params, outputs, inputs, nv.POINT_LOG_JOINT)

    def transform(self, params):
        return self._transform(params)


# ---- the likelihood object as the reference's callable (README.md:81-106) ------------------------------------------------------
class _LogLikFunction(torch.autograd.Function):
    """custom_vjp of log_likelihood_fn(latent_sample, outputs, inputs, **kwargs): one point evaluation gives the value and the
    derivatives with respect to the constrained latent vector and the trainable kwargs."""

    @staticmethod
    def forward(ctx, theta, kw, run):
        out = run(theta.detach(), None if kw is None else kw.detach())
        ctx.save_for_backward(out)
        ctx.n, ctx.k = theta.numel(), 0 if kw is None else kw.numel()
        return -out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        n, k = ctx.n, ctx.k
        return -g * out[1 : 1 + n], (-g * out[1 + 2 * n : 1 + 2 * n + k]) if k else None, None


def log_likelihood(L, latent_sample, outputs, inputs, kwargs) -> torch.Tensor:
    """The scalar log-likelihood of ``outputs`` at the constrained ``latent_sample`` (see likelihoods.Likelihood.__call__)."""
    import numpy as np

    from . import distributions as tfd

    names = sorted(L.latent_names())
    missing = [n for n in names if n not in latent_sample]
    if missing:
        raise KeyError(f"latent_sample has no entry {missing[0]!r}")
    knames = sorted(L.kwarg_names())
    for n in knames:
        if n not in kwargs:
            raise KeyError(f"the likelihood needs the keyword argument {n!r}")
    leaves = [torch.as_tensor(latent_sample[n]) for n in names]
    dev = next((l.device for l in leaves if l.is_cuda), torch.device("cuda", torch.cuda.current_device()))
    sig = (tuple((n, tuple(l.shape)) for n, l in zip(names, leaves)), tuple(knames), str(dev))
    cache = L.__dict__.setdefault("_point_engines", {})  # lives and dies with the likelihood object
    if sig not in cache:
        # only the leaves the likelihood reads, in the flat layout of core.py:168-176; the priors are placeholders (never evaluated)
        prior = {n: tfd.Normal(np.zeros(tuple(l.shape), np.float32), 1.0) for n, l in zip(names, leaves)}
        latents = compile_latents(prior, fill_in_bijector({}, prior))
        cache[sig] = ElboEngine(latents, L, "mean_field", None, [(n, 1) for n in knames], device=dev)
    eng = cache[sig]
    X = None
    if eng.lik_id != nv.LIK_BERNOULLI_PROBS:
        X = inputs[eng.x_name] if isinstance(inputs, dict) else inputs
    theta = torch.cat([l.to(dev, torch.float32).reshape(-1) for l in leaves])
    kw = torch.cat([torch.as_tensor(kwargs[n]).to(dev, torch.float32).reshape(-1) for n in knames]) if knames else None

    def run(t, k):
        return eng.point_step(t, X, outputs, 1.0, nv.POINT_LOGLIK, kwargs_vec=k)

    return _LogLikFunction.apply(theta, kw, run)

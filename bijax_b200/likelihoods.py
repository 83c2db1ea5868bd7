"""Recognised likelihood families (north_star item 5).  Each object is what a user passes as ``log_likelihood_fn``
(v1, README.md:81-106) / ``likelihood_fn`` (v2, advi.py:90-91): declarative, so the engine can run the S x N x D
contraction and its transpose as one fused pass over X instead of differentiating user code.

    BernoulliLogits(weights="weights", X="X")
        == lambda latent, outputs, inputs, **kw: tfd.Bernoulli(logits=inputs["X"] @ latent["weights"]).log_prob(outputs).sum()
    GaussianLinear(weights="weights", X="X", log_noise_scale="log_noise_scale")
        == README.md:99-106 (learnable noise passed as a trainable kwarg), with X laid out [N, D]
    BernoulliProbs(probs="p_of_heads")
        == README.md:91-95 (coin toss; no inputs)
"""
from __future__ import annotations

from typing import Optional


class Likelihood:
    """Base of the recognised families.  Besides describing the likelihood to the engine, an instance IS the callable of
    the reference's contract (README.md:81-106):

        log_likelihood_fn(latent_sample, outputs, inputs, **kwargs) -> scalar log-likelihood

    with ``latent_sample`` a dict of CONSTRAINED latent values (e.g. one posterior sample, ``{"weights": w}``) and ``kwargs`` the
    trainable likelihood parameters (e.g. ``log_noise_scale=...``).  The call runs on the same CUDA kernels as the ELBO step
    (a point evaluation, ``BJX_POINT_LOGLIK``) and is differentiable with respect to the latent leaves and the kwargs."""

    def latent_names(self):
        raise NotImplementedError

    def kwarg_names(self):
        return ()

    def __call__(self, latent_sample, outputs, inputs=None, **kwargs):
        from .point import log_likelihood

        return log_likelihood(self, latent_sample, outputs, inputs, kwargs)


class BernoulliLogits(Likelihood):
    def __init__(self, weights: str = "weights", X: str = "X"):
        self.weights, self.X = weights, X

    def latent_names(self):
        return (self.weights,)


class GaussianLinear(Likelihood):
    """Exactly one of: ``log_noise_scale`` (name of a trainable kwarg holding log tau), ``noise_latent`` (name of a scalar
    latent holding tau itself), ``noise_scale`` (a constant)."""

    def __init__(self, weights: str = "weights", X: str = "X", log_noise_scale: Optional[str] = None,
                 noise_latent: Optional[str] = None, noise_scale: Optional[float] = None):
        given = [v is not None for v in (log_noise_scale, noise_latent, noise_scale)]
        if sum(given) == 0:
            log_noise_scale = "log_noise_scale"
        elif sum(given) > 1:
            raise ValueError("give exactly one of log_noise_scale, noise_latent, noise_scale")
        self.weights, self.X = weights, X
        self.log_noise_scale, self.noise_latent, self.noise_scale = log_noise_scale, noise_latent, noise_scale

    def latent_names(self):
        return (self.weights,) + ((self.noise_latent,) if self.noise_latent is not None else ())

    def kwarg_names(self):
        return (self.log_noise_scale,) if self.log_noise_scale is not None else ()


class BernoulliProbs(Likelihood):
    def __init__(self, probs: str):
        self.probs = probs

    def latent_names(self):
        return (self.probs,)

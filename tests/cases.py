"""Matched (oracle, product) model builders for the parity tests.  The oracle (oracle/) is the checker only."""
from __future__ import annotations

import math

import numpy as np
import torch

import bijax_b200 as bj
from bijax_b200 import bijectors as tfb
from bijax_b200 import distributions as tfd
from bijax_b200 import likelihoods as lk
from oracle import advi_ref as R
from oracle import threefry as tf


def synth_X(N, D, seed=1234):
    """The bench's synthetic design matrix on the CPU: X[n, d] = normal(PRNGKey(seed))[n*D+d] / sqrt(D), partitionable."""
    x = tf.normal(tf.prng_key(seed), (N, D), tf.LAYOUT_PARTITIONABLE)
    return (x / np.float32(math.sqrt(D))).astype(np.float32)


def synth_logistic(N, D):
    X = synth_X(N, D)
    w = tf.normal(tf.prng_key(1235), (D,), tf.LAYOUT_PARTITIONABLE)
    u = tf.uniform(tf.prng_key(1236), (N,), tf.LAYOUT_PARTITIONABLE)
    p = 1.0 / (1.0 + np.exp(-(X.astype(np.float64) @ w.astype(np.float64))))
    y = (u < p).astype(np.float32)
    return X, y


def synth_linear(N, D, noise=0.1):
    X = synth_X(N, D)
    w = tf.normal(tf.prng_key(1235), (D,), tf.LAYOUT_PARTITIONABLE)
    e = tf.normal(tf.prng_key(1237), (N,), tf.LAYOUT_PARTITIONABLE)
    y = (X.astype(np.float64) @ w.astype(np.float64) + noise * e).astype(np.float32)
    return X, y


_ORACLE_BIJ = {"Identity": R.Identity, "Exp": R.Exp, "Softplus": R.Softplus, "Sigmoid": R.Sigmoid}


def oracle_bijector(b):
    if type(b).__name__ == "Chain":
        return R.Chain([oracle_bijector(x) for x in b.bijectors])
    return _ORACLE_BIJ[type(b).__name__]()


def oracle_prior(d):
    n = type(d).__name__
    t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64))
    if n == "Normal":
        return R.Normal(t(d.loc), t(d.scale))
    if n == "MultivariateNormalDiag":
        return R.MultivariateNormalDiag(t(d.loc), t(d.scale_diag))
    if n == "Beta":
        return R.Beta(t(d.concentration1), t(d.concentration0))
    if n == "Gamma":
        return R.Gamma(t(d.concentration), t(d.rate))
    raise NotImplementedError(n)


def oracle_likelihood(L):
    if isinstance(L, lk.BernoulliProbs):
        return R.bernoulli_probs_loglik(L.probs)
    if isinstance(L, lk.BernoulliLogits):
        return R.bernoulli_logits_loglik(L.weights)
    if isinstance(L, lk.GaussianLinear):
        if L.log_noise_scale is not None:
            return R.gaussian_linear_loglik(L.weights, L.log_noise_scale)
        if L.noise_latent is not None:
            name = L.noise_latent

            def fn(latent, outputs, inputs, **kw):
                loc = inputs["X"] @ latent[L.weights]
                scale = latent[name]
                zz = (outputs - loc) / scale
                return (-0.5 * zz * zz - 0.5 * R.LOG_2PI - torch.log(scale)).sum()

            return fn
        c = float(L.noise_scale)

        def fn2(latent, outputs, inputs, **kw):
            loc = inputs["X"] @ latent[L.weights]
            zz = (outputs - loc) / c
            return (-0.5 * zz * zz - 0.5 * R.LOG_2PI - math.log(c)).sum()

        return fn2
    raise NotImplementedError


def build_pair(prior, bijector, likelihood, vi_type="mean_field", scale_bijector=None, precision="auto", kernel=None, prepare_dataset=True):
    opb = None
    if scale_bijector is not None:
        opb = [tfb.Identity(), scale_bijector]
    model = bj.ADVI(prior, dict(bijector), likelihood, vi_type=vi_type, ordered_posterior_bijectors=opb, precision=precision, kernel=kernel,
                    prepare_dataset=prepare_dataset)
    oprior = {k: oracle_prior(v) for k, v in prior.items()}
    obij = {k: oracle_bijector(v) for k, v in model.prior_constraints.items()}
    ref = R.ADVIRef(oprior, obij, oracle_likelihood(likelihood), vi_type,
                    scale_bijector=oracle_bijector(scale_bijector) if scale_bijector is not None else None)
    return model, ref


def oracle_step(ref, loc, raw, kwargs, y, X, full_data_size, key_np, S, layout, dtype=torch.float64):
    """float64 oracle on float32 eps drawn by the oracle PRNG for the same key (``dtype=torch.float32``: the same
    restatement in the reference's own working precision, used to measure how well-conditioned an output is)."""
    D = ref.size
    eps = torch.tensor(tf.normal(key_np, (S, D), layout), dtype=dtype)
    t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dtype)
    kw = {k: t64(v) for k, v in kwargs.items()}
    inputs = {"X": t64(X)} if X is not None else None
    return ref.value_and_grad(t64(loc), t64(raw), kw, t64(y), inputs, full_data_size, eps), eps


def assert_close(got, want, rtol=1e-5, atol_scale=1e-5, what=""):
    """rtol 1e-5 as north_star states: ``tol = max(rtol*|want|, atol_scale*max|want|)`` -- entries that are tiny relative to
    the largest one of their block get the block's absolute floor, the large ones exactly rtol."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert np.isfinite(want).all(), f"{what}: oracle value is not finite"
    assert np.isfinite(got).all(), f"{what}: result is not finite: {got}"
    scale = max(float(np.max(np.abs(want))), 1e-30)
    err = np.abs(got - want)
    tol = np.maximum(rtol * np.abs(want), atol_scale * scale)
    bad = err > tol
    assert not bad.any(), f"{what}: max err {err.max():.3e} (rel-to-max {err.max() / scale:.3e}) at {np.argmax(err)}; want {want.ravel()[np.argmax(err)]:.6e} got {got.ravel()[np.argmax(err)]:.6e}"

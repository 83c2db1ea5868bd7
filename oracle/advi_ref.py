"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of bijax's reparameterised Monte-Carlo ELBO and of the
``jax.value_and_grad`` the reference takes of it:

  * loss ............ /root/reference/bijax/advi.py:80-95
  * unconstrained prior, log-prob sum, transforms, families and their default
    parameter bijectors .. /root/reference/bijax/core.py:114-121,133-165,168-212
  * value_and_grad ... /root/reference/bijax/utils.py:7,25 (restated with
    torch autograd on the CPU, any dtype; float64 is the checker, float32 with
    all host threads is the timed CPU baseline)
  * likelihood exemplars .. /root/reference/README.md:91-106,
    /root/reference/examples/advi/coin_toss.ipynb

The distribution / bijector arithmetic lives in tensorflow_probability
(unpinned, /root/reference/requirements.txt:1, absent from this image); the
formulas below restate TFP's published log_prob / forward /
forward_log_det_jacobian definitions.

epsilon is NOT drawn here: callers pass the standard-normal draws produced by
oracle/threefry.py for the same key, exactly as ``q.sample(seed=...)`` would
(MultivariateNormalDiag / TriL sample = loc + scale @ normal(seed, (S, D))).

PARITY STATUS: "parity unpinned" against a live bijax/JAX/TFP (none can be
installed here, the reference has no tests or golden vectors); pinned instead
against analytic results (conjugate posteriors, closed-form gradients checked
by central finite differences in tests/test_oracle.py).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence

import torch

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------- bijectors (tfb.*)
class Identity:
    def forward(self, x):
        return x

    def fldj(self, x):
        return torch.zeros_like(x)

    def inverse(self, y):
        return y


class Exp:
    def forward(self, x):
        return torch.exp(x)

    def fldj(self, x):
        return x

    def inverse(self, y):
        return torch.log(y)


class Softplus:
    def forward(self, x):
        return torch.nn.functional.softplus(x)

    def fldj(self, x):
        return -torch.nn.functional.softplus(-x)

    def inverse(self, y):
        return y + torch.log(-torch.expm1(-y))


class Sigmoid:
    def forward(self, x):
        return torch.sigmoid(x)

    def fldj(self, x):
        sp = torch.nn.functional.softplus
        return -sp(-x) - sp(x)

    def inverse(self, y):
        return torch.log(y) - torch.log1p(-y)


class Chain:
    """tfb.Chain([b1, b2, ...]) = b1 o b2 o ... (rightmost applied first)."""

    def __init__(self, bijectors: Sequence):
        self.bijectors = list(bijectors)

    def forward(self, x):
        for b in reversed(self.bijectors):
            x = b.forward(x)
        return x

    def fldj(self, x):
        total = torch.zeros_like(x)
        for b in reversed(self.bijectors):
            total = total + b.fldj(x)
            x = b.forward(x)
        return total

    def inverse(self, y):
        for b in self.bijectors:
            y = b.inverse(y)
        return y


# --------------------------------------------------------------------------- distributions (tfd.*)
def _t(v, like):
    return torch.as_tensor(v, dtype=like.dtype, device=like.device)


class Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def event_like_shape(self):
        return tuple(torch.broadcast_shapes(torch.as_tensor(self.loc).shape, torch.as_tensor(self.scale).shape))

    def log_prob(self, x):
        loc, scale = _t(self.loc, x), _t(self.scale, x)
        zz = (x - loc) / scale
        return -0.5 * zz * zz - 0.5 * LOG_2PI - torch.log(scale)


class MultivariateNormalDiag:
    def __init__(self, loc, scale_diag):
        self.loc, self.scale_diag = loc, scale_diag

    def event_like_shape(self):
        return tuple(torch.broadcast_shapes(torch.as_tensor(self.loc).shape, torch.as_tensor(self.scale_diag).shape))

    def log_prob(self, x):
        loc, scale = _t(self.loc, x), _t(self.scale_diag, x)
        zz = (x - loc) / scale
        return (-0.5 * zz * zz - 0.5 * LOG_2PI - torch.log(scale)).sum(-1)


class Beta:
    def __init__(self, concentration1, concentration0):
        self.concentration1, self.concentration0 = concentration1, concentration0

    def event_like_shape(self):
        return tuple(
            torch.broadcast_shapes(
                torch.as_tensor(self.concentration1).shape, torch.as_tensor(self.concentration0).shape
            )
        )

    def log_prob(self, x):
        a, b = _t(self.concentration1, x), _t(self.concentration0, x)
        lbeta = torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b)
        return torch.xlogy(a - 1.0, x) + torch.special.xlog1py(b - 1.0, -x) - lbeta


class Gamma:
    def __init__(self, concentration, rate):
        self.concentration, self.rate = concentration, rate

    def event_like_shape(self):
        return tuple(
            torch.broadcast_shapes(torch.as_tensor(self.concentration).shape, torch.as_tensor(self.rate).shape)
        )

    def log_prob(self, x):
        a, r = _t(self.concentration, x), _t(self.rate, x)
        return torch.xlogy(a - 1.0, x) - r * x - torch.lgamma(a) + a * torch.log(r)


# --------------------------------------------------------------------------- likelihood exemplars
def bernoulli_probs_loglik(name: str):
    """Coin toss: tfd.Bernoulli(probs=p).log_prob(outputs).sum() (README.md:91-95)."""

    def fn(latent, outputs, inputs, **kwargs):
        p = latent[name]
        k = outputs
        return ((1.0 - k) * torch.log1p(-p) + k * torch.log(p)).sum()

    return fn


def bernoulli_logits_loglik(name: str = "weights"):
    """Bayesian logistic regression: tfd.Bernoulli(logits=X@w).log_prob(y).sum()."""
    sp = torch.nn.functional.softplus

    def fn(latent, outputs, inputs, **kwargs):
        logits = inputs["X"] @ latent[name]
        k = outputs
        return (-(1.0 - k) * sp(logits) - k * sp(-logits)).sum()

    return fn


def gaussian_linear_loglik(name: str = "weights", noise: str = "log_noise_scale"):
    """Linear regression with learnable noise (README.md:99-106); X is [N, D]."""

    def fn(latent, outputs, inputs, **kwargs):
        loc = inputs["X"] @ latent[name]
        scale = torch.exp(kwargs[noise])
        zz = (outputs - loc) / scale
        return (-0.5 * zz * zz - 0.5 * LOG_2PI - torch.log(scale)).sum()

    return fn


# --------------------------------------------------------------------------- the model
class ADVIRef:
    """Restates ADVI.__init__ / loss_fn (advi.py:26-95) on torch-CPU tensors."""

    def __init__(
        self,
        prior: Dict[str, object],
        bijector: Dict[str, object],
        log_likelihood_fn: Callable,
        vi_type: str = "mean_field",
        scale_bijector=None,
    ):
        self.prior = dict(prior)
        self.bijector = dict(bijector)
        for k in self.prior:  # fill_in_bijector, core.py:107-111
            self.bijector.setdefault(k, Identity())
        assert self.bijector.keys() == self.prior.keys()
        self.log_likelihood_fn = log_likelihood_fn
        self.vi_type = vi_type
        # get_mean_field default [Identity, Exp] (core.py:180-181); full rank [Identity, Identity] (core.py:205-206)
        # (low rank: [Identity, Exp, Identity], core.py:191-192)
        self.scale_bijector = scale_bijector if scale_bijector is not None else (Identity() if vi_type == "full_rank" else Exp())
        # flat layout: sorted keys, row-major ravel (core.py:168-176, jax dict pytrees flatten in sorted-key order)
        self.names = sorted(self.prior)
        self.shapes = {k: self.prior[k].event_like_shape() for k in self.names}
        self.sizes = {k: int(math.prod(self.shapes[k])) if self.shapes[k] else 1 for k in self.names}
        self.size = sum(self.sizes.values())

    def unravel(self, z):
        out, o = {}, 0
        for k in self.names:
            n = self.sizes[k]
            out[k] = z[o : o + n].reshape(self.shapes[k])
            o += n
        return out

    def loss(self, loc, raw_scale, kwargs, outputs, inputs, full_data_size, eps):
        """advi.py:80-95. ``eps`` is [S, D] = normal(seed, (S, D)); returns the scalar loss."""
        dt = loc.dtype
        eps = eps.to(dt)
        D = self.size
        if self.vi_type == "mean_field":
            sigma = self.scale_bijector.forward(raw_scale)  # transform_dist_params, core.py:162-165
            samples = loc + sigma * eps  # q.sample, advi.py:83
        else:
            tril = torch.tril(self.scale_bijector.forward(raw_scale))  # MVNTriL masks with band_part
            samples = loc + eps @ tril.T

        def loss_no_batch(sample):
            if self.vi_type == "mean_field":
                zz = (sample - loc) / sigma
                q_log_prob = (-0.5 * zz * zz - torch.log(sigma)).sum() - 0.5 * D * LOG_2PI
            else:
                zz = torch.linalg.solve_triangular(tril, (sample - loc)[:, None], upper=False)[:, 0]
                q_log_prob = -0.5 * (zz * zz).sum() - torch.log(torch.abs(torch.diagonal(tril))).sum() - 0.5 * D * LOG_2PI
            tree = self.unravel(sample)
            p_log_prob = 0.0
            theta = {}
            for k in self.names:  # log_prob_dist(approx_normal_prior, tree): prior(b(z)) + fldj(z), core.py:114-121,133-151
                b = self.bijector[k]
                th = b.forward(tree[k])
                p_log_prob = p_log_prob + (self.prior[k].log_prob(th).sum() + b.fldj(tree[k]).sum())
                theta[k] = th  # transform_tree, core.py:154-155
            ll = self.log_likelihood_fn(theta, outputs, inputs, **kwargs)
            ll = (ll / len(outputs)) * full_data_size  # advi.py:92
            return q_log_prob - p_log_prob - ll

        return torch.stack([loss_no_batch(samples[s]) for s in range(samples.shape[0])]).mean()

    def posterior_log_prob(self, loc, raw_scale, sample: Dict[str, torch.Tensor]):
        """Posterior.log_prob (core.py:57-71) for ONE constrained sample tree: inverse-transform every leaf
        (inverse_transform_tree), ravel, q.log_prob of the fitted Gaussian, plus the sum of inverse_log_det_jacobian
        (= -fldj at the pre-image)."""
        zs, ildj = [], 0.0
        for k in self.names:
            b = self.bijector[k]
            th = torch.as_tensor(sample[k], dtype=loc.dtype).reshape(-1)
            z = b.inverse(th)
            zs.append(z)
            ildj = ildj - b.fldj(z).sum()
        z = torch.cat(zs)
        D = self.size
        if self.vi_type == "mean_field":
            sigma = self.scale_bijector.forward(raw_scale)
            zz = (z - loc) / sigma
            q = (-0.5 * zz * zz - torch.log(sigma)).sum() - 0.5 * D * LOG_2PI
        elif self.vi_type == "low_rank":  # raw_scale = (raw scale_diag, scale_perturb_factor)
            raw_d, U = raw_scale
            A = torch.diag(self.scale_bijector.forward(raw_d)) + U @ U.T
            zz = torch.linalg.solve(A, z - loc)
            q = -0.5 * (zz * zz).sum() - torch.linalg.slogdet(A)[1] - 0.5 * D * LOG_2PI
        else:
            tril = torch.tril(self.scale_bijector.forward(raw_scale))
            zz = torch.linalg.solve_triangular(tril, (z - loc)[:, None], upper=False)[:, 0]
            q = -0.5 * (zz * zz).sum() - torch.log(torch.abs(torch.diagonal(tril))).sum() - 0.5 * D * LOG_2PI
        return q + ildj

    def loss_low_rank(self, loc, raw_diag, U, kwargs, outputs, inputs, full_data_size, eps):
        """advi.py:80-95 for vi_type="low_rank" (core.py:190-201): q = MultivariateNormalDiagPlusLowRank(loc, d, U) with
        d = scale_bijector(raw_diag) (default Exp), i.e. scale operator A = diag(d) + U U^T; sample = loc + A eps;
        q.log_prob(z) = -1/2 |A^-1 (z - loc)|^2 - log|det A| - D/2 log 2 pi (dense solve / slogdet here, float64)."""
        dt = loc.dtype
        eps = eps.to(dt)
        D = self.size
        d = self.scale_bijector.forward(raw_diag)
        A = torch.diag(d) + U @ U.T
        samples = loc + eps @ A.T
        logdet = torch.linalg.slogdet(A)[1]

        def loss_no_batch(sample):
            zz = torch.linalg.solve(A, sample - loc)
            q_log_prob = -0.5 * (zz * zz).sum() - logdet - 0.5 * D * LOG_2PI
            tree = self.unravel(sample)
            p_log_prob = 0.0
            theta = {}
            for k in self.names:
                b = self.bijector[k]
                th = b.forward(tree[k])
                p_log_prob = p_log_prob + (self.prior[k].log_prob(th).sum() + b.fldj(tree[k]).sum())
                theta[k] = th
            ll = self.log_likelihood_fn(theta, outputs, inputs, **kwargs)
            ll = (ll / len(outputs)) * full_data_size
            return q_log_prob - p_log_prob - ll

        return torch.stack([loss_no_batch(samples[s]) for s in range(samples.shape[0])]).mean()

    def value_and_grad_low_rank(self, loc, raw_diag, U, kwargs, outputs, inputs, full_data_size, eps):
        loc = loc.detach().clone().requires_grad_(True)
        raw_diag = raw_diag.detach().clone().requires_grad_(True)
        U = U.detach().clone().requires_grad_(True)
        kw = {k: v.detach().clone().requires_grad_(True) for k, v in kwargs.items()}
        loss = self.loss_low_rank(loc, raw_diag, U, kw, outputs, inputs, full_data_size, eps)
        leaves = [loc, raw_diag, U] + list(kw.values())
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        gk = {k: (g if g is not None else torch.zeros_like(v)) for (k, v), g in zip(kw.items(), grads[3:])}
        return loss.detach(), grads[0], grads[1], grads[2], gk

    def map_neg_log_joint(self, z, outputs, inputs):
        """MAP.neg_log_joint (map.py:27-33) at the flat unconstrained vector z: transform_tree, log_prob_dist of the PRIOR at
        the constrained value (no log-det-Jacobian), likelihood.log_prob(outputs).sum(); returns -(log_prior + log_lik)."""
        tree = self.unravel(z)
        theta = {k: self.bijector[k].forward(tree[k]) for k in self.names}
        log_prior = sum(self.prior[k].log_prob(theta[k]).sum() for k in self.names)
        return -(log_prior + self.log_likelihood_fn(theta, outputs, inputs))

    def mcmc_log_joint(self, z, outputs, inputs):
        """MCMC.log_joint (mcmc.py:30-34): log_prob_dist(approx_normal_prior, z) (prior at b(z) plus fldj, core.py:114-121)
        plus the log-likelihood at the constrained value."""
        tree = self.unravel(z)
        theta = {k: self.bijector[k].forward(tree[k]) for k in self.names}
        p = sum(self.prior[k].log_prob(theta[k]).sum() + self.bijector[k].fldj(tree[k]).sum() for k in self.names)
        return p + self.log_likelihood_fn(theta, outputs, inputs)

    def value_and_grad_chunked(self, loc, raw_scale, kwargs, outputs, X, full_data_size, eps, chunk_rows=1 << 18,
                               n_total=None, dtype=torch.float64):
        """``value_and_grad`` for problems too large for one autograd graph (BASELINE.json's full sizes: c3 10M rows, a c5
        shard of 12.5M rows, c4 at D=2048 / S=256 / N=1M).  Same formulas as ``loss`` (advi.py:80-95), float64 throughout:

          * sampling, q.log_prob and the unconstrained-prior terms are evaluated for all S samples at once (the ``vmap`` of
            advi.py:95) -- they are O(S*D) / O(S*D^2);
          * the log-likelihood (advi.py:90-91) is the user's ``log_likelihood_fn`` itself, called on row blocks of
            (outputs, X) with the S constrained samples stacked on a trailing axis; because the loss is linear in the
            per-sample log-likelihoods, only sum_s ll_s and its cotangents are accumulated over the blocks;
          * the chain rule back to (loc, raw_scale, kwargs) goes through one autograd pass of the small graph.

        It runs wherever ``outputs`` / ``X`` live (``outputs.device``): the -m gpu tests keep the float32 dataset on the
        GPU and let torch do the float64 arithmetic there -- torch is the CHECKER's calculator here, not the product.
        ``n_total``: the global ``len(outputs)`` when (outputs, X) is one row shard of a larger dataset.
        ``dtype=torch.float32`` on CPU tensors is the timed CPU baseline of the full-rank configuration (bench.py).
        Returns loss, dloc, draw_scale, {dkwargs} like ``value_and_grad``."""
        dev = outputs.device
        f64 = lambda t: torch.as_tensor(t).detach().to(device=dev, dtype=dtype)
        loc = f64(loc).clone().requires_grad_(True)
        raw_scale = f64(raw_scale).clone().requires_grad_(True)
        kw = {k: f64(v).clone().requires_grad_(True) for k, v in kwargs.items()}
        eps = f64(eps)
        S, D = eps.shape[0], self.size
        if self.vi_type == "mean_field":
            sigma = self.scale_bijector.forward(raw_scale)
            samples = loc + sigma * eps
            zz = (samples - loc) / sigma
            q = (-0.5 * zz * zz - torch.log(sigma)).sum(-1) - 0.5 * D * LOG_2PI
        else:
            tril = torch.tril(self.scale_bijector.forward(raw_scale))
            samples = loc + eps @ tril.T
            zz = torch.linalg.solve_triangular(tril, (samples - loc).T, upper=False)  # [D, S]
            q = -0.5 * (zz * zz).sum(0) - torch.log(torch.abs(torch.diagonal(tril))).sum() - 0.5 * D * LOG_2PI
        p = torch.zeros(S, dtype=dtype, device=dev)
        theta, o = {}, 0
        for k in self.names:
            n = self.sizes[k]
            zk = samples[:, o : o + n].reshape((S,) + tuple(self.shapes[k]))
            o += n
            b = self.bijector[k]
            th = b.forward(zk)
            p = p + self.prior[k].log_prob(th).reshape(S, -1).sum(-1) + b.fldj(zk).reshape(S, -1).sum(-1)
            theta[k] = th.movedim(0, -1)  # sample axis last: X @ theta[weights] is [rows, S]
        th_leaf = {k: v.detach().clone().requires_grad_(True) for k, v in theta.items()}
        kw_leaf = {k: v.detach().clone().requires_grad_(True) for k, v in kw.items()}
        leaves = list(th_leaf.values()) + list(kw_leaf.values())
        acc = [torch.zeros_like(l) for l in leaves]
        ll_sum = torch.zeros((), dtype=dtype, device=dev)
        N = outputs.shape[0]
        for i in range(0, N, int(chunk_rows)):
            yc = outputs[i : i + chunk_rows].to(dtype)[:, None]
            inputs = {"X": X[i : i + chunk_rows].to(dtype)} if X is not None else None
            ll = self.log_likelihood_fn(th_leaf, yc, inputs, **kw_leaf)  # sum over the block's rows AND the S samples
            gs = torch.autograd.grad(ll, leaves, allow_unused=True)
            for a, g in zip(acc, gs):
                if g is not None:
                    a += g
            ll_sum += ll.detach()
        c = float(full_data_size) / float(n_total if n_total is not None else N)  # advi.py:92
        loss = (q - p).mean() - (c / S) * ll_sum
        surrogate = (q - p).mean() - (c / S) * sum((a * v).sum() for a, v in zip(acc, list(theta.values()) + list(kw.values())))
        grads = torch.autograd.grad(surrogate, [loc, raw_scale] + list(kw.values()), allow_unused=True)
        gk = {k: (g if g is not None else torch.zeros_like(v)) for (k, v), g in zip(kw.items(), grads[2:])}
        return loss.detach(), grads[0], grads[1], gk

    def value_and_grad(self, loc, raw_scale, kwargs, outputs, inputs, full_data_size, eps):
        """jax.value_and_grad(loss_fn) (utils.py:7): returns loss, dloc, draw_scale, {dkwargs}."""
        loc = loc.detach().clone().requires_grad_(True)
        raw_scale = raw_scale.detach().clone().requires_grad_(True)
        kw = {k: v.detach().clone().requires_grad_(True) for k, v in kwargs.items()}
        loss = self.loss(loc, raw_scale, kw, outputs, inputs, full_data_size, eps)
        leaves = [loc, raw_scale] + list(kw.values())
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        gk = {k: (g if g is not None else torch.zeros_like(v)) for (k, v), g in zip(kw.items(), grads[2:])}
        return loss.detach(), grads[0], grads[1], gk


# --------------------------------------------------------------------------- vectorised fast paths (same arithmetic, batched over S)
def logistic_meanfield_value_and_grad(loc, raw_scale, X, y, full_data_size, eps, prior_scale=1.0):
    """Batched restatement for Normal(0, prior_scale) iid prior + Identity bijector + Bernoulli-logits likelihood,
    mean-field with the default Exp scale bijector.  Same formulas as ADVIRef.loss, but the S samples go through
    one X @ theta^T so that it can be timed as the CPU baseline (this is what XLA's vmap of advi.py:95 lowers to)."""
    loc = loc.detach().clone().requires_grad_(True)
    raw_scale = raw_scale.detach().clone().requires_grad_(True)
    sp = torch.nn.functional.softplus
    D = loc.numel()
    sigma = torch.exp(raw_scale)
    z = loc + sigma * eps.to(loc.dtype)  # [S, D]
    zz = (z - loc) / sigma
    q = (-0.5 * zz * zz - torch.log(sigma)).sum(-1) - 0.5 * D * LOG_2PI
    ps = torch.as_tensor(prior_scale, dtype=loc.dtype)
    p = (-0.5 * (z / ps) ** 2 - 0.5 * LOG_2PI - torch.log(ps)).sum(-1)
    logits = X @ z.T  # [N, S]
    ll = (-(1.0 - y)[:, None] * sp(logits) - y[:, None] * sp(-logits)).sum(0)
    ll = ll / y.shape[0] * full_data_size
    loss = (q - p - ll).mean()
    g = torch.autograd.grad(loss, [loc, raw_scale])
    return loss.detach(), g[0], g[1]


def linreg_meanfield_value_and_grad(loc, raw_scale, log_noise_scale, X, y, full_data_size, eps, prior_scale=1.0):
    """Batched restatement for the Gaussian linear likelihood with learnable log_noise_scale (README.md:99-106)."""
    loc = loc.detach().clone().requires_grad_(True)
    raw_scale = raw_scale.detach().clone().requires_grad_(True)
    lns = log_noise_scale.detach().clone().requires_grad_(True)
    D = loc.numel()
    sigma = torch.exp(raw_scale)
    z = loc + sigma * eps.to(loc.dtype)
    zz = (z - loc) / sigma
    q = (-0.5 * zz * zz - torch.log(sigma)).sum(-1) - 0.5 * D * LOG_2PI
    ps = torch.as_tensor(prior_scale, dtype=loc.dtype)
    p = (-0.5 * (z / ps) ** 2 - 0.5 * LOG_2PI - torch.log(ps)).sum(-1)
    mean = X @ z.T
    tau = torch.exp(lns)
    r = (y[:, None] - mean) / tau
    ll = (-0.5 * r * r - 0.5 * LOG_2PI - torch.log(tau)).sum(0)
    ll = ll / y.shape[0] * full_data_size
    loss = (q - p - ll).mean()
    g = torch.autograd.grad(loss, [loc, raw_scale, lns])
    return loss.detach(), g[0], g[1], g[2]

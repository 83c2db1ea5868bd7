"""ADVI -- the drop-in shell around the fused CUDA ELBO engine.

Mirrors bijax's class surface:
  * v1 (README.md:113, the surface north_star names):  ADVI(prior, bijector, log_likelihood_fn, vi_type="mean_field")
  * v2 (bijax/advi.py:26-34):  ADVI(prior, likelihood_fn, prior_constraints, vi_type, rank, ordered_posterior_bijectors)
  * .init(seed[, initializer])                    advi.py:77-78  (+ utils.py:39-55)
  * .loss_fn(params, outputs, inputs, full_data_size, seed, n_samples=1)     advi.py:80-95
  * .apply(params) == .get_params_posterior(params)                          advi.py:97-100

What runs underneath is NOT the reference's op-by-op graph: ``loss_fn`` enqueues one fused step (bjx_elbo_step) that
returns the loss and every gradient; the returned scalar carries a custom autograd node (the jax.custom_vjp analogue)
whose backward just scales the saved gradients.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, Optional, Tuple

import torch

from . import _native as nv
from . import bijectors as tfb
from . import likelihoods as lk
from . import random as brandom
from .engine import ElboEngine
from .spec import compile_latents, fill_in_bijector
from .tree import tree_flatten, tree_unflatten

_POSTERIOR_KEYS = ("posterior", "approx_posterior")


def _scale_key(vi_type: str) -> str:
    return "scale_diag" if vi_type == "mean_field" else "scale_tril"


class _ElboFunction(torch.autograd.Function):
    """custom_vjp: forward computes loss AND all gradients in one launch sequence; backward = cotangent x residuals."""

    @staticmethod
    def forward(ctx, loc, raw, kw, run):
        out = run(loc.detach(), raw.detach(), kw.detach() if kw is not None else None)
        ctx.save_for_backward(out)
        ctx.shapes = (loc.shape, raw.shape, None if kw is None else kw.shape)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        ls, rs, ks = ctx.shapes
        D, P = math.prod(ls), math.prod(rs)
        d_loc = (g * out[1 : 1 + D]).reshape(ls)
        d_raw = (g * out[1 + D : 1 + D + P]).reshape(rs)
        d_kw = (g * out[1 + D + P :]).reshape(ks) if ks is not None else None
        return d_loc, d_raw, d_kw, None


class _FillTriangular(torch.autograd.Function):
    """packed [D(D+1)/2] -> dense lower-triangular [D, D] in tfp.math.fill_triangular order (bjx_fill_triangular)."""

    @staticmethod
    def forward(ctx, packed, D):
        packed = packed.detach().to(torch.float32).contiguous()
        dense = torch.empty((D, D), dtype=torch.float32, device=packed.device)
        with torch.cuda.device(packed.device):
            nv.check(nv.load().bjx_fill_triangular(packed.data_ptr(), D, dense.data_ptr(), nv.current_stream_ptr()))
        ctx.D = D
        return dense

    @staticmethod
    def backward(ctx, g):
        g = g.to(torch.float32).contiguous()
        D = ctx.D
        out = torch.empty(D * (D + 1) // 2, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            nv.check(nv.load().bjx_fill_triangular_vjp(g.data_ptr(), D, out.data_ptr(), nv.current_stream_ptr()))
        return out, None


def fill_triangular(packed: torch.Tensor) -> torch.Tensor:
    """tfp.math.fill_triangular(x) (lower) on the device, differentiable."""
    m = packed.numel()
    D = int((math.sqrt(8 * m + 1) - 1) / 2)
    if D * (D + 1) // 2 != m:
        raise ValueError(f"{m} is not a triangular number")
    return _FillTriangular.apply(packed.reshape(-1), D)


class ADVI:
    def __init__(self, prior: Dict[str, Any], *args, vi_type: str = "mean_field", rank: Optional[int] = None,
                 ordered_posterior_bijectors=None, bijector=None, log_likelihood_fn=None, likelihood_fn=None,
                 prior_constraints=None, device=None, precision: str = "auto", kernel: Optional[str] = None,
                 prepare_dataset: bool = True, packed_tril: bool = False):
        # positional forms: v1 (prior, bijector, log_likelihood_fn[, vi_type]); v2 (prior, likelihood_fn[, prior_constraints[, vi_type]])
        args = list(args)
        if args:
            if isinstance(args[0], dict):
                bijector = args.pop(0)
                if args:
                    log_likelihood_fn = args.pop(0)
            else:
                likelihood_fn = args.pop(0)
                if args and (isinstance(args[0], dict) or args[0] is None):
                    prior_constraints = args.pop(0)
            if args and isinstance(args[0], str):
                vi_type = args.pop(0)
            if args:
                raise TypeError("too many positional arguments")
        # which params key init() returns: the v2 signature is the reference's current source (advi.py:78,
        # {"approx_posterior": ...}); the v1 signature is the README / notebook surface (params["posterior"])
        self._posterior_key = "approx_posterior" if (bijector is None and log_likelihood_fn is None and
                                                      (likelihood_fn is not None or prior_constraints is not None)) else "posterior"
        constraints = bijector if bijector is not None else prior_constraints
        if constraints is None:
            constraints = {}
        likelihood = log_likelihood_fn if log_likelihood_fn is not None else likelihood_fn
        if likelihood is None:
            raise TypeError("a likelihood is required (log_likelihood_fn= / likelihood_fn=)")

        self.prior = prior
        self.prior_constraints = fill_in_bijector(constraints, self.prior)  # core.py:107-111
        self.bijector = self.prior_constraints
        assert self.prior_constraints.keys() == self.prior.keys(), (
            "The keys in `prior` and `prior_constraints` must be the same."
        )
        assert (vi_type == "low_rank") == (rank is not None), "`rank` must be specified only if `vi_type` is `low_rank`."
        if vi_type not in ("mean_field", "full_rank", "low_rank"):
            raise ValueError(f"unknown vi_type {vi_type!r}")
        if not isinstance(likelihood, lk.Likelihood):
            raise NotImplementedError(
                "arbitrary python log_likelihood_fn callables are not on the fused CUDA path; pass one of "
                "bijax_b200.likelihoods.{BernoulliLogits, GaussianLinear, BernoulliProbs}"
            )
        self.likelihood_fn = self.log_likelihood_fn = likelihood
        self.vi_type = vi_type
        self.rank = rank
        # default parameter bijectors: [Identity, Exp] mean field (core.py:180-181), [Identity, Identity] full rank (core.py:205-206)
        if ordered_posterior_bijectors is None:
            # core.py:180-181 (mean field), 191-192 (low rank: [Identity, Exp, Identity]), 205-206 (full rank)
            ordered_posterior_bijectors = [tfb.Identity(), tfb.Identity() if vi_type == "full_rank" else tfb.Exp()]
            if vi_type == "low_rank":
                ordered_posterior_bijectors.append(tfb.Identity())
        if vi_type == "low_rank" and len(ordered_posterior_bijectors) > 2 and type(ordered_posterior_bijectors[2]).__name__ != "Identity":
            raise NotImplementedError("only Identity is supported on scale_perturb_factor")
        if type(ordered_posterior_bijectors[0]).__name__ != "Identity":
            raise NotImplementedError("only Identity is supported on the location parameter")
        self.posterior_params_bijector = ordered_posterior_bijectors
        self.latents = compile_latents(self.prior, self.prior_constraints)
        self.size = self.latents.size
        self.unravel_fn = self.latents.unravel
        self.device = device
        self.precision = precision
        self.prepare_dataset = prepare_dataset  # keep X as split-bf16 operand planes, made once per dataset (engine.py)
        # extension (north_star item 2): full-rank scale_tril held as a packed D(D+1)/2 vector, L = fill_triangular(packed);
        # the reference's own parameterisation is the dense D x D raw matrix (core.py:205-209), which stays the default
        self.packed_tril = bool(packed_tril)
        if self.packed_tril and vi_type != "full_rank":
            raise ValueError("packed_tril applies to vi_type='full_rank' only")
        self.kernel = kernel  # None = the library picks the streaming kernel; a name from _native.KERNEL_NAMES forces one
        self._engines: Dict[Tuple, ElboEngine] = {}
        self._group = None  # torch.distributed process group for row-sharded data (see parallel.py)
        self._global_len: Optional[int] = None

    # ------------------------------------------------------------------------------------------------------------------
    def init(self, seed, initializer: Optional[Callable] = None):
        """advi.py:77-78: raw params = initializer(seed, (P,)), default normal(stddev=1), unraveled in (loc, scale) order.
        The top-level key is the reference's "approx_posterior" (advi.py:78) for the v2 constructor form and "posterior" for
        the v1 (README / notebook) form; loss_fn / apply / train accept either."""
        D = self.size
        if self.vi_type == "low_rank":
            # leaves in constructor-argument order: loc [D], scale_diag [D], scale_perturb_factor [D, rank]  (utils.py:53-55)
            r = int(self.rank)
            flat = brandom.normal(seed, (2 * D + D * r,)) if initializer is None else initializer(seed, (2 * D + D * r,))
            return {self._posterior_key: {"loc": flat[:D].clone(), "scale_diag": flat[D : 2 * D].clone(),
                                  "scale_perturb_factor": flat[2 * D :].reshape(D, r).clone()}}
        P = D if self.vi_type == "mean_field" else (D * (D + 1) // 2 if self.packed_tril else D * D)
        flat = brandom.normal(seed, (D + P,)) if initializer is None else initializer(seed, (D + P,))
        scale = flat[D:].reshape((P,) if (self.vi_type == "mean_field" or self.packed_tril) else (D, D))
        return {self._posterior_key: {"loc": flat[:D].clone(), _scale_key(self.vi_type): scale.clone()}}

    # ------------------------------------------------------------------------------------------------------------------
    def _split_params(self, params):
        key = next((k for k in _POSTERIOR_KEYS if k in params), None)
        if key is None:
            raise KeyError(f"params needs one of {_POSTERIOR_KEYS}")
        post = params[key]
        if isinstance(post, dict):
            loc = post["loc"]
            raw = post.get("scale_diag", post.get("scale_tril", post.get("scale")))
        else:  # an object with attributes (e.g. a tfd distribution holding raw leaves, as the reference's init returns)
            loc = post.loc
            raw = getattr(post, "scale_diag", None)
            if raw is None:
                raw = getattr(post, "scale_tril")
        kwargs = {k: v for k, v in params.items() if k != key}
        if self.vi_type == "low_rank":
            U = post["scale_perturb_factor"] if isinstance(post, dict) else post.scale_perturb_factor
            # the engine takes [raw scale_diag | scale_perturb_factor] as one vector; torch.cat keeps both differentiable
            raw = torch.cat([raw.reshape(-1).to(torch.float32), U.reshape(-1).to(torch.float32)])
        if self.packed_tril:
            raw = fill_triangular(raw)  # differentiable: gradients come back in the packed layout
        return loc, raw, kwargs

    def _engine_for(self, kwargs) -> Tuple[ElboEngine, list]:
        names = sorted(kwargs)
        leaves = []
        layout = []
        for n in names:
            ls, _ = tree_flatten(kwargs[n])
            numel = sum(int(l.numel()) for l in ls)
            layout.append((n, numel))
            leaves.extend(ls)
        sig = tuple(layout)
        if sig not in self._engines:
            self._engines[sig] = ElboEngine(self.latents, self.likelihood_fn, self.vi_type, self.posterior_params_bijector[1],
                                            layout, device=self.device, precision=self.precision, kernel=self.kernel,
                                            prepare_dataset=self.prepare_dataset, rank=self.rank)
        return self._engines[sig], leaves

    def distributed(self, group=None, global_len: Optional[int] = None, shard: str = "rows", p2p: bool = False):
        """Data parallelism with one all-reduce of the packed [loss | grads] buffer per step.
        ``shard="rows"``: ``outputs`` / ``inputs`` passed to loss_fn are this rank's row shard.  ``shard="samples"`` (for
        small N): every rank passes the full data and evaluates its block of the ``n_samples`` Monte-Carlo samples."""
        import torch.distributed as dist

        if shard not in ("rows", "samples"):
            raise ValueError("shard must be 'rows' or 'samples'")
        self._group = group if group is not None else dist.group.WORLD
        self._global_len = global_len
        self._shard = shard
        # p2p=True: the all-reduce is one kernel over NVLink peer memory (parallel.P2PAllReduce) instead of an NCCL call
        self._p2p = bool(p2p)
        self._p2p_op = None
        return self

    def _shard_context(self, engine, n_local: int, full_data_size: float):
        """(ll_scale, prior_weight, exchange) of a row-sharded model with the peer-memory exchange."""
        import torch.distributed as dist

        if self._global_len is None:
            t = torch.tensor([int(n_local)], dtype=torch.int64, device=engine.device)
            dist.all_reduce(t, group=self._group)
            self._global_len = int(t.item())
        n_out = 1 + engine.D + engine.P + engine.K
        if self._p2p_op is None or self._p2p_op.n != n_out:
            from .parallel import P2PAllReduce

            self._p2p_op = P2PAllReduce(n_out, self._group, engine.device)
        prior_weight = 1.0 if dist.get_rank(self._group) == 0 else 0.0
        return float(full_data_size) / float(self._global_len), prior_weight, self._p2p_op

    def _run(self, engine, seed, outputs, inputs, full_data_size, n_samples):
        X = None
        if engine.lik_id != nv.LIK_BERNOULLI_PROBS:
            X = inputs[engine.x_name] if isinstance(inputs, dict) else inputs
        n_local = len(outputs)
        prior_weight = 1.0
        n_global = n_local
        sample_shard = None
        if self._group is not None and getattr(self, "_shard", "rows") == "samples":
            import torch.distributed as dist

            from . import parallel as par

            rank, world = dist.get_rank(self._group), dist.get_world_size(self._group)
            s0, s1 = par.shard_samples(int(n_samples), rank, world)
            sample_shard = (s0, int(n_samples), par.entropy_weight(rank))
            n_samples = s1 - s0
        elif self._group is not None:
            import torch.distributed as dist

            if self._global_len is None:
                t = torch.tensor([n_local], dtype=torch.int64, device=engine.device)
                dist.all_reduce(t, group=self._group)
                self._global_len = int(t.item())
            n_global = self._global_len
            prior_weight = 1.0 if dist.get_rank(self._group) == 0 else 0.0
        ll_scale = float(full_data_size) / float(n_global)  # advi.py:92

        def run(loc, raw, kw):
            if sample_shard is not None and n_samples == 0:  # more ranks than samples: this rank only joins the reduction
                out = torch.zeros(1 + engine.D + engine.P + engine.K, dtype=torch.float32, device=engine.device)
            else:
                out = engine.step(seed, loc, raw, kw, X, outputs, ll_scale, n_samples, prior_weight, sample_shard=sample_shard)
            if self._group is not None:
                import torch.distributed as dist

                if getattr(self, "_p2p", False):
                    if self._p2p_op is None or self._p2p_op.n != out.numel():
                        from .parallel import P2PAllReduce

                        self._p2p_op = P2PAllReduce(out.numel(), self._group, engine.device)
                    self._p2p_op(out)
                else:
                    dist.all_reduce(out, group=self._group)
            return out

        return run

    def loss_fn(self, params, outputs, inputs, full_data_size, seed, n_samples=1):
        """Negative Monte-Carlo ELBO (advi.py:80-95); differentiable w.r.t. every tensor leaf of ``params``."""
        loc, raw, kwargs = self._split_params(params)
        engine, kw_leaves = self._engine_for(kwargs)
        kw = torch.cat([l.reshape(-1).to(torch.float32) for l in kw_leaves]) if kw_leaves else None
        run = self._run(engine, seed, outputs, inputs, full_data_size, n_samples)
        return _ElboFunction.apply(loc, raw, kw, run)

    def minibatch_loss_fn(self, outputs, inputs, batch_size, n_samples=1, full_data_size=None, copy_free=None):
        """``loss(params, seed)`` on a minibatch of ``batch_size`` rows drawn on the device from ``seed``
        (jax.random.choice semantics) with the full-data scaling of advi.py:92; see minibatch.MinibatchLoss."""
        from .minibatch import MinibatchLoss

        return MinibatchLoss(self, outputs, inputs, batch_size, n_samples, full_data_size, copy_free)

    def value_and_grad(self, params, outputs, inputs, full_data_size, seed, n_samples=1):
        """jax.value_and_grad(loss_fn)(params, ...) in one fused step: (loss, grads with the structure of params)."""
        loc, raw, kwargs = self._split_params(params)
        engine, kw_leaves = self._engine_for(kwargs)
        kw = torch.cat([l.detach().reshape(-1).to(torch.float32) for l in kw_leaves]) if kw_leaves else None
        out = self._run(engine, seed, outputs, inputs, full_data_size, n_samples)(loc.detach(), raw.detach(), kw)
        D, P = engine.D, engine.P
        key = next(k for k in _POSTERIOR_KEYS if k in params)
        d_raw = out[1 + D : 1 + D + P].reshape(raw.shape)
        if self.packed_tril:  # gather the dense gradient back into the packed layout (VJP of fill_triangular)
            packed = torch.empty(D * (D + 1) // 2, dtype=torch.float32, device=out.device)
            with torch.cuda.device(out.device):
                nv.check(nv.load().bjx_fill_triangular_vjp(d_raw.contiguous().data_ptr(), D, packed.data_ptr(), nv.current_stream_ptr()))
            d_raw = packed
        if self.vi_type == "low_rank":
            r = int(self.rank)
            grads: Dict[str, Any] = {key: {"loc": out[1 : 1 + D].reshape(loc.shape), "scale_diag": d_raw[:D].clone(),
                                           "scale_perturb_factor": d_raw[D:].reshape(D, r).clone()}}
        else:
            grads = {key: {"loc": out[1 : 1 + D].reshape(loc.shape), _scale_key(self.vi_type): d_raw}}
        o = 1 + D + P
        for n in sorted(kwargs):
            ls, td = tree_flatten(kwargs[n])
            gl = []
            for l in ls:
                gl.append(out[o : o + l.numel()].reshape(l.shape))
                o += l.numel()
            grads[n] = tree_unflatten(td, gl)
        return out[0], grads

    # ------------------------------------------------------------------------------------------------------------------
    def get_params_posterior(self, params):
        from .core import Posterior

        loc, raw, kwargs = self._split_params(params)
        engine, _ = self._engine_for(kwargs)
        return Posterior(self, engine, loc.detach(), raw.detach())

    apply = get_params_posterior

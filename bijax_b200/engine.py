"""Host-side driver of the fused ELBO+gradient step: owns the coordinate table, the workspace and the argument block, and
calls ``bjx_elbo_step`` through the C ABI.  One engine per (model, device)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nv
from . import likelihoods as lk
from . import random as brandom
from .dataset import PreparedX
from .spec import LatentSpec

_SCALE_IDS = {"Exp": nv.SCALE_EXP, "Softplus": nv.SCALE_SOFTPLUS, "Identity": nv.SCALE_IDENTITY}
_PRECISIONS = {"auto": nv.PREC_AUTO, "fp32": nv.PREC_FP32, "bf16x3": nv.PREC_BF16X3, "bf16x2": nv.PREC_BF16X2}


class ElboEngine:
    def __init__(
        self,
        latents: LatentSpec,
        likelihood,
        vi_type: str = "mean_field",
        scale_bijector=None,
        kwarg_layout: Sequence[Tuple[str, int]] = (),
        device=None,
        precision: str = "auto",
        kernel: Optional[str] = None,
        prepare_dataset: bool = True,
        rank: Optional[int] = None,
    ):
        nv.require_cuda()
        self.lib = nv.load()
        self.latents = latents
        self.likelihood = likelihood
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if vi_type not in ("mean_field", "full_rank", "low_rank"):
            raise ValueError(f"unknown vi_type {vi_type!r}")
        self.vi_type = {"mean_field": nv.VI_MEAN_FIELD, "full_rank": nv.VI_FULL_RANK, "low_rank": nv.VI_LOW_RANK}[vi_type]
        self.rank = int(rank) if vi_type == "low_rank" else 0
        if vi_type == "low_rank" and not 1 <= self.rank <= 64:
            raise NotImplementedError("the fused low-rank family supports 1 <= rank <= 64")
        # default scale bijectors: Exp on scale_diag (core.py:180-181, 191-192), Identity on scale_tril (core.py:205-206)
        sb = type(scale_bijector).__name__ if scale_bijector is not None else ("Identity" if vi_type == "full_rank" else "Exp")
        if sb not in _SCALE_IDS:
            raise NotImplementedError(f"scale bijector {sb} is not one of Exp / Softplus / Identity")
        self.scale_bijector = _SCALE_IDS[sb]
        self.D = latents.size
        self.P = {nv.VI_MEAN_FIELD: self.D, nv.VI_FULL_RANK: self.D * self.D, nv.VI_LOW_RANK: self.D * (1 + self.rank)}[self.vi_type]
        self.kwarg_layout = list(kwarg_layout)  # [(name, numel)] in sorted-key order
        self.K = sum(n for _, n in self.kwarg_layout)
        self.precision = _PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        # tests / benchmarks may force one streaming kernel by name (nv.KERNEL_NAMES); None = the library's own choice
        self.kernel_override = 0 if kernel is None else nv.KERNEL_IDS[kernel] + 1
        self.coords = torch.from_numpy(latents.coords.view(np.uint8).copy()).to(self.device)
        self._ws: Optional[torch.Tensor] = None
        # the dataset X pre-split into bf16 operand planes for the two-pass tcgen05 GEMM path, made once per dataset
        # (X is a constant of the training loop) and reused while the same, unmodified tensor is passed in
        self.prepare_dataset = prepare_dataset
        self._planes: Optional[torch.Tensor] = None
        self._planes_key = None
        self._planes_storage = None  # strong reference to the storage of the tensor the planes were made from
        self._planes_misses = 0
        self._last_raw = None
        self._point_key: Optional[torch.Tensor] = None
        self._resolve_likelihood()

    # ------------------------------------------------------------------------------------------------------------------
    def _resolve_likelihood(self):
        L = self.likelihood
        self.noise_src, self.noise_index, self.noise_value = nv.NOISE_CONST, 0, 1.0
        self.x_name = getattr(L, "X", None)
        if isinstance(L, lk.BernoulliProbs):
            off, n = self.latents.offset_of(L.probs)
            if n != 1:
                raise ValueError("BernoulliProbs needs a scalar latent")
            self.lik_id, self.w_offset, self.Dw = nv.LIK_BERNOULLI_PROBS, off, 0
        elif isinstance(L, (lk.BernoulliLogits, lk.GaussianLinear)):
            off, n = self.latents.offset_of(L.weights)
            self.w_offset, self.Dw = off, n
            self.lik_id = nv.LIK_BERNOULLI_LOGITS if isinstance(L, lk.BernoulliLogits) else nv.LIK_GAUSSIAN
            if isinstance(L, lk.GaussianLinear):
                if L.noise_scale is not None:
                    self.noise_src, self.noise_value = nv.NOISE_CONST, float(L.noise_scale)
                elif L.noise_latent is not None:
                    o, n2 = self.latents.offset_of(L.noise_latent)
                    if n2 != 1:
                        raise ValueError("noise_latent must be a scalar latent")
                    self.noise_src, self.noise_index = nv.NOISE_LATENT, o
                else:
                    pos, found = 0, False
                    for name, numel in self.kwarg_layout:
                        if name == L.log_noise_scale:
                            if numel != 1:
                                raise ValueError("log_noise_scale must be a scalar parameter")
                            self.noise_src, self.noise_index, found = nv.NOISE_KWARG, pos, True
                        pos += numel
                    if not found:
                        raise KeyError(f"params has no trainable entry {L.log_noise_scale!r} for the Gaussian noise scale")
        else:
            raise NotImplementedError(
                "only the recognised likelihood families (bijax_b200.likelihoods) run on the fused CUDA path"
            )

    # ------------------------------------------------------------------------------------------------------------------
    def _args(self, S, N, ldx, ll_scale, prior_weight, layout) -> nv.BjxElboArgs:
        a = nv.BjxElboArgs()
        a.S, a.N, a.D, a.Dw, a.w_offset, a.K, a.ldx = int(S), int(N), self.D, self.Dw, self.w_offset, self.K, int(ldx)
        a.vi_type, a.scale_bijector, a.likelihood = self.vi_type, self.scale_bijector, self.lik_id
        a.noise_src, a.noise_index, a.noise_value = self.noise_src, self.noise_index, self.noise_value
        a.threefry_layout = layout
        a.precision = self.precision
        a.rank = self.rank
        a.kernel_override = self.kernel_override
        a.ll_scale, a.prior_weight = float(ll_scale), float(prior_weight)
        a.coords = self.coords.data_ptr()
        return a

    def _workspace(self, a: nv.BjxElboArgs) -> torch.Tensor:
        need = int(self.lib.bjx_elbo_workspace_bytes(C.byref(a)))
        if need == 0:
            raise nv.BjxError(nv.ERR_INVALID_ARGUMENT, self.lib.bjx_last_error().decode())
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _f32(t: torch.Tensor, device) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        if t.device != device or t.dtype != torch.float32:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("inputs must be CUDA float32 on the engine's device before CUDA-graph capture")
            t = t.to(device=device, dtype=torch.float32)
        return t.contiguous()

    def _data(self, X, y, indexed: bool = False):
        """Validated (X, y, N, row pitch).  ``indexed``: the rows are X[row_index[r]] -- X is the full dataset, y the batch."""
        y = self._f32(y, self.device).reshape(-1)
        N = y.numel()
        if self.lik_id == nv.LIK_BERNOULLI_PROBS:
            return None, y, N, 0
        if isinstance(X, PreparedX):  # the dataset as planes only: nothing to convert, the kernels never read fp32 X
            X._require_complete()
            if X.device != self.device or (X.shape[0] != N and not indexed) or X.shape[1] != self.Dw:
                raise ValueError(f"PreparedX must be [N={N}, D={self.Dw}] on {self.device}; got {X.shape} on {X.device}")
            return X, y, N, self.Dw
        if not isinstance(X, torch.Tensor):
            X = torch.as_tensor(X)
        if X.device != self.device or X.dtype != torch.float32:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("X must be CUDA float32 on the engine's device before CUDA-graph capture")
            X = X.to(device=self.device, dtype=torch.float32)
        if X.dim() == 2 and X.stride(1) != 1:
            X = X.contiguous()  # rows may be pitched (ldx > Dw, e.g. a column slice of a wider matrix), columns must be dense
        if X.dim() != 2 or (X.shape[0] != N and not indexed) or X.shape[1] != self.Dw:
            raise ValueError(f"X must be [N={N}, D={self.Dw}] (rows = data, row-major); got {tuple(X.shape)}")
        return X, y, N, X.stride(0)

    @staticmethod
    def _check_row_index(row_index: torch.Tensor, N: int) -> int:
        if not (row_index.is_cuda and row_index.dtype == torch.int32 and row_index.is_contiguous() and row_index.numel() == N):
            raise ValueError(f"row_index must be a contiguous CUDA int32 tensor with one entry per row of the batch ({N})")
        return row_index.data_ptr()

    def supports_row_index(self, S, N, planes: bool = False) -> bool:
        """Can a step of this shape read its rows through a row index?  ``planes=False``: out of raw fp32 X (register-converting
        kernel, Dw in {128, 256, 384, 512}); ``planes=True``: out of the full dataset's prepared planes with 16-byte cp.async copies
        (any Dw in [16, 512])."""
        if self.Dw < 1:
            return False
        a = self._args(S, N, max(self.Dw, 1), 1.0, 1.0, nv.LAYOUT_LEGACY)
        if planes:
            a.x_planes, a.N_full = 1, max(int(N), 1)  # any non-NULL value: only the route is being asked about
        return self.lib.bjx_elbo_supports_row_index(C.byref(a)) == 1

    def wants_planes(self, S, N) -> bool:
        """True if a step of this shape reads the dataset as prepared split-bf16 planes (TMA-fed tcgen05 kernels)."""
        if not self.prepare_dataset or self.Dw < 1 or self.precision == nv.PREC_BF16X3:
            return False
        a = self._args(S, N, max(self.Dw, 1), 1.0, 1.0, nv.LAYOUT_LEGACY)
        return self.lib.bjx_elbo_kernel_choice(C.byref(a)) in (nv.KERNEL_IDS["tcgen05_gemm_two_pass"], nv.KERNEL_IDS["tcgen05_bf16_split"])

    def _x_planes(self, a: nv.BjxElboArgs, X: torch.Tensor) -> Optional[int]:
        """Device pointer of the prepared operand planes of X if this step takes the GEMM path, else None."""
        if isinstance(X, PreparedX):
            if self.lib.bjx_elbo_kernel_choice(C.byref(a)) not in (nv.KERNEL_IDS["tcgen05_gemm_two_pass"], nv.KERNEL_IDS["tcgen05_bf16_split"]) \
                    or self.precision == nv.PREC_BF16X3:
                raise TypeError("this shape runs on a kernel that reads fp32 X; pass the tensor itself instead of a PreparedX")
            return X.planes.data_ptr()
        if not self.prepare_dataset or X is None:
            return None
        if self.lib.bjx_elbo_kernel_choice(C.byref(a)) not in (nv.KERNEL_IDS["tcgen05_gemm_two_pass"], nv.KERNEL_IDS["tcgen05_bf16_split"]):
            return None
        # identity of the dataset the planes were made from: the engine HOLDS the tensor's storage, so the allocator cannot
        # hand the same address to another tensor while the planes live (a caller that passes y[idx], X[idx] temporaries,
        # the reference's minibatch pattern, would otherwise hit the cache with a recycled address and version 0); views
        # of one storage (X, X[:n]) share its version counter and compare by offset / shape / strides
        storage = X.untyped_storage()
        key = (storage._cdata, X.data_ptr(), X._version, tuple(X.shape), tuple(X.stride()))
        if self._planes_misses > 2 and self._last_raw is not None and key == self._last_raw[1]:
            self._planes_misses = 0  # the same tensor came back: it is a dataset after all
        if self._planes_key != key:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("dataset planes must be prepared before CUDA-graph capture (engine.prepare)")
            # a caller that passes a fresh minibatch every step (advi.py:92 scaling) would pay a preparation per step:
            # after two consecutive misses the engine stays on the kernels that read raw fp32 X until a tensor repeats
            self._planes_misses += 1
            if self._planes_misses > 2:
                self._planes, self._planes_key, self._planes_storage = None, None, None
                self._last_raw = (storage, key)  # held, so that "the same tensor came back" cannot be a recycled address
                return None
            self._planes, self._planes_storage = None, None  # release the old planes before allocating new ones
            n = int(self.lib.bjx_x_planes_bytes(int(a.N), int(a.Dw)))
            try:
                self._planes = torch.empty(n, dtype=torch.uint8, device=self.device)
            except torch.cuda.OutOfMemoryError:
                # no room for a second copy of the dataset: stay on the kernels that read raw fp32 X -- a slower path
                # (DESIGN.md section 3.1), so say so instead of flipping silently
                import warnings

                warnings.warn(f"bijax_b200: no room for the {n / 1e9:.1f} GB of prepared operand planes of X; this engine now "
                              "reads raw fp32 X every step (register-converting kernel, about 0.6 of the HBM roofline "
                              "instead of 0.8-0.9).  Free memory or pass prepare_dataset=False to silence this.",
                              RuntimeWarning, stacklevel=3)
                self.prepare_dataset = False
                self._planes, self._planes_key, self._planes_storage = None, None, None
                return None
            with torch.cuda.device(self.device):
                nv.check(self.lib.bjx_prepare_x_planes(X.data_ptr(), int(a.N), int(a.Dw), int(a.ldx), self._planes.data_ptr(),
                                                       nv.current_stream_ptr()))
            self._planes_key, self._planes_storage = key, storage
            self._last_raw = None
        else:
            self._planes_misses = 0
        return self._planes.data_ptr()

    @staticmethod
    def _x_ptr(X):
        return None if X is None or isinstance(X, PreparedX) else X.data_ptr()

    def release_planes(self) -> None:
        """Drop the prepared operand planes (and the hold on their source tensor's storage)."""
        self._planes, self._planes_key, self._planes_storage, self._last_raw = None, None, None, None
        self._planes_misses = 0

    def prepare(self, X, y, n_samples) -> bool:
        """Prepare the dataset's operand planes now (otherwise the first step does it).  Returns True if this engine keeps
        prepared planes for this problem, False if the selected kernel reads raw fp32 X."""
        X, y, N, ldx = self._data(X, y)
        a = self._args(n_samples, N, ldx if ldx else max(self.Dw, 1), 1.0, 1.0, nv.LAYOUT_LEGACY)
        return self._x_planes(a, X) is not None

    def kernel_choice(self, S, N) -> str:
        a = self._args(S, N, max(self.Dw, 1), 1.0, 1.0, nv.LAYOUT_LEGACY)
        return nv.KERNEL_NAMES[nv.check(self.lib.bjx_elbo_kernel_choice(C.byref(a)))]

    def step(self, key, loc, raw_scale, kwargs_vec, X, y, ll_scale, n_samples, prior_weight=1.0, out=None, eps_out=None,
             partitionable=None, sample_shard=None, row_index=None, planes=None, n_full=None) -> torch.Tensor:
        """Enqueue one ELBO+grad step; returns the packed device tensor [loss | d loc | d raw_scale | d kwargs].
        ``sample_shard = (s_offset, S_total, entropy_weight)`` evaluates samples [s_offset, s_offset + n_samples) of a
        step with S_total samples (shard the Monte-Carlo samples when N is small); the packed
        outputs of all shards sum to the unsharded step.
        ``row_index`` (int32 [len(y)]): row r of the batch is X[row_index[r]] -- X is the full dataset, read through the
        index (BjxElboArgs.row_index; only where ``supports_row_index``).  ``planes``: operand planes of X maintained by
        the caller (the minibatch sampler gathers plane rows itself) instead of the engine's own cache; together with
        ``row_index`` they are the FULL dataset's planes (``n_full`` rows) and the kernel gathers the batch's rows out of them."""
        dev = self.device
        key = brandom._check_key(key)
        loc = self._f32(loc, dev).reshape(-1)
        raw = self._f32(raw_scale, dev).reshape(-1)
        if loc.numel() != self.D or raw.numel() != self.P:
            raise ValueError(f"loc must have {self.D} and raw scale {self.P} elements")
        X, y, N, ldx = self._data(X, y, indexed=row_index is not None)
        a = self._args(n_samples, N, ldx if ldx else max(self.Dw, 1), ll_scale, prior_weight, brandom._layout(partitionable))
        a.key, a.loc, a.raw_scale = key.data_ptr(), loc.data_ptr(), raw.data_ptr()
        kw = None
        if self.K:
            kw = self._f32(kwargs_vec, dev).reshape(-1)
            if kw.numel() != self.K:
                raise ValueError(f"expected {self.K} kwarg values")
            a.kwargs = kw.data_ptr()
        a.X = self._x_ptr(X)
        a.y = y.data_ptr()
        if row_index is not None:
            a.row_index = self._check_row_index(row_index, N)
        if out is None:
            out = torch.empty(1 + self.D + self.P + self.K, dtype=torch.float32, device=dev)
        a.out = out.data_ptr()
        if eps_out is not None:
            a.eps_out = eps_out.data_ptr()
        if sample_shard is not None:
            a.s_offset, a.S_total, a.entropy_weight = int(sample_shard[0]), int(sample_shard[1]), float(sample_shard[2])
        if row_index is None:
            a.x_planes = planes.data_ptr() if planes is not None else self._x_planes(a, X)
        elif planes is not None or isinstance(X, PreparedX):
            a.x_planes = planes.data_ptr() if planes is not None else X.planes.data_ptr()
            a.N_full = int(n_full if n_full is not None else X.shape[0])
        ws = self._workspace(a)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(dev):
            nv.check(self.lib.bjx_elbo_step(C.byref(a), nv.current_stream_ptr()))
        return out

    def step_host(self, key_host: np.ndarray, params_host: torch.Tensor, out_host: torch.Tensor, X, y, ll_scale, n_samples,
                  prior_weight=1.0, partitionable=None) -> torch.Tensor:
        """The same step through HOST buffers (pinned): key/params copied in, packed result copied out, stream synchronised.
        X / y stay device-resident, as they do for the reference (closure constants of the jitted scan, utils.py:22-31)."""
        X, y, N, ldx = self._data(X, y)
        a = self._args(n_samples, N, ldx if ldx else max(self.Dw, 1), ll_scale, prior_weight, brandom._layout(partitionable))
        a.X = self._x_ptr(X)
        a.y = y.data_ptr()
        a.x_planes = self._x_planes(a, X)
        ws = self._workspace(a)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        key_host = np.ascontiguousarray(key_host, dtype=np.uint32)
        assert params_host.dtype == torch.float32 and params_host.numel() == self.D + self.P + self.K
        assert out_host.dtype == torch.float32 and out_host.numel() == 1 + self.D + self.P + self.K
        with torch.cuda.device(self.device):
            nv.check(
                self.lib.bjx_elbo_step_host(C.byref(a), key_host.ctypes.data, params_host.data_ptr(), out_host.data_ptr(), nv.current_stream_ptr())
            )
        return out_host

    def stepper(self, key, loc, raw_scale, kwargs_vec, X, y, ll_scale, n_samples, out, prior_weight=1.0, partitionable=None):
        """``step`` with fixed device buffers resolved ONCE: returns ``call()`` that enqueues one ELBO+grad step reading the
        CURRENT contents of ``key`` / ``loc`` / ``raw_scale`` / ``kwargs_vec`` and writing ``out`` -- one C call per step (what a
        host loop that refreshes those buffers with small copies pays)."""
        dev = self.device
        key = brandom._check_key(key)
        for t in (loc, raw_scale, out):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        X, y, N, ldx = self._data(X, y)
        a = self._args(n_samples, N, ldx if ldx else max(self.Dw, 1), ll_scale, prior_weight, brandom._layout(partitionable))
        a.key, a.loc, a.raw_scale = key.data_ptr(), loc.data_ptr(), raw_scale.data_ptr()
        if self.K:
            a.kwargs = kwargs_vec.data_ptr()
        a.X = self._x_ptr(X)
        a.y = y.data_ptr()
        a.out = out.data_ptr()
        a.x_planes = self._x_planes(a, X)
        ws = self._workspace(a)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        fn, ref = self.lib.bjx_elbo_step, C.byref(a)
        keep = (key, loc, raw_scale, kwargs_vec, X, y, out, ws, a)

        def call():
            assert keep is not None
            with torch.cuda.device(dev):
                nv.check(fn(ref, nv.current_stream_ptr()))
            return out

        return call

    def host_stepper(self, X, y, ll_scale, n_samples, prior_weight=1.0, partitionable=None):
        """``step_host`` with everything that does not change between steps resolved ONCE (argument block, kernel choice,
        planes, workspace): returns ``call(key_host uint32[2] ndarray, params_host pinned f32, out_host pinned f32)`` whose only
        work per step is the one C call ``bjx_elbo_step_host`` -- what a host-driven training loop pays per step."""
        X, y, N, ldx = self._data(X, y)
        a = self._args(n_samples, N, ldx if ldx else max(self.Dw, 1), ll_scale, prior_weight, brandom._layout(partitionable))
        a.X = self._x_ptr(X)
        a.y = y.data_ptr()
        a.x_planes = self._x_planes(a, X)
        ws = self._workspace(a)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        n_params = self.D + self.P + self.K
        fn, ref, dev = self.lib.bjx_elbo_step_host, C.byref(a), self.device
        keep = (X, y, ws, a)  # everything the argument block points at stays alive with the closure

        def call(key_host: np.ndarray, params_host: torch.Tensor, out_host: torch.Tensor) -> torch.Tensor:
            assert key_host.dtype == np.uint32 and key_host.size == 2 and key_host.flags.c_contiguous
            assert params_host.numel() == n_params and out_host.numel() == 1 + n_params and keep is not None
            with torch.cuda.device(dev):
                nv.check(fn(ref, key_host.ctypes.data, params_host.data_ptr(), out_host.data_ptr(), nv.current_stream_ptr()))
            return out_host

        return call

    def point_step(self, z, X, y, ll_scale, mode, out=None, kwargs_vec=None) -> torch.Tensor:
        """Point evaluation at the unconstrained value z [D] (BjxElboArgs.point_mode): packed [value | d value / d z | 0... |
        d kwargs], value = -(log prior + ll_scale * log likelihood) with or without the bijectors' log-det-Jacobian
        (MCMC.log_joint, mcmc.py:30-34 / MAP.neg_log_joint, map.py:27-33), or -ll_scale * log likelihood at the CONSTRAINED
        value z itself (POINT_LOGLIK: the user-facing log_likelihood_fn, README.md:81-106)."""
        if self.vi_type != nv.VI_MEAN_FIELD or (self.K and kwargs_vec is None):
            raise NotImplementedError("point evaluation runs on a mean-field engine; trainable kwargs must be passed as kwargs_vec")
        dev = self.device
        z = self._f32(z, dev).reshape(-1)
        if z.numel() != self.D:
            raise ValueError(f"z must have {self.D} elements")
        X, y, N, ldx = self._data(X, y)
        a = self._args(1, N, ldx if ldx else max(self.Dw, 1), ll_scale, 1.0, nv.LAYOUT_LEGACY)
        a.point_mode = int(mode)
        if self._point_key is None:
            self._point_key = torch.zeros(2, dtype=torch.uint32, device=dev)
        a.key, a.loc, a.raw_scale = self._point_key.data_ptr(), z.data_ptr(), z.data_ptr()  # raw_scale is ignored
        kw = None
        if self.K:
            kw = self._f32(kwargs_vec, dev).reshape(-1)
            if kw.numel() != self.K:
                raise ValueError(f"expected {self.K} kwarg values")
            a.kwargs = kw.data_ptr()
        a.X = self._x_ptr(X)
        a.y = y.data_ptr()
        if out is None:
            out = torch.empty(1 + 2 * self.D + self.K, dtype=torch.float32, device=dev)
        a.out = out.data_ptr()
        a.x_planes = self._x_planes(a, X)
        ws = self._workspace(a)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(dev):
            nv.check(self.lib.bjx_elbo_step(C.byref(a), nv.current_stream_ptr()))
        return out

    def point_transform(self, z) -> torch.Tensor:
        """theta = bijector(z) for one unconstrained vector (transform_tree, core.py:154-155): the fused prologue with a zero scale."""
        dev = self.device
        z = self._f32(z, dev).reshape(-1)
        zero_scale = torch.full((self.D,), float("-inf") if self.scale_bijector == nv.SCALE_EXP else 0.0, device=dev)
        if self._point_key is None:
            self._point_key = torch.zeros(2, dtype=torch.uint32, device=dev)
        return self.posterior_sample(self._point_key, z, zero_scale, 1).reshape(-1)

    def posterior_log_prob(self, theta: torch.Tensor, loc, raw_scale) -> torch.Tensor:
        """log q(theta_i) for constrained samples theta [n, D] in the flat latent layout (Posterior.log_prob, core.py:57-71)."""
        dev = self.device
        theta = self._f32(theta, dev).reshape(-1, self.D)
        loc = self._f32(loc, dev).reshape(-1)
        raw = self._f32(raw_scale, dev).reshape(-1)
        a = self._args(1, 0, max(self.Dw, 1), 1.0, 1.0, nv.LAYOUT_LEGACY)
        a.loc, a.raw_scale = loc.data_ptr(), raw.data_ptr()
        if self.vi_type == nv.VI_LOW_RANK:  # Woodbury factors
            scratch = torch.empty(self.D * self.rank + self.D + 64, dtype=torch.float32, device=dev)
            a.workspace, a.workspace_bytes = scratch.data_ptr(), scratch.numel() * 4
        out = torch.empty(theta.shape[0], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nv.check(self.lib.bjx_posterior_log_prob(C.byref(a), theta.data_ptr(), theta.shape[0], out.data_ptr(), nv.current_stream_ptr()))
        return out

    def train_steps(self, keys, n_steps, params_flat, adam_m, adam_v, hyper, steps_done, losses, out, X, y, ll_scale,
                    n_samples, partitionable=None, minibatch=None, prior_weight=1.0, exchange=None):
        """Enqueue ``n_steps`` whole training steps (ELBO + gradients -> Adam -> loss record) with no host round trip
        (bjx_train_steps; train_fn's scan body, utils.py:22-28).  ``keys`` is the [n, 2] uint32 array of per-step keys,
        ``steps_done`` a device int64 counter, ``params_flat`` = [loc | raw_scale | kwargs] updated in place.
        ``minibatch``: the state of a MinibatchLoss (batch buffers + BjxMinibatch) -- every step first draws a fresh batch
        (bjx_train_steps_minibatch); X / y are then its batch buffers, or the full X read through its index.
        ``exchange``: a parallel.P2PAllReduce -- X / y are this rank's ROW SHARD, ``ll_scale`` uses the global length,
        ``prior_weight`` is 1 on one rank, and every step sums the packed output over the ranks on the device before Adam
        (bjx_train_steps_sharded); the caller advances ``exchange.epoch`` by the steps it enqueued."""
        indexed = minibatch is not None and minibatch["indexed"]
        X, y, N, ldx = self._data(X, y, indexed=indexed)
        a = self._args(n_samples, N, ldx if ldx else max(self.Dw, 1), ll_scale, prior_weight, brandom._layout(partitionable))
        assert keys.dtype == torch.uint32 and keys.is_cuda and keys.is_contiguous()
        assert params_flat.numel() == self.D + self.P + self.K and params_flat.is_contiguous()
        a.key = keys.data_ptr()
        a.X = self._x_ptr(X)
        a.y = y.data_ptr()
        a.out = out.data_ptr()
        a.loc = a.raw_scale = params_flat.data_ptr()  # replaced inside the library by views into params_flat
        if indexed:
            a.row_index = self._check_row_index(minibatch["idx"], N)
            if minibatch.get("planes_full") is not None:  # gather the batch's rows out of the full dataset's planes (cp.async gather)
                a.x_planes, a.N_full = minibatch["planes_full"].data_ptr(), int(X.shape[0])
        elif minibatch is not None:  # the sampler fills the batch's planes itself (bjx_minibatch_gather)
            a.x_planes = minibatch["planes_b"].data_ptr() if minibatch["planes_b"] is not None else None
        else:
            a.x_planes = self._x_planes(a, X)
        ws = self._workspace(a)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        lr, b1, b2, eps = hyper
        if minibatch is not None:
            with torch.cuda.device(self.device):
                nv.check(self.lib.bjx_train_steps_minibatch(C.byref(a), C.byref(minibatch["mb"]), int(n_steps), params_flat.data_ptr(),
                                                            adam_m.data_ptr(), adam_v.data_ptr(), lr, b1, b2, eps, steps_done.data_ptr(),
                                                            losses.data_ptr() if losses is not None else None, nv.current_stream_ptr()))
            return
        if exchange is not None:
            assert exchange.n == out.numel(), "the exchange buffer must be sized for the packed step output"
            with torch.cuda.device(self.device):
                nv.check(self.lib.bjx_train_steps_sharded(C.byref(a), int(n_steps), params_flat.data_ptr(), adam_m.data_ptr(), adam_v.data_ptr(),
                                                          lr, b1, b2, eps, steps_done.data_ptr(), losses.data_ptr() if losses is not None else None,
                                                          exchange._ptrs, exchange.rank, exchange.world, exchange.epoch, nv.current_stream_ptr()))
            return
        with torch.cuda.device(self.device):
            nv.check(self.lib.bjx_train_steps(C.byref(a), int(n_steps), params_flat.data_ptr(), adam_m.data_ptr(), adam_v.data_ptr(),
                                              lr, b1, b2, eps, steps_done.data_ptr(), losses.data_ptr() if losses is not None else None,
                                              nv.current_stream_ptr()))

    def posterior_sample(self, key, loc, raw_scale, n_samples, partitionable=None):
        """theta[S, D] = bijector(loc + scale * normal(key, (S, D)))  (Posterior.sample, core.py:46-55)."""
        dev = self.device
        key = brandom._check_key(key)
        loc = self._f32(loc, dev).reshape(-1)
        raw = self._f32(raw_scale, dev).reshape(-1)
        a = self._args(n_samples, 0, max(self.Dw, 1), 1.0, 1.0, brandom._layout(partitionable))
        a.likelihood, a.w_offset, a.Dw = nv.LIK_BERNOULLI_PROBS, 0, 0
        a.key, a.loc, a.raw_scale = key.data_ptr(), loc.data_ptr(), raw.data_ptr()
        ws = self._workspace(a)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        theta = torch.empty((int(n_samples), self.D), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nv.check(self.lib.bjx_posterior_sample(C.byref(a), theta.data_ptr(), None, nv.current_stream_ptr()))
        return theta

"""On-device minibatch sampler for the ELBO step.

The reference's only concession to large N is the ``ll / len(outputs) * full_data_size`` scaling of ``loss_fn``
(bijax/advi.py:92); which rows make up ``outputs`` / ``inputs`` is left to the caller.  What a JAX user writes around it,

    def loss(params, seed):
        k_batch, k_elbo = jax.random.split(seed)
        idx = jax.random.choice(k_batch, N, (B,))
        return model.loss_fn(params, y[idx], {"X": X[idx]}, full_data_size=N, seed=k_elbo, n_samples=S)

is ``MinibatchLoss(model, y, {"X": X}, batch_size=B, n_samples=S)`` here: one gather kernel (csrc/bjx_minibatch.cu) draws
the indices from the key and copies those rows -- of the prepared split-bf16 planes when the step's kernel is TMA-fed,
of raw fp32 X otherwise -- into batch buffers the step then reads.  ``train_fn(loss, params, optim.adam(...), n, seed)``
recognises the object and keeps the whole loop on the device (bjx_train_steps_minibatch in replayed CUDA graphs).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native as nv
from . import random as brandom


COPY_FREE_MIN_ROWS = 32768


class MinibatchLoss:
    def __init__(self, model, outputs, inputs, batch_size: int, n_samples: int = 1, full_data_size: Optional[int] = None,
                 copy_free: Optional[bool] = None):
        nv.require_cuda()
        self.model = model
        self.B = int(batch_size)
        self.S = int(n_samples)
        self.outputs = outputs
        self.inputs = inputs
        self.N = len(outputs)
        if self.B < 1 or self.N < 1:
            raise ValueError("need batch_size >= 1 and a non-empty dataset")
        if getattr(model, "_group", None) is not None:
            raise NotImplementedError("MinibatchLoss draws from one device's dataset; it does not combine with model.distributed()")
        self.full_data_size = float(self.N if full_data_size is None else full_data_size)
        # True: the step reads the dataset through the row index whenever its kernel can (BjxElboArgs.row_index); False:
        # always gather the batch into its own buffers; None: by batch size -- measured on B200 at D = 512 (profiles/
        # r01_minibatch.json) the copy-free path wins from a few tiles per SM upwards (76 vs 102 us per epoch at 64 Ki rows,
        # 329 vs 562 us at 512 Ki) and loses below (46 vs 36 us at 4 Ki rows: one tile per CTA, latency-bound)
        self.copy_free = (self.B >= COPY_FREE_MIN_ROWS) if copy_free is None else bool(copy_free)
        self._state = None  # (engine, buffers) built on first use: the engine depends on the params' trainable kwargs

    # ---- buffers ----------------------------------------------------------------------------------------------------
    def _setup(self, engine):
        if self._state is not None and self._state["engine"] is engine:
            return self._state
        dev = engine.device
        X = None
        if engine.lik_id != nv.LIK_BERNOULLI_PROBS:
            X = self.inputs[engine.x_name] if isinstance(self.inputs, dict) else self.inputs
        X, y, N, ldx = engine._data(X, self.outputs)
        st = {"engine": engine, "X": X, "y": y, "ldx": ldx}
        st["yb"] = torch.empty(self.B, dtype=torch.float32, device=dev)
        st["idx"] = torch.empty(self.B, dtype=torch.int32, device=dev)
        st["elbo_key"] = torch.zeros(2, dtype=torch.uint32, device=dev)
        from .dataset import PreparedX

        prepared = isinstance(X, PreparedX)  # the dataset as planes only: every route below that needs fp32 X is closed
        tma_fed = X is not None and (engine.wants_planes(self.S, self.B) or prepared)
        if prepared and not engine.supports_row_index(self.S, self.B, planes=True) and not engine.wants_planes(self.S, self.B):
            raise TypeError("a PreparedX dataset needs a batch step that reads operand planes (tensor-core kernels)")
        # two copy-free routes: (a) the register-converting kernel reading raw fp32 X through the index (Dw in {128, 256,
        # 384, 512}), (b) the planes kernel gathering rows of the full dataset's planes with cp.async (any Dw in [16, 512]).
        # Measured on B200 (profiles/r02_minibatch.json): (b) wins at Dw = 512 (0.56 vs 0.49 of the HBM roofline at 512 Ki
        # rows), (a) at the narrower exact widths (Dw = 128: 271 vs 335 us per 1 Mi rows), and only (b) exists elsewhere.
        raw_ok = X is not None and not prepared and self.copy_free and engine.supports_row_index(self.S, self.B)
        planes_ok = tma_fed and self.copy_free and engine.supports_row_index(self.S, self.B, planes=True)
        via_planes = planes_ok and (engine.Dw > 384 or not raw_ok)
        st["indexed"] = X is not None and self.copy_free and (via_planes or raw_ok)
        st["Xb"] = torch.zeros((self.B, engine.Dw), dtype=torch.float32, device=dev) if X is not None and not st["indexed"] else None
        st["planes_full"] = st["planes_b"] = None
        if tma_fed and (via_planes or not st["indexed"]):
            # keep the FULL dataset as planes (made once): the step gathers its rows itself, or the sampler copies them
            lib = engine.lib
            if prepared:
                full = X.planes
            else:
                full = torch.empty(int(lib.bjx_x_planes_bytes(N, engine.Dw)), dtype=torch.uint8, device=dev)
                with torch.cuda.device(dev):
                    nv.check(lib.bjx_prepare_x_planes(X.data_ptr(), N, engine.Dw, ldx, full.data_ptr(), nv.current_stream_ptr()))
            st["planes_full"] = full
            if not st["indexed"]:
                st["planes_b"] = torch.empty(int(lib.bjx_x_planes_bytes(self.B, engine.Dw)), dtype=torch.uint8, device=dev)
        mb = nv.BjxMinibatch()
        mb.N_full, mb.B, mb.Dw, mb.ldx_full = N, self.B, max(engine.Dw, 1), ldx if ldx else max(engine.Dw, 1)
        mb.y_full, mb.yb = y.data_ptr(), st["yb"].data_ptr()
        mb.idx_out, mb.elbo_key = st["idx"].data_ptr(), st["elbo_key"].data_ptr()
        if st["planes_b"] is not None:
            mb.x_planes_full, mb.planes_b = st["planes_full"].data_ptr(), st["planes_b"].data_ptr()
        elif X is not None and not st["indexed"]:
            if prepared:
                raise TypeError("this batch step reads fp32 rows; a PreparedX dataset holds only operand planes")
            mb.X_full, mb.Xb = X.data_ptr(), st["Xb"].data_ptr()
        st["mb"] = mb
        self._state = st
        return st

    def _engine(self, params):
        _, _, kwargs = self.model._split_params(params)
        engine, _ = self.model._engine_for(kwargs)
        return engine

    # ---- one draw ---------------------------------------------------------------------------------------------------
    def sample(self, params, seed):
        """Draw the batch of ``seed``: returns (idx int32[B], y[idx], X[idx] or None, k_elbo).  The returned tensors are
        the sampler's own buffers (overwritten by the next draw); X[idx] is None when the step reads the dataset through
        idx (copy-free path) and stale when it reads gathered planes."""
        st = self._setup(self._engine(params))
        seed = brandom._check_key(seed)
        with torch.cuda.device(st["engine"].device):
            nv.check(st["engine"].lib.bjx_minibatch_gather(C.byref(st["mb"]), seed.data_ptr(), None, brandom._layout(None), nv.current_stream_ptr()))
        for t in (st["idx"], st["yb"], st["Xb"], st["elbo_key"]):
            if t is not None:  # written by a raw kernel: tell torch (and the engine's planes cache) that the contents changed
                torch.autograd.graph.increment_version(t)
        return st["idx"], st["yb"], st["Xb"], st["elbo_key"]

    def __call__(self, params, seed):
        """The loss of the docstring's ``loss(params, seed)``; differentiable like ``model.loss_fn``."""
        from .advi import _ElboFunction

        model = self.model
        loc, raw, kwargs = model._split_params(params)
        engine, kw_leaves = model._engine_for(kwargs)
        kw = torch.cat([l.reshape(-1).to(torch.float32) for l in kw_leaves]) if kw_leaves else None
        st = self._setup(engine)
        self.sample(params, seed)
        ll_scale = self.full_data_size / float(self.B)  # advi.py:92
        Xin = st["X"] if st["indexed"] else st["Xb"]

        def run(loc_, raw_, kw_):
            return engine.step(st["elbo_key"], loc_, raw_, kw_, Xin, st["yb"], ll_scale, self.S,
                               row_index=st["idx"] if st["indexed"] else None,
                               planes=st["planes_full"] if st["indexed"] else st["planes_b"], n_full=self.N)

        return _ElboFunction.apply(loc, raw, kw, run)

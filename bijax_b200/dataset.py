"""A design matrix that lives on the device ONLY as its split-bf16 operand planes.

The tensor-core kernels of the ELBO step never read fp32 X once its planes exist (``x = x1 + x2`` to 2^-17 relative, two bf16
planes = 4 bytes per element, exactly the footprint of fp32).  ``engine.prepare`` / the first step make the planes from a
device tensor the caller keeps, so the dataset is then resident twice (41 GB at c3, 51 GB per GPU at c5).  ``PreparedX``
streams X through a staging buffer instead -- from a host tensor / array, or from chunks produced on the fly -- and keeps the
planes alone; pass it wherever ``inputs["X"]`` goes:

    Xp = bj.PreparedX.from_tensor(X_host, chunk_rows=1 << 20)        # or .from_chunks(iterable_of_[rows, D]_blocks, N, D)
    loss_fn = partial(model.loss_fn, outputs=y, inputs={"X": Xp}, full_data_size=N, n_samples=32)

It is accepted by every shape whose kernel reads the planes (the fused single-pass kernel: 16 <= D <= 512 with S <= 32; the
two-pass GEMM path outside ``precision="bf16x3"``) and refused loudly elsewhere (the CUDA-core kernels read fp32 X).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _native as nv


class PreparedX:
    def __init__(self, n_rows: int, n_cols: int, device=None):
        nv.require_cuda()
        self.lib = nv.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.shape = (int(n_rows), int(n_cols))
        if self.shape[0] < 1 or self.shape[1] < 1:
            raise ValueError("PreparedX needs at least one row and one column")
        self.planes = torch.empty(int(self.lib.bjx_x_planes_bytes(*self.shape)), dtype=torch.uint8, device=self.device)
        self.rows_filled = 0

    def __len__(self) -> int:
        return self.shape[0]

    @property
    def nbytes(self) -> int:
        return self.planes.numel()

    def append(self, block) -> "PreparedX":
        """Split the next ``block`` ([rows, D], any device / float dtype) into rows [rows_filled, rows_filled + rows) of the planes."""
        if not isinstance(block, torch.Tensor):
            block = torch.as_tensor(block)
        if block.dim() != 2 or block.shape[1] != self.shape[1]:
            raise ValueError(f"block must be [rows, {self.shape[1]}]; got {tuple(block.shape)}")
        rows = int(block.shape[0])
        if self.rows_filled + rows > self.shape[0]:
            raise ValueError("more rows than the dataset was declared with")
        stage = block.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        with torch.cuda.device(self.device):
            nv.check(self.lib.bjx_prepare_x_planes_rows(stage.data_ptr(), rows, self.shape[1], stage.stride(0), self.planes.data_ptr(),
                                                        self.shape[0], self.rows_filled, nv.current_stream_ptr()))
        stage.record_stream(torch.cuda.current_stream(self.device))  # the split reads it asynchronously
        self.rows_filled += rows
        return self

    def _require_complete(self):
        if self.rows_filled != self.shape[0]:
            raise RuntimeError(f"PreparedX holds {self.rows_filled} of its {self.shape[0]} rows")

    @classmethod
    def from_chunks(cls, chunks: Iterable, n_rows: int, n_cols: int, device=None) -> "PreparedX":
        out = cls(n_rows, n_cols, device)
        for block in chunks:
            out.append(block)
        out._require_complete()
        return out

    @classmethod
    def from_tensor(cls, X, chunk_rows: int = 1 << 20, device: Optional[torch.device] = None) -> "PreparedX":
        if not isinstance(X, torch.Tensor):
            X = torch.as_tensor(X)
        if X.dim() != 2:
            raise ValueError("X must be [N, D]")
        return cls.from_chunks((X[i : i + chunk_rows] for i in range(0, X.shape[0], int(chunk_rows))), X.shape[0], X.shape[1], device)

"""jax.random on the B200: PRNGKey / split / bits / uniform / normal backed by the threefry2x32 CUDA kernel.

Mirrors the calls bijax makes (advi.py:77,83; utils.py:30,53-55,70; core.py:103,169).  Keys are uint32[2] device
tensors, exactly JAX's raw key data.  ``config.threefry_partitionable`` selects the counter layout
(``jax_threefry_partitionable``): False = the layout JAX used when bijax was written (default here), True = the JAX >= 0.5
default.
"""
from __future__ import annotations

import math
from typing import Sequence, Union

import torch

from . import _native as nv


class _Config:
    threefry_partitionable: bool = False


config = _Config()


def _layout(partitionable=None) -> int:
    if partitionable is None:
        partitionable = config.threefry_partitionable
    return nv.LAYOUT_PARTITIONABLE if partitionable else nv.LAYOUT_LEGACY


def _device(device=None) -> torch.device:
    nv.require_cuda()
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def PRNGKey(seed: int, device=None) -> torch.Tensor:
    """uint32[2] = [seed >> 32, seed & 0xffffffff] on the device."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    # torch.uint32 construction goes through int64 to avoid signed-overflow checks
    host = torch.tensor([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=torch.int64).to(torch.uint32)
    return host.to(_device(device))


def _check_key(key: torch.Tensor) -> torch.Tensor:
    if not (isinstance(key, torch.Tensor) and key.dtype == torch.uint32 and key.numel() == 2 and key.is_cuda):
        raise TypeError("key must be a CUDA uint32 tensor with 2 elements (use bijax_b200.random.PRNGKey)")
    return key.contiguous()


def _shape(shape) -> tuple:
    if isinstance(shape, int):
        return (shape,)
    return tuple(int(s) for s in shape)


def split(key: torch.Tensor, num: int = 2, partitionable=None) -> torch.Tensor:
    key = _check_key(key)
    out = torch.empty((int(num), 2), dtype=torch.uint32, device=key.device)
    with torch.cuda.device(key.device):
        nv.check(nv.load().bjx_random_split(key.data_ptr(), int(num), _layout(partitionable), out.data_ptr(), nv.current_stream_ptr()))
    return out


def bits(key: torch.Tensor, shape: Union[int, Sequence[int]] = (), partitionable=None, offset: int = 0, total=None) -> torch.Tensor:
    key = _check_key(key)
    shape = _shape(shape)
    n = math.prod(shape) if shape else 1
    total = n if total is None else int(total)
    out = torch.empty(shape, dtype=torch.uint32, device=key.device)
    with torch.cuda.device(key.device):
        nv.check(nv.load().bjx_random_bits(key.data_ptr(), total, int(offset), n, _layout(partitionable), out.data_ptr(), nv.current_stream_ptr()))
    return out


def uniform(key, shape=(), minval: float = 0.0, maxval: float = 1.0, partitionable=None, offset: int = 0, total=None, out=None) -> torch.Tensor:
    key = _check_key(key)
    shape = _shape(shape)
    n = math.prod(shape) if shape else 1
    total = n if total is None else int(total)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=key.device)
    with torch.cuda.device(key.device):
        nv.check(
            nv.load().bjx_random_uniform(key.data_ptr(), total, int(offset), n, _layout(partitionable), float(minval), float(maxval), out.data_ptr(), nv.current_stream_ptr())
        )
    return out


def normal(key, shape=(), partitionable=None, offset: int = 0, total=None, out=None) -> torch.Tensor:
    """jax.random.normal(key, shape, float32).  With the partitionable layout, ``offset``/``total`` regenerate any slice
    [offset, offset + prod(shape)) of a longer stream (used to shard synthetic data across GPUs)."""
    key = _check_key(key)
    shape = _shape(shape)
    n = math.prod(shape) if shape else 1
    total = n if total is None else int(total)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=key.device)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n
    with torch.cuda.device(key.device):
        nv.check(nv.load().bjx_random_normal(key.data_ptr(), total, int(offset), n, _layout(partitionable), out.data_ptr(), nv.current_stream_ptr()))
    return out


def randint(key, shape, minval: int, maxval: int, partitionable=None) -> torch.Tensor:
    """jax.random.randint(key, shape, minval, maxval) for the default int32 dtype (the index draw of the on-device
    minibatch sampler, csrc/bjx_minibatch.cu; restated in oracle/threefry.py::randint)."""
    key = _check_key(key)
    shape = _shape(shape)
    n = math.prod(shape) if shape else 1
    span = int(maxval) - int(minval)
    if span < 1:
        span = 1  # jax returns minval everywhere
    if span >= 2**31 or not (-(2**31) <= int(minval) < 2**31):
        raise ValueError("randint: int32 range only")
    out = torch.empty(n, dtype=torch.int32, device=key.device)
    if n:
        mb = nv.BjxMinibatch()
        mb.N_full, mb.B, mb.idx_out = span, n, out.data_ptr()
        with torch.cuda.device(key.device):
            nv.check(nv.load().bjx_minibatch_gather(mb, key.data_ptr(), None, _layout(partitionable), nv.current_stream_ptr()))
    if minval:
        out += int(minval)
    return out.reshape(shape)


def choice(key, a: int, shape=(), partitionable=None) -> torch.Tensor:
    """jax.random.choice(key, a, shape) for an integer ``a`` with the defaults replace=True, p=None: randint(key, shape, 0, a)."""
    return randint(key, shape, 0, int(a), partitionable)

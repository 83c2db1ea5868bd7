"""Data parallelism for the ELBO step: rows of X / y, or -- when N is small -- Monte-Carlo samples.

The log-likelihood is a sum over data rows (advi.py:91) and the loss a mean over samples (advi.py:95), so the path
shards by rows with ONE exchange step: a sum all-reduce of the packed [loss | d loc | d raw_scale | d kwargs] buffer.
Every rank regenerates the same eps from the same key (counter-based PRNG), uses the GLOBAL batch length in
``full_data_size / len(outputs)`` (advi.py:92), and exactly one rank contributes the q / prior terms
(``prior_weight`` = 1 on rank 0, 0 elsewhere), so the reduced buffer equals the single-device result.

Sample sharding (``shard_samples``): every rank holds ALL the data and evaluates samples [s0, s1) of the S-sample step;
eps comes from the global flat index of the (S, D) draw, so the shards reproduce the unsharded noise exactly; every rank
carries the q / prior terms of its own samples (``prior_weight`` = 1 everywhere), the per-step constant of the entropy
gradient is added by one rank (``entropy_weight``), means divide by the global S, and the same single all-reduce follows.

These helpers are host logic only (they work with gloo on CPU tensors for tests and with NCCL on CUDA tensors).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_rows(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rows owned by ``rank``; sizes differ by at most one row."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_samples(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [s0, s1) of Monte-Carlo samples owned by ``rank`` (sizes differ by at most one)."""
    return shard_rows(n_samples, rank, world)


def entropy_weight(rank: int) -> float:
    return 1.0 if rank == 0 else 0.0


def prior_weight(rank: int) -> float:
    return 1.0 if rank == 0 else 0.0


def ll_scale(full_data_size: float, global_len: int) -> float:
    return float(full_data_size) / float(global_len)


def global_len(local_len: int, group=None, device=None) -> int:
    t = torch.tensor([int(local_len)], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item())


def allreduce_packed(out: torch.Tensor, group=None) -> torch.Tensor:
    """The single exchange step of the path: in-place sum of the packed loss+gradient buffer."""
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out

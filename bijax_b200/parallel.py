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


class P2PAllReduce:
    """The exchange step as ONE kernel over NVLink peer memory (bjx_allreduce_packed_p2p) instead of an NCCL call: every
    rank stores its packed [loss | grads] vector into a slot of every peer's buffer, raises a flag, waits for the peers'
    flags and sums the slots in rank order.  One process per GPU of ONE box (CUDA IPC); at most 8 ranks.  The process
    group is used once, to exchange the 64-byte IPC handles."""

    def __init__(self, n: int, group=None, device=None):
        import ctypes as C

        from . import _native as nv

        self.lib = nv.load()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("the peer-memory all-reduce covers the <= 8 GPUs of one box")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n = int(n)
        nbytes = int(self.lib.bjx_p2p_buffer_bytes(self.n, self.world))
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        error = None
        try:
            with torch.cuda.device(self.device):
                nv.check(self.lib.bjx_p2p_alloc(nbytes, C.byref(own), handle))
        except Exception as exc:  # keep going: every rank must take part in the collectives below
            error = exc
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        self._own = own if error is None else None
        self._ptrs = (C.c_void_p * self.world)()
        self._opened = []
        if error is None:
            try:
                with torch.cuda.device(self.device):
                    for p in range(self.world):
                        if p == self.rank:
                            self._ptrs[p] = own.value
                        else:
                            q = C.c_void_p()
                            hb = (C.c_ubyte * 64).from_buffer_copy(handles[p])
                            nv.check(self.lib.bjx_p2p_open(hb, C.byref(q)))
                            self._ptrs[p] = q.value
                            self._opened.append(q)
            except Exception as exc:
                error = exc
        # agreement (also the barrier: every buffer is mapped everywhere before the first epoch): all ranks succeed or all raise
        ok = torch.tensor([0.0 if error is not None else 1.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() == 0.0:
            self._release()
            raise RuntimeError(f"peer-memory all-reduce could not be set up on every rank (this rank: {error})")
        self.epoch = 0

    def __call__(self, out: torch.Tensor) -> torch.Tensor:
        from . import _native as nv

        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == self.n
        self.epoch += 1
        with torch.cuda.device(self.device):
            nv.check(self.lib.bjx_allreduce_packed_p2p(out.data_ptr(), self.n, self._ptrs, self.rank, self.world, self.epoch,
                                                       nv.current_stream_ptr()))
        return out

    def _release(self):
        with torch.cuda.device(self.device):
            for q in self._opened:
                self.lib.bjx_p2p_close(q)
            self._opened = []
            if self._own is not None:
                self.lib.bjx_p2p_free(self._own)
                self._own = None

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody is still pushing into a buffer that is about to be unmapped
        self._release()

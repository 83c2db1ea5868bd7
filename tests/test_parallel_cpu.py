"""CPU, world_size 2 over gloo: the sharding arithmetic of the N>1 path (row blocks, global ll_scale, prior_weight on
one rank, one sum all-reduce of the packed buffer) reproduces the unsharded result.  The oracle stands in for the
kernels here (no GPU in this container); the same helpers drive NCCL on the GPU box."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bijax_b200 import parallel as par
from oracle import advi_ref as R
from oracle import threefry as tf
from tests import cases as C


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _packed_oracle(X, y, loc, raw, lns, full, n_global, weight, eps):
    """What one rank's bjx_elbo_step writes: loss/grad partials with ll_scale from the GLOBAL length and the q/prior
    terms scaled by prior_weight."""
    loc_t = torch.tensor(loc, dtype=torch.float64, requires_grad=True)
    raw_t = torch.tensor(raw, dtype=torch.float64, requires_grad=True)
    lns_t = torch.tensor(lns, dtype=torch.float64, requires_grad=True)
    D = loc_t.numel()
    sigma = torch.exp(raw_t)
    z = loc_t + sigma * eps
    q = (-0.5 * eps * eps - torch.log(sigma)).sum(-1) - 0.5 * D * R.LOG_2PI
    p = (-0.5 * z * z - 0.5 * R.LOG_2PI).sum(-1)
    tau = torch.exp(lns_t)
    r = (torch.tensor(y, dtype=torch.float64)[:, None] - torch.tensor(X, dtype=torch.float64) @ z.T) / tau
    ll = (-0.5 * r * r - 0.5 * R.LOG_2PI - torch.log(tau)).sum(0)
    loss = (weight * (q - p) - (full / n_global) * ll).mean()
    g = torch.autograd.grad(loss, [loc_t, raw_t, lns_t])
    return torch.cat([loss.detach().reshape(1), g[0], g[1], g[2].reshape(1)])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, D, S = 1001, 6, 5
        X, y = C.synth_linear(N, D)
        loc = 0.1 * tf.normal(tf.prng_key(3), (D,), 1).astype(np.float64)
        raw = np.full(D, math.log(0.2))
        eps = torch.tensor(tf.normal(tf.prng_key(9), (S, D), tf.LAYOUT_LEGACY), dtype=torch.float64)  # same key on every rank
        a, b = par.shard_rows(N, rank, world)
        n_glob = par.global_len(b - a)
        assert n_glob == N
        out = _packed_oracle(X[a:b], y[a:b], loc, raw, -0.5, 3.0 * N, n_glob, par.prior_weight(rank), eps)
        par.allreduce_packed(out)
        if rank == 0:
            want = _packed_oracle(X, y, loc, raw, -0.5, 3.0 * N, N, 1.0, eps)
            q.put((out.numpy(), want.numpy()))
    finally:
        dist.destroy_process_group()


def _packed_oracle_samples(X, y, loc, raw, lns, full, eps_all, s0, s1, entropy_w):
    """What one rank's bjx_elbo_step writes under SAMPLE sharding: sums over its samples [s0, s1) divided by the global S,
    q / prior terms of its own samples at full weight, and the per-step constant of the entropy gradient
    (-d sum log sigma / d raw = -1 under Exp) carried by one rank only."""
    S = eps_all.shape[0]
    loc_t = torch.tensor(loc, dtype=torch.float64, requires_grad=True)
    raw_t = torch.tensor(raw, dtype=torch.float64, requires_grad=True)
    lns_t = torch.tensor(lns, dtype=torch.float64, requires_grad=True)
    D = loc_t.numel()
    eps = eps_all[s0:s1]
    sigma = torch.exp(raw_t)
    z = loc_t + sigma * eps
    logsig_const = torch.log(sigma).sum().detach()  # value counted per sample; its gradient is the once-per-step constant
    q = (-0.5 * eps * eps).sum(-1) - logsig_const - 0.5 * D * R.LOG_2PI
    p = (-0.5 * z * z - 0.5 * R.LOG_2PI).sum(-1)
    tau = torch.exp(lns_t)
    r = (torch.tensor(y, dtype=torch.float64)[:, None] - torch.tensor(X, dtype=torch.float64) @ z.T) / tau
    ll = (-0.5 * r * r - 0.5 * R.LOG_2PI - torch.log(tau)).sum(0)
    loss = ((q - p) - (full / len(y)) * ll).sum() / S - entropy_w * (torch.log(sigma).sum() - logsig_const)
    g = torch.autograd.grad(loss, [loc_t, raw_t, lns_t])
    return torch.cat([loss.detach().reshape(1), g[0], g[1], g[2].reshape(1)])


def _worker_samples(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, D, S = 40, 6, 7  # small N: shard the Monte-Carlo samples instead of the rows
        X, y = C.synth_linear(N, D)
        loc = 0.1 * tf.normal(tf.prng_key(3), (D,), 1).astype(np.float64)
        raw = np.full(D, math.log(0.2))
        eps = torch.tensor(tf.normal(tf.prng_key(9), (S, D), tf.LAYOUT_LEGACY), dtype=torch.float64)
        s0, s1 = par.shard_samples(S, rank, world)
        out = _packed_oracle_samples(X, y, loc, raw, -0.5, 3.0 * N, eps, s0, s1, par.entropy_weight(rank))
        par.allreduce_packed(out)
        if rank == 0:
            want = _packed_oracle(X, y, loc, raw, -0.5, 3.0 * N, N, 1.0, eps)
            q.put((out.numpy(), want.numpy()))
    finally:
        dist.destroy_process_group()


def test_sample_sharded_allreduce_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_samples, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    assert par.shard_samples(7, 0, 2) == (0, 4) and par.shard_samples(7, 1, 2) == (4, 7)
    assert par.entropy_weight(0) == 1.0 and par.entropy_weight(1) == 0.0


def test_row_sharded_allreduce_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_shard_rows_partition():
    for n, w in ((10, 3), (100_000_000, 8), (7, 8), (0, 2)):
        spans = [par.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert par.shard_rows(100_000_000, 3, 8) == (37_500_000, 50_000_000)
    assert par.prior_weight(0) == 1.0 and par.prior_weight(5) == 0.0

"""torchrun --nproc-per-node N scripts/multi_gpu_train_check.py

Row-sharded training on N GPUs of one box, three ways, on the same synthetic logistic problem:
  (a) single-GPU device-resident loop on ALL the rows (rank 0 only)                     -- the reference trajectory
  (b) row-sharded HOST-driven loop: bjx_elbo_step + peer-memory all-reduce + bjx_adam_update per step
  (c) row-sharded DEVICE-resident loop: bjx_train_steps_sharded in replayed CUDA graphs (no host work per step)
(b) and (c) must agree with each other to float32 rounding and with (a) to the shard-additivity tolerance; parameters must
be bit-identical on every rank.  Also times (b) against (c) per epoch.  Exit code 0 only if every check holds."""
import json, math, os, sys, time
from functools import partial

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import distributions as tfd, likelihoods as lk, optim, random as br
from bijax_b200 import utils as U

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

N, D, S = int(os.environ.get("MG_ROWS", 2_000_000)), 512, 32
n_epochs = 3 * U.GRAPH_STEPS + 5
X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True)
X.mul_(1.0 / math.sqrt(D))
w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
y = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < torch.sigmoid(X @ w)).float()
lo, hi = rank * N // world, (rank + 1) * N // world
Xs, ys = X[lo:hi].contiguous(), y[lo:hi].contiguous()
prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}


def fresh(model):
    return model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))


def train(model, Xd, yd, fused, n):
    loss_fn = partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=S)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    res = bj.train(loss_fn, fresh(model), optim.adam(0.01), n_epochs=n, seed=br.PRNGKey(3), fused=fused)
    torch.cuda.synchronize(); dist.barrier()
    return res, 1e6 * (time.perf_counter() - t0) / n


single = None
if rank == 0:
    m0 = bj.ADVI(prior, {}, lk.BernoulliLogits())
    loss_fn = partial(m0.loss_fn, outputs=y, inputs={"X": X}, full_data_size=N, n_samples=S)
    single = bj.train(loss_fn, fresh(m0), optim.adam(0.01), n_epochs=n_epochs, seed=br.PRNGKey(3))
    for e in m0._engines.values():
        e.release_planes()
del X, y
torch.cuda.empty_cache()

mh = bj.ADVI(prior, {}, lk.BernoulliLogits()).distributed(shard="rows", p2p=True)
host, _ = train(mh, Xs, ys, False, n_epochs)
host2, us_host = train(mh, Xs, ys, False, n_epochs)
md = bj.ADVI(prior, {}, lk.BernoulliLogits()).distributed(shard="rows", p2p=True)
devr, _ = train(md, Xs, ys, True, n_epochs)
devr2, us_dev = train(md, Xs, ys, True, n_epochs)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


checks = {}
checks["device_loop_vs_host_loop_losses"] = rel(devr["losses"], host["losses"])
checks["device_loop_vs_host_loop_loc"] = rel(devr["params"]["posterior"]["loc"], host["params"]["posterior"]["loc"])
checks["device_loop_repeatable"] = bool(torch.equal(devr["losses"], devr2["losses"]))
flat = torch.cat([devr["params"]["posterior"]["loc"], devr["params"]["posterior"]["scale_diag"], devr["losses"]])
parts = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(parts, flat)
checks["params_and_losses_bit_identical_on_every_rank"] = bool(all(torch.equal(parts[0], p) for p in parts))
ok = (checks["device_loop_vs_host_loop_losses"] < 2e-5 and checks["device_loop_vs_host_loop_loc"] < 2e-4
      and checks["device_loop_repeatable"] and checks["params_and_losses_bit_identical_on_every_rank"])
if rank == 0:
    checks["sharded_vs_single_gpu_losses"] = rel(devr["losses"], single["losses"])
    checks["sharded_vs_single_gpu_loc"] = rel(devr["params"]["posterior"]["loc"], single["params"]["posterior"]["loc"])
    ok = ok and checks["sharded_vs_single_gpu_losses"] < 2e-5 and checks["sharded_vs_single_gpu_loc"] < 5e-4
    print(json.dumps({"world": world, "rows_total": N, "rows_per_gpu": hi - lo, "epochs": n_epochs, "ok": bool(ok), "checks": checks,
                      "us_per_epoch_host_driven": us_host, "us_per_epoch_device_resident": us_dev}), flush=True)
okt = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
for m in (mh, md):
    if m._p2p_op is not None:
        m._p2p_op.close()
dist.destroy_process_group()
sys.exit(0 if okt.item() == 1.0 else 1)

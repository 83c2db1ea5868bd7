#!/usr/bin/env python
"""bench.py -- ELBO+grad steps/sec of the B200 ADVI engine on BASELINE.json's workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c3|c1|c2|c4]
                    [--precision auto|fp32|bf16x3|bf16x2]

One "step" = one evaluation of the negative Monte-Carlo ELBO and all its gradients at a fresh PRNG key
(bijax/advi.py:80-95 under jax.value_and_grad, bijax/utils.py:7,25), optimiser update excluded.

Workloads (synthetic, generated on the device from threefry):
  --config c3 (default; the configuration BASELINE.json's metric is quoted on)
      N=1 : Bayesian logistic regression, mean-field, D=512, N=10M rows, S=32, X fp32 [N, D] row-major.
      N>1 : config c5 shard -- the same model with 12.5M rows per GPU (100M rows at 8 GPUs), rows sharded, one exchange
            of the packed [loss | grads] buffer per step.  Weak scaling: per-GPU rows are fixed.
      `value` is the whole-job rate in c3-equivalent steps/s: (S*N_total*D per second) / (32 * 10M * 512); at N=1 it IS
      the step rate of c3.
  --config c1 | c2 | c4   the other BASELINE.json configs at their full sizes on ONE GPU (parity-test cases given a
      driver-visible line): c1 coin toss (latency), c2 linear regression D=5 N=1M S=64 (fp32 FMA), c4 full-rank linear
      regression D=2048 N=1M S=256 (tensor pipe).  `value` is the plain step rate of that config.

--impl reference times the CPU restatement of the reference (oracle/, torch-CPU fp32, all host threads) on the same
config; JAX/TFP cannot be installed here.  Each of its steps is a bounded row sample of the workload (all of
it where that fits); `ms_per_step` is the MEASURED time of one sampled step, the sample's share of the workload is in
`cpu_baseline.sample_fraction`, and `value` is the work rate in the metric's unit.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

D_C3, S_C3, N_C3 = 512, 32, 10_000_000
ROWS_PER_GPU_MULTI = 12_500_000
METRIC = "ELBO+grad steps/sec (S x N x D per step)"
UNIT = "steps/s (c3-equivalent: 32 x 10M x 512 S*N*D per step)"
L2_NOTE = "inputs_larger_than_l2 (X shard is >= 20 GB, read once per step)"
SMALL = {
    "c1": dict(D=1, S=16, N=100, workload="c1: coin toss, Beta(0.5,0.5) prior + Sigmoid bijector, mean-field, 100 Bernoulli observations, S=16"),
    "c2": dict(D=5, S=64, N=1_000_000, workload="c2: Bayesian linear regression, mean-field, D=5 + learnable log_noise_scale, N=1M, S=64"),
    "c4": dict(D=2048, S=256, N=1_000_000, workload="c4: full-rank ADVI Bayesian linear regression, D=2048, N=1M, S=256, learnable log_noise_scale"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c1", "c2", "c4"])
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--rows-per-gpu", type=int, default=0, help="override the per-GPU row count (0 = config default)")
    ap.add_argument("--cpu-rows", type=int, default=2_000_000, help="row sample of the cpu_baseline leg of our own arm")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the cpu_baseline leg of our own arm")
    ap.add_argument("--ref-seconds", type=float, default=100.0,
                    help="--impl reference: target wall time of all (steps + warmup) sampled steps; the row sample is sized to it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 1 s timed region behind roofline.sustained")
    ap.add_argument("--no-same-shard", action="store_true", help="N=1: skip the extra run at 12.5M rows (n1_same_shard_value)")
    ap.add_argument("--no-multi-gpu-check", action="store_true")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="N > 1: use NCCL for the exchange step instead of the one-kernel all-reduce over NVLink peer memory "
                         "(bjx_allreduce_packed_p2p, the default)")
    ap.add_argument("--raw-fp32", action="store_true",
                    help="do not keep the dataset as prepared split-bf16 planes: the kernel converts raw fp32 X in registers every step")
    return ap.parse_args()


def c3_config(world: int, rows: int):
    """`config` of the JSON line -- identical in both arms (the arm-specific choices are under `impl_detail`)."""
    n_total = rows * world
    if world == 1 and rows == N_C3:
        workload = "c3: Bayesian logistic regression, mean-field, D=512, N=10M, S=32 on 1 B200"
    elif world == 1:
        workload = f"c3 shape: Bayesian logistic regression, mean-field, D=512, N={rows}, S=32 on 1 B200"
    else:
        workload = (f"c5 shard: Bayesian logistic regression, mean-field, D=512, S=32, {rows} rows per GPU x {world} GPUs, "
                    "one sum all-reduce of the packed [loss|grads] buffer per step")
    return {"workload": workload, "rows_per_gpu": rows, "rows_total": n_total, "D": D_C3, "S": S_C3, "l2": L2_NOTE,
            "threefry_layout": "legacy"}


# ----------------------------------------------------------------------------------------------------------------------
# CPU restatement legs (the ONLY places bench.py executes oracle/)
def _cpu_setup():
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return cores


def cpu_c3_steps(X, y, n_steps: int, warmup: int, seconds: float = 0.0):
    """oracle/advi_ref.py's batched logistic restatement (fp32, torch CPU, all threads) on the rows of X.
    Returns per-step wall times (s) of the timed steps.  seconds > 0: stop early once the budget is spent (>= 3 steps)."""
    import torch

    from oracle import advi_ref as R
    from oracle import threefry as tf

    loc = torch.zeros(D_C3, dtype=torch.float32)
    raw = torch.full((D_C3,), math.log(0.1), dtype=torch.float32)
    keys = tf.split(tf.prng_key(0), n_steps + warmup, tf.LAYOUT_LEGACY)
    times = []
    t_begin = time.perf_counter()
    for i in range(n_steps + warmup):
        eps = torch.from_numpy(tf.normal(keys[i], (S_C3, D_C3), tf.LAYOUT_LEGACY))
        t0 = time.perf_counter()
        R.logistic_meanfield_value_and_grad(loc, raw, X, y, float(N_C3), eps)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if seconds > 0 and i >= warmup + 2 and time.perf_counter() - t_begin > seconds:
            break
    return times


def cpu_c3_data(rows: int):
    import torch

    g = torch.Generator().manual_seed(1234)
    X = torch.empty(rows, D_C3, dtype=torch.float32)
    for i in range(0, rows, 1 << 18):  # blockwise: no second full-size temporary
        X[i : i + (1 << 18)] = torch.randn(min(1 << 18, rows - i), D_C3, generator=g, dtype=torch.float32) / math.sqrt(D_C3)
    w = torch.randn(D_C3, generator=g, dtype=torch.float32)
    y = (torch.rand(rows, generator=g) < torch.sigmoid(X @ w)).float()
    return X, y


def cpu_small_step_fn(cfg: str, rows: int):
    """A callable running ONE CPU-restatement step of c1 / c2 / c4 on `rows` rows (fp32, all threads) + a description."""
    import numpy as np
    import torch

    from oracle import advi_ref as R
    from oracle import threefry as tf

    c = SMALL[cfg]
    D, S = c["D"], c["S"]
    g = torch.Generator().manual_seed(1234)
    eps = torch.from_numpy(tf.normal(tf.prng_key(1), (S, D), tf.LAYOUT_LEGACY))
    if cfg == "c1":
        ref = R.ADVIRef({"p": R.Beta(0.5, 0.5)}, {"p": R.Sigmoid()}, R.bernoulli_probs_loglik("p"))
        y = torch.tensor([1.0 if n % 10 < 7 else 0.0 for n in range(c["N"])])
        z = torch.zeros(1)
        return (lambda: ref.value_and_grad(z, z, {}, y, None, float(c["N"]), eps)), "oracle ADVIRef.value_and_grad (per-sample restatement)"
    X = torch.randn(rows, D, generator=g, dtype=torch.float32) / math.sqrt(D)
    w = torch.randn(D, generator=g, dtype=torch.float32)
    y = X @ w + 0.1 * torch.randn(rows, generator=g)
    lns = torch.tensor(math.log(0.1))
    if cfg == "c2":
        loc, raw = torch.zeros(D), torch.full((D,), math.log(0.1))
        return (lambda: R.linreg_meanfield_value_and_grad(loc, raw, lns, X, y, float(c["N"]), eps)), "oracle linreg_meanfield_value_and_grad (batched restatement)"
    prior = {"weights": R.MultivariateNormalDiag(torch.zeros(D), torch.ones(D))}
    ref = R.ADVIRef(prior, {}, R.gaussian_linear_loglik(), "full_rank")
    loc, M = torch.zeros(D), 0.1 * torch.eye(D)
    return (lambda: ref.value_and_grad_chunked(loc, M, {"log_noise_scale": lns}, y, X, float(c["N"]), eps, chunk_rows=1 << 14,
                                                n_total=rows, dtype=torch.float32)), "oracle ADVIRef.value_and_grad_chunked (full-rank restatement, float32)"


def time_cpu_steps(fn, n_steps: int, warmup: int, seconds: float = 0.0):
    times = []
    t_begin = time.perf_counter()
    for i in range(n_steps + warmup):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if seconds > 0 and i >= warmup + 2 and time.perf_counter() - t_begin > seconds:
            break
    return times


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0: float, t1: float):
        """Clocks / throttle reasons of the samples taken inside [t0, t1] (the sampler keeps running)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for t, line in list(self.lines):
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def load_peaks():
    pk = ROOT / "MEASURED_PEAKS.json"
    return json.loads(pk.read_text()) if pk.exists() else {}


# ----------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: the CPU restatement on the box's host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    cores = _cpu_setup()
    K, W = args.steps, max(args.warmup, 1)
    if args.config == "c3":
        world = max(args.gpus, 1)
        rows_cfg = args.rows_per_gpu or (N_C3 if world == 1 else ROWS_PER_GPU_MULTI)
        n_total = rows_cfg * world
        # size the row sample so that (K + W) steps take about --ref-seconds, within the host's memory
        Xp, yp = cpu_c3_data(50_000)
        probe = cpu_c3_steps(Xp, yp, 2, 1)
        per_row = min(probe) / 50_000
        rows = int(args.ref_seconds / (K + W) / per_row)
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = 32 << 30
        rows = max(50_000, min(rows, n_total, int(0.45 * avail / (4 * D_C3 + 10 * 4 * S_C3))))
        rows -= rows % 1000
        X, y = cpu_c3_data(rows)
        times = cpu_c3_steps(X, y, K, W)
        ms = 1e3 * sum(times) / len(times)
        value = (rows / N_C3) / (ms * 1e-3)
        frac = rows / n_total
        sample = (f"{rows} rows x D=512 x S=32 per step = {frac:.4f} of the workload's {n_total} rows ({len(times)} timed steps, mean); "
                  "X ~ N(0, 1/D) drawn with torch.randn on the host (same distribution as the device generator)")
        config, unit, restatement = c3_config(world, rows_cfg), UNIT, "oracle logistic_meanfield_value_and_grad (batched restatement)"
    else:
        c = SMALL[args.config]
        rows = c["N"] if args.config != "c4" else 8192
        fn, restatement = cpu_small_step_fn(args.config, rows)
        t_probe = time_cpu_steps(fn, 1, 1)[0]
        n_run = K if t_probe * (K + W) <= max(args.ref_seconds, 30.0) else max(3, int(args.ref_seconds / t_probe) - W)
        times = time_cpu_steps(fn, n_run, W)
        ms = 1e3 * sum(times) / len(times)
        frac = rows / c["N"]
        value = frac / (ms * 1e-3)
        sample = f"{rows} of {c['N']} rows per step ({len(times)} timed steps, mean)"
        config, unit = small_config(args.config), f"steps/s ({args.config})"
        K = len(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample, "sample_fraction": frac,
                         "ms_per_sampled_step": ms, "restatement": restatement},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "impl_detail": {"note": "reference = CPU restatement of bijax/advi.py:80-95 + value_and_grad (oracle/, torch-CPU fp32 autograd, "
                                f"{cores} threads); jax/tfp/optax are not installable in this image so bijax itself cannot run. "
                                "ms_per_step is the measured time of one SAMPLED step; value = sample_fraction / that time."},
    }
    print(json.dumps(line), flush=True)


def small_config(cfg: str):
    c = SMALL[cfg]
    l2 = {"c1": "flushed between timed steps (256 MB memset outside the per-step event pairs); the step's data is 400 B",
          "c2": "flushed between timed steps (256 MB memset outside the per-step event pairs); X = 20 MB would otherwise be L2-resident",
          "c4": "inputs_larger_than_l2 (X = 8.2 GB, read once per pass)"}[cfg]
    return {"workload": c["workload"], "rows_per_gpu": c["N"], "rows_total": c["N"], "D": c["D"], "S": c["S"], "l2": l2,
            "threefry_layout": "legacy"}


def launches_per_step(kernel: str) -> int:
    # k_prologue + likelihood + k_reduce_partials + k_tail (mean field: its last block computes the loss and d kwargs, the former
    # k_scalars launch); the tiny coin-toss step is one launch
    # (two-pass GEMM path: + operand split of theta, pass 1, pass 2 instead of one likelihood launch, two partial reductions)
    return {"tcgen05_bf16_split": 4, "fp32_smalld": 4, "fp32_tiled": 5, "coin": 1, "tcgen05_gemm_two_pass": 7}[kernel]


def collect_kernel_ms(lib, nv, K):
    kms = (C.c_float * K)()
    kn = C.c_int32(0)
    nv.check(lib.bjx_profile_collect(kms, K, C.byref(kn)))
    nv.check(lib.bjx_profile_end())
    return [float(x) for x in kms[: kn.value]]


# ----------------------------------------------------------------------------------------------------------------------
def make_c3_data(br, torch, rows, row0, n_total, dev, planes_only=True, chunk=1 << 20):
    """The synthetic shard [row0, row0 + rows) of the n_total-row dataset (counter = global flat index, so a shard is a slice of
    the global stream).  planes_only: X is generated chunk by chunk and kept on the device ONLY as its split-bf16 operand planes
    (bj.PreparedX: 4 bytes per element, the footprint of fp32 -- the dataset is resident once, not twice); otherwise the fp32
    matrix is returned (the raw-fp32 kernels, or an engine that prepares the planes itself)."""
    import bijax_b200 as bj

    D = D_C3
    w_true = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    u = br.uniform(br.PRNGKey(1236), (rows,), partitionable=True, offset=row0, total=n_total)
    y = torch.empty(rows, dtype=torch.float32, device=dev)
    X = bj.PreparedX(rows, D, dev) if planes_only else torch.empty((rows, D), dtype=torch.float32, device=dev)
    buf = torch.empty((min(chunk, rows), D), dtype=torch.float32, device=dev) if planes_only else None
    for i in range(0, rows, chunk):
        n = min(chunk, rows - i)
        blk = buf[:n] if planes_only else X[i : i + n]
        br.normal(br.PRNGKey(1234), (n * D,), partitionable=True, offset=(row0 + i) * D, total=n_total * D, out=blk.view(-1))
        blk.mul_(1.0 / math.sqrt(D))
        y[i : i + n] = (u[i : i + n] < torch.sigmoid(blk @ w_true)).float()
        if planes_only:
            X.append(blk)
    return X, y


def multi_gpu_check(torch, dist, br, engine, p2p, world, rank, dev):
    """Correctness of the exchange and of row sharding on THIS box, before anything is timed:
    (1) the peer-memory all-reduce kernel vs the rank-ordered sum of the gathered vectors (bit-equal) and vs NCCL;
    (2) a 2M-row c3-shaped problem sharded over the ranks, summed by the exchange, vs the same step on one GPU."""
    res = {}
    n = 1 + 2 * D_C3
    v = br.normal(br.PRNGKey(100 + rank), (n,), partitionable=True) * (1.0 + rank)
    gathered = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(gathered, v)
    ordered = gathered[0].clone()
    for g in gathered[1:]:
        ordered += g
    nccl = v.clone()
    dist.all_reduce(nccl)
    if p2p is not None:
        mine = p2p(v.clone())
        res["p2p_bit_equal_to_rank_ordered_sum"] = bool(torch.equal(mine, ordered))
        res["p2p_vs_nccl_max_abs_diff"] = float((mine - nccl).abs().max())
    res["nccl_vs_rank_ordered_sum_max_abs_diff"] = float((nccl - ordered).abs().max())
    # (2) sharded step == single-GPU step
    n_chk, S = 2_000_000, S_C3
    lo, hi = rank * n_chk // world, (rank + 1) * n_chk // world
    Xs, ys = make_c3_data(br, torch, hi - lo, lo, n_chk, dev, planes_only=False)
    loc = 0.02 * br.normal(br.PRNGKey(5), (D_C3,), partitionable=True)
    raw = torch.full((D_C3,), math.log(0.1), device=dev)
    key = br.PRNGKey(77)
    part = engine.step(key, loc, raw, None, Xs, ys, 1.0, S, 1.0 if rank == 0 else 0.0, partitionable=False).clone()
    if p2p is not None:
        p2p(part)
    else:
        dist.all_reduce(part)
    engine.release_planes()
    del Xs, ys
    ok = torch.ones(1, device=dev)
    if rank == 0:
        Xf, yf = make_c3_data(br, torch, n_chk, 0, n_chk, dev, planes_only=False)
        single = engine.step(key, loc, raw, None, Xf, yf, 1.0, S, 1.0, partitionable=False).clone()
        engine.release_planes()
        del Xf, yf

        def rel(a, b):
            return float((a - b).abs().max() / b.abs().max())

        res["sharded_vs_single_gpu_rel_err"] = {"rows": n_chk, "loss": rel(part[:1], single[:1]),
                                                "d_loc": rel(part[1 : 1 + D_C3], single[1 : 1 + D_C3]),
                                                "d_raw_scale": rel(part[1 + D_C3 :], single[1 + D_C3 :])}
        worst = max(res["sharded_vs_single_gpu_rel_err"][k] for k in ("loss", "d_loc", "d_raw_scale"))
        res["ok"] = bool(worst < 5e-6 and res.get("p2p_bit_equal_to_rank_ordered_sum", True))
        ok[0] = 1.0 if res["ok"] else 0.0
    identical = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(identical, part)
    res["reduced_buffer_identical_on_every_rank"] = bool(all(torch.equal(identical[0], t) for t in identical))
    torch.cuda.empty_cache()
    return res


def run_c3(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import bijax_b200 as bj
    from bijax_b200 import _native as nv
    from bijax_b200 import distributions as tfd
    from bijax_b200 import likelihoods as lk
    from bijax_b200 import random as br

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    D, S = D_C3, S_C3
    rows = args.rows_per_gpu or (N_C3 if world == 1 else ROWS_PER_GPU_MULTI)
    n_total = rows * world
    row0 = rows * rank
    lib = nv.load()

    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model = bj.ADVI(prior, {}, lk.BernoulliLogits(), vi_type="mean_field", precision=args.precision,
                    prepare_dataset=not args.raw_fp32)
    engine, _ = model._engine_for({})
    out = torch.empty(1 + 2 * D, device=dev)

    p2p = None
    if world > 1 and not args.nccl_allreduce:
        from bijax_b200.parallel import P2PAllReduce

        try:
            p2p = P2PAllReduce(out.numel(), None, dev)  # raises on every rank together if any rank cannot map its peers
        except RuntimeError as exc:
            if rank == 0:
                print(f"[bench] peer-memory all-reduce unavailable ({exc}); using NCCL", file=sys.stderr, flush=True)
            p2p = None

    mg_check = None
    if world > 1 and not args.no_multi_gpu_check:
        mg_check = multi_gpu_check(torch, dist, br, engine, p2p, world, rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(t):
        if p2p is not None:
            p2p(t)
        else:
            dist.all_reduce(t)

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def measure(rows_, row0_, n_total_, want_e2e, want_sustained):
        """Data of `rows_` rows -> prepare -> W warm-ups -> K timed steps (+ e2e, + the >= 1 s sustained region)."""
        torch.cuda.synchronize()
        tp0 = time.perf_counter()
        # once per dataset, outside the step: X -> split-bf16 planes (4 B/element); with planes_only fp32 X is never resident
        X, y = make_c3_data(br, torch, rows_, row0_, n_total_, dev, planes_only=not args.raw_fp32)
        kernel = engine.kernel_choice(S, rows_)
        prepared = engine.prepare(X, y, S)
        torch.cuda.synchronize()
        prepare_ms = 1e3 * (time.perf_counter() - tp0)
        dataset_gb = (X.nbytes if prepared and not args.raw_fp32 else 4.0 * rows_ * D) / 1e9
        loc = torch.zeros(D, device=dev)
        raw = torch.full((D,), math.log(0.1), device=dev)
        ll_scale = 1.0  # full_data_size / len(outputs), both global
        prior_weight = 1.0 if rank == 0 else 0.0
        keys = br.split(br.PRNGKey(0), W + K + 2048)  # per-step key = split(seed, n)[i]  (utils.py:30)

        def step(i):
            engine.step(keys[i % keys.shape[0]], loc, raw, None, X, y, ll_scale, S, prior_weight, out=out, partitionable=False)
            if world > 1:
                allreduce(out)

        def timed(n_steps, first):
            nv.check(lib.bjx_profile_begin(n_steps))
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            ev0.record()
            for i in range(first, first + n_steps):
                step(i)
            ev1.record()
            barrier()
            t1 = time.perf_counter()
            kms = collect_kernel_ms(lib, nv, n_steps)
            return ev0.elapsed_time(ev1), sum(kms) / max(len(kms), 1), t0, t1

        for i in range(W):
            step(i)
        barrier()
        if rank == 0:
            time.sleep(0.3)
        elapsed_ms, kernel_ms, t0, t1 = timed(K, W)
        clocks = sampler.window(t0, t1) if rank == 0 else None
        t = torch.tensor([elapsed_ms, kernel_ms, -kernel_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r = {"kernel": kernel, "prepared": prepared, "prepare_ms": prepare_ms, "elapsed_ms": float(t[0]), "kernel_ms": float(t[1]),
             "kernel_ms_min_rank": -float(t[2]), "clocks": clocks, "loss": float(out[0]), "rows": rows_, "dataset_gb": dataset_gb}

        if want_e2e:
            params_host = torch.cat([loc.cpu(), raw.cpu()]).pin_memory()
            out_host = torch.empty(1 + 2 * D).pin_memory()
            keys_host = keys.cpu().numpy()
            stage = torch.empty_like(out)
            keys_host = np.ascontiguousarray(keys_host, dtype=np.uint32)
            host_step = engine.host_stepper(X, y, ll_scale, S, prior_weight, partitionable=False) if world == 1 else None
            keys_pinned = torch.from_numpy(keys_host.astype(np.int64)).to(torch.uint32).pin_memory()
            kdev, pdev = torch.zeros(2, dtype=torch.uint32, device=dev), torch.zeros(2 * D, device=dev)
            dev_step = engine.stepper(kdev, pdev[:D], pdev[D:], None, X, y, ll_scale, S, stage, prior_weight, partitionable=False) if world > 1 else None

            def e2e_step(i):
                if world == 1:
                    # the reference-facing C-ABI call (bjx_elbo_step_host): host key/params in, packed loss+grads out, synchronous
                    host_step(keys_host[i], params_host, out_host)
                else:  # pinned key + params -> device, one C call for the step, the exchange kernel, packed result -> pinned host
                    kdev.copy_(keys_pinned[i], non_blocking=True)
                    pdev.copy_(params_host, non_blocking=True)
                    dev_step()
                    allreduce(stage)
                    out_host.copy_(stage, non_blocking=False)

            # same starting point as the device-timed region: a 1 kW part drops its clocks some 100 ms into a burst, and this
            # region would otherwise inherit the previous one's power state -- let the device idle, then warm up again
            barrier()
            time.sleep(float(os.environ.get("BJX_BENCH_IDLE_S", "1.0")))
            for i in range(W + int(os.environ.get("BJX_BENCH_EXTRA_WARMUP", "0"))):
                e2e_step(i % W)
            barrier()
            nv.check(lib.bjx_profile_begin(K))
            tt0 = time.perf_counter()
            for i in range(W, W + K):
                e2e_step(i)
            barrier()
            tt1 = time.perf_counter()
            tt = torch.tensor([tt1 - tt0], dtype=torch.float64, device=dev)
            e2e_kms = collect_kernel_ms(lib, nv, K)
            e2e_clocks = sampler.window(tt0, tt1) if rank == 0 else None
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            r["e2e"] = {"value": K / float(tt[0]) * n_total_ / N_C3, "unit": UNIT, "h2d_bytes_per_step": 8 + params_host.numel() * 4,
                        "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": 1e3 * float(tt[0]) / K,
                        "kernel_ms": (sum(e2e_kms) / len(e2e_kms)) if e2e_kms else None,
                        "kernel_ms_min_max": [min(e2e_kms), max(e2e_kms)] if e2e_kms else None, "clocks": e2e_clocks,
                        "note": "host key+params -> device, loss+grads -> host every step (bjx_elbo_step_host at N=1, wall clock); X, y are "
                                "the device-resident dataset, as they are for the reference's jitted scan (utils.py:22-31); kernel_ms = the "
                                "likelihood kernel inside THIS region (this rank), so ms_per_step - kernel_ms is the whole host + copy + small-kernel share"}
        if want_sustained:
            n_sus = max(K, min(2000, int(math.ceil(1200.0 / max(r["elapsed_ms"] / K, 1e-3)))))
            s_ms, s_kernel_ms, s0, s1 = timed(n_sus, W + K)
            ts = torch.tensor([s_ms, s_kernel_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            r["sustained"] = {"steps": n_sus, "seconds": float(ts[0]) * 1e-3, "kernel_ms": float(ts[1]), "ms_per_step": float(ts[0]) / n_sus,
                              "clocks": sampler.window(s0, s1) if rank == 0 else None}
        del X, y
        engine.release_planes()
        torch.cuda.empty_cache()
        return r

    main = measure(rows, row0, n_total, not args.no_e2e, not args.no_sustained)
    same_shard = None
    if world == 1 and rows == N_C3 and not args.no_same_shard:
        same_shard = measure(ROWS_PER_GPU_MULTI, 0, ROWS_PER_GPU_MULTI, False, False)

    if rank == 0:
        sampler.stop()
        K_ms = main["elapsed_ms"] / K
        steps_per_s = 1e3 / K_ms
        value = steps_per_s * n_total / N_C3
        peaks = load_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg_bytes = 4.0 * rows * D + 4.0 * rows  # X once + y once, per launch
        achieved = alg_bytes / (main["kernel_ms"] * 1e-3) / 1e9
        kernel, prepared = main["kernel"], main["prepared"]
        traffic, traffic_source = None, None
        for tp in sorted((ROOT / "profiles").glob("r0*_traffic.json"), reverse=True):
            if kernel == "tcgen05_bf16_split" and rows == N_C3:
                ent = json.loads(tp.read_text()).get("k_lik_tc_tma<8>" if prepared else "k_lik_tc<8>", {})
                if ent.get("traffic_bytes_per_launch"):
                    traffic = ent["traffic_bytes_per_launch"]
                    traffic_source = (f"profiles/{tp.name}: dram__bytes_read.sum + dram__bytes_write.sum of one committed ncu --set full "
                                      "capture of this kernel on this workload; NOT measured in this run")
                    break
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "traffic_source": traffic_source,
                    "kernel": kernel + (" (k_lik_tc_tma: TMA-fed, prepared planes)" if prepared and kernel == "tcgen05_bf16_split" else ""),
                    "kernel_ms": main["kernel_ms"], "algorithmic_bytes": alg_bytes, "peak_source": peak_src}
        if world > 1:
            roofline["kernel_ms_slowest_rank"] = main["kernel_ms"]
            roofline["kernel_ms_fastest_rank"] = main["kernel_ms_min_rank"]
            roofline["non_kernel_ms_per_step"] = K_ms - main["kernel_ms"]
        if "sustained" in main:
            s = main["sustained"]
            s_ach = alg_bytes / (s["kernel_ms"] * 1e-3) / 1e9
            roofline["sustained"] = {"achieved": s_ach, "frac": s_ach / hbm_peak, "kernel_ms": s["kernel_ms"], "steps": s["steps"],
                                     "seconds": s["seconds"], "ms_per_step": s["ms_per_step"],
                                     "value": 1e3 / s["ms_per_step"] * n_total / N_C3, "clocks": s["clocks"],
                                     "note": "the same kernel timed over a >= 1 s region (power-capped clocks); peak is the same burst copy figure"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = _cpu_setup()
            Xc, yc = cpu_c3_data(args.cpu_rows)
            times = cpu_c3_steps(Xc, yc, 30, 1, args.cpu_seconds)
            cms = 1e3 * float(np.median(times))
            frac = args.cpu_rows / N_C3
            cpu = {"value": frac / (cms * 1e-3), "unit": UNIT, "cores": cores, "kind": "port", "sample_fraction": frac,
                   "ms_per_sampled_step": cms,
                   "sample": f"{args.cpu_rows} rows (D=512, S=32) = {frac:.4f} of the workload per step, {len(times)} steps, median {cms:.1f} ms/step; "
                             "value = sample_fraction / time (work rate in the metric's unit)"}
        n_launch = launches_per_step(kernel) + (1 if world > 1 else 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": K_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": c3_config(world, rows),
            "impl_detail": {
                "precision": args.precision, "kernel": kernel,
                "exchange": ("none (one GPU)" if world == 1 else
                             "one-shot all-reduce kernel over NVLink peer memory (bjx_allreduce_packed_p2p); NCCL only bootstraps and barriers"
                             if p2p is not None else "NCCL all-reduce"),
                "dataset_layout": ("split-bf16 planes x1+x2 only (bj.PreparedX: 4 B/element, %.1f GB resident per GPU; generated + split chunk by "
                                   "chunk in %.0f ms, fp32 X never resident as a whole; TMA-fed kernel)" % (main["dataset_gb"], main["prepare_ms"])
                                   if prepared else "raw fp32 X (converted in registers every step)"),
            },
            "clocks": main["clocks"], "e2e": main.get("e2e"), "gpu_launches": n_launch * K,
            "roofline": roofline, "cpu_baseline": cpu,
            "steps_per_s_actual": steps_per_s, "snd_per_s": steps_per_s * S * n_total * D, "loss_last_step": main["loss"],
        }
        if mg_check is not None:
            line["multi_gpu_check"] = mg_check
        if same_shard is not None:
            ss_ms = same_shard["elapsed_ms"] / K
            line["n1_same_shard_value"] = 1e3 / ss_ms * ROWS_PER_GPU_MULTI / N_C3
            line["n1_same_shard"] = {"rows": ROWS_PER_GPU_MULTI, "ms_per_step": ss_ms, "kernel_ms": same_shard["kernel_ms"],
                                     "roofline_frac": (4.0 * ROWS_PER_GPU_MULTI * (D + 1)) / (same_shard["kernel_ms"] * 1e-3) / 1e9 / hbm_peak,
                                     "clocks": same_shard["clocks"],
                                     "note": "one GPU on the per-GPU shard of the N>1 runs (12.5M rows): the weak-scaling anchor"}
        print(json.dumps(line), flush=True)
    if p2p is not None:
        p2p.close()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
def run_small(args):
    """c1 / c2 / c4 at full size on one GPU (rank 0 of a torchrun launch; the other ranks exit)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    import torch

    import bijax_b200 as bj
    from bijax_b200 import _native as nv
    from bijax_b200 import bijectors as tfb
    from bijax_b200 import distributions as tfd
    from bijax_b200 import likelihoods as lk
    from bijax_b200 import random as br

    cfg = args.config
    c = SMALL[cfg]
    D, S, N = c["D"], c["S"], c["N"]
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib = nv.load()
    K, W = args.steps, max(args.warmup, 3)
    keys = br.split(br.PRNGKey(0), W + K)
    kw = None
    if cfg == "c1":
        y = torch.tensor([1.0 if n % 10 < 7 else 0.0 for n in range(N)], device=dev)
        X = None
        model = bj.ADVI({"p_of_heads": tfd.Beta(0.5, 0.5)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
        engine, _ = model._engine_for({})
        loc, raw = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        P = 1
    else:
        X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True)
        X.mul_(1.0 / math.sqrt(D))
        w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
        y = X @ w + 0.1 * br.normal(br.PRNGKey(1237), (N,), partitionable=True)
        prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
        kw = torch.tensor([math.log(0.1)], device=dev)
        vi = "full_rank" if cfg == "c4" else "mean_field"
        model = bj.ADVI(prior, {}, lk.GaussianLinear(), vi_type=vi, precision=args.precision)
        engine, _ = model._engine_for({"log_noise_scale": kw})
        loc = torch.zeros(D, device=dev)
        raw = 0.1 * torch.eye(D, device=dev) if cfg == "c4" else torch.full((D,), math.log(0.1), device=dev)
        P = D * D if cfg == "c4" else D
    nk = 0 if kw is None else 1
    out = torch.empty(1 + D + P + nk, device=dev)
    kernel = engine.kernel_choice(S, N)
    if X is not None:
        engine.prepare(X, y, S)

    def step(i):
        engine.step(keys[i], loc, raw, kw, X, y, 1.0, S, out=out, partitionable=False)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if cfg in ("c1", "c2") else None
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    nv.check(lib.bjx_profile_begin(K))
    t0 = time.perf_counter()
    if flush is None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(W, W + K):
            step(i)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / K
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for j, i in enumerate(range(W, W + K)):
            flush.zero_()  # evict L2 (126 MB) outside the timed pair
            evs[j][0].record()
            step(i)
            evs[j][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / K
    t1 = time.perf_counter()
    kms = collect_kernel_ms(lib, nv, K)
    clocks = sampler.window(t0, t1)
    kernel_ms = sum(kms) / len(kms) if kms else None
    value = 1e3 / ms
    unit = f"steps/s ({cfg})"

    # e2e through the host-buffer C-ABI entry point (key + params in, loss + grads out, synchronous)
    e2e = None
    if not args.no_e2e:
        params_host = torch.cat([loc.reshape(-1).cpu(), raw.reshape(-1).cpu()] + ([kw.cpu()] if kw is not None else [])).pin_memory()
        out_host = torch.empty(out.numel()).pin_memory()
        keys_host = keys.cpu().numpy()
        for i in range(3):
            engine.step_host(keys_host[i], params_host, out_host, X, y, 1.0, S, partitionable=False)
        torch.cuda.synchronize()
        tt0 = time.perf_counter()
        for i in range(W, W + K):
            engine.step_host(keys_host[i], params_host, out_host, X, y, 1.0, S, partitionable=False)
        torch.cuda.synchronize()
        e_ms = 1e3 * (time.perf_counter() - tt0) / K
        e2e = {"value": 1e3 / e_ms, "unit": unit, "h2d_bytes_per_step": 8 + params_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
               "ms_per_step": e_ms, "note": "bjx_elbo_step_host: host key+params -> device, loss+grads -> host, synchronous; X, y device-resident"}

    peaks = load_peaks()
    extra = {}
    if cfg == "c4":
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        issued = 3 * 4.0 * S * N * D / (kernel_ms * 1e-3) / 1e12  # 3 bf16 products per fp32 product (2-term split)
        roofline = {"bound": "tensor", "achieved": issued, "peak": peak, "unit": "TFLOP/s", "frac": issued / peak, "traffic": None,
                    "kernel": "k_gemm_tc two passes (EPI_LIK1 + EPI_LIK2), cta_group::2", "kernel_ms": kernel_ms,
                    "useful_tflops_fp32_equivalent": 4.0 * S * N * D / (kernel_ms * 1e-3) / 1e12,
                    "issued_flops_per_step": 3 * 4.0 * S * N * D, "useful_flops_per_step": 4.0 * S * N * D + 4.0 * S * D * D,
                    "algorithmic_bytes": 4.0 * N * D + 4.0 * N,
                    "peak_source": ("measured (MEASURED_PEAKS.json bf16_tflops_sustained: cuBLAS bf16 back to back for 4 s; the kernel is timed inside a long step)"
                                    if "bf16_tflops_sustained" in peaks else "fallback 1400 TFLOP/s")}
        if "bf16_tflops" in peaks:  # the 20-step region is ~0.1 s, between the two cuBLAS figures: say both
            roofline["frac_of_burst_peak"] = issued / float(peaks["bf16_tflops"])
            roofline["burst_peak"] = float(peaks["bf16_tflops"])
    elif cfg == "c2":
        mhz = float(peaks.get("sm_max_mhz", 1965.0))
        peak = 148 * 128 * 2 * mhz * 1e6 / 1e12
        ach = 4.0 * S * N * D / (kernel_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                    "kernel": "k_lik_smalld", "kernel_ms": kernel_ms, "algorithmic_bytes": 4.0 * N * D + 4.0 * N,
                    "hbm_gbs_if_streamed": (4.0 * N * D + 4.0 * N) / (kernel_ms * 1e-3) / 1e9,
                    "peak_source": "derived: 148 SMs x 128 fp32 FMA lanes x 2 flops x sm_max_mhz (no measured fp32 figure in MEASURED_PEAKS.json); "
                                   "useful flops 4*S*N*D only"}
    else:
        # c1 is launch latency: also the back-to-back replay of a captured one-launch step, without the L2 flush
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                engine.step(keys[0], loc, raw, None, None, y, 1.0, S, out=out, partitionable=False)
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(20):
            g.replay()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(500):
            g.replay()
        ev1.record()
        torch.cuda.synchronize()
        extra["graph_replay_us_per_step_back_to_back"] = ev0.elapsed_time(ev1) * 1e3 / 500
        roofline = {"bound": "latency", "achieved": ms * 1e3, "peak": None, "unit": "us/step", "frac": None, "traffic": None,
                    "kernel": "k_tiny_coin_step (the whole step is one block, one launch)", "kernel_ms": kernel_ms,
                    "algorithmic_bytes": 4.0 * N, "peak_source": "none: 400 B of data, pure launch + dependency latency"}
    cpu = None
    if not args.no_cpu_baseline:
        cores = _cpu_setup()
        rows_cpu = N if cfg != "c4" else 8192
        fn, restatement = cpu_small_step_fn(cfg, rows_cpu)
        times = time_cpu_steps(fn, 30, 1, args.cpu_seconds)
        cms = 1e3 * float(np.median(times))
        frac = rows_cpu / N
        cpu = {"value": frac / (cms * 1e-3), "unit": unit, "cores": cores, "kind": "port", "sample_fraction": frac, "ms_per_sampled_step": cms,
               "sample": f"{rows_cpu} of {N} rows per step, {len(times)} steps, median {cms:.2f} ms/step", "restatement": restatement}
    sampler.stop()
    line = {
        "metric": METRIC, "value": value, "unit": unit, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": small_config(cfg), "impl_detail": {"precision": args.precision, "kernel": kernel},
        "clocks": clocks, "e2e": e2e, "gpu_launches": (15 if cfg == "c4" else launches_per_step(kernel)) * K,  # c4: prologue, k_eps, 5 operand splits, 4 GEMMs (L.eps, pass 1, pass 2, dM), two partial reductions, tail, scalars
        "roofline": roofline, "cpu_baseline": cpu, "snd_per_s": value * S * N * D, "loss_last_step": float(out[0]),
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c3":
        run_c3(args)
    else:
        run_small(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- ELBO+grad steps/sec of the B200 ADVI engine on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision auto|fp32|bf16x3|bf16x2]

One "step" = one evaluation of the negative Monte-Carlo ELBO and all its gradients at a fresh PRNG key
(bijax/advi.py:80-95 under jax.value_and_grad, bijax/utils.py:7,25), optimiser update excluded.

Workload (synthetic, generated on the device from threefry):
  N=1 : config c3 -- Bayesian logistic regression, mean-field, D=512, N=10M rows, S=32, X fp32 [N, D] row-major.
  N>1 : config c5 shard -- the same model with 12.5M rows per GPU (100M rows at 8 GPUs), rows sharded, one NCCL
        all-reduce of the packed [loss | grads] buffer per step.  Weak scaling: per-GPU rows are fixed.
`value` is the whole-job rate in c3-equivalent steps/s: (S*N_total*D processed per second) / (32 * 10M * 512); at N=1
it IS the step rate of c3.

--impl reference times the CPU restatement of the reference (oracle/advi_ref.py, torch-CPU fp32 autograd, all host
threads) on a bounded row sample of the same workload; JAX/TFP cannot be installed here.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--rows-per-gpu", type=int, default=0, help="override the per-GPU row count (0 = config default)")
    ap.add_argument("--cpu-rows", type=int, default=200_000, help="row sample of the CPU baseline")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline time budget at N=1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="N > 1: use NCCL for the exchange step instead of the one-kernel all-reduce over NVLink peer memory "
                         "(bjx_allreduce_packed_p2p, the default)")
    ap.add_argument("--raw-fp32", action="store_true",
                    help="do not keep the dataset as prepared split-bf16 planes: the kernel converts raw fp32 X in registers every step")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
def cpu_restatement_rate(rows: int, seconds: float, steps_cap: int, warmup: int = 1, X=None, y=None):
    """Times oracle/advi_ref.py's batched logistic restatement (fp32, torch CPU, all threads) on `rows` rows.
    Returns (c3-equivalent steps/s, ms per sampled step, steps timed, cores)."""
    import numpy as np
    import torch

    from oracle import advi_ref as R
    from oracle import threefry as tf

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if X is None:
        g = torch.Generator().manual_seed(1234)
        X = torch.randn(rows, D_C3, generator=g, dtype=torch.float32) / math.sqrt(D_C3)
        w = torch.randn(D_C3, generator=g, dtype=torch.float32)
        y = (torch.rand(rows, generator=g) < torch.sigmoid(X @ w)).float()
    loc = torch.zeros(D_C3, dtype=torch.float32)
    raw = torch.full((D_C3,), math.log(0.1), dtype=torch.float32)
    keys = tf.split(tf.prng_key(0), steps_cap + warmup, tf.LAYOUT_LEGACY)
    times = []
    t_begin = time.perf_counter()
    for i in range(steps_cap + warmup):
        eps = torch.from_numpy(tf.normal(keys[i], (S_C3, D_C3), tf.LAYOUT_LEGACY))
        t0 = time.perf_counter()
        R.logistic_meanfield_value_and_grad(loc, raw, X, y, float(N_C3), eps)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if i >= warmup + 2 and time.perf_counter() - t_begin > seconds:
            break
    ms = 1e3 * float(np.median(times))
    rate_c3 = (1e3 / ms) * rows / N_C3
    return rate_c3, ms, len(times), cores


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

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.lines:
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


# ----------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, ms, n, cores = cpu_restatement_rate(args.cpu_rows, max(args.cpu_seconds, 10.0), max(args.steps, 3), max(args.warmup, 1))
    sample = f"{args.cpu_rows} rows x D=512 x S=32 per step ({n} steps, median), scaled linearly in rows to N=10M"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": max(args.warmup, 1),
        "ms_per_step": ms * N_C3 / args.cpu_rows, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c3: Bayesian logistic regression, mean-field, D=512, N=10M, S=32 (CPU arm: bounded row sample)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = CPU restatement of bijax/advi.py:80-95 + value_and_grad (oracle/advi_ref.py, torch-CPU fp32 autograd); "
                "jax/tfp/optax are not installable in this image so bijax itself cannot run",
    }
    print(json.dumps(line), flush=True)


def launches_per_step(kernel: str) -> int:
    # k_prologue + likelihood + k_reduce_partials + k_tail + k_scalars (mean field); the tiny coin-toss step is one launch
    # (two-pass GEMM path: + operand split of theta, pass 1, pass 2 instead of one likelihood launch)
    return {"tcgen05_bf16_split": 5, "fp32_smalld": 5, "fp32_tiled": 6, "coin": 1, "tcgen05_gemm_two_pass": 7}[kernel]


def run_ours(args):
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

    # ---- synthetic data, generated on the device, shard-stable (counter = global flat index) ----------------------------
    X = torch.empty((rows, D), dtype=torch.float32, device=dev)
    br.normal(br.PRNGKey(1234), (rows * D,), partitionable=True, offset=row0 * D, total=n_total * D, out=X.view(-1))
    X.mul_(1.0 / math.sqrt(D))
    w_true = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    u = br.uniform(br.PRNGKey(1236), (rows,), partitionable=True, offset=row0, total=n_total)
    y = torch.empty(rows, dtype=torch.float32, device=dev)
    chunk = 1 << 20
    for i in range(0, rows, chunk):
        y[i : i + chunk] = (u[i : i + chunk] < torch.sigmoid(X[i : i + chunk] @ w_true)).float()
    del u

    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model = bj.ADVI(prior, {}, lk.BernoulliLogits(), vi_type="mean_field", precision=args.precision,
                    prepare_dataset=not args.raw_fp32)
    engine, _ = model._engine_for({})
    kernel = engine.kernel_choice(S, rows)
    # dataset preparation (once per dataset, outside the step): X -> split-bf16 operand planes, 4 bytes per element like fp32
    torch.cuda.synchronize()
    tp0 = time.perf_counter()
    prepared = engine.prepare(X, y, S)
    torch.cuda.synchronize()
    prepare_ms = 1e3 * (time.perf_counter() - tp0)
    loc = torch.zeros(D, device=dev)
    raw = torch.full((D,), math.log(0.1), device=dev)
    ll_scale = float(n_total) / float(n_total)  # full_data_size / len(outputs), both global
    prior_weight = 1.0 if rank == 0 else 0.0
    K, W = args.steps, max(args.warmup, 3)
    keys = br.split(br.PRNGKey(0), W + K)  # per-step key = split(seed, n)[i]  (utils.py:30)
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

    def allreduce(t):
        if p2p is not None:
            p2p(t)
        else:
            dist.all_reduce(t)

    def step(i):
        engine.step(keys[i], loc, raw, None, X, y, ll_scale, S, prior_weight, out=out, partitionable=False)
        if world > 1:
            allreduce(out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()

    lib = nv.load()
    nv.check(lib.bjx_profile_begin(K))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for i in range(W, W + K):
        step(i)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    elapsed_ms = ev0.elapsed_time(ev1)
    kms = (C.c_float * K)()
    kn = C.c_int32(0)
    nv.check(lib.bjx_profile_collect(kms, K, C.byref(kn)))
    nv.check(lib.bjx_profile_end())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    kernel_ms = sum(kms[: kn.value]) / max(kn.value, 1)
    t = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms = float(t[0]), float(t[1])
    loss_value = float(out[0])

    ms_per_step = elapsed_ms / K
    steps_per_s = 1e3 / ms_per_step
    value = steps_per_s * n_total / N_C3

    # ---- e2e: the same step through host buffers ---------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        params_host = torch.cat([loc.cpu(), raw.cpu()]).pin_memory()
        out_host = torch.empty(1 + 2 * D).pin_memory()
        keys_host = keys.cpu().numpy()
        h2d = 8 + params_host.numel() * 4
        d2h = out_host.numel() * 4
        stage = torch.empty_like(out)

        def e2e_step(i):
            if world == 1:
                # the reference-facing C-ABI call: host key/params in, packed loss+grads out, synchronous
                engine.step_host(keys_host[i], params_host, out_host, X, y, ll_scale, S, prior_weight, partitionable=False)
            else:
                p = params_host.to(dev, non_blocking=True)
                k = torch.from_numpy(keys_host[i].astype(np.int64)).to(torch.uint32).to(dev, non_blocking=True)
                engine.step(k, p[:D], p[D:], None, X, y, ll_scale, S, prior_weight, out=stage, partitionable=False)
                allreduce(stage)
                out_host.copy_(stage, non_blocking=False)

        for i in range(min(W, 3)):
            e2e_step(i)
        barrier()
        tt0 = time.perf_counter()
        for i in range(W, W + K):
            e2e_step(i)
        barrier()
        e2e_s = time.perf_counter() - tt0
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_rate = K / float(tt[0]) * n_total / N_C3
        e2e = {"value": e2e_rate, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * float(tt[0]) / K,
               "note": "host key+params -> device, loss+grads -> host every step (bjx_elbo_step_host at N=1); X, y are the "
                       "device-resident dataset, as they are for the reference's jitted scan (utils.py:22-31)"}

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg_bytes = 4.0 * rows * D + 4.0 * rows  # X once + y once, per launch
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None  # dram__bytes_read+write per launch from the committed ncu --set full capture of this workload
        tp = ROOT / "profiles" / "r01_traffic.json"
        if tp.exists() and kernel == "tcgen05_bf16_split" and rows == N_C3:
            traffic = json.loads(tp.read_text()).get("k_lik_tc_tma<8>" if prepared else "k_lik_tc<8>", {}).get("traffic_bytes_per_launch")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "kernel": kernel + (" (k_lik_tc_tma: TMA-fed, prepared planes)" if prepared and kernel == "tcgen05_bf16_split" else ""), "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            xs = X[: args.cpu_rows].cpu()
            ys = y[: args.cpu_rows].cpu()
            rate, cms, n, cores = cpu_restatement_rate(args.cpu_rows, args.cpu_seconds, 30, 1, xs, ys)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {args.cpu_rows} rows of the same X (D=512, S=32), {n} steps, median {cms:.1f} ms/step, scaled linearly in rows to N=10M"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": ("c3: Bayesian logistic regression, mean-field, D=512, N=10M, S=32 on 1 B200" if world == 1 else
                             f"c5 shard: Bayesian logistic regression, mean-field, D=512, S=32, {rows} rows per GPU x {world} GPUs, NCCL all-reduce of [loss|grads]"),
                "rows_per_gpu": rows, "rows_total": n_total, "D": D, "S": S, "precision": args.precision, "kernel": kernel,
                "l2": "inputs_larger_than_l2 (X shard is >= 20 GB, read once per step)", "threefry_layout": "legacy",
                "exchange": ("none (one GPU)" if world == 1 else
                             "one-shot all-reduce kernel over NVLink peer memory (bjx_allreduce_packed_p2p)" if p2p is not None else "NCCL all-reduce"),
                "dataset_layout": ("split-bf16 planes x1+x2 (4 B/element, prepared once per dataset in %.1f ms, TMA-fed kernel)" % prepare_ms
                                   if prepared else "raw fp32 X (converted in registers every step)"),
            },
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step(kernel) * K,
            "roofline": roofline, "cpu_baseline": cpu,
            "steps_per_s_actual": steps_per_s, "snd_per_s": steps_per_s * S * n_total * D, "loss_last_step": loss_value,
        }
        print(json.dumps(line), flush=True)
    if p2p is not None:
        p2p.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

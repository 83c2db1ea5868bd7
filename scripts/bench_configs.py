"""Step times of the other BASELINE.json configs (c1, c2, reduced c4) beside the CPU restatement (oracle, all threads).
These are parity-test cases, not the headline bench line (bench.py); results are recorded under profiles/."""
import json, math, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods as lk, random as br
from oracle import advi_ref as R, threefry as tf

dev = torch.device("cuda")
torch.set_num_threads(os.cpu_count())

def time_gpu(fn, steps=200, warmup=10):
    for i in range(warmup): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): fn(warmup + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / steps  # us

def time_cpu(fn, seconds=5.0, warmup=1):
    for _ in range(warmup): fn()
    ts = []
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end or len(ts) < 3:
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))

out = []
keys = br.split(br.PRNGKey(0), 400)

# ---- c1: coin toss, Beta(0.5,0.5) + Sigmoid, 100 Bernoulli observations, S=16
y = torch.tensor([1.0 if n % 10 < 7 else 0.0 for n in range(100)], device=dev)
m = bj.ADVI({"p_of_heads": tfd.Beta(0.5, 0.5)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
eng, _ = m._engine_for({})
loc, raw = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
o = torch.empty(3, device=dev)
us = time_gpu(lambda i: eng.step(keys[i], loc, raw, None, None, y, 1.0, 16, out=o))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.step(keys[0], loc, raw, None, None, y, 1.0, 16, out=o)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        eng.step(keys[0], loc, raw, None, None, y, 1.0, 16, out=o)
us_graph = time_gpu(lambda i: g.replay())
ref = R.ADVIRef({"p": R.Beta(0.5, 0.5)}, {"p": R.Sigmoid()}, R.bernoulli_probs_loglik("p"))
yc = y.cpu(); epsc = torch.tensor(tf.normal(tf.prng_key(1), (16, 1), 0))
cpu = time_cpu(lambda: ref.value_and_grad(torch.zeros(1), torch.zeros(1), {}, yc, None, 100, epsc), 2.0)
out.append({"config": "c1 coin toss D=1 N=100 S=16", "kernel": eng.kernel_choice(16, 100), "gpu_us_per_step": us, "gpu_us_per_step_cuda_graph": us_graph,
            "cpu_us_per_step": cpu, "bound": "launch latency (1 launch: k_tiny_coin_step)"})

# ---- c2: linear regression D=5 + learnable log_noise_scale, N=1M, S=64
N, D, S = 1_000_000, 5, 64
X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
yl = X @ w + 0.1 * br.normal(br.PRNGKey(1237), (N,), partitionable=True)
m = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}, {}, lk.GaussianLinear())
lns = torch.tensor([math.log(0.1)], device=dev)
eng, _ = m._engine_for({"log_noise_scale": lns})
loc, raw = torch.zeros(D, device=dev), torch.full((D,), math.log(0.1), device=dev)
o = torch.empty(1 + 2 * D + 1, device=dev)
us = time_gpu(lambda i: eng.step(keys[i], loc, raw, lns, X, yl, 1.0, S, out=o))
g2 = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    eng.step(keys[0], loc, raw, lns, X, yl, 1.0, S, out=o)
    torch.cuda.synchronize()
    with torch.cuda.graph(g2, stream=s):
        eng.step(keys[0], loc, raw, lns, X, yl, 1.0, S, out=o)
us_graph2 = time_gpu(lambda i: g2.replay())
import ctypes as C
from bijax_b200 import _native as nv
lib = nv.load(); nv.check(lib.bjx_profile_begin(64))
for i in range(64): eng.step(keys[i], loc, raw, lns, X, yl, 1.0, S, out=o)
kms = (C.c_float * 64)(); kn = C.c_int32(0); nv.check(lib.bjx_profile_collect(kms, 64, C.byref(kn))); nv.check(lib.bjx_profile_end())
k_us = 1e3 * sorted(kms[:kn.value])[kn.value // 2]
Xc, yc = X.cpu(), yl.cpu()
epsc = torch.tensor(tf.normal(tf.prng_key(1), (S, D), 0))
cpu = time_cpu(lambda: R.linreg_meanfield_value_and_grad(torch.zeros(D), torch.full((D,), math.log(0.1)), torch.tensor(math.log(0.1)), Xc, yc, float(N), epsc), 5.0)
flops = 4.0 * S * N * D
out.append({"config": "c2 lin-reg D=5(+log_noise_scale) N=1M S=64", "kernel": eng.kernel_choice(S, N), "gpu_us_per_step": us, "gpu_us_per_step_cuda_graph": us_graph2,
            "likelihood_kernel_us": k_us, "cpu_us_per_step": cpu, "gflops_fp32_kernel": flops / k_us / 1e3, "bound": "fp32 FMA / L2 (X = 20 MB is L2-resident; not larger than L2)",
            "algorithmic_bytes": 4.0 * N * D + 4.0 * N, "gbs": (4.0 * N * D + 4.0 * N) / us / 1e3})

# ---- the training loop itself (train_fn, utils.py:22-31): ELBO + gradients + Adam per step, device-resident (CUDA graphs)
from functools import partial
from bijax_b200 import optim
def time_train(model, params, loss_kw, n):
    lf = partial(model.loss_fn, **loss_kw)
    bj.train(lf, params, optim.adam(1e-3), n_epochs=64, seed=br.PRNGKey(5)); torch.cuda.synchronize()
    t0 = time.perf_counter(); bj.train(lf, params, optim.adam(1e-3), n_epochs=n, seed=br.PRNGKey(5)); torch.cuda.synchronize()
    fused = 1e6 * (time.perf_counter() - t0) / n
    nh = max(n // 10, 50)
    t0 = time.perf_counter(); bj.train(lf, params, optim.adam(1e-3), n_epochs=nh, seed=br.PRNGKey(5), fused=False); torch.cuda.synchronize()
    return fused, 1e6 * (time.perf_counter() - t0) / nh
m1 = bj.ADVI({"p_of_heads": tfd.Beta(0.5, 0.5)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
f1, h1 = time_train(m1, m1.init(br.PRNGKey(0)), dict(outputs=y, inputs=None, full_data_size=100, n_samples=16), 4000)
out.append({"config": "c1 training loop (step + Adam), 4000 epochs", "fused_loop_us_per_epoch": f1, "host_driven_loop_us_per_epoch": h1})
m2 = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}, {}, lk.GaussianLinear())
p2 = m2.init(br.PRNGKey(0)); p2["log_noise_scale"] = torch.tensor(math.log(0.1), device=dev)
f2, h2 = time_train(m2, p2, dict(outputs=yl, inputs={"X": X}, full_data_size=N, n_samples=S), 2000)
out.append({"config": "c2 training loop (step + Adam), 2000 epochs", "fused_loop_us_per_epoch": f2, "host_driven_loop_us_per_epoch": h2})
# c4 at full size: scripts/bench_c4.py
print(json.dumps({"cores": os.cpu_count(), "results": out}, indent=1))

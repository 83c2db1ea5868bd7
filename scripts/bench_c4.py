"""config c4: full-rank ADVI Bayesian linear regression D=2048, N=1M, S=256 (BASELINE.json configs[3]) on one B200.
Reports the step time, the streaming-likelihood share, fp32-equivalent TFLOP/s (4*S*N*D + 4*S*D^2 useful flops per step),
the bf16 tensor-core rate actually issued (3 products per useful product) against the measured dense bf16 peak, and a
parity check of the tensor-core step against the fp32 CUDA-core step on the same device data."""
import ctypes as C, json, math, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import _native as nv, distributions as tfd, likelihoods as lk, random as br

N = int(os.environ.get("C4_N", 1_000_000)); D = int(os.environ.get("C4_D", 2048)); S = int(os.environ.get("C4_S", 256))
steps = int(os.environ.get("C4_STEPS", 10)); check = int(os.environ.get("C4_CHECK", 1))
dev = torch.device("cuda")
X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
y = X @ w + 0.1 * br.normal(br.PRNGKey(1237), (N,), partitionable=True)
prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
lns = torch.tensor([math.log(0.1)], device=dev)
loc, M = torch.zeros(D, device=dev), 0.1 * torch.eye(D, device=dev)
keys = br.split(br.PRNGKey(0), steps + 8)
lib = nv.load()

prepare = int(os.environ.get("C4_PREPARE", 1))

def engine(precision):
    m = bj.ADVI(prior, {}, lk.GaussianLinear(), vi_type="full_rank", precision=precision)
    e = m._engine_for({"log_noise_scale": lns})[0]
    e.prepare_dataset = bool(prepare)
    return e

eng = engine(os.environ.get("C4_PRECISION", "auto"))
out = torch.empty(1 + D + D * D + 1, device=dev)
for i in range(3): eng.step(keys[i], loc, M, lns, X, y, 1.0, S, out=out)
torch.cuda.synchronize()
nv.check(lib.bjx_profile_begin(steps))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps): eng.step(keys[3 + i], loc, M, lns, X, y, 1.0, S, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
kms = (C.c_float * steps)(); kn = C.c_int32(0)
nv.check(lib.bjx_profile_collect(kms, steps, C.byref(kn))); nv.check(lib.bjx_profile_end())
lik_ms = float(np.median(kms[: kn.value]))
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
flops = 4.0 * S * N * D + 4.0 * S * D * D
res = {"config": f"c4 full-rank lin-reg D={D} N={N} S={S}", "kernel": eng.kernel_choice(S, N), "ms_per_step": ms,
       "steps_per_s": 1e3 / ms, "likelihood_ms": lik_ms, "useful_tflops": flops / ms / 1e9,
       "bf16_tflops_issued_likelihood": 3 * 4.0 * S * N * D / lik_ms / 1e9,
       "bf16_peak_sustained": peaks.get("bf16_tflops_sustained"), "bf16_peak_burst": peaks.get("bf16_tflops"),
       "hbm_bytes_algorithmic": 4.0 * N * D + 4.0 * N, "workspace_gb": eng._ws.numel() / 1e9, "dataset_planes_prepared_once": bool(prepare),
       "planes_gb": (eng._planes.numel() / 1e9 if eng._planes is not None else 0.0)}
if peaks.get("bf16_tflops_sustained"):
    res["tensor_frac_of_sustained_peak"] = res["bf16_tflops_issued_likelihood"] / peaks["bf16_tflops_sustained"]
if check:
    key = keys[0]
    a = eng.step(key, loc, M, lns, X, y, 1.0, S).clone()
    b = eng.step(key, loc, M, lns, X, y, 1.0, S).clone()
    res["bit_reproducible"] = bool(torch.equal(a, b))
    ref = engine("fp32")
    t0 = time.perf_counter(); r = ref.step(key, loc, M, lns, X, y, 1.0, S).clone(); torch.cuda.synchronize()
    res["fp32_cuda_core_step_ms"] = 1e3 * (time.perf_counter() - t0)
    def rel(u, v): return float((u - v).abs().max() / v.abs().max())
    res["rel_err_vs_fp32_kernel"] = {"loss": rel(a[:1], r[:1]), "d_loc": rel(a[1:1 + D], r[1:1 + D]),
                                     "d_M": rel(a[1 + D:1 + D + D * D], r[1 + D:1 + D + D * D]), "d_log_noise": rel(a[-1:], r[-1:])}
print(json.dumps(res, indent=1))

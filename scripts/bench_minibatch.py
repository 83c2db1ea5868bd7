"""Minibatch path at the headline shape: logistic regression D=512, S=32, dataset N rows in
HBM (raw fp32 + prepared planes), batches of B rows drawn on the device.  Times the gather kernel against its HBM
roofline (B rows read + B rows written), the batch step, and a whole training epoch (gather -> step -> Adam) in the
device-resident loop vs the host-driven loop.  Results go to profiles/ by hand; not a bench.py line."""
import json, math, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods as lk, optim, random as br

dev = torch.device("cuda")
N = int(os.environ.get("MB_N", 2_000_000))
D, S = int(os.environ.get("MB_D", 512)), 32
peak = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
hbm = None
for k, v in (peak.items() if isinstance(peak, dict) else []):
    if "hbm" in k.lower() and isinstance(v, (int, float)):
        hbm = float(v)
X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
y = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < torch.sigmoid(X @ w)).float()
prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits())
model_raw = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits(), prepare_dataset=False)
ROUTES = {"gather4": (model, True), "gathered": (model, False), "raw_index": (model_raw, True)}

def ev_time(fn, steps, warmup=5):
    for i in range(warmup): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): fn(warmup + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / steps

res = []
keys = br.split(br.PRNGKey(0), 300)
for B, route in [(b, r) for b in tuple(int(v) for v in os.environ.get("MB_BATCHES", "4096,65536,524288").split(",")) for r in os.environ.get("MB_ROUTES", "gather4,gathered,raw_index").split(",")]:
    if B > N: continue
    model, copy_free = ROUTES[route]
    loss = model.minibatch_loss_fn(y, {"X": X}, batch_size=B, n_samples=S, copy_free=copy_free)
    params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.01 * br.normal(k, s))
    us_gather = ev_time(lambda i: loss.sample(params, keys[i]), 200)
    st = loss._state
    uses_planes = st["planes_b"] is not None
    eng = st["engine"]
    loc, raw = params["posterior"]["loc"], params["posterior"]["scale_diag"]
    o = torch.empty(1 + 2 * D, device=dev)
    Xin = st["X"] if st["indexed"] else st["Xb"]
    us_step = ev_time(lambda i: eng.step(keys[i], loc, raw, None, Xin, st["yb"], N / B, S, out=o,
                                         row_index=st["idx"] if st["indexed"] else None,
                                         planes=st["planes_full"] if st["indexed"] else st["planes_b"], n_full=N), 200)
    n = 40 * 16
    def fresh():
        return model.init(br.PRNGKey(0), initializer=lambda k, s: 0.01 * br.normal(k, s))
    bj.train(loss, fresh(), optim.adam(0.01), n_epochs=40, seed=br.PRNGKey(1))  # warm
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = bj.train(loss, fresh(), optim.adam(0.01), n_epochs=n, seed=br.PRNGKey(1))
    torch.cuda.synchronize(); us_fused = (time.perf_counter() - t0) * 1e6 / n
    nh = 100
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bj.train(loss, fresh(), optim.adam(0.01), n_epochs=nh, seed=br.PRNGKey(1), fused=False)
    torch.cuda.synchronize(); us_host = (time.perf_counter() - t0) * 1e6 / nh
    if st["indexed"]:
        gather_bytes = 3.0 * B * 4  # y read + written, idx written
        layout = ("none: the TMA-fed step gathers the batch's rows out of the full dataset's planes (cp.async)" if st["planes_full"] is not None
                  else "none: the step reads X[idx[r]] through BjxElboArgs.row_index (raw fp32 rows, converted in registers)")
    else:
        gather_bytes = 2.0 * B * D * 4 + 3.0 * B * 4  # rows read + written (planes: 2 x bf16 = 4 B/elem), y read + written, idx written
        layout = "split-bf16 planes (gathered plane rows)" if uses_planes else "raw fp32 rows (gathered)"
    step_bytes = 4.0 * B * D + 4.0 * B
    rec = {"N": N, "B": B, "D": D, "S": S, "route": route, "batch_copy": layout,
           "step_kernel": eng.kernel_choice(S, B), "gather_us": us_gather, "gather_algorithmic_bytes": gather_bytes,
           "gather_GBps": gather_bytes / us_gather * 1e-3, "step_us": us_step, "step_algorithmic_bytes": step_bytes,
           "step_GBps": step_bytes / us_step * 1e-3,
           "fused_loop_us_per_epoch": us_fused, "epoch_GBps_on_algorithmic_bytes": step_bytes / us_fused * 1e-3,
           "host_driven_loop_us_per_epoch": us_host, "default_for_this_batch_size": route == ("gather4" if B >= 32768 else "gathered"), "final_loss": float(r["losses"][-16:].mean())}
    if hbm:
        rec["epoch_frac_of_hbm_peak"] = rec["epoch_GBps_on_algorithmic_bytes"] / hbm
    res.append(rec)
    del loss
# full-batch epoch for scale
from functools import partial
lf = partial(model.loss_fn, outputs=y, inputs={"X": X}, full_data_size=N, n_samples=S)
bj.train(lf, fresh(), optim.adam(0.01), n_epochs=40, seed=br.PRNGKey(1))
torch.cuda.synchronize(); t0 = time.perf_counter()
r = bj.train(lf, fresh(), optim.adam(0.01), n_epochs=160, seed=br.PRNGKey(1))
torch.cuda.synchronize()
res.append({"N": N, "B": "full batch", "fused_loop_us_per_epoch": (time.perf_counter() - t0) * 1e6 / 160, "final_loss": float(r["losses"][-16:].mean())})
print(json.dumps({"hbm_peak_GBps": hbm, "results": res}, indent=1))

"""The fused single-pass kernel (k_lik_tc_tma, prepared planes; ENV_RAW=1: k_lik_tc on raw fp32 X) across its envelope: D in {128, 256, 384, 512}, S in {8, 32},
both likelihood families, X sized to ~10 GB so that every point streams from HBM (>> L2).  Reports ms/step and the
achieved algorithmic bandwidth (4*N*D + 4*N bytes per step) against the measured HBM peak.  Not a bench.py line."""
import json, math, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods as lk, random as br

dev = torch.device("cuda")
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs") if os.path.exists("MEASURED_PEAKS.json") else None
GB = float(os.environ.get("ENV_GB", 10))
res = []
keys = br.split(br.PRNGKey(0), 64)
for D in tuple(int(v) for v in os.environ.get("ENV_D", "128,256,384,512").split(",")):
    N = int(GB * 1e9 / (4 * D)) // 64 * 64
    X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
    w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    eta = X @ w
    y_logit = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < torch.sigmoid(eta)).float()
    y_lin = eta + 0.1 * br.normal(br.PRNGKey(1237), (N,), partitionable=True)
    del eta
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    for fam in ("bernoulli_logits", "gaussian_linear"):
        like = lk.BernoulliLogits() if fam == "bernoulli_logits" else lk.GaussianLinear()
        model = bj.ADVI(prior, {"weights": tfb.Identity()}, like, prepare_dataset=os.environ.get("ENV_RAW", "0") != "1")
        kw = {} if fam == "bernoulli_logits" else {"log_noise_scale": torch.tensor(math.log(0.1), device=dev)}
        eng, _ = model._engine_for(kw)
        y = y_logit if fam == "bernoulli_logits" else y_lin
        kwv = None if fam == "bernoulli_logits" else torch.tensor([math.log(0.1)], device=dev)
        loc = torch.zeros(D, device=dev); raw = torch.full((D,), math.log(0.1), device=dev)
        for S in (8, 32):
            o = torch.empty(1 + 2 * D + (0 if kwv is None else 1), device=dev)
            for i in range(4): eng.step(keys[i], loc, raw, kwv, X, y, 1.0, S, out=o)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20): eng.step(keys[4 + i], loc, raw, kwv, X, y, 1.0, S, out=o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            nbytes = 4.0 * N * D + 4.0 * N
            rec = {"D": D, "S": S, "N": N, "likelihood": fam, "kernel": eng.kernel_choice(S, N), "ms_per_step": ms,
                   "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / ms * 1e-6, "finite": bool(torch.isfinite(o).all())}
            if peak: rec["frac_of_hbm_peak"] = rec["GBps"] / peak
            res.append(rec)
        del model, eng
    del X, y_logit, y_lin
    torch.cuda.empty_cache()
print(json.dumps({"hbm_peak_GBps": peak, "results": res}, indent=1))

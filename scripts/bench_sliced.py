"""32 < S <= 96 on a narrow X: the fused single pass run once per slice of 32 samples (automatic choice) against the two-pass
tcgen05 GEMM path (forced), same data, 6 GB of X per point.  Not a bench.py line."""
import json, math, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import distributions as tfd, likelihoods as lk, random as br

peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs") if os.path.exists("MEASURED_PEAKS.json") else None
keys = br.split(br.PRNGKey(0), 64)
res = []
for D, S in ((128, 64), (128, 96), (256, 64), (384, 64), (512, 64), (200, 64)):
    N = int(6e9 / (4 * D)) // 64 * 64
    X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
    y = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < 0.5).float()
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    loc, raw = torch.zeros(D, device="cuda"), torch.full((D,), math.log(0.1), device="cuda")
    rec = {"D": D, "S": S, "N": N}
    for name, kernel in (("auto", None), ("two_pass", "tcgen05_gemm_two_pass")):
        eng, _ = bj.ADVI(prior, {}, lk.BernoulliLogits(), kernel=kernel)._engine_for({})
        o = torch.empty(1 + 2 * D, device="cuda")
        for i in range(4): eng.step(keys[i], loc, raw, None, X, y, 1.0, S, out=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): eng.step(keys[4 + i], loc, raw, None, X, y, 1.0, S, out=o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        rec[name] = {"kernel": eng.kernel_choice(S, N), "ms_per_step": ms, "frac_of_hbm_peak_on_one_read": (4.0 * N * D + 4.0 * N) / ms * 1e-6 / peak if peak else None}
        del eng
    rec["speedup_vs_two_pass"] = rec["two_pass"]["ms_per_step"] / rec["auto"]["ms_per_step"]
    res.append(rec)
    del X, y
    torch.cuda.empty_cache()
print(json.dumps({"hbm_peak_GBps": peak, "results": res}, indent=1))

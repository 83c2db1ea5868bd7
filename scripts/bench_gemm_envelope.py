"""The two-pass TMA-fed tcgen05 GEMM path (k_gemm_tc, prepared planes) across its envelope: S in {64, 128, 256, 1024},
D in {64, 192, 512, 1024, 2048}, Gaussian-linear likelihood with learnable noise (the c4 family, mean field here so that
the full-rank contractions do not blur the picture).  Per point: ms/step, bf16 flops issued (3 products x 4*S*N*D) against
the sustained bf16 peak, and the HBM traffic of the two passes (X planes twice, g written and read: 2 x 4*N*D + 2 x 4*N*S) against the copy peak -- the larger of the two
fractions says how close the point is to ITS roofline.  Not a bench.py line."""
import json, math, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods as lk, random as br

dev = torch.device("cuda")
pk = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
hbm, tf_peak = pk.get("hbm_gbs"), pk.get("bf16_tflops_sustained")
GB = float(os.environ.get("ENV_GB", 3))
res = []
keys = br.split(br.PRNGKey(0), 32)
for D in (64, 192, 512, 1024, 2048):
    N = int(GB * 1e9 / (4 * D)) // 256 * 256
    X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
    w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    y = X @ w + 0.1 * br.normal(br.PRNGKey(1237), (N,), partitionable=True)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.GaussianLinear(), kernel="tcgen05_gemm_two_pass")
    eng, _ = model._engine_for({"log_noise_scale": torch.tensor(0.0, device=dev)})
    kwv = torch.tensor([math.log(0.1)], device=dev)
    loc = torch.zeros(D, device=dev); raw = torch.full((D,), math.log(0.1), device=dev)
    for S in (64, 128, 256, 1024):
        o = torch.empty(2 + 2 * D, device=dev)
        try:
            for i in range(3): eng.step(keys[i], loc, raw, kwv, X, y, 1.0, S, out=o)
        except Exception as e:  # outside the envelope
            res.append({"D": D, "S": S, "N": N, "error": str(e)[:120]}); continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10): eng.step(keys[3 + i], loc, raw, kwv, X, y, 1.0, S, out=o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        flops = 3 * 4.0 * S * N * D
        nbytes = 2 * 4.0 * N * D + 2 * 4.0 * N * S  # X planes in both passes + g (split-bf16 planes) written by pass 1, read by pass 2
        rec = {"D": D, "S": S, "N": N, "kernel": eng.kernel_choice(S, N), "ms_per_step": ms, "bf16_TFLOPs_issued": flops / ms * 1e-9,
               "GBps_X_and_g": nbytes / ms * 1e-6, "finite": bool(torch.isfinite(o).all())}
        if tf_peak: rec["frac_of_bf16_sustained_peak"] = rec["bf16_TFLOPs_issued"] / tf_peak
        if hbm: rec["frac_of_hbm_peak"] = rec["GBps_X_and_g"] / hbm
        res.append(rec)
    del X, y, model, eng
    torch.cuda.empty_cache()
print(json.dumps({"hbm_peak_GBps": hbm, "bf16_sustained_TFLOPs": tf_peak, "results": res}, indent=1))

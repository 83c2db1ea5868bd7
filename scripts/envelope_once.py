"""Three steps of the fused kernel at D = 128 and D = 256 (ENV_D overrides; S = 32, Bernoulli-logit, 4 GB of X each, prepared
planes) for an ncu capture of k_lik_tc_tma<2, 8> / <4, 8> (the slot-ring variants) or, at D > 512, of the paired variant; no
timing here."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods as lk, random as br

dev = torch.device("cuda")
keys = br.split(br.PRNGKey(0), 8)
for D in tuple(int(v) for v in os.environ.get("ENV_D", "128,256").split(",")):
    N = int(4e9 / (4 * D)) // 64 * 64
    X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
    w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    y = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < torch.sigmoid(X @ w)).float()
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits())
    eng, _ = model._engine_for({})
    loc = torch.zeros(D, device=dev); raw = torch.full((D,), math.log(0.1), device=dev)
    for i in range(3):
        o = eng.step(keys[i], loc, raw, None, X, y, 1.0, 32)
    torch.cuda.synchronize()
    print(D, N, eng.kernel_choice(32, N), float(o[0]))
    del X, y, model, eng
    torch.cuda.empty_cache()

"""One c2 step (D=5 + learnable noise, N=1M, S=64) a few times -- the ncu target for k_lik_smalld."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import distributions as tfd, likelihoods as lk, random as br
dev = torch.device("cuda")
N, D, S = 1_000_000, 5, 64
X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
y = X @ w + 0.1 * br.normal(br.PRNGKey(1237), (N,), partitionable=True)
m = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}, {}, lk.GaussianLinear())
lns = torch.tensor([math.log(0.1)], device=dev)
eng, _ = m._engine_for({"log_noise_scale": lns})
loc, raw = torch.zeros(D, device=dev), torch.full((D,), math.log(0.1), device=dev)
keys = br.split(br.PRNGKey(0), 8)
for i in range(6):
    out = eng.step(keys[i], loc, raw, lns, X, y, 1.0, S)
torch.cuda.synchronize()
print("loss", out[0].item())

"""A few copy-free minibatch steps (cp.async gather out of the full dataset's planes) at c3's shape, for ncu captures."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods as lk, random as br

N, B, D, S = int(os.environ.get("MB_N", 1_000_000)), int(os.environ.get("MB_B", 524288)), int(os.environ.get("MB_D", 512)), 32
X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
y = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < 0.5).float()
prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits(), prepare_dataset=os.environ.get("MB_RAW", "0") != "1")
loss = model.minibatch_loss_fn(y, {"X": X}, batch_size=B, n_samples=S, copy_free=True)
params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.01 * br.normal(k, s))
keys = br.split(br.PRNGKey(0), 8)
for i in range(5):
    v = loss(params, keys[i])
torch.cuda.synchronize()
print("ok", float(v), loss._state["indexed"], loss._state["planes_full"] is not None)

"""A few minibatch epochs at D=512, S=32, B=512 Ki rows drawn from N=2M, in both modes (copy-free through the row index,
then gathered planes), for an ncu capture of k_minibatch_gather / k_lik_tc<8, true> / k_lik_tc_tma (no timing here)."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods as lk, random as br

N, D, S, B = 2_000_000, 512, 32, 524288
X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
y = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < torch.sigmoid(X @ w)).float()
prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits())
params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.01 * br.normal(k, s))
keys = br.split(br.PRNGKey(0), 8)
for copy_free in (True, False):
    loss = model.minibatch_loss_fn(y, {"X": X}, batch_size=B, n_samples=S, copy_free=copy_free)
    for i in range(4):
        l = loss(params, keys[i])
    torch.cuda.synchronize()
    print("copy_free", copy_free, float(l))

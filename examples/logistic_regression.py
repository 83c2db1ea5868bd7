"""Bayesian logistic regression (BASELINE config c3 at a reduced size) with the three variational families, minibatches, then MAP.

    python examples/logistic_regression.py [N] [D]          (needs a CUDA device)
"""
import math
import sys
import time
from functools import partial

import numpy as np
import torch

sys_path_root = __import__("pathlib").Path(__file__).resolve().parent.parent
__import__("sys").path.insert(0, str(sys_path_root))  # run from a checkout without installing
import bijax_b200 as bj
from bijax_b200 import distributions as tfd, likelihoods, optim, random as jr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda")
X = jr.normal(jr.PRNGKey(1234), (N, D), partitionable=True) / math.sqrt(D)
w_true = jr.normal(jr.PRNGKey(1235), (D,), partitionable=True)
y = (jr.uniform(jr.PRNGKey(1236), (N,), partitionable=True) < torch.sigmoid(X @ w_true)).float()
prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}

for vi_type, extra in (("mean_field", {}), ("low_rank", {"rank": 4}), ("full_rank", {})):
    model = bj.ADVI(prior, {}, likelihoods.BernoulliLogits(), vi_type=vi_type, **extra)
    small = lambda k, shape: 0.01 * jr.normal(k, shape)
    params = model.init(jr.PRNGKey(0), initializer=small)
    if vi_type == "full_rank":  # a well-conditioned start: 0.05 I + small noise
        params["posterior"]["scale_tril"] = params["posterior"]["scale_tril"] + 0.05 * torch.eye(D, device=dev)
    elif vi_type == "mean_field" or vi_type == "low_rank":
        params["posterior"]["scale_diag"] = params["posterior"]["scale_diag"] - 3.0  # sigma = exp(-3)
    loss_fn = partial(model.loss_fn, outputs=y, inputs={"X": X}, full_data_size=N, n_samples=32)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = bj.train(loss_fn, params, optim.adam(0.02), n_epochs=300, seed=jr.PRNGKey(1))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    w = res["params"]["posterior"]["loc"]
    cos = torch.nn.functional.cosine_similarity(w, w_true, dim=0).item()
    kern = next(iter(model._engines.values())).kernel_choice(32, N)
    print(f"{vi_type:10s}  300 epochs in {dt:6.2f} s ({1e3 * dt / 300:6.2f} ms/epoch, kernel {kern}),  -ELBO {res['losses'][-1].item():.1f},  cos(loc, w*) = {cos:.4f}")

# stochastic VI: minibatches drawn on the device from the step key (jax.random.choice semantics), full-data scaling
# of advi.py:92; the whole loop (draw -> ELBO + gradients -> Adam) stays on the device
model = bj.ADVI(prior, {}, likelihoods.BernoulliLogits())
params = model.init(jr.PRNGKey(0), initializer=lambda k, shape: 0.01 * jr.normal(k, shape))
params["posterior"]["scale_diag"] = params["posterior"]["scale_diag"] - 3.0
B = min(65536, N)
loss = model.minibatch_loss_fn(y, {"X": X}, batch_size=B, n_samples=32)
bj.train(loss, params, optim.adam(0.02), n_epochs=48, seed=jr.PRNGKey(1))  # one-time costs: plane preparation, kernel load, graph capture
torch.cuda.synchronize(); t0 = time.perf_counter()
res = bj.train(loss, params, optim.adam(0.02), n_epochs=1500, seed=jr.PRNGKey(1))
torch.cuda.synchronize(); dt = time.perf_counter() - t0
cos = torch.nn.functional.cosine_similarity(res["params"]["posterior"]["loc"], w_true, dim=0).item()
print(f"minibatch   1500 epochs of {B} rows in {dt:6.2f} s ({1e6 * dt / 1500:6.1f} us/epoch),  -ELBO estimate {res['losses'][-50:].mean().item():.1f},  cos(loc, w*) = {cos:.4f}")

m = bj.MAP(prior, likelihoods.BernoulliLogits(), None)
p = m.init(jr.PRNGKey(0))
res = bj.train(partial(m.neg_log_joint, outputs=y, inputs={"X": X}), p, optim.adam(0.05), n_epochs=300)
cos = torch.nn.functional.cosine_similarity(res["params"]["params"]["weights"], w_true, dim=0).item()
print(f"MAP         -log joint {res['losses'][-1].item():.1f},  cos(w_map, w*) = {cos:.4f}")

"""The reference's examples/advi/coin_toss.ipynb on the B200 engine: Beta(1, 2) prior on p_of_heads, Sigmoid bijector,
tosses [0, 1, 1, 1, 1, 0]; the exact posterior is Beta(5, 4) and the log evidence -4.941642.

    python examples/coin_toss.py          (needs a CUDA device)
"""
import math
from functools import partial

import torch

sys_path_root = __import__("pathlib").Path(__file__).resolve().parent.parent
__import__("sys").path.insert(0, str(sys_path_root))  # run from a checkout without installing
import bijax_b200 as bj
from bijax_b200 import bijectors as tfb, distributions as tfd, likelihoods, optim, random as jr

tosses = torch.tensor([0, 1, 1, 1, 1, 0], dtype=torch.float32, device="cuda")
prior = {"p_of_heads": tfd.Beta(1.0, 2.0)}
bijector = {"p_of_heads": tfb.Sigmoid()}
model = bj.ADVI(prior, bijector, likelihoods.BernoulliProbs("p_of_heads"), vi_type="mean_field")  # README.md:113

params = model.init(jr.PRNGKey(0))  # advi.py:77-78
loss_fn = partial(model.loss_fn, outputs=tosses, inputs=None, full_data_size=len(tosses), n_samples=64)
result = bj.train(loss_fn, params, optim.adam(0.05), n_epochs=1000, seed=jr.PRNGKey(1))  # utils.py:6-36, device-resident loop

final = result["losses"][-200:].mean().item()
print(f"-ELBO after training: {final:.4f}   (-log evidence = 4.9416; PyMC ADVI 4.9947, NumPyro SVI 4.9520)")
posterior = model.apply(result["params"])  # advi.py:97-100
samples = posterior.sample(jr.PRNGKey(2), (10000,))["p_of_heads"]
print(f"posterior mean {samples.mean().item():.4f} (exact Beta(5,4): {5 / 9:.4f}), sd {samples.std().item():.4f} (exact {math.sqrt(20 / (81 * 10)):.4f})")
grid = torch.tensor([0.3, 0.55, 0.8], device="cuda")
lp = posterior.log_prob({"p_of_heads": grid}, (3,))
exact = [math.log(280.0) + 4 * math.log(p) + 3 * math.log(1 - p) for p in grid.tolist()]
print("log q(p) at 0.3/0.55/0.8:", [round(v, 3) for v in lp.tolist()], " exact log Beta(5,4):", [round(v, 3) for v in exact])

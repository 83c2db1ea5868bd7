"""Compare the tcgen05 kernel and the fp32 CUDA-core kernel against a float64 evaluation on the GPU, as N grows."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import distributions as tfd, likelihoods as lk, random as br

D, S = 512, 32
dev = torch.device("cuda")
def engine(prec):
    m = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}, {}, lk.BernoulliLogits(), precision=prec)
    return m._engine_for({})[0]
tc, fp = engine("auto"), engine("fp32")
for N in (65536, 524288, 2_000_000, 8_000_000):
    X = br.normal(br.PRNGKey(1234), (N, D), partitionable=True); X.mul_(1 / math.sqrt(D))
    w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    y = (br.uniform(br.PRNGKey(1236), (N,), partitionable=True) < torch.sigmoid(X @ w)).float()
    loc = 0.02 * br.normal(br.PRNGKey(5), (D,), partitionable=True); raw = torch.full((D,), math.log(0.1), device=dev)
    key = br.PRNGKey(123)
    eps = torch.empty(S * D, device=dev)
    o_tc = tc.step(key, loc, raw, None, X, y, 1.0, S, eps_out=eps).double()
    o_fp = fp.step(key, loc, raw, None, X, y, 1.0, S).double()
    # float64 reference of d loc: sum_s [ z/S - (1/S) * sum_n (y - sigmoid(x.z_s)) x ]
    z = (loc.double() + torch.exp(raw.double()) * eps.double().reshape(S, D))
    G = torch.zeros(S, D, dtype=torch.float64, device=dev); ll = torch.zeros(S, dtype=torch.float64, device=dev)
    for i in range(0, N, 1 << 19):
        xb = X[i:i + (1 << 19)].double(); eta = xb @ z.T
        yb = y[i:i + (1 << 19)].double()[:, None]
        G += (yb - torch.sigmoid(eta)).T @ xb
        ll += (yb * eta - torch.nn.functional.softplus(eta)).sum(0)
    dloc = (z / S - G / S).sum(0)
    q = (-0.5 * eps.double().reshape(S, D) ** 2 - raw.double() - 0.5 * math.log(2 * math.pi)).sum(1)
    p = (-0.5 * z * z - 0.5 * math.log(2 * math.pi)).sum(1)
    loss = (q - p - ll).mean()
    sc = dloc.abs().max()
    print(f"N={N:9d}  d loc max err / max|d loc|:  tcgen05 {((o_tc[1:1+D]-dloc).abs().max()/sc).item():.2e}   fp32 {((o_fp[1:1+D]-dloc).abs().max()/sc).item():.2e}"
          f"   signed mean rel err tc {(((o_tc[1:1+D]-dloc)/dloc.abs()).mean()).item():+.2e} fp32 {(((o_fp[1:1+D]-dloc)/dloc.abs()).mean()).item():+.2e}"
          f"   loss rel err tc {((o_tc[0]-loss)/loss).item():+.2e} fp32 {((o_fp[0]-loss)/loss).item():+.2e}")
    del X, y

"""Dump the per-tile pipeline timeline of the tcgen05 likelihood kernel (CTA 0, tiles 16..23), in SM cycles."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bijax_b200 as bj
from bijax_b200 import _native as nv, distributions as tfd, likelihoods as lk, random as br

D, S, N = int(os.environ.get("TRACE_D", 512)), int(os.environ.get("TRACE_S", 32)), int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dev = torch.device("cuda")
X = br.normal(br.PRNGKey(1), (N, D), partitionable=True) / math.sqrt(D)
y = (torch.rand(N, device=dev) < 0.5).float()
model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}, {}, lk.BernoulliLogits())
eng, _ = model._engine_for({})
loc, raw = torch.zeros(D, device=dev), torch.full((D,), math.log(0.1), device=dev)
trace = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
for _ in range(3):
    eng.step(br.PRNGKey(0), loc, raw, None, X, y, 1.0, S)
nv.load().bjx_debug_set_trace_buffer(trace.data_ptr())
eng.step(br.PRNGKey(0), loc, raw, None, X, y, 1.0, S)
torch.cuda.synchronize()
nv.load().bjx_debug_set_trace_buffer(None)
t = trace.cpu().numpy().reshape(8, 64)
t0 = t[0, 0]
if eng._planes is not None:
    print("TMA-fed variant (prepared planes)")
names = {**{b: f"ld wait_empty b{b}" for b in range(8)}, **{8 + b: f"ld stored b{b}" for b in range(8)},
         **{16 + b: f"mma x_full b{b}" for b in range(8)}, 24: "mma gemm1 issued+commit", 25: "mma dl_full", 26: "mma gemm2 issued",
         27: "mma x_full b0 (before tcgen05 fence)", **{40 + w: f"ld stored b0 quarter {w}" for w in range(4)},
         **{44 + w: f"ld stored b1 quarter {w}" for w in range(4)},
         48: "ld b4 loads issued", 49: "ld b4 data arrived",
         32: "epi acc1_full", 33: "epi tmem loaded", 34: "epi math done", 35: "epi dl stored"}
for it in range(1, int(os.environ.get("TRACE_TILES", 3))):
    ev = sorted((int(t[it, k] - t0), names[k]) for k in names if t[it, k])
    print(f"--- tile {16 + it} (tile period {int(t[it,0]-t[it-1,0])} cycles)")
    base = ev[0][0]
    for c, n in ev:
        print(f"{c - base:8d}  {n}")

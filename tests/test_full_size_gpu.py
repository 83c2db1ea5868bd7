"""GPU: BASELINE.json's headline configurations at their FULL sizes against the float64 restatement of advi.py:80-95.

The default (``precision="auto"``) path of c3 / c5 is the fused tcgen05 kernel on a 2-term bf16 split of X, of c4 the
two-pass tcgen05 GEMM on the same split; both accumulate in TMEM.  Small-size parity does not bound the error of a
10M-row accumulation, so these tests evaluate the oracle's own formulas (``ADVIRef.value_and_grad_chunked``: the
reference's ``log_likelihood_fn`` on row blocks, float64, torch on the GPU as the checker's calculator) over the very
dataset the step ran on, and hold loss and every gradient block to north_star's rtol 1e-5
(``tol = max(rtol*|want|, 1e-5*max|want|)``, tests/cases.py)."""
import math

import numpy as np
import pytest
import torch

from bijax_b200 import distributions as tfd
from bijax_b200 import likelihoods as lk
from oracle import threefry as tf
from tests import cases as C

pytestmark = pytest.mark.gpu


def _device_X(rows, D, row0=0, n_total=None):
    """bench.py's design matrix (counter = global flat index, so a shard is a slice of the global stream)."""
    from bijax_b200 import random as br

    n_total = n_total or rows
    X = torch.empty((rows, D), dtype=torch.float32, device="cuda")
    br.normal(br.PRNGKey(1234), (rows * D,), partitionable=True, offset=row0 * D, total=n_total * D, out=X.view(-1))
    X.mul_(1.0 / math.sqrt(D))
    return X


def _logistic_y(X, row0=0, n_total=None):
    from bijax_b200 import random as br

    rows, D = X.shape
    w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    u = br.uniform(br.PRNGKey(1236), (rows,), partitionable=True, offset=row0, total=n_total or rows)
    y = torch.empty(rows, dtype=torch.float32, device=X.device)
    for i in range(0, rows, 1 << 20):
        y[i : i + (1 << 20)] = (u[i : i + (1 << 20)] < torch.sigmoid(X[i : i + (1 << 20)] @ w)).float()
    return y


def _check(packed, want, D, P, names):
    wl, wgl, wgr, wgk = want
    got = packed.cpu().numpy()
    C.assert_close(got[0], wl.item(), what="loss")
    C.assert_close(got[1 : 1 + D], wgl.cpu().numpy().ravel(), what="d loc")
    C.assert_close(got[1 + D : 1 + D + P], wgr.cpu().numpy().ravel(), what="d raw scale")
    for i, k in enumerate(names):
        C.assert_close(got[1 + D + P + i], wgk[k].item(), what=f"d {k}")


@pytest.mark.timeout(900)
@pytest.mark.parametrize("rows,row0,n_total,D", [(10_000_000, 0, 10_000_000, 512), (12_500_000, 37_500_000, 100_000_000, 512),
                                                  (5_000_000, 0, 5_000_000, 1024), (6_000_000, 0, 6_000_000, 768)],
                         ids=["c3_10M_rows", "c5_shard_3_of_8", "d1024_5M_rows_cta_pairs", "d768_6M_rows_cta_pairs"])
def test_logistic_d512_s32_full_size_against_the_float64_oracle(rows, row0, n_total, D):
    """c3 (N=10M on one GPU) and one c5 shard (rows [37.5M, 50M) of the 100M-row dataset, ``ll_scale`` from the global
    length): default precision, prepared planes = the exact kernel and layout bench.py times.  The same at D = 1024 and
    D = 768 (20 GB / 18 GB of X): the single pass with a cluster of two CTAs per tile (each CTA half of the columns, ~1000
    tiles per cluster, partial logits summed across the pair on every tile)."""
    from bijax_b200 import random as br

    S = 32
    X = _device_X(rows, D, row0, n_total)
    y = _logistic_y(X, row0, n_total)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.BernoulliLogits())
    eng = model._engine_for({})[0]
    assert eng.kernel_choice(S, rows) == "tcgen05_bf16_split"
    loc = 0.02 * tf.normal(tf.prng_key(5), (D,), 1)
    raw = (math.log(0.1) + 0.1 * tf.normal(tf.prng_key(6), (D,), 1)).astype(np.float32)
    full = float(n_total)
    got = eng.step(br.PRNGKey(2024), torch.tensor(loc, device="cuda"), torch.tensor(raw, device="cuda"), None, X, y,
                   full / n_total, S, prior_weight=1.0).clone()
    assert eng._planes is not None  # the TMA-fed variant on the prepared planes ran
    eng.release_planes()  # make room for the checker's float64 blocks
    eps = torch.tensor(tf.normal(tf.prng_key(2024), (S, D), tf.LAYOUT_LEGACY), dtype=torch.float64)
    want = ref.value_and_grad_chunked(loc, raw, {}, y, X, full, eps, chunk_rows=1 << 18, n_total=n_total)
    _check(got, want, D, D, [])
    del X, y
    torch.cuda.empty_cache()


@pytest.mark.timeout(900)
def test_c4_full_rank_d2048_s256_n1m_against_the_float64_oracle():
    """c4 at full size: full-rank Gaussian linear regression, D=2048, S=256, N=1M, learnable log_noise_scale.  Stated
    conditioning of the raw scale_tril: 0.1*I + 0.001*lower-triangular normal noise (q.log_prob here is the analytic
    identity A^-1 (z - loc) = eps; the oracle solves the triangular system in float64)."""
    from bijax_b200 import random as br

    D, S, N = 2048, 256, 1_000_000
    X = _device_X(N, D)
    w = br.normal(br.PRNGKey(1235), (D,), partitionable=True)
    y = X @ w + 0.1 * br.normal(br.PRNGKey(1237), (N,), partitionable=True)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), vi_type="full_rank")
    lns = np.float32(math.log(0.1))
    eng = model._engine_for({"log_noise_scale": torch.tensor([lns], device="cuda")})[0]
    assert eng.kernel_choice(S, N) == "tcgen05_gemm_two_pass"
    loc = 0.1 * tf.normal(tf.prng_key(6), (D,), 1)
    M = (0.1 * np.eye(D) + 0.001 * tf.normal(tf.prng_key(8), (D, D), 1)).astype(np.float32)
    got = eng.step(br.PRNGKey(17), torch.tensor(loc, device="cuda"), torch.tensor(M, device="cuda"),
                   torch.tensor([lns], device="cuda"), X, y, 1.0, S).clone()
    eng.release_planes()
    eng._ws = None
    torch.cuda.empty_cache()
    eps = torch.tensor(tf.normal(tf.prng_key(17), (S, D), tf.LAYOUT_LEGACY), dtype=torch.float64)
    want = ref.value_and_grad_chunked(loc, M, {"log_noise_scale": torch.tensor(float(lns), dtype=torch.float64)}, y, X,
                                      float(N), eps, chunk_rows=1 << 16)
    _check(got, want, D, D * D, ["log_noise_scale"])
    gM = got[1 + D : 1 + D + D * D].reshape(D, D)
    assert torch.all(torch.triu(gM, 1) == 0)

"""CPU: pin the oracle against every known-answer vector we have, and its gradient against finite differences."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import advi_ref as R
from oracle import threefry as tf

KAT = json.loads((Path(__file__).parent / "golden" / "threefry_kat.json").read_text())


def test_threefry_block_kats():
    for v in KAT["threefry2x32"]:
        k = [int(x, 16) for x in v["key"]]
        c = [int(x, 16) for x in v["ctr"]]
        a, b = tf.threefry2x32(k[0], k[1], c[0], c[1])
        assert [int(a), int(b)] == [int(x, 16) for x in v["out"]]


def test_split_legacy_doc_value():
    assert tf.split(tf.prng_key(0), 2, tf.LAYOUT_LEGACY).tolist() == KAT["split_legacy_seed0_num2"]


def test_normal_and_uniform_doc_values():
    for v in KAT["normal"]:
        key = tf.prng_key(v["seed"])
        for name, layout in (("legacy", tf.LAYOUT_LEGACY), ("partitionable", tf.LAYOUT_PARTITIONABLE)):
            got = tf.normal(key, tuple(v["shape"]), layout).ravel()[0]
            ulp = abs(int(np.float32(got).view(np.int32)) - int(np.float32(v[name]).view(np.int32)))
            assert ulp <= v["max_ulp"][name], (v, name, got)
    assert tf.uniform(tf.prng_key(0), (), tf.LAYOUT_LEGACY) == np.float32(KAT["uniform_legacy_seed0"])


def test_shape_invariance_and_offsets():
    key = tf.prng_key(7)
    for layout in (0, 1):
        a = tf.normal(key, (6, 5), layout)
        b = tf.normal(key, (30,), layout)
        assert np.array_equal(a.ravel(), b)
    full = tf.random_bits(key, 1000, tf.LAYOUT_PARTITIONABLE)
    part = tf.random_bits(key, 100, tf.LAYOUT_PARTITIONABLE, offset=333)
    assert np.array_equal(full[333:433], part)
    # odd sizes exercise the zero pad of the legacy layout
    assert tf.random_bits(key, 7, tf.LAYOUT_LEGACY).shape == (7,)


def test_normal_moments():
    x = tf.normal(tf.prng_key(3), (200000,), tf.LAYOUT_PARTITIONABLE).astype(np.float64)
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    assert np.isfinite(x).all()


def _coin_model():
    return R.ADVIRef({"p": R.Beta(0.5, 0.5)}, {"p": R.Sigmoid()}, R.bernoulli_probs_loglik("p"))


def test_golden_coin_toss_vector():
    """SURVEY.md appendix C (legacy layout): restatement-derived known answer for config c1."""
    init = tf.normal(tf.prng_key(0), (2,), tf.LAYOUT_LEGACY)
    eps = tf.normal(tf.prng_key(1), (16, 1), tf.LAYOUT_LEGACY)
    y = torch.tensor([1.0 if n % 10 < 7 else 0.0 for n in range(100)], dtype=torch.float64)
    loc = torch.tensor([float(init[0])], dtype=torch.float64)
    raw = torch.tensor([float(init[1])], dtype=torch.float64)
    loss, gl, gr, _ = _coin_model().value_and_grad(loc, raw, {}, y, None, 100, torch.tensor(eps, dtype=torch.float64))
    assert abs(loss.item() - 200.757306865) < 1e-6
    assert abs(gl.item() - (-47.500152711)) < 1e-6
    assert abs(gr.item() - 133.430885412) < 1e-6


@pytest.mark.parametrize("vi_type", ["mean_field", "full_rank"])
def test_gradient_matches_finite_differences(vi_type):
    torch.manual_seed(0)
    D, N, S = 4, 50, 6
    X = torch.randn(N, D, dtype=torch.float64) / 2
    y = (torch.rand(N, dtype=torch.float64) < 0.5).double()
    m = R.ADVIRef({"weights": R.MultivariateNormalDiag(torch.zeros(D), torch.ones(D))}, {}, R.bernoulli_logits_loglik(), vi_type)
    loc = torch.randn(D, dtype=torch.float64) * 0.3
    raw = torch.randn(D, dtype=torch.float64) * 0.2 if vi_type == "mean_field" else (0.5 * torch.eye(D) + 0.1 * torch.randn(D, D)).double()
    eps = torch.randn(S, D, dtype=torch.float64)
    loss, gl, gr, _ = m.value_and_grad(loc, raw, {}, y, {"X": X}, 3 * N, eps)
    h = 1e-6
    for i in range(D):
        e = torch.zeros(D, dtype=torch.float64)
        e[i] = h
        fd = (m.loss(loc + e, raw, {}, y, {"X": X}, 3 * N, eps) - m.loss(loc - e, raw, {}, y, {"X": X}, 3 * N, eps)) / (2 * h)
        assert abs(fd.item() - gl[i].item()) < 1e-5 * max(1.0, abs(fd.item()))
    flat = raw.reshape(-1)
    for i in range(flat.numel()):
        e = torch.zeros_like(flat)
        e[i] = h
        fp = m.loss(loc, (flat + e).reshape(raw.shape), {}, y, {"X": X}, 3 * N, eps)
        fm = m.loss(loc, (flat - e).reshape(raw.shape), {}, y, {"X": X}, 3 * N, eps)
        fd = (fp - fm) / (2 * h)
        assert abs(fd.item() - gr.reshape(-1)[i].item()) < 1e-5 * max(1.0, abs(fd.item()))
    if vi_type == "full_rank":  # MultivariateNormalTriL masks the strict upper triangle: zero gradient there
        assert torch.all(torch.triu(gr, diagonal=1) == 0)


def test_batched_restatements_agree_with_the_per_sample_one():
    torch.manual_seed(1)
    D, N, S = 5, 40, 7
    X = torch.randn(N, D, dtype=torch.float64)
    eps = torch.randn(S, D, dtype=torch.float64)
    loc, raw = torch.randn(D, dtype=torch.float64) * 0.1, torch.randn(D, dtype=torch.float64) * 0.1 - 1
    prior = {"weights": R.MultivariateNormalDiag(torch.zeros(D), torch.ones(D))}
    y = (torch.rand(N, dtype=torch.float64) < 0.4).double()
    a = R.ADVIRef(prior, {}, R.bernoulli_logits_loglik()).value_and_grad(loc, raw, {}, y, {"X": X}, N, eps)
    b = R.logistic_meanfield_value_and_grad(loc, raw, X, y, N, eps)
    for u, v in zip(a[:3], b):
        assert torch.allclose(u, v, rtol=1e-12, atol=1e-12)
    yl = torch.randn(N, dtype=torch.float64)
    lns = torch.tensor(-0.7, dtype=torch.float64)
    a = R.ADVIRef(prior, {}, R.gaussian_linear_loglik()).value_and_grad(loc, raw, {"log_noise_scale": lns}, yl, {"X": X}, 2 * N, eps)
    b = R.linreg_meanfield_value_and_grad(loc, raw, lns, X, yl, 2 * N, eps)
    assert torch.allclose(a[0], b[0], rtol=1e-12) and torch.allclose(a[1], b[1], rtol=1e-12) and torch.allclose(a[2], b[2], rtol=1e-12)
    assert torch.allclose(a[3]["log_noise_scale"], b[3], rtol=1e-12)


@pytest.mark.parametrize("case", ["logistic", "linreg_kwarg", "full_rank", "coin", "noise_latent"])
def test_chunked_checker_is_the_per_sample_restatement(case):
    """``value_and_grad_chunked`` (the checker of the full-size GPU parity tests) against ``value_and_grad`` on problems
    where both run: ragged row blocks, every likelihood exemplar, constrained latents, a row shard with ``n_total``."""
    from tests import cases as C
    from bijax_b200 import likelihoods as lk

    torch.manual_seed(3)
    D, N, S = 6, 101, 5
    X = torch.randn(N, D, dtype=torch.float64) / 2
    eps = torch.randn(S, D, dtype=torch.float64)
    loc = torch.randn(D, dtype=torch.float64) * 0.2
    raw = torch.randn(D, dtype=torch.float64) * 0.1 - 1
    prior = {"weights": R.MultivariateNormalDiag(torch.zeros(D), torch.ones(D))}
    kwargs, vi = {}, "mean_field"
    if case == "logistic":
        m = R.ADVIRef(prior, {}, R.bernoulli_logits_loglik())
        y = (torch.rand(N, dtype=torch.float64) < 0.4).double()
    elif case == "linreg_kwarg":
        m = R.ADVIRef(prior, {}, R.gaussian_linear_loglik())
        y, kwargs = torch.randn(N, dtype=torch.float64), {"log_noise_scale": torch.tensor(-0.3, dtype=torch.float64)}
    elif case == "full_rank":
        vi = "full_rank"
        m = R.ADVIRef(prior, {}, R.gaussian_linear_loglik(), vi)
        y, kwargs = torch.randn(N, dtype=torch.float64), {"log_noise_scale": torch.tensor(-0.3, dtype=torch.float64)}
        raw = (0.4 * torch.eye(D) + 0.05 * torch.randn(D, D)).double()
    elif case == "coin":
        m, X, D = _coin_model(), None, 1
        y = (torch.rand(N, dtype=torch.float64) < 0.7).double()
        loc, raw, eps = loc[:1], raw[:1], eps[:, :1]
    else:  # a constrained scalar latent next to the weights (sorted keys: noise < weights)
        pr = {"noise": R.Gamma(2.0, 3.0), "weights": R.MultivariateNormalDiag(torch.zeros(D - 1), torch.ones(D - 1))}
        m = R.ADVIRef(pr, {"noise": R.Softplus()}, C.oracle_likelihood(lk.GaussianLinear(noise_latent="noise")))
        X, y = X[:, : D - 1], torch.randn(N, dtype=torch.float64)
    inputs = {"X": X} if X is not None else None
    a = m.value_and_grad(loc, raw, kwargs, y, inputs, 3 * N, eps)
    b = m.value_and_grad_chunked(loc, raw, kwargs, y, X, 3 * N, eps, chunk_rows=17)
    for u, v in zip(a[:3], b[:3]):
        assert torch.allclose(u, v, rtol=1e-11, atol=1e-11)
    for k in kwargs:
        assert torch.allclose(a[3][k], b[3][k], rtol=1e-11, atol=1e-11)
    # two row shards, each told the global length, sum to the whole once the q / prior terms are counted once
    cut = 37
    s0 = m.value_and_grad_chunked(loc, raw, kwargs, y[:cut], X[:cut] if X is not None else None, 3 * N, eps, chunk_rows=16, n_total=N)
    s1 = m.value_and_grad_chunked(loc, raw, kwargs, y[cut:], X[cut:] if X is not None else None, 3 * N, eps, chunk_rows=16, n_total=N)
    z = m.value_and_grad_chunked(loc, raw, kwargs, y[:0], X[:0] if X is not None else None, 3 * N, eps, n_total=N)  # q / prior terms alone
    for u, v0, v1, vz in zip(a[:3], s0[:3], s1[:3], z[:3]):
        assert torch.allclose(u, v0 + v1 - vz, rtol=1e-10, atol=1e-10)


def test_coin_toss_loss_is_bounded_by_the_log_evidence():
    """coin_toss.ipynb: tosses [0,1,1,1,1,0], prior Beta(1,2): log evidence -4.941642, so -ELBO >= 4.9416 for any q."""
    y = torch.tensor([0, 1, 1, 1, 1, 0], dtype=torch.float64)
    m = R.ADVIRef({"p": R.Beta(1.0, 2.0)}, {"p": R.Sigmoid()}, R.bernoulli_probs_loglik("p"))
    eps = torch.tensor(tf.normal(tf.prng_key(5), (4000, 1), tf.LAYOUT_PARTITIONABLE), dtype=torch.float64)
    # the exact posterior Beta(5,4) pushed through logit is close to N(0.24, 0.68): the bound must be near-tight there
    loss = m.loss(torch.tensor([0.24], dtype=torch.float64), torch.log(torch.tensor([0.68], dtype=torch.float64)), {}, y, None, 6, eps)
    assert 4.9416 - 0.02 < loss.item() < 5.0


def test_oracle_posterior_log_prob_is_the_logit_normal_density():
    """core.py:57-71 restated: for one Sigmoid-constrained latent the fitted posterior is logit-normal,
    log q(p) = Normal(logit p; m, s).log_prob - log(p (1 - p))."""
    import math

    import torch

    from oracle import advi_ref as R

    ref = R.ADVIRef({"p": R.Beta(0.5, 0.5)}, {"p": R.Sigmoid()}, R.bernoulli_probs_loglik("p"))
    m, rho = 0.3, -0.4
    s = math.exp(rho)
    for p in (0.1, 0.5, 0.83):
        got = ref.posterior_log_prob(torch.tensor([m], dtype=torch.float64), torch.tensor([rho], dtype=torch.float64),
                                     {"p": torch.tensor(p, dtype=torch.float64)}).item()
        z = math.log(p / (1 - p))
        want = -0.5 * ((z - m) / s) ** 2 - math.log(s) - 0.5 * math.log(2 * math.pi) - math.log(p * (1 - p))
        assert abs(got - want) < 1e-12


def test_oracle_map_and_mcmc_joints_differ_by_the_log_det_jacobian():
    """map.py:27-33 evaluates the prior density at the constrained value; mcmc.py:30-34 the density of z (core.py:117):
    for one Gamma latent under Exp they differ by fldj(z) = z, and both equal the hand-written formula."""
    import math

    import torch

    from oracle import advi_ref as R

    ref = R.ADVIRef({"noise": R.Gamma(2.0, 3.0), "weights": R.Normal(torch.zeros(2, dtype=torch.float64), 1.0)},
                    {"noise": R.Exp()}, R.gaussian_linear_loglik("weights", "log_noise_scale"))
    # replace the likelihood by one that reads the latent noise, as tests/cases.py does
    def lik(latent, outputs, inputs, **kw):
        loc = inputs["X"] @ latent["weights"]
        zz = (outputs - loc) / latent["noise"]
        return (-0.5 * zz * zz - 0.5 * R.LOG_2PI - torch.log(latent["noise"])).sum()
    ref.log_likelihood_fn = lik
    X = torch.tensor([[1.0, 2.0], [0.5, -1.0], [2.0, 0.0]], dtype=torch.float64)
    y = torch.tensor([0.3, -0.2, 1.1], dtype=torch.float64)
    z = torch.tensor([0.4, 0.7, -0.3], dtype=torch.float64)  # layout: noise, weights (sorted keys)
    tau, w = math.exp(0.4), z[1:]
    log_gamma = 2.0 * math.log(3.0) - math.lgamma(2.0) + (2.0 - 1.0) * math.log(tau) - 3.0 * tau
    log_norm = float((-0.5 * w * w - 0.5 * R.LOG_2PI).sum())
    r = (y - X @ w) / tau
    ll = float((-0.5 * r * r - 0.5 * R.LOG_2PI - math.log(tau)).sum())
    m = ref.map_neg_log_joint(z, y, {"X": X}).item()
    j = ref.mcmc_log_joint(z, y, {"X": X}).item()
    assert abs(m + (log_gamma + log_norm + ll)) < 1e-12
    assert abs(j - (log_gamma + 0.4 + log_norm + ll)) < 1e-12


def test_oracle_distribution_and_bijector_formulas_against_scipy():
    """The TFP formulas the oracle restates pinned to an independent implementation: scipy.stats for
    Normal / Beta / Gamma / multivariate-normal log densities, numerical derivatives for the bijectors' forward
    log-det-Jacobians, scipy.special.erfinv for the uniform -> normal map of jax.random.normal."""
    import numpy as np
    import scipy.special as sp
    import scipy.stats as st
    import torch

    from oracle import advi_ref as R
    from oracle import threefry as tf

    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    x = np.array([0.05, 0.3, 0.77, 0.999])
    assert np.allclose(R.Beta(2.5, 0.7).log_prob(t(x)).numpy(), st.beta(2.5, 0.7).logpdf(x), rtol=1e-12)
    g = np.array([1e-3, 0.4, 2.0, 30.0])
    assert np.allclose(R.Gamma(2.0, 3.0).log_prob(t(g)).numpy(), st.gamma(2.0, scale=1 / 3.0).logpdf(g), rtol=1e-12)
    n = np.array([-3.0, 0.1, 4.0])
    assert np.allclose(R.Normal(t(0.5), t(1.7)).log_prob(t(n)).numpy(), st.norm(0.5, 1.7).logpdf(n), rtol=1e-12)
    # bijectors: fldj = log |d forward / dx| and inverse(forward(x)) = x
    xs = np.linspace(-3.0, 3.0, 13)
    for b in (R.Exp(), R.Softplus(), R.Sigmoid(), R.Chain([R.Softplus(), R.Exp()])):
        f = lambda v: b.forward(t(v)).numpy()
        num = np.log(np.abs((f(xs + 1e-6) - f(xs - 1e-6)) / 2e-6))
        assert np.allclose(b.fldj(t(xs)).numpy(), num, atol=1e-6), type(b).__name__
        assert np.allclose(b.inverse(b.forward(t(xs))).numpy(), xs, atol=1e-9), type(b).__name__
    # the fitted Gaussians: q.log_prob of MultivariateNormalTriL and MultivariateNormalDiagPlusLowRank = N(loc, A A^T)
    rng = np.random.default_rng(0)
    D, r = 6, 2
    loc = rng.normal(size=D)
    ref = R.ADVIRef({"w": R.Normal(torch.zeros(D, dtype=torch.float64), 1.0)}, {}, R.bernoulli_logits_loglik("w"), "full_rank")
    M = 0.5 * np.eye(D) + 0.1 * rng.normal(size=(D, D))
    z = rng.normal(size=D)
    got = ref.posterior_log_prob(t(loc), t(M), {"w": t(z)}).item()
    L = np.tril(M)
    assert abs(got - st.multivariate_normal(loc, L @ L.T).logpdf(z)) < 1e-10
    ref.vi_type = "low_rank"
    ref.scale_bijector = R.Exp()
    raw_d, U = rng.normal(size=D) * 0.3 - 0.5, 0.4 * rng.normal(size=(D, r))
    A = np.diag(np.exp(raw_d)) + U @ U.T
    got = ref.posterior_log_prob(t(loc), (t(raw_d), t(U)), {"w": t(z)}).item()
    assert abs(got - st.multivariate_normal(loc, A @ A.T).logpdf(z)) < 1e-10
    # uniform -> normal: sqrt(2) * erfinv(u) with XLA's float32 (Giles) polynomial: relative error below 1e-5 everywhere
    # (a few float32 ulps in the bulk, 5e-6 in the far tail) against scipy's double-precision erfinv
    u = tf.uniform(tf.prng_key(3), (4096,), tf.LAYOUT_LEGACY, minval=np.nextafter(np.float32(-1), np.float32(0)), maxval=np.float32(1))
    nrm = tf.normal(tf.prng_key(3), (4096,), tf.LAYOUT_LEGACY)
    want = np.sqrt(2.0) * sp.erfinv(u.astype(np.float64))
    rel = np.abs(nrm - want) / np.maximum(np.abs(want), 1e-3)
    assert rel.max() < 1e-5 and np.median(rel) < 2e-7


@pytest.mark.parametrize("layout", [0, 1])
def test_randint_restatement_properties(layout):
    """oracle/threefry.py::randint (the index draw of the minibatch sampler): for span <= 2**16 the uint32 recipe equals
    the exact (higher * 2**32 + lower) mod span; above that the wrapped multiplier is 0 and it reduces to lower mod span;
    a degenerate range returns minval; draws are in range and reproducible."""
    key = tf.prng_key(123)
    k1, k2 = tf.split(key, 2, layout)
    n = 500
    hi = [int(v) for v in tf.random_bits(k1, n, layout)]
    lo = [int(v) for v in tf.random_bits(k2, n, layout)]
    for span in (1, 2, 10, 100, 4096, 65535, 65536):
        got = tf.randint(key, n, 0, span, layout)
        m = (65536 % span) ** 2 % span
        assert [int(g) for g in got] == [(((h % span) * m) % 2**32 + l % span) % 2**32 % span for h, l in zip(hi, lo)]
        if span <= 256:  # no intermediate overflow: the recipe is the exact double-width remainder
            assert [int(g) for g in got] == [((h << 32) + l) % span for h, l in zip(hi, lo)]
    for span in (65537, 10_000_000, 2**31 - 1):
        assert [int(g) for g in tf.randint(key, n, 0, span, layout)] == [l % span for l in lo]
    assert np.array_equal(tf.randint(key, 7, 5, 5, layout), np.full(7, 5, np.int32))
    got = tf.randint(key, n, -3, 4, layout)
    assert got.dtype == np.int32 and got.min() >= -3 and got.max() <= 3 and len(set(got.tolist())) == 7
    assert np.array_equal(got, tf.randint(key, n, -3, 4, layout))



def test_randint_regression_vectors():
    """tests/golden/randint_vectors.json (scripts/make_randint_fixture.py): the index draws the oracle and the CUDA sampler
    agreed on when the fixture was committed -- a change of either shows up here (CPU) or in tests/test_minibatch_gpu.py."""
    fx = json.loads((Path(__file__).parent / "golden" / "randint_vectors.json").read_text())
    assert len(fx["cases"]) == 14
    for c in fx["cases"]:
        got = tf.randint(tf.prng_key(c["seed"]), c["n"], c["minval"], c["maxval"], c["layout"])
        assert [int(v) for v in got] == c["values"], c


def test_optax_adam_restatement():
    """oracle/optax_ref.py: closed forms of optax.adam (first update; constant gradient) and an independent implementation
    of the same published algorithm (torch.optim.Adam in float64: eps outside the root, both bias corrections)."""
    from oracle import optax_ref as O

    rng = np.random.default_rng(0)
    n, lr, eps = 9, 0.03, 1e-8
    g = rng.normal(size=n)
    st = O.adam_init(n)
    u1 = O.adam_update(g, st, lr)
    assert np.allclose(u1, -lr * g / (np.abs(g) + eps), rtol=1e-14, atol=0)
    for _ in range(5):  # constant gradient: mu^ = g and nu^ = g^2 at every step
        assert np.allclose(O.adam_update(g, st, lr), u1, rtol=1e-12, atol=0)
    assert st.count == 6
    # a quadratic bowl, 60 steps, against torch.optim.Adam
    A = rng.normal(size=(n, n))
    A = A @ A.T / n + np.eye(n)
    b = rng.normal(size=n)
    res = O.train(lambda p, key: (0.5 * p @ A @ p - b @ p, A @ p - b), np.zeros(n), range(60), lr, 0.8, 0.95, 1e-6)
    p = torch.zeros(n, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([p], lr=lr, betas=(0.8, 0.95), eps=1e-6)
    At, bt = torch.tensor(A), torch.tensor(b)
    losses = []
    for _ in range(60):
        opt.zero_grad()
        loss = 0.5 * p @ At @ p - bt @ p
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.allclose(res["params"], p.detach().numpy(), rtol=1e-10, atol=1e-12)
    assert np.allclose(res["losses"], losses, rtol=1e-10, atol=1e-12)


def test_log1p_specs_against_the_documented_normals():
    """The default float spec (XLA's own Log1p expansion with the Cephes logf it inlines on CPU) reproduces all four
    documented jax.random.normal values exactly; so does the platform-libm float spec; the round-1 spec (correctly rounded
    log1p) misses one of them by one ulp -- log(1 - u*u) sits 0.004 ulp from a rounding tie there."""
    want = {"xla_cpu": [0, 0, 0, 0], "libm_f32": [0, 0, 0, 0], "correctly_rounded": [0, 1, 0, 0]}
    old = tf.LOG1P_SPEC
    try:
        for spec, ulps in want.items():
            tf.LOG1P_SPEC = spec
            got = []
            for v in KAT["normal"]:
                for name, layout in (("legacy", tf.LAYOUT_LEGACY), ("partitionable", tf.LAYOUT_PARTITIONABLE)):
                    x = tf.normal(tf.prng_key(v["seed"]), tuple(v["shape"]), layout).ravel()[0]
                    got.append(abs(int(np.float32(x).view(np.int32)) - int(np.float32(v[name]).view(np.int32))))
            assert got == ulps, (spec, got)
    finally:
        tf.LOG1P_SPEC = old
    assert tf.LOG1P_SPEC == "xla_cpu"


def test_float_spec_building_blocks():
    """logf_cephes within one ulp of the true logarithm; log1p_f32 within two ulps of log1p on (-1, 0]; the three specs
    agree to a few ulps on a million draws (they differ in ~1 % of them)."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(1e-7, 1.0, 200000), np.exp(rng.uniform(-80, 80, 200000))]).astype(np.float32)
    ref = np.log(x.astype(np.float64))
    err = np.abs(tf.logf_cephes(x).astype(np.float64) - ref) / np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    assert err.max() < 1.0
    assert tf.logf_cephes(np.zeros(1, np.float32))[0] == -np.inf
    t = -rng.uniform(0.0, 1.0, 400000).astype(np.float32) ** 2
    ref = np.log1p(t.astype(np.float64))
    got = tf.log1p_f32(t, "xla_cpu").astype(np.float64)
    assert (np.abs(got - ref) <= 2.0 * np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64) + 1e-45).all()
    bits = tf.random_bits(tf.prng_key(9), 1 << 20, tf.LAYOUT_PARTITIONABLE)
    u = tf.uniform_from_bits(bits, tf._LO, 1.0)
    a, b, c = (tf.erfinv_f32(u, s) for s in ("xla_cpu", "correctly_rounded", "libm_f32"))
    for other in (b, c):
        d = np.abs(a.view(np.int32).astype(np.int64) - other.view(np.int32).astype(np.int64))
        assert d.max() <= 4 and 0.0 < (d > 0).mean() < 0.05


def test_values_published_in_the_jax_documentation():
    """tests/golden/threefry_kat.json "jax_docs": split chains, normals of subkeys, a (3,) draw and a uniform printed in JAX's
    public docs, for both counter layouts -- all exact."""
    for name, layout in (("legacy", tf.LAYOUT_LEGACY), ("partitionable", tf.LAYOUT_PARTITIONABLE)):
        for v in KAT["jax_docs"][name]:
            key = tf.prng_key(v["seed"]) if "seed" in v else np.array(v["key"], dtype=np.uint32)
            if "split" in v:
                assert tf.split(key, len(v["split"]), layout).tolist() == v["split"], v["what"]
            if "normal" in v:
                got = tf.normal(key, tuple(v["shape"]), layout).ravel()
                assert np.array_equal(got, np.array(v["normal"], dtype=np.float32)), (v["what"], got)
            if "uniform" in v:
                got = tf.uniform(key, tuple(v["shape"]), layout).ravel()
                assert np.array_equal(got, np.array(v["uniform"], dtype=np.float32)), (v["what"], got)

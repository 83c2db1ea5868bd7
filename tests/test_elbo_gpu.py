"""GPU: loss and every gradient of the fused CUDA step against the float64 oracle on the same key and inputs.
Tolerance: rtol 1e-5 (north_star), with an absolute floor of 1e-5 x the largest entry of each gradient block."""
import math

import numpy as np
import pytest
import torch

from bijax_b200 import bijectors as tfb
from bijax_b200 import distributions as tfd
from bijax_b200 import likelihoods as lk
from oracle import threefry as tf
from tests import cases as C

pytestmark = pytest.mark.gpu


def _run_case(model, ref, loc, raw, kwargs, y, X, full, seed, S, layout, expect_kernel=None):
    from bijax_b200 import random as br

    dev = torch.device("cuda")
    key = br.PRNGKey(seed)
    params = {"posterior": {"loc": torch.tensor(loc, device=dev), ("scale_diag" if model.vi_type == "mean_field" else "scale_tril"): torch.tensor(raw, device=dev)}}
    for k, v in kwargs.items():
        params[k] = torch.tensor(v, dtype=torch.float32, device=dev)
    br.config.threefry_partitionable = bool(layout)
    try:
        yt = torch.tensor(y, device=dev)
        inputs = {"X": torch.tensor(X, device=dev)} if X is not None else None
        loss, grads = model.value_and_grad(params, yt, inputs, full, key, S)
        # the autograd route must give the same numbers
        p2 = {"posterior": {k: v.clone().requires_grad_(True) for k, v in params["posterior"].items()}}
        p2.update({k: torch.tensor(v, dtype=torch.float32, device=dev, requires_grad=True) for k, v in kwargs.items()})
        l2 = model.loss_fn(p2, yt, inputs, full, key, S)
        l2.backward()
    finally:
        br.config.threefry_partitionable = False
    (wl, wgl, wgr, wgk), eps = C.oracle_step(ref, loc, raw, kwargs, y, X, full, tf.prng_key(seed), S, layout)
    if expect_kernel is not None:
        eng = next(iter(model._engines.values()))
        assert eng.kernel_choice(S, len(y)) == expect_kernel
    skey = "scale_diag" if model.vi_type == "mean_field" else "scale_tril"
    C.assert_close(loss.item(), wl.item(), what="loss")
    C.assert_close(grads["posterior"]["loc"].cpu().numpy(), wgl.numpy(), what="d loc")
    C.assert_close(grads["posterior"][skey].cpu().numpy(), wgr.numpy(), what="d raw scale")
    for k in kwargs:
        C.assert_close(grads[k].cpu().numpy(), wgk[k].numpy(), what=f"d {k}")
        assert torch.equal(p2[k].grad, grads[k].reshape(p2[k].shape))
    assert l2.item() == loss.item()
    assert torch.equal(p2["posterior"]["loc"].grad, grads["posterior"]["loc"])
    assert torch.equal(p2["posterior"][skey].grad, grads["posterior"][skey])
    return loss, grads


@pytest.mark.parametrize("layout", [0, 1])
def test_c1_coin_toss(layout):
    """config c1: Beta(0.5,0.5) prior, Sigmoid bijector, 100 Bernoulli observations, S=16 (README.md:41,91-95)."""
    model, ref = C.build_pair({"p_of_heads": tfd.Beta(0.5, 0.5)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
    init = tf.normal(tf.prng_key(0), (2,), layout)
    y = np.array([1.0 if n % 10 < 7 else 0.0 for n in range(100)], dtype=np.float32)
    loss, _ = _run_case(model, ref, init[:1], init[1:], {}, y, None, 100, 1, 16, layout, expect_kernel="coin")
    if layout == 0:  # SURVEY.md appendix C known answer
        assert abs(loss.item() - 200.757306865) < 200.76 * 1e-5


@pytest.mark.parametrize("N,S", [(100, 256), (100, 257), (4096, 16), (4097, 16), (100_000, 33)])
def test_coin_toss_on_either_side_of_the_one_block_limits(N, S):
    """The one-block step (k_tiny_coin_step: S <= 256, N <= 4096) and the four-launch path give the oracle's numbers.
    The gradient of the single latent is sum_s (k - N*sigmoid(z_s)) * eps_s: thousands of tosses nearly cancel against N*theta,
    which is why the kernels sum the labels once in double instead of adding N rounded products per sample."""
    model, ref = C.build_pair({"p_of_heads": tfd.Beta(2.0, 3.0)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
    init = tf.normal(tf.prng_key(5), (2,), 0)
    y = np.array([1.0 if (n * 7) % 10 < 4 else 0.0 for n in range(N)], dtype=np.float32)
    _run_case(model, ref, init[:1], 0.3 * init[1:] - 1.0, {}, y, None, max(N, 1), 11, S, 0, expect_kernel="coin")


@pytest.mark.parametrize("S", [9, 300])
def test_coin_toss_with_other_latents_beside_the_probability(S):
    """D = 4: the toss probability is coordinate 3 of the flattened latent; the other coordinates only see their priors."""
    prior = {"a_scale": tfd.Gamma(2.0, 1.5), "b_vec": tfd.MultivariateNormalDiag(loc=np.zeros(2, np.float32), scale_diag=np.full(2, 0.7, np.float32)),
             "p_of_heads": tfd.Beta(1.5, 0.8)}
    model, ref = C.build_pair(prior, {"a_scale": tfb.Softplus(), "p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"),
                              scale_bijector=tfb.Softplus())
    init = tf.normal(tf.prng_key(6), (8,), 0)
    y = np.array([1.0 if n % 3 == 0 else 0.0 for n in range(250)], dtype=np.float32)
    _run_case(model, ref, 0.5 * init[:4], 0.3 * init[4:], {}, y, None, 500, 12, S, 0, expect_kernel="coin")


@pytest.mark.parametrize("N,S", [(4096, 64), (1000, 7), (1, 3), (300, 300)])
def test_c2_linear_regression_learnable_noise_small_d(N, S):
    """config c2 shape family: D=5, Gaussian likelihood with trainable log_noise_scale (README.md:99-106)."""
    D = 5
    X, y = C.synth_linear(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.GaussianLinear())
    loc = 0.1 * tf.normal(tf.prng_key(3), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32) + 0.05 * tf.normal(tf.prng_key(4), (D,), 1)
    _run_case(model, ref, loc, raw, {"log_noise_scale": np.float32(math.log(0.3))}, y, X, 10 * N, 7, S, 0, expect_kernel="fp32_smalld")


@pytest.mark.parametrize("D", [1, 3, 4, 8, 11, 15])
def test_small_d_every_record_width(D):
    N, S = 777, 16
    X, y = C.synth_logistic(N, D)
    prior = {"weights": tfd.Normal(np.zeros(D, np.float32), 2.0)}
    model, ref = C.build_pair(prior, {}, lk.BernoulliLogits())
    loc = 0.3 * tf.normal(tf.prng_key(13), (D,), 1)
    raw = -1.0 + 0.1 * tf.normal(tf.prng_key(14), (D,), 1)
    _run_case(model, ref, loc, raw, {}, y, X, N, 21, S, 1, expect_kernel="fp32_smalld")


@pytest.mark.parametrize("N,D,S", [(3000, 48, 8), (129, 16, 33), (5000, 200, 32), (640, 512, 16)])
def test_logistic_regression_fp32_tiled(N, D, S):
    """config c3 shape family at sizes the oracle finishes in seconds, forced onto the CUDA-core fallback."""
    X, y = C.synth_logistic(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), precision="fp32")
    loc = 0.05 * tf.normal(tf.prng_key(5), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32)
    _run_case(model, ref, loc, raw, {}, y, X, 4 * N, 11, S, 0, expect_kernel="fp32_tiled")


@pytest.mark.parametrize("D,S,lik", [(24, 10, "gauss"), (70, 5, "logit")])
def test_full_rank(D, S, lik):
    """config c4 shape family: dense D x D raw scale_tril (upper triangle ignored, zero gradient), well-conditioned M."""
    N = 500
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    if lik == "gauss":
        X, y = C.synth_linear(N, D)
        model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), vi_type="full_rank", precision="fp32")
        kwargs = {"log_noise_scale": np.float32(math.log(0.2))}
    else:
        X, y = C.synth_logistic(N, D)
        model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), vi_type="full_rank", precision="fp32")
        kwargs = {}
    loc = 0.1 * tf.normal(tf.prng_key(6), (D,), 1)
    M = (0.1 * np.eye(D) + 0.01 * tf.normal(tf.prng_key(8), (D, D), 1)).astype(np.float32)  # conditioning stated: 0.1 I + 0.01 noise
    loss, grads = _run_case(model, ref, loc, M, kwargs, y, X, N, 17, S, 0)
    g = grads["posterior"]["scale_tril"].cpu().numpy()
    assert np.all(np.triu(g, 1) == 0.0)


def test_constrained_latents_priors_and_chains():
    """Gamma + Exp, Beta + Sigmoid, Normal + Softplus chain, plus a latent noise scale feeding the Gaussian likelihood."""
    N, D, S = 600, 6, 12
    X, y = C.synth_linear(N, D, noise=0.5)
    prior = {
        "weights": tfd.Normal(np.zeros(D, np.float32), 1.5),
        "noise": tfd.Gamma(2.0, 3.0),
        "aux_beta": tfd.Beta(np.array([2.0, 0.5], np.float32), np.array([3.0, 1.0], np.float32)),
        "aux_pos": tfd.Gamma(1.0, 0.5),
        "aux_chain": tfd.Normal(1.0, 0.7),
    }
    bij = {
        "noise": tfb.Exp(),
        "aux_beta": tfb.Sigmoid(),
        "aux_pos": tfb.Softplus(),
        "aux_chain": tfb.Chain([tfb.Softplus(), tfb.Exp()]),
    }
    model, ref = C.build_pair(prior, bij, lk.GaussianLinear(noise_latent="noise"))
    Dtot = model.size
    loc = 0.3 * tf.normal(tf.prng_key(31), (Dtot,), 1)
    raw = -1.5 + 0.2 * tf.normal(tf.prng_key(32), (Dtot,), 1)
    _run_case(model, ref, loc, raw, {}, y, X, N, 33, S, 1)


def test_softplus_scale_bijector_and_constant_noise():
    N, D, S = 400, 5, 9
    X, y = C.synth_linear(N, D)
    prior = {"weights": tfd.Normal(np.zeros(D, np.float32), 1.0)}
    model, ref = C.build_pair(prior, {}, lk.GaussianLinear(noise_scale=0.25), scale_bijector=tfb.Softplus())
    loc = 0.2 * tf.normal(tf.prng_key(41), (D,), 1)
    raw = -2.0 + 0.3 * tf.normal(tf.prng_key(42), (D,), 1)
    _run_case(model, ref, loc, raw, {}, y, X, 3 * N, 43, S, 0)


def test_eps_is_bit_exact_and_step_is_deterministic():
    from bijax_b200 import random as br

    N, D, S = 2000, 64, 8
    X, y = C.synth_logistic(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.BernoulliLogits(), precision="fp32")
    dev = torch.device("cuda")
    params = {"posterior": {"loc": torch.zeros(D, device=dev), "scale_diag": torch.full((D,), -2.0, device=dev)}}
    eng, _ = model._engine_for({})
    key = br.PRNGKey(77)
    for layout in (0, 1):
        eps = torch.empty(S * D, device=dev)
        out1 = eng.step(key, params["posterior"]["loc"], params["posterior"]["scale_diag"], None, torch.tensor(X, device=dev),
                        torch.tensor(y, device=dev), 1.0, S, eps_out=eps, partitionable=bool(layout)).clone()
        out2 = eng.step(key, params["posterior"]["loc"], params["posterior"]["scale_diag"], None, torch.tensor(X, device=dev),
                        torch.tensor(y, device=dev), 1.0, S, partitionable=bool(layout))
        want = tf.normal(tf.prng_key(77), (S, D), layout).ravel()
        assert np.array_equal(eps.cpu().numpy().view(np.uint32), want.view(np.uint32))
        assert torch.equal(out1, out2), "two runs of the same step must agree bit for bit"


def test_host_buffer_entry_point_matches_device_entry_point():
    from bijax_b200 import random as br

    N, D, S = 1500, 5, 16
    X, y = C.synth_linear(N, D)
    prior = {"weights": tfd.Normal(np.zeros(D, np.float32), 1.0)}
    model, _ = C.build_pair(prior, {}, lk.GaussianLinear())
    dev = torch.device("cuda")
    loc = torch.linspace(-0.2, 0.2, D)
    raw = torch.full((D,), -2.0)
    lns = torch.tensor([-1.0])
    eng, _ = model._engine_for({"log_noise_scale": lns})
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    out_dev = eng.step(br.PRNGKey(9), loc.to(dev), raw.to(dev), lns.to(dev), Xd, yd, 2.0, S)
    ph = torch.cat([loc, raw, lns]).pin_memory()
    oh = torch.empty(1 + 2 * D + 1).pin_memory()
    eng.step_host(tf.prng_key(9), ph, oh, Xd, yd, 2.0, S)
    assert torch.equal(out_dev.cpu(), oh)


def test_error_paths():
    from bijax_b200 import _native as nv
    from bijax_b200 import random as br

    D = 4
    prior = {"weights": tfd.Normal(np.zeros(D, np.float32), 1.0)}
    model, _ = C.build_pair(prior, {}, lk.BernoulliLogits())
    dev = torch.device("cuda")
    params = {"posterior": {"loc": torch.zeros(D, device=dev), "scale_diag": torch.zeros(D, device=dev)}}
    with pytest.raises(ValueError):
        model.value_and_grad(params, torch.zeros(10, device=dev), {"X": torch.zeros(10, D + 1, device=dev)}, 10, br.PRNGKey(0), 4)
    with pytest.raises(TypeError):
        model.value_and_grad(params, torch.zeros(10, device=dev), {"X": torch.zeros(10, D, device=dev)}, 10, torch.zeros(2), 4)


@pytest.mark.parametrize(
    "N,D,S,lik",
    [(64, 128, 32, "logit"), (1, 128, 32, "logit"), (65, 256, 16, "logit"), (1000, 512, 32, "logit"),
     (20000, 512, 32, "logit"), (5000, 384, 7, "gauss"), (12345, 256, 32, "gauss"),
     (40000, 128, 32, "logit"), (30000, 384, 20, "gauss"), (50000, 256, 32, "logit"),  # several tiles per CTA for every block count
     # widths that are not a multiple of 128: planes padded to 64 columns, the missing blocks are TMA zero fill
     (20000, 500, 32, "logit"), (3000, 16, 32, "gauss"), (9000, 50, 5, "logit"), (4097, 64, 32, "logit"), (12000, 100, 17, "gauss"),
     (30000, 200, 32, "logit"), (25000, 320, 32, "gauss"), (15000, 448, 9, "logit"), (777, 511, 32, "gauss"), (20000, 129, 32, "logit"),
     # more than 32 samples on a narrow X: the single pass once per slice of 32 samples (each slice its own block of partials)
     (5000, 128, 64, "logit"), (3000, 256, 50, "gauss"), (20000, 384, 64, "logit"), (9000, 128, 96, "gauss"), (4000, 200, 40, "logit"),
     # 512 < D <= 1024: a cluster of two CTAs per tile, half of the columns each, partial logits exchanged through DSMEM
     (20000, 1024, 32, "logit"), (9000, 768, 32, "gauss"), (5000, 640, 20, "logit"), (12000, 1000, 32, "gauss"), (1, 1024, 32, "logit"),
     (65, 600, 8, "gauss"), (30000, 1024, 16, "logit"), (40000, 513, 32, "gauss"), (26000, 896, 32, "logit")],
)
@pytest.mark.parametrize("prepared", [True, False])
def test_tensor_core_path(N, D, S, lik, prepared):
    """config c3/c5 shape family on the fused tcgen05 kernel (bf16 split operands, fp32 accumulation in TMEM): the TMA-fed
    variant over the dataset's prepared split-bf16 planes, and the register-converting variant over raw fp32 X."""
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    if lik == "logit":
        X, y = C.synth_logistic(N, D)
        model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), precision="bf16x2", prepare_dataset=prepared)
        kwargs = {}
    else:
        X, y = C.synth_linear(N, D)
        model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), precision="bf16x2", prepare_dataset=prepared)
        kwargs = {"log_noise_scale": np.float32(math.log(0.5))}
    loc = 0.05 * tf.normal(tf.prng_key(5), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32)
    _run_case(model, ref, loc, raw, kwargs, y, X, 4 * N, 11, S, 0, expect_kernel="tcgen05_bf16_split")


@pytest.mark.parametrize(
    "N,D,S,lik",
    [(300, 64, 40, "gauss"), (1000, 200, 64, "logit"), (4097, 512, 130, "gauss"), (2500, 1000, 256, "logit"),
     (129, 2048, 256, "gauss"), (70000, 128, 33, "logit"), (1, 64, 1, "gauss"), (45000, 192, 100, "gauss"),
     (700, 64, 300, "logit"), (900, 128, 520, "gauss")],  # more samples than one 256-wide tile
)
def test_two_pass_tensor_core_gemm_path(N, D, S, lik):
    """config c4 shape family (S up to 256, D up to 2048): TMA-fed tcgen05 GEMMs with split-bf16 operand planes,
    pass 1 = logits + link function, pass 2 = the transposed contraction of the VJP."""
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    if lik == "logit":
        X, y = C.synth_logistic(N, D)
        model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), kernel="tcgen05_gemm_two_pass")
        kwargs = {}
    else:
        X, y = C.synth_linear(N, D)
        model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), kernel="tcgen05_gemm_two_pass")
        kwargs = {"log_noise_scale": np.float32(math.log(0.5))}
    loc = 0.05 * tf.normal(tf.prng_key(5), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32)
    _run_case(model, ref, loc, raw, kwargs, y, X, 4 * N, 11, S, 0, expect_kernel="tcgen05_gemm_two_pass")


@pytest.mark.parametrize("N,D,S,lik", [(3000, 256, 48, "logit"), (1500, 200, 130, "gauss")])
def test_three_term_split_is_fp32_accurate(N, D, S, lik):
    """precision="bf16x3": every operand split in three bf16 terms, six products per useful product -- the fp32-accurate
    tensor-core mode (north_star: "3xTF32 / fp32-accurate").  Held to a 4x tighter tolerance than the 2-term default."""
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    if lik == "logit":
        X, y = C.synth_logistic(N, D)
        model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), precision="bf16x3")
        kwargs = {}
    else:
        X, y = C.synth_linear(N, D)
        model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), precision="bf16x3")
        kwargs = {"log_noise_scale": np.float32(math.log(0.5))}
    loc = 0.05 * tf.normal(tf.prng_key(5), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32)
    loss, grads = _run_case(model, ref, loc, raw, kwargs, y, X, 4 * N, 11, S, 0, expect_kernel="tcgen05_gemm_two_pass")
    (wl, wgl, wgr, _), _ = C.oracle_step(ref, loc, raw, kwargs, y, X, 4 * N, tf.prng_key(11), S, 0)
    C.assert_close(loss.item(), wl.item(), rtol=2.5e-6, atol_scale=2.5e-6, what="loss (3-term)")
    C.assert_close(grads["posterior"]["loc"].cpu().numpy(), wgl.numpy(), rtol=2.5e-6, atol_scale=2.5e-6, what="d loc (3-term)")
    C.assert_close(grads["posterior"]["scale_diag"].cpu().numpy(), wgr.numpy(), rtol=2.5e-6, atol_scale=2.5e-6, what="d rho (3-term)")


@pytest.mark.parametrize("D,S,lik", [(128, 10, "gauss"), (200, 130, "logit"), (384, 64, "gauss")])
def test_full_rank_contractions_on_the_tensor_cores(D, S, lik):
    """z = loc + tril(M) eps and dM = tril(sum_s gz eps^T) as TMA-fed tcgen05 GEMMs with a 3-term bf16 split (fp32-accurate);
    taken for D >= 128 unless precision="fp32"."""
    N = 600
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    if lik == "gauss":
        X, y = C.synth_linear(N, D)
        model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), vi_type="full_rank")
        kwargs = {"log_noise_scale": np.float32(math.log(0.2))}
    else:
        X, y = C.synth_logistic(N, D)
        model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), vi_type="full_rank")
        kwargs = {}
    loc = 0.1 * tf.normal(tf.prng_key(6), (D,), 1)
    M = (0.1 * np.eye(D) + 0.01 * tf.normal(tf.prng_key(8), (D, D), 1)).astype(np.float32)  # conditioning stated: 0.1 I + 0.01 noise
    loss, grads = _run_case(model, ref, loc, M, kwargs, y, X, N, 17, S, 0)
    g = grads["posterior"]["scale_tril"].cpu().numpy()
    assert np.all(np.triu(g, 1) == 0.0)


@pytest.mark.parametrize("kernel,Dw,S,prepared", [("tcgen05_bf16_split", 256, 24, True), ("tcgen05_bf16_split", 256, 24, False),
                                                  ("tcgen05_gemm_two_pass", 192, 70, True), ("tcgen05_gemm_two_pass", 192, 70, False),
                                                  # a pair of CTAs per tile (the second CTA's theta columns start at w_offset + 384),
                                                  # 128-row tiles (two blocks / one block), each on prepared planes and on the in-step split
                                                  ("tcgen05_bf16_split", 700, 32, True), ("tcgen05_bf16_split", 700, 12, False),
                                                  ("tcgen05_bf16_split", 100, 32, True), ("tcgen05_bf16_split", 100, 9, False),
                                                  ("tcgen05_bf16_split", 40, 32, True), ("tcgen05_bf16_split", 40, 17, False)])
def test_tensor_core_paths_with_offset_weights_latent_noise_and_strided_X(kernel, Dw, S, prepared):
    """The weights are not the first latent (w_offset > 0: "a_first" sorts before "weights"), the Gaussian noise scale is a
    constrained latent (Gamma prior + Exp bijector), and X is a column slice of a wider matrix (row pitch ldx > Dw)."""
    from bijax_b200 import random as br

    N = 3001
    Xw, y = C.synth_linear(N, Dw + 8, noise=0.4)
    prior = {
        "a_first": tfd.Normal(np.zeros(3, np.float32), 2.0),
        "noise": tfd.Gamma(2.0, 3.0),
        "weights": tfd.MultivariateNormalDiag(loc=np.zeros(Dw, np.float32), scale_diag=np.ones(Dw, np.float32)),
    }
    model, ref = C.build_pair(prior, {"noise": tfb.Exp()}, lk.GaussianLinear(noise_latent="noise"), kernel=kernel,
                              prepare_dataset=prepared)
    D = model.size
    assert D == Dw + 4
    loc = (0.05 * tf.normal(tf.prng_key(5), (D,), 1)).astype(np.float32)
    raw = np.full(D, math.log(0.1), np.float32)
    dev = torch.device("cuda")
    Xfull = torch.tensor(Xw, device=dev)
    Xv = Xfull[:, :Dw]  # a view: stride(0) = Dw + 8
    assert Xv.stride(0) == Dw + 8
    eng = model._engine_for({})[0]
    # the engine's host layer makes inputs contiguous, so drive the C ABI stride through the engine's own data path
    out = eng.step(br.PRNGKey(11), torch.tensor(loc, device=dev), torch.tensor(raw, device=dev), None, Xv, torch.tensor(y, device=dev), 2.5, S)
    (wl, wgl, wgr, _), _ = C.oracle_step(ref, loc, raw, {}, y, Xw[:, :Dw], 2.5 * N, tf.prng_key(11), S, 0)
    assert eng.kernel_choice(S, N) == kernel
    C.assert_close(out[0].item(), wl.item(), what="loss")
    C.assert_close(out[1 : 1 + D].cpu().numpy(), wgl.numpy(), what="d loc")
    C.assert_close(out[1 + D :].cpu().numpy(), wgr.numpy(), what="d rho")


def test_prepared_dataset_planes_match_the_in_step_split():
    """The two-pass path with X pre-split once per dataset (bjx_prepare_x_planes, cached by the engine while the same
    unmodified tensor is passed) is bit-identical to splitting inside every step; an in-place update of X invalidates the cache."""
    from bijax_b200 import random as br
    from bijax_b200.engine import ElboEngine

    N, D, S = 3000, 192, 48
    X, y = _device_logistic(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.BernoulliLogits(), kernel="tcgen05_gemm_two_pass")
    warm = model._engine_for({})[0]
    cold = ElboEngine(model.latents, model.likelihood_fn, kernel="tcgen05_gemm_two_pass", prepare_dataset=False)
    loc, raw = torch.zeros(D, device=X.device), torch.full((D,), -2.0, device=X.device)
    a = warm.step(br.PRNGKey(4), loc, raw, None, X, y, 1.0, S).clone()
    b = cold.step(br.PRNGKey(4), loc, raw, None, X, y, 1.0, S).clone()
    assert warm._planes is not None and cold._planes is None
    assert torch.equal(a, b)
    X.mul_(0.5)
    a2 = warm.step(br.PRNGKey(4), loc, raw, None, X, y, 1.0, S).clone()
    b2 = cold.step(br.PRNGKey(4), loc, raw, None, X, y, 1.0, S).clone()
    assert torch.equal(a2, b2) and not torch.equal(a, a2)


@pytest.mark.parametrize("vi,layout", [("mean_field", 0), ("mean_field", 1), ("full_rank", 0)])
def test_sample_shards_sum_to_the_unsharded_step(vi, layout):
    """SURVEY.md section 8e, small-N case: the step evaluated as sample shards [0,5) + [5,12) (global eps indices, means over
    the global S, the entropy-gradient constant on one shard) sums to the unsharded 12-sample step, and the shards draw
    exactly the unsharded noise."""
    from bijax_b200 import random as br

    D, N, S = (40, 300, 12) if vi == "mean_field" else (130, 300, 12)
    X, y = C.synth_linear(N, D)
    dev = torch.device("cuda")
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.GaussianLinear(), vi_type=vi)
    lns = torch.tensor([-0.7], device=dev)
    eng = model._engine_for({"log_noise_scale": lns})[0]
    loc = torch.tensor(0.1 * tf.normal(tf.prng_key(6), (D,), 1), device=dev)
    if vi == "mean_field":
        raw = torch.full((D,), -1.2, device=dev)
    else:
        raw = torch.tensor((0.2 * np.eye(D) + 0.01 * tf.normal(tf.prng_key(8), (D, D), 1)).astype(np.float32), device=dev)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    key = br.PRNGKey(21)
    part = bool(layout)
    e_all = torch.empty(S * D, device=dev)
    full = eng.step(key, loc, raw, lns, Xd, yd, 2.0, S, eps_out=e_all, partitionable=part).clone()
    e_a, e_b = torch.empty(5 * D, device=dev), torch.empty(7 * D, device=dev)
    a = eng.step(key, loc, raw, lns, Xd, yd, 2.0, 5, eps_out=e_a, partitionable=part, sample_shard=(0, S, 1.0)).clone()
    b = eng.step(key, loc, raw, lns, Xd, yd, 2.0, 7, eps_out=e_b, partitionable=part, sample_shard=(5, S, 0.0)).clone()
    assert torch.equal(torch.cat([e_a, e_b]), e_all)
    C.assert_close((a + b).cpu().numpy(), full.cpu().numpy(), rtol=3e-6, atol_scale=3e-6, what="sample-shard additivity")


@pytest.mark.parametrize("D,r,S,lik,kernel", [(7, 1, 9, "gauss", None), (40, 3, 12, "logit", None), (192, 5, 33, "gauss", "tcgen05_gemm_two_pass"),
                                                (128, 64, 8, "logit", "tcgen05_bf16_split")])
def test_low_rank_family(D, r, S, lik, kernel):
    """vi_type="low_rank" (core.py:190-201): q = MultivariateNormalDiagPlusLowRank(loc, Exp(raw_d), U), scale operator
    diag(d) + U U^T; loss and every gradient (loc, raw scale_diag, scale_perturb_factor, kwargs) against the float64 oracle
    that uses a dense solve and slogdet; init draws 2D + D*r values in (loc, scale_diag, scale_perturb_factor) order."""
    import bijax_b200 as bj
    from bijax_b200 import random as br

    N = 900
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    if lik == "gauss":
        X, y = C.synth_linear(N, D)
        L = lk.GaussianLinear()
        kwargs = {"log_noise_scale": np.float32(math.log(0.3))}
    else:
        X, y = C.synth_logistic(N, D)
        L = lk.BernoulliLogits()
        kwargs = {}
    model = bj.ADVI(prior, {}, L, vi_type="low_rank", rank=r, kernel=kernel)
    _, ref = C.build_pair(prior, {}, L)
    p0 = model.init(br.PRNGKey(3))
    flat = tf.normal(tf.prng_key(3), (2 * D + D * r,), tf.LAYOUT_LEGACY)
    assert np.array_equal(p0["posterior"]["loc"].cpu().numpy(), flat[:D])
    assert np.array_equal(p0["posterior"]["scale_perturb_factor"].cpu().numpy(), flat[2 * D :].reshape(D, r))
    loc = (0.1 * tf.normal(tf.prng_key(6), (D,), 1)).astype(np.float32)
    raw_d = (-1.5 + 0.2 * tf.normal(tf.prng_key(7), (D,), 1)).astype(np.float32)
    U = (0.08 * tf.normal(tf.prng_key(8), (D, r), 1)).astype(np.float32)
    dev = torch.device("cuda")
    params = {"posterior": {"loc": torch.tensor(loc, device=dev), "scale_diag": torch.tensor(raw_d, device=dev),
                            "scale_perturb_factor": torch.tensor(U, device=dev)}}
    for k, v in kwargs.items():
        params[k] = torch.tensor(v, device=dev)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    loss, grads = model.value_and_grad(params, yd, {"X": Xd}, 3 * N, br.PRNGKey(17), S)
    t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    eps = torch.tensor(tf.normal(tf.prng_key(17), (S, D), 0), dtype=torch.float64)
    wl, wgl, wgd, wgU, wgk = ref.value_and_grad_low_rank(t64(loc), t64(raw_d), t64(U), {k: t64(v) for k, v in kwargs.items()},
                                                         t64(y), {"X": t64(X)}, 3 * N, eps)
    C.assert_close(loss.item(), wl.item(), what="loss")
    C.assert_close(grads["posterior"]["loc"].cpu().numpy(), wgl.numpy(), what="d loc")
    C.assert_close(grads["posterior"]["scale_diag"].cpu().numpy(), wgd.numpy(), what="d raw scale_diag")
    C.assert_close(grads["posterior"]["scale_perturb_factor"].cpu().numpy(), wgU.numpy(), what="d scale_perturb_factor")
    for k in kwargs:
        C.assert_close(grads[k].cpu().numpy(), wgk[k].numpy(), what=f"d {k}")
    # autograd route through loss_fn
    q = {"posterior": {k: v.clone().requires_grad_(True) for k, v in params["posterior"].items()}}
    q.update({k: params[k].clone().requires_grad_(True) for k in kwargs})
    model.loss_fn(q, yd, {"X": Xd}, 3 * N, br.PRNGKey(17), S).backward()
    assert torch.equal(q["posterior"]["scale_perturb_factor"].grad, grads["posterior"]["scale_perturb_factor"])
    assert torch.equal(q["posterior"]["scale_diag"].grad, grads["posterior"]["scale_diag"])
    # posterior samples have the covariance A A^T
    if D <= 40:
        post = model.apply(params)
        smp = post.sample(br.PRNGKey(5), (20000,))["weights"].cpu().double().numpy()
        A = np.diag(np.exp(raw_d.astype(np.float64))) + U.astype(np.float64) @ U.astype(np.float64).T
        cov = A @ A.T
        emp = np.cov(smp.T)
        assert np.abs(emp - cov).max() < 0.08 * np.abs(cov).max() + 2e-3
        assert np.abs(smp.mean(0) - loc).max() < 0.02


def test_c4_shape_family_full_rank_wide():
    """config c4's shape family at reduced N: full-rank Gaussian linear regression with D = 1024 and S = 256 -- CTA-pair
    (cta_group::2) GEMM passes with 256-wide tiles, the tensor-core L.eps and dM contractions -- against the float64 oracle."""
    D, S, N = 1024, 256, 1500
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    X, y = C.synth_linear(N, D)
    model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), vi_type="full_rank", kernel="tcgen05_gemm_two_pass")
    kwargs = {"log_noise_scale": np.float32(math.log(0.2))}
    loc = 0.1 * tf.normal(tf.prng_key(6), (D,), 1)
    M = (0.1 * np.eye(D) + 0.002 * tf.normal(tf.prng_key(8), (D, D), 1)).astype(np.float32)  # conditioning stated: 0.1 I + 0.002 noise
    loss, grads = _run_case(model, ref, loc, M, kwargs, y, X, N, 17, S, 0, expect_kernel="tcgen05_gemm_two_pass")
    g = grads["posterior"]["scale_tril"].cpu().numpy()
    assert np.all(np.triu(g, 1) == 0.0)


def _np_fill_triangular(x):
    """tfp.math.fill_triangular(x, upper=False) restated in NumPy (documented example: [1..6] -> [[4,0,0],[6,5,0],[3,2,1]])."""
    m = x.shape[-1]
    n = int((math.sqrt(8 * m + 1) - 1) / 2)
    full = np.concatenate([x[n:], x[::-1]]).reshape(n, n)
    return np.tril(full)


def test_packed_fill_triangular_parameterisation():
    """north_star item 2: L = fill_triangular(packed).  The packed model must give the dense model's loss and d loc, and
    a d packed that is the gather of the dense gradient over the lower triangle, through both gradient routes."""
    from bijax_b200 import random as br
    from bijax_b200.advi import fill_triangular

    assert np.array_equal(_np_fill_triangular(np.arange(1.0, 7.0)), np.array([[4, 0, 0], [6, 5, 0], [3, 2, 1]], float))
    dev = torch.device("cuda")
    got = fill_triangular(torch.arange(1.0, 7.0, device=dev)).cpu().numpy()
    assert np.array_equal(got, np.array([[4, 0, 0], [6, 5, 0], [3, 2, 1]], np.float32))
    D, N, S = 33, 400, 6
    X, y = C.synth_logistic(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    dense_model, _ = C.build_pair(prior, {}, lk.BernoulliLogits(), vi_type="full_rank")
    import bijax_b200 as bj

    packed_model = bj.ADVI(prior, {}, lk.BernoulliLogits(), vi_type="full_rank", packed_tril=True)
    m = D * (D + 1) // 2
    p0 = packed_model.init(br.PRNGKey(4))
    assert p0["posterior"]["scale_tril"].shape == (m,)
    packed = (0.05 * tf.normal(tf.prng_key(8), (m,), 1)).astype(np.float32)
    dense = _np_fill_triangular(packed).astype(np.float32)
    assert np.array_equal(fill_triangular(torch.tensor(packed, device=dev)).cpu().numpy(), dense)
    loc = (0.1 * tf.normal(tf.prng_key(6), (D,), 1)).astype(np.float32)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    pd = {"posterior": {"loc": torch.tensor(loc, device=dev), "scale_tril": torch.tensor(dense + np.eye(D, dtype=np.float32), device=dev)}}
    # same matrix in packed form: add the identity through the packed entries that land on the diagonal
    eye_packed = np.zeros(m, np.float32)
    for i in range(D):
        k = i * D + i
        eye_packed[D + k if k < m - D else m - 1 - (k - (m - D))] = 1.0
    pp = {"posterior": {"loc": torch.tensor(loc, device=dev), "scale_tril": torch.tensor(packed + eye_packed, device=dev)}}
    ld, gd = dense_model.value_and_grad(pd, yd, {"X": Xd}, N, br.PRNGKey(2), S)
    lp, gp = packed_model.value_and_grad(pp, yd, {"X": Xd}, N, br.PRNGKey(2), S)
    assert ld.item() == lp.item()
    assert torch.equal(gd["posterior"]["loc"], gp["posterior"]["loc"])
    gdn = gd["posterior"]["scale_tril"].cpu().numpy()
    want = np.zeros(m, np.float32)
    for i in range(D):
        for j in range(i + 1):
            k = i * D + j
            want[D + k if k < m - D else m - 1 - (k - (m - D))] = gdn[i, j]
    assert np.array_equal(gp["posterior"]["scale_tril"].cpu().numpy(), want)
    # autograd route
    q = {"posterior": {k: v.clone().requires_grad_(True) for k, v in pp["posterior"].items()}}
    packed_model.loss_fn(q, yd, {"X": Xd}, N, br.PRNGKey(2), S).backward()
    assert torch.equal(q["posterior"]["scale_tril"].grad, gp["posterior"]["scale_tril"])


def test_reference_style_minibatch_temporaries_never_reuse_stale_planes():
    """The reference's minibatch pattern ``loss_fn(params, y[idx], X[idx], N, seed)`` frees X[idx] after every call, and the
    caching allocator hands the next X[idx] the same address (version 0, same shape): the prepared-planes cache must not
    take that for the previous batch.  Every step against an engine that never prepares planes and against the float64
    oracle; also a buffer rewritten in place by a raw kernel after ``increment_version``."""
    from bijax_b200 import random as br

    N, B, D, S = 8192, 1024, 128, 16
    X, y = _device_logistic(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.BernoulliLogits())
    raw_model, _ = C.build_pair(prior, {}, lk.BernoulliLogits(), prepare_dataset=False)
    eng = model._engine_for({})[0]
    assert eng.kernel_choice(S, B) == "tcgen05_bf16_split"
    loc = 0.05 * br.normal(br.PRNGKey(5), (D,), partitionable=True)
    rawp = torch.full((D,), -2.0, device=X.device)
    params = {"posterior": {"loc": loc, "scale_diag": rawp}}
    addresses = set()
    g = torch.Generator(device="cuda").manual_seed(0)
    for i in range(5):
        idx = torch.randint(0, N, (B,), device="cuda", generator=g)
        xb = X[idx]
        addresses.add(xb.data_ptr())
        got, grads = model.value_and_grad(params, y[idx], {"X": xb}, N, br.PRNGKey(i), S)
        want, wgrads = raw_model.value_and_grad(params, y[idx], {"X": X[idx]}, N, br.PRNGKey(i), S)
        C.assert_close(got.item(), want.item(), rtol=2e-6, atol_scale=2e-6, what=f"loss of batch {i}")
        C.assert_close(grads["posterior"]["loc"].cpu().numpy(), wgrads["posterior"]["loc"].cpu().numpy(), rtol=1e-5, what=f"d loc of batch {i}")
        (wl, wgl, _, _), _ = C.oracle_step(ref, loc.cpu().numpy(), rawp.cpu().numpy(), {}, y[idx].cpu().numpy(), xb.cpu().numpy(), N,
                                           tf.prng_key(i), S, 0)
        C.assert_close(got.item(), wl.item(), what=f"loss of batch {i} vs oracle")
        C.assert_close(grads["posterior"]["loc"].cpu().numpy(), wgl.numpy(), what=f"d loc of batch {i} vs oracle")
        del xb
    # one persistent batch buffer rewritten in place (what a sampler's own buffers are): the version counter is the signal
    buf, ybuf = X[:B].clone(), y[:B].clone()
    for i in range(3):
        idx = torch.randint(0, N, (B,), device="cuda", generator=g)
        buf.copy_(X[idx])
        ybuf.copy_(y[idx])
        got, _ = model.value_and_grad(params, ybuf, {"X": buf}, N, br.PRNGKey(i), S)
        want, _ = raw_model.value_and_grad(params, ybuf, {"X": buf}, N, br.PRNGKey(i), S)
        C.assert_close(got.item(), want.item(), rtol=2e-6, atol_scale=2e-6, what=f"in-place batch {i}")


def test_fresh_minibatches_fall_back_to_raw_fp32_and_a_repeated_dataset_is_prepared_again():
    """advi.py:92 scales a minibatch to the full data size; a caller passing a NEW tensor every step must not pay a dataset
    preparation per step: after two consecutive cache misses the engine reads raw fp32 X, and prepares again once a tensor
    repeats.  Results do not depend on which route was taken (two-pass path: bit-identical)."""
    from bijax_b200 import random as br

    N, D, S = 2048, 128, 40
    X, y = _device_logistic(4 * N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.BernoulliLogits(), kernel="tcgen05_gemm_two_pass")
    eng = model._engine_for({})[0]
    ref_model, _ = C.build_pair(prior, {}, lk.BernoulliLogits(), kernel="tcgen05_gemm_two_pass", prepare_dataset=False)
    ref = ref_model._engine_for({})[0]
    loc, raw = torch.zeros(D, device=X.device), torch.full((D,), -2.0, device=X.device)
    used_planes = []
    for i in range(4):
        xb, yb = X[i * N : (i + 1) * N].clone(), y[i * N : (i + 1) * N].clone()  # a fresh tensor per step
        a = eng.step(br.PRNGKey(i), loc, raw, None, xb, yb, 4.0, S).clone()
        used_planes.append(eng._planes is not None)
        b = ref.step(br.PRNGKey(i), loc, raw, None, xb, yb, 4.0, S).clone()
        assert torch.equal(a, b)
    assert used_planes == [True, True, False, False]
    xb, yb = X[:N].clone(), y[:N].clone()
    for i in range(3):  # the same tensor again and again: a dataset
        eng.step(br.PRNGKey(9), loc, raw, None, xb, yb, 4.0, S)
    assert eng._planes is not None


def _tc_engine(D, lik=None, precision="auto"):
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lik or lk.BernoulliLogits(), precision=precision)
    return model._engine_for({})[0]


def _device_logistic(N, D, seed=1234):
    """The bench's synthetic data, generated on the device (same generator as bench.py)."""
    from bijax_b200 import random as br

    X = br.normal(br.PRNGKey(seed), (N, D), partitionable=True)
    X.mul_(1.0 / math.sqrt(D))
    w = br.normal(br.PRNGKey(seed + 1), (D,), partitionable=True)
    u = br.uniform(br.PRNGKey(seed + 2), (N,), partitionable=True)
    y = (u < torch.sigmoid(X @ w)).float()
    return X, y


def test_full_size_c2_against_the_oracle():
    """config c2 at BASELINE.json's full size: D=5 (+ learnable log_noise_scale), N=1M, S=64 -- the float64 oracle still
    finishes in seconds at this size, so parity is checked directly."""
    N, D, S = 1_000_000, 5, 64
    X, y = C.synth_linear(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.GaussianLinear())
    loc = np.zeros(D, np.float32)
    raw = np.full(D, math.log(0.1), np.float32)
    _run_case(model, ref, loc, raw, {"log_noise_scale": np.float32(math.log(0.1))}, y, X, N, 0, S, 0, expect_kernel="fp32_smalld")


def test_large_shape_properties_of_the_tensor_core_path():
    """config c3 shape (D=512, S=32) at 2M rows, where no CPU oracle is affordable: size-independent properties.
    (1) idempotence: the step is bit-reproducible; (2) shard additivity: the packed outputs of row shards, with the
    q/prior terms on one shard only, sum to the unsharded output; (3) agreement with the independent fp32 CUDA-core
    kernel on the same device data."""
    from bijax_b200 import random as br

    N, D, S = 2_000_000, 512, 32
    X, y = _device_logistic(N, D)
    eng = _tc_engine(D)
    assert eng.kernel_choice(S, N) == "tcgen05_bf16_split"
    dev = X.device
    loc = 0.02 * br.normal(br.PRNGKey(5), (D,), partitionable=True)
    raw = torch.full((D,), math.log(0.1), device=dev)
    key = br.PRNGKey(123)
    full = eng.step(key, loc, raw, None, X, y, 1.0, S).clone()
    again = eng.step(key, loc, raw, None, X, y, 1.0, S).clone()
    assert torch.equal(full, again)
    cut = 777_777  # deliberately not a multiple of the 64-row tile
    a = eng.step(key, loc, raw, None, X[:cut], y[:cut], 1.0, S, prior_weight=1.0).clone()
    b = eng.step(key, loc, raw, None, X[cut:], y[cut:], 1.0, S, prior_weight=0.0).clone()
    C.assert_close((a + b).cpu().numpy(), full.cpu().numpy(), rtol=2e-6, atol_scale=2e-6, what="shard additivity")
    ref_eng = _tc_engine(D, precision="fp32")
    assert ref_eng.kernel_choice(S, N) == "fp32_tiled"
    other = ref_eng.step(key, loc, raw, None, X, y, 1.0, S)
    C.assert_close(full[:1].cpu().numpy(), other[:1].cpu().numpy(), what="loss vs fp32 kernel")
    C.assert_close(full[1 : 1 + D].cpu().numpy(), other[1 : 1 + D].cpu().numpy(), what="d loc vs fp32 kernel")
    C.assert_close(full[1 + D :].cpu().numpy(), other[1 + D :].cpu().numpy(), what="d rho vs fp32 kernel")


@pytest.mark.parametrize("D,prepared", [(128, True), (384, True), (256, True), (128, False), (384, False),
                                        (768, True), (1024, True), (640, False)])  # > 512: a pair of CTAs per tile, 3-4 chunks each
def test_accumulator_folding_with_few_blocks_per_tile(D, prepared):
    """Enough tiles per CTA (21+) that the periodic folding of the GEMM2 accumulators runs for every block count
    (D = 128, 256, 384 have 1, 2, 3 chunks, so a different folding rotation than D = 512): the fused tensor-core kernels
    against the independent fp32 CUDA-core kernel on the same device data, plus bit-reproducibility."""
    from bijax_b200 import random as br

    N, S = 200_000, 32
    X, y = _device_logistic(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.BernoulliLogits(), prepare_dataset=prepared)
    eng = model._engine_for({})[0]
    assert eng.kernel_choice(S, N) == "tcgen05_bf16_split"
    loc = 0.02 * br.normal(br.PRNGKey(5), (D,), partitionable=True)
    raw = torch.full((D,), math.log(0.1), device=X.device)
    key = br.PRNGKey(77)
    a = eng.step(key, loc, raw, None, X, y, 1.0, S).clone()
    b = eng.step(key, loc, raw, None, X, y, 1.0, S).clone()
    assert torch.equal(a, b)
    ref_eng = _tc_engine(D, precision="fp32")
    other = ref_eng.step(key, loc, raw, None, X, y, 1.0, S)
    C.assert_close(a[:1].cpu().numpy(), other[:1].cpu().numpy(), what="loss vs fp32 kernel")
    C.assert_close(a[1 : 1 + D].cpu().numpy(), other[1 : 1 + D].cpu().numpy(), what="d loc vs fp32 kernel")
    C.assert_close(a[1 + D :].cpu().numpy(), other[1 + D :].cpu().numpy(), what="d rho vs fp32 kernel")


def test_sample_count_invariance_of_eps_layout():
    """eps for (S, D) is the flat stream reshaped, so the first S' < S samples of a partitionable draw are the S'-sample
    draw: sharding the Monte-Carlo samples (small-N case) reproduces the same noise."""
    from bijax_b200 import random as br

    D = 128
    eng = _tc_engine(D)
    dev = torch.device("cuda")
    X, y = _device_logistic(4096, D)
    loc, raw = torch.zeros(D, device=dev), torch.full((D,), -1.0, device=dev)
    e32, e8 = torch.empty(32 * D, device=dev), torch.empty(8 * D, device=dev)
    eng.step(br.PRNGKey(3), loc, raw, None, X, y, 1.0, 32, eps_out=e32, partitionable=True)
    eng.step(br.PRNGKey(3), loc, raw, None, X, y, 1.0, 8, eps_out=e8, partitionable=True)
    assert torch.equal(e32[: 8 * D], e8)


def test_dataset_held_only_as_planes():
    """bj.PreparedX: X streamed through a staging buffer (host tensor in ragged chunks) and kept on the device only as its
    split-bf16 planes.  The planes equal bjx_prepare_x_planes of the whole matrix bit for bit; the step, the host-buffer step,
    the two-pass GEMM path and the device-resident training loop give the very numbers they give on the fp32 tensor; shapes
    whose kernel reads fp32 X refuse it."""
    from functools import partial

    import bijax_b200 as bj
    from bijax_b200 import _native as nv
    from bijax_b200 import optim
    from bijax_b200 import random as br

    N, D, S = 10_000, 200, 32
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    Xp = bj.PreparedX.from_tensor(torch.tensor(X), chunk_rows=3333)  # from the HOST, ragged chunks
    assert Xp.shape == (N, D) and Xp.rows_filled == N
    lib = nv.load()
    whole = torch.empty(int(lib.bjx_x_planes_bytes(N, D)), dtype=torch.uint8, device="cuda")
    nv.check(lib.bjx_prepare_x_planes(Xd.data_ptr(), N, D, D, whole.data_ptr(), nv.current_stream_ptr()))
    assert torch.equal(whole, Xp.planes)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.BernoulliLogits())
    eng = model._engine_for({})[0]
    loc = 0.05 * br.normal(br.PRNGKey(5), (D,), partitionable=True)
    raw = torch.full((D,), -2.0, device="cuda")
    key = br.PRNGKey(9)
    want = eng.step(key, loc, raw, None, Xd, yd, 1.0, S).clone()
    got = eng.step(key, loc, raw, None, Xp, yd, 1.0, S).clone()
    assert eng.kernel_choice(S, N) == "tcgen05_bf16_split" and torch.equal(got, want)
    # through the reference-facing surface, and the autograd route
    params = {"posterior": {"loc": loc, "scale_diag": raw}}
    l1, g1 = model.value_and_grad(params, yd, {"X": Xp}, N, key, S)
    l2, g2 = model.value_and_grad(params, yd, {"X": Xd}, N, key, S)
    assert l1.item() == l2.item() and torch.equal(g1["posterior"]["loc"], g2["posterior"]["loc"])
    # the two-pass GEMM path (S = 128) reads only planes as well
    assert eng.kernel_choice(128, N) == "tcgen05_gemm_two_pass"
    w128 = eng.step(key, loc, raw, None, Xd, yd, 1.0, 128).clone()
    g128 = eng.step(key, loc, raw, None, Xp, yd, 1.0, 128).clone()
    assert torch.equal(g128, w128)
    # the device-resident training loop
    fresh = lambda: model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
    a = bj.train(partial(model.loss_fn, outputs=yd, inputs={"X": Xp}, full_data_size=N, n_samples=S), fresh(), optim.adam(0.02), 40, br.PRNGKey(3))
    b = bj.train(partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=S), fresh(), optim.adam(0.02), 40, br.PRNGKey(3))
    assert torch.equal(a["losses"], b["losses"])
    # minibatches drawn on the device out of the planes alone
    mb_p = model.minibatch_loss_fn(yd, {"X": Xp}, batch_size=2048, n_samples=S, copy_free=True)
    mb_t = model.minibatch_loss_fn(yd, {"X": Xd}, batch_size=2048, n_samples=S, copy_free=True)
    vp, _ = bj.value_and_grad(mb_p)(params, br.PRNGKey(21))
    vt, _ = bj.value_and_grad(mb_t)(params, br.PRNGKey(21))
    assert vp.item() == vt.item() and mb_p._state["planes_full"] is Xp.planes
    # a shape whose kernel reads fp32 X
    small = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(5, np.float32), scale_diag=np.ones(5, np.float32))}, {}, lk.BernoulliLogits())
    X5 = bj.PreparedX.from_tensor(torch.zeros(100, 5))
    with pytest.raises(TypeError):
        small.value_and_grad({"posterior": {"loc": torch.zeros(5, device="cuda"), "scale_diag": torch.zeros(5, device="cuda")}},
                             torch.zeros(100, device="cuda"), {"X": X5}, 100, key, 4)
    with pytest.raises(RuntimeError):  # incomplete
        half = bj.PreparedX(10, D)
        half.append(torch.zeros(4, D))
        eng.step(key, loc, raw, None, half, yd[:10], 1.0, S)


def test_sample_slices_under_sample_sharding():
    """Two sample shards of 48 + 48 samples (each more than one slice of 32) at D = 128 sum to the unsharded 96-sample step, and
    the unsharded step matches the float64 oracle: the per-slice partial blocks, the global eps index and the sharding weights
    compose."""
    from bijax_b200 import random as br

    N, D, S = 6000, 128, 96
    X, y = C.synth_logistic(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.BernoulliLogits())
    eng = model._engine_for({})[0]
    assert eng.kernel_choice(S, N) == "tcgen05_bf16_split" and eng.kernel_choice(48, N) == "tcgen05_bf16_split"
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    loc = 0.05 * tf.normal(tf.prng_key(5), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32)
    lt, rt = torch.tensor(loc, device="cuda"), torch.tensor(raw, device="cuda")
    key = br.PRNGKey(13)
    full = eng.step(key, lt, rt, None, Xd, yd, 1.0, S, partitionable=True).clone()
    a = eng.step(key, lt, rt, None, Xd, yd, 1.0, 48, prior_weight=1.0, sample_shard=(0, S, 1.0), partitionable=True).clone()
    b = eng.step(key, lt, rt, None, Xd, yd, 1.0, 48, prior_weight=1.0, sample_shard=(48, S, 0.0), partitionable=True).clone()
    C.assert_close((a + b).cpu().numpy(), full.cpu().numpy(), rtol=2e-6, atol_scale=2e-6, what="sample shards")
    (wl, wgl, wgr, _), _ = C.oracle_step(ref, loc, raw, {}, y, X, N, tf.prng_key(13), S, 1)
    got = full.cpu().numpy()
    C.assert_close(got[0], wl.item(), what="loss")
    C.assert_close(got[1 : 1 + D], wgl.numpy(), what="d loc")
    C.assert_close(got[1 + D :], wgr.numpy(), what="d raw scale")

"""GPU: the bijax-shaped surface -- init / loss_fn / apply / train -- behaves like the reference's."""
import math

import numpy as np
import pytest
import torch

import bijax_b200 as bj
from bijax_b200 import bijectors as tfb
from bijax_b200 import distributions as tfd
from bijax_b200 import likelihoods as lk
from bijax_b200 import optim
from oracle import threefry as tf

pytestmark = pytest.mark.gpu


def test_init_matches_initialize_params():
    """advi.py:77-78 + utils.py:39-55: params = normal(seed, (P,)) unraveled in (loc, scale) order."""
    from bijax_b200 import random as br

    D = 5
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D), scale_diag=np.ones(D))}
    for vi, P in (("mean_field", D), ("full_rank", D * D)):
        model = bj.ADVI(prior, {}, lk.GaussianLinear(), vi_type=vi)
        params = model.init(br.PRNGKey(0))
        flat = tf.normal(tf.prng_key(0), (D + P,), tf.LAYOUT_LEGACY)
        assert np.array_equal(params["posterior"]["loc"].cpu().numpy(), flat[:D])
        skey = "scale_diag" if vi == "mean_field" else "scale_tril"
        assert np.array_equal(params["posterior"][skey].cpu().numpy().ravel(), flat[D:])
        assert params["posterior"][skey].shape == ((D,) if vi == "mean_field" else (D, D))


def test_coin_toss_trains_to_the_analytic_posterior():
    """coin_toss.ipynb: tosses [0,1,1,1,1,0], prior Beta(1,2) -> posterior Beta(5,4), log evidence -4.941642.
    PyMC's ADVI reports 4.9947 and NumPyro's 4.9520 on the same model; the converged -ELBO must sit in that band."""
    from bijax_b200 import random as br
    from functools import partial

    y = torch.tensor([0, 1, 1, 1, 1, 0], dtype=torch.float32, device="cuda")
    model = bj.ADVI({"p_of_heads": tfd.Beta(1.0, 2.0)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
    params = model.init(br.PRNGKey(0))
    loss_fn = partial(model.loss_fn, outputs=y, inputs=None, full_data_size=6, n_samples=64)
    res = bj.train(loss_fn, params, optim.adam(0.05), n_epochs=400, seed=br.PRNGKey(1), return_args={"params_list"})
    losses = res["losses"].cpu().numpy()
    assert losses.shape == (400,) and len(res["params_list"]) == 400
    final = float(losses[-100:].mean())
    assert 4.9416 <= final + 0.03 and final < 5.05, final
    post = model.apply(res["params"])
    samples = post.sample(br.PRNGKey(2), (4000,))["p_of_heads"].cpu().numpy()
    assert samples.shape == (4000,) and 0.0 < samples.min() and samples.max() < 1.0
    assert abs(samples.mean() - 5.0 / 9.0) < 0.03  # Beta(5,4) mean
    lp = post.log_prob({"p_of_heads": torch.tensor([0.3, 0.55, 0.8], device="cuda")}, (3,))
    # the fitted logit-normal must be close to the exact Beta(5,4) density near its bulk
    exact = [math.log(280.0) + 4 * math.log(p) + 3 * math.log(1 - p) for p in (0.3, 0.55, 0.8)]
    assert np.allclose(lp.cpu().numpy(), exact, atol=0.25)


def test_linear_regression_recovers_weights_and_noise():
    from bijax_b200 import random as br
    from functools import partial
    from tests import cases as C

    N, D = 20000, 5
    X, y = C.synth_linear(N, D, noise=0.1)
    w_true = tf.normal(tf.prng_key(1235), (D,), tf.LAYOUT_PARTITIONABLE)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D), scale_diag=np.ones(D) * 3)}, {}, lk.GaussianLinear())
    params = model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
    params["log_noise_scale"] = torch.tensor(0.0, device="cuda")
    loss_fn = partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=8)
    res = bj.train(loss_fn, params, optim.adam(0.05), n_epochs=600, seed=br.PRNGKey(3))
    w = res["params"]["posterior"]["loc"].cpu().numpy()
    assert np.allclose(w, w_true, atol=0.05), (w, w_true)
    assert abs(math.exp(res["params"]["log_noise_scale"].item()) - 0.1) < 0.02


def test_posterior_sample_matches_prologue_arithmetic():
    from bijax_b200 import random as br

    prior = {"a": tfd.Gamma(2.0, 2.0), "b": tfd.Normal(np.zeros(3), 1.0)}
    model = bj.ADVI(prior, {"a": tfb.Exp()}, lk.GaussianLinear(weights="b", noise_latent="a"))
    params = model.init(br.PRNGKey(4))
    s = model.apply(params).sample(br.PRNGKey(5), (7, 2))
    assert s["a"].shape == (7, 2) and s["b"].shape == (7, 2, 3)
    eps = tf.normal(tf.prng_key(5), (14, 4), tf.LAYOUT_LEGACY)
    loc = params["posterior"]["loc"].cpu().numpy()
    sig = np.exp(params["posterior"]["scale_diag"].cpu().numpy())
    z = loc + sig * eps
    assert np.allclose(s["a"].cpu().numpy().ravel(), np.exp(z[:, 0]), rtol=2e-6)
    assert np.allclose(s["b"].cpu().numpy().reshape(14, 3), z[:, 1:], rtol=2e-6, atol=1e-7)


def test_device_resident_training_loop_matches_the_host_driven_loop():
    """train_fn's scan (utils.py:22-31) as bjx_train_steps inside replayed CUDA graphs: same keys, same Adam, so the
    trajectory must match the generic loop (value_and_grad -> optimizer.update -> apply_updates on the host) to float
    rounding of the bias correction, and the graph-replayed loop must equal the eagerly enqueued one bit for bit."""
    from bijax_b200 import random as br
    from bijax_b200 import utils as U
    from functools import partial
    from tests import cases as C

    N, D = 4096, 5
    X, y = C.synth_linear(N, D, noise=0.1)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D), scale_diag=np.ones(D))}, {}, lk.GaussianLinear())

    def fresh():
        p = model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
        p["log_noise_scale"] = torch.tensor(-1.0, device="cuda")
        return p

    loss_fn = partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=8)
    n = 3 * U.GRAPH_STEPS + 5
    fused = bj.train(loss_fn, fresh(), optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3))
    host = bj.train(loss_fn, fresh(), optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3), fused=False)
    assert fused["losses"].shape == host["losses"].shape == (n,)
    assert torch.allclose(fused["losses"], host["losses"], rtol=2e-5, atol=1e-5)
    for k in ("loc", "scale_diag"):
        assert torch.allclose(fused["params"]["posterior"][k], host["params"]["posterior"][k], rtol=1e-4, atol=1e-5)
    assert torch.allclose(fused["params"]["log_noise_scale"], host["params"]["log_noise_scale"], rtol=1e-4, atol=1e-5)
    assert fused["params"]["log_noise_scale"].shape == ()
    old = U.GRAPH_STEPS
    try:
        U.GRAPH_STEPS = 10 ** 9  # no graph: everything enqueued eagerly
        eager = bj.train(loss_fn, fresh(), optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3))
    finally:
        U.GRAPH_STEPS = old
    assert torch.equal(eager["losses"], fused["losses"])
    assert torch.equal(eager["params"]["posterior"]["loc"], fused["params"]["posterior"]["loc"])


@pytest.mark.parametrize("D", [100, 640])
def test_graph_replayed_training_loop_on_the_tall_tile_and_cta_pair_kernels(D):
    """The device-resident loop captured in CUDA graphs when the step's likelihood kernel is k_lik_tc_tall (D = 100) or the
    cluster launch of the CTA-pair variant (D = 640: cudaLaunchKernelEx inside stream capture): graph replay equals the
    eagerly enqueued loop bit for bit, and both follow the host-driven loop."""
    from bijax_b200 import random as br
    from bijax_b200 import utils as U
    from functools import partial
    from tests import cases as C

    N = 6000
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D), scale_diag=np.ones(D))}, {}, lk.BernoulliLogits())
    loss_fn = partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=16)
    fresh = lambda: model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
    n = 2 * U.GRAPH_STEPS + 3
    fused = bj.train(loss_fn, fresh(), optim.adam(0.01), n_epochs=n, seed=br.PRNGKey(3))
    eng = next(iter(model._engines.values()))
    assert eng.kernel_choice(16, N) == "tcgen05_bf16_split"
    host = bj.train(loss_fn, fresh(), optim.adam(0.01), n_epochs=n, seed=br.PRNGKey(3), fused=False)
    assert torch.allclose(fused["losses"], host["losses"], rtol=2e-5, atol=1e-5)
    for k in ("loc", "scale_diag"):
        assert torch.allclose(fused["params"]["posterior"][k], host["params"]["posterior"][k], rtol=1e-4, atol=1e-5)
    old = U.GRAPH_STEPS
    try:
        U.GRAPH_STEPS = 10 ** 9
        eager = bj.train(loss_fn, fresh(), optim.adam(0.01), n_epochs=n, seed=br.PRNGKey(3))
    finally:
        U.GRAPH_STEPS = old
    assert torch.equal(eager["losses"], fused["losses"])
    assert torch.equal(eager["params"]["posterior"]["loc"], fused["params"]["posterior"]["loc"])


@pytest.mark.parametrize("vi", ["mean_field", "full_rank", "low_rank"])
def test_posterior_log_prob_kernel_against_the_oracle(vi):
    """Posterior.log_prob (core.py:57-71) as one fused kernel: inverse bijector chains + log-det-Jacobians + diagonal /
    triangular-solve Gaussian density, against the float64 restatement, for constrained latents of every supported kind."""
    from bijax_b200 import random as br
    from tests import cases as C

    Dw = 70 if vi != "mean_field" else 6
    prior = {
        "weights": tfd.Normal(np.zeros(Dw, np.float32), 1.5),
        "noise": tfd.Gamma(2.0, 3.0),
        "aux_beta": tfd.Beta(np.array([2.0, 0.5], np.float32), np.array([3.0, 1.0], np.float32)),
        "aux_pos": tfd.Gamma(1.0, 0.5),
        "aux_chain": tfd.Normal(1.0, 0.7),
    }
    bij = {"noise": tfb.Exp(), "aux_beta": tfb.Sigmoid(), "aux_pos": tfb.Softplus(), "aux_chain": tfb.Chain([tfb.Softplus(), tfb.Exp()])}
    if vi == "low_rank":
        model = bj.ADVI(prior, bij, lk.GaussianLinear(noise_latent="noise"), vi_type=vi, rank=4)
        _, ref = C.build_pair(prior, bij, lk.GaussianLinear(noise_latent="noise"))
        ref.vi_type = "low_rank"
    else:
        model, ref = C.build_pair(prior, bij, lk.GaussianLinear(noise_latent="noise"), vi_type=vi)
    D = model.size
    loc = (0.3 * tf.normal(tf.prng_key(31), (D,), 1)).astype(np.float32)
    raw64 = None
    if vi == "low_rank":
        raw = (-1.0 + 0.2 * tf.normal(tf.prng_key(32), (D,), 1)).astype(np.float32)
        Uf = (0.2 * tf.normal(tf.prng_key(33), (D, 4), 1)).astype(np.float32)
        skey = "scale_diag"
        raw64 = (torch.tensor(raw, dtype=torch.float64), torch.tensor(Uf, dtype=torch.float64))
    elif vi == "mean_field":
        raw = (-1.0 + 0.2 * tf.normal(tf.prng_key(32), (D,), 1)).astype(np.float32)
        skey = "scale_diag"
    else:
        raw = (0.3 * np.eye(D) + 0.02 * tf.normal(tf.prng_key(32), (D, D), 1)).astype(np.float32)
        skey = "scale_tril"
    params = {"posterior": {"loc": torch.tensor(loc, device="cuda"), skey: torch.tensor(raw, device="cuda")}}
    if vi == "low_rank":
        params["posterior"]["scale_perturb_factor"] = torch.tensor(Uf, device="cuda")
    post = model.apply(params)
    n = 37
    samples = post.sample(br.PRNGKey(9), (n,))
    got = post.log_prob(samples, (n,)).cpu().numpy()
    assert got.shape == (n,)
    t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    cpu = {k: v.cpu().double() for k, v in samples.items()}
    want = np.array([ref.posterior_log_prob(t64(loc), raw64 if raw64 is not None else t64(raw), {k: v[i] for k, v in cpu.items()}).item()
                     for i in range(n)])
    assert np.allclose(got, want, rtol=2e-5, atol=2e-4), np.abs(got - want).max()
    p = post.prob(samples, (n,)).cpu().numpy()
    assert np.allclose(p, np.exp(want), rtol=1e-3)
    # batch of shape (2, 3): same numbers, reshaped
    s6 = {k: v[:6].reshape(2, 3, *v.shape[1:]) for k, v in samples.items()}
    assert np.allclose(post.log_prob(s6, (2, 3)).cpu().numpy().ravel(), got[:6])


@pytest.mark.parametrize("D,kernel", [(6, None), (192, "tcgen05_gemm_two_pass"), (128, "tcgen05_bf16_split")])
def test_map_and_mcmc_log_joint_on_the_likelihood_kernels(D, kernel):
    """bijax/map.py:27-33 and bijax/mcmc.py:30-34 as point evaluations of the fused step (S = 1, eps = 0, point_mode):
    value and gradient with respect to the unconstrained tree against the float64 restatement, constrained latents
    included; MAP uses the prior density of the constrained value, MCMC the density of z (with the log-det-Jacobian)."""
    from tests import cases as C

    N = 700
    X, y = C.synth_linear(N, D, noise=0.5)
    prior = {
        "weights": tfd.Normal(np.zeros(D, np.float32), 1.5),
        "noise": tfd.Gamma(2.0, 3.0),
        "aux_beta": tfd.Beta(np.array([2.0, 0.5], np.float32), np.array([3.0, 1.0], np.float32)),
        "aux_chain": tfd.Normal(1.0, 0.7),
    }
    bij = {"noise": tfb.Exp(), "aux_beta": tfb.Sigmoid(), "aux_chain": tfb.Chain([tfb.Softplus(), tfb.Exp()])}
    lik = lk.GaussianLinear(noise_latent="noise")
    _, ref = C.build_pair(prior, bij, lik)
    dev = torch.device("cuda")
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    zflat = (0.3 * tf.normal(tf.prng_key(31), (ref.size,), 1)).astype(np.float32)
    tree64 = ref.unravel(torch.tensor(zflat, dtype=torch.float64))
    for cls, mode in ((bj.MAP, "map"), (bj.MCMC, "mcmc")):
        model = cls(prior, lik, bij, kernel=kernel) if cls is bj.MAP else cls(prior, bij, lik, kernel=kernel)
        ztree = {k: v.to(torch.float32).to(dev).clone().requires_grad_(True) for k, v in tree64.items()}
        if mode == "map":
            val = model.neg_log_joint({"params": ztree}, yd, {"X": Xd})
        else:
            val = model.log_joint(ztree, yd, {"X": Xd})
        val.backward()
        z64 = torch.tensor(zflat, dtype=torch.float64, requires_grad=True)
        t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
        want = ref.map_neg_log_joint(z64, t64(y), {"X": t64(X)}) if mode == "map" else ref.mcmc_log_joint(z64, t64(y), {"X": t64(X)})
        (gw,) = torch.autograd.grad(want, z64)
        C.assert_close(val.item(), want.item(), what=f"{mode} value")
        got = torch.cat([ztree[k].grad.reshape(-1) for k in ref.names]).cpu().numpy()
        C.assert_close(got, gw.numpy(), what=f"{mode} gradient")
    # transform_tree
    th = model.transform({k: v.detach() for k, v in ztree.items()})
    assert torch.allclose(th["noise"].cpu().double(), torch.exp(tree64["noise"]), rtol=1e-6)
    assert torch.allclose(th["aux_beta"].cpu().double(), torch.sigmoid(tree64["aux_beta"]), rtol=1e-6)


def test_map_training_recovers_the_ridge_solution():
    """MAP for Gaussian linear regression with a Normal(0, s) prior is ridge regression: train_fn in its seed=None form
    (utils.py:10-19) on MAP.neg_log_joint converges to the closed-form solution."""
    from bijax_b200 import random as br
    from functools import partial
    from tests import cases as C

    N, D, tau, s0 = 5000, 8, 0.3, 2.0
    X, y = C.synth_linear(N, D, noise=tau)
    dev = torch.device("cuda")
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    model = bj.MAP({"weights": tfd.Normal(np.zeros(D, np.float32), s0)}, lk.GaussianLinear(noise_scale=tau), None)
    params = model.init(br.PRNGKey(0))
    assert params["params"]["weights"].shape == (D,)
    loss_fn = partial(model.neg_log_joint, outputs=yd, inputs={"X": Xd})
    res = bj.train(loss_fn, params, optim.adam(0.05), n_epochs=800)
    w = res["params"]["params"]["weights"].cpu().numpy().astype(np.float64)
    X64, y64 = X.astype(np.float64), y.astype(np.float64)
    ridge = np.linalg.solve(X64.T @ X64 / tau**2 + np.eye(D) / s0**2, X64.T @ y64 / tau**2)
    assert np.allclose(w, ridge, atol=2e-3), np.abs(w - ridge).max()
    assert res["losses"][-1] < res["losses"][0]
    assert torch.equal(model.apply(res["params"])["weights"].cpu(), res["params"]["params"]["weights"].cpu())


def test_device_resident_loop_with_the_low_rank_family():
    """The fused training loop carries [loc | raw scale_diag | scale_perturb_factor | kwargs] as one flat vector."""
    from bijax_b200 import random as br
    from functools import partial
    from tests import cases as C

    N, D, r = 2048, 12, 3
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D), scale_diag=np.ones(D))}, {}, lk.BernoulliLogits(),
                    vi_type="low_rank", rank=r)
    fresh = lambda: model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.05 * br.normal(k, shape))
    loss_fn = partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=8)
    fused = bj.train(loss_fn, fresh(), optim.adam(0.02), n_epochs=40, seed=br.PRNGKey(3))
    host = bj.train(loss_fn, fresh(), optim.adam(0.02), n_epochs=40, seed=br.PRNGKey(3), fused=False)
    assert fused["params"]["posterior"]["scale_perturb_factor"].shape == (D, r)
    assert torch.allclose(fused["losses"], host["losses"], rtol=2e-5, atol=1e-4)
    for k in ("loc", "scale_diag", "scale_perturb_factor"):
        assert torch.allclose(fused["params"]["posterior"][k], host["params"]["posterior"][k], rtol=1e-4, atol=2e-5), k


def test_peer_memory_allreduce_kernel_single_rank(tmp_path):
    """The one-kernel exchange step (bjx_allreduce_packed_p2p) on a world of one: slot push, epoch flag, rank-order sum and
    the alternation of the two slot sets over many epochs leave the vector unchanged.  (Worlds of 2 and 8 are exercised by
    scripts/p2p_check.py under torchrun: bit-equal to the rank-order sum, profiles/ and DESIGN.md section 6.)"""
    import torch.distributed as dist

    from bijax_b200.parallel import P2PAllReduce

    dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        op = P2PAllReduce(1025, None, torch.device("cuda"))
        for i in range(5):
            v = torch.randn(1025, device="cuda")
            w = v.clone()
            op(w)
            assert torch.equal(v, w)
        assert op.epoch == 5
        op.close()
    finally:
        dist.destroy_process_group()


def test_device_resident_loop_captures_the_two_pass_tensor_core_path():
    """CUDA-graph capture of bjx_train_steps over the TMA-fed two-pass path (tensor-map encodes on the host at capture time,
    memset nodes, CTA-pair cluster launches): graph replay equals the host-driven loop."""
    from bijax_b200 import random as br
    from functools import partial
    from tests import cases as C

    N, D, S = 3000, 128, 160  # S > 128: 256-wide tiles, cta_group::2
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D), scale_diag=np.ones(D))}, {}, lk.BernoulliLogits(),
                    kernel="tcgen05_gemm_two_pass")
    fresh = lambda: model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.05 * br.normal(k, shape))
    loss_fn = partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=S)
    fused = bj.train(loss_fn, fresh(), optim.adam(0.02), n_epochs=36, seed=br.PRNGKey(3))
    host = bj.train(loss_fn, fresh(), optim.adam(0.02), n_epochs=36, seed=br.PRNGKey(3), fused=False)
    assert torch.allclose(fused["losses"], host["losses"], rtol=2e-5, atol=1e-4)
    assert torch.allclose(fused["params"]["posterior"]["loc"], host["params"]["posterior"]["loc"], rtol=1e-4, atol=2e-5)


def test_row_sharded_device_resident_loop_on_a_world_of_one(tmp_path):
    """bjx_train_steps_sharded (ELBO -> peer-memory all-reduce with the epoch read from the device step counter -> Adam, in
    replayed CUDA graphs): on a world of one the exchange is the identity, so the trajectory must equal the single-device
    loop bit for bit, and the exchange's epoch counter must have advanced by one per step.  (Two ranks:
    scripts/multi_gpu_train_check.py under torchrun; tests/test_multi_gpu.py runs it when the box has two GPUs.)"""
    import torch.distributed as dist
    from functools import partial
    from bijax_b200 import random as br
    from bijax_b200 import utils as U
    from tests import cases as C

    N, D, S = 4096, 128, 8
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}

    def run(sharded):
        model = bj.ADVI(prior, {}, lk.BernoulliLogits())
        if sharded:
            model.distributed(shard="rows", p2p=True)
        params = model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
        loss_fn = partial(model.loss_fn, outputs=yd, inputs={"X": Xd}, full_data_size=N, n_samples=S)
        n = 2 * U.GRAPH_STEPS + 3
        res = bj.train(loss_fn, params, optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3))
        return model, res, n

    _, single, n = run(False)
    dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        model, sharded, _ = run(True)
        assert model._p2p_op is not None and model._p2p_op.epoch == n + 1  # + the warm-up step before capture
        assert torch.equal(single["losses"], sharded["losses"])
        for k in ("loc", "scale_diag"):
            assert torch.equal(single["params"]["posterior"][k], sharded["params"]["posterior"][k])
        model._p2p_op.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["bernoulli_logits_d128", "gaussian_kwarg_d5", "gaussian_latent_noise_d40", "coin"])
def test_likelihood_objects_are_the_reference_callable(case):
    """README.md:81-106: ``log_likelihood_fn(latent_sample, outputs, inputs, **kwargs) -> scalar``.  The declarative objects
    ARE that callable (a point evaluation on the same CUDA kernels, BJX_POINT_LOGLIK): value and gradients with respect to the
    constrained latent leaves and the trainable kwargs against the float64 restatement of the same exemplar."""
    from oracle import advi_ref as R
    from tests import cases as C

    dev = torch.device("cuda")
    kwargs, okw = {}, {}
    if case == "coin":
        L, ofn = lk.BernoulliProbs("p"), R.bernoulli_probs_loglik("p")
        y = np.array([1.0 if n % 10 < 7 else 0.0 for n in range(777)], dtype=np.float32)
        X, latent = None, {"p": np.float32(0.37)}
    else:
        D = {"bernoulli_logits_d128": 128, "gaussian_kwarg_d5": 5, "gaussian_latent_noise_d40": 40}[case]
        N = 3001
        w = (0.3 * tf.normal(tf.prng_key(4), (D,), 1)).astype(np.float32)
        if case == "bernoulli_logits_d128":
            X, y = C.synth_logistic(N, D)
            L, ofn, latent = lk.BernoulliLogits(), R.bernoulli_logits_loglik(), {"weights": w}
        elif case == "gaussian_kwarg_d5":
            X, y = C.synth_linear(N, D)
            L, ofn, latent = lk.GaussianLinear(), R.gaussian_linear_loglik(), {"weights": w}
            kwargs = {"log_noise_scale": np.float32(-0.4)}
        else:
            X, y = C.synth_linear(N, D)
            L = lk.GaussianLinear(noise_latent="noise")
            ofn, latent = C.oracle_likelihood(L), {"weights": w, "noise": np.float32(0.8), "unused": np.float32(3.0)}
    t = lambda a, req=True: torch.tensor(np.asarray(a, dtype=np.float32), device=dev, requires_grad=req)
    lat_t = {k: t(v, k != "unused") for k, v in latent.items()}
    kw_t = {k: t(v) for k, v in kwargs.items()}
    inputs = {"X": torch.tensor(X, device=dev)} if X is not None else None
    got = L(lat_t, torch.tensor(y, device=dev), inputs, **kw_t)
    assert got.shape == ()
    got.backward()
    t64 = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    lat_o = {k: t64(v) for k, v in latent.items() if k != "unused"}
    kw_o = {k: t64(v) for k, v in kwargs.items()}
    want = ofn(lat_o, torch.tensor(y, dtype=torch.float64), {"X": torch.tensor(X, dtype=torch.float64)} if X is not None else None, **kw_o)
    want.backward()
    C.assert_close(got.item(), want.item(), what="log-likelihood")
    for k in lat_o:
        C.assert_close(lat_t[k].grad.cpu().numpy(), lat_o[k].grad.numpy(), what=f"d / d {k}")
    for k in kw_o:
        C.assert_close(kw_t[k].grad.cpu().numpy(), kw_o[k].grad.numpy(), what=f"d / d {k}")
    with pytest.raises(KeyError):
        L({}, torch.tensor(y, device=dev), inputs, **kw_t)

"""GPU: the optimiser side of train_fn (utils.py:22-31) against an independent float64 restatement -- oracle ELBO
(oracle/advi_ref.py) + oracle optax.adam (oracle/optax_ref.py) + oracle key split -- instead of against the same
``bjx_adam_update`` kernel driven from the host."""
from functools import partial

import numpy as np
import pytest
import torch

import bijax_b200 as bj
from bijax_b200 import _native as nv
from bijax_b200 import bijectors as tfb
from bijax_b200 import distributions as tfd
from bijax_b200 import likelihoods as lk
from bijax_b200 import optim
from bijax_b200 import random as br
from oracle import optax_ref as O
from oracle import threefry as tf
from tests import cases as C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count", [1, 2, 10, 1000, 100000])
def test_adam_kernel_against_the_optax_restatement(count):
    """One bjx_adam_update at step ``count`` from arbitrary moments: new moments and the applied update."""
    rng = np.random.default_rng(count)
    n, lr, b1, b2, eps = 1037, 0.02, 0.9, 0.999, 1e-8
    g = rng.normal(size=n).astype(np.float32) * np.float32(10.0) ** rng.integers(-4, 3, size=n).astype(np.float32)
    m0 = (0.3 * g + 0.1 * rng.normal(size=n)).astype(np.float32)
    v0 = (np.abs(rng.normal(size=n)) * g.astype(np.float64) ** 2).astype(np.float32)
    p0 = rng.normal(size=n).astype(np.float32)
    st = O.adam_init(n)
    st.count, st.mu, st.nu = count - 1, m0.astype(np.float64), v0.astype(np.float64)
    u = O.adam_update(g, st, lr, b1, b2, eps)
    p, m, v, gd = (torch.tensor(a, device="cuda") for a in (p0, m0, v0, g))
    nv.check(nv.load().bjx_adam_update(p.data_ptr(), m.data_ptr(), v.data_ptr(), gd.data_ptr(), n, count, lr, b1, b2, eps,
                                       nv.current_stream_ptr()))
    # mu mixes signs (cancellation between b1*m and (1-b1)*g): float32 rounding of the larger term is the floor;
    # nu is a sum of positives: a few float32 ulps relative, which also proves (1 - b2) is the float nearest to 0.001
    C.assert_close(m.cpu().numpy(), st.mu, rtol=2e-6, atol_scale=2e-7, what="mu")
    C.assert_close(v.cpu().numpy(), st.nu, rtol=5e-7, atol_scale=0.0, what="nu")
    C.assert_close(p.cpu().numpy(), p0 + u, rtol=1e-5, atol_scale=2e-7, what="params")


def _oracle_training(ref, flat0, D, P, kw_names, y, X, full, seed, n, S, lr, layout=tf.LAYOUT_LEGACY):
    keys = tf.split(tf.prng_key(seed), n, layout)  # utils.py:30
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    inputs = {"X": t(X)} if X is not None else None
    yt = t(y)

    def vg(p, key):
        eps = torch.tensor(tf.normal(key, (S, D), layout), dtype=torch.float64)
        kw = {k: t(p[D + P + i]) for i, k in enumerate(kw_names)}
        raw = p[D : D + P] if P == D else p[D : D + P].reshape(D, D)
        loss, gl, gr, gk = ref.value_and_grad(t(p[:D]), t(raw), kw, yt, inputs, full, eps)
        return loss.item(), np.concatenate([gl.numpy().ravel(), gr.numpy().ravel()] + [gk[k].numpy().ravel() for k in kw_names])

    return O.train(vg, flat0, keys, lr)


@pytest.mark.parametrize("fused", [True, False], ids=["device_resident_loop", "host_driven_loop"])
def test_c1_training_against_oracle_elbo_and_oracle_adam(fused):
    """config c1 (coin toss, S=16): 48 epochs of train_fn -- as bjx_train_steps in replayed CUDA graphs, and host-driven --
    against the float64 trajectory."""
    model, ref = C.build_pair({"p_of_heads": tfd.Beta(0.5, 0.5)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
    y = np.array([1.0 if n % 10 < 7 else 0.0 for n in range(100)], dtype=np.float32)
    n, S, lr = 48, 16, 0.05
    params = model.init(br.PRNGKey(0))
    post = params[next(k for k in params if "posterior" in k)]
    flat0 = np.concatenate([post["loc"].cpu().numpy().ravel(), post["scale_diag"].cpu().numpy().ravel()])
    loss_fn = partial(model.loss_fn, outputs=torch.tensor(y, device="cuda"), inputs=None, full_data_size=100, n_samples=S)
    res = bj.train(loss_fn, params, optim.adam(lr), n_epochs=n, seed=br.PRNGKey(1), fused=fused)
    want = _oracle_training(ref, flat0, 1, 1, [], y, None, 100, 1, n, S, lr)
    got_post = res["params"][next(k for k in res["params"] if "posterior" in k)]
    got = np.concatenate([got_post["loc"].cpu().numpy().ravel(), got_post["scale_diag"].cpu().numpy().ravel()])
    C.assert_close(res["losses"].cpu().numpy(), want["losses"], rtol=2e-5, atol_scale=2e-5, what="losses")
    assert np.abs(got - want["params"]).max() < 2e-4, (got, want["params"])


@pytest.mark.parametrize("fused", [True, False], ids=["device_resident_loop", "host_driven_loop"])
def test_small_c2_training_against_oracle_elbo_and_oracle_adam(fused):
    """config c2's model (D=5 + learnable log_noise_scale, S=64) on 4096 rows: 40 epochs against the float64 trajectory."""
    N, D, S, n, lr = 4096, 5, 64, 40, 0.02
    X, y = C.synth_linear(N, D, noise=0.1)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.GaussianLinear())
    params = model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
    params["log_noise_scale"] = torch.tensor(-1.0, device="cuda")
    post = params[next(k for k in params if "posterior" in k)]
    flat0 = np.concatenate([post["loc"].cpu().numpy().ravel(), post["scale_diag"].cpu().numpy().ravel(), [-1.0]])
    loss_fn = partial(model.loss_fn, outputs=torch.tensor(y, device="cuda"), inputs={"X": torch.tensor(X, device="cuda")},
                      full_data_size=N, n_samples=S)
    res = bj.train(loss_fn, params, optim.adam(lr), n_epochs=n, seed=br.PRNGKey(3), fused=fused)
    want = _oracle_training(ref, flat0, D, D, ["log_noise_scale"], y, X, N, 3, n, S, lr)
    got_post = res["params"][next(k for k in res["params"] if "posterior" in k)]
    got = np.concatenate([got_post["loc"].cpu().numpy().ravel(), got_post["scale_diag"].cpu().numpy().ravel(),
                          res["params"]["log_noise_scale"].cpu().numpy().ravel()])
    C.assert_close(res["losses"].cpu().numpy(), want["losses"], rtol=5e-5, atol_scale=2e-5, what="losses")
    assert np.abs(got - want["params"]).max() < 3e-4, (got, want["params"])

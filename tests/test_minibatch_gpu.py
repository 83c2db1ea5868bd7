"""GPU: the on-device minibatch sampler: indices bit-exact against the oracle's restatement of
jax.random.randint / choice, gathered rows (raw fp32 and prepared planes) equal to X[idx], the minibatch loss equal to
loss_fn on those rows with the split-off seed, and the device-resident minibatch loop equal to the host-driven one."""
import ctypes as C_
import math

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
from oracle import threefry as tf
from tests import cases as C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("N,B", [(1, 5), (100, 1), (100, 4096), (65536, 1000), (65537, 1000), (10_000_000, 4099), (2**31 - 1, 257)])
def test_randint_is_bit_exact_against_the_oracle(layout, N, B):
    got = br.randint(br.PRNGKey(42), (B,), 0, N, partitionable=bool(layout)).cpu().numpy()
    want = tf.randint(tf.prng_key(42), B, 0, N, layout)
    assert got.dtype == np.int32 and np.array_equal(got, want)
    assert got.min() >= 0 and got.max() < N


def test_randint_offsets_shapes_and_choice():
    k = br.PRNGKey(9)
    got = br.randint(k, (3, 7), -5, 12).cpu().numpy()
    assert got.shape == (3, 7) and np.array_equal(got.reshape(-1), tf.randint(tf.prng_key(9), 21, -5, 12))
    assert np.array_equal(br.choice(k, 1000, (64,)).cpu().numpy(), tf.randint(tf.prng_key(9), 64, 0, 1000))
    assert np.array_equal(br.randint(k, (4,), 3, 3).cpu().numpy(), np.full(4, 3, np.int32))  # empty range: minval everywhere
    assert br.randint(k, (0,), 0, 10).numel() == 0


def _logistic_model(D, prepare_dataset=True):
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    return bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits(), vi_type="mean_field", prepare_dataset=prepare_dataset)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("case", ["raw_small_d", "indexed_tc", "indexed_planes_tc", "indexed_planes_d200", "planes_tc", "coin"])
def test_gathered_batch_and_loss_match_indexing_on_the_host(case, layout):
    """sample(): idx / y[idx] / X[idx] (or the planes of X[idx]) / k_elbo against the oracle's split + randint; the
    minibatch loss and gradients equal loss_fn evaluated on the rows indexed by torch, bit for bit.  Copy-free routes:
    indexed_tc = the register-converting kernel reading raw fp32 X through the index (engine without prepared planes);
    indexed_planes_* = the TMA-fed kernel gathering rows of the FULL dataset's planes with 16-byte cp.async copies."""
    dev = torch.device("cuda")
    br.config.threefry_partitionable = bool(layout)
    try:
        if case == "raw_small_d":
            N, D, B, S = 5000, 5, 333, 6
            X, y = C.synth_linear(N, D)
            model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}, {}, lk.GaussianLinear())
            params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.05 * br.normal(k, s))
            params["log_noise_scale"] = torch.tensor(-0.7, device=dev)
        elif case in ("planes_tc", "indexed_tc", "indexed_planes_tc", "indexed_planes_d200"):
            N, D, B, S = 3000, {"indexed_planes_d200": 200, "indexed_planes_tc": 512}.get(case, 128), 700, 32
            X, y = C.synth_logistic(N, D)
            model = _logistic_model(D, prepare_dataset=case != "indexed_tc")
            params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.05 * br.normal(k, s))
        else:
            N, B, S = 777, 64, 16
            X, y = None, np.array([1.0 if n % 10 < 7 else 0.0 for n in range(N)], dtype=np.float32)
            model = bj.ADVI({"p_of_heads": tfd.Beta(0.5, 0.5)}, {"p_of_heads": tfb.Sigmoid()}, lk.BernoulliProbs("p_of_heads"))
            params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.05 * br.normal(k, s))
        yd = torch.tensor(y, device=dev)
        Xd = torch.tensor(X, device=dev) if X is not None else None
        inputs = {"X": Xd} if Xd is not None else None
        loss = model.minibatch_loss_fn(yd, inputs, batch_size=B, n_samples=S, full_data_size=3 * N, copy_free={"planes_tc": False, "indexed_tc": True, "indexed_planes_tc": True, "indexed_planes_d200": True}.get(case))
        seed = br.PRNGKey(77)
        idx, yb, Xb, ke = loss.sample(params, seed)
        kb_np, ke_np = tf.split(tf.prng_key(77), 2, layout)
        want_idx = tf.randint(kb_np, B, 0, N, layout)
        assert np.array_equal(idx.cpu().numpy(), want_idx)
        assert np.array_equal(ke.cpu().numpy(), ke_np)
        assert np.array_equal(yb.cpu().numpy(), y[want_idx])
        uses_planes = case == "planes_tc"
        st = loss._state
        assert (st["planes_b"] is not None) == uses_planes
        assert st["indexed"] == case.startswith("indexed")  # the fused tcgen05 kernel reads the dataset through idx: no copy
        if st["indexed"]:
            assert Xb is None and st["mb"].Xb is None and st["mb"].planes_b is None
            assert (st["planes_full"] is not None) == case.startswith("indexed_planes")
        if X is not None and not uses_planes and not st["indexed"]:
            assert np.array_equal(Xb.cpu().numpy(), X[want_idx])
        if uses_planes:  # gathered plane rows == the planes of the gathered rows
            Xg = Xd[torch.as_tensor(want_idx, device=dev).long()].contiguous()
            ref = torch.empty_like(st["planes_b"])
            nv.check(nv.load().bjx_prepare_x_planes(Xg.data_ptr(), B, D, D, ref.data_ptr(), nv.current_stream_ptr()))
            assert torch.equal(ref, st["planes_b"])
        # the loss: same numbers as loss_fn on rows indexed on the host side
        got, g_got = bj.value_and_grad(loss)(params, seed)
        li = torch.as_tensor(want_idx, device=dev).long()
        ref_inputs = {"X": Xd[li].contiguous()} if Xd is not None else None
        ke_t = torch.tensor(ke_np.astype(np.int64), dtype=torch.int64).to(torch.uint32).to(dev)
        # the same kernel without the index, on rows gathered by torch: raw fp32 rows converted in registers (indexed_tc), or
        # the TMA-fed kernel on the planes of the gathered rows (every other tensor-core case) -- same tiles, same order
        want, g_want = model.value_and_grad(params, yd[li].contiguous(), ref_inputs, 3 * N, ke_t, S)
        assert got.item() == want.item()
        for k in ("loc", "scale_diag"):
            assert torch.equal(g_got["posterior"][k], g_want["posterior"][k])
        if case == "indexed_tc":  # and the TMA-fed kernel on the gathered planes agrees to rounding
            w2, g2 = _logistic_model(D).value_and_grad(params, yd[li].contiguous(), ref_inputs, 3 * N, ke_t, S)
            C.assert_close(got.item(), w2.item(), what="loss, indexed vs planes")
            C.assert_close(g_got["posterior"]["loc"].cpu().numpy(), g2["posterior"]["loc"].cpu().numpy(), what="d loc, indexed vs planes")
    finally:
        br.config.threefry_partitionable = False


def test_minibatch_loss_against_the_float64_oracle():
    """End to end against the CPU restatement: split -> randint -> index -> ELBO and gradients."""
    N, D, B, S = 2000, 5, 256, 8
    X, y = C.synth_linear(N, D)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.GaussianLinear())
    loc = 0.1 * tf.normal(tf.prng_key(3), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32)
    dev = torch.device("cuda")
    params = {"posterior": {"loc": torch.tensor(loc, device=dev), "scale_diag": torch.tensor(raw, device=dev)},
              "log_noise_scale": torch.tensor(math.log(0.3), dtype=torch.float32, device=dev)}
    loss = model.minibatch_loss_fn(torch.tensor(y, device=dev), {"X": torch.tensor(X, device=dev)}, batch_size=B, n_samples=S)
    got, grads = bj.value_and_grad(loss)(params, br.PRNGKey(5))
    kb, ke = tf.split(tf.prng_key(5), 2, 0)
    idx = tf.randint(kb, B, 0, N, 0)
    (wl, wgl, wgr, wgk), _ = C.oracle_step(ref, loc, raw, {"log_noise_scale": np.float32(math.log(0.3))}, y[idx], X[idx], N, ke, S, 0)
    C.assert_close(got.item(), wl.item(), what="loss")
    C.assert_close(grads["posterior"]["loc"].cpu().numpy(), wgl.numpy(), what="d loc")
    C.assert_close(grads["posterior"]["scale_diag"].cpu().numpy(), wgr.numpy(), what="d raw scale")
    C.assert_close(grads["log_noise_scale"].cpu().numpy(), wgk["log_noise_scale"].numpy(), what="d log_noise_scale")


@pytest.mark.parametrize("case", ["raw_small_d", "indexed_tc", "indexed_planes_tc", "planes_tc"])
def test_device_resident_minibatch_loop_matches_the_host_driven_loop(case):
    """bjx_train_steps_minibatch in replayed CUDA graphs (gather -> step -> Adam per epoch, batch drawn from
    split(seed, n)[i] on the device) against the generic loop calling loss(params, seed=seeds[i])."""
    from bijax_b200 import utils as U

    dev = torch.device("cuda")
    if case == "raw_small_d":
        N, D, B, S = 20000, 5, 512, 8
        X, y = C.synth_linear(N, D, noise=0.1)
        model = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D), scale_diag=np.ones(D))}, {}, lk.GaussianLinear())
    else:
        N, D, B, S = 20000, (200 if case == "indexed_planes_tc" else 128), 1024, 32  # 200: only the planes kernel gathers it
        X, y = C.synth_logistic(N, D)
        model = _logistic_model(D)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)

    def fresh():
        p = model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
        if case == "raw_small_d":
            p["log_noise_scale"] = torch.tensor(-1.0, device=dev)
        return p

    loss = model.minibatch_loss_fn(yd, {"X": Xd}, batch_size=B, n_samples=S,
                                   copy_free={"planes_tc": False, "indexed_tc": True, "indexed_planes_tc": True}.get(case))
    n = 2 * U.GRAPH_STEPS + 3
    fused = bj.train(loss, fresh(), optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3))
    host = bj.train(loss, fresh(), optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3), fused=False)
    assert fused["losses"].shape == host["losses"].shape == (n,)
    assert torch.allclose(fused["losses"], host["losses"], rtol=2e-5, atol=1e-5)
    for k in ("loc", "scale_diag"):
        assert torch.allclose(fused["params"]["posterior"][k], host["params"]["posterior"][k], rtol=1e-4, atol=1e-5)
    if case == "indexed_planes_tc":
        assert loss._state["indexed"] and loss._state["planes_full"] is not None
    # different epochs see different batches
    assert fused["losses"].unique().numel() == n
    # the idx buffer holds the LAST epoch's draw
    keys = tf.split(tf.prng_key(3), n, 0)
    kb, _ = tf.split(keys[n - 1], 2, 0)
    assert np.array_equal(loss._state["idx"].cpu().numpy(), tf.randint(kb, B, 0, N, 0))


def test_minibatch_argument_checks():
    lib = nv.load()
    mb = nv.BjxMinibatch()
    k = br.PRNGKey(0)
    mb.N_full, mb.B = 0, 4
    assert lib.bjx_minibatch_gather(C_.byref(mb), k.data_ptr(), None, 0, None) < 0  # empty dataset
    mb.N_full, mb.B = 2**31, 4
    assert lib.bjx_minibatch_gather(C_.byref(mb), k.data_ptr(), None, 0, None) < 0  # int32 indices only
    mb.N_full, mb.B = 10, 4
    buf = torch.empty(4, device="cuda")
    mb.yb = buf.data_ptr()
    assert lib.bjx_minibatch_gather(C_.byref(mb), k.data_ptr(), None, 0, None) < 0  # yb without y_full
    assert b"y_full" in lib.bjx_last_error()
    with pytest.raises(ValueError):
        br.randint(k, (3,), 0, 2**31 + 5)


@pytest.mark.parametrize("D,S,N,B", [(512, 32, 30000, 4097), (256, 7, 9000, 1), (384, 20, 5000, 5000)])
def test_step_through_a_row_index_equals_the_step_on_gathered_rows(D, S, N, B):
    """BjxElboArgs.row_index (ABI 3): the fused tcgen05 kernel reading X[idx[r]] gives, bit for bit, what it gives on the
    gathered rows (repeated indices, B not a multiple of the 64-row tile, every block count); shapes outside that kernel
    say so instead of ignoring the index."""
    dev = torch.device("cuda")
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits(), prepare_dataset=False)
    eng, _ = model._engine_for({})
    assert eng.supports_row_index(S, B)
    idx = br.randint(br.PRNGKey(D + B), (B,), 0, N)
    li = idx.long()
    loc = 0.05 * br.normal(br.PRNGKey(1), (D,))
    raw = torch.full((D,), -2.0, device=dev)
    key = br.PRNGKey(4)
    want = eng.step(key, loc, raw, None, Xd[li].contiguous(), yd[li].contiguous(), N / B, S).clone()
    got = eng.step(key, loc, raw, None, Xd, yd[li].contiguous(), N / B, S, row_index=idx).clone()
    assert torch.equal(got, want)
    # a shape the fused kernel does not take
    small = bj.ADVI({"weights": tfd.MultivariateNormalDiag(loc=np.zeros(5, np.float32), scale_diag=np.ones(5, np.float32))},
                    {"weights": tfb.Identity()}, lk.BernoulliLogits())
    e2, _ = small._engine_for({})
    assert not e2.supports_row_index(S, B)
    X5 = torch.zeros((N, 5), device=dev)
    with pytest.raises(nv.BjxError):
        e2.step(key, torch.zeros(5, device=dev), torch.zeros(5, device=dev), None, X5, yd[li].contiguous(), 1.0, S, row_index=idx)
    with pytest.raises(ValueError):  # one index per batch row
        eng.step(key, loc, raw, None, Xd, yd[li].contiguous(), N / B, S, row_index=idx[:-1].contiguous() if B > 1 else idx.long())



@pytest.mark.parametrize("D,S,N,B", [(512, 32, 30000, 4097), (256, 7, 9000, 1), (384, 20, 5000, 5000), (128, 32, 20000, 9001),
                                     (200, 32, 7000, 3333), (64, 16, 4000, 700), (500, 32, 6000, 6000)])
def test_step_gathering_plane_rows_equals_the_step_on_gathered_rows(D, S, N, B):
    """BjxElboArgs.row_index + x_planes + N_full (ABI 4): the TMA-fed kernel gathering rows of the full dataset's planes with
    16-byte cp.async copies gives, bit for bit (D > 128: the same tile schedule), what it gives on the planes of the gathered rows -- repeated indices, B not a
    multiple of the 64-row tile, every block count (ring and spare-slot schedules), widths padded by TMA zero fill."""
    dev = torch.device("cuda")
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits())
    eng, _ = model._engine_for({})
    assert eng.supports_row_index(S, B, planes=True)
    lib = nv.load()
    full = torch.empty(int(lib.bjx_x_planes_bytes(N, D)), dtype=torch.uint8, device=dev)
    nv.check(lib.bjx_prepare_x_planes(Xd.data_ptr(), N, D, D, full.data_ptr(), nv.current_stream_ptr()))
    idx = br.randint(br.PRNGKey(D + B), (B,), 0, N)
    li = idx.long()
    loc = 0.05 * br.normal(br.PRNGKey(1), (D,))
    raw = torch.full((D,), -2.0, device=dev)
    key = br.PRNGKey(4)
    Xg, yg = Xd[li].contiguous(), yd[li].contiguous()
    want = eng.step(key, loc, raw, None, Xg, yg, N / B, S).clone()  # prepares the planes of the gathered rows
    assert eng._planes is not None and eng.kernel_choice(S, B) == "tcgen05_bf16_split"
    got = eng.step(key, loc, raw, None, Xd, yg, N / B, S, row_index=idx, planes=full, n_full=N).clone()
    if D > 128:
        assert torch.equal(got, want)
    else:
        # D <= 128: the contiguous step runs on 128-row tiles (k_lik_tc_tall), the gathering one on 64-row tiles -- the same
        # products summed in a different order
        C.assert_close(got.cpu().numpy(), want.cpu().numpy(), rtol=2e-6, atol_scale=2e-6, what="gathered vs contiguous")
    with pytest.raises(nv.BjxError):  # planes of the full dataset without its row count
        eng.step(key, loc, raw, None, Xd, yg, N / B, S, row_index=idx, planes=full, n_full=0)


def test_minibatch_on_the_two_pass_gemm_path():
    """A batch step outside the fused kernel's envelope (S = 128: four slices of 32 samples would cost more than two passes) runs the TMA-fed two-pass GEMM on gathered plane rows:
    same numbers (to the rounding of its atomic reduction order) as loss_fn on host-indexed rows."""
    dev = torch.device("cuda")
    N, D, B, S = 40000, 192, 16384, 128
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model = bj.ADVI(prior, {"weights": tfb.Identity()}, lk.BernoulliLogits())
    params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.05 * br.normal(k, s))
    loss = model.minibatch_loss_fn(yd, {"X": Xd}, batch_size=B, n_samples=S)
    seed = br.PRNGKey(31)
    got, g_got = bj.value_and_grad(loss)(params, seed)
    st = loss._state
    assert st["engine"].kernel_choice(S, B) == "tcgen05_gemm_two_pass"
    assert st["planes_b"] is not None and not st["indexed"]
    kb, ke = tf.split(tf.prng_key(31), 2, 0)
    li = torch.as_tensor(tf.randint(kb, B, 0, N, 0), device=dev).long()
    ke_t = torch.tensor(ke.astype(np.int64), dtype=torch.int64).to(torch.uint32).to(dev)
    want, g_want = model.value_and_grad(params, yd[li].contiguous(), {"X": Xd[li].contiguous()}, N, ke_t, S)
    C.assert_close(got.item(), want.item(), rtol=2e-6, atol_scale=2e-6, what="loss")
    for k in ("loc", "scale_diag"):
        C.assert_close(g_got["posterior"][k].cpu().numpy(), g_want["posterior"][k].cpu().numpy(), rtol=2e-6, atol_scale=2e-6, what=k)


def test_randint_against_the_committed_vectors():
    """tests/golden/randint_vectors.json: the CUDA sampler against the committed index draws (both layouts)."""
    import json
    from pathlib import Path

    fx = json.loads((Path(__file__).parent / "golden" / "randint_vectors.json").read_text())
    for c in fx["cases"]:
        got = br.randint(br.PRNGKey(c["seed"]), (c["n"],), c["minval"], c["maxval"], partitionable=bool(c["layout"]))
        assert got.cpu().tolist() == c["values"], c


@pytest.mark.parametrize("D", [128, 200])
def test_more_than_32_samples_through_the_index(D):
    """S = 64 on a narrow X runs the fused single pass once per slice of 32 samples; through a row index every slice gathers the
    batch's rows again (raw fp32 X at D = 128, the dataset's planes at D = 200).  Loss and gradients equal the step on rows gathered
    by torch bit for bit, and the float64 oracle to rtol 1e-5; the device-resident loop equals the host-driven one."""
    from bijax_b200 import utils as U

    dev = torch.device("cuda")
    N, B, S = 20000, 4097, 64
    X, y = C.synth_logistic(N, D)
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, ref = C.build_pair(prior, {}, lk.BernoulliLogits())
    params = model.init(br.PRNGKey(0), initializer=lambda k, s: 0.05 * br.normal(k, s))
    loss = model.minibatch_loss_fn(yd, {"X": Xd}, batch_size=B, n_samples=S, copy_free=True)
    seed = br.PRNGKey(31)
    got, g_got = bj.value_and_grad(loss)(params, seed)
    st = loss._state
    assert st["indexed"] and st["engine"].kernel_choice(S, B) == "tcgen05_bf16_split"
    kb, ke = tf.split(tf.prng_key(31), 2, 0)
    idx = tf.randint(kb, B, 0, N, 0)
    li = torch.as_tensor(idx, device=dev).long()
    ke_t = torch.tensor(ke.astype(np.int64), dtype=torch.int64).to(torch.uint32).to(dev)
    same_route = model if st["planes_full"] is not None else bj.ADVI(prior, {}, lk.BernoulliLogits(), prepare_dataset=False)
    want, g_want = same_route.value_and_grad(params, yd[li].contiguous(), {"X": Xd[li].contiguous()}, N, ke_t, S)
    assert got.item() == want.item() and torch.equal(g_got["posterior"]["loc"], g_want["posterior"]["loc"])
    loc, raw = params["posterior"]["loc"].cpu().numpy(), params["posterior"]["scale_diag"].cpu().numpy()
    (wl, wgl, wgr, _), _ = C.oracle_step(ref, loc, raw, {}, y[idx], X[idx], N, ke, S, 0)
    C.assert_close(got.item(), wl.item(), what="loss")
    C.assert_close(g_got["posterior"]["loc"].cpu().numpy(), wgl.numpy(), what="d loc")
    C.assert_close(g_got["posterior"]["scale_diag"].cpu().numpy(), wgr.numpy(), what="d raw scale")
    n = 2 * U.GRAPH_STEPS + 1
    fresh = lambda: model.init(br.PRNGKey(0), initializer=lambda k, shape: 0.01 * br.normal(k, shape))
    fused = bj.train(loss, fresh(), optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3))
    host = bj.train(loss, fresh(), optim.adam(0.02), n_epochs=n, seed=br.PRNGKey(3), fused=False)
    assert torch.allclose(fused["losses"], host["losses"], rtol=2e-5, atol=1e-5)

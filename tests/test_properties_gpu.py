"""GPU property tests (hypothesis): randomly drawn small problems against the float64 oracle,
row-shard and sample-shard additivity with arbitrary (also empty) shards, and layout invariance of the noise."""
import math

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from bijax_b200 import bijectors as tfb
from bijax_b200 import distributions as tfd
from bijax_b200 import likelihoods as lk
from bijax_b200 import random as br
from oracle import threefry as tf
from tests import cases as C
from tests.test_elbo_gpu import _run_case

pytestmark = pytest.mark.gpu
_SETTINGS = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))


@settings(max_examples=30, **_SETTINGS)
@given(N=st.integers(1, 700), D=st.integers(1, 70), S=st.integers(1, 40), gauss=st.booleans(), layout=st.integers(0, 1),
       softplus=st.booleans(), seed=st.integers(0, 2**31 - 1))
def test_random_small_problems_match_the_oracle(N, D, S, gauss, layout, softplus, seed):
    """Any (N, D, S) in the CUDA-core envelope (small-D register kernel for D <= 15, tiled kernel above), both likelihood
    families, both scale bijectors, both counter layouts: loss and gradients within rtol 1e-5 of the oracle."""
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.full(D, 1.5, np.float32))}
    sb = tfb.Softplus() if softplus else None
    if gauss:
        X, y = C.synth_linear(N, D)
        model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), scale_bijector=sb)
        kwargs = {"log_noise_scale": np.float32(math.log(0.4))}
    else:
        X, y = C.synth_logistic(N, D)
        model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), scale_bijector=sb)
        kwargs = {}
    loc = 0.2 * tf.normal(tf.prng_key(seed), (D,), 1)
    raw = -1.0 + 0.2 * tf.normal(tf.prng_key(seed + 1), (D,), 1)
    _run_case(model, ref, loc, raw, kwargs, y, X, 3 * N, seed % 1000, S, layout)


@settings(max_examples=25, **_SETTINGS)
@given(N=st.integers(1, 2500), D=st.integers(16, 1024), S=st.integers(1, 32), gauss=st.booleans(), prepared=st.booleans(),
       seed=st.integers(0, 2**31 - 1))
def test_random_widths_on_the_fused_tensor_core_kernels_match_the_oracle(N, D, S, gauss, prepared, seed):
    """Any width in [16, 1024] with S <= 32 runs ONE pass over X on the tensor cores: 128-row tiles of one or two blocks (D <= 64,
    D <= 128), 64-row tiles in a ring or with a spare slot (D <= 512), a cluster of two CTAs per tile (D <= 1024) -- ragged N,
    widths that are not a multiple of anything, prepared planes or the in-step split, both likelihood families."""
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    if gauss:
        X, y = C.synth_linear(N, D)
        model, ref = C.build_pair(prior, {}, lk.GaussianLinear(), precision="bf16x2", prepare_dataset=prepared)
        kwargs = {"log_noise_scale": np.float32(math.log(0.5))}
    else:
        X, y = C.synth_logistic(N, D)
        model, ref = C.build_pair(prior, {}, lk.BernoulliLogits(), precision="bf16x2", prepare_dataset=prepared)
        kwargs = {}
    loc = 0.05 * tf.normal(tf.prng_key(seed), (D,), 1)
    raw = np.full(D, math.log(0.1), np.float32)
    _run_case(model, ref, loc, raw, kwargs, y, X, 4 * N, seed % 1000, S, 0, expect_kernel="tcgen05_bf16_split")


@settings(max_examples=20, **_SETTINGS)
@given(N=st.integers(1, 3000), D=st.sampled_from([3, 5, 15, 40, 128]), S=st.integers(1, 32), G=st.integers(2, 8), data=st.data())
def test_row_shards_with_arbitrary_cuts_sum_to_the_unsharded_step(N, D, S, G, data):
    """contiguous row shards (any cut points, empty shards included) with ll_scale from the GLOBAL
    length and the q / prior terms on one shard only sum to the unsharded packed output."""
    cuts = sorted(data.draw(st.lists(st.integers(0, N), min_size=G - 1, max_size=G - 1)))
    bounds = [0] + cuts + [N]
    X, y = C.synth_logistic(N, D)
    dev = torch.device("cuda")
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.BernoulliLogits())
    eng, _ = model._engine_for({})
    loc = 0.1 * br.normal(br.PRNGKey(1), (D,))
    raw = torch.full((D,), -1.5, device=dev)
    key = br.PRNGKey(N * 31 + S)
    full = eng.step(key, loc, raw, None, Xd, yd, 2.5, S).clone()
    total = torch.zeros_like(full)
    for g in range(G):
        a, b = bounds[g], bounds[g + 1]
        total += eng.step(key, loc, raw, None, Xd[a:b], yd[a:b], 2.5, S, prior_weight=1.0 if g == 0 else 0.0)
    C.assert_close(total.cpu().numpy(), full.cpu().numpy(), rtol=5e-6, atol_scale=5e-6, what="row-shard additivity")


@settings(max_examples=20, **_SETTINGS)
@given(N=st.integers(1, 400), D=st.integers(1, 24), S=st.integers(2, 48), G=st.integers(2, 8), layout=st.integers(0, 1), data=st.data())
def test_sample_shards_with_arbitrary_cuts_sum_to_the_unsharded_step(N, D, S, G, layout, data):
    """SURVEY.md section 8e, small-N case: sample shards [s0, s1) (global eps indices, means over the global S, the
    per-step entropy constant on one shard) sum to the unsharded step; empty shards are skipped as the host layer does."""
    cuts = sorted(data.draw(st.lists(st.integers(0, S), min_size=G - 1, max_size=G - 1)))
    bounds = [0] + cuts + [S]
    X, y = C.synth_linear(N, D)
    dev = torch.device("cuda")
    Xd, yd = torch.tensor(X, device=dev), torch.tensor(y, device=dev)
    prior = {"weights": tfd.MultivariateNormalDiag(loc=np.zeros(D, np.float32), scale_diag=np.ones(D, np.float32))}
    model, _ = C.build_pair(prior, {}, lk.GaussianLinear())
    eng, _ = model._engine_for({"log_noise_scale": torch.tensor(0.0)})
    loc = 0.1 * br.normal(br.PRNGKey(2), (D,))
    raw = torch.full((D,), -1.0, device=dev)
    lns = torch.tensor([-0.5], device=dev)
    key = br.PRNGKey(S * 17 + D)
    part = bool(layout)
    full = eng.step(key, loc, raw, lns, Xd, yd, 1.5, S, partitionable=part).clone()
    total = torch.zeros_like(full)
    first = True
    for g in range(G):
        s0, s1 = bounds[g], bounds[g + 1]
        if s1 == s0:
            continue
        total += eng.step(key, loc, raw, lns, Xd, yd, 1.5, s1 - s0, partitionable=part, sample_shard=(s0, S, 1.0 if first else 0.0))
        first = False
    C.assert_close(total.cpu().numpy(), full.cpu().numpy(), rtol=5e-6, atol_scale=5e-6, what="sample-shard additivity")


@settings(max_examples=25, **_SETTINGS)
@given(S=st.integers(1, 300), D=st.integers(1, 300), layout=st.integers(0, 1), seed=st.integers(0, 2**63 - 1), data=st.data())
def test_noise_is_layout_invariant_and_sliceable(S, D, layout, seed, data):
    """eps of shape (S, D) is the flat stream of S * D draws; with the partitionable layout any slice regenerates from
    (offset, total); both equal the oracle bit for bit."""
    part = bool(layout)
    key = br.PRNGKey(seed)
    a = br.normal(key, (S, D), partitionable=part)
    b = br.normal(key, (S * D,), partitionable=part)
    assert torch.equal(a.reshape(-1), b)
    want = tf.normal(tf.prng_key(seed), (S * D,), layout)
    assert np.array_equal(b.cpu().numpy().view(np.uint32), want.view(np.uint32))
    off = data.draw(st.integers(0, S * D - 1))
    n = data.draw(st.integers(1, S * D - off))
    if part:
        c = br.normal(key, (n,), partitionable=True, offset=off, total=S * D)
        assert torch.equal(c, b[off : off + n])
    elif off > 0 or n < S * D:  # the legacy layout pairs counter j with j + ceil(total / 2): a slice is refused, not guessed
        from bijax_b200 import _native as nv

        with pytest.raises(nv.BjxError):
            br.normal(key, (n,), partitionable=False, offset=off, total=S * D)

"""CPU property tests (hypothesis) of the oracle itself: the restated PRNG is a flat, sliceable stream in both layouts, the
index draw stays in range, and the float64 ELBO restatement is additive over row shards -- the properties the GPU suite
then checks on the CUDA path (tests/test_properties_gpu.py)."""
import math

import numpy as np
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import advi_ref as R
from oracle import threefry as tf
from tests import cases as C

_SETTINGS = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))


@settings(max_examples=40, **_SETTINGS)
@given(S=st.integers(1, 60), D=st.integers(1, 60), layout=st.integers(0, 1), seed=st.integers(0, 2**63 - 1), data=st.data())
def test_oracle_noise_is_a_flat_sliceable_stream(S, D, layout, seed, data):
    key = tf.prng_key(seed)
    flat = tf.normal(key, (S * D,), layout)
    assert np.array_equal(tf.normal(key, (S, D), layout).reshape(-1).view(np.uint32), flat.view(np.uint32))
    bits = tf.random_bits(key, S * D, layout)
    off = data.draw(st.integers(0, S * D - 1))
    n = data.draw(st.integers(1, S * D - off))
    if layout == tf.LAYOUT_PARTITIONABLE:  # counter = global flat index: a slice needs no knowledge of the total
        assert np.array_equal(tf.random_bits(key, n, layout, offset=off), bits[off : off + n])
    assert np.isfinite(flat).all() and np.abs(flat).max() < 6.0


@settings(max_examples=40, **_SETTINGS)
@given(n=st.integers(1, 200), lo=st.integers(-1000, 1000), span=st.integers(1, 2**31 - 1001), layout=st.integers(0, 1),
       seed=st.integers(0, 2**63 - 1))
def test_oracle_randint_stays_in_range(n, lo, span, layout, seed):
    got = tf.randint(tf.prng_key(seed), n, lo, lo + span, layout)
    assert got.dtype == np.int32 and got.shape == (n,)
    assert got.min() >= lo and got.max() < lo + span


@settings(max_examples=15, **_SETTINGS)
@given(N=st.integers(2, 200), D=st.integers(1, 12), S=st.integers(1, 9), G=st.integers(2, 5), data=st.data())
def test_oracle_elbo_is_additive_over_row_shards(N, D, S, G, data):
    """loss / grads of the float64 restatement: sum over shards of [weight_g * (q - p) - (full / N) * ll_g] == unsharded."""
    cuts = sorted(data.draw(st.lists(st.integers(0, N), min_size=G - 1, max_size=G - 1)))
    bounds = [0] + cuts + [N]
    X, y = C.synth_logistic(N, D)
    ref = R.ADVIRef({"weights": R.MultivariateNormalDiag(torch.zeros(D, dtype=torch.float64), torch.ones(D, dtype=torch.float64))},
                    {"weights": R.Identity()}, R.bernoulli_logits_loglik("weights"), "mean_field")
    loc = torch.tensor(0.1 * tf.normal(tf.prng_key(1), (D,), 1), dtype=torch.float64)
    raw = torch.full((D,), math.log(0.3), dtype=torch.float64)
    eps = torch.tensor(tf.normal(tf.prng_key(2), (S, D), 0), dtype=torch.float64)
    t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    full_size = 2.0 * N
    wl, wgl, wgr, _ = ref.value_and_grad(loc, raw, {}, t64(y), {"X": t64(X)}, full_size, eps)
    # the q / prior part alone (what the other shards must NOT add again): the loss is affine in full_data_size
    w2 = ref.value_and_grad(loc, raw, {}, t64(y), {"X": t64(X)}, 2.0 * full_size, eps)
    pl, pgl, pgr = 2.0 * wl - w2[0], 2.0 * wgl - w2[1], 2.0 * wgr - w2[2]
    tl, tgl, tgr = pl.clone(), pgl.clone(), pgr.clone()
    for g in range(G):
        a, b = bounds[g], bounds[g + 1]
        if a == b:
            continue
        # a shard sees ll_scale = full / N_global: pass full_data_size scaled to its own length
        sl, sgl, sgr, _ = ref.value_and_grad(loc, raw, {}, t64(y[a:b]), {"X": t64(X[a:b])}, full_size * (b - a) / N, eps)
        tl, tgl, tgr = tl + (sl - pl), tgl + (sgl - pgl), tgr + (sgr - pgr)
    assert torch.allclose(tl, wl, rtol=1e-11, atol=1e-11)
    assert torch.allclose(tgl, wgl, rtol=1e-10, atol=1e-10) and torch.allclose(tgr, wgr, rtol=1e-10, atol=1e-10)

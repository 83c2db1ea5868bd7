"""CPU: host-side model compilation (flat layout, bijector chains, prior table) -- core.py:107-121,168-176."""
import math

import numpy as np
import pytest

import bijax_b200 as bj
from bijax_b200 import _native as nv
from bijax_b200 import bijectors as tfb
from bijax_b200 import distributions as tfd
from bijax_b200 import likelihoods as lk
from bijax_b200.spec import bijector_ops, compile_latents, fill_in_bijector, pack_chain


def test_sorted_key_row_major_layout():
    prior = {
        "weights": tfd.MultivariateNormalDiag(loc=np.zeros(3), scale_diag=np.array([1.0, 2.0, 3.0])),
        "bias": tfd.Normal(0.5, 2.0),
        "net": {"b": tfd.Gamma(2.0, 3.0), "a": tfd.Normal(np.zeros((2, 2)), 1.0)},
    }
    bij = fill_in_bijector({"net": {"b": tfb.Exp()}}, prior)
    spec = compile_latents(prior, bij)
    assert ["/".join(p) for p in spec.paths] == ["bias", "net/a", "net/b", "weights"]
    assert spec.offsets == [0, 1, 5, 6] and spec.size == 9 and spec.shapes[1] == (2, 2)
    c = spec.coords
    assert c["prior"].tolist() == [0, 0, 0, 0, 0, 2, 0, 0, 0]
    assert c["chain"][5] == nv.BIJ_EXP and c["chain"][0] == 0
    assert np.allclose(c["p1"][6:], [1, 2, 3])
    assert math.isclose(c["lognorm"][5], 2.0 * math.log(3.0) - math.lgamma(2.0), rel_tol=1e-6)
    assert math.isclose(c["lognorm"][0], -0.5 * math.log(2 * math.pi) - math.log(2.0), rel_tol=1e-6)
    assert spec.offset_of("weights") == (6, 3) and spec.offset_of(("net", "b")) == (5, 1)


def test_chain_order_is_right_to_left():
    ch = tfb.Chain([tfb.Exp(), tfb.Chain([tfb.Identity(), tfb.Softplus()]), tfb.Sigmoid()])
    assert bijector_ops(ch) == [nv.BIJ_SIGMOID, nv.BIJ_SOFTPLUS, nv.BIJ_EXP]
    assert pack_chain(bijector_ops(ch)) == nv.BIJ_SIGMOID | nv.BIJ_SOFTPLUS << 4 | nv.BIJ_EXP << 8


def test_fill_in_bijector_mutates_like_the_reference():
    user = {}
    prior = {"a": tfd.Normal(0.0, 1.0)}
    out = fill_in_bijector(user, prior)
    assert out is user and type(user["a"]).__name__ == "Identity"


def test_constructor_forms_and_errors():
    prior = {"p": tfd.Beta(0.5, 0.5)}
    m1 = bj.ADVI(prior, {"p": tfb.Sigmoid()}, lk.BernoulliProbs("p"), vi_type="mean_field")  # v1 (README.md:113)
    m2 = bj.ADVI(prior, lk.BernoulliProbs("p"), {"p": tfb.Sigmoid()}, "full_rank")  # v2 (advi.py:26-34)
    assert m1.size == m2.size == 1 and m2.vi_type == "full_rank"
    with pytest.raises(AssertionError):
        bj.ADVI(prior, {"p": tfb.Sigmoid()}, lk.BernoulliProbs("p"), vi_type="mean_field", rank=2)
    with pytest.raises(NotImplementedError):
        bj.ADVI(prior, {"p": tfb.Sigmoid()}, lambda *a, **k: 0.0)
    m3 = bj.ADVI(prior, {"p": tfb.Sigmoid()}, lk.BernoulliProbs("p"), vi_type="low_rank", rank=2)  # core.py:190-201
    assert m3.vi_type == "low_rank" and m3.rank == 2 and len(m3.posterior_params_bijector) == 3
    with pytest.raises(AssertionError):
        bj.ADVI(prior, {"p": tfb.Sigmoid()}, lk.BernoulliProbs("p"), vi_type="low_rank")  # rank is required (advi.py:55-57)
    with pytest.raises(ValueError):
        bj.ADVI(prior, {"p": tfb.Sigmoid()}, lk.BernoulliProbs("p"), vi_type="banded")


def test_duck_typed_foreign_distribution_objects():
    class Normal:  # e.g. a tfd.Normal where TFP exists
        def __init__(self):
            self.parameters = {"loc": 1.0, "scale": 4.0}
            self.loc, self.scale = None, None

    spec = compile_latents({"x": Normal()}, {"x": tfb.Identity()})
    assert spec.coords["p0"][0] == 1.0 and spec.coords["p1"][0] == 4.0

"""CPU: host-side logic of the train_fn mirror (bijax/utils.py:6-36) that needs no device -- which calls are recognised as
the device-resident loop, the generic loop's contract (losses, params_list, states, seeds), and the pytree helpers."""
from functools import partial

import pytest
import torch

from bijax_b200 import optim
from bijax_b200 import utils as U
from bijax_b200.tree import ravel_pytree, tree_flatten, tree_map, tree_unflatten


class _FakeModel:
    """Has the attributes _fused_loop_spec looks at; never touches CUDA."""

    _group = None
    packed_tril = False

    def _engine_for(self, kwargs):  # pragma: no cover - presence only
        raise AssertionError("not called")

    def loss_fn(self, params, outputs, inputs, full_data_size, seed, n_samples=1):  # pragma: no cover - presence only
        raise AssertionError("not called")


def test_which_calls_run_as_the_device_resident_loop():
    m = _FakeModel()
    adam = optim.adam(0.1)
    canonical = partial(m.loss_fn, outputs=[1.0], inputs=None, full_data_size=1, n_samples=4)
    spec = U._fused_loop_spec(canonical, adam, seed=object(), return_args=frozenset())
    assert spec is not None and spec[0] is m and spec[1]["n_samples"] == 4
    assert U._fused_loop_spec(canonical, adam, seed=None, return_args=frozenset()) is None          # no per-step keys
    assert U._fused_loop_spec(canonical, adam, seed=object(), return_args={"params_list"}) is None    # per-step output asked back
    assert U._fused_loop_spec(canonical, optim.sgd(0.1), seed=object(), return_args=frozenset()) is None  # not adam
    assert U._fused_loop_spec(m.loss_fn, adam, seed=object(), return_args=frozenset()) is None       # not a partial
    assert U._fused_loop_spec(partial(m.loss_fn, [1.0]), adam, seed=object(), return_args=frozenset()) is None  # positional
    extra = partial(m.loss_fn, outputs=[1.0], inputs=None, full_data_size=1, seed=3)
    assert U._fused_loop_spec(extra, adam, seed=object(), return_args=frozenset()) is None            # seed fixed by the caller
    missing = partial(m.loss_fn, outputs=[1.0], inputs=None)
    assert U._fused_loop_spec(missing, adam, seed=object(), return_args=frozenset()) is None
    other = partial(lambda params, **kw: 0.0, outputs=[1.0], inputs=None, full_data_size=1)
    assert U._fused_loop_spec(other, adam, seed=object(), return_args=frozenset()) is None            # not a model's loss_fn
    m2 = _FakeModel()
    m2._group = object()  # row-sharded: host-driven step + all-reduce
    assert U._fused_loop_spec(partial(m2.loss_fn, outputs=[1.0], inputs=None, full_data_size=1), adam, object(), frozenset()) is None
    m2b = _FakeModel()  # row-sharded WITH the peer-memory exchange: the all-reduce runs on the device too (bjx_train_steps_sharded)
    m2b._group, m2b._shard, m2b._p2p = object(), "rows", True
    assert U._fused_loop_spec(partial(m2b.loss_fn, outputs=[1.0], inputs=None, full_data_size=1), adam, object(), frozenset()) is not None
    m2c = _FakeModel()  # sample sharding stays host-driven
    m2c._group, m2c._shard, m2c._p2p = object(), "samples", True
    assert U._fused_loop_spec(partial(m2c.loss_fn, outputs=[1.0], inputs=None, full_data_size=1), adam, object(), frozenset()) is None
    m3 = _FakeModel()
    m3.packed_tril = True
    assert U._fused_loop_spec(partial(m3.loss_fn, outputs=[1.0], inputs=None, full_data_size=1), adam, object(), frozenset()) is None


def test_generic_loop_contract_on_cpu_tensors():
    """utils.py:6-36 with a seedless loss: losses [n_epochs], params_list / states on request, params pytree preserved."""
    target = torch.tensor([3.0, -1.0])

    def loss_fn(params):
        return ((params["a"]["w"] - target) ** 2).sum() + (params["b"] - 0.5) ** 2

    params = {"a": {"w": torch.zeros(2)}, "b": torch.tensor(0.0)}
    res = U.train_fn(loss_fn, params, optim.sgd(0.1), n_epochs=60, return_args={"params_list", "states", "seeds"})
    assert set(res) == {"params", "losses", "params_list", "states", "seeds"}
    assert res["losses"].shape == (60,) and res["losses"][-1] < 1e-8 and res["losses"][0] == pytest.approx(10.25)
    assert torch.allclose(res["params"]["a"]["w"], target, atol=1e-4) and abs(res["params"]["b"].item() - 0.5) < 1e-4
    assert len(res["params_list"]) == 60 and res["seeds"] is None and res["states"] == ()
    assert torch.equal(res["params_list"][-1]["a"]["w"], res["params"]["a"]["w"])
    assert params["a"]["w"].abs().sum() == 0  # the caller's pytree is not modified in place
    empty = U.train_fn(loss_fn, params, optim.sgd(0.1), n_epochs=0)
    assert empty["losses"].numel() == 0 and empty["params"] is params


def test_value_and_grad_and_pytree_helpers():
    params = {"z": torch.tensor([1.0, 2.0]), "a": {"k": torch.tensor(3.0), "unused": torch.tensor([5.0])}}
    loss, grads = U.value_and_grad(lambda p: (p["z"] ** 2).sum() * p["a"]["k"])(params)
    assert loss.item() == 15.0
    assert torch.equal(grads["z"], torch.tensor([6.0, 12.0])) and grads["a"]["k"].item() == 5.0
    assert torch.equal(grads["a"]["unused"], torch.zeros(1))  # jax gives zeros for unused leaves, not None
    leaves, td = tree_flatten(params)
    assert [tuple(l.shape) for l in leaves] == [(), (1,), (2,)]  # sorted keys, depth first ("a" before "z")
    assert tree_unflatten(td, leaves)["a"]["k"] is params["a"]["k"]
    flat, unravel = ravel_pytree(params)
    assert flat.tolist() == [3.0, 5.0, 1.0, 2.0]
    back = unravel(flat * 2)
    assert back["z"].tolist() == [2.0, 4.0] and back["a"]["k"].shape == ()
    doubled = tree_map(lambda a, b: a + b, params, params)
    assert doubled["a"]["unused"].item() == 10.0


def test_adam_has_no_cpu_path_and_schedules_stay_on_the_host_loop():
    """optim.adam updates on the device only (no CPU fallback on the product path); an optax-style schedule is accepted and
    keeps train_fn on the generic loop (no static hyper-parameters to hand to bjx_train_steps)."""
    adam = optim.adam(0.1)
    state = adam.init({"w": torch.zeros(3)})
    with pytest.raises(TypeError, match="no CPU path"):
        adam.update({"w": torch.ones(3)}, state)
    sched = optim.adam(lambda count: 0.1 / (1 + count))
    assert sched.hyper is None and adam.hyper == (0.1, 0.9, 0.999, 1e-8)
    m = _FakeModel()
    canonical = partial(m.loss_fn, outputs=[1.0], inputs=None, full_data_size=1, n_samples=4)
    assert U._fused_loop_spec(canonical, sched, seed=object(), return_args=frozenset()) is None


def test_return_args_cover_the_reference_locals():
    """utils.py:33-35 returns ``locals()[key]`` for any requested name: ``state`` is the optimiser state before the scan."""
    res = U.train_fn(lambda p: (p["w"] ** 2).sum(), {"w": torch.ones(2)}, optim.sgd(0.1), n_epochs=3,
                     return_args={"state", "states", "n_epochs", "losses"})
    assert res["state"] == () and res["states"] == () and res["n_epochs"] == 3 and res["losses"].shape == (3,)

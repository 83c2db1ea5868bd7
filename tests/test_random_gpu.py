"""GPU: the threefry kernel against the CPU oracle -- bit-exact bits, uniforms and normals, both counter layouts."""
import numpy as np
import pytest
import torch

from oracle import threefry as tf

pytestmark = pytest.mark.gpu


def _key(seed):
    from bijax_b200 import random as br

    return br.PRNGKey(seed), tf.prng_key(seed)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 7, 16, 1000, 65537, 1 << 20])
def test_bits_uniform_normal_bit_exact(layout, n):
    from bijax_b200 import random as br

    k, kn = _key(20221019 + n)
    part = bool(layout)
    got_bits = br.bits(k, (n,), partitionable=part).cpu().numpy()
    assert np.array_equal(got_bits, tf.random_bits(kn, n, layout))
    got_u = br.uniform(k, (n,), partitionable=part).cpu().numpy()
    assert np.array_equal(got_u.view(np.uint32), tf.uniform(kn, (n,), layout).view(np.uint32))
    got_n = br.normal(k, (n,), partitionable=part).cpu().numpy()
    want_n = tf.normal(kn, (n,), layout)
    assert np.array_equal(got_n.view(np.uint32), want_n.view(np.uint32)), f"{(got_n != want_n).sum()} of {n} normals differ"


def test_known_answers_on_device():
    from bijax_b200 import random as br

    k0 = br.PRNGKey(0)
    assert br.split(k0, 2, partitionable=False).cpu().numpy().tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert br.normal(k0, (1,), partitionable=False).item() == np.float32(-0.20584226)
    assert br.normal(br.PRNGKey(42), (), partitionable=False).item() == np.float32(-0.18471177)
    assert br.normal(br.PRNGKey(42), (), partitionable=True).item() == np.float32(-0.028304616)
    assert br.uniform(k0, (), partitionable=False).item() == np.float32(0.41845703)
    assert br.normal(k0, (1,), partitionable=True).item() == np.float32(1.6226422)  # exact under the XLA Log1p / Cephes logf spec


def test_values_published_in_the_jax_documentation_on_device():
    """tests/golden/threefry_kat.json "jax_docs" on the CUDA generator: split chains, normals of subkeys, a (3,) draw, a uniform."""
    import json
    from pathlib import Path

    import torch

    from bijax_b200 import random as br

    kat = json.loads((Path(__file__).parent / "golden" / "threefry_kat.json").read_text())["jax_docs"]
    for name, part in (("legacy", False), ("partitionable", True)):
        for v in kat[name]:
            if "seed" in v:
                key = br.PRNGKey(v["seed"])
            else:
                key = torch.tensor(np.array(v["key"], dtype=np.int64), dtype=torch.int64).to(torch.uint32).cuda()
            if "split" in v:
                assert br.split(key, len(v["split"]), partitionable=part).cpu().numpy().tolist() == v["split"], v["what"]
            if "normal" in v:
                got = br.normal(key, tuple(v["shape"]), partitionable=part).cpu().numpy().ravel()
                assert np.array_equal(got, np.array(v["normal"], dtype=np.float32)), (v["what"], got)
            if "uniform" in v:
                got = br.uniform(key, tuple(v["shape"]), partitionable=part).cpu().numpy().ravel()
                assert np.array_equal(got, np.array(v["uniform"], dtype=np.float32)), (v["what"], got)


@pytest.mark.parametrize("layout", [0, 1])
def test_split_matches_oracle(layout):
    from bijax_b200 import random as br

    k, kn = _key(99)
    for num in (1, 3, 1000):
        got = br.split(k, num, partitionable=bool(layout)).cpu().numpy()
        assert np.array_equal(got, tf.split(kn, num, layout))


def test_partitionable_slices_are_shard_stable():
    from bijax_b200 import random as br

    k, kn = _key(5)
    total = 100000
    full = br.normal(k, (total,), partitionable=True).cpu().numpy()
    part = br.normal(k, (777,), partitionable=True, offset=12345, total=total).cpu().numpy()
    assert np.array_equal(full[12345 : 12345 + 777], part)
    # counters above 2^32 (the synthetic 5e9-element design matrix relies on them)
    off = (1 << 33) + 17
    big = br.bits(k, (64,), partitionable=True, offset=off, total=off + 64).cpu().numpy()
    assert np.array_equal(big, tf.random_bits(kn, 64, tf.LAYOUT_PARTITIONABLE, offset=off))


def test_shape_does_not_change_the_stream_and_legacy_rejects_offsets():
    from bijax_b200 import _native as nv
    from bijax_b200 import random as br

    k, _ = _key(11)
    a = br.normal(k, (32, 16)).cpu().numpy().ravel()
    b = br.normal(k, (512,)).cpu().numpy()
    assert np.array_equal(a, b)
    with pytest.raises(nv.BjxError):
        br.normal(k, (10,), partitionable=False, offset=5, total=100)

"""Writes tests/golden/randint_vectors.json: index draws of oracle/threefry.py::randint (= jax.random.randint / choice with
the defaults) for both counter layouts.  These are REGRESSION vectors of the restatement (its parity against a live JAX is
unpinned, see the function's docstring): they freeze what the oracle and the CUDA sampler agreed on when committed."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import threefry as tf

cases = []
for layout in (0, 1):
    for seed, n, lo, hi in ((0, 8, 0, 10), (42, 8, 0, 100), (7, 6, -5, 12), (1, 6, 0, 65536), (2, 6, 0, 65537), (3, 6, 0, 10_000_000), (4, 4, 0, 2**31 - 1)):
        cases.append({"layout": layout, "seed": seed, "n": n, "minval": lo, "maxval": hi,
                      "values": [int(v) for v in tf.randint(tf.prng_key(seed), n, lo, hi, layout)]})
out = Path(__file__).resolve().parent.parent / "tests" / "golden" / "randint_vectors.json"
out.write_text(json.dumps({"generator": "scripts/make_randint_fixture.py", "status": "regression vectors of the restatement; unpinned against JAX", "cases": cases}, indent=1) + "\n")
print(out, len(cases))

"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in NumPy, of the counter-based PRNG that sits under
``jax.random`` -- the generator bijax reaches through
``q.sample(seed=...)`` (/root/reference/bijax/advi.py:83), ``initialize_params``
(/root/reference/bijax/utils.py:53-55, default initialiser advi.py:77) and
``jax.random.split`` (/root/reference/bijax/utils.py:30,70).

JAX / jaxlib are third-party, unpinned dependencies of the reference
(/root/reference/requirements.txt:2-3) and are absent from this image, so
this file restates their *published* algorithm:

  * Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as
    easy as 1, 2, 3", SC'11; the Random123 known-answer vectors pin it);
  * ``jax._src.prng`` counter layouts: the legacy one (default before JAX
    0.5: iota split in halves) and the "partitionable" one (64-bit flat index
    split hi/lo, outputs XOR-ed);
  * ``jax.random.uniform`` bit trick (mantissa fill, minus one) and
    ``jax.random.normal`` = sqrt(2) * erf_inv(uniform(-1, 1)), with XLA's
    single-precision ``ErfInv`` (Giles' polynomial).

Pinned spec of the float pipeline ("the spec" in DESIGN.md): every operation
is IEEE-754 binary32 round-to-nearest with NO fused multiply-add, and
``w = -log1p(-u*u)`` follows XLA's own single-precision ``Log1p`` expansion
(xla/service/elemental_ir_emitter.cc, EmitLog1p -- what the reference's
jit-compiled CPU path executes): for |x| < sqrt(2) - 1 the Cephes rational
``x + (-x^2/2 + x*x^2 * P(x)/Q(x))`` evaluated in binary32 (Horner, separate
multiply and add, one division), otherwise ``log(1 + x)`` of the ROUNDED
binary32 sum, with ``log`` the Cephes ``logf`` polynomial that XLA:CPU inlines
for vectorised f32 logarithms (xla/service/cpu/llvm_ir_runtime.cc,
GenerateVF32Log: mantissa in [sqrt(1/2), sqrt(2)), degree-8 polynomial,
``e * ln 2`` added in two parts) -- again binary32 with separate multiply and
add.  Nothing in the pipeline calls a libm, so NumPy and the CUDA kernel agree
bit for bit by construction.  (Both expansions are restated from memory of
XLA's sources, which are not on this box; what pins them is below.)
``LOG1P_SPEC`` selects two alternatives kept
for the record: "correctly_rounded" (``float32(-log1p(float64(u*u)))``, the
round-1 spec, which misses one published value by one ulp) and "libm_f32"
(binary32 ``log1p`` of the platform's libm, what XLA:GPU's libdevice call
amounts to).

PARITY STATUS: pinned against the Random123 KATs and the publicly documented
``jax.random`` outputs listed in tests/golden/threefry_kat.json -- bits, split,
uniform and ALL FOUR documented normals exactly under the default spec; NOT
pinned against a live JAX (none available here).  The published partitionable
value for seed 0 is a sharp discriminator only by luck: log(1 - u*u) there sits
0.004 ulp from a binary32 rounding tie, so the correctly rounded logarithm and
the Cephes polynomial (0.79 ulp worst case) land on different floats, and the
published value is the polynomial's.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_MASK = np.uint64(0xFFFFFFFF)

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))

LAYOUT_LEGACY = 0
LAYOUT_PARTITIONABLE = 1


def _rotl(x, r):
    return ((x << U32(r)) | (x >> U32(32 - r))).astype(U32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 block function on uint32 arrays (broadcasting)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, (k0 ^ k1 ^ U32(0x1BD11BDA)).astype(U32))
        x0 = (x0 + ks[0]).astype(U32)
        x1 = (x1 + ks[1]).astype(U32)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(U32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(g + 2) % 3] + U32(g + 1)).astype(U32)
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` -> uint32[2] = [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def random_bits(key, n: int, layout: int = LAYOUT_LEGACY, offset: int = 0) -> np.ndarray:
    """n 32-bit draws for flat indices [offset, offset+n).

    ``offset`` is only meaningful for the partitionable layout (counter =
    global flat index, so any slice of a stream can be regenerated alone).
    """
    key = np.asarray(key, dtype=U32)
    n = int(n)
    if n == 0:
        return np.zeros((0,), dtype=U32)
    if layout == LAYOUT_LEGACY:
        if offset:
            raise ValueError("legacy layout is not prefix/shard stable; offset must be 0")
        h = (n + 1) // 2
        j = np.arange(h, dtype=np.uint64)
        c0 = (j & _MASK).astype(U32)
        c1 = ((j + np.uint64(h)) & _MASK).astype(U32)
        if n % 2:
            c1[-1] = 0  # the zero pad appended to the odd-sized iota
        a, b = threefry2x32(key[0], key[1], c0, c1)
        return np.concatenate([a, b])[:n]
    if layout == LAYOUT_PARTITIONABLE:
        i = np.arange(n, dtype=np.uint64) + np.uint64(offset)
        hi = (i >> np.uint64(32)).astype(U32)
        lo = (i & _MASK).astype(U32)
        a, b = threefry2x32(key[0], key[1], hi, lo)
        return (a ^ b).astype(U32)
    raise ValueError(f"unknown layout {layout}")


def split(key, num: int, layout: int = LAYOUT_LEGACY) -> np.ndarray:
    """``jax.random.split(key, num)`` -> uint32[num, 2]."""
    key = np.asarray(key, dtype=U32)
    if layout == LAYOUT_LEGACY:
        return random_bits(key, 2 * num, LAYOUT_LEGACY).reshape(num, 2)
    i = np.arange(num, dtype=np.uint64)
    a, b = threefry2x32(key[0], key[1], (i >> np.uint64(32)).astype(U32), (i & _MASK).astype(U32))
    return np.stack([a, b], axis=-1)


def bits_to_unit_float(bits: np.ndarray) -> np.ndarray:
    """Mantissa fill: float32 in [0, 1) from the top 23 bits."""
    bits = np.asarray(bits, dtype=U32)
    f = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32)
    return (f - np.float32(1.0)).astype(np.float32)


def uniform_from_bits(bits, minval=0.0, maxval=1.0) -> np.ndarray:
    """``jax.random.uniform`` float32: f*(maxval-minval)+minval, clamped below."""
    f = bits_to_unit_float(bits)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    scale = np.float32(hi - lo)
    u = (f * scale).astype(np.float32)
    u = (u + lo).astype(np.float32)
    return np.maximum(lo, u).astype(np.float32)


_ERFINV_SMALL = [
    2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
    -0.00125372503, -0.00417768164, 0.246640727, 1.50140941,
]
_ERFINV_LARGE = [
    -0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
    -0.0076224613, 0.00943887047, 1.00167406, 2.83297682,
]


# Cephes log1p rational (unity.c), the constants XLA's EmitLog1p carries: log1p(x) = x - x^2/2 + x^3 P(x)/Q(x)
_LOG1P_P = [4.5270000862445199635215e-5, 4.9854102823193375972212e-1, 6.5787325942061044846969e0,
            2.9911919328553073277375e1, 6.0949667980987787057556e1, 5.7112963590585538103336e1, 2.0039553499201281259648e1]
_LOG1P_Q = [1.0, 1.5062909083469192043167e1, 8.3047565967967209469434e1, 2.2176239823732856465394e2,
            3.0909872225312059774938e2, 2.1642788614495947685003e2, 6.0118660497603843919306e1]
_LOG1P_SMALL = np.float32(0.41421356237309504880)  # sqrt(2) - 1
LOG1P_SPEC = "xla_cpu"  # "xla_cpu" (default) | "correctly_rounded" | "libm_f32"


_LOGF_P = [7.0376836292e-2, -1.1514610310e-1, 1.1676998740e-1, -1.2420140846e-1, 1.4249322787e-1, -1.6668057665e-1,
           2.0000714765e-1, -2.4999993993e-1, 3.3333331174e-1]
_LOGF_Q1, _LOGF_Q2 = np.float32(-2.12194440e-4), np.float32(0.693359375)  # ln 2 = q1 + q2
_SQRTHF = np.float32(0.707106781186547524)


def logf_cephes(x: np.ndarray) -> np.ndarray:
    """Cephes logf as XLA:CPU inlines it, for positive normal binary32 x (0 -> -inf): all operations binary32, no FMA."""
    x = np.asarray(x, dtype=np.float32)
    bits = x.view(np.uint32)
    e = (((bits >> U32(23)) & U32(0xFF)).astype(np.int32) - 126).astype(np.float32)
    m = ((bits & U32(0x007FFFFF)) | U32(0x3F000000)).view(np.float32)  # mantissa in [0.5, 1)
    lt = m < _SQRTHF
    e = np.where(lt, e - np.float32(1.0), e).astype(np.float32)
    m = (np.where(lt, (m + m).astype(np.float32), m) - np.float32(1.0)).astype(np.float32)
    z = (m * m).astype(np.float32)
    y = _horner_f32(m, _LOGF_P)
    y = ((y * m).astype(np.float32) * z).astype(np.float32)
    y = (y + (e * _LOGF_Q1).astype(np.float32)).astype(np.float32)
    y = (y - (np.float32(0.5) * z).astype(np.float32)).astype(np.float32)
    r = (m + y).astype(np.float32)
    r = (r + (e * _LOGF_Q2).astype(np.float32)).astype(np.float32)
    return np.where(x == np.float32(0.0), np.float32(-np.inf), r).astype(np.float32)


def _horner_f32(x: np.ndarray, coeffs) -> np.ndarray:
    """XLA's EvaluatePolynomial in binary32: poly = poly * x + c, multiply and add rounded separately."""
    p = np.full_like(x, np.float32(coeffs[0]))
    for c in coeffs[1:]:
        p = (p * x).astype(np.float32)
        p = (p + np.float32(c)).astype(np.float32)
    return p


def log1p_f32(x: np.ndarray, spec: str = None) -> np.ndarray:
    """Single-precision log1p under the selected spec (module docstring)."""
    spec = LOG1P_SPEC if spec is None else spec
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        if spec == "correctly_rounded":
            return np.log1p(x.astype(np.float64)).astype(np.float32)
        if spec == "libm_f32":
            return np.log1p(x).astype(np.float32)
        if spec != "xla_cpu":
            raise ValueError(f"unknown log1p spec {spec!r}")
        big = logf_cephes((np.float32(1.0) + x).astype(np.float32))
        num, den = _horner_f32(x, _LOG1P_P), _horner_f32(x, _LOG1P_Q)
        x2 = (x * x).astype(np.float32)
        r = (num / den).astype(np.float32)
        q = ((x * x2).astype(np.float32) * r).astype(np.float32)
        z = ((np.float32(-0.5) * x2).astype(np.float32) + q).astype(np.float32)
        small = (x + z).astype(np.float32)
        return np.where(np.abs(x) < _LOG1P_SMALL, small, big).astype(np.float32)


def erfinv_f32(u: np.ndarray, spec: str = None) -> np.ndarray:
    """XLA ErfInv (f32, Giles) under the pinned spec in the module docstring."""
    u = np.asarray(u, dtype=np.float32)
    t = (u * u).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-log1p_f32(-t, spec)).astype(np.float32)
        small = w < np.float32(5.0)
        ws = (w - np.float32(2.5)).astype(np.float32)
        wl = (np.sqrt(w).astype(np.float32) - np.float32(3.0)).astype(np.float32)
        ww = np.where(small, ws, wl).astype(np.float32)
        p = np.where(small, np.float32(_ERFINV_SMALL[0]), np.float32(_ERFINV_LARGE[0])).astype(np.float32)
        for cs, cl in zip(_ERFINV_SMALL[1:], _ERFINV_LARGE[1:]):
            c = np.where(small, np.float32(cs), np.float32(cl)).astype(np.float32)
            p = (p * ww).astype(np.float32)  # separate multiply ...
            p = (c + p).astype(np.float32)   # ... and add: no FMA
        r = (p * u).astype(np.float32)
        r = np.where(np.abs(u) == np.float32(1.0), np.float32(np.inf) * u, r).astype(np.float32)
    return r


_LO = np.nextafter(np.float32(-1.0), np.float32(0.0))
_SQRT2 = np.float32(np.sqrt(2.0))


def normal_from_bits(bits) -> np.ndarray:
    u = uniform_from_bits(bits, _LO, 1.0)
    return (_SQRT2 * erfinv_f32(u)).astype(np.float32)


def normal(key, shape, layout: int = LAYOUT_LEGACY, offset: int = 0) -> np.ndarray:
    """``jax.random.normal(key, shape, float32)``."""
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    return normal_from_bits(random_bits(key, n, layout, offset)).reshape(shape)


def uniform(key, shape, layout: int = LAYOUT_LEGACY, minval=0.0, maxval=1.0, offset: int = 0) -> np.ndarray:
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    return uniform_from_bits(random_bits(key, n, layout, offset), minval, maxval).reshape(shape)


def randint(key, n: int, minval: int, maxval: int, layout: int = LAYOUT_LEGACY) -> np.ndarray:
    """``jax.random.randint(key, (n,), minval, maxval)`` for the default int32 dtype -> int32[n]; also what
    ``jax.random.choice(key, N, (n,))`` (replace=True, p=None) returns for ``minval, maxval = 0, N``.

    Restates ``jax._src.random._randint``: ``k1, k2 = split(key)``, 32 "higher" bits from k1 and 32 "lower" bits from
    k2, ``multiplier = (2**16 % span)**2 % span`` and ``offset = ((higher % span) * multiplier + lower % span) % span``,
    every operation in uint32 and wrapping exactly as XLA's integer ops do (for span > 2**16 the squared multiplier wraps
    to 0 and the offset reduces to ``lower % span``).  Used by the on-device minibatch sampler (SURVEY.md section 8(f)
    row 3); the reference itself never draws indices.

    PARITY STATUS of this function: UNPINNED against JAX -- no JAX here and no published known-answer vector; it pins
    the CUDA sampler (bijax_b200/csrc/bjx_minibatch.cu) against an independent NumPy restatement of the same published
    algorithm, on top of the pinned threefry / split."""
    k1, k2 = split(key, 2, layout)
    hi = random_bits(k1, n, layout).astype(U32)
    lo = random_bits(k2, n, layout).astype(U32)
    span = np.array([maxval - minval if maxval > minval else 1], dtype=np.uint64).astype(U32)
    with np.errstate(over="ignore"):
        mult = np.array([1 << 16], dtype=U32) % span
        mult = (mult * mult) % span
        off = ((hi % span) * mult + lo % span) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


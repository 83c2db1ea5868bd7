"""CPU: the XLA FFI shim (bijax_b200/csrc/xla_ffi_shim.cc) type-checks against a minimal declaration stub of xla::ffi and
binds every entry point of include/bjx.h that a JAX host needs.  jaxlib is not installable here, so the shim can be neither
linked nor run: what this pins is (1) syntax and types of every handler, (2) that each handler's parameter list matches its
Bind() chain (the stub's binder static_asserts it), (3) which C entry points are reachable through the shim."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "bijax_b200" / "csrc" / "xla_ffi_shim.cc"
STUB = ROOT / "tests" / "ffi_stub"
CUDA_INC = Path("/usr/local/cuda/include")


def _compile(src: Path, extra=()):
    gxx = shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime_api.h").exists():
        pytest.skip("g++ or the CUDA runtime headers are missing")
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function", f"-I{STUB}", f"-I{ROOT / 'include'}",
           f"-I{CUDA_INC}", *extra, str(src)]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_shim_type_checks_against_the_ffi_stub():
    r = _compile(SHIM)
    assert r.returncode == 0, r.stderr[-4000:]


def test_stub_binder_rejects_a_mismatched_handler(tmp_path):
    """The check above means something: a handler whose parameters disagree with its Bind() chain must NOT compile."""
    bad = tmp_path / "bad.cc"
    bad.write_text('''
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;
static ffi::Error Impl(int stream, ffi::Buffer<ffi::F32> x, ffi::ResultBuffer<ffi::F32> out, int64_t n) { return ffi::Error::Success(); }
XLA_FFI_DEFINE_HANDLER_SYMBOL(Bad, Impl, ffi::Ffi::Bind().Ctx<ffi::PlatformStream<int>>().Arg<ffi::Buffer<ffi::F32>>()
                                            .Ret<ffi::Buffer<ffi::F32>>().Attr<float>("n").Attr<int32_t>("extra"));
''')
    r = _compile(bad)
    assert r.returncode != 0 and "does not match its Ffi::Bind() chain" in r.stderr


def test_shim_covers_the_entry_points_a_jax_host_needs():
    src = SHIM.read_text()
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src))
    assert handlers == {"BjxElboStep", "BjxPrepareXPlanes", "BjxPosteriorSample", "BjxPosteriorLogProb", "BjxRandomBits",
                        "BjxRandomUniform", "BjxRandomNormal", "BjxRandomSplit", "BjxAdamUpdate", "BjxTrainSteps"}
    called = set(re.findall(r"\b(bjx_[a-z0-9_]+)\(", src))
    assert {"bjx_elbo_step", "bjx_prepare_x_planes", "bjx_x_planes_bytes", "bjx_posterior_sample", "bjx_posterior_log_prob",
            "bjx_random_bits", "bjx_random_uniform", "bjx_random_normal", "bjx_random_split", "bjx_adam_update", "bjx_train_steps",
            "bjx_last_error"} <= called
    header = (ROOT / "include" / "bjx.h").read_text()
    for fn in called:
        assert re.search(rf"\b{fn}\(", header), f"{fn} is not declared in include/bjx.h"
    # every BjxElboArgs field a JAX host can vary is an operand or an attribute of BjxElboStep
    elbo = src[src.index("XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxElboStep"):src.index("XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxPrepareXPlanes")]
    for attr in ("S", "w_offset", "vi_type", "scale_bijector", "likelihood", "noise_src", "noise_index", "threefry_layout", "precision",
                 "noise_value", "ll_scale", "prior_weight", "rank", "S_total", "s_offset", "entropy_weight", "point_mode", "N_full"):
        assert f'"{attr}"' in elbo, attr
    assert "row_index" in elbo and "planes" in elbo

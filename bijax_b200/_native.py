"""ctypes binding of libbjx.so (the C ABI declared in include/bjx.h).

There is NO fallback: if the CUDA library is missing and cannot be built, importing the symbols raises.  torch is
used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_lib" / "libbjx.so"

# ---- constants mirrored from include/bjx.h ----------------------------------------------------------------------------
ABI_VERSION = 4
OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, -1, -2, -3, -4
LAYOUT_LEGACY, LAYOUT_PARTITIONABLE = 0, 1
VI_MEAN_FIELD, VI_FULL_RANK, VI_LOW_RANK = 0, 1, 2
SCALE_EXP, SCALE_SOFTPLUS, SCALE_IDENTITY = 0, 1, 2
BIJ_IDENTITY, BIJ_EXP, BIJ_SOFTPLUS, BIJ_SIGMOID = 0, 1, 2, 3
PRIOR_NORMAL, PRIOR_BETA, PRIOR_GAMMA = 0, 1, 2
LIK_BERNOULLI_LOGITS, LIK_GAUSSIAN, LIK_BERNOULLI_PROBS = 0, 1, 2
NOISE_CONST, NOISE_KWARG, NOISE_LATENT = 0, 1, 2
PREC_AUTO, PREC_FP32, PREC_BF16X3, PREC_BF16X2 = 0, 1, 2, 3
POINT_OFF, POINT_LOG_JOINT, POINT_MAP, POINT_LOGLIK = 0, 1, 2, 3
KERNEL_NAMES = {0: "fp32_tiled", 1: "fp32_smalld", 2: "tcgen05_bf16_split", 3: "coin", 4: "tcgen05_gemm_two_pass"}
KERNEL_IDS = {v: k for k, v in KERNEL_NAMES.items()}

EXPORTED_SYMBOLS = (
    "bjx_abi_version",
    "bjx_last_error",
    "bjx_device_sm_count",
    "bjx_random_bits",
    "bjx_random_uniform",
    "bjx_random_normal",
    "bjx_random_split",
    "bjx_elbo_workspace_bytes",
    "bjx_elbo_step",
    "bjx_elbo_step_host",
    "bjx_elbo_kernel_choice",
    "bjx_elbo_supports_row_index",
    "bjx_train_steps",
    "bjx_minibatch_gather",
    "bjx_train_steps_minibatch",
    "bjx_train_steps_sharded",
    "bjx_allreduce_packed_p2p_dev",
    "bjx_p2p_set_timeout_ms",
    "bjx_p2p_buffer_bytes",
    "bjx_p2p_alloc",
    "bjx_p2p_open",
    "bjx_p2p_close",
    "bjx_p2p_free",
    "bjx_allreduce_packed_p2p",
    "bjx_x_planes_bytes",
    "bjx_prepare_x_planes",
    "bjx_prepare_x_planes_rows",
    "bjx_posterior_sample",
    "bjx_posterior_log_prob",
    "bjx_fill_triangular",
    "bjx_fill_triangular_vjp",
    "bjx_adam_update",
    "bjx_profile_begin",
    "bjx_profile_collect",
    "bjx_profile_end",
    "bjx_debug_set_trace_buffer",
)


class BjxCoord(C.Structure):
    _fields_ = [("prior", C.c_uint32), ("chain", C.c_uint32), ("p0", C.c_float), ("p1", C.c_float), ("lognorm", C.c_float)]


class BjxElboArgs(C.Structure):
    _fields_ = [
        ("S", C.c_int64),
        ("N", C.c_int64),
        ("D", C.c_int64),
        ("Dw", C.c_int64),
        ("w_offset", C.c_int64),
        ("K", C.c_int64),
        ("ldx", C.c_int64),
        ("vi_type", C.c_int32),
        ("scale_bijector", C.c_int32),
        ("likelihood", C.c_int32),
        ("noise_src", C.c_int32),
        ("noise_index", C.c_int32),
        ("threefry_layout", C.c_int32),
        ("precision", C.c_int32),
        ("kernel_override", C.c_int32),
        ("noise_value", C.c_float),
        ("ll_scale", C.c_float),
        ("prior_weight", C.c_float),
        ("reserved1", C.c_float),
        ("key", C.c_void_p),
        ("loc", C.c_void_p),
        ("raw_scale", C.c_void_p),
        ("kwargs", C.c_void_p),
        ("X", C.c_void_p),
        ("y", C.c_void_p),
        ("coords", C.c_void_p),
        ("out", C.c_void_p),
        ("eps_out", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("x_planes", C.c_void_p),
        ("key_index", C.c_void_p),
        ("S_total", C.c_int64),
        ("s_offset", C.c_int64),
        ("entropy_weight", C.c_float),
        ("point_mode", C.c_int32),
        ("reserved3", C.c_int32),
        ("rank", C.c_int64),
        ("row_index", C.c_void_p),
        ("N_full", C.c_int64),
    ]


class BjxMinibatch(C.Structure):
    """include/bjx.h BjxMinibatch (the on-device minibatch sampler)."""

    _fields_ = [
        ("N_full", C.c_int64), ("B", C.c_int64), ("Dw", C.c_int64), ("ldx_full", C.c_int64),
        ("X_full", C.c_void_p), ("y_full", C.c_void_p), ("x_planes_full", C.c_void_p),
        ("Xb", C.c_void_p), ("yb", C.c_void_p), ("planes_b", C.c_void_p), ("idx_out", C.c_void_p), ("elbo_key", C.c_void_p),
    ]


class BjxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbjx error {code}: {message}")
        self.code = code


_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libbjx.so, building it in-tree first if it is absent and nvcc is available.  Raises otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        if os.environ.get("BJX_NO_AUTOBUILD"):
            raise ImportError(f"{_LIB_PATH} is missing and BJX_NO_AUTOBUILD is set; run `python -m bijax_b200.build`")
        from . import build as _build

        _build.build()
    lib = C.CDLL(str(_LIB_PATH))
    i32, i64, f32, f64, vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p
    lib.bjx_abi_version.restype = C.c_int
    lib.bjx_last_error.restype = C.c_char_p
    lib.bjx_device_sm_count.restype = C.c_int
    lib.bjx_random_bits.argtypes = [vp, i64, i64, i64, i32, vp, vp]
    lib.bjx_random_uniform.argtypes = [vp, i64, i64, i64, i32, f32, f32, vp, vp]
    lib.bjx_random_normal.argtypes = [vp, i64, i64, i64, i32, vp, vp]
    lib.bjx_random_split.argtypes = [vp, i64, i32, vp, vp]
    lib.bjx_elbo_workspace_bytes.argtypes = [C.POINTER(BjxElboArgs)]
    lib.bjx_elbo_workspace_bytes.restype = C.c_size_t
    lib.bjx_elbo_step.argtypes = [C.POINTER(BjxElboArgs), vp]
    lib.bjx_elbo_step_host.argtypes = [C.POINTER(BjxElboArgs), vp, vp, vp, vp]
    lib.bjx_elbo_kernel_choice.argtypes = [C.POINTER(BjxElboArgs)]
    lib.bjx_elbo_supports_row_index.argtypes = [C.POINTER(BjxElboArgs)]
    lib.bjx_train_steps.argtypes = [C.POINTER(BjxElboArgs), i64, vp, vp, vp, f64, f64, f64, f64, vp, vp, vp]
    lib.bjx_minibatch_gather.argtypes = [C.POINTER(BjxMinibatch), vp, vp, i32, vp]
    lib.bjx_train_steps_minibatch.argtypes = [C.POINTER(BjxElboArgs), C.POINTER(BjxMinibatch), i64, vp, vp, vp, f64, f64, f64, f64, vp, vp, vp]
    lib.bjx_train_steps_sharded.argtypes = [C.POINTER(BjxElboArgs), i64, vp, vp, vp, f64, f64, f64, f64, vp, vp, vp, i32, i32, C.c_uint64, vp]
    lib.bjx_allreduce_packed_p2p_dev.argtypes = [vp, i64, vp, i32, i32, C.c_uint64, vp, vp]
    lib.bjx_p2p_set_timeout_ms.argtypes = [i64]
    lib.bjx_p2p_buffer_bytes.argtypes = [i64, i32]
    lib.bjx_p2p_buffer_bytes.restype = C.c_size_t
    lib.bjx_p2p_alloc.argtypes = [C.c_size_t, C.POINTER(vp), vp]
    lib.bjx_p2p_open.argtypes = [vp, C.POINTER(vp)]
    lib.bjx_p2p_close.argtypes = [vp]
    lib.bjx_p2p_free.argtypes = [vp]
    lib.bjx_allreduce_packed_p2p.argtypes = [vp, i64, C.POINTER(vp), i32, i32, C.c_uint64, vp]
    lib.bjx_x_planes_bytes.argtypes = [i64, i64]
    lib.bjx_x_planes_bytes.restype = C.c_size_t
    lib.bjx_prepare_x_planes.argtypes = [vp, i64, i64, i64, vp, vp]
    lib.bjx_prepare_x_planes_rows.argtypes = [vp, i64, i64, i64, vp, i64, i64, vp]
    lib.bjx_posterior_sample.argtypes = [C.POINTER(BjxElboArgs), vp, vp, vp]
    lib.bjx_posterior_log_prob.argtypes = [C.POINTER(BjxElboArgs), vp, i64, vp, vp]
    lib.bjx_fill_triangular.argtypes = [vp, i64, vp, vp]
    lib.bjx_fill_triangular_vjp.argtypes = [vp, i64, vp, vp]
    lib.bjx_adam_update.argtypes = [vp, vp, vp, vp, i64, i64, f64, f64, f64, f64, vp]
    lib.bjx_profile_begin.argtypes = [i32]
    lib.bjx_profile_collect.argtypes = [vp, i32, vp]
    lib.bjx_profile_end.argtypes = []
    lib.bjx_debug_set_trace_buffer.argtypes = [vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)  # AttributeError here = the library does not export what bjx.h declares
        if name not in ("bjx_last_error", "bjx_elbo_workspace_bytes", "bjx_abi_version", "bjx_device_sm_count", "bjx_x_planes_bytes", "bjx_p2p_buffer_bytes"):
            fn.restype = C.c_int
    if lib.bjx_abi_version() != ABI_VERSION:
        raise ImportError(f"libbjx ABI {lib.bjx_abi_version()} != expected {ABI_VERSION}; rebuild with `python -m bijax_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        raise BjxError(rc, load().bjx_last_error().decode("utf-8", "replace"))
    return rc


def current_stream_ptr() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("bijax_b200 runs its hot path on CUDA (sm_100a) only; no CUDA device is visible and there is no CPU fallback")

// XLA FFI shim: exposes bjx_elbo_step as an XLA custom call so that bijax's own host (JAX) can reach the engine from
// inside jit + lax.scan + value_and_grad (bijax/utils.py:22-31).  Compiled ONLY where jaxlib's headers exist
// (jax.ffi.include_dir()); neither this image nor the GPU box has them, so this file is not part of libbjx.so and is
// untested here -- all logic lives below the C ABI (include/bjx.h), which IS tested.  INTEGRATION.md shows the Python side.
//
//   g++ -O2 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//       bijax_b200/csrc/xla_ffi_shim.cc -Lbijax_b200/_lib -lbjx -o bijax_b200/_lib/libbjx_xla.so
#if __has_include("xla/ffi/api/ffi.h")
#include "xla/ffi/api/ffi.h"

#include <cuda_runtime_api.h>

#include "bjx.h"

namespace ffi = xla::ffi;

// Operands: key u32[2], loc f32[D], raw_scale f32[P], kwargs f32[K], X f32[N, Dw], y f32[N], coords u8[D*20],
//           planes u8[bjx_x_planes_bytes(N, Dw)] (output of BjxPrepareXPlanes, computed once outside the scan) or u8[0].
// Result:   out f32[1 + D + P + K] (loss, then every gradient) and a u8 workspace sized by bjx_elbo_workspace_bytes.
// Attributes carry the static configuration (S, family ids, ll_scale, prior_weight, ...).
static ffi::Error ElboStepImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> loc,
                               ffi::Buffer<ffi::F32> raw_scale, ffi::Buffer<ffi::F32> kwargs, ffi::Buffer<ffi::F32> X,
                               ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::U8> coords, ffi::Buffer<ffi::U8> planes,
                               ffi::ResultBuffer<ffi::F32> out,
                               ffi::ResultBuffer<ffi::U8> workspace, int64_t S, int64_t w_offset, int32_t vi_type,
                               int32_t scale_bijector, int32_t likelihood, int32_t noise_src, int32_t noise_index,
                               int32_t threefry_layout, int32_t precision, float noise_value, float ll_scale,
                               float prior_weight) {
  BjxElboArgs a{};
  a.S = S;
  a.N = static_cast<int64_t>(y.element_count());
  a.D = static_cast<int64_t>(loc.element_count());
  a.K = static_cast<int64_t>(kwargs.element_count());
  const auto xd = X.dimensions();
  a.Dw = xd.size() == 2 ? xd[1] : 0;
  a.ldx = a.Dw;
  a.w_offset = w_offset;
  a.vi_type = vi_type;
  a.scale_bijector = scale_bijector;
  a.likelihood = likelihood;
  a.noise_src = noise_src;
  a.noise_index = noise_index;
  a.threefry_layout = threefry_layout;
  a.precision = precision;
  a.noise_value = noise_value;
  a.ll_scale = ll_scale;
  a.prior_weight = prior_weight;
  a.key = key.typed_data();
  a.loc = loc.typed_data();
  a.raw_scale = raw_scale.typed_data();
  a.kwargs = a.K ? kwargs.typed_data() : nullptr;
  a.X = a.Dw ? X.typed_data() : nullptr;
  a.y = y.typed_data();
  a.coords = reinterpret_cast<const BjxCoord*>(coords.typed_data());
  a.x_planes = planes.element_count() ? planes.typed_data() : nullptr;  // dataset prepared once (bjx_prepare_x_planes)
  a.out = out->typed_data();
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  const int rc = bjx_elbo_step(&a, stream);  // enqueue only: no sync, no allocation -> command-buffer capturable
  if (rc == BJX_OK) return ffi::Error::Success();
  return ffi::Error(rc == BJX_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    bjx_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxElboStep, ElboStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("S")
                                  .Attr<int64_t>("w_offset")
                                  .Attr<int32_t>("vi_type")
                                  .Attr<int32_t>("scale_bijector")
                                  .Attr<int32_t>("likelihood")
                                  .Attr<int32_t>("noise_src")
                                  .Attr<int32_t>("noise_index")
                                  .Attr<int32_t>("threefry_layout")
                                  .Attr<int32_t>("precision")
                                  .Attr<float>("noise_value")
                                  .Attr<float>("ll_scale")
                                  .Attr<float>("prior_weight"));

// planes u8[bjx_x_planes_bytes(N, Dw)] = split-bf16 operand planes of X f32[N, Dw]: a pure function of X, called once per
// dataset outside train_fn's scan.
static ffi::Error PrepareXPlanesImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> X, ffi::ResultBuffer<ffi::U8> planes) {
  const auto xd = X.dimensions();
  if (xd.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "X must be [N, Dw]");
  if (planes->element_count() < bjx_x_planes_bytes(xd[0], xd[1]))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "planes buffer smaller than bjx_x_planes_bytes(N, Dw)");
  const int rc = bjx_prepare_x_planes(X.typed_data(), xd[0], xd[1], xd[1], planes->typed_data(), stream);
  if (rc == BJX_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, bjx_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxPrepareXPlanes, PrepareXPlanesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
#else
// jaxlib headers not found: nothing to build (see the header comment).
#endif

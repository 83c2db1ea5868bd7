// XLA FFI shim: exposes the C ABI of include/bjx.h as XLA custom calls so that bijax's own host (JAX) can reach the engine
// from inside jit + lax.scan + value_and_grad (bijax/utils.py:22-31).  One handler per C entry point a JAX host needs:
//
//   BjxElboStep           bjx_elbo_step            ADVI.loss_fn + jax.value_and_grad        advi.py:80-95, utils.py:7
//   BjxPrepareXPlanes     bjx_prepare_x_planes     dataset set-up (X is a scan constant)    utils.py:22-31
//   BjxPosteriorSample    bjx_posterior_sample     Posterior.sample                         core.py:46-55
//   BjxPosteriorLogProb   bjx_posterior_log_prob   Posterior.log_prob / .prob               core.py:57-75
//   BjxRandomBits/Uniform/Normal/Split             jax.random.* (for hosts that want the device generator)
//   BjxAdamUpdate         bjx_adam_update          optimizer.update + apply_updates         utils.py:26-27
//   BjxTrainSteps         bjx_train_steps          the scan body of train_fn, n times       utils.py:22-31
//
// Compiled into a separate library ONLY where jaxlib's headers exist (jax.ffi.include_dir()):
//
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude
//       -I/usr/local/cuda/include bijax_b200/csrc/xla_ffi_shim.cc -Lbijax_b200/_lib -lbjx -o bijax_b200/_lib/libbjx_xla.so
//
// Neither this image nor the GPU box has jaxlib, so the shim cannot be linked or run here; tests/test_ffi_shim_cpu.py
// type-checks it (g++ -fsyntax-only) against tests/ffi_stub/xla/ffi/api/ffi.h, a minimal declaration stub of the
// xla::ffi surface used below whose binder verifies every handler's signature against its Bind() chain.  All logic lives
// below the C ABI, which IS tested.  INTEGRATION.md shows the Python side (ffi_call + custom_vjp).
//
// Conventions: absent optional operands are passed as zero-length buffers (kwargs f32[0], planes u8[0], row_index s32[0]);
// every handler only enqueues on the stream XLA supplies (no sync, no allocation: command-buffer capturable); scratch is a
// u8 result buffer sized by bjx_elbo_workspace_bytes at trace time; state updated in place (BjxAdamUpdate, BjxTrainSteps)
// is declared with input_output_aliases on the Python side -- if XLA hands over distinct buffers anyway, the operand is
// copied into the result first.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include <cstdint>

#include "bjx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int rc) {
  if (rc == BJX_OK) return ffi::Error::Success();
  const ffi::ErrorCode code = rc == BJX_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument
                              : rc == BJX_ERR_UNSUPPORTED    ? ffi::ErrorCode::kUnimplemented
                              : rc == BJX_ERR_WORKSPACE      ? ffi::ErrorCode::kResourceExhausted
                                                             : ffi::ErrorCode::kInternal;
  return ffi::Error(code, bjx_last_error());
}

// bring an in/out state buffer into its result buffer when XLA did not alias them
template <typename In, typename Out>
int CarryOver(cudaStream_t stream, In& in, Out& out, size_t bytes) {
  if (static_cast<const void*>(in.typed_data()) == static_cast<const void*>(out->typed_data())) return BJX_OK;
  return cudaMemcpyAsync(out->typed_data(), in.typed_data(), bytes, cudaMemcpyDeviceToDevice, stream) == cudaSuccess ? BJX_OK
                                                                                                                     : BJX_ERR_CUDA;
}

// ---- the hot path -------------------------------------------------------------------------------------------------------
// Operands: key u32[2], loc f32[D], raw_scale f32[P], kwargs f32[K], X f32[N, Dw] (or f32[0]), y f32[N], coords u8[D*20],
//           planes u8[bjx_x_planes_bytes(N, Dw)] | u8[0], row_index s32[N] | s32[0] (then X is the FULL dataset [N_full, Dw]).
// Results:  out f32[1 + D + P + K] (loss, then every gradient), workspace u8[bjx_elbo_workspace_bytes].
ffi::Error ElboStepImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> loc, ffi::Buffer<ffi::F32> raw_scale,
                        ffi::Buffer<ffi::F32> kwargs, ffi::Buffer<ffi::F32> X, ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::U8> coords,
                        ffi::Buffer<ffi::U8> planes, ffi::Buffer<ffi::S32> row_index, ffi::ResultBuffer<ffi::F32> out,
                        ffi::ResultBuffer<ffi::U8> workspace, int64_t S, int64_t w_offset, int32_t vi_type, int32_t scale_bijector,
                        int32_t likelihood, int32_t noise_src, int32_t noise_index, int32_t threefry_layout, int32_t precision,
                        float noise_value, float ll_scale, float prior_weight, int64_t rank, int64_t S_total, int64_t s_offset,
                        float entropy_weight, int32_t point_mode, int64_t N_full) {
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
  a.rank = rank;
  a.S_total = S_total;
  a.s_offset = s_offset;
  a.entropy_weight = entropy_weight;
  a.point_mode = point_mode;
  a.N_full = N_full;  // rows of the full dataset behind row_index + planes (0: not that route)
  a.key = key.typed_data();
  a.loc = loc.typed_data();
  a.raw_scale = raw_scale.typed_data();
  a.kwargs = a.K ? kwargs.typed_data() : nullptr;
  a.X = a.Dw ? X.typed_data() : nullptr;
  a.y = y.typed_data();
  a.coords = reinterpret_cast<const BjxCoord*>(coords.typed_data());
  a.x_planes = planes.element_count() ? planes.typed_data() : nullptr;  // dataset prepared once (BjxPrepareXPlanes)
  if (row_index.element_count()) {
    if (static_cast<int64_t>(row_index.element_count()) != a.N)
      return ffi::Error(ffi::ErrorCode::kInvalidArgument, "row_index must have one entry per row of the batch (len(y))");
    a.row_index = row_index.typed_data();
  }
  a.out = out->typed_data();
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  return Status(bjx_elbo_step(&a, stream));
}

// planes u8[bjx_x_planes_bytes(N, Dw)] = split-bf16 operand planes of X f32[N, Dw]: a pure function of X, called once per
// dataset outside train_fn's scan.
ffi::Error PrepareXPlanesImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> X, ffi::ResultBuffer<ffi::U8> planes) {
  const auto xd = X.dimensions();
  if (xd.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "X must be [N, Dw]");
  if (planes->element_count() < bjx_x_planes_bytes(xd[0], xd[1]))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "planes buffer smaller than bjx_x_planes_bytes(N, Dw)");
  return Status(bjx_prepare_x_planes(X.typed_data(), xd[0], xd[1], xd[1], planes->typed_data(), stream));
}

// ---- posterior access ---------------------------------------------------------------------------------------------------
void FillPosteriorArgs(BjxElboArgs& a, const ffi::Buffer<ffi::F32>& loc, const ffi::Buffer<ffi::F32>& raw_scale,
                       const ffi::Buffer<ffi::U8>& coords, int32_t vi_type, int32_t scale_bijector, int64_t rank) {
  a.D = static_cast<int64_t>(loc.element_count());
  a.vi_type = vi_type;
  a.scale_bijector = scale_bijector;
  a.rank = rank;
  a.likelihood = BJX_LIK_BERNOULLI_PROBS;  // data-free: no likelihood scratch is planned
  a.ll_scale = a.prior_weight = 1.0f;
  a.loc = loc.typed_data();
  a.raw_scale = raw_scale.typed_data();
  a.coords = reinterpret_cast<const BjxCoord*>(coords.typed_data());
}

// theta f32[S, D] = bijector(loc + scale * normal(key, (S, D)))
ffi::Error PosteriorSampleImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> loc, ffi::Buffer<ffi::F32> raw_scale,
                               ffi::Buffer<ffi::U8> coords, ffi::ResultBuffer<ffi::F32> theta, ffi::ResultBuffer<ffi::U8> workspace,
                               int64_t S, int32_t vi_type, int32_t scale_bijector, int32_t threefry_layout, int64_t rank) {
  BjxElboArgs a{};
  FillPosteriorArgs(a, loc, raw_scale, coords, vi_type, scale_bijector, rank);
  a.S = S;
  a.threefry_layout = threefry_layout;
  a.key = key.typed_data();
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  if (static_cast<int64_t>(theta->element_count()) != S * a.D)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "theta must be [S, D]");
  return Status(bjx_posterior_sample(&a, theta->typed_data(), nullptr, stream));
}

// out f32[n] = log q(theta_i) for constrained samples theta f32[n, D]
ffi::Error PosteriorLogProbImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> theta, ffi::Buffer<ffi::F32> loc,
                                ffi::Buffer<ffi::F32> raw_scale, ffi::Buffer<ffi::U8> coords, ffi::ResultBuffer<ffi::F32> out,
                                ffi::ResultBuffer<ffi::U8> workspace, int32_t vi_type, int32_t scale_bijector, int64_t rank) {
  BjxElboArgs a{};
  FillPosteriorArgs(a, loc, raw_scale, coords, vi_type, scale_bijector, rank);
  a.S = 1;
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  const int64_t n = static_cast<int64_t>(out->element_count());
  if (static_cast<int64_t>(theta.element_count()) != n * a.D)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "theta must be [n, D] with n = len(out)");
  return Status(bjx_posterior_log_prob(&a, theta.typed_data(), n, out->typed_data(), stream));
}

// ---- jax.random on the device ------------------------------------------------------------------------------------------
ffi::Error RandomBitsImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::ResultBuffer<ffi::U32> out, int64_t total,
                          int64_t offset, int32_t layout) {
  return Status(bjx_random_bits(key.typed_data(), total, offset, static_cast<int64_t>(out->element_count()), layout,
                                out->typed_data(), stream));
}

ffi::Error RandomUniformImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::ResultBuffer<ffi::F32> out, int64_t total,
                             int64_t offset, int32_t layout, float minval, float maxval) {
  return Status(bjx_random_uniform(key.typed_data(), total, offset, static_cast<int64_t>(out->element_count()), layout, minval,
                                   maxval, out->typed_data(), stream));
}

ffi::Error RandomNormalImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::ResultBuffer<ffi::F32> out, int64_t total,
                            int64_t offset, int32_t layout) {
  return Status(bjx_random_normal(key.typed_data(), total, offset, static_cast<int64_t>(out->element_count()), layout,
                                  out->typed_data(), stream));
}

// out u32[num, 2] = jax.random.split(key, num)
ffi::Error RandomSplitImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::ResultBuffer<ffi::U32> out, int32_t layout) {
  return Status(bjx_random_split(key.typed_data(), static_cast<int64_t>(out->element_count() / 2), layout, out->typed_data(), stream));
}

// ---- optimiser and the device-resident loop -------------------------------------------------------------------------------
// (params, m, v) are updated in place: declare input_output_aliases={0: 0, 1: 1, 2: 2} in the ffi_call
ffi::Error AdamUpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> m, ffi::Buffer<ffi::F32> v,
                          ffi::Buffer<ffi::F32> grads, ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> m_out,
                          ffi::ResultBuffer<ffi::F32> v_out, int64_t step, double lr, double b1, double b2, double eps) {
  const size_t n = params.element_count();
  if (m.element_count() != n || v.element_count() != n || grads.element_count() != n)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "params, m, v and grads must have the same length");
  int rc = CarryOver(stream, params, params_out, n * sizeof(float));
  if (rc == BJX_OK) rc = CarryOver(stream, m, m_out, n * sizeof(float));
  if (rc == BJX_OK) rc = CarryOver(stream, v, v_out, n * sizeof(float));
  if (rc != BJX_OK) return ffi::Error(ffi::ErrorCode::kInternal, "device-to-device copy of the optimiser state failed");
  return Status(bjx_adam_update(params_out->typed_data(), m_out->typed_data(), v_out->typed_data(), grads.typed_data(),
                                static_cast<int64_t>(n), step, lr, b1, b2, eps, stream));
}

// n_steps x { ELBO + grads at keys[*steps_done] -> Adam -> losses[*steps_done], ++*steps_done }: train_fn's whole scan as ONE
// custom call.  keys u32[n, 2]; state (params_flat, m, v, steps_done s64[1], losses f32[n_total]) aliased in -> out.
ffi::Error TrainStepsImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> keys, ffi::Buffer<ffi::F32> params_flat, ffi::Buffer<ffi::F32> m,
                          ffi::Buffer<ffi::F32> v, ffi::Buffer<ffi::S64> steps_done, ffi::Buffer<ffi::F32> losses,
                          ffi::Buffer<ffi::F32> X, ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::U8> coords, ffi::Buffer<ffi::U8> planes,
                          ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> m_out, ffi::ResultBuffer<ffi::F32> v_out,
                          ffi::ResultBuffer<ffi::S64> steps_out, ffi::ResultBuffer<ffi::F32> losses_out,
                          ffi::ResultBuffer<ffi::F32> step_out, ffi::ResultBuffer<ffi::U8> workspace, int64_t n_steps, int64_t S,
                          int64_t D, int64_t K, int64_t w_offset, int32_t vi_type, int32_t scale_bijector, int32_t likelihood,
                          int32_t noise_src, int32_t noise_index, int32_t threefry_layout, int32_t precision, float noise_value,
                          float ll_scale, int64_t rank, double lr, double b1, double b2, double eps) {
  BjxElboArgs a{};
  a.S = S;
  a.N = static_cast<int64_t>(y.element_count());
  a.D = D;
  a.K = K;
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
  a.prior_weight = 1.0f;
  a.rank = rank;
  a.key = keys.typed_data();
  a.X = a.Dw ? X.typed_data() : nullptr;
  a.y = y.typed_data();
  a.coords = reinterpret_cast<const BjxCoord*>(coords.typed_data());
  a.x_planes = planes.element_count() ? planes.typed_data() : nullptr;
  a.out = step_out->typed_data();
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  const size_t n = params_flat.element_count();
  int rc = CarryOver(stream, params_flat, params_out, n * sizeof(float));
  if (rc == BJX_OK) rc = CarryOver(stream, m, m_out, n * sizeof(float));
  if (rc == BJX_OK) rc = CarryOver(stream, v, v_out, n * sizeof(float));
  if (rc == BJX_OK) rc = CarryOver(stream, steps_done, steps_out, sizeof(int64_t));
  if (rc == BJX_OK) rc = CarryOver(stream, losses, losses_out, losses.element_count() * sizeof(float));
  if (rc != BJX_OK) return ffi::Error(ffi::ErrorCode::kInternal, "device-to-device copy of the training state failed");
  a.loc = a.raw_scale = params_out->typed_data();  // replaced inside the library by views into params_flat
  return Status(bjx_train_steps(&a, n_steps, params_out->typed_data(), m_out->typed_data(), v_out->typed_data(), lr, b1, b2, eps,
                                steps_out->typed_data(), losses_out->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxElboStep, ElboStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()   // key
                                  .Arg<ffi::Buffer<ffi::F32>>()   // loc
                                  .Arg<ffi::Buffer<ffi::F32>>()   // raw_scale
                                  .Arg<ffi::Buffer<ffi::F32>>()   // kwargs
                                  .Arg<ffi::Buffer<ffi::F32>>()   // X
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y
                                  .Arg<ffi::Buffer<ffi::U8>>()    // coords
                                  .Arg<ffi::Buffer<ffi::U8>>()    // planes
                                  .Arg<ffi::Buffer<ffi::S32>>()   // row_index
                                  .Ret<ffi::Buffer<ffi::F32>>()   // out
                                  .Ret<ffi::Buffer<ffi::U8>>()    // workspace
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
                                  .Attr<float>("prior_weight")
                                  .Attr<int64_t>("rank")
                                  .Attr<int64_t>("S_total")
                                  .Attr<int64_t>("s_offset")
                                  .Attr<float>("entropy_weight")
                                  .Attr<int32_t>("point_mode")
                                  .Attr<int64_t>("N_full"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxPrepareXPlanes, PrepareXPlanesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxPosteriorSample, PosteriorSampleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("S")
                                  .Attr<int32_t>("vi_type")
                                  .Attr<int32_t>("scale_bijector")
                                  .Attr<int32_t>("threefry_layout")
                                  .Attr<int64_t>("rank"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxPosteriorLogProb, PosteriorLogProbImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("vi_type")
                                  .Attr<int32_t>("scale_bijector")
                                  .Attr<int64_t>("rank"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxRandomBits, RandomBitsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Attr<int64_t>("total")
                                  .Attr<int64_t>("offset")
                                  .Attr<int32_t>("layout"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxRandomUniform, RandomUniformImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("total")
                                  .Attr<int64_t>("offset")
                                  .Attr<int32_t>("layout")
                                  .Attr<float>("minval")
                                  .Attr<float>("maxval"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxRandomNormal, RandomNormalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("total")
                                  .Attr<int64_t>("offset")
                                  .Attr<int32_t>("layout"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxRandomSplit, RandomSplitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Attr<int32_t>("layout"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxAdamUpdate, AdamUpdateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("step")
                                  .Attr<double>("lr")
                                  .Attr<double>("b1")
                                  .Attr<double>("b2")
                                  .Attr<double>("eps"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(BjxTrainSteps, TrainStepsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()   // keys [n, 2]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params_flat
                                  .Arg<ffi::Buffer<ffi::F32>>()   // adam m
                                  .Arg<ffi::Buffer<ffi::F32>>()   // adam v
                                  .Arg<ffi::Buffer<ffi::S64>>()   // steps_done [1]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // losses
                                  .Arg<ffi::Buffer<ffi::F32>>()   // X
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y
                                  .Arg<ffi::Buffer<ffi::U8>>()    // coords
                                  .Arg<ffi::Buffer<ffi::U8>>()    // planes
                                  .Ret<ffi::Buffer<ffi::F32>>()   // params_flat (aliased)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // m (aliased)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // v (aliased)
                                  .Ret<ffi::Buffer<ffi::S64>>()   // steps_done (aliased)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // losses (aliased)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // packed output of the last step (scratch)
                                  .Ret<ffi::Buffer<ffi::U8>>()    // workspace
                                  .Attr<int64_t>("n_steps")
                                  .Attr<int64_t>("S")
                                  .Attr<int64_t>("D")
                                  .Attr<int64_t>("K")
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
                                  .Attr<int64_t>("rank")
                                  .Attr<double>("lr")
                                  .Attr<double>("b1")
                                  .Attr<double>("b2")
                                  .Attr<double>("eps"));
#else
// jaxlib headers not found: nothing to build (see the header comment).
#endif

// Shared host/device helpers for libbjx (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bjx.h"

namespace bjx {

// ---- host-side error plumbing (thread-local message, see bjx_api.cu) --------------------------------------------
int fail(int code, const char* fmt, ...);

#define BJX_CUDA_TRY(expr)                                                                             \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) return ::bjx::fail(BJX_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define BJX_LAUNCH_CHECK(what)                                                                          \
  do {                                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                                \
    if (_e != cudaSuccess) return ::bjx::fail(BJX_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(_e)); \
  } while (0)

#define BJX_REQUIRE(cond, ...)                                             \
  do {                                                                     \
    if (!(cond)) return ::bjx::fail(BJX_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

int sm_count();

// per-device "this function's attribute has been set" flags (function attributes are per device)
struct DeviceOnce {
  bool done[64] = {};
  bool test_and_set() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;  // unknown device: set the attribute every time
    const bool was = done[dev];
    done[dev] = true;
    return was;
  }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- Threefry-2x32-20 (Salmon et al. SC'11; the generator under jax.random) ---------------------------------------
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define BJX_TF_ROUND(r) \
  x0 += x1;             \
  x1 = rotl32(x1, r);   \
  x1 ^= x0;
  BJX_TF_ROUND(13) BJX_TF_ROUND(15) BJX_TF_ROUND(26) BJX_TF_ROUND(6)
  x0 += k1;
  x1 += ks2 + 1u;
  BJX_TF_ROUND(17) BJX_TF_ROUND(29) BJX_TF_ROUND(16) BJX_TF_ROUND(24)
  x0 += ks2;
  x1 += k0 + 2u;
  BJX_TF_ROUND(13) BJX_TF_ROUND(15) BJX_TF_ROUND(26) BJX_TF_ROUND(6)
  x0 += k0;
  x1 += k1 + 3u;
  BJX_TF_ROUND(17) BJX_TF_ROUND(29) BJX_TF_ROUND(16) BJX_TF_ROUND(24)
  x0 += k1;
  x1 += ks2 + 4u;
  BJX_TF_ROUND(13) BJX_TF_ROUND(15) BJX_TF_ROUND(26) BJX_TF_ROUND(6)
  x0 += ks2;
  x1 += k0 + 5u;
#undef BJX_TF_ROUND
}

// 32 random bits for flat index i of a stream of `total` draws (jax._src.prng counter layouts).
__device__ __forceinline__ uint32_t random_bits_at(uint32_t k0, uint32_t k1, uint64_t i, uint64_t total, int layout) {
  uint32_t a, b;
  if (layout == BJX_LAYOUT_PARTITIONABLE) {
    a = (uint32_t)(i >> 32);
    b = (uint32_t)i;
    threefry2x32(k0, k1, a, b);
    return a ^ b;
  }
  // legacy: iota(total) (+ one zero pad if odd) split in halves; pair j = (j, j + h)
  const uint64_t h = (total + 1) >> 1;
  const bool second = i >= h;
  const uint64_t j = second ? i - h : i;
  const uint64_t c1 = j + h;
  a = (uint32_t)j;
  b = (c1 < total) ? (uint32_t)c1 : 0u;  // the pad element
  threefry2x32(k0, k1, a, b);
  return second ? b : a;
}

// jax.random.uniform bit trick followed by the affine map; all roundings explicit (no FMA contraction).
__device__ __forceinline__ float bits_to_unit(uint32_t bits) {
  return __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
}

__device__ __forceinline__ float bits_to_uniform(uint32_t bits, float lo, float hi) {
  const float scale = __fsub_rn(hi, lo);
  const float u = __fadd_rn(__fmul_rn(bits_to_unit(bits), scale), lo);
  return fmaxf(lo, u);
}

// ---- the pinned float spec of oracle/threefry.py: binary32 round-to-nearest, every multiply and add rounded separately ----
#define BJX_MADD(p, x, c) __fadd_rn(__fmul_rn((p), (x)), (c))

// Cephes logf as XLA:CPU inlines it (GenerateVF32Log), for positive normal x; 0 -> -inf
__device__ __forceinline__ float logf_spec(float x) {
  const uint32_t bits = __float_as_uint(x);
  float e = (float)((int)((bits >> 23) & 0xFFu) - 126);
  float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F000000u);  // mantissa in [0.5, 1)
  const bool lt = m < 0.707106781186547524f;
  e = lt ? __fsub_rn(e, 1.0f) : e;
  m = __fsub_rn(lt ? __fadd_rn(m, m) : m, 1.0f);
  const float z = __fmul_rn(m, m);
  float y = 7.0376836292e-2f;
  y = BJX_MADD(y, m, -1.1514610310e-1f);
  y = BJX_MADD(y, m, 1.1676998740e-1f);
  y = BJX_MADD(y, m, -1.2420140846e-1f);
  y = BJX_MADD(y, m, 1.4249322787e-1f);
  y = BJX_MADD(y, m, -1.6668057665e-1f);
  y = BJX_MADD(y, m, 2.0000714765e-1f);
  y = BJX_MADD(y, m, -2.4999993993e-1f);
  y = BJX_MADD(y, m, 3.3333331174e-1f);
  y = __fmul_rn(__fmul_rn(y, m), z);
  y = __fadd_rn(y, __fmul_rn(e, -2.12194440e-4f));
  y = __fsub_rn(y, __fmul_rn(0.5f, z));
  float r = __fadd_rn(m, y);
  r = __fadd_rn(r, __fmul_rn(e, 0.693359375f));
  return x == 0.0f ? __int_as_float(0xff800000) : r;
}

// XLA's single-precision Log1p (EmitLog1p): Cephes rational below sqrt(2) - 1, log(1 + x) of the rounded sum above
__device__ __forceinline__ float log1p_spec(float x) {
  if (fabsf(x) < 0.41421356237309504880f) {
    float num = 4.5270000862445199635215e-5f;
    num = BJX_MADD(num, x, 4.9854102823193375972212e-1f);
    num = BJX_MADD(num, x, 6.5787325942061044846969e0f);
    num = BJX_MADD(num, x, 2.9911919328553073277375e1f);
    num = BJX_MADD(num, x, 6.0949667980987787057556e1f);
    num = BJX_MADD(num, x, 5.7112963590585538103336e1f);
    num = BJX_MADD(num, x, 2.0039553499201281259648e1f);
    float den = 1.0f;
    den = BJX_MADD(den, x, 1.5062909083469192043167e1f);
    den = BJX_MADD(den, x, 8.3047565967967209469434e1f);
    den = BJX_MADD(den, x, 2.2176239823732856465394e2f);
    den = BJX_MADD(den, x, 3.0909872225312059774938e2f);
    den = BJX_MADD(den, x, 2.1642788614495947685003e2f);
    den = BJX_MADD(den, x, 6.0118660497603843919306e1f);
    const float x2 = __fmul_rn(x, x);
    const float q = __fmul_rn(__fmul_rn(x, x2), __fdiv_rn(num, den));
    const float z = __fadd_rn(__fmul_rn(-0.5f, x2), q);
    return __fadd_rn(x, z);
  }
  return logf_spec(__fadd_rn(1.0f, x));
}

// XLA ErfInv (f32, Giles' polynomial): w = -log1p(-u*u), then one of two degree-8 polynomials
__device__ __forceinline__ float erfinv_spec(float u) {
  const float t = __fmul_rn(u, u);
  float w = -log1p_spec(-t);
  float p;
  if (w < 5.0f) {
    w = __fsub_rn(w, 2.5f);
    p = 2.81022636e-08f;
    p = __fadd_rn(3.43273939e-07f, __fmul_rn(p, w));
    p = __fadd_rn(-3.5233877e-06f, __fmul_rn(p, w));
    p = __fadd_rn(-4.39150654e-06f, __fmul_rn(p, w));
    p = __fadd_rn(0.00021858087f, __fmul_rn(p, w));
    p = __fadd_rn(-0.00125372503f, __fmul_rn(p, w));
    p = __fadd_rn(-0.00417768164f, __fmul_rn(p, w));
    p = __fadd_rn(0.246640727f, __fmul_rn(p, w));
    p = __fadd_rn(1.50140941f, __fmul_rn(p, w));
  } else {
    w = __fsub_rn(__fsqrt_rn(w), 3.0f);
    p = -0.000200214257f;
    p = __fadd_rn(0.000100950558f, __fmul_rn(p, w));
    p = __fadd_rn(0.00134934322f, __fmul_rn(p, w));
    p = __fadd_rn(-0.00367342844f, __fmul_rn(p, w));
    p = __fadd_rn(0.00573950773f, __fmul_rn(p, w));
    p = __fadd_rn(-0.0076224613f, __fmul_rn(p, w));
    p = __fadd_rn(0.00943887047f, __fmul_rn(p, w));
    p = __fadd_rn(1.00167406f, __fmul_rn(p, w));
    p = __fadd_rn(2.83297682f, __fmul_rn(p, w));
  }
  const float r = __fmul_rn(p, u);
  return (fabsf(u) == 1.0f) ? __fmul_rn(__int_as_float(0x7f800000), u) : r;
}

__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = __int_as_float(0xBF7FFFFF);  // nextafter(-1, 0)
  const float u = bits_to_uniform(bits, lo, 1.0f);
  return __fmul_rn(1.41421354f, erfinv_spec(u));  // float32(sqrt(2))
}

// ---- numerically careful elementwise pieces --------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float x) {
  // log(1 + e^x) = max(x, 0) + log1p(e^{-|x|})
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float sigmoid_f(float x) {
  // 1 / (1 + e^{-x}) without overflow on either side
  const float e = expf(-fabsf(x));
  const float s = 1.0f / (1.0f + e);
  return x >= 0.0f ? s : e * s;
}

// ---- reductions -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over a block of up to 1024 threads; result valid in thread 0. `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : 0.0f;
    v = warp_sum(v);
  }
  return v;
}

// ---- launchers implemented across the .cu files ---------------------------------------------------------------------
struct Workspace;  // laid out in bjx_elbo.cu

}  // namespace bjx

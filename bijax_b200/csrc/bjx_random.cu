// Kernel 1: counter-based threefry2x32 that reproduces jax.random.{bits,uniform,normal,split}
// (call sites in the reference: bijax/advi.py:77,83; bijax/utils.py:30,53-55,70).
#include "bjx_common.cuh"

namespace bjx {

enum { MODE_BITS = 0, MODE_UNIFORM = 1, MODE_NORMAL = 2 };

template <int MODE>
__device__ __forceinline__ void store_draw(void* out, uint64_t idx, uint32_t bits, float lo, float hi) {
  if (MODE == MODE_BITS) {
    ((uint32_t*)out)[idx] = bits;
  } else if (MODE == MODE_UNIFORM) {
    ((float*)out)[idx] = bits_to_uniform(bits, lo, hi);
  } else {
    ((float*)out)[idx] = bits_to_normal(bits);
  }
}

// Partitionable layout: one threefry block per element, counter = global flat index. Pure streaming store.
template <int MODE>
__global__ void __launch_bounds__(256) k_random_partitionable(const uint32_t* __restrict__ key, uint64_t offset,
                                                               uint64_t n, void* __restrict__ out, float lo, float hi) {
  const uint32_t k0 = key[0], k1 = key[1];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t g = i + offset;
    uint32_t a = (uint32_t)(g >> 32), b = (uint32_t)g;
    threefry2x32(k0, k1, a, b);
    store_draw<MODE>(out, i, a ^ b, lo, hi);
  }
}

// Legacy layout: one threefry block per PAIR (j, j+h); both outputs are used, so no work is wasted.
template <int MODE>
__global__ void __launch_bounds__(256) k_random_legacy(const uint32_t* __restrict__ key, uint64_t total,
                                                        void* __restrict__ out, float lo, float hi) {
  const uint32_t k0 = key[0], k1 = key[1];
  const uint64_t h = (total + 1) >> 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < h; j += stride) {
    const uint64_t c1 = j + h;
    uint32_t a = (uint32_t)j, b = c1 < total ? (uint32_t)c1 : 0u;
    threefry2x32(k0, k1, a, b);
    store_draw<MODE>(out, j, a, lo, hi);
    if (c1 < total) store_draw<MODE>(out, c1, b, lo, hi);
  }
}

// jax.random.split: legacy = bits(2*num) reshaped (num, 2); partitionable = both threefry words of counter i.
__global__ void __launch_bounds__(256) k_split_partitionable(const uint32_t* __restrict__ key, uint64_t num,
                                                              uint32_t* __restrict__ out) {
  const uint32_t k0 = key[0], k1 = key[1];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num; i += stride) {
    uint32_t a = (uint32_t)(i >> 32), b = (uint32_t)i;
    threefry2x32(k0, k1, a, b);
    out[2 * i] = a;
    out[2 * i + 1] = b;
  }
}

static int grid_for(uint64_t work, int block) {
  const int sms = sm_count();
  uint64_t g = (work + block - 1) / block;
  const uint64_t cap = (uint64_t)(sms > 0 ? sms : 148) * 16;  // persistent grid-stride: a whole number of waves
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <int MODE>
static int launch_random(const uint32_t* key, int64_t total, int64_t offset, int64_t n, int32_t layout, void* out,
                         float lo, float hi, cudaStream_t st) {
  BJX_REQUIRE(key != nullptr && (out != nullptr || n == 0), "random: null pointer");
  BJX_REQUIRE(total >= 0 && n >= 0 && offset >= 0 && offset + n <= total, "random: need 0 <= offset, offset+n <= total");
  if (n == 0) return BJX_OK;
  if (layout == BJX_LAYOUT_PARTITIONABLE) {
    k_random_partitionable<MODE><<<grid_for(n, 256), 256, 0, st>>>(key, (uint64_t)offset, (uint64_t)n, out, lo, hi);
  } else if (layout == BJX_LAYOUT_LEGACY) {
    BJX_REQUIRE(offset == 0 && n == total, "random: the legacy layout is not shard-stable; need offset == 0 and n == total");
    BJX_REQUIRE(total <= 0xFFFFFFFFll, "random: legacy layout supports at most 2^32-1 draws per key");
    k_random_legacy<MODE><<<grid_for((n + 1) / 2, 256), 256, 0, st>>>(key, (uint64_t)total, out, lo, hi);
  } else {
    return fail(BJX_ERR_INVALID_ARGUMENT, "random: unknown layout %d", layout);
  }
  BJX_LAUNCH_CHECK("k_random");
  return BJX_OK;
}

}  // namespace bjx

extern "C" {

int bjx_random_bits(const uint32_t* key, int64_t total, int64_t offset, int64_t n, int32_t layout, uint32_t* out,
                    void* stream) {
  return bjx::launch_random<bjx::MODE_BITS>(key, total, offset, n, layout, out, 0.f, 1.f, (cudaStream_t)stream);
}

int bjx_random_uniform(const uint32_t* key, int64_t total, int64_t offset, int64_t n, int32_t layout, float minval,
                       float maxval, float* out, void* stream) {
  return bjx::launch_random<bjx::MODE_UNIFORM>(key, total, offset, n, layout, out, minval, maxval, (cudaStream_t)stream);
}

int bjx_random_normal(const uint32_t* key, int64_t total, int64_t offset, int64_t n, int32_t layout, float* out,
                      void* stream) {
  return bjx::launch_random<bjx::MODE_NORMAL>(key, total, offset, n, layout, out, 0.f, 1.f, (cudaStream_t)stream);
}

int bjx_random_split(const uint32_t* key, int64_t num, int32_t layout, uint32_t* out, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(key != nullptr && (out != nullptr || num == 0), "split: null pointer");
  BJX_REQUIRE(num >= 0, "split: num must be >= 0");
  if (num == 0) return BJX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == BJX_LAYOUT_LEGACY) {
    return launch_random<MODE_BITS>(key, 2 * num, 0, 2 * num, BJX_LAYOUT_LEGACY, out, 0.f, 1.f, st);
  } else if (layout == BJX_LAYOUT_PARTITIONABLE) {
    k_split_partitionable<<<grid_for(num, 256), 256, 0, st>>>(key, (uint64_t)num, out);
    BJX_LAUNCH_CHECK("k_split");
    return BJX_OK;
  }
  return fail(BJX_ERR_INVALID_ARGUMENT, "split: unknown layout %d", layout);
}

}  // extern "C"

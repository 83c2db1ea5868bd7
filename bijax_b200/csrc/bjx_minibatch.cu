// On-device minibatch sampler (the reference's only concession to large N is the
// full_data_size / len(outputs) scaling of bijax/advi.py:92 -- the rows themselves are picked by the caller).
//
// What a JAX user writes around loss_fn,
//     k_batch, k_elbo = jax.random.split(key)
//     idx = jax.random.choice(k_batch, N, (B,))            # replace=True, p=None  ==  jax.random.randint(k_batch, (B,), 0, N)
//     loss_fn(params, y[idx], {"X": X[idx]}, full_data_size=N, seed=k_elbo, n_samples=S)
// runs here as ONE kernel in front of the step: the row indices are recomputed from the key (a handful of threefry blocks
// per row) and those rows of X (raw fp32 or the prepared split-bf16 planes) and y are copied into the batch buffers.
// HBM traffic: B rows read + B rows written; the full dataset is never touched.  When the step's kernel can read the
// dataset through the index itself (BjxElboArgs.row_index) only idx and y are produced here.
//
// jax.random.randint (jax/_src/random.py, _randint; int32 => 32-bit draws, all arithmetic in uint32 and wrapping as XLA's):
//     k1, k2 = split(key);  hi = bits(k1, (B,));  lo = bits(k2, (B,));  span = N
//     m = (2^16 mod span)^2 mod span;   idx = ((hi mod span) * m + lo mod span) mod span
#include "bjx_common.cuh"

namespace bjx {

// jax.random.split(key, 2) (same layouts as k_split_* in bjx_random.cu)
__device__ __forceinline__ void split2_keys(uint32_t k0, uint32_t k1, int layout, uint32_t (&ka)[2], uint32_t (&kb)[2]) {
  if (layout == BJX_LAYOUT_PARTITIONABLE) {  // key i = both threefry words of counter i
    uint32_t a = 0u, b = 0u;
    threefry2x32(k0, k1, a, b);
    ka[0] = a, ka[1] = b;
    a = 0u, b = 1u;
    threefry2x32(k0, k1, a, b);
    kb[0] = a, kb[1] = b;
  } else {  // bits(4) reshaped (2, 2): blocks (0, 2) and (1, 3) give [a0, a1, b0, b1]
    uint32_t a0 = 0u, b0 = 2u, a1 = 1u, b1 = 3u;
    threefry2x32(k0, k1, a0, b0);
    threefry2x32(k0, k1, a1, b1);
    ka[0] = a0, ka[1] = a1;
    kb[0] = b0, kb[1] = b1;
  }
}

__device__ __forceinline__ uint32_t randint_at(const uint32_t (&k1)[2], const uint32_t (&k2)[2], uint64_t b, uint64_t B, uint32_t span,
                                               uint32_t mult, int layout) {
  const uint32_t hi = random_bits_at(k1[0], k1[1], b, B, layout);
  const uint32_t lo = random_bits_at(k2[0], k2[1], b, B, layout);
  return ((hi % span) * mult + lo % span) % span;  // uint32, wrapping
}

struct GatherArgs {
  int64_t N, B, Dw, ldx, Dp;
  const float* X;
  const float* y;
  const uint4* planes;  // [2][N][Dp] bf16
  float* Xb;
  float* yb;
  uint4* planes_b;  // [2][B][Dp] bf16
  int32_t* idx_out;
  uint32_t* elbo_key_out;
  int layout, do_split, vec4;
};

// A warp owns `rpw` consecutive batch rows (a power of two <= 32): lane l < rpw draws the index of row l (one set of
// threefry blocks per row, not per lane) and writes idx / y coalesced; the warp then copies the rows one after the other,
// the index broadcast by shuffle.  rpw grows with B so that small batches still spread over all SMs.
__global__ void __launch_bounds__(256) k_minibatch_gather(const uint32_t* __restrict__ key, const int64_t* __restrict__ key_index,
                                                          GatherArgs g, int rpw) {
  const int lane = threadIdx.x & 31;
  const int64_t base = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * rpw;
  // the two key splits are the same for every row: once per block
  __shared__ uint32_t s_keys[4];
  if (threadIdx.x == 0) {
    const uint32_t* k = key_index ? key + 2 * (*key_index) : key;
    uint32_t kbatch[2] = {k[0], k[1]}, kelbo[2] = {0u, 0u};
    if (g.do_split) split2_keys(k[0], k[1], g.layout, kbatch, kelbo);
    if (g.elbo_key_out && blockIdx.x == 0) {
      g.elbo_key_out[0] = kelbo[0];
      g.elbo_key_out[1] = kelbo[1];
    }
    uint32_t k1[2], k2[2];
    split2_keys(kbatch[0], kbatch[1], g.layout, k1, k2);
    s_keys[0] = k1[0], s_keys[1] = k1[1], s_keys[2] = k2[0], s_keys[3] = k2[1];
  }
  __syncthreads();
  if (base >= g.B) return;
  int64_t mine = 0;
  if (lane < rpw && base + lane < g.B) {
    const uint32_t k1[2] = {s_keys[0], s_keys[1]}, k2[2] = {s_keys[2], s_keys[3]};
    const uint32_t span = (uint32_t)g.N;
    uint32_t mult = 65536u % span;
    mult = (mult * mult) % span;
    mine = (int64_t)randint_at(k1, k2, (uint64_t)(base + lane), (uint64_t)g.B, span, mult, g.layout);
    if (g.idx_out) g.idx_out[base + lane] = (int32_t)mine;
    if (g.yb) g.yb[base + lane] = g.y[mine];
  }
  if (!g.Xb && !g.planes_b) return;
  const int64_t row16 = g.Dp / 8;  // uint4 (8 bf16) per plane row
  for (int j = 0; j < rpw; ++j) {
    const int64_t b = base + j;
    if (b >= g.B) break;
    const int64_t r = __shfl_sync(0xffffffffu, mine, j);
    if (g.Xb) {
      const float* src = g.X + r * g.ldx;
      float* dst = g.Xb + b * g.Dw;
      if (g.vec4) {
        const float4* s4 = (const float4*)src;
        float4* d4 = (float4*)dst;
        for (int64_t c = lane; c < g.Dw / 4; c += 32) d4[c] = __ldg(s4 + c);
      } else {
        for (int64_t c = lane; c < g.Dw; c += 32) dst[c] = __ldg(src + c);
      }
    }
    if (g.planes_b) {
      for (int t = 0; t < 2; ++t) {
        const uint4* src = g.planes + ((int64_t)t * g.N + r) * row16;
        uint4* dst = g.planes_b + ((int64_t)t * g.B + b) * row16;
        for (int64_t c = lane; c < row16; c += 32) dst[c] = __ldg(src + c);
      }
    }
  }
}

int launch_minibatch_gather(const BjxMinibatch& mb, const uint32_t* key, const int64_t* key_index, int layout, cudaStream_t st) {
  BJX_REQUIRE(key != nullptr, "key is null");
  BJX_REQUIRE(mb.N_full >= 1 && mb.N_full < (1ll << 31), "minibatch: N_full must be in [1, 2^31) (int32 indices; got %lld)", (long long)mb.N_full);
  BJX_REQUIRE(mb.B >= 1 && mb.B < (1ll << 31), "minibatch: batch size must be in [1, 2^31)");
  BJX_REQUIRE(layout == BJX_LAYOUT_LEGACY || layout == BJX_LAYOUT_PARTITIONABLE, "unknown threefry layout %d", layout);
  BJX_REQUIRE(mb.yb == nullptr || mb.y_full != nullptr, "minibatch: yb given without y_full");
  BJX_REQUIRE(mb.Xb == nullptr || (mb.X_full != nullptr && mb.Dw >= 1 && mb.ldx_full >= mb.Dw), "minibatch: Xb needs X_full, Dw >= 1, ldx_full >= Dw");
  BJX_REQUIRE(mb.planes_b == nullptr || (mb.x_planes_full != nullptr && mb.Dw >= 1), "minibatch: planes_b needs x_planes_full and Dw");
  BJX_REQUIRE((((uintptr_t)mb.planes_b | (uintptr_t)mb.x_planes_full) & 15) == 0, "minibatch: planes must be 16-byte aligned");
  GatherArgs g;
  g.N = mb.N_full, g.B = mb.B, g.Dw = mb.Dw, g.ldx = mb.ldx_full, g.Dp = (mb.Dw + 63) / 64 * 64;
  g.X = mb.X_full, g.y = mb.y_full, g.planes = (const uint4*)mb.x_planes_full;
  g.Xb = mb.Xb, g.yb = mb.yb, g.planes_b = (uint4*)mb.planes_b;
  g.idx_out = mb.idx_out, g.elbo_key_out = mb.elbo_key;
  g.layout = layout, g.do_split = mb.elbo_key != nullptr ? 1 : 0;
  g.vec4 = (mb.Xb && mb.Dw % 4 == 0 && mb.ldx_full % 4 == 0 && (((uintptr_t)mb.X_full | (uintptr_t)mb.Xb) & 15) == 0) ? 1 : 0;
  int rpw = 32;  // index-only calls: one thread per row
  if (g.Xb || g.planes_b) {
    rpw = 1;
    while (rpw < 32 && mb.B / (2 * rpw) >= 4096) rpw *= 2;  // keep >= 4096 warps in flight before growing a warp's share
  }
  const int64_t warps = (mb.B + rpw - 1) / rpw;
  k_minibatch_gather<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(key, key_index, g, rpw);
  BJX_LAUNCH_CHECK("k_minibatch_gather");
  return BJX_OK;
}

}  // namespace bjx

extern "C" {

int bjx_minibatch_gather(const BjxMinibatch* mb, const uint32_t* key, const int64_t* key_index, int32_t layout, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(mb != nullptr, "mb is null");
  return launch_minibatch_gather(*mb, key, key_index, layout, (cudaStream_t)stream);
}

}  // extern "C"

// Posterior.log_prob (core.py:57-71): for constrained samples theta,
//     log q(theta) = q.log_prob(z) + sum_d inverse_log_det_jacobian_d(theta_d),   z = b^{-1}(theta)
// with q = MultivariateNormalDiag(loc, sigma) or MultivariateNormalTriL(loc, tril(L)) in the unconstrained space.
//   mean field: one warp per sample.
//   full rank : one block per sample; blocked forward substitution L w = z - loc (32-row panels: a block-wide GEMV against
//               the already solved part, then the 32 x 32 diagonal block solved by one warp with shuffles).
#include "bjx_elbo.cuh"
#include "bjx_variational.cuh"

namespace bjx {

// z = b^{-1}(theta) for a packed chain (lowest nibble applied first), and fldj_b(z) summed along the chain
__device__ __forceinline__ void inverse_coord(float theta, uint32_t chain, float& z_out, float& fldj_out) {
  uint32_t ops[8];
  int n = 0;
  while ((chain & 0xFu) && n < 8) {
    ops[n++] = chain & 0xFu;
    chain >>= 4;
  }
  float x = theta;
  for (int i = n - 1; i >= 0; --i) {
    const uint32_t op = ops[i];
    if (op == BJX_BIJ_EXP) {
      x = logf(x);
    } else if (op == BJX_BIJ_SOFTPLUS) {
      x = x + logf(-expm1f(-x));  // tfb.Softplus inverse: log(expm1(y))
    } else {                       // sigmoid^{-1} = logit
      x = logf(x) - log1pf(-x);
    }
  }
  z_out = x;
  float ldj = 0.0f;
  for (int i = 0; i < n; ++i) {
    const uint32_t op = ops[i];
    if (op == BJX_BIJ_EXP) {
      ldj += x;
      x = expf(x);
    } else if (op == BJX_BIJ_SOFTPLUS) {
      ldj += -softplus_f(-x);
      x = softplus_f(x);
    } else {
      ldj += -softplus_f(-x) - softplus_f(x);
      x = sigmoid_f(x);
    }
  }
  fldj_out = ldj;
}

__global__ void __launch_bounds__(256) k_post_logprob_mf(const float* __restrict__ theta, int64_t n, int64_t D,
                                                         const float* __restrict__ loc, const float* __restrict__ raw,
                                                         const BjxCoord* __restrict__ coords, int scale_bij,
                                                         float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (s >= n) return;
  float acc = 0.0f;
  for (int64_t d = lane; d < D; d += 32) {
    float z, fldj;
    inverse_coord(theta[s * D + d], coords[d].chain, z, fldj);
    const ScaleEval sc = eval_scale(raw[d], scale_bij);
    const float zz = (z - loc[d]) / sc.sigma;
    const float logsig = scale_bij == BJX_SCALE_EXP ? raw[d] : logf(sc.sigma);
    acc += -0.5f * zz * zz - logsig - BJX_HALF_LOG_2PI - fldj;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[s] = acc;
}

__global__ void __launch_bounds__(256) k_post_logprob_fr(const float* __restrict__ theta, int64_t D,
                                                         const float* __restrict__ loc, const float* __restrict__ M,
                                                         const BjxCoord* __restrict__ coords, int scale_bij,
                                                         float* __restrict__ out) {
  extern __shared__ float sm[];
  float* w = sm;            // [D] solved part of L w = z - loc
  float* diff = sm + D;     // [D]
  __shared__ float rhs[32];
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const int64_t s = blockIdx.x;
  float part = 0.0f;  // - sum fldj - sum log|L_ii|
  for (int64_t d = tid; d < D; d += 256) {
    float z, fldj;
    inverse_coord(theta[s * D + d], coords[d].chain, z, fldj);
    diff[d] = z - loc[d];
    part += -fldj - logf(fabsf(eval_scale(M[d * D + d], scale_bij).sigma));
  }
  __syncthreads();
  for (int64_t r0 = 0; r0 < D; r0 += 32) {
    // rows r0 .. r0+31 against the solved w[0 .. r0): four rows per warp, lanes stride over j
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int64_t i = r0 + wq * 4 + rr;
      float acc = 0.0f;
      if (i < D)
        for (int64_t j = lane; j < r0; j += 32) acc = fmaf(eval_scale(M[i * D + j], scale_bij).sigma, w[j], acc);
      acc = warp_sum(acc);
      if (lane == 0) rhs[wq * 4 + rr] = i < D ? diff[i] - acc : 0.0f;
    }
    __syncthreads();
    if (wq == 0) {
      // 32 x 32 lower-triangular diagonal block: lane l owns row r0 + l
      const int64_t i = r0 + lane;
      float b = rhs[lane];
      for (int k = 0; k < 32; ++k) {
        const int64_t col = r0 + k;
        if (col >= D) break;  // warp-uniform
        const float lkk = eval_scale(M[col * D + col], scale_bij).sigma;
        const float xk = __shfl_sync(0xffffffffu, b, k) / lkk;
        if (lane == k) w[col] = xk;
        if (lane > k && i < D) b = fmaf(-eval_scale(M[i * D + col], scale_bij).sigma, xk, b);
      }
    }
    __syncthreads();
  }
  for (int64_t d = tid; d < D; d += 256) part += -0.5f * w[d] * w[d] - BJX_HALF_LOG_2PI;
  const float tot = block_sum(part, red);
  if (tid == 0) out[s] = tot;
}

// tfp.math.fill_triangular (lower): the D x D matrix is reshape(concat(x[D:], reverse(x)), (D, D)) with the strict upper
// triangle zeroed; element (i, j), j <= i, therefore comes from x[src] with k = i*D + j, m = D(D+1)/2:
//     src = D + k            if k < m - D        else        src = m - 1 - (k - (m - D))
__device__ __forceinline__ int64_t fill_tril_src(int64_t i, int64_t j, int64_t D) {
  const int64_t m = D * (D + 1) / 2, k = i * D + j;
  return k < m - D ? D + k : m - 1 - (k - (m - D));
}

__global__ void __launch_bounds__(256) k_fill_tril(const float* __restrict__ packed, int64_t D, float* __restrict__ dense) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= D * D) return;
  const int64_t i = idx / D, j = idx % D;
  dense[idx] = j <= i ? packed[fill_tril_src(i, j, D)] : 0.0f;
}

// VJP: every packed entry appears exactly once in the lower triangle
__global__ void __launch_bounds__(256) k_fill_tril_vjp(const float* __restrict__ d_dense, int64_t D, float* __restrict__ d_packed) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= D * D) return;
  const int64_t i = idx / D, j = idx % D;
  if (j <= i) d_packed[fill_tril_src(i, j, D)] = d_dense[idx];
}

}  // namespace bjx

extern "C" int bjx_fill_triangular(const float* packed, int64_t D, float* dense, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(packed && dense && D >= 1 && D < (1ll << 30), "bad arguments");
  k_fill_tril<<<(unsigned)((D * D + 255) / 256), 256, 0, (cudaStream_t)stream>>>(packed, D, dense);
  BJX_LAUNCH_CHECK("k_fill_tril");
  return BJX_OK;
}

extern "C" int bjx_fill_triangular_vjp(const float* d_dense, int64_t D, float* d_packed, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(d_dense && d_packed && D >= 1 && D < (1ll << 30), "bad arguments");
  k_fill_tril_vjp<<<(unsigned)((D * D + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_dense, D, d_packed);
  BJX_LAUNCH_CHECK("k_fill_tril_vjp");
  return BJX_OK;
}

extern "C" int bjx_posterior_log_prob(const BjxElboArgs* args, const float* theta, int64_t n, float* out, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(args != nullptr && theta != nullptr && out != nullptr, "null pointer");
  const BjxElboArgs& a = *args;
  BJX_REQUIRE(a.D >= 1 && n >= 0, "need D >= 1 and n >= 0");
  BJX_REQUIRE(a.loc && a.raw_scale && a.coords, "null loc / raw_scale / coords pointer");
  BJX_REQUIRE(a.vi_type == BJX_VI_MEAN_FIELD || a.vi_type == BJX_VI_FULL_RANK || a.vi_type == BJX_VI_LOW_RANK, "unknown vi_type %d", a.vi_type);
  BJX_REQUIRE(a.scale_bijector >= BJX_SCALE_EXP && a.scale_bijector <= BJX_SCALE_IDENTITY, "unknown scale bijector %d", a.scale_bijector);
  if (n == 0) return BJX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.vi_type == BJX_VI_LOW_RANK) {
    BJX_REQUIRE(a.rank >= 1 && a.rank <= 64, "low rank needs 1 <= rank <= 64");
    BJX_REQUIRE(a.workspace != nullptr && a.workspace_bytes >= lowrank_posterior_bytes(a),
                "low-rank Posterior.log_prob needs a workspace of (D * rank + D + 4) floats");
    BJX_REQUIRE(((size_t)a.D + a.rank) * sizeof(float) <= 200 * 1024 && n <= 0x7FFFFFFFll, "D too large for the low-rank log_prob kernel");
    return launch_lowrank_posterior_log_prob(a, theta, n, out, st);
  }
  if (a.vi_type == BJX_VI_MEAN_FIELD) {
    k_post_logprob_mf<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(theta, n, a.D, a.loc, a.raw_scale, a.coords, a.scale_bijector, out);
    BJX_LAUNCH_CHECK("k_post_logprob_mf");
  } else {
    const size_t smem = (size_t)2 * a.D * sizeof(float);
    BJX_REQUIRE(smem <= 200 * 1024, "full-rank Posterior.log_prob supports D <= 25600 (two fp32 vectors in shared memory)");
    BJX_REQUIRE(n <= 0x7FFFFFFFll, "too many samples for one launch");
    if (smem > 48 * 1024) BJX_CUDA_TRY(cudaFuncSetAttribute(k_post_logprob_fr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_post_logprob_fr<<<(unsigned)n, 256, smem, st>>>(theta, a.D, a.loc, a.raw_scale, a.coords, a.scale_bijector, out);
    BJX_LAUNCH_CHECK("k_post_logprob_fr");
  }
  return BJX_OK;
}

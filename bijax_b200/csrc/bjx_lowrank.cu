// The low-rank variational family (core.py:190-201): q = MultivariateNormalDiagPlusLowRank(loc, scale_diag = d,
// scale_perturb_factor = U [D, r]), i.e. the scale operator  A = diag(d) + U U^T  (d = Exp(raw_d) by default).
//   sample (advi.py:83):     z_s = loc + A eps_s = loc + d * eps_s + U (U^T eps_s),     eps_s ~ N(0, I_D)
//   q.log_prob (advi.py:86): -1/2 |A^-1 (z_s - loc)|^2 - log|det A| - D/2 log 2pi, and A^-1 (z_s - loc) == eps_s
//   log|det A| = sum log d_i + log det C,   C = I_r + U^T diag(1/d) U     (matrix determinant lemma)
//   d log|det A| / d d_i = [A^-1]_ii = 1/d_i - sum_k W_ik U_ik / d_i,    d log|det A| / d U = 2 W,    W = diag(1/d) U C^-1 = A^-1 U
//   reparameterisation gradient: d L / d d_i = sum_s gz_si eps_si,   d L / d U = sum_s gz_s (U^T eps_s)^T + eps_s (U^T gz_s)^T
// Everything here is O(S D r + D r^2) with r <= 64: small next to the likelihood pass, which is unchanged (it consumes theta).
#include "bjx_elbo.cuh"
#include "bjx_variational.cuh"

namespace bjx {

// out[s, k] = sum_i U[i, k] v[s, i]      (v = eps -> a;  v = gz -> b).  One block per sample, one warp per k (strided).
__global__ void __launch_bounds__(256) k_lr_project(const float* __restrict__ U, const float* __restrict__ v, int64_t D, int r,
                                                    float* __restrict__ out) {
  const int64_t s = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = w; k < r; k += 8) {
    float acc = 0.0f;
    for (int64_t i = lane; i < D; i += 32) acc = fmaf(U[i * r + k], v[s * D + i], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[s * r + k] = acc;
  }
}

// z[s, i] = loc[i] + d_i eps[s, i] + sum_k U[i, k] a[s, k]
__global__ void __launch_bounds__(256) k_lr_z(const float* __restrict__ loc, const float* __restrict__ raw_d,
                                              const float* __restrict__ U, const float* __restrict__ eps,
                                              const float* __restrict__ a, int64_t S, int64_t D, int r, int scale_bij,
                                              float* __restrict__ z) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= S * D) return;
  const int64_t s = idx / D, i = idx % D;
  float acc = fmaf(eval_scale(raw_d[i], scale_bij).sigma, eps[idx], loc[i]);
  for (int k = 0; k < r; ++k) acc = fmaf(U[i * r + k], a[s * r + k], acc);
  z[idx] = acc;
}

// One block: C = I + U^T diag(1/d) U, its Cholesky factor and inverse in shared memory; then
//   scal[0] = log|det A|,  W[D, r] = diag(1/d) U C^-1,  ddiag[i] = [A^-1]_ii
__global__ void __launch_bounds__(256) k_lr_logdet(const float* __restrict__ raw_d, const float* __restrict__ U, int64_t D, int r,
                                                   int scale_bij, float* __restrict__ scal, float* __restrict__ W,
                                                   float* __restrict__ ddiag) {
  extern __shared__ float sm[];
  float* Cm = sm;                  // [r][r + 1]  C, then its Cholesky factor L (lower)
  float* Ci = Cm + r * (r + 1);    // [r][r + 1]  C^-1
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int ld = r + 1;
  for (int p = tid; p < r * r; p += 256) {
    const int k = p / r, l = p % r;
    float acc = k == l ? 1.0f : 0.0f;
    if (l <= k) {
      for (int64_t i = 0; i < D; ++i) acc = fmaf(U[i * r + k] / eval_scale(raw_d[i], scale_bij).sigma, U[i * r + l], acc);
      Cm[k * ld + l] = acc;
    }
  }
  __syncthreads();
  // Cholesky (lower), column by column
  for (int j = 0; j < r; ++j) {
    if (tid == 0) {
      float v = Cm[j * ld + j];
      for (int m = 0; m < j; ++m) v -= Cm[j * ld + m] * Cm[j * ld + m];
      Cm[j * ld + j] = sqrtf(v);
    }
    __syncthreads();
    for (int k = j + 1 + tid; k < r; k += 256) {
      float v = Cm[k * ld + j];
      for (int m = 0; m < j; ++m) v -= Cm[k * ld + m] * Cm[j * ld + m];
      Cm[k * ld + j] = v / Cm[j * ld + j];
    }
    __syncthreads();
  }
  // C^-1: thread c solves L L^T x = e_c
  if (tid < r) {
    const int c = tid;
    float x[64];
    for (int k = 0; k < r; ++k) {  // forward: L y = e_c
      float v = k == c ? 1.0f : 0.0f;
      for (int m = 0; m < k; ++m) v -= Cm[k * ld + m] * x[m];
      x[k] = v / Cm[k * ld + k];
    }
    for (int k = r - 1; k >= 0; --k) {  // backward: L^T x = y
      float v = x[k];
      for (int m = k + 1; m < r; ++m) v -= Cm[m * ld + k] * x[m];
      x[k] = v / Cm[k * ld + k];
    }
    for (int k = 0; k < r; ++k) Ci[k * ld + c] = x[k];
  }
  __syncthreads();
  float part = 0.0f;
  for (int64_t i = tid; i < D; i += 256) {
    const float di = eval_scale(raw_d[i], scale_bij).sigma;
    part += logf(fabsf(di));
    float dd = 1.0f / di;
    for (int k = 0; k < r; ++k) {
      float wk = 0.0f;
      for (int l = 0; l < r; ++l) wk = fmaf(U[i * r + l] / di, Ci[l * ld + k], wk);
      W[i * r + k] = wk;
      dd -= wk * U[i * r + k] / di;
    }
    ddiag[i] = dd;
  }
  for (int k = tid; k < r; k += 256) part += 2.0f * logf(Cm[k * ld + k]);
  const float tot = block_sum(part, red);
  if (tid == 0) scal[0] = tot;
}

// d raw_d[i] = sigma'(raw_i) (sum_s gz_si eps_si - ew [A^-1]_ii),   dU[i, k] = sum_s (gz_si a_sk + eps_si b_sk) - 2 ew W_ik
__global__ void __launch_bounds__(256) k_lr_grads(const float* __restrict__ gz, const float* __restrict__ eps,
                                                  const float* __restrict__ a, const float* __restrict__ b,
                                                  const float* __restrict__ raw_d, const float* __restrict__ W,
                                                  const float* __restrict__ ddiag, int64_t S, int64_t D, int r, int scale_bij,
                                                  float entropy_weight, float* __restrict__ d_raw, float* __restrict__ dU) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= D * (r + 1)) return;
  const int64_t i = idx / (r + 1);
  const int k = (int)(idx % (r + 1)) - 1;
  if (k < 0) {
    float acc = 0.0f;
    for (int64_t s = 0; s < S; ++s) acc = fmaf(gz[s * D + i], eps[s * D + i], acc);
    d_raw[i] = eval_scale(raw_d[i], scale_bij).dsigma * (acc - entropy_weight * ddiag[i]);
  } else {
    float acc = 0.0f;
    for (int64_t s = 0; s < S; ++s) acc += gz[s * D + i] * a[s * r + k] + eps[s * D + i] * b[s * r + k];
    dU[i * r + k] = acc - 2.0f * entropy_weight * W[i * r + k];
  }
}

// Posterior.log_prob for the low-rank family: one block per sample, Woodbury solve
//   A^-1 v = v / d - W (U^T (v / d)),   v = b^-1(theta) - loc,   W = diag(1/d) U C^-1  (k_lr_logdet)
__device__ __forceinline__ void lr_inverse_coord(float theta, uint32_t chain, float& z_out, float& fldj_out);

__global__ void __launch_bounds__(256) k_post_logprob_lr(const float* __restrict__ theta, int64_t D, int r,
                                                         const float* __restrict__ loc, const float* __restrict__ raw_d,
                                                         const float* __restrict__ U, const float* __restrict__ W,
                                                         const float* __restrict__ scal, const BjxCoord* __restrict__ coords,
                                                         int scale_bij, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* vd = sm;        // [D]  (z - loc) / d
  float* t = sm + D;     // [r]  U^T vd
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t s = blockIdx.x;
  float part = 0.0f;  // - sum fldj
  for (int64_t i = tid; i < D; i += 256) {
    float z, fldj;
    lr_inverse_coord(theta[s * D + i], coords[i].chain, z, fldj);
    vd[i] = (z - loc[i]) / eval_scale(raw_d[i], scale_bij).sigma;
    part -= fldj;
  }
  __syncthreads();
  for (int k = w; k < r; k += 8) {
    float acc = 0.0f;
    for (int64_t i = lane; i < D; i += 32) acc = fmaf(U[i * r + k], vd[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) t[k] = acc;
  }
  __syncthreads();
  for (int64_t i = tid; i < D; i += 256) {
    float zz = vd[i];
    for (int k = 0; k < r; ++k) zz = fmaf(-W[i * r + k], t[k], zz);
    part += -0.5f * zz * zz - BJX_HALF_LOG_2PI;
  }
  const float tot = block_sum(part, red);
  if (tid == 0) out[s] = tot - scal[0];
}

// the inverse of a packed bijector chain and the forward log-det-Jacobian at the pre-image (same as bjx_posterior.cu)
__device__ __forceinline__ void lr_inverse_coord(float theta, uint32_t chain, float& z_out, float& fldj_out) {
  uint32_t ops[8];
  int n = 0;
  while ((chain & 0xFu) && n < 8) {
    ops[n++] = chain & 0xFu;
    chain >>= 4;
  }
  float x = theta;
  for (int i = n - 1; i >= 0; --i) {
    const uint32_t op = ops[i];
    if (op == BJX_BIJ_EXP) x = logf(x);
    else if (op == BJX_BIJ_SOFTPLUS) x = x + logf(-expm1f(-x));
    else x = logf(x) - log1pf(-x);
  }
  z_out = x;
  float ldj = 0.0f;
  for (int i = 0; i < n; ++i) {
    const uint32_t op = ops[i];
    if (op == BJX_BIJ_EXP) { ldj += x; x = expf(x); }
    else if (op == BJX_BIJ_SOFTPLUS) { ldj += -softplus_f(-x); x = softplus_f(x); }
    else { ldj += -softplus_f(-x) - softplus_f(x); x = sigmoid_f(x); }
  }
  fldj_out = ldj;
}

size_t lowrank_posterior_bytes(const BjxElboArgs& a) { return align_up(((size_t)a.D * a.rank + a.D + 4) * sizeof(float), 256); }

int launch_lowrank_posterior_log_prob(const BjxElboArgs& a, const float* theta, int64_t n, float* out, cudaStream_t st) {
  const int r = (int)a.rank;
  const float* raw_d = a.raw_scale;
  const float* U = a.raw_scale + a.D;
  float* W = (float*)a.workspace;
  float* ddiag = W + a.D * r;
  float* scal = ddiag + a.D;
  k_lr_logdet<<<1, 256, (size_t)2 * r * (r + 1) * sizeof(float), st>>>(raw_d, U, a.D, r, a.scale_bijector, scal, W, ddiag);
  BJX_LAUNCH_CHECK("k_lr_logdet");
  const size_t smem = ((size_t)a.D + r) * sizeof(float);
  if (smem > 48 * 1024) BJX_CUDA_TRY(cudaFuncSetAttribute(k_post_logprob_lr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_post_logprob_lr<<<(unsigned)n, 256, smem, st>>>(theta, a.D, r, a.loc, raw_d, U, W, scal, a.coords, a.scale_bijector, out);
  BJX_LAUNCH_CHECK("k_post_logprob_lr");
  return BJX_OK;
}

size_t lowrank_bytes(const BjxElboArgs& a) {
  const size_t S = (size_t)a.S, D = (size_t)a.D, r = (size_t)a.rank;
  return align_up((2 * S * r + D * r + D + 4) * sizeof(float), 256);
}

// eps (already drawn) -> a = U^T eps, z; and log|det A| with its derivatives (they do not depend on the samples)
int launch_lowrank_sample(const BjxElboArgs& a, const float* eps, float* z, char* scratch, cudaStream_t st) {
  const int r = (int)a.rank;
  const float* raw_d = a.raw_scale;
  const float* U = a.raw_scale + a.D;
  float* av = (float*)scratch;
  float* W = av + 2 * a.S * r;
  float* ddiag = W + a.D * r;
  float* scal = ddiag + a.D;
  k_lr_project<<<(unsigned)a.S, 256, 0, st>>>(U, eps, a.D, r, av);
  BJX_LAUNCH_CHECK("k_lr_project(eps)");
  k_lr_z<<<(unsigned)((a.S * a.D + 255) / 256), 256, 0, st>>>(a.loc, raw_d, U, eps, av, a.S, a.D, r, a.scale_bijector, z);
  BJX_LAUNCH_CHECK("k_lr_z");
  k_lr_logdet<<<1, 256, (size_t)2 * r * (r + 1) * sizeof(float), st>>>(raw_d, U, a.D, r, a.scale_bijector, scal, W, ddiag);
  BJX_LAUNCH_CHECK("k_lr_logdet");
  return BJX_OK;
}

const float* lowrank_logdet_ptr(const BjxElboArgs& a, const char* scratch) {
  return (const float*)scratch + 2 * a.S * a.rank + a.D * a.rank + a.D;
}

int launch_lowrank_grads(const BjxElboArgs& a, const float* gz, const float* eps, char* scratch, float entropy_weight,
                         float* d_raw, cudaStream_t st) {
  const int r = (int)a.rank;
  const float* U = a.raw_scale + a.D;
  float* av = (float*)scratch;
  float* bv = av + a.S * r;
  float* W = av + 2 * a.S * r;
  float* ddiag = W + a.D * r;
  k_lr_project<<<(unsigned)a.S, 256, 0, st>>>(U, gz, a.D, r, bv);
  BJX_LAUNCH_CHECK("k_lr_project(gz)");
  const int64_t n = a.D * (r + 1);
  k_lr_grads<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gz, eps, av, bv, a.raw_scale, W, ddiag, a.S, a.D, r, a.scale_bijector,
                                                          entropy_weight, d_raw, d_raw + a.D);
  BJX_LAUNCH_CHECK("k_lr_grads");
  return BJX_OK;
}

}  // namespace bjx

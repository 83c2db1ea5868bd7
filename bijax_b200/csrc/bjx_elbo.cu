// The ELBO + gradient step (advi.py:80-95 under utils.py:7): plan, prologue, tail and the C entry points.
//
// Launch sequence of one step (all on the caller's stream, no host sync, graph-capturable):
//   [full rank only: k_eps, k_fullrank_z]
//   k_prologue        eps = normal(key); z; theta = b(z); per-(s,d) prior/entropy terms and their z-derivatives
//   <likelihood>      one of: k_lik_tc (tcgen05), k_lik_smalld, k_lik_tiled_A+B, k_lik_coin  -> partial sums
//   k_reduce_partials deterministic reduction of the per-CTA partials
//   k_tail            chain rule through the bijector and the reparameterisation: d loc, d raw_scale (+ gz for full rank)
//   [full rank only: k_fullrank_dM]
//   k_scalars         loss and d kwargs
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bjx_elbo.cuh"
#include "bjx_sgemm.cuh"
#include "bjx_variational.cuh"

namespace bjx {

// ---- error plumbing ---------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

// ---- optional event pairs around the streaming-likelihood launch (bench.py's roofline measurement) -------------------
struct Profile {
  std::vector<cudaEvent_t> ev;  // 2 * capacity
  int capacity = 0, used = 0;
};
static thread_local Profile g_prof;

static void prof_mark(cudaStream_t st, int which) {
  if (g_prof.capacity == 0 || g_prof.used >= g_prof.capacity) return;
  cudaEventRecord(g_prof.ev[2 * g_prof.used + which], st);
  if (which == 1) ++g_prof.used;
}

// ---- plan ---------------------------------------------------------------------------------------------------------------
static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int make_plan(const BjxElboArgs& a, Plan& p) {
  memset(&p, 0, sizeof(p));
  BJX_REQUIRE(a.S >= 1 && a.S <= (1 << 20), "S (n_samples) must be in [1, 2^20] (got %lld)", (long long)a.S);
  BJX_REQUIRE(a.D >= 1, "D must be >= 1");
  BJX_REQUIRE(a.N >= 0, "N must be >= 0");
  BJX_REQUIRE(a.K >= 0, "K must be >= 0");
  BJX_REQUIRE(a.vi_type == BJX_VI_MEAN_FIELD || a.vi_type == BJX_VI_FULL_RANK || a.vi_type == BJX_VI_LOW_RANK,
              "unknown vi_type %d", a.vi_type);
  BJX_REQUIRE(a.vi_type != BJX_VI_LOW_RANK || (a.rank >= 1 && a.rank <= 64), "low rank needs 1 <= rank <= 64 (got %lld)", (long long)a.rank);
  BJX_REQUIRE(a.scale_bijector >= BJX_SCALE_EXP && a.scale_bijector <= BJX_SCALE_IDENTITY, "unknown scale bijector %d",
              a.scale_bijector);
  BJX_REQUIRE(a.threefry_layout == BJX_LAYOUT_LEGACY || a.threefry_layout == BJX_LAYOUT_PARTITIONABLE,
              "unknown threefry layout %d", a.threefry_layout);
  BJX_REQUIRE(a.point_mode >= BJX_POINT_OFF && a.point_mode <= BJX_POINT_LOGLIK, "unknown point_mode %d", a.point_mode);
  BJX_REQUIRE(a.point_mode == BJX_POINT_OFF || (a.S == 1 && a.S_total == 0 && a.vi_type == BJX_VI_MEAN_FIELD),
              "point evaluation needs S == 1, no sample sharding and vi_type = mean field (raw_scale is ignored)");
  BJX_REQUIRE(a.S_total == 0 || (a.s_offset >= 0 && a.s_offset + a.S <= a.S_total), "sample shard [s_offset, s_offset+S) must lie in [0, S_total)");
  BJX_REQUIRE((a.S_total ? a.S_total : a.S) * a.D <= 0xFFFFFFFFll, "S*D must fit the 32-bit legacy counter");
  const int sms = sm_count();
  if (sms <= 0) return fail(BJX_ERR_CUDA, "no CUDA device");

  if (a.likelihood == BJX_LIK_BERNOULLI_PROBS) {
    BJX_REQUIRE(a.w_offset >= 0 && a.w_offset < a.D, "w_offset out of range");
    p.kernel = KERNEL_COIN;
    p.tiny = (a.vi_type == BJX_VI_MEAN_FIELD && a.point_mode == BJX_POINT_OFF && a.S * a.D <= TINY_SD && a.S <= TINY_S &&
              a.N <= TINY_N && a.kernel_override == 0)
                 ? 1
                 : 0;
  } else if (a.likelihood == BJX_LIK_BERNOULLI_LOGITS || a.likelihood == BJX_LIK_GAUSSIAN) {
    BJX_REQUIRE(a.Dw >= 1 && a.w_offset >= 0 && a.w_offset + a.Dw <= a.D, "weights [w_offset, w_offset+Dw) must lie in [0, D)");
    BJX_REQUIRE(a.ldx >= a.Dw, "ldx must be >= Dw");
    if (a.likelihood == BJX_LIK_GAUSSIAN) {
      BJX_REQUIRE(a.noise_src >= BJX_NOISE_CONST && a.noise_src <= BJX_NOISE_LATENT, "unknown noise_src %d", a.noise_src);
      if (a.noise_src == BJX_NOISE_KWARG) BJX_REQUIRE(a.noise_index >= 0 && a.noise_index < a.K, "noise kwarg index out of range");
      if (a.noise_src == BJX_NOISE_LATENT)
        BJX_REQUIRE(a.noise_index >= 0 && a.noise_index < a.D && !(a.noise_index >= a.w_offset && a.noise_index < a.w_offset + a.Dw),
                    "latent noise coordinate out of range or inside the weights");
    }
    const bool want_tc = a.precision == BJX_PREC_BF16X2 || a.precision == BJX_PREC_BF16X3 || a.precision == BJX_PREC_AUTO;
    if (a.precision != BJX_PREC_FP32 && a.precision != BJX_PREC_AUTO && !tc_eligible(a) && !gemm_eligible(a))
      return fail(BJX_ERR_UNSUPPORTED, "tensor-core precision requested but the shape is outside the tcgen05 kernel's envelope");
    // the two-pass GEMM path pays for operand planes and two launches: use it when the contraction is large
    const bool gemm_worth = (double)a.N * (double)a.S * (double)a.Dw >= 134217728.0 || a.precision == BJX_PREC_BF16X2 ||
                            a.precision == BJX_PREC_BF16X3;
    if (want_tc && tc_eligible(a))
      p.kernel = KERNEL_TC_BF16;
    else if (want_tc && gemm_eligible(a) && gemm_worth)
      p.kernel = KERNEL_TC_GEMM;
    else if (a.Dw <= 15)
      p.kernel = KERNEL_FP32_SMALLD;
    else
      p.kernel = KERNEL_FP32_TILED;
    if (a.kernel_override != 0) {  // tests and benchmarks: force a specific streaming kernel (KERNEL_* + 1)
      const int k = a.kernel_override - 1;
      BJX_REQUIRE(k == KERNEL_FP32_TILED || k == KERNEL_FP32_SMALLD || k == KERNEL_TC_BF16 || k == KERNEL_TC_GEMM,
                  "kernel_override %d is not a streaming-likelihood kernel", a.kernel_override);
      if (k == KERNEL_TC_BF16 && !tc_eligible(a)) return fail(BJX_ERR_UNSUPPORTED, "shape outside the fused tcgen05 kernel's envelope");
      if (k == KERNEL_TC_GEMM && !gemm_eligible(a)) return fail(BJX_ERR_UNSUPPORTED, "shape outside the tcgen05 GEMM path's envelope");
      if (k == KERNEL_FP32_SMALLD && a.Dw > 15) return fail(BJX_ERR_UNSUPPORTED, "small-D kernel needs Dw <= 15");
      p.kernel = k;
    }
  } else {
    return fail(BJX_ERR_INVALID_ARGUMENT, "unknown likelihood family %d", a.likelihood);
  }

  if (a.row_index != nullptr) {
    if (p.kernel != KERNEL_TC_BF16 || a.N < 1 || (a.x_planes != nullptr && a.N_full < 1))
      return fail(BJX_ERR_UNSUPPORTED, "row_index: only the fused tcgen05 kernel reads through an index (raw fp32 X, or x_planes of the "
                                       "full dataset together with N_full); gather the rows instead");
  }
  p.Pn = a.vi_type == BJX_VI_MEAN_FIELD ? a.D : (a.vi_type == BJX_VI_FULL_RANK ? a.D * a.D : a.D + a.D * a.rank);
  p.n_qp_blocks = (int)((a.S * a.D + 255) / 256);
  const int64_t Dw = a.likelihood == BJX_LIK_BERNOULLI_PROBS ? 1 : a.Dw;

  switch (p.kernel) {
    case KERNEL_COIN:
      p.PG = p.PS = 1;
      break;
    case KERNEL_FP32_SMALLD: {
      p.sd_rec4 = (int)((a.Dw + 1 + 3) / 4);
      p.sd_SB = next_pow2((int)(a.S < 256 ? a.S : 256));
      p.sd_grid_y = (int)((a.S + p.sd_SB - 1) / p.sd_SB);
      int64_t gx = (a.N + SD_CHUNK_ROWS - 1) / SD_CHUNK_ROWS;
      const int64_t cap = (int64_t)sms * smalld_occupancy(p.sd_rec4) / p.sd_grid_y;  // exactly one resident wave
      if (gx > cap) gx = cap;
      if (gx < 1) gx = 1;
      p.sd_grid_x = (int)gx;
      p.PG = p.PS = p.sd_grid_x;
      break;
    }
    case KERNEL_FP32_TILED: {
      p.tA_grid_y = (int)((a.S + 31) / 32);
      int64_t mt = (a.N + 127) / 128;
      int64_t gx = (int64_t)sms * 2 / p.tA_grid_y;
      if (gx < 1) gx = 1;
      if (gx > mt) gx = mt;
      if (gx < 1) gx = 1;
      p.tA_grid_x = (int)gx;
      p.PS = p.tA_grid_x;
      p.tB_tiles = (int)(((a.S + 31) / 32) * ((a.Dw + 127) / 128));
      int64_t ks = ((int64_t)sms * 2 + p.tB_tiles - 1) / p.tB_tiles;
      int64_t kmax = (a.N + 255) / 256;  // at least 256 rows per slot
      if (ks > kmax) ks = kmax;
      if (ks < 1) ks = 1;
      p.tB_kslots = (int)ks;
      p.tB_rows_per_slot = (a.N + ks - 1) / ks;
      p.tB_rows_per_slot = (p.tB_rows_per_slot + SG_BK - 1) / SG_BK * SG_BK;
      p.PG = p.tB_kslots;
      break;
    }
    case KERNEL_TC_BF16: {
      int rc = tc_plan(a, p);
      if (rc) return rc;
      p.PG = p.PS = p.tc_ctas;
      break;
    }
    case KERNEL_TC_GEMM: {
      int rc = gemm_plan(a, p);
      if (rc) return rc;
      p.PG = p.gm_splits;
      p.PS = 4 * p.gm_grid1;
      break;
    }
  }
  p.fr_tc = (a.vi_type == BJX_VI_FULL_RANK && a.precision != BJX_PREC_FP32 && fullrank_tc_eligible(a)) ? 1 : 0;

  // workspace carve-up (256-byte aligned regions)
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t SD = (size_t)a.S * a.D * sizeof(float);
  p.o_eps = take(SD);
  p.o_z = take(a.vi_type != BJX_VI_MEAN_FIELD ? SD : 0);
  p.o_theta = take(SD);
  p.o_bprime = take(SD);
  p.o_gprior = take(SD);
  p.o_gz = take(a.vi_type != BJX_VI_MEAN_FIELD ? SD : 0);
  p.o_qp = take((size_t)p.n_qp_blocks * sizeof(float));
  p.o_aux = take((size_t)a.S * sizeof(float4));
  p.o_G = take((size_t)a.S * Dw * sizeof(float));
  p.o_stat = take((size_t)a.S * sizeof(float));
  p.o_Gpart = take((size_t)p.PG * a.S * Dw * sizeof(float));
  p.o_statpart = take((size_t)p.PS * a.S * sizeof(float));
  p.o_gbuf = take(p.kernel == KERNEL_FP32_TILED ? (size_t)a.N * a.S * sizeof(float) : 0);
  p.o_tc_theta = take(p.kernel == KERNEL_TC_BF16 ? tc_theta_bytes(a, p) : 0);
  const bool gm = p.kernel == KERNEL_TC_GEMM;
  p.o_gm_x = take(gm && (a.x_planes == nullptr || p.gm_nt == 3) ? x_planes_bytes(a.N, a.Dw) / 2 * p.gm_nt : 0);
  p.o_gm_t = take(gm ? (size_t)p.gm_nt * a.S * p.gm_Dp * 2 : 0);
  p.o_gm_g = take(gm ? (size_t)p.gm_nt * a.N * p.gm_Sp * 2 : 0);
  p.o_fr_tc = take(p.fr_tc ? fullrank_tc_bytes(a) : 0);
  p.o_lr = take(a.vi_type == BJX_VI_LOW_RANK ? lowrank_bytes(a) : 0);
  p.o_stage_key = take(16);
  p.o_stage_params = take((size_t)(a.D + p.Pn + a.K) * sizeof(float));
  p.o_stage_out = take((size_t)(1 + a.D + p.Pn + a.K) * sizeof(float));
  p.total = off;
  return BJX_OK;
}

// ---- prologue -----------------------------------------------------------------------------------------------------------
// One thread per (s, d).  Mean field: draws eps, forms z = loc + sigma * eps.  Full rank: eps and z were produced by
// k_eps / k_fullrank_z.  Then theta = b(z), and the per-coordinate pieces of q.log_prob - log p~ with d/dz.
template <bool FULL>
__global__ void __launch_bounds__(256) k_prologue(const uint32_t* __restrict__ key, const int64_t* __restrict__ key_index,
                                                  const float* __restrict__ loc,
                                                  const float* __restrict__ raw, const BjxCoord* __restrict__ coords,
                                                  int64_t S, int64_t D, int scale_bij, int layout,
                                                  float* __restrict__ eps_buf, const float* __restrict__ z_buf,
                                                  float* __restrict__ theta, float* __restrict__ bprime,
                                                  float* __restrict__ gprior, float* __restrict__ qp_part,
                                                  float* __restrict__ eps_out, int64_t probs_coord,
                                                  float4* __restrict__ probs_aux, int64_t s_offset, int64_t S_total,
                                                  int point_mode) {
  __shared__ float red[32];
  const int64_t total = S * D;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  float term = 0.0f;
  if (i < total) {
    const int64_t d = i % D;
    float eps, z, logsig;
    if (FULL) {
      eps = eps_buf[i];
      z = z_buf[i];
      if (raw != nullptr) {
        const ScaleEval sc = eval_scale(raw[d * D + d], scale_bij);
        logsig = logf(fabsf(sc.sigma));
      } else {
        logsig = 0.0f;  // low rank: log|det(diag(d) + U U^T)| is not a sum over coordinates; k_scalars adds it per sample
      }
    } else {
      const uint32_t* k = key_index ? key + 2 * (*key_index) : key;  // device-resident loop: key = keys[step counter]
      // sample sharding: the flat index and the stream length are those of the unsharded (S_total, D) draw
      const uint32_t bits = random_bits_at(k[0], k[1], (uint64_t)(i + s_offset * D), (uint64_t)(S_total * D), layout);
      eps = point_mode ? 0.0f : bits_to_normal(bits);
      eps_buf[i] = eps;
      const float rho = raw[d];
      const ScaleEval sc = eval_scale(rho, scale_bij);
      z = point_mode ? loc[d] : loc[d] + sc.sigma * eps;
      logsig = scale_bij == BJX_SCALE_EXP ? rho : logf(sc.sigma);
    }
    if (eps_out) eps_out[i] = eps;
    CoordEval ce;
    if (point_mode == BJX_POINT_LOGLIK) {
      // loc IS the constrained value: no bijector, no prior; derivatives are with respect to theta itself
      ce.theta = z;
      ce.dtheta = 1.0f;
      ce.logp = ce.dlogp = ce.ldj = ce.dldj = 0.0f;
      ce.lt = logf(z);
      ce.l1t = log1pf(-z);
      ce.dlt = 1.0f / z;
      ce.dl1t = -1.0f / (1.0f - z);
    } else {
      ce = eval_coord(z, coords[d]);
    }
    theta[i] = ce.theta;
    bprime[i] = ce.dtheta;
    gprior[i] = point_mode == BJX_POINT_MAP ? -(ce.dlogp - ce.dldj) : -ce.dlogp;
    if (d == probs_coord) probs_aux[i / D] = make_float4(ce.lt, ce.l1t, ce.dlt, ce.dl1t);
    // q.log_prob per coordinate (z - loc)/sigma == eps, minus the unconstrained prior's log density
    term = (-0.5f * eps * eps - logsig - BJX_HALF_LOG_2PI) - ce.logp;
    if (point_mode == BJX_POINT_LOG_JOINT) term = -ce.logp;               // no variational term
    if (point_mode == BJX_POINT_MAP) term = -(ce.logp - ce.ldj);          // prior density of theta = b(z), map.py:30
    if (point_mode == BJX_POINT_LOGLIK) term = 0.0f;                      // the likelihood alone
  }
  const float t = block_sum(term, red);
  if (threadIdx.x == 0) qp_part[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) k_eps(const uint32_t* __restrict__ key, const int64_t* __restrict__ key_index,
                                             int64_t total, int64_t offset, int64_t global_total, int layout,
                                             float* __restrict__ eps) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const uint32_t* k = key_index ? key + 2 * (*key_index) : key;
  if (i < total) eps[i] = bits_to_normal(random_bits_at(k[0], k[1], (uint64_t)(i + offset), (uint64_t)global_total, layout));
}

// ---- full-rank contractions (fp32 FFMA; S x D x D) -----------------------------------------------------------------
struct LoadEps {  // A(m = s, k = j) = eps[s, j]
  static constexpr bool k_fast = true;
  const float* eps;
  int64_t S, D;
  __device__ __forceinline__ float operator()(int64_t s, int64_t j) const { return (s < S && j < D) ? eps[s * D + j] : 0.0f; }
};
struct LoadTril {  // B(n = i, k = j) = scale(M[i, j]) for j <= i, else 0   (MultivariateNormalTriL masks with band_part)
  static constexpr bool k_fast = true;
  const float* M;
  int64_t D;
  int scale_bij;
  __device__ __forceinline__ float operator()(int64_t i, int64_t j) const {
    return (i < D && j <= i) ? eval_scale(M[i * D + j], scale_bij).sigma : 0.0f;
  }
};
struct LoadColMajor {  // A(m, k) = base[k * D + m]
  static constexpr bool k_fast = false;
  const float* base;
  int64_t rows, D, kmax;
  __device__ __forceinline__ float operator()(int64_t m, int64_t k) const { return (m < rows && k < kmax) ? base[k * D + m] : 0.0f; }
};

// z[s, i] = loc[i] + sum_{j <= i} L[i, j] eps[s, j]
__global__ void __launch_bounds__(256) k_fullrank_z(const float* __restrict__ eps, const float* __restrict__ M,
                                                    const float* __restrict__ loc, int64_t S, int64_t D, int scale_bij,
                                                    float* __restrict__ z) {
  constexpr int BM = 64, BN = 64;
  __shared__ SgemmSmem<BM, BN> sm;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  float acc[4][4];
  const int64_t kend = min(D, n0 + BN);  // columns j > i contribute nothing
  sgemm_tile<BM, BN>(acc, LoadEps{eps, S, D}, LoadTril{M, D, scale_bij}, m0, n0, 0, kend, sm);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t s = m0 + ty + 16 * i;
    if (s >= S) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = n0 + tx + 16 * j;
      if (c < D) z[s * D + c] = loc[c] + acc[i][j];
    }
  }
}

// d raw[i, j] = [i >= j] * ( sum_s gz[s, i] eps[s, j] * scale'(raw_ij) ) - [i == j] * prior_weight * dlog|L_ii|/draw_ii
__global__ void __launch_bounds__(256) k_fullrank_dM(const float* __restrict__ gz, const float* __restrict__ eps,
                                                     const float* __restrict__ M, int64_t S, int64_t D, int scale_bij,
                                                     float prior_weight, float* __restrict__ dM) {
  constexpr int BM = 64, BN = 64;
  __shared__ SgemmSmem<BM, BN> sm;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  float acc[4][4];
  const bool live = n0 <= m0 + BM - 1;  // tile touches the lower triangle
  if (live) {
    sgemm_tile<BM, BN>(acc, LoadColMajor{gz, D, D, S}, LoadColMajor{eps, D, D, S}, m0, n0, 0, S, sm);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = m0 + ty + 16 * i;
    if (r >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = n0 + tx + 16 * j;
      if (c >= D) continue;
      float v = 0.0f;
      if (c <= r) {
        const ScaleEval sc = eval_scale(M[r * D + c], scale_bij);
        v = acc[i][j] * sc.dsigma;
        if (c == r) v -= prior_weight * sc.dlogsigma;
      }
      dM[r * D + c] = v;
    }
  }
}

// ---- coin toss (tfd.Bernoulli(probs=p).log_prob(y).sum(), README.md:91-95): one block per sample ---------------------
// aux[s] = (log p, log(1-p), d log p / dz, d log(1-p) / dz) from the prologue; G[s] is d ll_s / d z (not d/d theta).
// The coin-toss log-likelihood is LINEAR in the labels: sum_n [(1 - y_n) A + y_n B] = (N - sum y) A + (sum y) B, for A, B the
// per-sample log(1 - theta), log(theta) (or their z-derivatives).  Summing the N rounded products instead adds the SAME
// two values thousands of times, and the rounding of such a sum is not random: it drifts systematically (4e-6 relative at
// N = 4096, which the near-cancellation k - N*theta of the gradient amplifies past rtol 1e-5).  So: sum y once, in double.
__device__ __forceinline__ double block_sum_f64(double v, double* red /* [8] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];  // fixed order: every thread gets the same bits
  return t;
}

__global__ void __launch_bounds__(256) k_lik_coin(const float* __restrict__ y, int64_t N, const float4* __restrict__ aux,
                                                  float* __restrict__ G, float* __restrict__ stat) {
  __shared__ double red[8];
  const int s = blockIdx.x;
  double sy = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += 256) sy += (double)y[n];
  sy = block_sum_f64(sy, red);
  if (threadIdx.x == 0) {
    const float4 a = aux[s];
    const double k0 = (double)N - sy;
    stat[s] = (float)(k0 * (double)a.y + sy * (double)a.x);
    G[s] = (float)(k0 * (double)a.w + sy * (double)a.z);
  }
}

// ---- tail ---------------------------------------------------------------------------------------------------------------
template <int W>  // warps per block: 8, or 32 when there are few outputs and many slots (c2: 384 outputs x 592 slots)
__global__ void __launch_bounds__(W * 32) k_reduce_partials_w(const float* __restrict__ Gpart, int PG, const float* __restrict__ statpart,
                                                              int PS, int64_t SDw, int64_t S, float* __restrict__ G,
                                                              float* __restrict__ stat) {
  // block = 32 consecutive outputs (coalesced) x W warps striding over the partial slots; the per-warp sums are
  // combined in a fixed order through shared memory, so the result does not depend on timing
  __shared__ float red[W][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const float* src = nullptr;
  int P = 0;
  int64_t stride = 0;
  float* dst = nullptr;
  if (i < SDw) {
    src = Gpart + i; P = PG; stride = SDw; dst = G + i;
  } else if (i < SDw + S) {
    src = statpart + (i - SDw); P = PS; stride = S; dst = stat + (i - SDw);
  }
  // four independent chains per warp: the loop is a chain of L2 latencies (148 slots at c3, 592 at c2), not of bytes
  float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
  if (src != nullptr) {
    int p = w;
    for (; p + 3 * W < P; p += 4 * W) {
      t0 += src[(size_t)p * stride];
      t1 += src[(size_t)(p + W) * stride];
      t2 += src[(size_t)(p + 2 * W) * stride];
      t3 += src[(size_t)(p + 3 * W) * stride];
    }
    for (; p < P; p += W) t0 += src[(size_t)p * stride];
  }
  red[w][lane] = (t0 + t1) + (t2 + t3);
  __syncthreads();
  if (w == 0 && dst != nullptr) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < W; ++k) t += red[k][lane];
    *dst = t;
  }
}
// picks the block width: W = 32 when the grid would not even fill half of the SMs and the slots are many
static void launch_reduce_partials(const float* Gpart, int PG, const float* statpart, int PS, int64_t SDw, int64_t S, float* G, float* stat,
                                   cudaStream_t st) {
  const unsigned blocks = (unsigned)((SDw + S + 31) / 32);
  if (blocks < 64 && (PG > 64 || PS > 64))
    k_reduce_partials_w<32><<<blocks, 1024, 0, st>>>(Gpart, PG, statpart, PS, SDw, S, G, stat);
  else
    k_reduce_partials_w<8><<<blocks, 256, 0, st>>>(Gpart, PG, statpart, PS, SDw, S, G, stat);
}

// Few partial slots and many outputs (the two-pass GEMM path: 9 K splits x S x Dw, 0.5 M outputs at c4): one thread per output
// walks the slots in order -- eight warps striding over nine slots leave most of a block idle (30 us at c4 against 6 us here).
// The stat partials of that path have many slots (4 per pass-1 CTA) and few outputs: they keep the block-per-32-outputs kernel.
__global__ void __launch_bounds__(256) k_reduce_partials_few(const float* __restrict__ Gpart, int PG, int64_t SDw, float* __restrict__ G) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= SDw) return;
  float t = 0.0f;
  for (int p = 0; p < PG; ++p) t += Gpart[(size_t)p * SDw + i];
  G[i] = t;
}

struct TailArgs {
  int64_t S, D, Dw, w_off, N;
  int vi_type, scale_bij, likelihood, noise_src, noise_index;
  float noise_value, ll_scale, prior_weight;
  int point_mode;        // BJX_POINT_*: d raw_scale is zero
  float entropy_weight;  // weight of the per-step constant -d log sigma / d raw (once per step across sample shards)
  int64_t S_total;       // samples of the whole step (means divide by this)
};

__device__ __forceinline__ float tau_of(const TailArgs& t, const float* kwargs, const float* theta, int64_t s) {
  if (t.noise_src == BJX_NOISE_KWARG) return expf(kwargs[t.noise_index]);
  if (t.noise_src == BJX_NOISE_LATENT) return theta[s * t.D + t.noise_index];
  return t.noise_value;
}

// loss = prior_weight * mean_s(q_s - p~_s) - ll_scale * mean_s ll_s ;  d kwargs          (one block of 256 threads)
__device__ __forceinline__ void scalars_body(const TailArgs& t, int64_t K, int64_t Pn, const float* __restrict__ kwargs,
                                             const float* __restrict__ theta, const float* __restrict__ qp_part, int n_qp,
                                             const float* __restrict__ stat, const float* __restrict__ logdet,
                                             float* __restrict__ out) {
  __shared__ float red[32];
  float qp = 0.0f;
  for (int i = threadIdx.x; i < n_qp; i += 256) qp += qp_part[i];
  qp = block_sum(qp, red);
  float ll = 0.0f, dk = 0.0f;
  for (int64_t s = threadIdx.x; s < t.S; s += 256) {
    if (t.likelihood == BJX_LIK_GAUSSIAN) {
      const float tau = tau_of(t, kwargs, theta, s);
      const float q = stat[s] / (tau * tau);
      ll += -0.5f * q - (float)t.N * (logf(tau) + BJX_HALF_LOG_2PI);
      dk += q - (float)t.N;  // d ll_s / d log tau
    } else {
      ll += stat[s];
    }
  }
  ll = block_sum(ll, red);
  dk = block_sum(dk, red);
  if (threadIdx.x == 0) {
    const float invS = 1.0f / (float)t.S_total;
    // low rank: q.log_prob carries -log|det A| once per sample of this call
    const float qextra = logdet ? -logdet[0] * (float)t.S : 0.0f;
    out[0] = t.prior_weight * (qp + qextra) * invS - t.ll_scale * ll * invS;
    float* dkw = out + 1 + t.D + Pn;
    for (int64_t k = 0; k < K; ++k) dkw[k] = 0.0f;
    if (t.likelihood == BJX_LIK_GAUSSIAN && t.noise_src == BJX_NOISE_KWARG) dkw[t.noise_index] = -t.ll_scale * dk * invS;
  }
}
__global__ void __launch_bounds__(256) k_scalars(TailArgs t, int64_t K, int64_t Pn, const float* __restrict__ kwargs,
                                                 const float* __restrict__ theta, const float* __restrict__ qp_part,
                                                 int n_qp, const float* __restrict__ stat, const float* __restrict__ logdet,
                                                 float* __restrict__ out) {
  scalars_body(t, K, Pn, kwargs, theta, qp_part, n_qp, stat, logdet, out);
}

// One WARP per latent coordinate d; lanes stride over the S samples, fixed-shape shuffle reduction (deterministic).
//   gz[s,d] = prior_weight * (-d log p~ / dz) / S  -  (ll_scale / S) * dll_s/dtheta_d * b'(z)
//   d loc[d] = sum_s gz;   d rho[d] = sigma'(rho) * sum_s gz * eps  -  prior_weight * dlog sigma / d rho     (mean field)
__global__ void __launch_bounds__(256) k_tail(TailArgs t, const float* __restrict__ raw, const float* __restrict__ kwargs,
                                              const float* __restrict__ eps, const float* __restrict__ theta,
                                              const float* __restrict__ bprime, const float* __restrict__ gprior,
                                              const float* __restrict__ G, const float* __restrict__ stat,
                                              float* __restrict__ gz_out, float* __restrict__ out,
                                              int fuse_scalars, int64_t K, int64_t Pn, const float* __restrict__ qp_part, int n_qp) {
  // fuse_scalars (mean field): one extra block at the end of the grid computes the loss and d kwargs -- they depend on the
  // same reduced partials as this kernel, not on its results -- which saves the k_scalars launch
  if (fuse_scalars && blockIdx.x == gridDim.x - 1) {
    scalars_body(t, K, Pn, kwargs, theta, qp_part, n_qp, stat, nullptr, out);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int64_t d = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (d >= t.D) return;
  const float invS = 1.0f / (float)t.S_total;
  const float cS = t.ll_scale * invS;
  const bool is_w = t.likelihood == BJX_LIK_BERNOULLI_PROBS ? (d == t.w_off) : (d >= t.w_off && d < t.w_off + t.Dw);
  const bool is_noise = t.likelihood == BJX_LIK_GAUSSIAN && t.noise_src == BJX_NOISE_LATENT && d == t.noise_index;
  const int64_t Dw = t.likelihood == BJX_LIK_BERNOULLI_PROBS ? 1 : t.Dw;
  float acc_mu = 0.0f, acc_rho = 0.0f;
  for (int64_t s = lane; s < t.S; s += 32) {
    const int64_t i = s * t.D + d;
    float dll = 0.0f;  // d ll_s / d theta[s, d]
    if (is_w) {
      dll = G[s * Dw + (d - t.w_off)];
      if (t.likelihood == BJX_LIK_GAUSSIAN) {
        const float tau = tau_of(t, kwargs, theta, s);
        dll = dll / (tau * tau);
      }
    } else if (is_noise) {
      const float tau = theta[i];
      dll = (stat[s] / (tau * tau) - (float)t.N) / tau;  // d/dtau [ -Q/(2 tau^2) - N log tau ]
    }
    // the coin kernel already differentiated with respect to z
    const float bp = t.likelihood == BJX_LIK_BERNOULLI_PROBS ? 1.0f : bprime[i];
    const float gz = t.prior_weight * gprior[i] * invS - cS * dll * bp;
    if (gz_out) gz_out[i] = gz;
    acc_mu += gz;
    acc_rho = fmaf(gz, eps[i], acc_rho);
  }
  acc_mu = warp_sum(acc_mu);
  acc_rho = warp_sum(acc_rho);
  if (lane == 0) {
    out[1 + d] = acc_mu;
    if (t.vi_type == BJX_VI_MEAN_FIELD) {
      const ScaleEval sc = eval_scale(raw[d], t.scale_bij);
      out[1 + t.D + d] = t.point_mode ? 0.0f : acc_rho * sc.dsigma - t.entropy_weight * sc.dlogsigma;
    }
  }
}

// ---- optax.adam on one element (eps outside the square root, bias-corrected) ------------------------------------------------
// Hyper-parameters arrive as doubles and are rounded ONCE: (1 - b) is the float nearest to the exact complement (1.0f - 0.999f
// is 1.3e-5 off 0.001), and the bias corrections 1 - b^t come from expm1(t log b) in double (1 - powf(b, t) loses 1e-5
// relative while b^t is close to 1, i.e. during the first steps).
struct AdamHyper {
  float lr, b1, omb1, b2, omb2, eps;
  double lb1, lb2;  // log b1, log b2
};
static AdamHyper adam_hyper(double lr, double b1, double b2, double eps) {
  return AdamHyper{(float)lr, (float)b1, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)eps, log(b1), log(b2)};
}
__device__ __forceinline__ void adam_bias(const AdamHyper& h, int64_t t, float& c1, float& c2) {  // t = 1-based step number
  c1 = (float)(-expm1((double)t * h.lb1));
  c2 = (float)(-expm1((double)t * h.lb2));
}
__device__ __forceinline__ void adam_one(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float gi, int64_t i,
                                         const AdamHyper& h, float c1, float c2) {
  const float mi = h.b1 * m[i] + h.omb1 * gi;
  const float vi = h.b2 * v[i] + h.omb2 * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= h.lr * (mi / c1) / (sqrtf(vi / c2) + h.eps);
}

// the rest of a training step (Adam, loss record, step counter) for kernels that can finish it themselves
struct TrainTail {
  float *params, *m, *v;  // params == nullptr: not a training step
  AdamHyper h;
  int64_t* steps_done;
  float* losses;
  int64_t n;
};

// ---- tiny problems: the whole step in ONE block ---------------------------------------------------------------------------
// Config c1 (coin toss: D = 1, S = 16, N = 100) is pure launch latency -- four launches of a few hundred threads.  This
// kernel runs the same four stages (k_prologue -> k_lik_coin -> k_tail -> k_scalars, same formulas) back to back in one
// block with shared memory in place of the global workspace.  Mean field, tfd.Bernoulli(probs = theta) likelihood,
// S * D <= 1024, S <= 256, N <= 4096.

__global__ void __launch_bounds__(256) k_tiny_coin_step(const uint32_t* __restrict__ key, const int64_t* __restrict__ key_index,
                                                        const float* __restrict__ loc, const float* __restrict__ raw,
                                                        const BjxCoord* __restrict__ coords, int S, int D, int scale_bij,
                                                        int layout, const float* __restrict__ y, int N, int probs_coord,
                                                        float prior_weight, float entropy_weight, float ll_scale,
                                                        int64_t s_offset, int64_t S_total, int64_t K,
                                                        float* __restrict__ eps_out, float* __restrict__ out, TrainTail tt) {
  __shared__ float s_eps[TINY_SD], s_gprior[TINY_SD];
  __shared__ float4 s_aux[TINY_S];
  __shared__ float s_G[TINY_S], s_stat[TINY_S];
  __shared__ float red[32];
  __shared__ double s_red64[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = S * D;
  const uint32_t* k = key_index ? key + 2 * (*key_index) : key;
  // ---- stage 1 (k_prologue): eps, z, bijector, prior and entropy terms
  float term = 0.0f;
  for (int i = tid; i < total; i += 256) {
    const int d = i % D;
    const uint32_t bits = random_bits_at(k[0], k[1], (uint64_t)(i + s_offset * D), (uint64_t)(S_total * D), layout);
    const float eps = bits_to_normal(bits);
    const float rho = raw[d];
    const ScaleEval sc = eval_scale(rho, scale_bij);
    const float z = loc[d] + sc.sigma * eps;
    const float logsig = scale_bij == BJX_SCALE_EXP ? rho : logf(sc.sigma);
    const CoordEval ce = eval_coord(z, coords[d]);
    s_eps[i] = eps;
    s_gprior[i] = -ce.dlogp;
    if (eps_out) eps_out[i] = eps;
    if (d == probs_coord) s_aux[i / D] = make_float4(ce.lt, ce.l1t, ce.dlt, ce.dl1t);
    term += (-0.5f * eps * eps - logsig - BJX_HALF_LOG_2PI) - ce.logp;
  }
  const float qp_sum = block_sum(term, red);  // valid in thread 0
  __syncthreads();
  // ---- stage 2 (k_lik_coin): the labels are summed ONCE (the log-likelihood is linear in them, see k_lik_coin)
  double sy = 0.0;
  for (int n = tid; n < N; n += 256) sy += (double)y[n];
  sy = block_sum_f64(sy, s_red64);
  for (int s = tid; s < S; s += 256) {
    const float4 a = s_aux[s];
    const double k0 = (double)N - sy;
    s_stat[s] = N > 0 ? (float)(k0 * (double)a.y + sy * (double)a.x) : 0.0f;
    s_G[s] = N > 0 ? (float)(k0 * (double)a.w + sy * (double)a.z) : 0.0f;
  }
  __syncthreads();
  // ---- stage 3 (k_tail): reparameterisation gradient, one warp per latent coordinate
  const float invS = 1.0f / (float)S_total;
  const float cS = ll_scale * invS;
  for (int d = warp; d < D; d += 8) {
    float acc_mu = 0.0f, acc_rho = 0.0f;
    for (int s2 = lane; s2 < S; s2 += 32) {
      const int i = s2 * D + d;
      const float dll = d == probs_coord ? s_G[s2] : 0.0f;  // already differentiated with respect to z
      const float gz = prior_weight * s_gprior[i] * invS - cS * dll;
      acc_mu += gz;
      acc_rho = fmaf(gz, s_eps[i], acc_rho);
    }
    acc_mu = warp_sum(acc_mu);
    acc_rho = warp_sum(acc_rho);
    if (lane == 0) {
      const ScaleEval sc = eval_scale(raw[d], scale_bij);
      out[1 + d] = acc_mu;
      out[1 + D + d] = acc_rho * sc.dsigma - entropy_weight * sc.dlogsigma;
    }
  }
  // ---- stage 4 (k_scalars): loss and (zero) kwarg gradients
  float ll = 0.0f;
  for (int s2 = tid; s2 < S; s2 += 256) ll += s_stat[s2];
  ll = block_sum(ll, red);
  if (tid == 0) {
    out[0] = prior_weight * qp_sum * invS - ll_scale * ll * invS;
    for (int64_t kk = 0; kk < K; ++kk) out[1 + 2 * D + kk] = 0.0f;
  }
  // ---- training loop: Adam on [loc | raw | kwargs], loss record, ++step -- the whole epoch is this one launch
  if (tt.params != nullptr) {
    __syncthreads();  // out[] written above by this block
    const int64_t t0 = *tt.steps_done;
    float c1, c2;
    adam_bias(tt.h, t0 + 1, c1, c2);
    for (int64_t i = tid; i < tt.n; i += 256) adam_one(tt.params, tt.m, tt.v, out[1 + i], i, tt.h, c1, c2);
    __syncthreads();  // everybody has read the step number (and the key it selects)
    if (tid == 0) {
      if (tt.losses) tt.losses[t0] = out[0];
      *tt.steps_done = t0 + 1;
    }
  }
}

// ---- orchestration ------------------------------------------------------------------------------------------------------
static int run_prologue(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st) {
  float* eps = (float*)(ws + p.o_eps);
  float* z = (float*)(ws + p.o_z);
  float* theta = (float*)(ws + p.o_theta);
  float* bprime = (float*)(ws + p.o_bprime);
  float* gprior = (float*)(ws + p.o_gprior);
  float* qp = (float*)(ws + p.o_qp);
  float4* aux = (float4*)(ws + p.o_aux);
  const int64_t probs_coord = (a.likelihood == BJX_LIK_BERNOULLI_PROBS && a.N > 0) ? a.w_offset : -1;
  const int64_t total = a.S * a.D;
  const int64_t S_total = a.S_total ? a.S_total : a.S, s_offset = a.S_total ? a.s_offset : 0;
  if (a.vi_type == BJX_VI_LOW_RANK) {
    k_eps<<<p.n_qp_blocks, 256, 0, st>>>(a.key, a.key_index, total, s_offset * a.D, S_total * a.D, a.threefry_layout, eps);
    BJX_LAUNCH_CHECK("k_eps");
    int rc = launch_lowrank_sample(a, eps, z, ws + p.o_lr, st);
    if (rc) return rc;
    k_prologue<true><<<p.n_qp_blocks, 256, 0, st>>>(a.key, a.key_index, a.loc, nullptr, a.coords, a.S, a.D, a.scale_bijector,
                                                    a.threefry_layout, eps, z, theta, bprime, gprior, qp, a.eps_out,
                                                    probs_coord, aux, s_offset, S_total, a.point_mode);
  } else if (a.vi_type == BJX_VI_FULL_RANK) {
    k_eps<<<p.n_qp_blocks, 256, 0, st>>>(a.key, a.key_index, total, s_offset * a.D, S_total * a.D, a.threefry_layout, eps);
    BJX_LAUNCH_CHECK("k_eps");
    if (p.fr_tc) {
      int rc = launch_fullrank_z_tc(a, eps, z, ws + p.o_fr_tc, st);
      if (rc) return rc;
    } else {
      dim3 grid((unsigned)((a.D + 63) / 64), (unsigned)((a.S + 63) / 64));
      k_fullrank_z<<<grid, 256, 0, st>>>(eps, a.raw_scale, a.loc, a.S, a.D, a.scale_bijector, z);
      BJX_LAUNCH_CHECK("k_fullrank_z");
    }
    k_prologue<true><<<p.n_qp_blocks, 256, 0, st>>>(a.key, a.key_index, a.loc, a.raw_scale, a.coords, a.S, a.D, a.scale_bijector,
                                                    a.threefry_layout, eps, z, theta, bprime, gprior, qp, a.eps_out,
                                                    probs_coord, aux, s_offset, S_total, a.point_mode);
  } else {
    k_prologue<false><<<p.n_qp_blocks, 256, 0, st>>>(a.key, a.key_index, a.loc, a.raw_scale, a.coords, a.S, a.D, a.scale_bijector,
                                                     a.threefry_layout, eps, nullptr, theta, bprime, gprior, qp, a.eps_out,
                                                     probs_coord, aux, s_offset, S_total, a.point_mode);
  }
  BJX_LAUNCH_CHECK("k_prologue");
  return BJX_OK;
}

static int check_ptrs(const BjxElboArgs& a, const Plan& p) {
  BJX_REQUIRE(a.key && a.loc && a.raw_scale && a.coords && a.out, "null key/loc/raw_scale/coords/out pointer");
  BJX_REQUIRE(a.K == 0 || a.kwargs, "K > 0 but kwargs is null");
  BJX_REQUIRE(a.N == 0 || a.y, "y is null");
  // X may be absent when the dataset lives on the device only as its prepared planes and the planned kernel reads nothing else
  const bool planes_only_ok = a.x_planes != nullptr && (p.kernel == KERNEL_TC_BF16 || (p.kernel == KERNEL_TC_GEMM && p.gm_nt == 2));
  BJX_REQUIRE(a.likelihood == BJX_LIK_BERNOULLI_PROBS || a.N == 0 || a.X || planes_only_ok,
              "X is null (allowed only with x_planes on the tensor-core kernels that read the planes)");
  BJX_REQUIRE(a.workspace != nullptr, "workspace is null");
  if (a.workspace_bytes < p.total)
    return fail(BJX_ERR_WORKSPACE, "workspace too small: have %zu bytes, need %zu", a.workspace_bytes, p.total);
  return BJX_OK;
}

// tt (training loops only): kernels that can finish the epoch themselves do so and set *tt_done
static int run_step(const BjxElboArgs& a, const Plan& p, cudaStream_t st, const TrainTail* tt = nullptr, bool* tt_done = nullptr) {
  char* ws = (char*)a.workspace;
  if (p.tiny) {  // the whole step in one block (launch latency is all there is at this size)
    const int probs_coord = a.N > 0 ? (int)a.w_offset : -1;
    k_tiny_coin_step<<<1, 256, 0, st>>>(a.key, a.key_index, a.loc, a.raw_scale, a.coords, (int)a.S, (int)a.D, a.scale_bijector,
                                        a.threefry_layout, a.y, (int)a.N, probs_coord, a.prior_weight,
                                        a.S_total ? a.entropy_weight : a.prior_weight, a.ll_scale, a.S_total ? a.s_offset : 0,
                                        a.S_total ? a.S_total : a.S, a.K, a.eps_out, a.out, tt ? *tt : TrainTail{});
    BJX_LAUNCH_CHECK("k_tiny_coin_step");
    if (tt && tt_done) *tt_done = true;
    return BJX_OK;
  }
  int rc = run_prologue(a, p, ws, st);
  if (rc) return rc;

  float* G = (float*)(ws + p.o_G);
  float* stat = (float*)(ws + p.o_stat);
  const int64_t Dw = a.likelihood == BJX_LIK_BERNOULLI_PROBS ? 1 : a.Dw;
  if (p.kernel == KERNEL_COIN) {
    k_lik_coin<<<(unsigned)a.S, 256, 0, st>>>(a.y, a.N, (const float4*)(ws + p.o_aux), G, stat);
    BJX_LAUNCH_CHECK("k_lik_coin");
  } else {
    if (a.N == 0) {
      BJX_CUDA_TRY(cudaMemsetAsync(ws + p.o_Gpart, 0, (size_t)p.PG * a.S * Dw * sizeof(float), st));
      BJX_CUDA_TRY(cudaMemsetAsync(ws + p.o_statpart, 0, (size_t)p.PS * a.S * sizeof(float), st));
    } else {
      prof_mark(st, 0);
      if (p.kernel == KERNEL_FP32_SMALLD) {
        rc = launch_lik_smalld(a, p, ws, st);
      } else if (p.kernel == KERNEL_FP32_TILED) {
        rc = launch_lik_tiled(a, p, ws, st);
      } else if (p.kernel == KERNEL_TC_GEMM) {
        rc = launch_lik_gemm(a, p, ws, st);
      } else {
        // the fused single pass takes 32 samples at a time: one launch per slice, each with its own block of the partials
        for (int64_t s0 = 0; s0 < a.S && rc == BJX_OK; s0 += 32) rc = launch_lik_tc(a, p, ws, st, s0, a.S - s0 < 32 ? a.S - s0 : 32);
      }
      prof_mark(st, 1);
    }
    if (rc) return rc;
    if (p.kernel == KERNEL_TC_BF16 && a.S > 32 && a.N > 0) {
      for (int64_t s0 = 0; s0 < a.S; s0 += 32) {  // partials are laid out [slice][CTA][S_slice, Dw]
        const int64_t Ss = a.S - s0 < 32 ? a.S - s0 : 32;
        launch_reduce_partials((const float*)(ws + p.o_Gpart) + (size_t)p.PG * s0 * Dw, p.PG,
                               (const float*)(ws + p.o_statpart) + (size_t)p.PS * s0, p.PS, Ss * Dw, Ss, G + s0 * Dw, stat + s0, st);
        BJX_LAUNCH_CHECK("k_reduce_partials");
      }
    } else {
      const int64_t SDw = a.S * Dw;
      if (p.PG <= 16 && SDw >= 65536) {
        k_reduce_partials_few<<<(unsigned)((SDw + 255) / 256), 256, 0, st>>>((const float*)(ws + p.o_Gpart), p.PG, SDw, G);
        // the stat partials alone: outputs [SDw, SDw + S) of the general kernel, i.e. an empty G range
        launch_reduce_partials(nullptr, 0, (const float*)(ws + p.o_statpart), p.PS, 0, a.S, G, stat, st);
      } else {
        launch_reduce_partials((const float*)(ws + p.o_Gpart), p.PG, (const float*)(ws + p.o_statpart), p.PS, SDw, a.S, G, stat, st);
      }
      BJX_LAUNCH_CHECK("k_reduce_partials");
    }
  }

  TailArgs t;
  t.S = a.S; t.D = a.D; t.Dw = a.Dw; t.w_off = a.w_offset; t.N = a.N;
  t.vi_type = a.vi_type; t.scale_bij = a.scale_bijector; t.likelihood = a.likelihood;
  t.noise_src = a.noise_src; t.noise_index = a.noise_index; t.noise_value = a.noise_value;
  t.ll_scale = a.ll_scale; t.prior_weight = a.prior_weight;
  t.entropy_weight = a.point_mode ? 0.0f : (a.S_total ? a.entropy_weight : a.prior_weight);
  t.point_mode = a.point_mode;
  t.S_total = a.S_total ? a.S_total : a.S;
  float* gz = a.vi_type != BJX_VI_MEAN_FIELD ? (float*)(ws + p.o_gz) : nullptr;
  const int fuse_scalars = a.vi_type == BJX_VI_MEAN_FIELD ? 1 : 0;
  k_tail<<<(unsigned)((a.D + 7) / 8) + fuse_scalars, 256, 0, st>>>(t, a.raw_scale, a.kwargs, (const float*)(ws + p.o_eps),
                                                                      (const float*)(ws + p.o_theta), (const float*)(ws + p.o_bprime),
                                                                      (const float*)(ws + p.o_gprior), G, stat, gz, a.out, fuse_scalars,
                                                                      a.K, p.Pn, (const float*)(ws + p.o_qp), p.n_qp_blocks);
  BJX_LAUNCH_CHECK("k_tail");
  if (fuse_scalars) return BJX_OK;
  if (a.vi_type == BJX_VI_FULL_RANK) {
    if (p.fr_tc) {
      rc = launch_fullrank_dM_tc(a, gz, a.out + 1 + a.D, ws + p.o_fr_tc, t.entropy_weight, st);
      if (rc) return rc;
    } else {
      dim3 grid((unsigned)((a.D + 63) / 64), (unsigned)((a.D + 63) / 64));
      k_fullrank_dM<<<grid, 256, 0, st>>>(gz, (const float*)(ws + p.o_eps), a.raw_scale, a.S, a.D, a.scale_bijector,
                                          t.entropy_weight, a.out + 1 + a.D);
      BJX_LAUNCH_CHECK("k_fullrank_dM");
    }
  }
  if (a.vi_type == BJX_VI_LOW_RANK) {
    rc = launch_lowrank_grads(a, gz, (const float*)(ws + p.o_eps), ws + p.o_lr, t.entropy_weight, a.out + 1 + a.D, st);
    if (rc) return rc;
  }
  k_scalars<<<1, 256, 0, st>>>(t, a.K, p.Pn, a.kwargs, (const float*)(ws + p.o_theta), (const float*)(ws + p.o_qp),
                               p.n_qp_blocks, stat, a.vi_type == BJX_VI_LOW_RANK ? lowrank_logdet_ptr(a, ws + p.o_lr) : nullptr,
                               a.out);
  BJX_LAUNCH_CHECK("k_scalars");
  return BJX_OK;
}

// ---- fused optimiser step (optax.adam, eps outside the square root, bias-corrected) ---------------------------------
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                              const float* __restrict__ g, int64_t n, AdamHyper h, float c1, float c2) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  adam_one(p, m, v, g[i], i, h, c1, c2);
}

// device-resident training loop: the step number lives on the device, so that a captured CUDA graph of whole training
// steps (ELBO + gradient -> Adam -> bookkeeping) can be replayed without any per-replay update
__global__ void __launch_bounds__(256) k_adam_dev(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                  const float* __restrict__ g, int64_t n, AdamHyper h,
                                                  const int64_t* __restrict__ steps_done) {
  __shared__ float c[2];
  if (threadIdx.x == 0) adam_bias(h, *steps_done + 1, c[0], c[1]);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  adam_one(p, m, v, g[i], i, h, c[0], c[1]);
}

// few parameters (n <= 1024): Adam and the bookkeeping of k_train_bump in one block, one launch
__global__ void __launch_bounds__(1024) k_adam_bump_small(TrainTail tt, const float* __restrict__ out) {
  __shared__ float c[2];
  const int64_t t0 = *tt.steps_done;
  if (threadIdx.x == 0) adam_bias(tt.h, t0 + 1, c[0], c[1]);
  __syncthreads();
  const int64_t i = threadIdx.x;
  if (i < tt.n) adam_one(tt.params, tt.m, tt.v, out[1 + i], i, tt.h, c[0], c[1]);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (tt.losses) tt.losses[t0] = out[0];
    *tt.steps_done = t0 + 1;
  }
}

__global__ void k_train_bump(int64_t* __restrict__ steps_done, const float* __restrict__ out, float* __restrict__ losses) {
  const int64_t t = *steps_done;
  if (losses) losses[t] = out[0];
  *steps_done = t + 1;
}

}  // namespace bjx

// =========================================================================================================================
extern "C" {

int bjx_abi_version(void) { return BJX_ABI_VERSION; }

const char* bjx_last_error(void) { return bjx::g_last_error.c_str(); }

int bjx_device_sm_count(void) { return bjx::sm_count(); }

size_t bjx_elbo_workspace_bytes(const BjxElboArgs* args) {
  if (!args) return 0;
  bjx::Plan p;
  if (bjx::make_plan(*args, p) != BJX_OK) return 0;
  return p.total;
}

int bjx_elbo_kernel_choice(const BjxElboArgs* args) {
  using namespace bjx;
  BJX_REQUIRE(args != nullptr, "args is null");
  Plan p;
  int rc = make_plan(*args, p);
  if (rc) return rc;
  return p.kernel;
}

int bjx_elbo_supports_row_index(const BjxElboArgs* args) {
  using namespace bjx;
  BJX_REQUIRE(args != nullptr, "args is null");
  BjxElboArgs a = *args;
  static const int32_t dummy = 0;
  if (a.row_index == nullptr) a.row_index = &dummy;
  if (a.x_planes == nullptr) a.N_full = 0;  // asked for the raw-fp32 route
  Plan p;
  return make_plan(a, p) == BJX_OK ? 1 : 0;
}

int bjx_elbo_step(const BjxElboArgs* args, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(args != nullptr, "args is null");
  Plan p;
  int rc = make_plan(*args, p);
  if (rc) return rc;
  rc = check_ptrs(*args, p);
  if (rc) return rc;
  return run_step(*args, p, (cudaStream_t)stream);
}

int bjx_elbo_step_host(const BjxElboArgs* args, const uint32_t* key_host, const float* params_host, float* out_host,
                       void* stream) {
  using namespace bjx;
  BJX_REQUIRE(args != nullptr && key_host && params_host && out_host, "null pointer");
  Plan p;
  int rc = make_plan(*args, p);
  if (rc) return rc;
  BjxElboArgs a = *args;
  char* ws = (char*)a.workspace;
  BJX_REQUIRE(ws != nullptr, "workspace is null");
  if (a.workspace_bytes < p.total)
    return fail(BJX_ERR_WORKSPACE, "workspace too small: have %zu bytes, need %zu", a.workspace_bytes, p.total);
  cudaStream_t st = (cudaStream_t)stream;
  float* sp = (float*)(ws + p.o_stage_params);
  a.key = (const uint32_t*)(ws + p.o_stage_key);
  a.loc = sp;
  a.raw_scale = sp + a.D;
  a.kwargs = a.K ? sp + a.D + p.Pn : nullptr;
  a.out = (float*)(ws + p.o_stage_out);
  rc = check_ptrs(a, p);
  if (rc) return rc;
  const size_t n_params = (size_t)(a.D + p.Pn + a.K);
  BJX_CUDA_TRY(cudaMemcpyAsync((void*)a.key, key_host, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  BJX_CUDA_TRY(cudaMemcpyAsync(sp, params_host, n_params * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = run_step(a, p, st);
  if (rc) return rc;
  BJX_CUDA_TRY(cudaMemcpyAsync(out_host, a.out, (1 + n_params) * sizeof(float), cudaMemcpyDeviceToHost, st));
  BJX_CUDA_TRY(cudaStreamSynchronize(st));
  return BJX_OK;
}

int bjx_posterior_sample(const BjxElboArgs* args, float* theta_out, float* z_out, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(args != nullptr && theta_out != nullptr, "null pointer");
  BjxElboArgs a = *args;
  // sampling needs no data: plan as a data-free coin problem so that no likelihood scratch is required
  a.likelihood = BJX_LIK_BERNOULLI_PROBS;
  a.w_offset = 0;
  a.N = 0;
  Plan p;
  int rc = make_plan(a, p);
  if (rc) return rc;
  BJX_REQUIRE(a.key && a.loc && a.raw_scale && a.coords && a.workspace, "null key/loc/raw_scale/coords/workspace pointer");
  if (a.workspace_bytes < p.total)
    return fail(BJX_ERR_WORKSPACE, "workspace too small: have %zu bytes, need %zu", a.workspace_bytes, p.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)a.workspace;
  a.eps_out = nullptr;
  rc = run_prologue(a, p, ws, st);
  if (rc) return rc;
  const size_t SD = (size_t)a.S * a.D * sizeof(float);
  BJX_CUDA_TRY(cudaMemcpyAsync(theta_out, ws + p.o_theta, SD, cudaMemcpyDeviceToDevice, st));
  if (z_out) {
    if (a.vi_type != BJX_VI_MEAN_FIELD) {
      BJX_CUDA_TRY(cudaMemcpyAsync(z_out, ws + p.o_z, SD, cudaMemcpyDeviceToDevice, st));
    } else {
      return fail(BJX_ERR_UNSUPPORTED, "z_out is only materialised for the full-rank family; invert the bijector instead");
    }
  }
  return BJX_OK;
}

struct ShardExchange {  // row-sharded training: the packed step output is summed over the ranks before Adam
  void* const* bufs_host;
  int32_t rank, world;
  uint64_t epoch_base;
};

static int train_steps_impl(const BjxElboArgs* args, const BjxMinibatch* mb, int64_t n_steps, float* params_flat, float* adam_m,
                            float* adam_v, double lr, double b1, double b2, double eps, int64_t* steps_done, float* losses,
                            void* stream, const ShardExchange* xch = nullptr) {
  using namespace bjx;
  BJX_REQUIRE(args != nullptr && params_flat && adam_m && adam_v && steps_done, "null pointer");
  BJX_REQUIRE(n_steps >= 0, "n_steps must be >= 0");
  Plan p;
  int rc = make_plan(*args, p);
  if (rc) return rc;
  BjxElboArgs a = *args;
  a.loc = params_flat;
  a.raw_scale = params_flat + a.D;
  a.kwargs = a.K ? params_flat + a.D + p.Pn : nullptr;
  a.key_index = steps_done;
  if (mb) {  // the step's seed is the half of split(key[*steps_done]) that the gather kernel leaves in mb->elbo_key
    BJX_REQUIRE(mb->elbo_key != nullptr, "minibatch training needs mb->elbo_key");
    BJX_REQUIRE(mb->B == a.N, "minibatch: args->N (%lld) must equal the batch size (%lld)", (long long)a.N, (long long)mb->B);
    a.key = mb->elbo_key;
    a.key_index = nullptr;
  }
  rc = check_ptrs(a, p);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = a.D + p.Pn + a.K;
  const TrainTail tt{params_flat, adam_m, adam_v, adam_hyper(lr, b1, b2, eps), steps_done, losses, n};
  for (int64_t i = 0; i < n_steps; ++i) {
    if (mb) {
      rc = launch_minibatch_gather(*mb, args->key, steps_done, a.threefry_layout, st);
      if (rc) return rc;
    }
    bool finished = false;
    rc = run_step(a, p, st, xch ? nullptr : &tt, &finished);
    if (rc) return rc;
    if (finished) continue;  // the step kernel did Adam and the bookkeeping itself
    if (xch) {  // every rank ends up with the same bits in a.out, so the replicated Adam states stay identical
      rc = bjx_allreduce_packed_p2p_dev(a.out, 1 + n, xch->bufs_host, xch->rank, xch->world, xch->epoch_base, steps_done, stream);
      if (rc) return rc;
    }
    if (n <= 1024) {
      k_adam_bump_small<<<1, 1024, 0, st>>>(tt, a.out);
      BJX_LAUNCH_CHECK("k_adam_bump_small");
      continue;
    }
    k_adam_dev<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params_flat, adam_m, adam_v, a.out + 1, n, tt.h, steps_done);
    BJX_LAUNCH_CHECK("k_adam_dev");
    k_train_bump<<<1, 1, 0, st>>>(steps_done, a.out, losses);
    BJX_LAUNCH_CHECK("k_train_bump");
  }
  return BJX_OK;
}

int bjx_train_steps(const BjxElboArgs* args, int64_t n_steps, float* params_flat, float* adam_m, float* adam_v,
                    double lr, double b1, double b2, double eps, int64_t* steps_done, float* losses, void* stream) {
  return train_steps_impl(args, nullptr, n_steps, params_flat, adam_m, adam_v, lr, b1, b2, eps, steps_done, losses, stream);
}

int bjx_train_steps_sharded(const BjxElboArgs* args, int64_t n_steps, float* params_flat, float* adam_m, float* adam_v, double lr,
                            double b1, double b2, double eps, int64_t* steps_done, float* losses, void* const* p2p_bufs_host,
                            int32_t rank, int32_t world, uint64_t epoch_base, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(p2p_bufs_host != nullptr && world >= 1 && world <= 8 && rank >= 0 && rank < world, "bad exchange arguments");
  const ShardExchange xch{p2p_bufs_host, rank, world, epoch_base};
  return train_steps_impl(args, nullptr, n_steps, params_flat, adam_m, adam_v, lr, b1, b2, eps, steps_done, losses, stream, &xch);
}

int bjx_train_steps_minibatch(const BjxElboArgs* args, const BjxMinibatch* mb, int64_t n_steps, float* params_flat, float* adam_m,
                              float* adam_v, double lr, double b1, double b2, double eps, int64_t* steps_done, float* losses,
                              void* stream) {
  using namespace bjx;
  BJX_REQUIRE(mb != nullptr, "mb is null");
  return train_steps_impl(args, mb, n_steps, params_flat, adam_m, adam_v, lr, b1, b2, eps, steps_done, losses, stream);
}

size_t bjx_x_planes_bytes(int64_t N, int64_t Dw) {
  if (N < 0 || Dw < 1) return 0;
  return bjx::x_planes_bytes(N, Dw);
}

int bjx_prepare_x_planes_rows(const float* X_rows, int64_t rows, int64_t Dw, int64_t ldx, void* planes, int64_t N_total, int64_t row0,
                              void* stream) {
  using namespace bjx;
  BJX_REQUIRE(X_rows != nullptr && planes != nullptr, "null pointer");
  BJX_REQUIRE(rows >= 0 && row0 >= 0 && row0 + rows <= N_total && Dw >= 1 && ldx >= Dw, "need 0 <= row0, row0 + rows <= N_total, Dw >= 1, ldx >= Dw");
  BJX_REQUIRE(((uintptr_t)planes & 15) == 0, "planes must be 16-byte aligned");
  return split_x_planes_rows(X_rows, rows, Dw, ldx, planes, N_total, row0, (cudaStream_t)stream);
}

int bjx_prepare_x_planes(const float* X, int64_t N, int64_t Dw, int64_t ldx, void* planes, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(X != nullptr && planes != nullptr, "null pointer");
  BJX_REQUIRE(N >= 0 && Dw >= 1 && ldx >= Dw, "need N >= 0, Dw >= 1, ldx >= Dw");
  BJX_REQUIRE(((uintptr_t)planes & 15) == 0, "planes must be 16-byte aligned");
  return split_x_planes(X, N, Dw, ldx, planes, (cudaStream_t)stream);
}

int bjx_profile_begin(int32_t capacity) {
  using namespace bjx;
  BJX_REQUIRE(capacity > 0 && capacity <= (1 << 16), "capacity must be in [1, 65536]");
  bjx_profile_end();
  g_prof.ev.resize(2 * (size_t)capacity);
  for (auto& e : g_prof.ev) BJX_CUDA_TRY(cudaEventCreate(&e));
  g_prof.capacity = capacity;
  g_prof.used = 0;
  return BJX_OK;
}

int bjx_profile_collect(float* ms_out_host, int32_t max_out, int32_t* n_out_host) {
  using namespace bjx;
  BJX_REQUIRE(ms_out_host && n_out_host && max_out >= 0, "null pointer");
  int n = g_prof.used < max_out ? g_prof.used : max_out;
  for (int i = 0; i < n; ++i) {
    BJX_CUDA_TRY(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
    BJX_CUDA_TRY(cudaEventElapsedTime(&ms_out_host[i], g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
  }
  *n_out_host = n;
  g_prof.used = 0;
  return BJX_OK;
}

int bjx_profile_end(void) {
  using namespace bjx;
  for (auto& e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  g_prof.capacity = 0;
  g_prof.used = 0;
  return BJX_OK;
}

int bjx_adam_update(float* params, float* m, float* v, const float* grads, int64_t n, int64_t step, double lr, double b1,
                    double b2, double eps, void* stream) {
  using namespace bjx;
  BJX_REQUIRE(params && m && v && grads, "null pointer");
  BJX_REQUIRE(n >= 0 && step >= 1, "need n >= 0 and step >= 1");
  BJX_REQUIRE(b1 > 0.0 && b1 < 1.0 && b2 > 0.0 && b2 < 1.0, "need 0 < b1, b2 < 1");
  if (n == 0) return BJX_OK;
  const AdamHyper h = adam_hyper(lr, b1, b2, eps);
  const float c1 = (float)(-expm1((double)step * h.lb1)), c2 = (float)(-expm1((double)step * h.lb2));
  k_adam<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, m, v, grads, n, h, c1, c2);
  BJX_LAUNCH_CHECK("k_adam");
  return BJX_OK;
}

}  // extern "C"

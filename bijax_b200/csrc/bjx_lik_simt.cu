// Kernel 5+6, CUDA-core (fp32 FFMA) variants of the streaming likelihood pass:
//   eta[n, s] = X[n, :] . theta[s, w]      (user likelihood_fn, advi.py:90; exemplars README.md:99-106)
//   stat[s]   = sum_n f(y_n, eta[n, s])    (likelihood.log_prob(outputs).sum(), advi.py:91)
//   G[s, d]   = sum_n g(y_n, eta[n, s]) X[n, d]   (the transposed dot_general of the VJP, utils.py:7)
// Two kernels: a single-pass register-resident one for a handful of columns (config c2: D = 5), and a generic
// two-pass tiled fallback for any shape the tensor-core kernel (bjx_lik_tc.cu) does not take.
#include <cstdlib>
#include <type_traits>

#include "bjx_elbo.cuh"
#include "bjx_sgemm.cuh"

namespace bjx {

// =====================================================================================================================
// small-D: each thread owns one Monte-Carlo sample s (theta[s, :] and G[s, :] live in registers) and a lane of rows;
// rows are staged through shared memory as padded records [x_0 .. x_{Dw-1}, y, 0...] so that a warp reads one record
// with REC4 broadcast LDS.128.  X is read from HBM exactly once, coalesced.
// =====================================================================================================================
constexpr int SD_CHUNK = SD_CHUNK_ROWS;

// SPT = Monte-Carlo samples per thread.  With two, a record read from shared memory (LDS.128 broadcast) and its address
// arithmetic serve two samples: 21 -> ~15 instructions per (row, sample) pair of which 12 are the arithmetic itself (config c2 is
// bound by instruction issue, profiles/r01_k_lik_smalld_c2_ncu_full.txt).  Rows per inner iteration shrink to two, so the
// register count -- and with it the number of resident blocks -- stays where it was.
template <int REC4, int DWT, int SPT>  // DWT = the column count as a compile-time constant: the padding slots cost no FMAs
__global__ void __launch_bounds__(256, REC4 <= 2 ? 4 : 1) k_lik_smalld(const float* __restrict__ X, int64_t ldx, const float* __restrict__ y,
                                                    int64_t N, int Dw, const float* __restrict__ theta, int64_t D,
                                                    int64_t w_off, int S, int family, int SB, int64_t rows_per_cta,
                                                    float* __restrict__ Gpart, float* __restrict__ statpart) {
  constexpr int REC = REC4 * 4;    // record = [y, x_0 .. x_{Dw-1}, 0 ...]: y sits at a fixed slot
  constexpr int ROWS = 4 / SPT;    // rows per inner iteration: ROWS * SPT independent eta chains give the FMA pipe its ILP
  __shared__ __align__(16) float sm[256 * (REC + 1)];
  const int tid = threadIdx.x;
  const int SBT = SB / SPT;        // threads along the samples of this block (SB samples)
  const int RL = 256 / SBT;
  const int sl = tid % SBT, rl = tid / SBT;
  int sidx[SPT];
  bool s_ok[SPT];
#pragma unroll
  for (int k = 0; k < SPT; ++k) {
    sidx[k] = blockIdx.y * SB + sl + k * SBT;
    s_ok[k] = sidx[k] < S;
  }

  float th[SPT][REC], G[SPT][REC];  // th[.][0] = 0 (the y slot), th[.][1 + d] = theta[s, d]
  float stat[SPT];
#pragma unroll
  for (int k = 0; k < SPT; ++k) {
#pragma unroll
    for (int d = 0; d < REC; ++d) {
      th[k][d] = (s_ok[k] && d >= 1 && d <= Dw) ? theta[(int64_t)sidx[k] * D + w_off + d - 1] : 0.0f;
      G[k][d] = 0.0f;
    }
    stat[k] = 0.0f;
  }

  // ROWS rows x SPT samples; MASKED: rows past nvalid contribute nothing (only the last group of a chunk can hold such rows)
  auto group = [&](int r0, int nvalid, auto masked_tag) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    float rec[ROWS][REC];
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const int r = MASKED ? min(r0 + u * RL, SD_CHUNK - 1) : r0 + u * RL;
      const float4* rp = reinterpret_cast<const float4*>(sm + r * REC);
#pragma unroll
      for (int q = 0; q < REC4; ++q) {
        const float4 v = rp[q];
        rec[u][4 * q + 0] = v.x;
        rec[u][4 * q + 1] = v.y;
        rec[u][4 * q + 2] = v.z;
        rec[u][4 * q + 3] = v.w;
      }
    }
    float eta[ROWS][SPT];
#pragma unroll
    for (int u = 0; u < ROWS; ++u)
#pragma unroll
      for (int k = 0; k < SPT; ++k) eta[u][k] = 0.0f;
#pragma unroll
    for (int d = 1; d <= DWT; ++d)
#pragma unroll
      for (int u = 0; u < ROWS; ++u)
#pragma unroll
        for (int k = 0; k < SPT; ++k) eta[u][k] = fmaf(rec[u][d], th[k][d], eta[u][k]);
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const bool ok = !MASKED || (r0 + u * RL < nvalid);
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        float f, g;
        link_eval(family, rec[u][0], eta[u][k], f, g);
        if (MASKED) {
          f = ok ? f : 0.0f;
          g = ok ? g : 0.0f;
        }
        stat[k] += f;
#pragma unroll
        for (int d = 1; d <= DWT; ++d) G[k][d] = fmaf(g, rec[u][d], G[k][d]);
      }
    }
  };

  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_end = min(N, row_begin + rows_per_cta);
  // this thread's row of a chunk, fetched into registers one chunk ahead: the global-load latency of chunk c + 1 runs under
  // the arithmetic of chunk c instead of in front of a barrier
  float nx[REC];
  auto fetch = [&](int64_t base) {
    const int64_t row = base + tid;
    if (row < row_end) {
      const float* xr = X + row * ldx;
      nx[0] = y[row];
#pragma unroll
      for (int d = 1; d < REC; ++d) nx[d] = d <= Dw ? xr[d - 1] : 0.0f;
    } else {  // rows past the end: eta = 0 and y = 0 would still contribute f(0, 0); they are masked below
#pragma unroll
      for (int d = 0; d < REC; ++d) nx[d] = 0.0f;
    }
  };
  if (row_begin < row_end) fetch(row_begin);
  for (int64_t base = row_begin; base < row_end; base += SD_CHUNK) {
    const int nvalid = (int)min((int64_t)SD_CHUNK, row_end - base);
    __syncthreads();  // previous chunk fully consumed
    {
      float4* rec = reinterpret_cast<float4*>(sm + tid * REC);
#pragma unroll
      for (int q = 0; q < REC4; ++q) rec[q] = make_float4(nx[4 * q], nx[4 * q + 1], nx[4 * q + 2], nx[4 * q + 3]);
    }
    __syncthreads();
    if (base + SD_CHUNK < row_end) fetch(base + SD_CHUNK);
    int r0 = rl;
    for (; r0 + (ROWS - 1) * RL < nvalid; r0 += ROWS * RL) group(r0, nvalid, std::false_type{});  // every row of the group is valid
    if (r0 < nvalid) group(r0, nvalid, std::true_type{});
  }

  // deterministic reduction over the RL row lanes of this CTA, one sample slot of the threads at a time (RL * SBT = 256 entries)
#pragma unroll
  for (int k = 0; k < SPT; ++k) {
    __syncthreads();
    float* red = sm + (size_t)(rl * SBT + sl) * (REC + 1);
#pragma unroll
    for (int d = 0; d < REC; ++d) red[d] = G[k][d];
    red[REC] = stat[k];
    __syncthreads();
    if (rl == 0 && s_ok[k]) {
      float acc[REC + 1];
#pragma unroll
      for (int d = 0; d <= REC; ++d) acc[d] = 0.0f;
      for (int l = 0; l < RL; ++l) {
        const float* rr = sm + (size_t)(l * SBT + sl) * (REC + 1);
#pragma unroll
        for (int d = 0; d <= REC; ++d) acc[d] += rr[d];
      }
      float* gp = Gpart + ((size_t)blockIdx.x * S + sidx[k]) * Dw;
#pragma unroll
      for (int d = 1; d < REC; ++d)
        if (d <= Dw) gp[d - 1] = acc[d];
      statpart[(size_t)blockIdx.x * S + sidx[k]] = acc[REC];
    }
  }
}

int smalld_occupancy(int rec4) {
  int occ = 0;
  cudaError_t e = cudaErrorInvalidValue;
  switch (rec4) {
    // (the widest column count of each record width: the most registers, a safe lower bound for the others)
    case 1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lik_smalld<1, 3, 2>, 256, 0); break;
    case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lik_smalld<2, 7, 2>, 256, 0); break;
    case 3: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lik_smalld<3, 11, 2>, 256, 0); break;
    case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lik_smalld<4, 15, 2>, 256, 0); break;
    default: break;
  }
  return (e == cudaSuccess && occ > 0) ? occ : 1;
}

int launch_lik_smalld(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st) {
  const float* theta = (const float*)(ws + p.o_theta);
  float* Gpart = (float*)(ws + p.o_Gpart);
  float* statpart = (float*)(ws + p.o_statpart);
  const int64_t rows_per_cta = (a.N + p.sd_grid_x - 1) / p.sd_grid_x;
  dim3 grid(p.sd_grid_x, p.sd_grid_y);
  // two samples per thread whenever the block has at least two (BJX_SD_SPT=1 keeps one: development switch)
  static int spt_env = -1;
  if (spt_env < 0) {
    const char* e = getenv("BJX_SD_SPT");
    spt_env = e ? atoi(e) : 2;
  }
  const bool two = p.sd_SB >= 2 && spt_env == 2;
#define BJX_SD_LAUNCH(R4, DW)                                                                                            \
  do {                                                                                                                   \
    if (two)                                                                                                             \
      k_lik_smalld<R4, DW, 2><<<grid, 256, 0, st>>>(a.X, a.ldx, a.y, a.N, (int)a.Dw, theta, a.D, a.w_offset, (int)a.S,   \
                                                    a.likelihood, p.sd_SB, rows_per_cta, Gpart, statpart);               \
    else                                                                                                                 \
      k_lik_smalld<R4, DW, 1><<<grid, 256, 0, st>>>(a.X, a.ldx, a.y, a.N, (int)a.Dw, theta, a.D, a.w_offset, (int)a.S,   \
                                                    a.likelihood, p.sd_SB, rows_per_cta, Gpart, statpart);               \
  } while (0)
  switch ((int)a.Dw) {
    case 1: BJX_SD_LAUNCH(1, 1); break;
    case 2: BJX_SD_LAUNCH(1, 2); break;
    case 3: BJX_SD_LAUNCH(1, 3); break;
    case 4: BJX_SD_LAUNCH(2, 4); break;
    case 5: BJX_SD_LAUNCH(2, 5); break;
    case 6: BJX_SD_LAUNCH(2, 6); break;
    case 7: BJX_SD_LAUNCH(2, 7); break;
    case 8: BJX_SD_LAUNCH(3, 8); break;
    case 9: BJX_SD_LAUNCH(3, 9); break;
    case 10: BJX_SD_LAUNCH(3, 10); break;
    case 11: BJX_SD_LAUNCH(3, 11); break;
    case 12: BJX_SD_LAUNCH(4, 12); break;
    case 13: BJX_SD_LAUNCH(4, 13); break;
    case 14: BJX_SD_LAUNCH(4, 14); break;
    case 15: BJX_SD_LAUNCH(4, 15); break;
    default: return fail(BJX_ERR_UNSUPPORTED, "small-D kernel supports Dw <= 15 (got %lld)", (long long)a.Dw);
  }
#undef BJX_SD_LAUNCH
  BJX_LAUNCH_CHECK("k_lik_smalld");
  return BJX_OK;
}

// =====================================================================================================================
// generic fp32 fallback, two passes:
//   A: eta tile = X @ theta^T (BM=128 rows x BN=32 samples), link function in the epilogue, g[n, s] -> scratch
//   B: G = g^T @ X, split over row ranges into partial slots (reduced deterministically by the tail kernel)
// =====================================================================================================================
struct LoadRowMajorK {  // M(row, k) = base[row * ld + k]
  static constexpr bool k_fast = true;
  const float* base;
  int64_t ld, rows, cols;
  __device__ __forceinline__ float operator()(int64_t r, int64_t k) const {
    return (r < rows && k < cols) ? base[r * ld + k] : 0.0f;
  }
};
struct LoadTransposed {  // M(row, k) = base[k * ld + row]
  static constexpr bool k_fast = false;
  const float* base;
  int64_t ld, rows, kmax;
  __device__ __forceinline__ float operator()(int64_t r, int64_t k) const {
    return (r < rows && k < kmax) ? base[k * ld + r] : 0.0f;
  }
};

__global__ void __launch_bounds__(256) k_lik_tiled_A(const float* __restrict__ X, int64_t ldx, const float* __restrict__ y,
                                                     int64_t N, int64_t Dw, const float* __restrict__ theta, int64_t D,
                                                     int64_t w_off, int S, int family, float* __restrict__ gbuf,
                                                     float* __restrict__ statpart) {
  constexpr int BM = 128, BN = 32;
  __shared__ SgemmSmem<BM, BN> sm;
  __shared__ float red[16][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  LoadRowMajorK la{X, ldx, N, Dw};
  LoadRowMajorK lb{theta + w_off, D, S, Dw};
  float stat_acc[BN / 16] = {0.0f, 0.0f};
  const int64_t m_tiles = (N + BM - 1) / BM;
  for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
    const int64_t m0 = mt * BM;
    float acc[BM / 16][BN / 16];
    sgemm_tile<BM, BN>(acc, la, lb, m0, n0, 0, Dw, sm);
#pragma unroll
    for (int i = 0; i < BM / 16; ++i) {
      const int64_t m = m0 + ty + 16 * i;
      if (m >= N) continue;
      const float yv = y[m];
#pragma unroll
      for (int j = 0; j < BN / 16; ++j) {
        const int64_t s = n0 + tx + 16 * j;
        if (s >= S) continue;
        float f, g;
        link_eval(family, yv, acc[i][j], f, g);
        gbuf[m * S + s] = g;
        stat_acc[j] += f;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < BN / 16; ++j) red[ty][tx + 16 * j] = stat_acc[j];
  __syncthreads();
  if (tid < BN && n0 + tid < S) {
    float t = 0.0f;
    for (int r = 0; r < 16; ++r) t += red[r][tid];
    statpart[(size_t)blockIdx.x * S + n0 + tid] = t;
  }
}

__global__ void __launch_bounds__(256) k_lik_tiled_B(const float* __restrict__ X, int64_t ldx, int64_t N, int64_t Dw,
                                                     const float* __restrict__ gbuf, int S, int64_t rows_per_slot,
                                                     int n_tiles, float* __restrict__ Gpart) {
  constexpr int BM = 32, BN = 128;
  __shared__ SgemmSmem<BM, BN> sm;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ks = blockIdx.x;
  const int mt = blockIdx.y / n_tiles, nt = blockIdx.y % n_tiles;
  const int64_t m0 = (int64_t)mt * BM, n0 = (int64_t)nt * BN;
  const int64_t k_begin = (int64_t)ks * rows_per_slot;
  const int64_t k_end = min(N, k_begin + rows_per_slot);
  LoadTransposed la{gbuf, S, S, N};    // A(m = s, k = row) = g[row, s]
  LoadTransposed lb{X, ldx, Dw, N};    // B(n = d, k = row) = X[row, d]
  float acc[BM / 16][BN / 16];
  sgemm_tile<BM, BN>(acc, la, lb, m0, n0, k_begin, k_end > k_begin ? k_end : k_begin, sm);
#pragma unroll
  for (int i = 0; i < BM / 16; ++i) {
    const int64_t s = m0 + ty + 16 * i;
    if (s >= S) continue;
#pragma unroll
    for (int j = 0; j < BN / 16; ++j) {
      const int64_t d = n0 + tx + 16 * j;
      if (d < Dw) Gpart[((size_t)ks * S + s) * Dw + d] = acc[i][j];
    }
  }
}

int launch_lik_tiled(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st) {
  const float* theta = (const float*)(ws + p.o_theta);
  float* gbuf = (float*)(ws + p.o_gbuf);
  float* Gpart = (float*)(ws + p.o_Gpart);
  float* statpart = (float*)(ws + p.o_statpart);
  k_lik_tiled_A<<<dim3(p.tA_grid_x, p.tA_grid_y), 256, 0, st>>>(a.X, a.ldx, a.y, a.N, a.Dw, theta, a.D, a.w_offset,
                                                                (int)a.S, a.likelihood, gbuf, statpart);
  BJX_LAUNCH_CHECK("k_lik_tiled_A");
  const int n_tiles = (int)((a.Dw + 127) / 128);
  k_lik_tiled_B<<<dim3(p.tB_kslots, p.tB_tiles), 256, 0, st>>>(a.X, a.ldx, a.N, a.Dw, gbuf, (int)a.S, p.tB_rows_per_slot,
                                                               n_tiles, Gpart);
  BJX_LAUNCH_CHECK("k_lik_tiled_B");
  return BJX_OK;
}

}  // namespace bjx

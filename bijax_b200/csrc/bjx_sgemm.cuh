// A small fp32 (FFMA) tiled contraction used by the CUDA-core paths: the generic likelihood fallback and the
// full-rank S x D x D contractions.  C[m, n] = sum_k A(m, k) * B(n, k), operands read through functors so the same
// core serves row-major, transposed and masked (lower-triangular) inputs.  256 threads, (BM/16) x (BN/16) per thread.
#pragma once
#include "bjx_common.cuh"

namespace bjx {

constexpr int SG_BK = 16;
constexpr int SG_PAD = 2;  // row stride == 2 (mod 32): conflict-free for both the k-fast and the m-fast store patterns

template <int BM, int BN>
struct SgemmSmem {
  float As[SG_BK][BM + SG_PAD];
  float Bs[SG_BK][BN + SG_PAD];
};

// LA / LB expose:  static constexpr bool k_fast;  __device__ float operator()(int64_t row, int64_t k) const
// (returning 0 outside the matrix).
template <int BM, int BN, class LA, class LB>
__device__ __forceinline__ void sgemm_tile(float (&acc)[BM / 16][BN / 16], const LA& la, const LB& lb, int64_t m0,
                                           int64_t n0, int64_t k_begin, int64_t k_end, SgemmSmem<BM, BN>& sm) {
  constexpr int TM = BM / 16, TN = BN / 16;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  for (int64_t kk = k_begin; kk < k_end; kk += SG_BK) {
#pragma unroll
    for (int e = tid; e < BM * SG_BK; e += 256) {
      const int k = LA::k_fast ? (e % SG_BK) : (e / BM);
      const int m = LA::k_fast ? (e / SG_BK) : (e % BM);
      sm.As[k][m] = (kk + k < k_end) ? la(m0 + m, kk + k) : 0.0f;
    }
#pragma unroll
    for (int e = tid; e < BN * SG_BK; e += 256) {
      const int k = LB::k_fast ? (e % SG_BK) : (e / BN);
      const int n = LB::k_fast ? (e / SG_BK) : (e % BN);
      sm.Bs[k][n] = (kk + k < k_end) ? lb(n0 + n, kk + k) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = sm.As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = sm.Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

}  // namespace bjx

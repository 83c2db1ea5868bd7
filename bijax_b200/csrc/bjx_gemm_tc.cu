// TMA-fed tcgen05 GEMM with split-bf16 operands (fp32 in, fp32-accurate out) and fused epilogues.
//
// One warp-specialised persistent kernel serves the four dense contractions of the path that do not fit the fused
// single-pass kernel of bjx_lik_tc.cu:
//
//   EPI_LIK1  eta[n, s] = sum_d X[n, d] theta[s, d]   -> link function in the epilogue: g = dll/deta (written as bf16 split
//                                                        planes for the next pass), stat[s] += f           (advi.py:90-91)
//   EPI_LIK2  G^T[d, s] = sum_n X[n, d] g[n, s]       -> the transposed contraction of the VJP             (utils.py:7)
//   EPI_Z     z[s, i]   = loc[i] + sum_{j<=i} L[i, j] eps[s, j]   full-rank sample, tfd.MultivariateNormalTriL (advi.py:83)
//   EPI_DM    dM[i, j]  = [j<=i] scale'(M_ij) sum_s gz[s, i] eps[s, j] - [i==j] prior_weight dlog|L_ii|     its VJP
//             (computed as the transposed tile sum_s eps[s, j] gz[s, i] so that the stores are coalesced)
//
// fp32 operands are pre-split into NT bf16 planes (x = x1 + x2 (+ x3), bjx_split_planes) so that every operand tile is a
// plain TMA box; the products a_i * b_j with i + j <= NT + 1 accumulate in fp32 in TMEM (NT = 2: 3 products, 2^-17 relative;
// NT = 3: 6 products, fp32-accurate).  Warp roles: warp 0 TMA producer, warp 1 MMA issuer (owns TMEM), warps 2-5 epilogue
// (one TMEM sub-partition each).  Two accumulator stages in TMEM let the epilogue of one virtual tile overlap the MMAs of
// the next.  Long accumulations over data rows (EPI_LIK2) are cut into virtual tiles of `kb_per_vtile` k-blocks whose
// accumulators are added into an fp32 partial by the epilogue, because the tensor core's fp32 accumulate truncates
// (DESIGN.md 3.1).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "bjx_elbo.cuh"
#include "bjx_tc_ptx.cuh"
#include "bjx_variational.cuh"

namespace bjx {
namespace gemm {

constexpr int BM = 128;            // UMMA M
constexpr int BK = 64;             // bf16 elements per k-block = one 128-byte swizzle row
constexpr int A_PLANE = BM * 128;  // bytes of one A plane tile
constexpr int NACC = 2;            // accumulator stages in TMEM
// Epilogue warps: four (one per TMEM sub-partition), except the transposed contraction on 256-wide tiles, which keeps its
// 128 x 256 fp32 partial of a whole K split in REGISTERS -- 128 values per thread need two warps per sub-partition.
__host__ __device__ constexpr int epi_warps(int epi, int bn) { return (epi == 1 /* EPI_LIK2 */ && bn == 256) ? 8 : 4; }
__host__ __device__ constexpr int threads_of(int epi, int bn) { return (2 + epi_warps(epi, bn)) * 32; }

enum { EPI_LIK1 = 0, EPI_LIK2 = 1, EPI_Z = 2, EPI_DM = 3 };

struct Sched {
  int tiles_m, tiles_n;  // output tiles (BM x BN)
  int splits;            // K splits; unit u = tile + n_tiles * split
  int kb_total;          // k-blocks in the whole K extent
  int kb_per_split;
  int kb_per_vtile;      // k-blocks accumulated in TMEM before the epilogue takes the accumulator
  int tril;              // 0 none; 1 EPI_Z: k <= last column index of the tile; 2 EPI_DM (transposed tiles: rows j, columns i): tiles with all j > i are empty
};

struct EpiArgs {
  // EPI_LIK1
  const float* y;
  int64_t n_rows;
  int S;
  int family;
  __nv_bfloat16* g1;
  __nv_bfloat16* g2;
  __nv_bfloat16* g3;  // third split term (NT = 3, BJX_PREC_BF16X3) or null
  int64_t g_pitch;    // elements
  float* statpart;  // [4 * grid][S]
  // EPI_LIK2
  float* Gpart;  // [splits][S][Dw]
  int64_t Dw;
  // EPI_Z / EPI_DM
  const float* loc;
  const float* M;
  float* out;  // z [S, D] or dM [D, D]
  int64_t D;
  int scale_bij;
  float prior_weight;
};

struct Params {
  CUtensorMap mapA[3];
  CUtensorMap mapB[3];
  Sched sched;
  EpiArgs epi;
};

struct Bars {
  unsigned long long full[4], empty[4], tmem_full[NACC], tmem_empty[NACC];
  uint32_t tmem_base;
};

template <int BN, int NT, int CG = 1>
struct Cfg {
  static constexpr int B_PLANE = (BN / CG) * 128;  // per CTA: a pair splits the B tile
  static constexpr int STAGE_BYTES = NT * (A_PLANE + B_PLANE);
  static constexpr int NSTAGE = (200 * 1024 / STAGE_BYTES) < 4 ? (200 * 1024 / STAGE_BYTES) : 4;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
  static_assert(NSTAGE >= 2, "need at least two pipeline stages");
  static_assert(TMEM_COLS == 32 || TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "power of two");
};

// Walks the virtual tiles of this CTA in the one order every warp role agrees on.
template <int BN, int CG, class F>
__device__ __forceinline__ void for_each_vtile(const Sched& sc, F&& f) {
  constexpr int BMC = BM * CG;  // rows of one output tile (a CTA pair works on 256)
  const int n_tiles = sc.tiles_m * sc.tiles_n;
  const int n_units = n_tiles * sc.splits;
  for (int u = blockIdx.x / CG; u < n_units; u += gridDim.x / CG) {
    const int tile = u % n_tiles, split = u / n_tiles;
    const int m_blk = tile / sc.tiles_n, n_blk = tile % sc.tiles_n;
    int kb_lo = split * sc.kb_per_split;
    int kb_hi = kb_lo + sc.kb_per_split;
    if (kb_hi > sc.kb_total) kb_hi = sc.kb_total;
    if (sc.tril == 1) {
      const int lim = ((n_blk + 1) * BN + BK - 1) / BK;
      if (kb_hi > lim) kb_hi = lim;
    } else if (sc.tril == 2) {
      if (m_blk * BMC > n_blk * BN + BN - 1) kb_hi = kb_lo;  // every j of the tile is above every i: strictly upper triangle
    }
    if (kb_hi < kb_lo) kb_hi = kb_lo;
    int kb0 = kb_lo;
    do {
      int kb1 = kb0 + sc.kb_per_vtile;
      if (kb1 > kb_hi) kb1 = kb_hi;
      f(m_blk, n_blk, split, kb0, kb1, kb0 == kb_lo, kb1 >= kb_hi);  // + first / last virtual tile of this output unit
      kb0 = kb1;
    } while (kb0 < kb_hi);
  }
}

// Sum of v[j] over the 32 lanes for every j: afterwards lane l holds the total of column l (31 shuffles).
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int BN, int NT, bool A_MN, bool B_MN, int EPI, int CG>
__global__ void __launch_bounds__(threads_of(EPI, BN), 1) k_gemm_tc(const __grid_constant__ Params P) {
  using namespace tc;
  using C = Cfg<BN, NT, CG>;
  constexpr int BMC = BM * CG;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;  // 0 = leader: issues the pair's MMAs, owns full / tmem_empty barriers
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  Bars* bars = (Bars*)(smem + C::NSTAGE * C::STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Sched& sc = P.sched;

  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], epi_warps(EPI, BN) * CG);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      tma_prefetch_desc(&P.mapA[t]);
      tma_prefetch_desc(&P.mapB[t]);
    }
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_cg2(&bars->tmem_base, C::TMEM_COLS); else tmem_alloc(&bars->tmem_base, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before multicast commits / remote arrivals reach them
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer ================================================
    // (pair: both CTAs load their own 128 rows of A and their half of B; every byte is counted on the leader's barrier)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](void* dst, const CUtensorMap* map, int32_t c0, int32_t c1, void* bar) {
        if (CG == 2) tma_load_2d_cg2(dst, map, c0, c1, bar); else tma_load_2d(dst, map, c0, c1, bar);
      };
      for_each_vtile<BN, CG>(sc, [&](int m_blk, int n_blk, int, int kb0, int kb1, bool, bool) {
        const int m0 = m_blk * BMC + rank * BM, n0 = n_blk * BN + rank * (BN / CG);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          if (rank == 0) mbar_expect_tx(&bars->full[stage], (uint32_t)(CG * C::STAGE_BYTES));
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + NT * A_PLANE;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (!A_MN) {
              load(sa + t * A_PLANE, &P.mapA[t], kb * BK, m0, &bars->full[stage]);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) load(sa + t * A_PLANE + i * 8192, &P.mapA[t], m0 + 64 * i, kb * BK, &bars->full[stage]);
            }
            if (!B_MN) {
              load(sb + t * C::B_PLANE, &P.mapB[t], kb * BK, n0, &bars->full[stage]);
            } else {
#pragma unroll
              for (int i = 0; i < BN / CG / 64; ++i) load(sb + t * C::B_PLANE + i * 8192, &P.mapB[t], n0 + 64 * i, kb * BK, &bars->full[stage]);
            }
          }
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1u;
          }
        }
      });
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (the leader CTA of a pair) ========================
    if (rank == 0) {
    constexpr uint32_t IDESC = make_idesc(BMC, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    // K-major tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart, K advances 32 bytes per UMMA (K = 16)
    // MN-major tile: blocks of [64 k-rows][64 mn] = 8 KB, 8-row atoms 1024 bytes apart, K advances 16 rows = 2048 bytes
    constexpr uint32_t A_LBO = A_MN ? 8192 : 16, B_LBO = B_MN ? 8192 : 16;
    constexpr uint64_t A_KADV = A_MN ? 128 : 2, B_KADV = B_MN ? 128 : 2;  // in 16-byte units
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for_each_vtile<BN, CG>(sc, [&](int, int, int, int kb0, int kb1, bool, bool) {
      mbar_wait(&bars->tmem_empty[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem + (uint32_t)(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + NT * A_PLANE;
          uint64_t da[NT], db[NT];
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            da[t] = make_desc(sa + t * A_PLANE, A_LBO, 1024);
            db[t] = make_desc(sb + t * C::B_PLANE, B_LBO, 1024);
          }
          bool first = kb == kb0;
          // products a_i * b_j with i + j <= NT - 1 (0-based), smallest terms first
#pragma unroll
          for (int sum = NT - 1; sum >= 0; --sum) {
#pragma unroll
            for (int i = 0; i <= sum; ++i) {
              const int j = sum - i;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                if (CG == 2) umma_bf16_cg2(d_tmem, da[i] + A_KADV * k, db[j] + B_KADV * k, IDESC, first ? 0u : 1u);
                else umma_bf16(d_tmem, da[i] + A_KADV * k, db[j] + B_KADV * k, IDESC, first ? 0u : 1u);
                first = false;
              }
            }
          }
          if (CG == 2) umma_commit_cg2(&bars->empty[stage]); else umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) {
        if (CG == 2) umma_commit_cg2(&bars->tmem_full[as]); else umma_commit(&bars->tmem_full[as]);
      }
      __syncwarp();
      if (++as == NACC) {
        as = 0;
        aphase ^= 1u;
      }
    });
    }
  } else {
    // ===================================================== epilogue =====================================================
    const int q = warp & 3;  // TMEM sub-partition of this warp: accumulator rows 32q .. 32q+31
    const EpiArgs& e = P.epi;
    int as = 0;
    uint32_t aphase = 0;
    // EPI_LIK2: the partial of this thread's accumulator row over the K split (CW columns: the whole tile, or half of a
    // 256-wide one -- warps 6-9 take columns [128, 256))
    constexpr int CW = EPI == EPI_LIK2 ? BN / (epi_warps(EPI, BN) / 4) : 1;
    const int cw0 = EPI == EPI_LIK2 ? ((warp - 2) >> 2) * CW : 0;
    float part[CW];
    for_each_vtile<BN, CG>(sc, [&](int m_blk, int n_blk, int split, int kb0, int kb1, bool first_vt, bool last_vt) {
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_acc = tmem + (uint32_t)(as * BN) + ((uint32_t)(32 * q) << 16);
      const bool empty = kb1 == kb0;
      const int64_t r = (int64_t)m_blk * BMC + rank * BM + 32 * q + lane;  // accumulator row of this thread
      if (EPI == EPI_LIK1) {
        const bool row_ok = r < e.n_rows;
        const float yv = row_ok ? e.y[r] : 0.0f;
        const bool logit = e.family == BJX_LIK_BERNOULLI_LOGITS;
        float* statp = e.statpart + ((size_t)blockIdx.x * 4 + q) * e.S;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int s0 = n_blk * BN + c * 32;
          if (s0 >= e.S) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(t_acc + c * 32, v);
          tmem_ld_wait();
          float f[32];
          uint32_t p1[16], p2[16], p3[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float g[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const bool ok = row_ok && (s0 + j + u) < e.S;
              const float x = __uint_as_float(v[j + u]);
              if (logit) {
                const float ex = __expf(-fabsf(x));
                const float ope = 1.0f + ex;
                const float rc = __fdividef(1.0f, ope);
                f[j + u] = ok ? (yv * x - fmaxf(x, 0.0f) - __logf(ope)) : 0.0f;
                g[u] = ok ? (yv - (x >= 0.0f ? rc : ex * rc)) : 0.0f;
              } else {
                const float rr = yv - x;
                f[j + u] = ok ? rr * rr : 0.0f;
                g[u] = ok ? rr : 0.0f;
              }
            }
            split2(g[0], g[1], p1[j >> 1], p2[j >> 1]);
            if (NT == 3) {  // third term: what the first two left over
              const float q0 = (g[0] - __uint_as_float(p1[j >> 1] << 16)) - __uint_as_float(p2[j >> 1] << 16);
              const float q1 = (g[1] - __uint_as_float(p1[j >> 1] & 0xFFFF0000u)) - __uint_as_float(p2[j >> 1] & 0xFFFF0000u);
              p3[j >> 1] = pack_bf16x2(q0, q1);
            }
          }
          if (row_ok) {
            uint4* d1 = reinterpret_cast<uint4*>(e.g1 + r * e.g_pitch + s0);
            uint4* d2 = reinterpret_cast<uint4*>(e.g2 + r * e.g_pitch + s0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              d1[k] = make_uint4(p1[4 * k], p1[4 * k + 1], p1[4 * k + 2], p1[4 * k + 3]);
              d2[k] = make_uint4(p2[4 * k], p2[4 * k + 1], p2[4 * k + 2], p2[4 * k + 3]);
            }
            if (NT == 3) {
              uint4* d3 = reinterpret_cast<uint4*>(e.g3 + r * e.g_pitch + s0);
#pragma unroll
              for (int k = 0; k < 4; ++k) d3[k] = make_uint4(p3[4 * k], p3[4 * k + 1], p3[4 * k + 2], p3[4 * k + 3]);
            }
          }
          const float tot = transpose_reduce32(f, lane);
          // only this thread ever adds to this address and its reductions are applied in program order: deterministic
          if (s0 + lane < e.S) atomicAdd(statp + s0 + lane, tot);
        }
      } else if (EPI == EPI_LIK2) {
        // accumulator rows = weight column d, columns = sample s.  The accumulator of every virtual tile (the tensor core's
        // fp32 accumulate truncates, so the chain is cut every kb_per_vtile k-blocks) is added into a partial that stays in
        // REGISTERS for the whole K split and is stored once: Gpart[split][s][d] = partial (lanes contiguous in d: coalesced).
        // The same fp32 additions in the same order as the red.global.add updates this replaces (bit-identical), without
        // their DRAM traffic -- 1.7 GB of writes and 3.7 GB of extra reads per c4 step.
        if (first_vt) {
#pragma unroll
          for (int j = 0; j < CW; ++j) part[j] = 0.0f;
        }
        if (!empty) {  // warp-uniform; tcgen05.ld is warp-collective
#pragma unroll
          for (int c = 0; c < CW / 32; ++c) {
            if (n_blk * BN + cw0 + c * 32 < e.S) {  // warp-uniform
              uint32_t v[32];
              tmem_ld32(t_acc + cw0 + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) part[c * 32 + j] += __uint_as_float(v[j]);
            }
          }
        }
        if (last_vt && r < e.Dw) {
          float* gp = e.Gpart + (size_t)split * e.S * e.Dw + r;
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int sj = n_blk * BN + cw0 + j;
            if (sj < e.S) gp[(size_t)sj * e.Dw] = part[j];
          }
        }
      } else if (EPI == EPI_Z) {
        // z[s, i] = loc[i] + acc   (rows = samples, columns = latent coordinate i)
        {
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            const int64_t i0 = (int64_t)n_blk * BN + c * 32;
            if (i0 >= e.D) break;
            uint32_t v[32];
            if (!empty) {  // warp-uniform
              tmem_ld32(t_acc + c * 32, v);
              tmem_ld_wait();
            }
            if (r < e.S) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (i0 + j < e.D) e.out[r * e.D + i0 + j] = e.loc[i0 + j] + (empty ? 0.0f : __uint_as_float(v[j]));
            }
          }
        }
      } else {
        // transposed tile of dM: accumulator rows = j (column of dM), columns = i (row of dM), so that for a fixed i the 32
        // lanes write -- and read M at -- consecutive addresses
        {
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            const int64_t i0 = (int64_t)n_blk * BN + c * 32;
            if (i0 >= e.D) break;
            uint32_t v[32];
            if (!empty) {  // warp-uniform
              tmem_ld32(t_acc + c * 32, v);
              tmem_ld_wait();
            }
            if (r < e.D) {
              // The raw scale enters only through d sigma / d raw (1 for the Identity bijector of get_full_rank's default:
              // M is then read on the diagonal alone) and, on the diagonal, through d log sigma / d raw.  With four warps
              // per SM this loop is a latency chain, so the division stays off the 16383 off-diagonal entries of a tile.
              const bool need_m = e.scale_bij != BJX_SCALE_IDENTITY;
              float mv[32];
              if (need_m) {  // all 32 loads first (independent, in flight together): the stores below may alias them for the compiler
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int64_t i = i0 + j;
                  mv[j] = (i < e.D && r <= i) ? __ldg(e.M + i * e.D + r) : 0.0f;
                }
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int64_t i = i0 + j;
                if (i < e.D) {
                  float val = 0.0f;
                  if (r <= i) {
                    val = empty ? 0.0f : __uint_as_float(v[j]);
                    if (need_m) val *= eval_scale(mv[j], e.scale_bij).dsigma;
                    if (r == i) {
                      const ScaleEval se = eval_scale(need_m ? mv[j] : __ldg(e.M + i * e.D + r), e.scale_bij);
                      val -= e.prior_weight * se.dlogsigma;
                    }
                  }
                  e.out[i * e.D + r] = val;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_leader(&bars->tmem_empty[as]); else mbar_arrive(&bars->tmem_empty[as]);
      }
      if (++as == NACC) {
        as = 0;
        aphase ^= 1u;
      }
    });
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer may still be reading accumulators the leader's MMAs wrote
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_cg2(tmem, C::TMEM_COLS); else tmem_dealloc(tmem, C::TMEM_COLS);
  }
}

// ---- fp32 -> bf16 split planes --------------------------------------------------------------------------------------
// out planes are [rows, pitch] row-major bf16 with columns [cols, pitch) zero-filled; one thread per 8 columns.
struct SrcPlain {  // src[r * ld + c]
  const float* p;
  int64_t ld;
  __device__ __forceinline__ float operator()(int64_t r, int64_t c) const { return p[r * ld + c]; }
};
struct SrcTril {  // scale(M[i, j]) for j <= i, else 0   (tfd.MultivariateNormalTriL masks with band_part)
  const float* M;
  int64_t D;
  int scale_bij;
  __device__ __forceinline__ float operator()(int64_t i, int64_t j) const {
    return j <= i ? eval_scale(M[i * D + j], scale_bij).sigma : 0.0f;
  }
};

template <int NT, class Src>
__global__ void __launch_bounds__(256) k_split_planes(Src src, int64_t rows, int64_t cols, int64_t pitch, __nv_bfloat16* o1,
                                                      __nv_bfloat16* o2, __nv_bfloat16* o3) {
  const int64_t per_row = pitch / 8;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * per_row) return;
  const int64_t r = idx / per_row, c0 = (idx % per_row) * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? src(r, c0 + j) : 0.0f;
  uint32_t a[4], b[4], c[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x0 = v[2 * j], x1 = v[2 * j + 1];
    a[j] = tc::pack_bf16x2(x0, x1);
    const float r0 = x0 - __uint_as_float(a[j] << 16), r1 = x1 - __uint_as_float(a[j] & 0xFFFF0000u);
    b[j] = tc::pack_bf16x2(r0, r1);
    if (NT == 3) {
      const float q0 = r0 - __uint_as_float(b[j] << 16), q1 = r1 - __uint_as_float(b[j] & 0xFFFF0000u);
      c[j] = tc::pack_bf16x2(q0, q1);
    }
  }
  const size_t o = (size_t)r * pitch + c0;
  *reinterpret_cast<uint4*>(o1 + o) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(o2 + o) = make_uint4(b[0], b[1], b[2], b[3]);
  if (NT == 3) *reinterpret_cast<uint4*>(o3 + o) = make_uint4(c[0], c[1], c[2], c[3]);
}

// specialised fast path for the big operand: X [rows, ld] with 16-byte aligned rows, cols % 8 == 0 handled by the tail mask
__global__ void __launch_bounds__(256) k_split_x(const float* __restrict__ X, int64_t ld, int64_t rows, int64_t cols, int64_t pitch,
                                                 __nv_bfloat16* __restrict__ o1, __nv_bfloat16* __restrict__ o2) {
  const int64_t per_row = pitch / 8;
  const int64_t total = rows * per_row;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int64_t r = idx / per_row, c0 = (idx % per_row) * 8;
    float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
    if (c0 + 8 <= cols) {
      const float4* p = reinterpret_cast<const float4*>(X + r * ld + c0);
      u0 = __ldcs(p);
      u1 = __ldcs(p + 1);
    } else {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = (c0 + j < cols) ? X[r * ld + c0 + j] : 0.0f;
      u0 = make_float4(t[0], t[1], t[2], t[3]);
      u1 = make_float4(t[4], t[5], t[6], t[7]);
    }
    uint32_t a[4], b[4];
    tc::split2(u0.x, u0.y, a[0], b[0]);
    tc::split2(u0.z, u0.w, a[1], b[1]);
    tc::split2(u1.x, u1.y, a[2], b[2]);
    tc::split2(u1.z, u1.w, a[3], b[3]);
    const size_t o = (size_t)r * pitch + c0;
    *reinterpret_cast<uint4*>(o1 + o) = make_uint4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<uint4*>(o2 + o) = make_uint4(b[0], b[1], b[2], b[3]);
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 matrix [rows, pitch] row-major, logical width `cols`; box = box_cols x box_rows, 128-byte swizzle, zero OOB fill
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t pitch, int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(BJX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(BJX_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
  return BJX_OK;
}

}  // namespace gemm
int make_bf16_tensor_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t pitch, int box_cols, int box_rows) {
  return gemm::make_map(m, base, rows, cols, pitch, box_cols, box_rows);
}
namespace gemm {

template <int BN, int NT, bool A_MN, bool B_MN, int EPI, int CG = 1>
static int launch(const Params& P, int grid, cudaStream_t st) {
  using C = Cfg<BN, NT, CG>;
  static DeviceOnce attr_set;
  if (!attr_set.test_and_set())
    BJX_CUDA_TRY(cudaFuncSetAttribute(k_gemm_tc<BN, NT, A_MN, B_MN, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  if (CG == 1) {
    k_gemm_tc<BN, NT, A_MN, B_MN, EPI, CG><<<grid, threads_of(EPI, BN), C::SMEM_BYTES, st>>>(P);
  } else {
    // CTA pairs: a cluster of two CTAs (same TPC) per 256-row tile
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(threads_of(EPI, BN));
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BJX_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_gemm_tc<BN, NT, A_MN, B_MN, EPI, CG>, P));
  }
  BJX_LAUNCH_CHECK("k_gemm_tc");
  return BJX_OK;
}

static int64_t pad64(int64_t v) { return (v + 63) / 64 * 64; }

}  // namespace gemm

// X [N, ldx] fp32 -> two bf16 planes [N, pad64(Dw)] back to back (x = x1 + x2 to 2^-17 relative)
int split_x_planes(const float* X, int64_t N, int64_t Dw, int64_t ldx, void* planes, cudaStream_t st) {
  using namespace gemm;
  if (N <= 0) return BJX_OK;
  const int64_t Dp = pad64(Dw);
  __nv_bfloat16* x1 = (__nv_bfloat16*)planes;
  __nv_bfloat16* x2 = x1 + (size_t)N * Dp;
  const int64_t total = N * (Dp / 8);
  if ((ldx % 4) == 0 && ((uintptr_t)X & 15) == 0) {
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_split_x<<<(unsigned)blocks, 256, 0, st>>>(X, ldx, N, Dw, Dp, x1, x2);
  } else {
    k_split_planes<2, SrcPlain><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(SrcPlain{X, ldx}, N, Dw, Dp, x1, x2, nullptr);
  }
  BJX_LAUNCH_CHECK("k_split_x");
  return BJX_OK;
}

// rows [row0, row0 + rows) of the planes of an N_total-row dataset, from a chunk X_rows [rows, ldx] holding just those rows
int split_x_planes_rows(const float* X_rows, int64_t rows, int64_t Dw, int64_t ldx, void* planes, int64_t N_total, int64_t row0,
                        cudaStream_t st) {
  using namespace gemm;
  if (rows <= 0) return BJX_OK;
  const int64_t Dp = pad64(Dw);
  __nv_bfloat16* x1 = (__nv_bfloat16*)planes + (size_t)row0 * Dp;
  __nv_bfloat16* x2 = (__nv_bfloat16*)planes + (size_t)N_total * Dp + (size_t)row0 * Dp;
  const int64_t total = rows * (Dp / 8);
  if ((ldx % 4) == 0 && ((uintptr_t)X_rows & 15) == 0) {
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_split_x<<<(unsigned)blocks, 256, 0, st>>>(X_rows, ldx, rows, Dw, Dp, x1, x2);
  } else {
    k_split_planes<2, SrcPlain><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(SrcPlain{X_rows, ldx}, rows, Dw, Dp, x1, x2, nullptr);
  }
  BJX_LAUNCH_CHECK("k_split_x");
  return BJX_OK;
}

size_t x_planes_bytes(int64_t N, int64_t Dw) { return (size_t)2 * (size_t)N * (size_t)gemm::pad64(Dw) * 2; }

// =========================================================================================================================
// Likelihood, two passes: eta + link (EPI_LIK1), then the transposed contraction (EPI_LIK2).
bool gemm_eligible(const BjxElboArgs& a) {
  if (a.likelihood != BJX_LIK_BERNOULLI_LOGITS && a.likelihood != BJX_LIK_GAUSSIAN) return false;
  if (a.Dw < 64 || a.S < 1 || a.S > 4096 || a.N < 1) return false;
  if (a.N >= (1ll << 31) - 256 || a.Dw >= (1ll << 31) - 256) return false;  // 32-bit TMA coordinates
  return true;  // BJX_PREC_BF16X3 (three-term split, fp32-accurate) runs here with 128-wide tiles
}

static int gemm_nt(const BjxElboArgs& a) { return a.precision == BJX_PREC_BF16X3 ? 3 : 2; }
static int lik1_bn(int64_t S, int nt = 2) { return nt == 3 ? 128 : (S <= 64 ? 64 : (S <= 128 ? 128 : 256)); }

// CTA pairs (cta_group::2) for the widest tiles: the pair shares the B operand, which halves its shared-memory reads and
// its TMA writes per SM -- the port both compete for (BJX_GEMM_CG=1 keeps single-CTA tiles, for A/B measurements)
static int gemm_cg(int bn) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("BJX_GEMM_CG");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 1 || forced == 2) return bn == 256 ? forced : 1;
  return bn == 256 ? 2 : 1;
}

int gemm_plan(const BjxElboArgs& a, Plan& p) {
  const int sms = sm_count();
  if (sms <= 0) return fail(BJX_ERR_CUDA, "no CUDA device");
  const int nt = gemm_nt(a);
  const int bn = lik1_bn(a.S, nt);
  const int cg = nt == 3 ? 1 : gemm_cg(bn);
  const int groups = sms / cg;  // CTAs, or CTA pairs
  const int bmc = 128 * cg;
  const int64_t tiles1 = ((a.N + bmc - 1) / bmc) * ((a.S + bn - 1) / bn);
  p.gm_grid1 = cg * (int)(tiles1 < groups ? tiles1 : groups);
  const int64_t tiles2 = ((a.Dw + bmc - 1) / bmc) * ((a.S + bn - 1) / bn);
  const int64_t kb_total = (a.N + 63) / 64;
  int64_t splits = tiles2 >= groups ? 1 : groups / tiles2;
  const int64_t max_splits = (kb_total + 15) / 16;  // at least one full (2-term) virtual tile per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.gm_splits = (int)splits;
  int64_t units = tiles2 * splits;
  p.gm_grid2 = cg * (int)(units < groups ? units : groups);
  p.gm_cg = cg;
  p.gm_nt = nt;
  p.gm_Dp = gemm::pad64(a.Dw);
  p.gm_Sp = gemm::pad64(a.S);
  return BJX_OK;
}

int launch_lik_gemm(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st) {
  using namespace gemm;
  const float* theta = (const float*)(ws + p.o_theta);
  float* Gpart = (float*)(ws + p.o_Gpart);
  float* statpart = (float*)(ws + p.o_statpart);
  const int64_t Dp = p.gm_Dp, Sp = p.gm_Sp;
  const int nt = p.gm_nt;  // split terms per operand: 2 (BJX_PREC_BF16X2 / auto) or 3 (BJX_PREC_BF16X3, fp32-accurate)
  // prepared planes hold two terms: the three-term mode always splits inside the step
  const bool prepared = a.x_planes != nullptr && nt == 2;
  __nv_bfloat16* xp[3];
  __nv_bfloat16* tp[3];
  __nv_bfloat16* gp[3];
  for (int t = 0; t < 3; ++t) {
    xp[t] = (prepared ? (__nv_bfloat16*)a.x_planes : (__nv_bfloat16*)(ws + p.o_gm_x)) + (size_t)t * a.N * Dp;
    tp[t] = (__nv_bfloat16*)(ws + p.o_gm_t) + (size_t)t * a.S * Dp;
    gp[t] = (__nv_bfloat16*)(ws + p.o_gm_g) + (size_t)t * a.N * Sp;
  }

  // operand planes
  if (!prepared) {
    if (nt == 2) {
      int rc = split_x_planes(a.X, a.N, a.Dw, a.ldx, xp[0], st);
      if (rc) return rc;
    } else {
      const int64_t total = a.N * (Dp / 8);
      k_split_planes<3, SrcPlain><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(SrcPlain{a.X, a.ldx}, a.N, a.Dw, Dp, xp[0], xp[1], xp[2]);
      BJX_LAUNCH_CHECK("k_split_planes(X)");
    }
  }
  {
    const int64_t tt = a.S * (Dp / 8);
    if (nt == 2)
      k_split_planes<2, SrcPlain><<<(unsigned)((tt + 255) / 256), 256, 0, st>>>(SrcPlain{theta + a.w_offset, a.D}, a.S, a.Dw, Dp, tp[0], tp[1], nullptr);
    else
      k_split_planes<3, SrcPlain><<<(unsigned)((tt + 255) / 256), 256, 0, st>>>(SrcPlain{theta + a.w_offset, a.D}, a.S, a.Dw, Dp, tp[0], tp[1], tp[2]);
    BJX_LAUNCH_CHECK("k_split_planes(theta)");
  }
  BJX_CUDA_TRY(cudaMemsetAsync(statpart, 0, (size_t)p.PS * a.S * sizeof(float), st));
  // (Gpart needs no clearing: every [split][s][d] entry is stored exactly once by the unit that owns it)
  const int bn = lik1_bn(a.S, nt);
  const int cg = p.gm_cg, bmc = 128 * cg;

  // ---- pass 1: eta, link, g planes, stat
  {
    Params P;
    memset(&P, 0, sizeof(P));
    int rc;
    for (int t = 0; t < nt; ++t) {
      if ((rc = make_map(&P.mapA[t], xp[t], a.N, a.Dw, Dp, 64, 128))) return rc;
      if ((rc = make_map(&P.mapB[t], tp[t], a.S, a.Dw, Dp, 64, bn / cg))) return rc;  // a pair loads half of the B tile per CTA
    }
    P.sched.tiles_m = (int)((a.N + bmc - 1) / bmc);
    P.sched.tiles_n = (int)((a.S + bn - 1) / bn);
    P.sched.splits = 1;
    P.sched.kb_total = (int)(Dp / 64);
    P.sched.kb_per_split = P.sched.kb_total;
    P.sched.kb_per_vtile = P.sched.kb_total;
    P.sched.tril = 0;
    P.epi.y = a.y;
    P.epi.n_rows = a.N;
    P.epi.S = (int)a.S;
    P.epi.family = a.likelihood;
    P.epi.g1 = gp[0];
    P.epi.g2 = gp[1];
    P.epi.g3 = nt == 3 ? gp[2] : nullptr;
    P.epi.g_pitch = Sp;
    P.epi.statpart = statpart;
    if (nt == 3) rc = launch<128, 3, false, false, EPI_LIK1>(P, p.gm_grid1, st);
    else if (bn == 64) rc = launch<64, 2, false, false, EPI_LIK1>(P, p.gm_grid1, st);
    else if (bn == 128) rc = launch<128, 2, false, false, EPI_LIK1>(P, p.gm_grid1, st);
    else if (cg == 2) rc = launch<256, 2, false, false, EPI_LIK1, 2>(P, p.gm_grid1, st);
    else rc = launch<256, 2, false, false, EPI_LIK1>(P, p.gm_grid1, st);
    if (rc) return rc;
  }
  // ---- pass 2: G^T[d, s] = sum_n X[n, d] g[n, s]   (A = X planes, B = g planes, both MN-major; K = data rows)
  {
    Params P;
    memset(&P, 0, sizeof(P));
    int rc;
    for (int t = 0; t < nt; ++t) {
      if ((rc = make_map(&P.mapA[t], xp[t], a.N, a.Dw, Dp, 64, 64))) return rc;
      if ((rc = make_map(&P.mapB[t], gp[t], a.N, a.S, Sp, 64, 64))) return rc;
    }
    P.sched.tiles_m = (int)((a.Dw + bmc - 1) / bmc);
    P.sched.tiles_n = (int)((a.S + bn - 1) / bn);
    P.sched.splits = p.gm_splits;
    P.sched.kb_total = (int)((a.N + 63) / 64);
    P.sched.kb_per_split = (P.sched.kb_total + p.gm_splits - 1) / p.gm_splits;
    // the accumulation chain in TMEM truncates (about 1.4e-7 relative per 64-row k-block and product triple): the
    // fp32-accurate mode hands the accumulator to the epilogue four times as often
    P.sched.kb_per_vtile = nt == 3 ? 4 : 16;
    P.sched.tril = 0;
    P.epi.S = (int)a.S;
    P.epi.Gpart = Gpart;
    P.epi.Dw = a.Dw;
    if (nt == 3) rc = launch<128, 3, true, true, EPI_LIK2>(P, p.gm_grid2, st);
    else if (bn == 64) rc = launch<64, 2, true, true, EPI_LIK2>(P, p.gm_grid2, st);
    else if (bn == 128) rc = launch<128, 2, true, true, EPI_LIK2>(P, p.gm_grid2, st);
    else if (cg == 2) rc = launch<256, 2, true, true, EPI_LIK2, 2>(P, p.gm_grid2, st);
    else rc = launch<256, 2, true, true, EPI_LIK2>(P, p.gm_grid2, st);
    if (rc) return rc;
  }
  return BJX_OK;
}

// =========================================================================================================================
// Full-rank contractions on the tensor cores (3-term split: fp32-accurate).
bool fullrank_tc_eligible(const BjxElboArgs& a) { return a.vi_type == BJX_VI_FULL_RANK && a.D >= 128 && a.D < (1ll << 30) && a.S < (1ll << 30); }

size_t fullrank_tc_bytes(const BjxElboArgs& a) {
  const int64_t Dp = gemm::pad64(a.D);
  // eps planes [S, Dp] x 3, L planes [D, Dp] x 3 (reused for gz planes [S, Dp] x 3)
  const size_t big = (size_t)(a.D > a.S ? a.D : a.S);
  return align_up((size_t)3 * a.S * Dp * 2, 256) + align_up((size_t)3 * big * Dp * 2, 256);
}

// z[s, i] = loc[i] + sum_{j <= i} scale(M[i, j]) eps[s, j]
int launch_fullrank_z_tc(const BjxElboArgs& a, const float* eps, float* z, char* scratch, cudaStream_t st) {
  using namespace gemm;
  const int64_t Dp = pad64(a.D);
  __nv_bfloat16* e0 = (__nv_bfloat16*)scratch;
  __nv_bfloat16* l0 = (__nv_bfloat16*)(scratch + align_up((size_t)3 * a.S * Dp * 2, 256));
  const size_t eplane = (size_t)a.S * Dp, lplane = (size_t)a.D * Dp;
  const int64_t te = a.S * (Dp / 8), tl = a.D * (Dp / 8);
  k_split_planes<3, SrcPlain><<<(unsigned)((te + 255) / 256), 256, 0, st>>>(SrcPlain{eps, a.D}, a.S, a.D, Dp, e0, e0 + eplane, e0 + 2 * eplane);
  BJX_LAUNCH_CHECK("k_split_planes(eps)");
  k_split_planes<3, SrcTril><<<(unsigned)((tl + 255) / 256), 256, 0, st>>>(SrcTril{a.raw_scale, a.D, a.scale_bijector}, a.D, a.D, Dp, l0, l0 + lplane, l0 + 2 * lplane);
  BJX_LAUNCH_CHECK("k_split_planes(L)");
  Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  for (int t = 0; t < 3; ++t) {
    if ((rc = make_map(&P.mapA[t], e0 + t * eplane, a.S, a.D, Dp, 64, 128))) return rc;
    if ((rc = make_map(&P.mapB[t], l0 + t * lplane, a.D, a.D, Dp, 64, 128))) return rc;
  }
  P.sched.tiles_m = (int)((a.S + 127) / 128);
  P.sched.tiles_n = (int)((a.D + 127) / 128);
  P.sched.splits = 1;
  P.sched.kb_total = (int)(Dp / 64);
  P.sched.kb_per_split = P.sched.kb_total;
  P.sched.kb_per_vtile = P.sched.kb_total;
  P.sched.tril = 1;
  P.epi.S = (int)a.S;
  P.epi.D = a.D;
  P.epi.loc = a.loc;
  P.epi.out = z;
  const int64_t tiles = (int64_t)P.sched.tiles_m * P.sched.tiles_n;
  const int sms = sm_count();
  return launch<128, 3, false, false, EPI_Z>(P, (int)(tiles < sms ? tiles : sms), st);
}

// dM[i, j] = [j <= i] scale'(M_ij) sum_s gz[s, i] eps[s, j] - [i == j] prior_weight dlog|L_ii|/draw_ii
// (eps planes are still in `scratch` from launch_fullrank_z_tc of the same step)
int launch_fullrank_dM_tc(const BjxElboArgs& a, const float* gz, float* dM, char* scratch, float entropy_weight, cudaStream_t st) {
  using namespace gemm;
  const int64_t Dp = pad64(a.D);
  __nv_bfloat16* e0 = (__nv_bfloat16*)scratch;
  __nv_bfloat16* g0 = (__nv_bfloat16*)(scratch + align_up((size_t)3 * a.S * Dp * 2, 256));
  const size_t plane = (size_t)a.S * Dp;
  const int64_t tg = a.S * (Dp / 8);
  k_split_planes<3, SrcPlain><<<(unsigned)((tg + 255) / 256), 256, 0, st>>>(SrcPlain{gz, a.D}, a.S, a.D, Dp, g0, g0 + plane, g0 + 2 * plane);
  BJX_LAUNCH_CHECK("k_split_planes(gz)");
  Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  for (int t = 0; t < 3; ++t) {  // A = eps (rows of the tile = j), B = gz (columns = i): the transposed tile of dM
    if ((rc = make_map(&P.mapA[t], e0 + t * plane, a.S, a.D, Dp, 64, 64))) return rc;
    if ((rc = make_map(&P.mapB[t], g0 + t * plane, a.S, a.D, Dp, 64, 64))) return rc;
  }
  P.sched.tiles_m = (int)((a.D + 127) / 128);
  P.sched.tiles_n = (int)((a.D + 127) / 128);
  P.sched.splits = 1;
  P.sched.kb_total = (int)((a.S + 63) / 64);
  P.sched.kb_per_split = P.sched.kb_total;
  P.sched.kb_per_vtile = P.sched.kb_total;
  P.sched.tril = 2;
  P.epi.S = (int)a.S;
  P.epi.D = a.D;
  P.epi.M = a.raw_scale;
  P.epi.out = dM;
  P.epi.scale_bij = a.scale_bijector;
  P.epi.prior_weight = entropy_weight;  // weight of the -d log|L_ii| / d raw_ii constant
  const int64_t tiles = (int64_t)P.sched.tiles_m * P.sched.tiles_n;
  const int sms = sm_count();
  return launch<128, 3, true, true, EPI_DM>(P, (int)(tiles < sms ? tiles : sms), st);
}

}  // namespace bjx

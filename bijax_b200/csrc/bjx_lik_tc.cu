// Kernel 5+6 on the 5th-generation tensor cores (tcgen05 + TMEM), one pass over X.
//
// For a tile of 64 data rows the CTA computes
//     eta[n, s]  = sum_d X[n, d] theta[s, d]              GEMM1  (M = 64 rows, N = samples, K = D)
//     g[n, s]    = d/d eta log p(y_n | eta)                epilogue from TMEM (sigmoid / residual), stat[s] += f
//     G^T[d, s] += sum_n X[n, d] g[n, s]                   GEMM2  (M = 128-column chunk of D, N = samples, K = 64 rows)
// so the transposed contraction of the VJP (utils.py:7) reuses the X tile while it is still in shared memory: X is
// read from HBM exactly once per step (4*N*D + 4*N algorithmic bytes).
//
// fp32 inputs are not representable on the tensor cores, so every operand is split on the fly into bf16 terms
// (x = x1 + x2 + ..., |x2| <= 2^-9 |x1|): products x1*t1 + x1*t2 + x2*t1 (+ x2*t2) accumulate in fp32 in TMEM.  The X
// tile is converted in registers by the loader warps straight after the coalesced 128-bit global loads and stored in the
// canonical UMMA shared-memory layout (128-byte swizzle), which serves GEMM1 as a K-major A operand and GEMM2 as an
// MN-major A operand without a second copy or a transpose.
//
// Warp roles (800 threads, 1 CTA per SM, persistent over row tiles):
//   warps 0-15  loaders: 8 x LDG.128 per lane (a quarter block per warp; 16 warps = 64 KB in flight per SM)
//               -> bf16 split -> STS.128 swizzled -> mbarrier.  No software pipelining inside a warp: ptxas shares
//               scoreboards between prefetch stages, which serialises them; depth comes from the number of warps.
//   warp  16    MMA issuer (converged warp, elect.sync), owns the TMEM allocation
//   warps 17-24 epilogue: tcgen05.ld logits -> link function -> bf16 split of g -> shared B operand of GEMM2
//
// Kernels in this file (launch_lik_tc picks one per shape; all read X exactly once):
//   k_lik_tc<NBLK, IDX>                  raw fp32 X, converted to the split-bf16 image in registers (described above)
//   k_lik_tc_tma<NBLK, EW, IDX, PAIR, NG> prepared planes, TMA-fed, 64-row tiles: 128 < D <= 512 (ring of tiles at D <= 256, spare
//                                        slot above); IDX = rows gathered through an index with cp.async; PAIR = a cluster of two
//                                        CTAs per tile, half of the columns each, for 512 < D <= 1024; NG = 2 epilogue groups
//   k_lik_tc_tall<NBLK>                  prepared planes, TMA-fed, 128-row tiles without the partner swap: D <= 64 / D <= 128
#include <cuda.h>

#include <cstdlib>

#include "bjx_elbo.cuh"
#include "bjx_tc_ptx.cuh"

namespace bjx {

namespace tc {

constexpr int TILE_ROWS = 64;      // rows per tile = UMMA M of GEMM1
constexpr int BLK_COLS = 64;       // columns per shared-memory block (64 bf16 = one 128-byte swizzle row)
constexpr int BLK_BYTES = TILE_ROWS * 128;  // 8 KB
constexpr int NS = 32;             // samples per pass (padded)
constexpr int NB = 64;             // UMMA N when both split terms of the B operand are used ([t1; t2] / [e1; e2])
constexpr int LOADER_WARPS = 16;   // each owns a quarter block (16 rows x 64 cols = 4 KB) at a time, no intra-warp pipelining
constexpr int MMA_WARP = 16;
constexpr int EPI_WARP0 = 17, EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;  // 800
constexpr int TMEM_COLS = 512;
constexpr int ACC1_COL = 0;        // 64 columns: logits [x*t1 | x1*t2]
constexpr int ACC2_COL = 128;      // 4 chunks x 64 columns: G^T [x*e1 | x1*e2]
// The tensor core's fp32 accumulate truncates: a TMEM accumulation chain picks up a systematic bias of ~1.4e-7 (relative)
// per tile (measured, scripts/accum_check.py: 1.2e-4 on the gradient after 845 tiles).  The GEMM2 accumulators are
// therefore flushed into an fp32 partial in global memory every ACC2_FLUSH tiles and the chain is restarted, which
// bounds the bias near 2e-6.  The four 128-column chunks are drained in rotation (one every ACC2_FLUSH/4 tiles) so that
// the work fits in the epilogue warps' idle window.
constexpr int ACC2_FLUSH = 16;

struct Bars {
  unsigned long long x_full[8], x_empty[8];
  unsigned long long acc1_full, dl_full, dl_empty, done;
  uint32_t tmem_base;
};

// Link functions with ONE exponential per datum (MUFU.EX2 / LG2 / RCP): e = exp(-|eta|),
// softplus(eta) = max(eta, 0) + log(1 + e), sigmoid(eta) = (eta >= 0 ? 1 : e) / (1 + e).  Absolute error ~1e-7 per datum.
__device__ __forceinline__ void link_fast(int family, float y, float eta, float& f, float& g) {
  if (family == BJX_LIK_BERNOULLI_LOGITS) {
    const float e = __expf(-fabsf(eta));
    const float ope = 1.0f + e;
    const float r = __frcp_rn(ope);
    f = y * eta - (fmaxf(eta, 0.0f) + __logf(ope));
    g = y - (eta >= 0.0f ? r : e * r);
  } else {
    const float r = y - eta;
    f = r * r;
    g = r;
  }
}

}  // namespace tc

// =========================================================================================================================
// IDX: row r of the problem is row ridx[r] of X (minibatches drawn on the device, BjxElboArgs.row_index): the loaders read
// the dataset THROUGH the index, so a batch costs one pass over its own rows and nothing else.  y is already gathered.
template <int NBLK, bool IDX>  // D / 64
__global__ void __launch_bounds__(tc::THREADS, 1)
k_lik_tc(const float* __restrict__ X, int64_t ldx, const float* __restrict__ y, int64_t N, const float* __restrict__ theta,
         int64_t Dtot, int64_t w_off, int S, int family, float* __restrict__ Gpart, float* __restrict__ statpart,
         long long* __restrict__ trace, int stagger_cycles, int ablate, int pf_dist, int flush_tiles, int pf_mode,
         const int32_t* __restrict__ ridx) {
  if (IDX) pf_mode = 0;  // address-range prefetches make no sense through an index
  using namespace tc;
  // optional timeline of CTA 0, tiles [TRACE_T0, TRACE_T0 + 8): 64 clock64 slots per tile (bjx_debug_set_trace_buffer)
  constexpr int64_t TRACE_T0 = 16;
#define BJX_TRACE(it_, slot_)                                                                            \
  do {                                                                                                   \
    if (trace != nullptr && blockIdx.x == 0 && (it_) >= TRACE_T0 && (it_) < TRACE_T0 + 8)                \
      trace[((it_)-TRACE_T0) * 64 + (slot_)] = clock64();                                                \
  } while (0)
  constexpr int D = NBLK * BLK_COLS;
  constexpr int NCH = NBLK / 2;  // 128-column chunks (UMMA M of GEMM2)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sx1 = smem;                              // NBLK blocks of [64 rows][64 bf16], x1 term
  uint8_t* sx2 = sx1 + NBLK * BLK_BYTES;            // x2 term
  uint8_t* sth = sx2 + NBLK * BLK_BYTES;            // NBLK blocks of [64 rows: t1(32) | t2(32)][64 bf16]
  uint8_t* sdl = sth + NBLK * BLK_BYTES;            // [64 rows n][64 bf16: e1(32 s) | e2(32 s)]  (MN-major B of GEMM2)
  Bars* bars = (Bars*)(sdl + BLK_BYTES);
  float* sred = (float*)(bars + 1);                 // [4][32] stat reduction scratch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (N + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // ---- one-time setup ---------------------------------------------------------------------------------------------------
  if (tid == 0) {
    for (int b = 0; b < 8; ++b) {
      mbar_init(&bars->x_full[b], 4);  // the four quarter-block warps
      mbar_init(&bars->x_empty[b], 1);
    }
    mbar_init(&bars->acc1_full, 1);
    mbar_init(&bars->dl_full, EPI_WARPS);
    mbar_init(&bars->dl_empty, 1);
    mbar_init(&bars->done, 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  // theta -> bf16 split B operand of GEMM1 (K-major, 128B swizzle): row n = s (t1) / 32 + s (t2)
  for (int idx = tid; idx < NS * (D / 8); idx += THREADS) {
    const int s = idx / (D / 8), d8 = idx % (D / 8);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = s < S ? theta[(int64_t)s * Dtot + w_off + d8 * 8 + j] : 0.0f;
    uint32_t p1[4], p2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], p1[j], p2[j]);
    const int b = d8 / 8, c = d8 % 8;
    uint8_t* blk = sth + b * BLK_BYTES;
    *reinterpret_cast<uint4*>(blk + s * 128 + ((c ^ (s & 7)) << 4)) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    const int n2 = 32 + s;
    *reinterpret_cast<uint4*>(blk + n2 * 128 + ((c ^ (n2 & 7)) << 4)) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < LOADER_WARPS) {
    // ===================================================== loaders =====================================================
    // unit = quarter of a block: rows 16*qd .. 16*qd+15, 64 columns.  lane -> (row 4*i + lane/8, 16-byte chunk lane%8).
    const int qd = warp & 3, bl = warp >> 2;
    const int rl = lane >> 3, c = lane & 7;
    // De-phase the CTAs: with identical work per tile all SMs would issue their 64 KB load bursts in lockstep (a 9.5 MB
    // burst every tile period, ~3500-cycle load latency measured); a one-off start-up skew spreads the HBM demand.
    if (stagger_cycles > 0 && my_tiles > 4) {
      const long long t_start = clock64();
      const long long delay = (long long)((blockIdx.x * 37) % 16) * stagger_cycles;
      while (clock64() - t_start < delay) {
      }
    }
    // A warp set owns FIXED block indices (bl, bl + 4, ...) in every tile: the parity waits on x_empty[b] below are only
    // unambiguous if the same waiter meets every phase of the barrier.  (A flat round-robin over blocks made a set meet a
    // block only every other tile for D = 128 and 384, two phases apart: the wait could pass on a stale phase.)  For
    // D = 128 the sets 2 and 3 therefore idle.
    // IDX: source rows of this lane's four rows, for the current tile and (loaded one tile ahead) for the next one
    int32_t cur[4] = {0, 0, 0, 0}, nxt[4] = {0, 0, 0, 0};
    if (IDX && my_tiles > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = (int64_t)blockIdx.x * TILE_ROWS + 16 * qd + rl + 4 * i;
        nxt[i] = row < N ? __ldg(ridx + row) : 0;
      }
    }
    for (int64_t u = 0; u < my_tiles * 2; ++u) {
      const int64_t it = u >> 1;
      const int b = bl + 4 * (int)(u & 1);
      if (IDX && (u & 1) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          cur[i] = nxt[i];
          const int64_t row = (blockIdx.x + (it + 1) * gridDim.x) * TILE_ROWS + 16 * qd + rl + 4 * i;
          nxt[i] = (it + 1 < my_tiles && row < N) ? __ldg(ridx + row) : 0;
        }
      }
      if (b >= NBLK) continue;
      const int64_t row0 = (blockIdx.x + it * gridDim.x) * TILE_ROWS + 16 * qd + rl;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + 4 * i;
        if (row < N) {
          const int64_t src = IDX ? (int64_t)cur[i] : row;
          const float4* p = reinterpret_cast<const float4*>(X + src * ldx + b * BLK_COLS + c * 8);
          v[2 * i] = __ldg(p);
          v[2 * i + 1] = __ldg(p + 1);
        } else {
          v[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
          v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      // fire-and-forget L2 prefetch of the SAME unit of the next tile (16 rows x 256 B = 32 lines, one per lane): HBM keeps
      // streaming into the 126 MB L2 while this warp's registers are tied up waiting for the shared-memory slot
      if (IDX) {
        // the same unit of the NEXT tile, through its (already loaded) indices: lanes c = 0 and c = 4 own the two 128-byte
        // lines of a row's 256-byte block slice
        if (pf_dist > 0 && it + 1 < my_tiles && (c & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int64_t prow = (blockIdx.x + (it + 1) * gridDim.x) * TILE_ROWS + 16 * qd + rl + 4 * i;
            if (prow < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(X + (int64_t)nxt[i] * ldx + b * BLK_COLS + c * 8));
          }
        }
      } else if (pf_dist > 0 && pf_mode < 2) {  // pf_dist is in blocks ahead in this warp-set's own sequence (8 = the same unit of the next tile)
        const int64_t itp = it + (pf_dist >= 8 ? pf_dist / 8 : 1);  // the same unit, pf_dist / 8 tiles ahead
        const int64_t prow = (blockIdx.x + itp * gridDim.x) * TILE_ROWS + 16 * qd + (lane >> 1);
        if (itp < my_tiles && prow < N) {
          const float* pp = X + prow * ldx + b * BLK_COLS + (lane & 1) * 32;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
          if (pf_mode == 1) {  // touch every 32-byte sector of the 128-byte line
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + 8));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + 16));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + 24));
          }
        }
      }
      if (tid == 0 && b == 4) BJX_TRACE(it, 48);                // loads issued
      mbar_wait(&bars->x_empty[b], (uint32_t)((it & 1) ^ 1));  // GEMM2 of the previous tile released this block
      if (tid == 0) BJX_TRACE(it, b);
      if (trace != nullptr) {  // trace builds only: stamp the arrival of the last load of the unit
        asm volatile("" ::"f"(v[7].w), "f"(v[0].x) : "memory");
        if (tid == 0 && b == 4) BJX_TRACE(it, 49);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 16 * qd + rl + 4 * i;
        if ((ablate & 8) && v[2 * i].x != 12345.678f) continue;  // keep the loads alive, skip convert + store
        uint32_t p1[4], p2[4];
        split2(v[2 * i].x, v[2 * i].y, p1[0], p2[0]);
        split2(v[2 * i].z, v[2 * i].w, p1[1], p2[1]);
        split2(v[2 * i + 1].x, v[2 * i + 1].y, p1[2], p2[2]);
        split2(v[2 * i + 1].z, v[2 * i + 1].w, p1[3], p2[3]);
        const uint32_t off = b * BLK_BYTES + r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sx1 + off) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
        *reinterpret_cast<uint4*>(sx2 + off) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
      }
      // writer-side generic->async proxy fence.  It lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, which is cheap HERE
      // because this warp has no global loads in flight at this point (a MEMBAR drains them).
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->x_full[b]);
      if (tid == 0) BJX_TRACE(it, 8 + b);
      if (lane == 0 && b == 0) BJX_TRACE(it, 40 + warp);   // all four quarter warps of block 0
      if (lane == 0 && b == 1) BJX_TRACE(it, 44 + (warp & 3));
    }
  } else if (warp == MMA_WARP) {
    // ===================================================== MMA issuer ==================================================
    // The whole warp walks the pipeline (waits are warp-uniform); one elected lane issues.  Issued this way a small
    // UMMA is bound only by its operand bytes at 128 B/clk of shared-memory bandwidth (scripts/ubench/umma_rate.cu).
    constexpr uint32_t ID_G1_N64 = make_idesc(64, NB, 0, 0);   // x1 * [t1; t2]
    constexpr uint32_t ID_G1_N32 = make_idesc(64, NS, 0, 0);   // x2 * t1
    constexpr uint32_t ID_G2 = make_idesc(128, NB, 1, 1);      // x1^T * [e1; e2], both operands MN-major
    constexpr uint32_t ID_G2_N32 = make_idesc(128, NS, 1, 1);  // x2^T * e1
    const uint64_t dx1 = make_desc(smem_u32(sx1), 16, 1024), dx2 = make_desc(smem_u32(sx2), 16, 1024);
    const uint64_t dth = make_desc(smem_u32(sth), 16, 1024);
    const uint64_t tx1 = make_desc(smem_u32(sx1), BLK_BYTES, 1024), tx2 = make_desc(smem_u32(sx2), BLK_BYTES, 1024);
    const uint64_t tdl = make_desc(smem_u32(sdl), BLK_BYTES, 1024);
    const int flush_step = flush_tiles / NCH > 0 ? flush_tiles / NCH : 1;  // chunk ch is drained when phase == ch * flush_step
    int phase = 0;                                                          // it % (NCH * flush_step)
    for (int64_t it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      if (pf_mode == 2 && pf_dist > 0) {
        // asynchronous bulk prefetch (TMA engine, no LSU / scoreboard involvement) of a whole tile pf_dist/NBLK tiles ahead
        const int64_t itp = it + pf_dist / NBLK;
        if (itp < my_tiles && elect_one()) {
          const int64_t row0 = (blockIdx.x + itp * gridDim.x) * TILE_ROWS;
          const int64_t rows = (N - row0) < TILE_ROWS ? (N - row0) : TILE_ROWS;
          if (ldx == D) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(X + row0 * ldx), "r"((uint32_t)(rows * D * 4)) : "memory");
          } else {
            for (int64_t rr = 0; rr < rows; ++rr)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(X + (row0 + rr) * ldx), "r"((uint32_t)(D * 4)) : "memory");
          }
        }
        __syncwarp();
      }
      // ---- GEMM1: logits tile, block by block as the loaders publish them
      for (int b = 0; b < NBLK; ++b) {
        // (no tcgen05.fence here: the operands come from generic-proxy stores ordered by the mbarrier + proxy fence, and
        // tcgen05.fence::after_thread_sync would wait for the previous tile's GEMM2 to drain: ~900 cycles measured)
        mbar_wait(&bars->x_full[b], ph);
        if (lane == 0) BJX_TRACE(it, 16 + b);
        if (!(ablate & 1) && elect_one()) {
          const uint64_t boff = (uint64_t)((b * BLK_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle row
            umma_bf16(tmem + ACC1_COL, dx1 + boff + 2 * k, dth + boff + 2 * k, ID_G1_N64, (b | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + ACC1_COL, dx2 + boff + 2 * k, dth + boff + 2 * k, ID_G1_N32, 1);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bars->acc1_full);
      __syncwarp();
      if (lane == 0) BJX_TRACE(it, 24);
      // ---- GEMM2: G^T chunks; K = the 64 rows of the tile
      mbar_wait(&bars->dl_full, ph);
      tc_fence_after();
      if (lane == 0) BJX_TRACE(it, 25);
      for (int ch = 0; ch < NCH; ++ch) {
        if (elect_one()) {
          const uint64_t coff = (uint64_t)((2 * ch * BLK_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 16 rows = two 8-row swizzle atoms, 1024 bytes apart)
            if (ablate & 2) continue;
            umma_bf16(tmem + ACC2_COL + ch * NB, tx1 + coff + 128 * k, tdl + 128 * k, ID_G2, (k != 0) || !(it == 0 || phase == ch * flush_step));
            umma_bf16(tmem + ACC2_COL + ch * NB, tx2 + coff + 128 * k, tdl + 128 * k, ID_G2_N32, 1);
          }
          umma_commit(&bars->x_empty[2 * ch]);
          umma_commit(&bars->x_empty[2 * ch + 1]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bars->dl_empty);
      __syncwarp();
      if (lane == 0) BJX_TRACE(it, 26);
      if (++phase == NCH * flush_step) phase = 0;
    }
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else {
    // ===================================================== epilogue =====================================================
    const int q = warp & 3;                 // TMEM sub-partition this warp may access
    const int h = (warp - EPI_WARP0) >> 2;  // which half of the 32 samples
    // UMMA M = 64: row r lives in lane (r % 16) of sub-partition r / 16, so tcgen05.ld fills lanes 0-15 only.  The upper
    // half-warp takes over 8 of the 16 columns by shuffle, so all 32 lanes (and the whole MUFU pipe) do useful work.
    const int hi = lane >> 4;
    const int r = 16 * q + (lane & 15);
    const int s0 = h * 16 + hi * 8;         // this lane's 8 samples
    float stat[8], prod[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      stat[j] = 0.0f;
      prod[j] = 1.0f;
    }
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
    const bool logit = family == BJX_LIK_BERNOULLI_LOGITS;
    float* gp = Gpart + (size_t)blockIdx.x * S * D;
    // Move the GEMM2 accumulators (TMEM lanes = d, columns = [x*e1 | x1*e2]) into the per-CTA fp32 partial G[s, d], eight
    // samples at a time, with fire-and-forget reductions.  Only this thread ever touches a given address and reductions of
    // one thread to one address are applied in program order, so the sum is deterministic.
    auto drain_chunk = [&](int ch, bool first) {
      {
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
          const int sb = h * 16 + part * 8;
          uint32_t fa[8], fb[8];
          tmem_ld8(t_lane + ACC2_COL + ch * NB + sb, fa);
          tmem_ld8(t_lane + ACC2_COL + ch * NB + NS + sb, fb);
          float* base = gp + (size_t)sb * D + ch * 128 + 32 * q + lane;
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (sb + j < S) {
              const float val = __uint_as_float(fa[j]) + __uint_as_float(fb[j]);
              if (first)
                __stcg(base + (size_t)j * D, val);
              else
                atomicAdd(base + (size_t)j * D, val);  // result unused -> RED.ADD.F32: no round trip
            }
          }
        }
      }
    };
    const int flush_step = flush_tiles / NCH > 0 ? flush_tiles / NCH : 1;
    int phase = 0;
    uint32_t flushed = 0;  // bit ch: chunk ch has been stored to the partial at least once
    for (int64_t it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      if (it > 0 && phase % flush_step == 0 && phase / flush_step < NCH) {
        // GEMM2 of the previous tile has retired (dl_empty); this warp would otherwise idle until the next logits arrive
        const int ch = phase / flush_step;
        mbar_wait(&bars->dl_empty, ph ^ 1);
        tc_fence_after();
        drain_chunk(ch, !((flushed >> ch) & 1u));  // the MMA warp restarts this chunk's chain with this tile's GEMM2
        tc_fence_before();
        flushed |= 1u << ch;
      }
      if (++phase == NCH * flush_step) phase = 0;
      const int64_t row = (blockIdx.x + it * gridDim.x) * TILE_ROWS + r;
      const bool row_ok = row < N;
      const float yv = row_ok ? y[row] : 0.0f;
      mbar_wait(&bars->acc1_full, ph);
      tc_fence_after();
      if (tid == EPI_WARP0 * 32) BJX_TRACE(it, 32);
      uint32_t va[16], vb[16];
      tmem_ld16(t_lane + ACC1_COL + h * 16, va);
      tmem_ld16(t_lane + ACC1_COL + NS + h * 16, vb);
      tmem_ld_wait();
      if (tid == EPI_WARP0 * 32) BJX_TRACE(it, 33);
      float eta[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float lo = __uint_as_float(va[j]) + __uint_as_float(vb[j]);
        const float up = __shfl_sync(0xffffffffu, __uint_as_float(va[8 + j]) + __uint_as_float(vb[8 + j]), lane & 15);
        eta[j] = hi ? up : lo;
      }
      uint32_t e1[4], e2[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float g[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool ok = row_ok && (s0 + j + u) < S;
          const float x = eta[j + u];
          if (ablate & 4) {
            g[u] = ok ? x : 0.0f;
          } else if (logit) {
            // one exponential per datum: e = exp(-|x|); sigmoid = (x >= 0 ? 1 : e) / (1 + e);
            // softplus = max(x, 0) + log(1 + e), with the logs taken of a running product (flushed every 32 tiles)
            const float e = __expf(-fabsf(x));
            const float ope = 1.0f + e;
            const float rc = __fdividef(1.0f, ope);  // MUFU.RCP
            stat[j + u] += ok ? (yv * x - fmaxf(x, 0.0f)) : 0.0f;
            prod[j + u] *= ok ? ope : 1.0f;
            g[u] = ok ? (yv - (x >= 0.0f ? rc : e * rc)) : 0.0f;
          } else {
            const float rr = yv - x;
            stat[j + u] += ok ? rr * rr : 0.0f;
            g[u] = ok ? rr : 0.0f;
          }
        }
        split2(g[0], g[1], e1[j >> 1], e2[j >> 1]);
      }
      if ((it & 31) == 31) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          stat[j] -= logf(prod[j]);
          prod[j] = 1.0f;
        }
      }
      if (tid == EPI_WARP0 * 32) BJX_TRACE(it, 34);
      mbar_wait(&bars->dl_empty, ph ^ 1);  // GEMM2 of the previous tile has consumed the dl buffer
      {
        uint8_t* rowp = sdl + r * 128;
        const int sw = r & 7;
        *reinterpret_cast<uint4*>(rowp + (((2 * h + hi) ^ sw) << 4)) = make_uint4(e1[0], e1[1], e1[2], e1[3]);
        *reinterpret_cast<uint4*>(rowp + (((4 + 2 * h + hi) ^ sw) << 4)) = make_uint4(e2[0], e2[1], e2[2], e2[3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->dl_full);
      if (tid == EPI_WARP0 * 32) BJX_TRACE(it, 35);
    }
    // ---- per-CTA partial of stat[s]: reduce over the 64 rows (16 lanes x 4 sub-partitions)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = stat[j] - logf(prod[j]);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((lane & 15) == 0) sred[q * NS + s0 + j] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
    if (warp == EPI_WARP0 && lane < S) {
      statpart[(size_t)blockIdx.x * S + lane] = sred[lane] + sred[NS + lane] + sred[2 * NS + lane] + sred[3 * NS + lane];
    }
    // ---- per-CTA partial of G[s, d]: add what is left in the GEMM2 accumulators once all MMAs have retired
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    if (my_tiles > 0) {
      for (int ch = 0; ch < NCH; ++ch) drain_chunk(ch, !((flushed >> ch) & 1u));
    } else {
      for (int i = (warp - EPI_WARP0) * 32 + lane; i < S * D; i += EPI_WARPS * 32) gp[i] = 0.0f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

#undef BJX_TRACE

// =========================================================================================================================
// The same single pass with the dataset held in HBM as split-bf16 planes (bjx_prepare_x_planes: x = x1 + x2, 4 bytes per
// element like fp32): every 64 x 64 block of the tile is one TMA box per plane, written by the copy engine straight into
// the swizzled UMMA image.  No loader warps, no registers in flight, no conversion and no generic-proxy stores or fences
// on the shared-memory port: one producer thread, one MMA warp, eight epilogue warps (320 threads).
namespace tcp {
constexpr int PROD_WARP = 0, MMA_WARP = 1, EPI_WARP0 = 2;  // then EW (8 or 16) epilogue warps
constexpr int GATHER_WARPS = 4;  // IDX: producer warps of the gather (warp 0 and GATHER_WARPS - 1 more after the epilogue warps)
struct Maps {
  CUtensorMap x1, x2;
};
struct BarsP {  // nine block slots: index 8 is the spare slot of block 0
  unsigned long long x_full[9], x_empty[9];
  unsigned long long acc1_full[2], dl_full[2], dl_empty[2], done;  // one logits barrier per accumulator (tile parity); g hand-over per epilogue group
  unsigned long long px_full[2];  // PAIR: the peer CTA's partial logits of a tile (by tile parity) have landed (transaction bytes)
  uint32_t tmem_base;
};
}  // namespace tcp

// IDX: row r of the problem is row ridx[r] of the planes (a minibatch drawn on the device, BjxElboArgs.row_index + x_planes +
// N_full): the producer warp gathers the batch's rows straight out of the FULL dataset's planes with 16-byte cp.async copies
// into the swizzled image, so a batch costs one pass over its own rows and nothing else.
//
// PAIR: a cluster of two CTAs works on one 64-row tile, each on HALF of the columns (CTA rank r owns columns [r * 64 * NBLK,
// (r + 1) * 64 * NBLK)): every CTA runs GEMM1 over its own K range, the epilogue warps send their partial logits to the same
// thread of the peer CTA through distributed shared memory (st.async + the peer's transaction barrier) and add what they
// receive, both CTAs evaluate the link function on the (bit-identical) sums, and each runs GEMM2 for its own columns.  This
// doubles the width one pass can hold -- TMEM has room for G^T of 512 columns x 32 samples per SM -- to D <= 1024, and at
// D <= 512 it halves the bytes of a tile per SM, so that two tiles fit the slots (the ring mode of the narrow widths).
//
// NG = 2 (D <= 128 only: the ring holds four tiles): the sixteen epilogue warps form TWO groups that work on alternate tiles,
// each with its own logits accumulator, g buffer and hand-over barriers.  A tile of 64 rows x 128 columns is 32 KB -- 1360
// cycles of HBM time per SM -- while one tile's epilogue is a dependent chain of ~1250 cycles (tcgen05.ld, partner swap,
// MUFU, g store, fence, arrive): with one group the chain is exposed on every tile, with two it overlaps the other group's.
template <int NBLK, int EW, bool IDX, bool PAIR, int NG = 1>  // per-CTA D / 64; epilogue warps (8 or 16); IDX adds a second producer warp
__global__ void __launch_bounds__((tcp::EPI_WARP0 + EW + (IDX ? tcp::GATHER_WARPS - 1 : 0)) * 32, 1)
k_lik_tc_tma(const __grid_constant__ tcp::Maps maps, const __nv_bfloat16* __restrict__ planes, const float* __restrict__ y,
             int64_t N, const float* __restrict__ theta, int64_t Dtot, int64_t w_off, int Dw, int Dp, int S, int family,
             float* __restrict__ Gpart, float* __restrict__ statpart, int pf_tiles, int flush_tiles, int ablate,
             long long* __restrict__ trace, const int32_t* __restrict__ ridx, int64_t plane_rows) {
#define BJX_TRACE(it_, slot_)                                                                 \
  do {                                                                                        \
    if (trace != nullptr && blockIdx.x == 0 && (it_) >= 16 && (it_) < 24) trace[((it_)-16) * 64 + (slot_)] = clock64(); \
  } while (0)
  using namespace tc;
  constexpr int NTHREADS = (tcp::EPI_WARP0 + EW + (IDX ? tcp::GATHER_WARPS - 1 : 0)) * 32;
  constexpr int PROD_WARP2 = tcp::EPI_WARP0 + EW;  // IDX only: the further producer warps start here
  constexpr int GW = tcp::GATHER_WARPS;            // producer warps of the gather: (plane, group of rows) each
  constexpr int GPW = 16 / (GW / 2);               // four-row groups per producer warp (16 groups = 64 rows per plane)
  constexpr int EWG = EW / NG;  // epilogue warps that work on one tile
  constexpr int HG = EWG / 4;   // groups of samples among the epilogue warps of one TMEM sub-partition
  constexpr int SG = NS / HG;   // samples whose partial logits a thread loads (16 or 8)
  constexpr int VPT = SG / 2;   // (row, sample) pairs a thread finishes after the swap with its partner (8 or 4)
  static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
  static_assert(!(PAIR && IDX), "the gather variant runs one CTA per tile");
  constexpr int D = NBLK * BLK_COLS;  // columns of THIS CTA
  constexpr int NCH = NBLK / 2;
  // PAIR: rank in the cluster, tile sequence of the cluster (both CTAs walk the same tiles), first column of this CTA
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int cid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ncl = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int col0 = rank * D;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // Block 0 of a tile alternates between its regular slot and a spare one (index 8) placed one block BEFORE it, so the
  // producer can fetch block 0 of tile t+1 while tile t is still being consumed: GEMM1 of the next tile then starts the
  // moment GEMM2 retires instead of waiting a TMA latency.  For GEMM2 (MN-major A, two 64-column blocks per chunk) the
  // spare slot is reached by a leading-dimension byte offset of two blocks instead of one.
  // X slots are 16 KB: [x1 block | x2 block], i.e. 128 rows of 128 bytes -- GEMM1 takes the stacked pair as ONE UMMA A
  // operand with M = 128 (rows 0-63 the x1 terms, rows 64-127 the x2 terms of the same 64 data rows); slot -1 is the spare.
  // Narrow problems (D = 128, 256): a tile is only 2 or 4 blocks, so the eight slots form a RING of R = 4 or 2 tiles
  // (slot = (tile % R) * NBLK + block).  The producer then runs R tiles ahead of the tensor core, and the whole GEMM1 of
  // tile t+1 is issued under the epilogue of tile t -- with one tile of slots the loads of tile t+1 could only start when
  // GEMM2(t) released them and every tile paid a TMA round trip (0.43 of the HBM roofline at D = 128).
  constexpr int SLOT = 2 * BLK_BYTES;
  constexpr int R = 8 / NBLK;
  constexpr bool RING = R > 1;
  constexpr int NSLOT = RING ? 8 : NBLK;
  // the spare slot of block 0 exists outside ring mode, except for a pair of 512-column CTAs: there its 16 KB are the two
  // receive buffers of the partial-logit exchange (shared memory is full: 128 KB tile + 64 KB theta + 8 KB g + 8 KB swap)
  constexpr bool SPARE = !RING && !(PAIR && NBLK == 8);
  static_assert(NG == 1 || (NG == 2 && EW == 16 && R >= 4 && !IDX && !PAIR), "two epilogue groups need a ring of four tiles");
  uint8_t* sx = smem + SLOT;
  uint8_t* sth = sx + NSLOT * SLOT;
  uint8_t* sdl = sth + NBLK * BLK_BYTES;  // NG buffers of g
  tcp::BarsP* bars = (tcp::BarsP*)(sdl + NG * BLK_BYTES);
  float* sred = (float*)(bars + 1);       // [NG][4][NS]
  float* xch = (float*)(((uintptr_t)(sred + NG * 4 * NS) + 15) & ~(uintptr_t)15);  // [8 warps][8][32 lanes] partial-logit exchange
  constexpr int XCH_FLOATS = EWG * VPT * 32;                                   // (two of them in ring mode, see the epilogue)
  int32_t* sidx = reinterpret_cast<int32_t*>(xch + (RING ? 2 : 1) * XCH_FLOATS);  // IDX only: [4 tiles][64] ring of row indices
  // PAIR: [tile parity][VPT / 4][EW * 32 threads] float4 written by the peer CTA; in the unused spare slot where there is one
  float* pxch = SPARE ? xch + (RING ? 2 : 1) * XCH_FLOATS : reinterpret_cast<float*>(smem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (N + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t my_tiles = (n_tiles > cid) ? (n_tiles - cid + ncl - 1) / ncl : 0;

  if (tid == 0) {
    for (int b = 0; b < 9; ++b) {
      mbar_init(&bars->x_full[b], IDX ? 32 * GW : 1);  // TMA: the producer's expect_tx arrival, completion counted in bytes;
                                                  // gather: one cp.async.mbarrier.arrive per lane of the two producer warps
      mbar_init(&bars->x_empty[b], 1);
    }
    mbar_init(&bars->px_full[0], 1);
    mbar_init(&bars->px_full[1], 1);
    mbar_init(&bars->acc1_full[0], 1);
    mbar_init(&bars->acc1_full[1], 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars->dl_full[g], EWG);
      mbar_init(&bars->dl_empty[g], 1);
    }
    mbar_init(&bars->done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&maps.x1);
    tma_prefetch_desc(&maps.x2);
  }
  if (tid < NG * 4 * NS) sred[tid] = 0.0f;
  if (warp == tcp::MMA_WARP) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  for (int idx = tid; idx < NS * (D / 8); idx += NTHREADS) {
    const int s = idx / (D / 8), d8 = idx % (D / 8);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (s < S && col0 + d8 * 8 + j < Dw) ? theta[(int64_t)s * Dtot + w_off + col0 + d8 * 8 + j] : 0.0f;
    uint32_t p1[4], p2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], p1[j], p2[j]);
    const int b = d8 / 8, c = d8 % 8;
    uint8_t* blk = sth + b * BLK_BYTES;
    *reinterpret_cast<uint4*>(blk + s * 128 + ((c ^ (s & 7)) << 4)) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    const int n2 = 32 + s;
    *reinterpret_cast<uint4*>(blk + n2 * 128 + ((c ^ (n2 & 7)) << 4)) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before any store of mine can reach them
  const uint32_t tmem = bars->tmem_base;

  if (warp == tcp::PROD_WARP || (IDX && warp >= PROD_WARP2)) {
    // ===================================================== TMA producer ================================================
    if constexpr (IDX) {
      // Gather variant: TWO producer warps (one per operand plane: a single warp keeps only ~50 KB of copies in flight, 0.5
      // of the roofline with everything else switched off) copy the batch's plane rows with 16-byte cp.async (LDGSTS): no
      // registers in flight, any source address per lane, and the destination is the swizzled UMMA image itself.  Lane =
      // (r4, c): row 4g + r4 of the tile for g = 0..15, 16-byte chunk c of the block's 128-byte row; one instruction covers
      // four rows x 128 bytes, 16 instructions a block.  Each lane keeps the byte address of its sixteen source rows and
      // arrives on the block's barrier through cp.async.mbarrier.arrive (64 arrivals per phase); the MMA warp adds the
      // generic->async proxy fence after its wait.  Rows past N read row 0 of the dataset: the
      // epilogue zeroes their g, so they contribute nothing.  Blocks beyond the planes' extent are zero-filled (src size 0).
      // (tile::gather4 TMA loads were measured first: the copy engine retires one 512-byte gather4 per ~75 cycles, 0.27 of the
      // HBM roofline -- profiles/r02_minibatch_gather4_rejected.json.)
      const int r4 = lane >> 3, c = lane & 7;
      const int pwi = warp == tcp::PROD_WARP ? 0 : warp - PROD_WARP2 + 1;  // producer warp index 0 .. GW - 1
      const int pw = pwi & 1;                                               // which plane this warp copies
      const int g0 = (pwi >> 1) * GPW;                                      // its first four-row group
      // (two warps, one per plane, keep ~100 KB of copies in flight; four were measured next -- profiles/r02_minibatch.json)
      const int64_t pitch = (int64_t)Dp * 2;
      const char* pl = reinterpret_cast<const char*>(planes) + (int64_t)pw * plane_rows * pitch;
      const int dst_plane = pw * BLK_BYTES;
      // Row indices travel through a four-tile ring in shared memory: the warp loads a tile's 64 indices coalesced (two per
      // lane) THREE tiles ahead into registers, parks them in the ring one iteration later, and everything that needs an
      // address -- the L2 touch two tiles ahead, the source rows one tile ahead -- reads the ring, never global memory, so
      // no dependent global load sits in front of a copy.
      int32_t stage0 = 0, stage1 = 0;
      auto fetch_idx = [&](int64_t t) {  // -> registers (producer warp 0)
        if (pwi != 0) return;
        const int64_t r0 = (blockIdx.x + t * gridDim.x) * TILE_ROWS + 2 * lane;
        stage0 = (t < my_tiles && r0 < N) ? __ldg(ridx + r0) : 0;
        stage1 = (t < my_tiles && r0 + 1 < N) ? __ldg(ridx + r0 + 1) : 0;
      };
      auto park_idx = [&](int64_t t) {  // registers -> ring slot t % 4, visible to both producer warps
        if (pwi == 0) {
          int32_t* slot = sidx + (int)(t & 3) * TILE_ROWS;
          slot[2 * lane] = stage0;
          slot[2 * lane + 1] = stage1;
        }
        asm volatile("bar.sync 14, %0;" ::"r"(32 * GW) : "memory");
      };
      auto load_rows = [&](int64_t t, const char*(&rb)[GPW]) {
        const int32_t* slot = sidx + (int)(t & 3) * TILE_ROWS;
#pragma unroll
        for (int g = 0; g < GPW; ++g) rb[g] = pl + (int64_t)slot[4 * (g0 + g) + r4] * pitch + c * 16;
      };
      // A row's eight 128-byte pieces (one per block) are copied at different times, and the X slots have no slack, so the
      // copies of tile t+1 can only be issued as GEMM2(t) releases its blocks: without help every tile pays a full random-
      // access HBM latency between GEMM2 and the next GEMM1.  Touching ALL lines of a tile's rows two tiles ahead (per-lane
      // prefetch.global.L2: lane c takes line c of row 4g + r4, both planes) turns that into an L2 hit.
      const int pf_kind = pf_tiles & 255;  // 0 off | 1 one touch per 128-byte line | 2 one per 32-byte sector | 3 bulk prefetch per row
      auto prefetch_rows = [&](int64_t t) {
        if (pf_kind == 0 || t >= my_tiles) return;
        const int32_t* slot = sidx + (int)(t & 3) * TILE_ROWS;
        const int64_t r0 = (blockIdx.x + t * gridDim.x) * TILE_ROWS;
        if (pf_kind == 3) {  // two rows per lane, the whole plane row (pitch bytes) in one request
          if (pwi >= 2) return;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int r = 2 * lane + u;
            if (r0 + r < N)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pl + (int64_t)slot[r] * pitch), "r"((uint32_t)pitch) : "memory");
          }
          return;
        }
        if (c * 128 >= pitch) return;
#pragma unroll
        for (int gg = 0; gg < GPW; ++gg) {
          const int g = g0 + gg;
          if (r0 + r4 + 4 * g < N) {
            const char* p = pl + (int64_t)slot[4 * g + r4] * pitch + c * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            if (pf_kind == 2) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 64));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 96));
            }
          }
        }
      };
      auto issue = [&](uint8_t* slot, unsigned long long* bar, int b, const char* const(&rb)[GPW]) {
        const uint32_t nbytes = (b * BLK_COLS < Dp) ? 16u : 0u;
        const int64_t boff = nbytes ? (int64_t)b * 128 : 0;
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
          const int r = 4 * (g0 + g) + r4;
          cp_async16(slot + dst_plane + r * 128 + ((c ^ (r & 7)) << 4), rb[g] + boff, nbytes);
        }
        cp_async_mbar_arrive(bar);
      };
      for (int64_t t = 0; t < 3; ++t) {  // prologue: tiles 0, 1, 2 into the ring (the only exposed index latency)
        fetch_idx(t);
        park_idx(t);
      }
      fetch_idx(3);
      const char *cur[GPW], *nxt[GPW];
      load_rows(0, cur);
      prefetch_rows(0);
      prefetch_rows(1);
      if (!RING && my_tiles > 0) {
        mbar_wait(&bars->x_empty[0], 1u);
        issue(sx, &bars->x_full[0], 0, cur);
      }
      for (int64_t it = 0; it < my_tiles; ++it) {
        load_rows(it + 1, nxt);
        if constexpr (RING) {
          const int q0 = (int)(it % R) * NBLK;
          const uint32_t use = (uint32_t)((it / R) & 1);
          for (int b = 0; b < NBLK; ++b) {
            const int q = q0 + b;
            mbar_wait(&bars->x_empty[q], use ^ 1);
            issue(sx + q * SLOT, &bars->x_full[q], b, cur);
          }
        } else {
          for (int b = 1; b < NBLK; ++b) {
            mbar_wait(&bars->x_empty[b], (uint32_t)((it & 1) ^ 1));
            issue(sx + b * SLOT, &bars->x_full[b], b, cur);
            if (b == 1 && it + 1 < my_tiles) {  // block 0 of the next tile into the other block-0 slot (see below)
              const int64_t t = it + 1;
              const int alt = (int)(t & 1);
              const int idx = alt ? 8 : 0;
              mbar_wait(&bars->x_empty[idx], (uint32_t)(((t >> 1) & 1) ^ 1));
              issue(alt ? sx - SLOT : sx, &bars->x_full[idx], 0, nxt);
            }
          }
        }
        prefetch_rows(it + 2);  // AFTER this tile's copies: touches issued first would queue ahead of the copies GEMM1 is waiting for
#pragma unroll
        for (int g = 0; g < GPW; ++g) cur[g] = nxt[g];
        park_idx(it + 3);   // fetched during the previous iteration; its ring slot held tile it - 1
        fetch_idx(it + 4);
      }
    } else {
      const __nv_bfloat16* p1 = planes;
      const __nv_bfloat16* p2 = planes + (size_t)N * Dp;  // Dp = columns of a plane (Dw padded to 64); blocks beyond it are TMA zero fill
      const int pf_mode = pf_tiles >> 8;  // development switch: 0 bulk prefetch before the tile's loads, 1 after them
      const int pf_dist = pf_tiles & 255;
      if (lane == 0) {
        auto issue_block0 = [&](int64_t t) {
          const int alt = (int)(t & 1);
          const int idx = alt ? 8 : 0;
          const int64_t row0 = ((int64_t)cid + t * ncl) * TILE_ROWS;
          mbar_wait(&bars->x_empty[idx], (uint32_t)(((t >> 1) & 1) ^ 1));  // released by chunk 0 of GEMM2 two tiles ago
          BJX_TRACE(t, 0);
          mbar_expect_tx(&bars->x_full[idx], 2u * BLK_BYTES);
          uint8_t* slot = alt ? sx - SLOT : sx;
          tma_load_2d(slot, &maps.x1, col0, (int32_t)row0, &bars->x_full[idx]);
          tma_load_2d(slot + BLK_BYTES, &maps.x2, col0, (int32_t)row0, &bars->x_full[idx]);
        };
        const int pf_ahead = pf_dist * R;  // the same distance in bytes for every D
        if (SPARE && my_tiles > 0) issue_block0(0);
        for (int64_t it = 0; it < my_tiles; ++it) {
          const int64_t row0 = ((int64_t)cid + it * ncl) * TILE_ROWS;
          auto bulk_pf = [&]() {
            // pull a later tile of both planes into L2 (rows of a tile are contiguous: 64 KB per plane at D = 512); in a
            // pair each CTA takes one plane of the whole tile
            const int64_t rowp = ((int64_t)cid + (it + pf_ahead) * ncl) * TILE_ROWS;
            const int64_t rows = (N - rowp) < TILE_ROWS ? (N - rowp) : TILE_ROWS;
            const uint32_t bytes = (uint32_t)(rows * Dp * 2);
            if (!PAIR || rank == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p1 + rowp * Dp), "r"(bytes) : "memory");
            if (!PAIR || rank == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p2 + rowp * Dp), "r"(bytes) : "memory");
          };
          if (pf_mode == 0 && pf_dist > 0 && it + pf_ahead < my_tiles) bulk_pf();
          if constexpr (RING) {
            const int q0 = (int)(it % R) * NBLK;
            const uint32_t use = (uint32_t)((it / R) & 1);
            for (int b = 0; b < NBLK; ++b) {
              const int q = q0 + b;
              mbar_wait(&bars->x_empty[q], use ^ 1);  // GEMM2 of tile it - R released this slot
              BJX_TRACE(it, b);
              mbar_expect_tx(&bars->x_full[q], 2u * BLK_BYTES);
              tma_load_2d(sx + q * SLOT, &maps.x1, col0 + b * BLK_COLS, (int32_t)row0, &bars->x_full[q]);
              tma_load_2d(sx + q * SLOT + BLK_BYTES, &maps.x2, col0 + b * BLK_COLS, (int32_t)row0, &bars->x_full[q]);
            }
          } else {
            for (int b = SPARE ? 1 : 0; b < NBLK; ++b) {
              mbar_wait(&bars->x_empty[b], (uint32_t)((it & 1) ^ 1));  // GEMM2 of the previous tile released this block
              BJX_TRACE(it, b);
              mbar_expect_tx(&bars->x_full[b], 2u * BLK_BYTES);
              tma_load_2d(sx + b * SLOT, &maps.x1, col0 + b * BLK_COLS, (int32_t)row0, &bars->x_full[b]);
              tma_load_2d(sx + b * SLOT + BLK_BYTES, &maps.x2, col0 + b * BLK_COLS, (int32_t)row0, &bars->x_full[b]);
              // block 0 of the NEXT tile goes into the other slot, free since chunk 0 of GEMM2(it - 1) like block 1: issue
              // it now, so that it has landed when the MMA warp wants to run it under this tile's epilogue
              if (SPARE && b == 1 && it + 1 < my_tiles) issue_block0(it + 1);
            }
          }
          if (pf_mode == 1 && pf_dist > 0 && it + pf_ahead < my_tiles) bulk_pf();
        }
      }
    }
  } else if (warp == tcp::MMA_WARP) {
    // ===================================================== MMA issuer ==================================================
    constexpr uint32_t ID_G1 = make_idesc(128, NB, 0, 0);      // [x1; x2] * [t1; t2]: all four partial products in one UMMA
    constexpr uint32_t ID_G2 = make_idesc(128, NB, 1, 1);
    constexpr uint32_t ID_G2_N32 = make_idesc(128, NS, 1, 1);
    const uint64_t dxa = make_desc(smem_u32(sx), 16, 1024);
    const uint64_t dth = make_desc(smem_u32(sth), 16, 1024);
    // GEMM2 (MN-major A): the two 64-column blocks of a chunk are one slot (16 KB) apart
    const uint64_t tx1 = make_desc(smem_u32(sx), SLOT, 1024), tx2 = make_desc(smem_u32(sx + BLK_BYTES), SLOT, 1024);
    const uint64_t tdl = make_desc(smem_u32(sdl), BLK_BYTES, 1024);
    // chunk 0 of GEMM2 when block 0 sits in the spare slot: slots {-1, 1} are two slot strides apart
    const uint64_t tx1_alt = make_desc(smem_u32(sx - SLOT), 2 * SLOT, 1024);
    const uint64_t tx2_alt = make_desc(smem_u32(sx - SLOT + BLK_BYTES), 2 * SLOT, 1024);
    const int flush_step = flush_tiles / NCH > 0 ? flush_tiles / NCH : 1;
    int phase = 0;
    // GEMM1 of one block into the logits accumulator of tile t's parity (two accumulators: block 0 of the next tile is
    // issued right after this tile's logits are committed, so it runs while the epilogue warps work on them)
    auto gemm1_block = [&](int64_t t, int b) {
      const bool alt = SPARE && (t & 1) != 0;  // block 0 sits in the spare slot
      const bool acc_b = (t & 1) != 0;         // which logits accumulator
      if constexpr (RING) {
        const int q = (int)(t % R) * NBLK + b;
        mbar_wait(&bars->x_full[q], (uint32_t)((t / R) & 1));
        if constexpr (IDX) { if (!(ablate & 16)) fence_proxy_async(); }  // the block was written by cp.async (generic proxy); UMMA reads through the async proxy
        if (lane == 0) BJX_TRACE(t, 16 + b);
        if (!(ablate & 1) && elect_one()) {
          const uint64_t toff = (uint64_t)((b * BLK_BYTES) >> 4);
          const uint64_t boff = (uint64_t)((q * SLOT) >> 4);
          const uint32_t acc = tmem + ACC1_COL + (acc_b ? NB : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc, dxa + boff + 2 * k, dth + toff + 2 * k, ID_G1, (b | k) != 0);
        }
        __syncwarp();
        return;
      }
      if (b == 0 && SPARE)
        mbar_wait(&bars->x_full[alt ? 8 : 0], (uint32_t)((t >> 1) & 1));
      else
        mbar_wait(&bars->x_full[b], (uint32_t)(t & 1));
      if constexpr (IDX) { if (!(ablate & 16)) fence_proxy_async(); }
      if (lane == 0) BJX_TRACE(t, 16 + b);
      if (!(ablate & 1) && elect_one()) {
        const uint64_t toff = (uint64_t)((b * BLK_BYTES) >> 4);  // theta block
        const uint64_t boff = (b == 0 && alt) ? (uint64_t)(-(int64_t)(SLOT >> 4)) : (uint64_t)((b * SLOT) >> 4);  // X slot
        const uint32_t acc = tmem + ACC1_COL + (acc_b ? NB : 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(acc, dxa + boff + 2 * k, dth + toff + 2 * k, ID_G1, (b | k) != 0);
      }
      __syncwarp();
    };
    if (my_tiles > 0) {
      if constexpr (RING) {
        for (int64_t t = 0; t < NG && t < my_tiles; ++t) {  // one tile per epilogue group
          for (int b = 0; b < NBLK; ++b) gemm1_block(t, b);
          if (elect_one()) umma_commit(&bars->acc1_full[t & 1]);
          __syncwarp();
        }
      } else if (SPARE) {
        gemm1_block(0, 0);
      }
    }
    for (int64_t it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const bool alt = SPARE ? (it & 1) != 0 : false;
      const int idx0 = alt ? 8 : 0;
      const int q0 = RING ? (int)(it % R) * NBLK : 0;  // first slot of this tile
      const int eg = NG == 2 ? (int)(it & 1) : 0;       // the epilogue group of this tile
      if constexpr (NG == 2) {
        // two groups: GEMM1 of tiles it and it + 1 are already issued (one logits accumulator each); GEMM1(it + 2) follows
        // this tile's GEMM2 below -- the same event (dl_full of this group) frees its accumulator
      } else if constexpr (RING) {
        // the whole GEMM1 of the next tile (other logits accumulator, slots already in the ring) under this tile's epilogue
        // Its barrier is the OTHER one (one per accumulator): the previous phase of that barrier was tile it - 1, which
        // every epilogue warp consumed before dl_full(it - 1) -- so it may complete right away, and the epilogue goes from
        // tile `it` straight into tile it + 1 without a hand-over through this warp.
        if (lane == 0) BJX_TRACE(it, 24);
        if (it + 1 < my_tiles) {
          for (int b = 0; b < NBLK; ++b) gemm1_block(it + 1, b);
          if (elect_one()) umma_commit(&bars->acc1_full[(it + 1) & 1]);
          __syncwarp();
        }
      } else {
        for (int b = SPARE ? 1 : 0; b < NBLK; ++b) gemm1_block(it, b);
        if (elect_one()) umma_commit(&bars->acc1_full[it & 1]);
        __syncwarp();
        if (lane == 0) BJX_TRACE(it, 24);
        if (SPARE && it + 1 < my_tiles) gemm1_block(it + 1, 0);  // its slot (the other one) was filled a tile ago
      }
      mbar_wait(&bars->dl_full[eg], NG == 2 ? (uint32_t)((it >> 1) & 1) : ph);
      tc_fence_after();
      if (lane == 0) BJX_TRACE(it, 25);
      const uint64_t tdl_g = tdl + (uint64_t)((eg * BLK_BYTES) >> 4);
      for (int ch = 0; ch < NCH; ++ch) {
        if (elect_one()) {
          const uint64_t coff = (uint64_t)(((q0 + 2 * ch) * SLOT) >> 4);
          const uint64_t a1 = (ch == 0 && alt) ? tx1_alt : tx1 + coff;
          const uint64_t a2 = (ch == 0 && alt) ? tx2_alt : tx2 + coff;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ablate & 2) continue;
            umma_bf16(tmem + ACC2_COL + ch * NB, a1 + 128 * k, tdl_g + 128 * k, ID_G2,
                      (k != 0) || !(it == 0 || phase == ch * flush_step));
            umma_bf16(tmem + ACC2_COL + ch * NB, a2 + 128 * k, tdl_g + 128 * k, ID_G2_N32, 1);
          }
          umma_commit(&bars->x_empty[(ch == 0 && !RING) ? idx0 : q0 + 2 * ch]);
          umma_commit(&bars->x_empty[q0 + 2 * ch + 1]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bars->dl_empty[eg]);
      __syncwarp();
      if (lane == 0) BJX_TRACE(it, 26);
      if constexpr (NG == 2) {
        if (it + 2 < my_tiles) {
          for (int b = 0; b < NBLK; ++b) gemm1_block(it + 2, b);
          if (elect_one()) umma_commit(&bars->acc1_full[eg]);
          __syncwarp();
          if (lane == 0) BJX_TRACE(it + 2, 24);
        }
      }
      if (++phase == NCH * flush_step) phase = 0;
    }
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else {
    // ===================================================== epilogue =====================================================
    // GEMM1 is an M = 128 UMMA: TMEM lane 64 * part + r holds, for data row r, the partial logits of the x1 terms
    // (part = 0) or of the x2 terms (part = 1), columns [x*t1 (32) | x*t2 (32)].  The warp of sub-partition q (lanes
    // 32q..32q+31) and its partner in sub-partition q ^ 2 hold the two parts of the same 32 rows; they swap halves of
    // their 16 samples through shared memory, so that every thread finishes 8 (row, sample) pairs.
    const int q = warp & 3;
    const int eg = (warp - tcp::EPI_WARP0) / EWG;      // epilogue group (NG = 2: this group takes the tiles of parity eg)
    const int wg = (warp - tcp::EPI_WARP0) - eg * EWG;  // warp within the group
    const int h = wg >> 2;
    const int part = q >> 1;
    const int r = 32 * (q & 1) + lane;
    const int s0 = h * SG + part * VPT;
    float* xch_mine = xch + (size_t)wg * (VPT * 32);             // what the partner wrote for me
    float* xch_peer = xch + (size_t)(wg ^ 2) * (VPT * 32);       // (warp index ^ 2 flips q's high bit)
    const int pair_bar = 2 + 2 * HG * eg + HG * (q & 1) + h;     // named barriers 2.., 64 threads each
    uint8_t* sdl_g = sdl + eg * BLK_BYTES;
    float stat[VPT], prod[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      stat[j] = 0.0f;
      prod[j] = 1.0f;
    }
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
    const bool logit = family == BJX_LIK_BERNOULLI_LOGITS;
    float* gp = Gpart + (size_t)cid * S * Dw;  // the partial keeps the caller's [S, Dw] layout; a pair shares one, by columns
    // The GEMM2 accumulators are folded every few tiles (the tensor core's fp32 accumulate truncates) into an fp32 partial
    // that lives in the 128 TMEM columns left over ([384, 512): 4 chunks x 32 samples, lane = weight column): registers
    // add, tcgen05.st writes back -- no global-memory traffic until the single store at the end of the kernel.
    constexpr int PART_COL = 384;
    auto drain_chunk = [&](int ch, bool first, bool final) {
#pragma unroll 1
      for (int part = 0; part < SG / 8; ++part) {
        const int sb = h * SG + part * 8;
        uint32_t fa[8], fb[8], fp[8];
        tmem_ld8(t_lane + ACC2_COL + ch * NB + sb, fa);
        tmem_ld8(t_lane + ACC2_COL + ch * NB + NS + sb, fb);
        if (!first) tmem_ld8(t_lane + PART_COL + ch * NS + sb, fp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float val = __uint_as_float(fa[j]) + __uint_as_float(fb[j]);
          if (!first) val += __uint_as_float(fp[j]);
          fp[j] = __float_as_uint(val);
        }
        if (final) {
          const int d = col0 + ch * 128 + 32 * q + lane;
          float* base = gp + (size_t)sb * Dw + d;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (sb + j < S && d < Dw) __stcg(base + (size_t)j * Dw, __uint_as_float(fp[j]));
        } else {
          tmem_st8(t_lane + PART_COL + ch * NS + sb, fp);
        }
      }
      if (!final) tmem_st_wait();
    };
    const int flush_step = flush_tiles / NCH > 0 ? flush_tiles / NCH : 1;
    const int flush_period = NCH * flush_step;
    // chunk ch is folded for the first time at tile ch * flush_step (chunk 0: one period in), then once per period
    auto first_fold = [&](int ch) { return (int64_t)(ch == 0 ? flush_period : ch * flush_step); };
    int64_t row = ((int64_t)cid + (int64_t)eg * ncl) * TILE_ROWS + r;
    float yv = (my_tiles > eg && row < N) ? y[row] : 0.0f;
    uint32_t va[SG], vb[SG];  // partial logits of one tile
    // PAIR: my slot in the peer's receive buffers and the peer's transaction barriers (shared::cluster addresses)
    const int te = tid - tcp::EPI_WARP0 * 32;
    constexpr int PX_FLOATS = EW * 32 * VPT;
    uint32_t px_remote = 0, px_rbar = 0;
    if constexpr (PAIR) {
      px_remote = mapa_cluster(smem_u32(pxch) + (uint32_t)te * 16u, (uint32_t)(rank ^ 1));
      px_rbar = mapa_cluster(smem_u32(&bars->px_full[0]), (uint32_t)(rank ^ 1));
    }
    auto load_logits = [&](int64_t t) {
      mbar_wait(&bars->acc1_full[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const uint32_t acc1 = t_lane + ACC1_COL + ((t & 1) ? NB : 0);  // the logits accumulator of the tile's parity
      if constexpr (SG == 16) {
        tmem_ld16(acc1 + h * SG, va);
        tmem_ld16(acc1 + NS + h * SG, vb);
      } else {
        tmem_ld8(acc1 + h * SG, va);
        tmem_ld8(acc1 + NS + h * SG, vb);
      }
    };
    for (int64_t it = eg; it < my_tiles; it += NG) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int phase = (int)(it % flush_period);
      if (it > 0 && phase % flush_step == 0 && phase / flush_step < NCH) {
        // GEMM2 of the previous tile (the other group's, with two groups) has retired and this tile's cannot start before
        // this group hands over g: nobody touches the accumulators now
        const int ch = phase / flush_step;
        if constexpr (NG == 2)
          mbar_wait(&bars->dl_empty[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
        else
          mbar_wait(&bars->dl_empty[0], ph ^ 1);
        tc_fence_after();
        drain_chunk(ch, it == first_fold(ch), false);
        tc_fence_before();
      }
      const bool row_ok = row < N;
      const float ycur = yv;
      // fetch the next tile's label now: its latency hides behind this tile's GEMM1
      row += (int64_t)NG * ncl * TILE_ROWS;
      yv = (it + NG < my_tiles && row < N) ? y[row] : 0.0f;
      if constexpr (PAIR) {  // this tile's receive phase: EW * 32 threads x VPT floats from the peer
        if (te == 0) mbar_expect_tx(&bars->px_full[it & 1], (uint32_t)(PX_FLOATS * sizeof(float)));
      }
      // (Loading the next tile's logits early, under this tile's link function, was tried in ring mode and lost 10 %: the
      // wait for GEMM1(t+1) -- queued behind GEMM2(t-1) on the tensor pipe -- then sits in front of this tile's hand-over.)
      load_logits(it);
      if (tid == (tcp::EPI_WARP0 + eg * EWG) * 32) BJX_TRACE(it, 32);
      tmem_ld_wait();
      // ring mode: the exchange buffers alternate with the tile (nothing orders my write for tile t+1 after the partner's
      // read for tile t any more, now that the next logits do not wait for this tile's hand-over)
      const int xoff = RING ? (int)(it & 1) * XCH_FLOATS : 0;
      float eta[VPT], g[VPT];
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        // samples [VPT * (1 - part), +VPT) of my SG go to the partner, [VPT * part, +VPT) stay
        const float lo = __uint_as_float(va[j]) + __uint_as_float(vb[j]);
        const float up = __uint_as_float(va[VPT + j]) + __uint_as_float(vb[VPT + j]);
        xch_peer[xoff + j * 32 + lane] = part ? lo : up;
        eta[j] = part ? up : lo;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
#pragma unroll
      for (int j = 0; j < VPT; ++j) eta[j] += xch_mine[xoff + j * 32 + lane];
      if constexpr (PAIR) {
        // eta holds the logits over THIS CTA's columns: send them to the same thread of the peer, add what it sent.  Two
        // receive buffers by tile parity: the peer writes tile t + 2 only after it has received my tile t + 1, which every
        // warp of mine sends after reading buffer t.  a + b == b + a exactly, so both CTAs get the same bits.
        const uint32_t po = (uint32_t)(it & 1) * (uint32_t)(PX_FLOATS * sizeof(float));
        const uint32_t rb = px_rbar + (uint32_t)(it & 1) * 8u;
#pragma unroll
        for (int j4 = 0; j4 < VPT / 4; ++j4)
          st_async_v4(px_remote + po + (uint32_t)j4 * (EW * 32 * 16), eta[4 * j4], eta[4 * j4 + 1], eta[4 * j4 + 2], eta[4 * j4 + 3], rb);
        mbar_wait_cluster(&bars->px_full[it & 1], (uint32_t)((it >> 1) & 1));
        const float4* mine = reinterpret_cast<const float4*>(pxch + (it & 1) * PX_FLOATS) + te;
#pragma unroll
        for (int j4 = 0; j4 < VPT / 4; ++j4) {
          const float4 o = mine[j4 * (EW * 32)];
          eta[4 * j4] += o.x;
          eta[4 * j4 + 1] += o.y;
          eta[4 * j4 + 2] += o.z;
          eta[4 * j4 + 3] += o.w;
        }
      }
      if (tid == (tcp::EPI_WARP0 + eg * EWG) * 32) BJX_TRACE(it, 33);
      // This stage is bound by instruction issue (2048 values per tile over four schedulers), so the common case -- a full
      // tile and S = 32 -- runs without the validity selects, and the reciprocal goes to the special-function pipe.
      const bool all_ok = (S == NS) && (row - (int64_t)NG * ncl * TILE_ROWS - r + TILE_ROWS <= N);  // warp-uniform
      // Only g is on the critical path (GEMM2 waits for it): the loss bookkeeping (stat, prod) is done after the hand-over.
      float ope[VPT];
      if (ablate & 4) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) g[j] = (row_ok && (s0 + j) < S) ? eta[j] : 0.0f;
      } else if (logit) {
        // one exponential per datum: e = exp(-|x|); sigmoid = (x >= 0 ? 1 : e) / (1 + e); softplus = max(x, 0) + log(1 + e),
        // the logs taken of a running product (every 32 tiles)
        float e[VPT], rc[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) e[j] = __expf(-fabsf(eta[j]));
#pragma unroll
        for (int j = 0; j < VPT; ++j) ope[j] = 1.0f + e[j];
#pragma unroll
        for (int j = 0; j < VPT; ++j) rc[j] = __fdividef(1.0f, ope[j]);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const float gj = ycur - (eta[j] >= 0.0f ? rc[j] : e[j] * rc[j]);
          g[j] = (all_ok || (row_ok && (s0 + j) < S)) ? gj : 0.0f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const float rr = ycur - eta[j];
          g[j] = (all_ok || (row_ok && (s0 + j) < S)) ? rr : 0.0f;
        }
      }
      uint32_t e1[VPT / 2], e2[VPT / 2];
#pragma unroll
      for (int j = 0; j < VPT / 2; ++j) split2(g[2 * j], g[2 * j + 1], e1[j], e2[j]);
      if (tid == (tcp::EPI_WARP0 + eg * EWG) * 32) BJX_TRACE(it, 34);
      // GEMM2 of this group's previous tile has consumed its g buffer
      mbar_wait(&bars->dl_empty[eg], NG == 2 ? (uint32_t)(((it >> 1) & 1) ^ 1) : (ph ^ 1));
      {
        uint8_t* rowp = sdl_g + r * 128;
        const int sw = r & 7;
        // row layout: [e1: 32 samples x 2 B | e2: 32 x 2 B], 16-byte chunks XOR-swizzled by the row
        if constexpr (VPT == 8) {
          *reinterpret_cast<uint4*>(rowp + (((2 * h + part) ^ sw) << 4)) = make_uint4(e1[0], e1[1], e1[2], e1[3]);
          *reinterpret_cast<uint4*>(rowp + (((4 + 2 * h + part) ^ sw) << 4)) = make_uint4(e2[0], e2[1], e2[2], e2[3]);
        } else {
          *reinterpret_cast<uint2*>(rowp + ((h ^ sw) << 4) + 8 * part) = make_uint2(e1[0], e1[1]);
          *reinterpret_cast<uint2*>(rowp + (((4 + h) ^ sw) << 4) + 8 * part) = make_uint2(e2[0], e2[1]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->dl_full[eg]);
      // ---- off the critical path: log-likelihood bookkeeping
      if (!(ablate & 4)) {
        if (logit) {
#pragma unroll
          for (int j = 0; j < VPT; ++j) {
            const bool ok = all_ok || (row_ok && (s0 + j) < S);
            stat[j] += ok ? (ycur * eta[j] - fmaxf(eta[j], 0.0f)) : 0.0f;
            prod[j] *= ok ? ope[j] : 1.0f;
          }
          if (((it / NG) & 31) == 31) {
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
              stat[j] -= logf(prod[j]);
              prod[j] = 1.0f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < VPT; ++j) stat[j] = fmaf(g[j], g[j], stat[j]);  // g = y - eta, already masked
        }
      }
      if (tid == (tcp::EPI_WARP0 + eg * EWG) * 32) BJX_TRACE(it, 35);
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float v = stat[j] - logf(prod[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sred[eg * 4 * NS + q * NS + s0 + j] = v;  // the other (q, sample) slots stay zero
    }
    asm volatile("bar.sync 1, %0;" ::"r"(EW * 32) : "memory");
    if (warp == tcp::EPI_WARP0 && lane < S && rank == 0) {  // both CTAs of a pair hold the same sums: one reports them
      float v = sred[lane] + sred[NS + lane] + sred[2 * NS + lane] + sred[3 * NS + lane];
      if constexpr (NG == 2) v += sred[4 * NS + lane] + sred[5 * NS + lane] + sred[6 * NS + lane] + sred[7 * NS + lane];
      statpart[(size_t)cid * S + lane] = v;
    }
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    if (my_tiles > 0) {
      if (eg == 0)
        for (int ch = 0; ch < NCH; ++ch) drain_chunk(ch, my_tiles <= first_fold(ch), true);
    } else {
      for (int i = (warp - tcp::EPI_WARP0) * 32 + lane; i < S * Dw; i += EW * 32) gp[i] = 0.0f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == tcp::MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
  if constexpr (PAIR) cluster_sync_all();  // no CTA leaves while the peer may still store into its shared memory
#undef BJX_TRACE
}

// =========================================================================================================================
// Narrow problems (D <= 128) on TALL tiles: 128 data rows per tile instead of 64.
//
// At D = 128 a 64-row tile is only 32 KB -- 1360 cycles of HBM time per SM -- while the fixed costs of a tile (barrier
// hand-overs, tcgen05.ld, the partner swap, the proxy fence in front of GEMM2) add up to a dependent chain of ~2000 cycles
// (profiles/r02_envelope_narrow_groups.json), so the 64-row kernel stops at 0.70-0.77 of the roofline there however the
// epilogue warps are arranged.  With 128 rows per tile the same fixed costs buy twice the bytes, and the accumulator layout
// gets simpler: GEMM1 is  x1[128 rows] * [t1 | t2]  (M = 128, N = 64)  +  x2[128 rows] * t1  (M = 128, N = 32) into the SAME
// TMEM columns, so TMEM lane r holds data row r completely (eta = acc[s] + acc[32 + s]) and the partner swap through
// shared memory is gone, together with the wasted x2 * t2 quarter of the stacked M = 128 form.
//   shared memory: 6 slots of 32 KB ([x1 block | x2 block], 128 rows x 64 columns each) = a ring of three tiles,
//                  theta 16 KB, g 16 KB ([128 rows][e1 (32 samples) | e2 (32)], MN-major B of GEMM2 with K = 128 rows)
//   TMEM: two logits accumulators (2 x 64 columns), G^T of the one 128-column chunk (64), its folded partial (32)
//   warps: producer, MMA issuer, sixteen epilogue warps (sub-partition q = warp % 4, sample group h: 8 samples per thread)
namespace tct {
constexpr int ROWS = 128;
constexpr int XBLK = ROWS * 128;  // one plane block: 16 KB
constexpr int SLOT = 2 * XBLK;    // 32 KB
constexpr int NSLOT = 6;  // a ring of three tiles of two blocks (64 < D <= 128) or of six tiles of one block (D <= 64)
constexpr int PROD_WARP = 0, MMA_WARP = 1, EPI_WARP0 = 2, EW = 16;
constexpr int NTHREADS = (EPI_WARP0 + EW) * 32;
constexpr int TMEM_COLS = 256, ACC1_COL = 0, ACC2_COL = 128, PART_COL = 192;
struct Bars {
  unsigned long long x_full[NSLOT], x_empty[NSLOT];
  unsigned long long acc1_full[2], dl_full, dl_empty, done;
  uint32_t tmem_base;
};
constexpr size_t smem_bytes(int nblk) { return (size_t)NSLOT * SLOT + nblk * tc::BLK_BYTES + XBLK + sizeof(Bars) + 4 * tc::NS * sizeof(float) + 1024; }
}  // namespace tct

// NBLK = 1 (D <= 64): the planes are 64 columns wide, so a tile is ONE block; GEMM2 then runs as M = 64 UMMAs (accumulator row d
// in lane d % 16 of sub-partition d / 16) instead of dragging a zero block through shared memory and the tensor pipe -- at
// 32 KB per tile the MMA time of two blocks alone (1408 cycles) exceeded the tile's HBM time (1360).
template <int NBLK>
__global__ void __launch_bounds__(tct::NTHREADS, 1)
k_lik_tc_tall(const __grid_constant__ tcp::Maps maps, const __nv_bfloat16* __restrict__ planes, const float* __restrict__ y, int64_t N,
              const float* __restrict__ theta, int64_t Dtot, int64_t w_off, int Dw, int Dp, int S, int family,
              float* __restrict__ Gpart, float* __restrict__ statpart, int pf_tiles, int flush_tiles, int ablate,
              long long* __restrict__ trace) {
#define BJX_TRACE(it_, slot_)                                                                 \
  do {                                                                                        \
    if (trace != nullptr && blockIdx.x == 0 && (it_) >= 16 && (it_) < 24) trace[((it_)-16) * 64 + (slot_)] = clock64(); \
  } while (0)
  using namespace tc;
  using namespace tct;
  constexpr int R = NSLOT / NBLK;
  constexpr int D = NBLK * BLK_COLS;       // 128 or 64
  constexpr int SG = NS / (EW / 4);        // samples per thread: 8
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sx = smem;
  uint8_t* sth = sx + NSLOT * SLOT;            // 2 blocks of [64 rows: t1 (32) | t2 (32)][64 bf16]
  uint8_t* sdl = sth + NBLK * BLK_BYTES;  // [128 rows][64 bf16: e1 (32 samples) | e2 (32)]
  tct::Bars* bars = (tct::Bars*)(sdl + XBLK);
  float* sred = (float*)(bars + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (N + ROWS - 1) / ROWS;
  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int b = 0; b < NSLOT; ++b) {
      mbar_init(&bars->x_full[b], 1);
      mbar_init(&bars->x_empty[b], 1);
    }
    mbar_init(&bars->acc1_full[0], 1);
    mbar_init(&bars->acc1_full[1], 1);
    mbar_init(&bars->dl_full, EW);
    mbar_init(&bars->dl_empty, 1);
    mbar_init(&bars->done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&maps.x1);
    tma_prefetch_desc(&maps.x2);
  }
  if (tid < 4 * NS) sred[tid] = 0.0f;
  if (warp == tct::MMA_WARP) tmem_alloc(&bars->tmem_base, tct::TMEM_COLS);
  for (int idx = tid; idx < NS * (D / 8); idx += NTHREADS) {
    const int s = idx / (D / 8), d8 = idx % (D / 8);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (s < S && d8 * 8 + j < Dw) ? theta[(int64_t)s * Dtot + w_off + d8 * 8 + j] : 0.0f;
    uint32_t p1[4], p2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], p1[j], p2[j]);
    const int b = d8 / 8, c = d8 % 8;
    uint8_t* blk = sth + b * BLK_BYTES;
    *reinterpret_cast<uint4*>(blk + s * 128 + ((c ^ (s & 7)) << 4)) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    const int n2 = 32 + s;
    *reinterpret_cast<uint4*>(blk + n2 * 128 + ((c ^ (n2 & 7)) << 4)) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == tct::PROD_WARP) {
    // ===================================================== TMA producer ================================================
    if (lane == 0) {
      const __nv_bfloat16* p1 = planes;
      const __nv_bfloat16* p2 = planes + (size_t)N * Dp;
      const int pf_mode = pf_tiles >> 8, pf_ahead = (pf_tiles & 255) * R;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t row0 = (blockIdx.x + it * gridDim.x) * ROWS;
        auto bulk_pf = [&]() {
          const int64_t rowp = (blockIdx.x + (it + pf_ahead) * gridDim.x) * ROWS;
          const int64_t rows = (N - rowp) < ROWS ? (N - rowp) : ROWS;
          const uint32_t bytes = (uint32_t)(rows * Dp * 2);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p1 + rowp * Dp), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p2 + rowp * Dp), "r"(bytes) : "memory");
        };
        if (pf_mode == 0 && pf_ahead > 0 && it + pf_ahead < my_tiles) bulk_pf();
        const int q0 = (int)(it % R) * NBLK;
        const uint32_t use = (uint32_t)((it / R) & 1);
        for (int b = 0; b < NBLK; ++b) {
          const int q = q0 + b;
          mbar_wait(&bars->x_empty[q], use ^ 1);  // GEMM2 of tile it - R released this slot
          BJX_TRACE(it, b);
          mbar_expect_tx(&bars->x_full[q], (uint32_t)SLOT);
          tma_load_2d(sx + q * SLOT, &maps.x1, b * BLK_COLS, (int32_t)row0, &bars->x_full[q]);
          tma_load_2d(sx + q * SLOT + XBLK, &maps.x2, b * BLK_COLS, (int32_t)row0, &bars->x_full[q]);
        }
        if (pf_mode == 1 && pf_ahead > 0 && it + pf_ahead < my_tiles) bulk_pf();
      }
    }
  } else if (warp == tct::MMA_WARP) {
    // ===================================================== MMA issuer ==================================================
    constexpr uint32_t ID_G1_N64 = make_idesc(128, NB, 0, 0);   // x1 * [t1 | t2]
    constexpr uint32_t ID_G1_N32 = make_idesc(128, NS, 0, 0);   // x2 * t1 into the first 32 columns
    constexpr uint32_t ID_G2 = make_idesc(64 * NBLK, NB, 1, 1);       // x1^T * [e1 | e2]
    constexpr uint32_t ID_G2_N32 = make_idesc(64 * NBLK, NS, 1, 1);   // x2^T * e1
    const uint64_t dx1 = make_desc(smem_u32(sx), 16, 1024), dx2 = make_desc(smem_u32(sx + XBLK), 16, 1024);
    const uint64_t dth = make_desc(smem_u32(sth), 16, 1024);
    // GEMM2 (MN-major A): the two 64-column blocks of the chunk are one slot apart
    const uint64_t tx1 = make_desc(smem_u32(sx), SLOT, 1024), tx2 = make_desc(smem_u32(sx + XBLK), SLOT, 1024);
    const uint64_t tdl = make_desc(smem_u32(sdl), XBLK, 1024);
    auto gemm1 = [&](int64_t t) {
      const uint32_t acc = tmem + tct::ACC1_COL + ((t & 1) ? NB : 0);
      for (int b = 0; b < NBLK; ++b) {
        const int q = (int)(t % R) * NBLK + b;
        mbar_wait(&bars->x_full[q], (uint32_t)((t / R) & 1));
        if (lane == 0) BJX_TRACE(t, 16 + b);
        if (!(ablate & 1) && elect_one()) {
          const uint64_t xoff = (uint64_t)((q * SLOT) >> 4), toff = (uint64_t)((b * BLK_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(acc, dx1 + xoff + 2 * k, dth + toff + 2 * k, ID_G1_N64, (b | k) != 0);
            umma_bf16(acc, dx2 + xoff + 2 * k, dth + toff + 2 * k, ID_G1_N32, 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bars->acc1_full[t & 1]);
      __syncwarp();
      if (lane == 0) BJX_TRACE(t, 24);
    };
    int phase = 0;
    if (my_tiles > 0) gemm1(0);
    for (int64_t it = 0; it < my_tiles; ++it) {
      if (it + 1 < my_tiles) gemm1(it + 1);  // other logits accumulator, slots already in the ring: runs under this tile's epilogue
      mbar_wait(&bars->dl_full, (uint32_t)(it & 1));
      tc_fence_after();
      if (lane == 0) BJX_TRACE(it, 25);
      if (elect_one()) {
        const int q0 = (int)(it % R) * NBLK;
        const uint64_t coff = (uint64_t)((q0 * SLOT) >> 4);
#pragma unroll
        for (int k = 0; k < ROWS / 16; ++k) {  // K = 16 rows = two 8-row swizzle atoms, 2048 bytes
          if (ablate & 2) continue;
          umma_bf16(tmem + tct::ACC2_COL, tx1 + coff + 128 * k, tdl + 128 * k, ID_G2, (k != 0) || !(it == 0 || phase == 0));
          umma_bf16(tmem + tct::ACC2_COL, tx2 + coff + 128 * k, tdl + 128 * k, ID_G2_N32, 1);
        }
        umma_commit(&bars->x_empty[q0]);
        if (NBLK == 2) umma_commit(&bars->x_empty[q0 + 1]);
        umma_commit(&bars->dl_empty);
      }
      __syncwarp();
      if (lane == 0) BJX_TRACE(it, 26);
      if (++phase == flush_tiles) phase = 0;
    }
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else {
    // ===================================================== epilogue =====================================================
    const int q = warp & 3;                       // TMEM sub-partition: data rows 32q .. 32q + 31 of the tile
    const int h = (warp - tct::EPI_WARP0) >> 2;   // sample group: samples 8h .. 8h + 7
    const int r = 32 * q + lane;
    const int s0 = h * SG;
    float stat[SG], prod[SG];
#pragma unroll
    for (int j = 0; j < SG; ++j) {
      stat[j] = 0.0f;
      prod[j] = 1.0f;
    }
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
    const bool logit = family == BJX_LIK_BERNOULLI_LOGITS;
    float* gp = Gpart + (size_t)blockIdx.x * S * Dw;
    // fold the GEMM2 accumulator (lanes = weight column d, columns [x*e1 (32) | x1*e2 (32)]) into the TMEM-resident partial
    auto drain = [&](bool first, bool final) {
      uint32_t fa[8], fb[8], fp[8];
      tmem_ld8(t_lane + tct::ACC2_COL + s0, fa);
      tmem_ld8(t_lane + tct::ACC2_COL + NS + s0, fb);
      if (!first) tmem_ld8(t_lane + tct::PART_COL + s0, fp);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float val = __uint_as_float(fa[j]) + __uint_as_float(fb[j]);
        if (!first) val += __uint_as_float(fp[j]);
        fp[j] = __float_as_uint(val);
      }
      if (final) {
        // lane of the accumulator = weight column (M = 128), or lanes 0-15 of each sub-partition hold columns 16q .. 16q + 15 (M = 64)
        const int d = NBLK == 2 ? r : 16 * q + (lane & 15);
        const bool lane_ok = NBLK == 2 || lane < 16;
        float* base = gp + (size_t)s0 * Dw + d;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (lane_ok && s0 + j < S && d < Dw) __stcg(base + (size_t)j * Dw, __uint_as_float(fp[j]));
      } else {
        tmem_st8(t_lane + tct::PART_COL + s0, fp);
        tmem_st_wait();
      }
    };
    int64_t row = (int64_t)blockIdx.x * ROWS + r;
    float yv = (my_tiles > 0 && row < N) ? y[row] : 0.0f;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      if (it > 0 && it % flush_tiles == 0) {
        mbar_wait(&bars->dl_empty, ph ^ 1);  // GEMM2 of the previous tile has retired; this tile's restarts the chain
        tc_fence_after();
        drain(it == flush_tiles, false);
        tc_fence_before();
      }
      const bool row_ok = row < N;
      const float ycur = yv;
      row += (int64_t)gridDim.x * ROWS;
      yv = (it + 1 < my_tiles && row < N) ? y[row] : 0.0f;
      mbar_wait(&bars->acc1_full[it & 1], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (tid == tct::EPI_WARP0 * 32) BJX_TRACE(it, 32);
      uint32_t va[SG], vb[SG];
      const uint32_t acc1 = t_lane + tct::ACC1_COL + ((it & 1) ? NB : 0);
      tmem_ld8(acc1 + s0, va);
      tmem_ld8(acc1 + NS + s0, vb);
      tmem_ld_wait();
      if (tid == tct::EPI_WARP0 * 32) BJX_TRACE(it, 33);
      float eta[SG], g[SG], ope[SG];
#pragma unroll
      for (int j = 0; j < SG; ++j) eta[j] = __uint_as_float(va[j]) + __uint_as_float(vb[j]);
      const bool all_ok = (S == NS) && (row - (int64_t)gridDim.x * ROWS - r + ROWS <= N);  // warp-uniform
      if (s0 >= S) {  // warp-uniform: this sample group does not exist (S <= 24) -- its g is zero, no link maths
#pragma unroll
        for (int j = 0; j < SG; ++j) {
          g[j] = 0.0f;
          ope[j] = 1.0f;
        }
      } else if (ablate & 4) {
#pragma unroll
        for (int j = 0; j < SG; ++j) g[j] = (row_ok && (s0 + j) < S) ? eta[j] : 0.0f;
      } else if (logit) {
        float e[SG], rc[SG];
#pragma unroll
        for (int j = 0; j < SG; ++j) e[j] = __expf(-fabsf(eta[j]));
#pragma unroll
        for (int j = 0; j < SG; ++j) ope[j] = 1.0f + e[j];
#pragma unroll
        for (int j = 0; j < SG; ++j) rc[j] = __fdividef(1.0f, ope[j]);
#pragma unroll
        for (int j = 0; j < SG; ++j) {
          const float gj = ycur - (eta[j] >= 0.0f ? rc[j] : e[j] * rc[j]);
          g[j] = (all_ok || (row_ok && (s0 + j) < S)) ? gj : 0.0f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < SG; ++j) {
          const float rr = ycur - eta[j];
          g[j] = (all_ok || (row_ok && (s0 + j) < S)) ? rr : 0.0f;
        }
      }
      uint32_t e1[SG / 2], e2[SG / 2];
#pragma unroll
      for (int j = 0; j < SG / 2; ++j) split2(g[2 * j], g[2 * j + 1], e1[j], e2[j]);
      if (tid == tct::EPI_WARP0 * 32) BJX_TRACE(it, 34);
      mbar_wait(&bars->dl_empty, ph ^ 1);  // GEMM2 of the previous tile has consumed the g buffer
      {
        uint8_t* rowp = sdl + r * 128;
        const int sw = r & 7;
        // row layout: [e1: 32 samples x 2 B | e2: 32 x 2 B]; this thread's 8 samples are one 16-byte chunk of each half
        *reinterpret_cast<uint4*>(rowp + ((h ^ sw) << 4)) = make_uint4(e1[0], e1[1], e1[2], e1[3]);
        *reinterpret_cast<uint4*>(rowp + (((4 + h) ^ sw) << 4)) = make_uint4(e2[0], e2[1], e2[2], e2[3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->dl_full);
      // ---- off the critical path: log-likelihood bookkeeping
      if (!(ablate & 4) && s0 < S) {
        if (logit) {
#pragma unroll
          for (int j = 0; j < SG; ++j) {
            const bool ok = all_ok || (row_ok && (s0 + j) < S);
            stat[j] += ok ? (ycur * eta[j] - fmaxf(eta[j], 0.0f)) : 0.0f;
            prod[j] *= ok ? ope[j] : 1.0f;
          }
          if ((it & 31) == 31) {
#pragma unroll
            for (int j = 0; j < SG; ++j) {
              stat[j] -= logf(prod[j]);
              prod[j] = 1.0f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < SG; ++j) stat[j] = fmaf(g[j], g[j], stat[j]);  // g = y - eta, already masked
        }
      }
      if (tid == tct::EPI_WARP0 * 32) BJX_TRACE(it, 35);
    }
#pragma unroll
    for (int j = 0; j < SG; ++j) {
      float v = stat[j] - logf(prod[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sred[q * NS + s0 + j] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(EW * 32) : "memory");
    if (warp == tct::EPI_WARP0 && lane < S) {
      statpart[(size_t)blockIdx.x * S + lane] = sred[lane] + sred[NS + lane] + sred[2 * NS + lane] + sred[3 * NS + lane];
    }
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    if (my_tiles > 0) {
      drain(my_tiles <= flush_tiles, true);
    } else {
      for (int i = (warp - tct::EPI_WARP0) * 32 + lane; i < S * Dw; i += EW * 32) gp[i] = 0.0f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == tct::MMA_WARP) tmem_dealloc(tmem, tct::TMEM_COLS);
#undef BJX_TRACE
}

// =========================================================================================================================
static int g_stagger = 600;          // start-up skew per phase group in SM cycles (BJX_TC_STAGGER overrides)
static int g_ablate = 0;             // timing experiments only (BJX_TC_ABLATE bitmask: 1 no GEMM1, 2 no GEMM2, 4 no link math, 8 no convert/store)
static int g_pf_mode = 0;            // 0 per-lane line prefetch by the loaders, 1 per-sector, 2 bulk prefetch by the MMA warp (BJX_TC_PFMODE)
static int g_pf = 8;                 // L2 prefetch distance in blocks (BJX_TC_PF overrides; 0 = off)
static int g_flush = tc::ACC2_FLUSH;  // tiles between flushes of the GEMM2 accumulators (BJX_TC_FLUSH overrides)
static long long* g_trace = nullptr;  // device buffer of 8 x 64 clock64 slots, or null (production)

static size_t tc_smem_bytes(int nblk) {
  return (size_t)(3 * nblk + 1) * tc::BLK_BYTES + sizeof(tc::Bars) + 4 * tc::NS * sizeof(float) + 1024;
}

static int64_t pad64(int64_t v) { return (v + 63) / 64 * 64; }
// columns are a multiple of 128 in [128, 512]: both variants of the kernel apply (raw fp32 X or prepared planes)
static bool tc_exact_width(const BjxElboArgs& a) { return a.Dw >= 128 && a.Dw <= 512 && a.Dw % 128 == 0; }

// Two CTAs per tile, each on half of the columns (k_lik_tc_tma<..., PAIR>): every width in (512, 1024]; BJX_TC_PAIR=1 also
// sends 256 < D <= 512 that way (development switch: two half-width tiles in the slots instead of one full-width tile).
static int tc_pair_forced() {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("BJX_TC_PAIR");
    forced = e ? atoi(e) : 0;
  }
  return forced;
}
static bool tc_pair(const BjxElboArgs& a) {
  if (a.row_index != nullptr) return false;
  return a.Dw > 512 || (tc_pair_forced() == 1 && a.Dw > 256 && a.x_planes != nullptr);
}
static int tc_max_clusters();

bool tc_eligible(const BjxElboArgs& a) {
  if (a.likelihood != BJX_LIK_BERNOULLI_LOGITS && a.likelihood != BJX_LIK_GAUSSIAN) return false;
  if (a.S < 1) return false;
  if (a.S > tc::NS) {
    // More than 32 samples: TMEM holds G^T for 32 samples x 512 columns, so the single pass runs once per slice of 32 samples
    // (X is read once per slice).  That beats the two-pass GEMM path -- X read twice plus the round trip of g through HBM,
    // 8 * pad64(S) bytes per row -- only for a few slices of a narrow X: k * Dp * 1.15 <= 2 * Dp + 2 * Sp, Dp >= 128, k <= 3.
    const int64_t k = (a.S + tc::NS - 1) / tc::NS, Dp = pad64(a.Dw), Sp = pad64(a.S);
    if (k > 3 || Dp < 128 || 115 * k * Dp > 100 * (2 * Dp + 2 * Sp)) return false;
  }
  if (a.precision == BJX_PREC_BF16X3) return false;  // the three-term split runs on the two-pass GEMM path
  // rows read through an index (minibatches): either the register-converting kernel on raw fp32 X, or -- with the FULL
  // dataset's planes and its row count -- the TMA-fed kernel gathering plane rows (tile::gather4)
  if (a.row_index != nullptr && a.x_planes != nullptr && a.N_full < 1) return false;
  if (tc_exact_width(a)) {
    if (a.x_planes == nullptr && (a.ldx % 4 != 0 || ((uintptr_t)a.X & 15) != 0)) return false;  // 128-bit row loads of raw X
    return true;
  }
  // Any other width in [16, 512] runs on the TMA-fed variant over the split-bf16 planes (columns padded to 64 by
  // bjx_prepare_x_planes): the kernel is instantiated for the next even number of 64-column blocks, and the blocks beyond
  // the plane's extent are zero-filled by the copy engine (out-of-bounds TMA boxes cost no HBM traffic).  The raw-fp32 /
  // row-index variant has no such fill, so a step without prepared planes splits X into the workspace first.
  if (a.Dw < 16 || a.Dw > 1024) return false;
  if (a.Dw > 512 && (a.row_index != nullptr || tc_pair_forced() < 0)) return false;  // the pair kernel has no gather variant (BJX_TC_PAIR=-1: off)
  if (a.Dw > 512 && tc_max_clusters() <= 0) return false;  // no cluster of two CTAs can be resident here: the two-pass GEMM takes the shape
  if (a.row_index != nullptr && a.x_planes == nullptr) return false;  // no planes to gather from, and no in-step split of a full dataset
  return true;
}

int tc_plan(const BjxElboArgs& a, Plan& p) {
  const int sms = sm_count();
  if (sms <= 0) return fail(BJX_ERR_CUDA, "no CUDA device");
  const int64_t tiles = (a.N + tc::TILE_ROWS - 1) / tc::TILE_ROWS;
  int64_t g = tiles < sms ? tiles : sms;
  if (tc_pair(a)) {  // one partial slot per CLUSTER of two CTAs
    const int mc = tc_max_clusters();
    if (mc <= 0) return fail(BJX_ERR_CUDA, "no cluster of two CTAs can be resident");
    g = tiles < mc ? tiles : mc;
  }
  if (g < 1) g = 1;
  p.tc_ctas = (int)g;
  p.tc_terms = 2;
  return BJX_OK;
}

// workspace for the operand planes when a width outside {128, 256, 384, 512} arrives without prepared planes
size_t tc_theta_bytes(const BjxElboArgs& a, const Plan&) {
  return (!tc_exact_width(a) && a.x_planes == nullptr && a.row_index == nullptr) ? x_planes_bytes(a.N, a.Dw) : 0;
}

// how many clusters of two CTAs of the paired kernel can be resident at once (the largest instance: 227 KB, 320 threads)
static int tc_max_clusters() {
  static int cached = -1;
  if (cached >= 0) return cached;
  const int smem = 227 * 1024;
  if (cudaFuncSetAttribute(k_lik_tc_tma<8, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2u * 148u);
  cfg.blockDim = dim3((tcp::EPI_WARP0 + 8) * 32);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_lik_tc_tma<8, 8, false, true>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  const int sms = sm_count();
  if (n > sms / 2) n = sms / 2;
  cached = n;
  return n;
}

int launch_lik_tc(const BjxElboArgs& a_in, const Plan& p, char* ws, cudaStream_t st, int64_t s0, int64_t S_slice) {
  // one sample slice [s0, s0 + S_slice) of the step (S_slice <= 32): its theta rows, and its own block of the partial buffers
  BjxElboArgs a = a_in;
  a.S = S_slice;
  static bool env_read = false;
  if (!env_read) {
    if (const char* e = getenv("BJX_TC_STAGGER")) g_stagger = atoi(e);
    if (const char* e = getenv("BJX_TC_ABLATE")) g_ablate = atoi(e);
    if (const char* e = getenv("BJX_TC_PF")) g_pf = atoi(e);
    if (const char* e = getenv("BJX_TC_PFMODE")) g_pf_mode = atoi(e);
    if (const char* e = getenv("BJX_TC_FLUSH")) g_flush = atoi(e) > 0 ? atoi(e) : 1;
    env_read = true;
  }
  const float* theta = (const float*)(ws + p.o_theta) + s0 * a.D;
  float* Gpart = (float*)(ws + p.o_Gpart) + (size_t)p.PG * s0 * a.Dw;
  float* statpart = (float*)(ws + p.o_statpart) + (size_t)p.PS * s0;
  const int64_t Dp = pad64(a.Dw);
  const bool pair = tc_pair(a);
  // 64-column blocks per CTA, even (GEMM2 works on 128-column chunks); a pair splits the blocks in two halves
  const int nblk = pair ? (int)((((Dp / tc::BLK_COLS + 1) / 2) + 1) & ~1ll) : (int)(((Dp / tc::BLK_COLS) + 1) & ~1ll);
  const size_t smem = tc_smem_bytes(nblk);
  const void* x_planes = a.x_planes;
  if (x_planes == nullptr && !tc_exact_width(a)) {
    if (s0 == 0) {  // the in-step split is made once, by the first slice
      int rc = split_x_planes(a.X, a.N, a.Dw, a.ldx, ws + p.o_tc_theta, st);
      if (rc) return rc;
    }
    x_planes = ws + p.o_tc_theta;
  }
  if (x_planes != nullptr) {
    // dataset prepared as split-bf16 planes: the TMA-fed variant (two extra block slots, see the kernel)
    const int nslot = 8 / nblk > 1 ? 8 : nblk;  // D = 128, 256: the eight slots are a ring of tiles (see the kernel)
    const size_t smem = (size_t)(2 * nslot + nblk + 3) * tc::BLK_BYTES + sizeof(tcp::BarsP) + 4 * tc::NS * sizeof(float) + 16 +
                        (nslot > nblk ? 2 : 1) * 8 * 256 * sizeof(float) + 1024 +
                        (a.row_index != nullptr ? 4 * tc::TILE_ROWS * sizeof(int32_t) : 0) +  // + the gather variant's index ring
                        (pair && nblk == 6 ? 2 * 8 * 256 * sizeof(float) : 0);  // + a pair's receive buffers where no spare slot is left over
    static int pf_tiles = -1;
    if (pf_tiles < 0) {
      const char* e = getenv("BJX_TC_PF_TILES");
      pf_tiles = e ? atoi(e) : 257;  // low byte: distance in tiles; bits 8+: mode (see the producer warp); 257 = one tile ahead, issued after the tile's loads
    }
    // folding the GEMM2 accumulators into the TMEM-resident partial is free of memory traffic: do it twice as often as
    // the raw-fp32 variant (bias of the truncating accumulate ~0.9e-6 instead of 1.7e-6)
    static int pf_idx = -1;  // gather variant: how many tiles ahead whole plane rows are touched into L2 (BJX_TC_PF_IDX; 0 = off)
    if (pf_idx < 0) {
      const char* e = getenv("BJX_TC_PF_IDX");
      pf_idx = e ? atoi(e) : 0;  // 0 off (default: no mode gained anything, profiles/r02_minibatch.json) | 1 per line | 2 per sector | 3 bulk per row
    }
    const int flush_tma = getenv("BJX_TC_FLUSH") ? g_flush : 8;
    tcp::Maps maps;
    const __nv_bfloat16* planes = (const __nv_bfloat16*)x_planes;
    int rc;
    const bool idx = a.row_index != nullptr;  // the planes are then the FULL dataset's (N_full rows), read through the index
    const int64_t plane_rows = idx ? a.N_full : a.N;
    if ((rc = make_bf16_tensor_map(&maps.x1, planes, plane_rows, Dp, Dp, tc::BLK_COLS, tc::TILE_ROWS))) return rc;
    if ((rc = make_bf16_tensor_map(&maps.x2, planes + (size_t)plane_rows * Dp, plane_rows, Dp, Dp, tc::BLK_COLS, tc::TILE_ROWS))) return rc;
#define BJX_TCP_LAUNCH_I(NBLK, EW, IDX)                                                                                        \
  do {                                                                                                                    \
    static DeviceOnce attr_set;                                                                                         \
    if (!attr_set.test_and_set()) {                                                                                                      \
      BJX_CUDA_TRY(cudaFuncSetAttribute(k_lik_tc_tma<NBLK, EW, IDX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    }                                                                                                                     \
    k_lik_tc_tma<NBLK, EW, IDX, false><<<p.tc_ctas, (tcp::EPI_WARP0 + EW + (IDX ? tcp::GATHER_WARPS - 1 : 0)) * 32, smem, st>>>(maps, planes, a.y, a.N, theta, a.D, a.w_offset, (int)a.Dw, (int)Dp, (int)a.S, \
                                                              a.likelihood, Gpart, statpart, idx ? pf_idx : pf_tiles, flush_tma, g_ablate, g_trace, a.row_index, plane_rows); \
  } while (0)
    // a pair of CTAs per tile: p.tc_ctas counts clusters
#define BJX_TCP_LAUNCH_PAIR(NBLK)                                                                                                \
  do {                                                                                                                            \
    static DeviceOnce attr_set;                                                                                                   \
    if (!attr_set.test_and_set()) {                                                                                               \
      BJX_CUDA_TRY(cudaFuncSetAttribute(k_lik_tc_tma<NBLK, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    }                                                                                                                             \
    cudaLaunchConfig_t cfg = {};                                                                                                  \
    cfg.gridDim = dim3(2u * (unsigned)p.tc_ctas);                                                                                 \
    cfg.blockDim = dim3((tcp::EPI_WARP0 + 8) * 32);                                                                               \
    cfg.dynamicSmemBytes = smem;                                                                                                  \
    cfg.stream = st;                                                                                                              \
    cudaLaunchAttribute attr[1];                                                                                                  \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                                             \
    attr[0].val.clusterDim.x = 2;                                                                                                 \
    attr[0].val.clusterDim.y = 1;                                                                                                 \
    attr[0].val.clusterDim.z = 1;                                                                                                 \
    cfg.attrs = attr;                                                                                                             \
    cfg.numAttrs = 1;                                                                                                             \
    BJX_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_lik_tc_tma<NBLK, 8, false, true>, maps, planes, a.y, a.N, theta, a.D, a.w_offset,     \
                                    (int)a.Dw, (int)Dp, (int)a.S, (int)a.likelihood, Gpart, statpart, pf_tiles, flush_tma, g_ablate,   \
                                    g_trace, (const int32_t*)nullptr, plane_rows));                                               \
  } while (0)
    if (pair) {
      switch (nblk) {
        case 4: BJX_TCP_LAUNCH_PAIR(4); break;
        case 6: BJX_TCP_LAUNCH_PAIR(6); break;
        case 8: BJX_TCP_LAUNCH_PAIR(8); break;
        default: return fail(BJX_ERR_UNSUPPORTED, "paired tcgen05 kernel needs 256 < D <= 1024");
      }
      BJX_LAUNCH_CHECK("k_lik_tc_tma<pair>");
      return BJX_OK;
    }
#define BJX_TCP_LAUNCH(NBLK, EW)                  \
  do {                                            \
    if (idx) BJX_TCP_LAUNCH_I(NBLK, EW, true);    \
    else BJX_TCP_LAUNCH_I(NBLK, EW, false);       \
  } while (0)
    // Epilogue warps: eight everywhere except D = 128 with the Bernoulli-logit link, where the tile is so short that the
    // epilogue's latency chain is the bound and sixteen warps (four finished values per thread) hide more of it: 0.65 ->
    // 0.75-0.77 of the roofline in an A/B on one box; the Gaussian link and D >= 256 do not gain.  BJX_TC_EPI_WARPS forces.
    static int forced_ew = -1;
    if (forced_ew < 0) {
      const char* e = getenv("BJX_TC_EPI_WARPS");
      forced_ew = e ? (atoi(e) == 16 ? 16 : 8) : 0;
    }
    const int epi_warps = forced_ew ? forced_ew : ((nblk == 2 && a.likelihood == BJX_LIK_BERNOULLI_LOGITS) ? 16 : 8);
#define BJX_TCP_LAUNCH_EW(NBLK)                 \
  do {                                          \
    if (epi_warps == 16) BJX_TCP_LAUNCH(NBLK, 16); \
    else BJX_TCP_LAUNCH(NBLK, 8);               \
  } while (0)
    // D <= 128: tall tiles (128 rows, no partner swap; see k_lik_tc_tall).  BJX_TC_TALL=0 keeps the 64-row kernel.
    static int tall = -1;
    if (tall < 0) {
      const char* e = getenv("BJX_TC_TALL");
      tall = e ? atoi(e) : 1;
    }
    if (nblk == 2 && !idx && tall && !forced_ew) {
      const bool one_block = Dp <= tc::BLK_COLS;
      const size_t smem_tall = tct::smem_bytes(one_block ? 1 : 2);
      tcp::Maps mt;
      if ((rc = make_bf16_tensor_map(&mt.x1, planes, plane_rows, Dp, Dp, tc::BLK_COLS, tct::ROWS))) return rc;
      if ((rc = make_bf16_tensor_map(&mt.x2, planes + (size_t)plane_rows * Dp, plane_rows, Dp, Dp, tc::BLK_COLS, tct::ROWS))) return rc;
      static DeviceOnce attr_set;
      if (!attr_set.test_and_set()) {
        BJX_CUDA_TRY(cudaFuncSetAttribute(k_lik_tc_tall<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tct::smem_bytes(1)));
        BJX_CUDA_TRY(cudaFuncSetAttribute(k_lik_tc_tall<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tct::smem_bytes(2)));
      }
      const int flush_tall = getenv("BJX_TC_FLUSH") ? g_flush : 4;  // 4 tiles of 128 rows = the chain length of 8 tiles of 64
      if (one_block)
        k_lik_tc_tall<1><<<p.tc_ctas, tct::NTHREADS, smem_tall, st>>>(mt, planes, a.y, a.N, theta, a.D, a.w_offset, (int)a.Dw, (int)Dp, (int)a.S,
                                                                      a.likelihood, Gpart, statpart, pf_tiles, flush_tall, g_ablate, g_trace);
      else
        k_lik_tc_tall<2><<<p.tc_ctas, tct::NTHREADS, smem_tall, st>>>(mt, planes, a.y, a.N, theta, a.D, a.w_offset, (int)a.Dw, (int)Dp, (int)a.S,
                                                                      a.likelihood, Gpart, statpart, pf_tiles, flush_tall, g_ablate, g_trace);
      BJX_LAUNCH_CHECK("k_lik_tc_tall");
      return BJX_OK;
    }
    // D <= 128: two epilogue groups on alternate tiles (see the kernel).  A/B on one box (profiles/r02_envelope_narrow_groups.json):
    // the Gaussian link gains (0.71-0.72 -> 0.75-0.77 of the roofline at D = 128), the Bernoulli-logit link does not (0.70 either
    // way: sixteen warps on ONE tile shorten its longer MUFU chain just as much), so it keeps one group.  BJX_TC_EPI_GROUPS=1 / 2 forces.
    static int groups = -1;
    if (groups < 0) {
      const char* e = getenv("BJX_TC_EPI_GROUPS");
      groups = e ? atoi(e) : 0;
    }
    const bool two_groups = groups ? groups == 2 : a.likelihood == BJX_LIK_GAUSSIAN;
    if (nblk == 2 && !idx && two_groups && !forced_ew) {
      const size_t smem2 = smem + tc::BLK_BYTES + 4 * tc::NS * sizeof(float);
      static DeviceOnce attr_set;
      if (!attr_set.test_and_set()) {
        BJX_CUDA_TRY(cudaFuncSetAttribute(k_lik_tc_tma<2, 16, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      }
      k_lik_tc_tma<2, 16, false, false, 2><<<p.tc_ctas, (tcp::EPI_WARP0 + 16) * 32, smem2, st>>>(
          maps, planes, a.y, a.N, theta, a.D, a.w_offset, (int)a.Dw, (int)Dp, (int)a.S, a.likelihood, Gpart, statpart, pf_tiles, flush_tma,
          g_ablate, g_trace, a.row_index, plane_rows);
      BJX_LAUNCH_CHECK("k_lik_tc_tma<two groups>");
      return BJX_OK;
    }
    switch (nblk) {
      case 2: BJX_TCP_LAUNCH_EW(2); break;
      case 4: BJX_TCP_LAUNCH_EW(4); break;
      case 6: BJX_TCP_LAUNCH_EW(6); break;
      case 8: BJX_TCP_LAUNCH_EW(8); break;
      default: return fail(BJX_ERR_UNSUPPORTED, "tcgen05 kernel needs D in {128, 256, 384, 512}");
    }
#undef BJX_TCP_LAUNCH_EW
#undef BJX_TCP_LAUNCH
#undef BJX_TCP_LAUNCH_I
#undef BJX_TCP_LAUNCH_PAIR
    BJX_LAUNCH_CHECK("k_lik_tc_tma");
    return BJX_OK;
  }
#define BJX_TC_LAUNCH(NBLK)                                                                                           \
  do {                                                                                                                \
    static DeviceOnce attr_set;                                                                                     \
    if (!attr_set.test_and_set()) {                                                                                                  \
      BJX_CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<NBLK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
      BJX_CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<NBLK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    }                                                                                                                 \
    if (a.row_index != nullptr)                                                                                       \
      k_lik_tc<NBLK, true><<<p.tc_ctas, tc::THREADS, smem, st>>>(a.X, a.ldx, a.y, a.N, theta, a.D, a.w_offset, (int)a.S, a.likelihood, Gpart, \
                                                                 statpart, g_trace, g_stagger, g_ablate, g_pf, g_flush, g_pf_mode, a.row_index); \
    else                                                                                                              \
      k_lik_tc<NBLK, false><<<p.tc_ctas, tc::THREADS, smem, st>>>(a.X, a.ldx, a.y, a.N, theta, a.D, a.w_offset, (int)a.S, a.likelihood, Gpart, \
                                                                  statpart, g_trace, g_stagger, g_ablate, g_pf, g_flush, g_pf_mode, nullptr); \
  } while (0)
  switch (nblk) {
    case 2: BJX_TC_LAUNCH(2); break;
    case 4: BJX_TC_LAUNCH(4); break;
    case 6: BJX_TC_LAUNCH(6); break;
    case 8: BJX_TC_LAUNCH(8); break;
    default: return fail(BJX_ERR_UNSUPPORTED, "tcgen05 kernel needs D in {128, 256, 384, 512}");
  }
#undef BJX_TC_LAUNCH
  BJX_LAUNCH_CHECK("k_lik_tc");
  return BJX_OK;
}

}  // namespace bjx

extern "C" int bjx_debug_set_trace_buffer(void* device_buffer_4096_bytes) {
  bjx::g_trace = (long long*)device_buffer_4096_bytes;
  return BJX_OK;
}

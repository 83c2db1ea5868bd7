// Internal plan / workspace layout shared by the ELBO step orchestrator and the streaming-likelihood kernels.
#pragma once
#include "bjx_common.cuh"

namespace bjx {

enum { KERNEL_FP32_TILED = 0, KERNEL_FP32_SMALLD = 1, KERNEL_TC_BF16 = 2, KERNEL_COIN = 3, KERNEL_TC_GEMM = 4 };

constexpr int SD_CHUNK_ROWS = 256;  // rows staged per shared-memory chunk by the small-D kernel

// limits of the one-block step for tiny problems (k_tiny_coin_step): S * D, S, N
constexpr int TINY_SD = 1024, TINY_S = 256, TINY_N = 4096;

struct Plan {
  int kernel;        // KERNEL_*
  int n_qp_blocks;   // prologue blocks = partial sums of (q - p~)
  int PG, PS;        // partial slots for G[S,Dw] and stat[S]
  int64_t Pn;        // number of raw-scale parameters (D or D*D)
  // small-D
  int sd_rec4, sd_SB, sd_grid_x, sd_grid_y;
  // fp32 tiled
  int tA_grid_x, tA_grid_y, tB_kslots, tB_tiles;
  int64_t tB_rows_per_slot;
  // tensor-core
  int tc_ctas, tc_terms;
  // two-pass TMA-fed tensor-core GEMM path (bjx_gemm_tc.cu)
  int gm_grid1, gm_grid2, gm_splits, gm_cg, gm_nt;  // gm_cg = 2: CTA pairs (cta_group::2); gm_nt = split terms per operand
  int64_t gm_Dp, gm_Sp;  // padded plane pitches (elements)
  int fr_tc;             // full-rank contractions on the tensor cores
  int tiny;              // the whole step in one block (k_tiny_coin_step)
  // byte offsets into the workspace
  size_t o_eps, o_z, o_theta, o_bprime, o_gprior, o_gz, o_qp, o_aux, o_G, o_stat, o_Gpart, o_statpart, o_gbuf, o_tc_theta, o_gm_x, o_gm_t, o_gm_g, o_fr_tc, o_lr,
      o_stage_key, o_stage_params, o_stage_out, total;
};

int make_plan(const BjxElboArgs& a, Plan& p);

// streaming likelihood launchers: fill Gpart[PG][S][Dw] (sum_n dll/deta * X) and statpart[PS][S]
int smalld_occupancy(int rec4);  // resident CTAs per SM of k_lik_smalld<rec4>
int launch_lik_smalld(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st);
int launch_lik_tiled(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st);
int launch_lik_tc(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st, int64_t s0, int64_t S_slice);
bool tc_eligible(const BjxElboArgs& a);
int tc_plan(const BjxElboArgs& a, Plan& p);          // fills tc_ctas / tc_terms
size_t tc_theta_bytes(const BjxElboArgs& a, const Plan& p);
// two-pass GEMM path for shapes outside the fused kernel's envelope (S > 32 or D > 512), and the full-rank contractions
bool gemm_eligible(const BjxElboArgs& a);
int gemm_plan(const BjxElboArgs& a, Plan& p);  // fills gm_*
int launch_lik_gemm(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st);
int split_x_planes(const float* X, int64_t N, int64_t Dw, int64_t ldx, void* planes, cudaStream_t st);
size_t x_planes_bytes(int64_t N, int64_t Dw);
int split_x_planes_rows(const float* X_rows, int64_t rows, int64_t Dw, int64_t ldx, void* planes, int64_t N_total, int64_t row0, cudaStream_t st);
bool fullrank_tc_eligible(const BjxElboArgs& a);
size_t fullrank_tc_bytes(const BjxElboArgs& a);
int launch_fullrank_z_tc(const BjxElboArgs& a, const float* eps, float* z, char* scratch, cudaStream_t st);
int launch_fullrank_dM_tc(const BjxElboArgs& a, const float* gz, float* dM, char* scratch, float entropy_weight, cudaStream_t st);

// low-rank variational family (bjx_lowrank.cu)
size_t lowrank_bytes(const BjxElboArgs& a);
int launch_lowrank_sample(const BjxElboArgs& a, const float* eps, float* z, char* scratch, cudaStream_t st);
const float* lowrank_logdet_ptr(const BjxElboArgs& a, const char* scratch);
int launch_lowrank_grads(const BjxElboArgs& a, const float* gz, const float* eps, char* scratch, float entropy_weight,
                         float* d_raw, cudaStream_t st);

size_t lowrank_posterior_bytes(const BjxElboArgs& a);
int launch_lowrank_posterior_log_prob(const BjxElboArgs& a, const float* theta, int64_t n, float* out, cudaStream_t st);

// per-datum link functions (the elementwise part of advi.py:91 for the recognised families)
//   Bernoulli(logits=eta): f = y*eta - softplus(eta)   (== -(1-y) softplus(eta) - y softplus(-eta)),  g = y - sigmoid(eta)
//   Normal(eta, tau):      f = (y - eta)^2 (tau applied in the tail),                                   g = y - eta
__device__ __forceinline__ void link_eval(int family, float y, float eta, float& f, float& g) {
  if (family == BJX_LIK_BERNOULLI_LOGITS) {
    f = y * eta - softplus_f(eta);
    g = y - sigmoid_f(eta);
  } else {
    const float r = y - eta;
    f = r * r;
    g = r;
  }
}

// bjx_minibatch.cu
int launch_minibatch_gather(const BjxMinibatch& mb, const uint32_t* key, const int64_t* key_index, int layout, cudaStream_t st);

}  // namespace bjx

// Internal plan / workspace layout shared by the ELBO step orchestrator and the streaming-likelihood kernels.
#pragma once
#include "bjx_common.cuh"

namespace bjx {

enum { KERNEL_FP32_TILED = 0, KERNEL_FP32_SMALLD = 1, KERNEL_TC_BF16 = 2, KERNEL_COIN = 3 };

constexpr int SD_CHUNK_ROWS = 256;  // rows staged per shared-memory chunk by the small-D kernel

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
  // byte offsets into the workspace
  size_t o_eps, o_z, o_theta, o_bprime, o_gprior, o_gz, o_qp, o_aux, o_G, o_stat, o_Gpart, o_statpart, o_gbuf, o_tc_theta,
      o_stage_key, o_stage_params, o_stage_out, total;
};

int make_plan(const BjxElboArgs& a, Plan& p);

// streaming likelihood launchers: fill Gpart[PG][S][Dw] (sum_n dll/deta * X) and statpart[PS][S]
int smalld_occupancy(int rec4);  // resident CTAs per SM of k_lik_smalld<rec4>
int launch_lik_smalld(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st);
int launch_lik_tiled(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st);
int launch_lik_tc(const BjxElboArgs& a, const Plan& p, char* ws, cudaStream_t st);
bool tc_eligible(const BjxElboArgs& a);
int tc_plan(const BjxElboArgs& a, Plan& p);          // fills tc_ctas / tc_terms
size_t tc_theta_bytes(const BjxElboArgs& a, const Plan& p);

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

}  // namespace bjx

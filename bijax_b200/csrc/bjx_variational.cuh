// Kernels 2-4 and 6 (the O(S*D) parts): fused variational sampling, bijector forward + log|det J|, log-prior and
// entropy terms, and the reparameterisation-gradient tail.  Restates, per coordinate, what the reference computes in
//   advi.py:82-83 (transform_dist_params + q.sample), advi.py:86 (q.log_prob), advi.py:88 (log_prob_dist of the
//   unconstrained prior, core.py:114-121,133-151), advi.py:89 (transform_tree) and the VJP of all of it (utils.py:7).
#pragma once
#include "bjx_common.cuh"

namespace bjx {

#define BJX_HALF_LOG_2PI 0.91893853320467274178f

struct ScaleEval {
  float sigma, dsigma, dlogsigma;  // sigma(rho), d sigma / d rho, d log sigma / d rho
};

__device__ __forceinline__ ScaleEval eval_scale(float rho, int bij) {
  ScaleEval r;
  if (bij == BJX_SCALE_EXP) {
    r.sigma = expf(rho);
    r.dsigma = r.sigma;
    r.dlogsigma = 1.0f;
  } else if (bij == BJX_SCALE_SOFTPLUS) {
    r.sigma = softplus_f(rho);
    r.dsigma = sigmoid_f(rho);
    r.dlogsigma = r.dsigma / r.sigma;
  } else {
    r.sigma = rho;
    r.dsigma = 1.0f;
    r.dlogsigma = 1.0f / rho;
  }
  return r;
}

struct CoordEval {
  float theta;    // b(z)
  float dtheta;   // d theta / d z
  float logp;     // log p(b(z)) + fldj_b(z)   (density of z under the unconstrained prior, core.py:117)
  float dlogp;    // d logp / d z
  float ldj;      // fldj_b(z) alone and its z-derivative (MAP.neg_log_joint uses the prior density WITHOUT it, map.py:30)
  float dldj;
  // log(theta), log(1 - theta) and their z-derivatives, free of the saturation of sigmoid in fp32 (used when theta
  // feeds tfd.Bernoulli(probs=theta): the coin-toss likelihood)
  float lt, l1t, dlt, dl1t;
};

// Forward pass of a bijector chain with its log-det-Jacobian and both first derivatives, then the prior at theta.
__device__ __forceinline__ CoordEval eval_coord(float z, const BjxCoord c) {
  float x = z, ldj = 0.0f, dx = 1.0f, dldj = 0.0f;
  float x_in_last = z;  // input of the LAST op (for the overflow-free closed forms below)
  float dx_in_last = 1.0f;
  uint32_t last_op = BJX_BIJ_IDENTITY;
  uint32_t chain = c.chain;
#pragma unroll 1
  while (chain & 0xFu) {
    const uint32_t op = chain & 0xFu;
    chain >>= 4;
    x_in_last = x;
    dx_in_last = dx;
    last_op = op;
    float y, dy, f, df;
    if (op == BJX_BIJ_EXP) {
      y = expf(x);
      dy = y;
      f = x;
      df = 1.0f;
    } else if (op == BJX_BIJ_SOFTPLUS) {
      const float sg = sigmoid_f(x);
      y = softplus_f(x);
      dy = sg;
      f = -softplus_f(-x);
      df = 1.0f - sg;
    } else {  // sigmoid
      const float sg = sigmoid_f(x);
      y = sg;
      dy = sg * sigmoid_f(-x);
      f = -softplus_f(-x) - softplus_f(x);
      df = 1.0f - 2.0f * sg;
    }
    ldj += f;
    dldj += df * dx;
    dx *= dy;
    x = y;
  }
  CoordEval r;
  r.theta = x;
  r.dtheta = dx;
  float lp, dlp_dz;
  if (c.prior == BJX_PRIOR_NORMAL) {
    const float t = (x - c.p0) / c.p1;
    lp = -0.5f * t * t + c.lognorm;
    dlp_dz = -(t / c.p1) * dx;
  } else if (c.prior == BJX_PRIOR_BETA) {
    const float am1 = c.p0 - 1.0f, bm1 = c.p1 - 1.0f;
    if (last_op == BJX_BIJ_SIGMOID) {
      // log(theta) = -softplus(-x), log(1-theta) = -softplus(x): no cancellation when theta rounds to 0 or 1
      const float lt = -softplus_f(-x_in_last), l1t = -softplus_f(x_in_last);
      lp = (am1 == 0.0f ? 0.0f : am1 * lt) + (bm1 == 0.0f ? 0.0f : bm1 * l1t) + c.lognorm;
      dlp_dz = (am1 * sigmoid_f(-x_in_last) - bm1 * x) * dx_in_last;
    } else {
      lp = (am1 == 0.0f ? 0.0f : am1 * logf(x)) + (bm1 == 0.0f ? 0.0f : bm1 * log1pf(-x)) + c.lognorm;
      dlp_dz = (am1 / x - bm1 / (1.0f - x)) * dx;
    }
  } else {  // gamma(concentration p0, rate p1)
    const float am1 = c.p0 - 1.0f;
    if (last_op == BJX_BIJ_EXP) {
      lp = (am1 == 0.0f ? 0.0f : am1 * x_in_last) - c.p1 * x + c.lognorm;
      dlp_dz = (am1 - c.p1 * x) * dx_in_last;
    } else {
      lp = (am1 == 0.0f ? 0.0f : am1 * logf(x)) - c.p1 * x + c.lognorm;
      dlp_dz = (am1 / x - c.p1) * dx;
    }
  }
  r.logp = lp + ldj;
  r.dlogp = dlp_dz + dldj;
  r.ldj = ldj;
  r.dldj = dldj;
  if (last_op == BJX_BIJ_SIGMOID) {
    r.lt = -softplus_f(-x_in_last);
    r.l1t = -softplus_f(x_in_last);
    r.dlt = sigmoid_f(-x_in_last) * dx_in_last;
    r.dl1t = -x * dx_in_last;
  } else {
    r.lt = logf(x);
    r.l1t = log1pf(-x);
    r.dlt = dx / x;
    r.dl1t = -dx / (1.0f - x);
  }
  return r;
}

}  // namespace bjx

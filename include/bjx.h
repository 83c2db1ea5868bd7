/* bjx.h -- C ABI of the B200-native ADVI engine (libbjx.so).
 *
 * This is the drop-in boundary for bijax's one hot path: the reparameterised
 * Monte-Carlo ELBO and its gradient,
 *     ADVI.loss_fn                     /root/reference/bijax/advi.py:80-95
 *     under jax.value_and_grad          /root/reference/bijax/utils.py:7,25
 * plus the PRNG calls that path makes (jax.random.normal / split / PRNGKey,
 * call sites advi.py:77,83 and utils.py:30,53-55,70).
 *
 * Everything is plain pointers and sizes: no torch, no XLA types.  Three
 * callers bind it:  (i) the XLA-FFI shim (bijax_b200/csrc/xla_ffi_shim.cc,
 * built only where jaxlib's headers exist), (ii) the ctypes host layer
 * (bijax_b200/_native.py) used by every test and by bench.py, (iii) any
 * C/C++ driver.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - all array pointers are DEVICE pointers unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; every call only ENQUEUES
 *     work on it (no device/stream synchronisation, no allocation), so a call
 *     sequence is CUDA-graph capturable -- except the *_host entry points,
 *     which copy and synchronise by design;
 *   - return value 0 = success, negative = BJX_ERR_*; bjx_last_error() gives
 *     the message for the calling thread.  Nothing throws across the ABI.
 *   - float32 everywhere except PRNG keys / bits (uint32).
 */
#ifndef BJX_H_
#define BJX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BJX_ABI_VERSION 4

#if defined(__GNUC__)
#define BJX_API __attribute__((visibility("default")))
#else
#define BJX_API
#endif

/* ---- error codes ------------------------------------------------------- */
#define BJX_OK 0
#define BJX_ERR_INVALID_ARGUMENT (-1)
#define BJX_ERR_CUDA (-2)
#define BJX_ERR_UNSUPPORTED (-3)
#define BJX_ERR_WORKSPACE (-4)

/* ---- jax.random counter layouts (jax_threefry_partitionable) ---------- */
#define BJX_LAYOUT_LEGACY 0        /* default before JAX 0.5 */
#define BJX_LAYOUT_PARTITIONABLE 1 /* default from JAX 0.5; counter = 64-bit flat index */

/* ---- variational family: core.py:179-187 (mean field), 204-212 (full rank) */
#define BJX_VI_MEAN_FIELD 0
#define BJX_VI_FULL_RANK 1
#define BJX_VI_LOW_RANK 2 /* core.py:190-201: MultivariateNormalDiagPlusLowRank, scale = diag(d) + U U^T, rank <= 64 */

/* ---- bijector on the raw scale parameter (ordered_posterior_bijectors[1],
 *      advi.py:33,45,65-67; default Exp, core.py:180-181) ------------------ */
#define BJX_SCALE_EXP 0
#define BJX_SCALE_SOFTPLUS 1
#define BJX_SCALE_IDENTITY 2 /* full-rank default: identity on scale_tril */

/* ---- bijector ops of prior_constraints (tfb.*), packed 4 bits each into
 *      BjxCoord.chain, lowest nibble applied FIRST; 0 terminates.
 *      tfb.Chain([b1, b2]) = b1 o b2  ->  chain = op(b2) | op(b1) << 4 ------ */
#define BJX_BIJ_IDENTITY 0
#define BJX_BIJ_EXP 1
#define BJX_BIJ_SOFTPLUS 2
#define BJX_BIJ_SIGMOID 3

/* ---- prior families (tfd.*), one table row per latent coordinate -------- */
#define BJX_PRIOR_NORMAL 0 /* p0 = loc, p1 = scale   (also each coordinate of MultivariateNormalDiag) */
#define BJX_PRIOR_BETA 1   /* p0 = concentration1, p1 = concentration0 */
#define BJX_PRIOR_GAMMA 2  /* p0 = concentration,  p1 = rate */

typedef struct BjxCoord {
  uint32_t prior; /* BJX_PRIOR_* */
  uint32_t chain; /* packed BJX_BIJ_* ops */
  float p0, p1;
  float lognorm; /* additive normaliser of log p, computed on the host in double:
                    Normal: -0.5*log(2pi) - log(scale); Beta: -lbeta(p0,p1);
                    Gamma: p0*log(p1) - lgamma(p0) */
} BjxCoord;

/* ---- recognised likelihood families (advi.py:90-91; README.md:91-106) --- */
#define BJX_LIK_BERNOULLI_LOGITS 0 /* tfd.Bernoulli(logits = X @ w).log_prob(y).sum() */
#define BJX_LIK_GAUSSIAN 1         /* tfd.Normal(X @ w, tau).log_prob(y).sum(), tau from noise_src */
#define BJX_LIK_BERNOULLI_PROBS 2  /* tfd.Bernoulli(probs = theta[w_offset]).log_prob(y).sum(); no X (coin toss) */

/* where the Gaussian noise scale tau comes from */
#define BJX_NOISE_CONST 0  /* tau = noise_value */
#define BJX_NOISE_KWARG 1  /* tau = exp(kwargs[noise_index])  -- trainable log_noise_scale, README.md:104 */
#define BJX_NOISE_LATENT 2 /* tau = theta[s, noise_index]     -- a constrained latent (e.g. Gamma prior + Exp) */

/* ---- precision of the streaming contraction ----------------------------- */
#define BJX_PREC_AUTO 0   /* tensor-core split path when the shape allows, else fp32 CUDA cores */
#define BJX_PREC_FP32 1   /* fp32 FFMA on CUDA cores only */
#define BJX_PREC_BF16X3 2 /* tcgen05, X/theta/dlogit each split in 3 bf16 terms, 6 products (fp32-accurate);
                             runs on the two-pass GEMM path with 128-wide tiles */
#define BJX_PREC_BF16X2 3 /* tcgen05, 2-term split (about 2^-17 relative per product) */

typedef struct BjxElboArgs {
  /* sizes */
  int64_t S;        /* n_samples (advi.py:80) */
  int64_t N;        /* rows of X / entries of y held by THIS call (local shard) */
  int64_t D;        /* size of the flat latent vector (core.py:168-176) */
  int64_t Dw;       /* columns of X = number of regression weights (0 for BERNOULLI_PROBS) */
  int64_t w_offset; /* first latent coordinate of the weights inside the flat vector */
  int64_t K;        /* number of trainable likelihood kwargs (v1 surface, README.md:86) */
  int64_t ldx;      /* row pitch of X in floats (>= Dw) */

  /* static attributes */
  int32_t vi_type;         /* BJX_VI_* */
  int32_t scale_bijector;  /* BJX_SCALE_* */
  int32_t likelihood;      /* BJX_LIK_* */
  int32_t noise_src;       /* BJX_NOISE_* (Gaussian only) */
  int32_t noise_index;     /* kwarg index or latent coordinate */
  int32_t threefry_layout; /* BJX_LAYOUT_* */
  int32_t precision;       /* BJX_PREC_* */
  int32_t kernel_override; /* 0 = automatic choice; KERNEL id + 1 forces a streaming kernel (tests / benchmarks):
                              1 fp32 tiled, 2 fp32 small-D, 3 fused tcgen05 single pass, 5 two-pass tcgen05 GEMM */
  float noise_value;   /* BJX_NOISE_CONST */
  float ll_scale;      /* full_data_size / len(outputs)   (advi.py:92; len = GLOBAL batch under sharding) */
  float prior_weight;  /* weight of the q / prior terms in THIS call's outputs: 1 on one rank, 0 on the
                          others (or 1/G everywhere) so that a sum over ranks is the reference's loss */
  float reserved1;

  /* inputs (device) */
  const uint32_t* key;     /* [2]   per-step PRNG key (advi.py:83 seed) */
  const float* loc;        /* [D]   params["posterior"].loc */
  const float* raw_scale;  /* [D] raw scale_diag  |  [D*D] row-major raw scale_tril (upper part ignored)  |
                              low rank: [D] raw scale_diag followed by [D, rank] row-major scale_perturb_factor */
  const float* kwargs;     /* [K]   or NULL */
  const float* X;          /* [N, ldx] row-major, or NULL */
  const float* y;          /* [N] */
  const BjxCoord* coords;  /* [D]   prior + bijector table */

  /* outputs (device) */
  float* out;      /* packed [1 + D + P + K]: loss, d loc, d raw_scale (P = D, D*D or D + D*rank), d kwargs.
                      This buffer is the all-reduce message under data sharding. */
  float* eps_out;  /* optional [S*D]: the standard-normal draws used (NULL to skip) */

  /* scratch (device) */
  void* workspace;
  size_t workspace_bytes; /* >= bjx_elbo_workspace_bytes(args) */

  /* optional (ABI 2): the dataset X pre-split into bf16 operand planes by bjx_prepare_x_planes, for the two-pass
   * tcgen05 GEMM path.  X is a constant of the training loop (a closure constant of the reference's jitted scan,
   * utils.py:22-31), so the split is done once per dataset instead of once per step.  NULL = split inside the step
   * (into the workspace).  The caller promises the planes were made from exactly this X / N / Dw. */
  const void* x_planes;

  /* optional (ABI 2): device-resident step counter.  When set, `key` is the base of a [n, 2] array of per-step keys
   * (jax.random.split(seed, n_epochs), utils.py:30) and the step uses key[*key_index]; see bjx_train_steps. */
  const int64_t* key_index;

  /* optional (ABI 2): Monte-Carlo SAMPLE sharding ("where N is small, shard S").  This call
   * evaluates samples [s_offset, s_offset + S) of a step with S_total samples in all: eps uses the global flat index
   * (so every shard draws exactly the noise of the unsharded step), means divide by S_total, and the per-step constant
   * of the entropy gradient (-d log sigma / d raw) is weighted by entropy_weight (1 on one shard, 0 on the others).
   * S_total = 0 (default) means no sample sharding: S_total = S, s_offset = 0, entropy_weight = prior_weight. */
  int64_t S_total;
  int64_t s_offset;
  float entropy_weight;

  /* optional (ABI 2): POINT evaluation at z = loc (no sampling, no variational term) -- the sibling inference classes on
   * the same likelihood kernels.  S must be 1; raw_scale is ignored and its gradient block is zero.
   *   BJX_POINT_LOG_JOINT : out = -(log p~(z) + ll_scale * ll(b(z)))   = -MCMC.log_joint    (mcmc.py:30-34)
   *   BJX_POINT_MAP       : out = -(log p(b(z)) + ll_scale * ll(b(z))) = MAP.neg_log_joint  (map.py:27-33; prior density
   *                         of the constrained value, no log-det-Jacobian)
   *   BJX_POINT_LOGLIK    : out = -ll_scale * ll(theta) with theta = loc ITSELF (the constrained value; no bijector, no prior)
   *                         = minus what the user-facing log_likelihood_fn(latent_sample, outputs, inputs, **kwargs) returns
   *                         (README.md:81-106), with d out / d theta in the `d loc` block and d kwargs in theirs (ABI 4)
   * and d out / d z in the `d loc` block. */
  int32_t point_mode;
  int32_t reserved3;

  /* optional (ABI 2): rank of the low-rank family (BJX_VI_LOW_RANK), 1..64 */
  int64_t rank;

  /* optional (ABI 3): row indirection for minibatches drawn on the device.  When set, row r of this call's data is
   * X[row_index[r]]: X is the FULL dataset, N the number of indexed rows (the batch size), and the kernel reads the dataset
   * through the index -- one pass over the batch's own rows, no gathered copy.  y is NOT indexed (pass y[row_index[r]],
   * e.g. BjxMinibatch.yb); x_planes must be NULL.  Only the fused tcgen05 kernel reads through an index (ask
   * bjx_elbo_supports_row_index); every other shape returns BJX_ERR_UNSUPPORTED and the caller gathers rows instead
   * (bjx_minibatch_gather).
   * ABI 4: x_planes may be set together with row_index when N_full is given -- the planes are then those of the FULL dataset
   * (bjx_prepare_x_planes(X_full, N_full, Dw)) and the TMA-fed kernel gathers the batch's plane rows with 16-byte cp.async copies (LDGSTS)
   * (any Dw in [16, 512]); X itself is then not read. */
  const int32_t* row_index;
  int64_t N_full; /* rows of the full dataset behind row_index + x_planes (0 = not given) */
} BjxElboArgs;

#define BJX_POINT_OFF 0
#define BJX_POINT_LOG_JOINT 1
#define BJX_POINT_MAP 2
#define BJX_POINT_LOGLIK 3


/* ---- library ------------------------------------------------------------ */
BJX_API int bjx_abi_version(void);
BJX_API const char* bjx_last_error(void);
BJX_API int bjx_device_sm_count(void); /* SM count of the current device, cached; negative on error */

/* ---- jax.random on the device (kernel 1) --------------------------------
 * Flat indices [offset, offset+n) of the stream of `total` draws under `key`;
 * `total` matters only to the legacy layout (the half split depends on it),
 * where offset must be 0 and n == total. */
BJX_API int bjx_random_bits(const uint32_t* key, int64_t total, int64_t offset, int64_t n, int32_t layout, uint32_t* out, void* stream);
BJX_API int bjx_random_uniform(const uint32_t* key, int64_t total, int64_t offset, int64_t n, int32_t layout, float minval,
                       float maxval, float* out, void* stream);
BJX_API int bjx_random_normal(const uint32_t* key, int64_t total, int64_t offset, int64_t n, int32_t layout, float* out, void* stream);
/* jax.random.split(key, num) -> out[num, 2]   (utils.py:30) */
BJX_API int bjx_random_split(const uint32_t* key, int64_t num, int32_t layout, uint32_t* out, void* stream);

/* ---- the hot path: loss and all gradients in one enqueue (kernels 2-6) --- */
BJX_API size_t bjx_elbo_workspace_bytes(const BjxElboArgs* args);
BJX_API int bjx_elbo_step(const BjxElboArgs* args, void* stream);
/* Same step through HOST buffers: copies key / params in, packed result out, and synchronises the stream.
 * params_host = [loc (D) | raw_scale (P) | kwargs (K)] contiguous; out_host = [1 + D + P + K].
 * args->key / loc / raw_scale / kwargs / out are ignored (staging lives in the workspace); X, y, coords stay
 * device-resident, as they are for the reference (closure constants of the jitted scan, utils.py:22-31). */
BJX_API int bjx_elbo_step_host(const BjxElboArgs* args, const uint32_t* key_host, const float* params_host, float* out_host,
                       void* stream);

/* Dataset preparation for the two-pass tcgen05 GEMM path: X [N, ldx] fp32 -> two bf16 planes (x = x1 + x2 to 2^-17
 * relative), each [N, round_up(Dw, 64)] row-major, back to back in `planes` (bjx_x_planes_bytes bytes, 16-byte aligned). */
BJX_API size_t bjx_x_planes_bytes(int64_t N, int64_t Dw);
BJX_API int bjx_prepare_x_planes(const float* X, int64_t N, int64_t Dw, int64_t ldx, void* planes, void* stream);
/* The same split chunk by chunk (ABI 4): X_rows [rows, ldx] holds rows [row0, row0 + rows) of an N_total-row dataset whose planes
 * (bjx_x_planes_bytes(N_total, Dw) bytes) are being filled -- fp32 X then never has to be resident on the device as a whole:
 * stream it through a staging buffer and keep only the planes (4 bytes per element).  A step may then pass X = NULL with
 * x_planes set wherever the planned kernel reads the planes (bjx_elbo_kernel_choice 2, or 4 outside BJX_PREC_BF16X3). */
BJX_API int bjx_prepare_x_planes_rows(const float* X_rows, int64_t rows, int64_t Dw, int64_t ldx, void* planes, int64_t N_total,
                                      int64_t row0, void* stream);

/* which streaming kernel a given problem resolves to: 0 = fp32 tiled, 1 = fp32 small-D, 2 = fused tcgen05 single pass
 * (split-bf16), 3 = coin toss, 4 = two-pass TMA-fed tcgen05 GEMM (split-bf16) */
BJX_API int bjx_elbo_kernel_choice(const BjxElboArgs* args);
/* 1 if a step of this shape can read its rows through BjxElboArgs.row_index, 0 if the rows must be gathered first.  Evaluated
 * for the route the arguments describe: x_planes == NULL asks about raw fp32 X, x_planes != NULL (any non-NULL value) with
 * N_full >= 1 about gathering plane rows. */
BJX_API int bjx_elbo_supports_row_index(const BjxElboArgs* args);

/* ---- posterior access (.apply -> Posterior.sample / log_prob, core.py:39-75) */
/* theta[S, D] = bijector(loc + scale * eps) for eps = normal(key, (S, D)); z_out optional */
BJX_API int bjx_posterior_sample(const BjxElboArgs* args, float* theta_out, float* z_out, void* stream);
/* out[i] = log q(theta_i) = q.log_prob(b^{-1}(theta_i)) + sum inverse_log_det_jacobian(theta_i)   (core.py:57-71) for n
 * constrained samples theta [n, D] in the flat latent layout; uses args->D, vi_type, scale_bijector, loc, raw_scale, coords
 * (low rank: also args->rank and a workspace of (D * rank + D + 4) floats for the Woodbury factors) */
BJX_API int bjx_posterior_log_prob(const BjxElboArgs* args, const float* theta, int64_t n, float* out, void* stream);

/* ---- packed full-rank parameterisation (north_star item 2: "L built by fill-triangular") ------------------------------
 * dense [D, D] = tfp.math.fill_triangular(packed [D(D+1)/2]) (lower; strict upper triangle = 0), and its VJP
 * d_packed = gather of d_dense over the lower triangle.  The reference itself keeps a dense D x D raw scale_tril
 * (core.py:205-209); the packed form is an extension of the host layer (ADVI(..., packed_tril=True)). */
BJX_API int bjx_fill_triangular(const float* packed, int64_t D, float* dense, void* stream);
BJX_API int bjx_fill_triangular_vjp(const float* d_dense, int64_t D, float* d_packed, void* stream);

/* ---- optimiser step fused on the device (optax.adam as used by train_fn, utils.py:26-27).  Hyper-parameters are doubles
 * (ABI 4) and rounded once inside: (1 - b) and the bias corrections 1 - b^step are formed in double. */
BJX_API int bjx_adam_update(float* params, float* m, float* v, const float* grads, int64_t n, int64_t step, double lr, double b1,
                    double b2, double eps, void* stream);

/* ---- the training loop of train_fn (utils.py:22-31) resident on the device -----------------------------------------
 * Enqueues n_steps x { ELBO + gradients at key[*steps_done] -> Adam update of params_flat in place -> losses[*steps_done]
 * = loss, ++*steps_done } with no host round trip; the sequence is CUDA-graph capturable and a captured graph can be
 * replayed as is, because the step number (bias correction, key index, loss slot) lives in `steps_done` on the device.
 * params_flat = [loc (D) | raw_scale (P) | kwargs (K)] (args->loc / raw_scale / kwargs are ignored), adam_m / adam_v the
 * optax.adam moments (zero-initialised), args->key the [n, 2] key array, args->out the packed step output (scratch),
 * losses (optional) at least as long as the total number of steps.  Single device; bjx_train_steps_sharded is the row-sharded form. */
BJX_API int bjx_train_steps(const BjxElboArgs* args, int64_t n_steps, float* params_flat, float* adam_m, float* adam_v, double lr,
                    double b1, double b2, double eps, int64_t* steps_done, float* losses, void* stream);

/* ---- on-device minibatch sampler (the scaling it feeds is advi.py:92) ----------------------
 * What a JAX user writes around loss_fn,
 *     k_batch, k_elbo = jax.random.split(key)
 *     idx = jax.random.choice(k_batch, N_full, (B,))     (replace=True: == jax.random.randint(k_batch, (B,), 0, N_full))
 *     loss_fn(params, y[idx], {"X": X[idx]}, full_data_size, seed=k_elbo, n_samples)
 * as one gather kernel: idx[b] from the key (threefry; uint32 arithmetic of jax's _randint), then row idx[b] of X (raw
 * fp32 and / or the prepared split-bf16 planes) and of y copied into contiguous batch buffers that bjx_elbo_step reads with
 * N = B and ll_scale = full_data_size / B.  Every destination is optional (NULL = skip).  elbo_key != NULL: `key` is first
 * split in two, the indices come from the first half and the second half is written to elbo_key[2]; elbo_key == NULL: the
 * indices come from `key` itself.  key_index (optional) selects key[2 * *key_index] on the device, as in BjxElboArgs. */
typedef struct BjxMinibatch {
  int64_t N_full;            /* rows of the full dataset (< 2^31) */
  int64_t B;                 /* rows per batch */
  int64_t Dw;                /* columns of X (ignored when only y is gathered) */
  int64_t ldx_full;          /* row pitch of X_full in floats */
  const float* X_full;       /* [N_full, ldx_full] or NULL */
  const float* y_full;       /* [N_full] or NULL */
  const void* x_planes_full; /* bjx_prepare_x_planes(X_full, N_full, Dw) or NULL */
  float* Xb;                 /* out [B, Dw] dense, or NULL */
  float* yb;                 /* out [B], or NULL */
  void* planes_b;            /* out bjx_x_planes_bytes(B, Dw) bytes in the layout bjx_prepare_x_planes(Xb, B, Dw) would give, or NULL */
  int32_t* idx_out;          /* out [B] row indices, or NULL */
  uint32_t* elbo_key;        /* out [2] second half of split(key), or NULL (see above) */
} BjxMinibatch;
BJX_API int bjx_minibatch_gather(const BjxMinibatch* mb, const uint32_t* key, const int64_t* key_index, int32_t layout, void* stream);

/* bjx_train_steps on minibatches: per step { bjx_minibatch_gather at key[*steps_done] -> ELBO + gradients on the batch with
 * seed mb->elbo_key -> Adam -> losses[*steps_done], ++*steps_done }.  `args` describes the BATCH problem (N = mb->B, X =
 * mb->Xb, y = mb->yb, x_planes = mb->planes_b or NULL, ll_scale = full_data_size / B); args->key is the [n, 2] array of
 * per-step keys; mb->elbo_key must be set.  Graph-capturable like bjx_train_steps. */
BJX_API int bjx_train_steps_minibatch(const BjxElboArgs* args, const BjxMinibatch* mb, int64_t n_steps, float* params_flat,
                                      float* adam_m, float* adam_v, double lr, double b1, double b2, double eps, int64_t* steps_done,
                                      float* losses, void* stream);

/* ---- the exchange step over NVLink peer memory (one kernel instead of an NCCL call) ---------------------------------
 * One-shot all-reduce (sum, in place) of the packed [loss | grads] buffer across the <= 8 GPUs of a box: every rank
 * stores its vector into a slot of every peer's buffer over NVLink, raises a flag, waits for the peers' flags and sums
 * the slots in rank order (deterministic).  Set-up, once: bjx_p2p_alloc (cudaMalloc + cudaIpcGetMemHandle), exchange the
 * 64-byte handles between the processes, bjx_p2p_open the peers'.  Per step: bjx_allreduce_packed_p2p with epoch = 1, 2,
 * 3, ... in lockstep on every rank, on the stream that ran bjx_elbo_step.  The replaced call is the psum / ncclAllReduce of
 * SURVEY.md section 8e (the reference itself is single-device). */
BJX_API size_t bjx_p2p_buffer_bytes(int64_t n, int32_t world);
BJX_API int bjx_p2p_alloc(size_t bytes, void** ptr_out, void* handle_out_64_bytes);
BJX_API int bjx_p2p_open(const void* handle_64_bytes, void** ptr_out);
BJX_API int bjx_p2p_close(void* peer_ptr);
BJX_API int bjx_p2p_free(void* own_ptr);
BJX_API int bjx_allreduce_packed_p2p(float* out, int64_t n, void* const* bufs_host, int32_t rank, int32_t world, uint64_t epoch,
                             void* stream);

/* The same exchange with the epoch read from a DEVICE step counter (epoch = epoch_base + *steps_done + 1): what a captured
 * CUDA graph of whole training steps needs (ABI 4). */
BJX_API int bjx_allreduce_packed_p2p_dev(float* out, int64_t n, void* const* bufs_host, int32_t rank, int32_t world,
                                         uint64_t epoch_base, const int64_t* steps_done, void* stream);
/* A rank waiting for a peer's flag gives up after this long (default 30 s; 0 restores it), prints the rank it waited for and
 * traps: the stream then reports a launch failure instead of hanging on a dead peer (ABI 4). */
BJX_API int bjx_p2p_set_timeout_ms(int64_t ms);

/* bjx_train_steps under ROW SHARDING (the scan of utils.py:22-31 as one device-resident executable on every
 * rank): per step { ELBO + gradients of this rank's rows at key[*steps_done] (args->prior_weight = 1 on one rank, ll_scale from the
 * GLOBAL length) -> bjx_allreduce_packed_p2p_dev of the packed [loss | grads] buffer -> Adam on the replicated params_flat ->
 * losses[*steps_done], ++*steps_done }.  Every rank computes the same reduced bits, so the replicated parameters and moments never
 * diverge.  The caller advances its own epoch counter by the number of steps enqueued (epoch_base + total steps).  ABI 4. */
BJX_API int bjx_train_steps_sharded(const BjxElboArgs* args, int64_t n_steps, float* params_flat, float* adam_m, float* adam_v,
                                    double lr, double b1, double b2, double eps, int64_t* steps_done, float* losses,
                                    void* const* p2p_bufs_host, int32_t rank, int32_t world, uint64_t epoch_base, void* stream);

/* ---- in-library timing of the dominant (streaming likelihood) kernel -------
 * bjx_profile_begin(capacity): from now on every bjx_elbo_step on this thread records a CUDA event pair around its
 * streaming-likelihood launch(es) on the caller's stream (no synchronisation); bjx_profile_collect synchronises on the
 * recorded events and returns their durations in milliseconds; bjx_profile_end releases the events. */
BJX_API int bjx_profile_begin(int32_t capacity);
BJX_API int bjx_profile_collect(float* ms_out_host, int32_t max_out, int32_t* n_out_host);
BJX_API int bjx_profile_end(void);

/* Development hook: when set to a device buffer of 8*64 int64 slots, CTA 0 of the tcgen05 kernel records clock64()
 * timestamps of its pipeline events for tiles 16..23; NULL (default) disables it. */
BJX_API int bjx_debug_set_trace_buffer(void* device_buffer_4096_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BJX_H_ */

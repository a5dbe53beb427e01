/*
 * cbo.h -- C ABI of libcbo: the B200-native full-conformal acquisition hot path.
 *
 * The reference (samuelstanton/conformal-bayesopt) is pure Python and has no FFI; these entry
 * points are what a binding for this path replaces (file:line in the reference tree):
 *
 *   cbo_conformal_acq_f64      ConformalAcquisition._conformal_fwd + forward
 *                              (conformalbo/acquisitions.py:53-76), i.e. fused
 *                              construct_conformal_bands (conformalbo/helpers.py:94-202),
 *                              conformal_gp_regression (helpers.py:218-333),
 *                              est_train_post_var (helpers.py:390-412),
 *                              RatioEstimator.forward (conformalbo/ratio_estimation.py:97-99),
 *                              _nonconformal_fwd of cUCB / cEI (acquisitions.py:151-159, 88-98)
 *                              and _conformal_integration (acquisitions.py:225-248).
 *   cbo_conf_mask_to_bounds_f64  conf_mask_to_bounds (helpers.py:15-91).
 *   cbo_ratio_forward_f64      RatioEstimator.forward (ratio_estimation.py:97-99).
 *   cbo_conformal_acq_grad_f64 the same path under autograd in optimize_acqf_sgld (conformalbo/optim/optimize.py:224-235).
 *   cbo_cehvi_integrate_f64    qConformalExpectedHypervolumeImprovement (conformalbo/mobo_acquisitions.py:24-45).
 *   cbo_cehvi_integrate_q_f64 / _grad_f64   the same for joint batches, cNEHVI (mobo_acquisitions.py:184-219), and its reverse pass.
 *   cbo_model_build_f64 / cbo_model_append_f64   the model refit before the path (conformalbo/utils.py:382-390) and the
 *                              observations a BO round adds (scripts/black_box_opt.py:154-155).
 *   cbo_ratio_train_step_f64   RatioEstimator.optimize_callback (ratio_estimation.py:110-144).
 *   cbo_sample_subset_i64      the minibatch draw of the same callback (ratio_estimation.py:129-131) on the device.
 *   cbo_sgld_update_f64        the SGLD / SGHMC step of optimize_acqf_sgld (optim/sgld.py:9-47, optim/optimize.py:224-235).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer; the library allocates no device memory, ever;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); the only host
 *     synchronisation is one per grid-refinement pass when max_grid_refinements > 0 (the
 *     reference synchronises there too, helpers.py:151-152,177).  A call with at most 32 candidate
 *     points and no grid refinement (an SGLD step) also uses two library-private side streams (probe
 *     solves; ratio estimator) that are forked from `stream` and joined back into it through
 *     events before the call returns: the caller still only sees work ordered on `stream`, and the
 *     call can be captured into a CUDA graph (the side streams become branches of it).  The side
 *     streams, their events and the kernels' shared-memory attributes are created on the FIRST call
 *     per device: make one un-captured call of the same shape before capturing;
 *   - every entry point returns a cbo_status (0 = ok); no C++ exception crosses the ABI;
 *   - NaN inputs propagate to NaN outputs (no validation on device);
 *   - re-entrant for distinct workspaces; the model arrays are read-only.
 *
 * Randomness of the reference is injected: eps_grid (grid sampler base
 * samples), z (Hutchinson probes, +-1), eps_mc (acquisition sampler base samples), u_rand
 * (uniforms behind Bernoulli.sample()).
 */
#ifndef CBO_H_
#define CBO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBO_VERSION 104

typedef enum {
    CBO_OK = 0,
    CBO_ERR_INVALID_ARG = 1,     /* null pointer / non-positive size / bad enum */
    CBO_ERR_UNSUPPORTED = 2,     /* shape outside what the kernels cover (see cbo_limits) */
    CBO_ERR_WORKSPACE = 3,       /* workspace too small or misaligned */
    CBO_ERR_CUDA = 4,            /* a CUDA runtime call / launch failed (cbo_last_cuda_error) */
    CBO_ERR_NO_DEVICE = 5        /* no sm_100 device visible */
} cbo_status;

/* CBO_ACQ_JOINT: the joint-posterior samples of f(X) over [X_baseline; X] under every refit model themselves (what cNEI turns
 * into improvements); they are integrated like cUCB samples and exported through cbo_outputs_f64.mc_values -- the per-output
 * call of cNEHVI (conformalbo/mobo_acquisitions.py:196-213). */
typedef enum { CBO_ACQ_UCB = 0, CBO_ACQ_EI = 1, CBO_ACQ_NEI = 2, CBO_ACQ_JOINT = 3 } cbo_acq_kind;

/* One output's exact GP, prepared once per model fit (all device pointers, fp64, row-major).
 * Np = cbo_padded_n(N) (N rounded up to a multiple of 128); rows/cols >= N are zero. */
typedef struct {
    int64_t N;                 /* training points */
    int64_t d;                 /* input dimension (<= 64) */
    int64_t Np;                /* padded N: leading dimension of linv / linv_t */
    const double* X;           /* N x d training inputs                              */
    const double* alpha;       /* Np    A^{-1}(y - mean), A = K + noise I            */
    const double* linv;        /* Np x Np  explicit inverse of chol(A), lower        */
    const double* linv_t;      /* Np x Np  its transpose (upper)                     */
    const double* inv_lengthscale; /* d   1 / ARD lengthscale                        */
    double mean;               /* constant prior mean                                */
    double outputscale;        /* Matern-5/2 output scale                            */
    double noise;              /* homoskedastic noise variance                       */
} cbo_gp_f64;

/* RatioEstimator target network d -> H -> ReLU -> H -> ReLU -> 1 (torch nn.Linear layout:
 * weight[out][in]).  Pass NULL for "no ratio estimator" (unit importance weights). */
typedef struct {
    int64_t hidden;            /* H <= 64 */
    const double* W1; const double* b1;   /* H x d, H */
    const double* W2; const double* b2;   /* H x H, H */
    const double* W3; const double* b3;   /* 1 x H, 1 */
} cbo_ratio_mlp_f64;

/* cNEI baseline (conformalbo/acquisitions.py:115-134), prepared once per acquisition object; nb + q <= 64 */
typedef struct {
    int64_t nb;                /* baseline points (after botorch's prune_baseline)              */
    const double* Xb;          /* nb x d                                                        */
    const double* KbX;         /* nb x Np   k(X_b, X), zero-padded columns                      */
    const double* Sbb;         /* nb x nb   noise-free posterior covariance of f(X_b) given D   */
    const double* mub;         /* nb        posterior mean of f(X_b)                            */
} cbo_nei_baseline_f64;

/* Adjoints handed in by an outer integration over several outputs (cEHVI / cNEHVI, cbo_cehvi_integrate_q_grad_f64): the
 * gradient call is then the vector-Jacobian product of this output's (conf_pred_mask, ood_mask, posterior samples at the
 * candidates under the refit models) with respect to Xq; the call's own integration is not differentiated. */
typedef struct {
    const double* mask_bar;    /* b x G x q       d L / d conf_pred_mask                                             */
    const double* ood_bar;     /* b x q           d L / d ood_mask                                                   */
    const double* f_bar;       /* b x S x G x q   d L / d samples: of the refit posterior drawn with eps_mc (S x G x q;
                                * kind UCB / EI), or of the joint samples mc_values (kind CBO_ACQ_JOINT); may be NULL */
} cbo_ext_adjoints_f64;

typedef struct {
    double alpha;              /* miscoverage level                         (acquisitions.py:32) */
    double temp;               /* sigmoid temperature; <= 0 selects the hard path (helpers.py:277,316) */
    int64_t grid_res;          /* G >= 2                                    (helpers.py:113) */
    int64_t q;                 /* joint batch size, 1..4 (q = 1 is the throughput path) */
    int64_t num_mc;            /* S, MC samples of the acquisition sampler  */
    int64_t num_probes;        /* S_p Hutchinson probes (reference: 8; <= 16) (helpers.py:390) */
    int64_t max_grid_refinements;  /*                                       (helpers.py:129) */
    int32_t kind;              /* cbo_acq_kind */
    int32_t optimistic;        /* opt_mask = [y_g >= mu]                    (acquisitions.py:57-62) */
    int32_t randomized;        /* randomised tie-break, needs u_rand        (helpers.py:306-312) */
    int32_t eps_per_candidate; /* 0: eps_grid is (G) shared (collapse_batch_dims), 1: (b x G) */
    double beta_prime;         /* cUCB: sqrt(beta*pi/2)                     */
    double best_f;             /* cEI incumbent                             */
    const cbo_nei_baseline_f64* nei;   /* kind == CBO_ACQ_NEI / CBO_ACQ_JOINT only (eps_mc is then S x G x (nb+q)), else NULL */
    /* max_grid_refinements > 0 and the candidate batch is sharded over several ranks: called on the HOST once per
     * refinement pass with stats = {min accepted, max accepted, sum of (accepted/G - 1/2)^2, candidates} of this rank's
     * shard; must overwrite them with the values over all ranks (min / max / sum / sum).  NULL: single batch. */
    void (*refine_reduce)(double* stats, void* ctx);
    void* refine_ctx;
    /* Initial row scales of the grid covariance, b x q (device) or NULL (= 1): the covar_scale products a caller-driven
     * refinement loop has accumulated (helpers.py:166-182) -- multi-output bands take the loop's decisions over ALL outputs,
     * so the t per-output calls of one step run with max_grid_refinements = 0 and the scales of the step. */
    const double* grid_cov_scale;
    /* cbo_conformal_acq_grad_f64 only: NULL = differentiate the call's own acquisition value. */
    const cbo_ext_adjoints_f64* ext_adjoints;
    /* cbo_conformal_acq_grad_f64 only, joint batches (q >= 2) without grid refinement (every kind; with ext_adjoints per output).  0: the target grid
     * is a constant of the reverse pass (what the reference's torch.no_grad() block, helpers.py:118-119, intends).  1: add
     * the gradient the reference's autograd actually propagates: the prior block k(x_j, x_k) of its lazily evaluated
     * predictive covariance is evaluated when the grid is sampled, outside the no_grad block, so the grid depends on X
     * through the candidates' mutual prior covariances (reproduces autograd through the reference's own classes,
     * tests/golden/ref_path_grad.pt). */
    int32_t grid_grad_leak;
} cbo_params;

/* Any pointer may be NULL (that output is skipped) except acq. */
typedef struct {
    double* acq;               /* b           acquisition value (max over q, mean over S)        */
    double* target_grid;       /* b x G x q   y_g                                                */
    double* grid_logp;         /* b x G       log N(y_g; mu, c S)  (joint over q)                */
    double* conf_pred_mask;    /* b x G x q   conformal (soft) mask                              */
    double* ood_mask;          /* b x q       sigma((omega_c - alpha)/temp)                      */
    double* q_conf_scores;     /* b x G x q   the candidates' own conformal scores               */
    double* post_mean;         /* b x q       mu(x)                                              */
    double* post_var;          /* q = 1: b, noisy predictive variance s(x); q > 1: b x q x q covariance S */
    double* num_accepted;      /* b x q       #{g : mask >= 0.5}                                 */
    double* refit_mean;        /* b x G       q = 1: posterior mean at x under the model refit on (x, y_g)  */
    double* refit_sd;          /* b           q = 1: its (noise-free) posterior sd                          */
    double* mc_values;         /* b x S x G x q   kind NEI / JOINT: the per-sample values the integration consumed  */
} cbo_outputs_f64;

int cbo_version(void);
/* Kernels this library has launched since it was loaded (all entry points, all threads): what bench.py reports as
 * gpu_launches for its timed region. */
int64_t cbo_launch_count(void);
const char* cbo_status_str(int status);
const char* cbo_last_cuda_error(void);
int64_t cbo_padded_n(int64_t N);

/* Bytes of workspace cbo_conformal_acq_f64 needs for b candidates (chunked internally when
 * max_grid_refinements == 0; `chunk` <= 0 selects the default chunk).  256-byte aligned. */
size_t cbo_workspace_bytes(int64_t N, int64_t d, int64_t b, const cbo_params* params, int64_t chunk);

/* The hot path.  Xq: b x q x d candidates; eps_grid: G x q (or b x G x q); z: (N+q) x S_p of +-1;
 * eps_mc: S x G x q; u_rand: b x G x q or NULL; chunk: candidates per internal pass (<= 0: default). */
int cbo_conformal_acq_f64(const cbo_gp_f64* gp, const cbo_ratio_mlp_f64* ratio, const cbo_params* params,
                          const double* Xq, int64_t b, const double* eps_grid, const double* z,
                          const double* eps_mc, const double* u_rand, const cbo_outputs_f64* out,
                          void* workspace, size_t workspace_bytes, int64_t chunk, void* stream);

/* Value and gradient (cUCB / cEI / cNEI, q <= 4): additionally writes dacq_dX (b x q x d) = d acq[c] / d Xq[c, :, :], what the reference
 * obtains by autograd through the same path inside optimize_acqf_sgld (conformalbo/optim/optimize.py:224-235:
 * acq_function(X).sum().backward()); the target grid, its log-density and the optimistic mask are constants, as they are
 * built under torch.no_grad() there (helpers.py:118-119, acquisitions.py:58-59).  temp <= 0: the masks are step functions
 * (zero gradient through them), the value path still differentiates.  cNEI (and CBO_ACQ_JOINT with external adjoints)
 * add the reverse pass of the joint posterior over [X_baseline; X] (Cholesky reverse of the (nb+q)^2 factor).  With
 * params->ext_adjoints the result is the vector-Jacobian product described there.  Workspace: cbo_workspace_bytes_grad. */
size_t cbo_workspace_bytes_grad(int64_t N, int64_t d, int64_t b, const cbo_params* params, int64_t chunk);
int cbo_conformal_acq_grad_f64(const cbo_gp_f64* gp, const cbo_ratio_mlp_f64* ratio, const cbo_params* params,
                               const double* Xq, int64_t b, const double* eps_grid, const double* z,
                               const double* eps_mc, const double* u_rand, const cbo_outputs_f64* out,
                               double* dacq_dX, void* workspace, size_t workspace_bytes, int64_t chunk, void* stream);

/* Same call with HOST candidate / result buffers (pinned for full speed), pipelined per internal chunk: the candidates of
 * chunk k+1 go host -> device and the values of chunk k-1 device -> host on a library-private copy stream (one per
 * device, created on first use) while chunk k computes on `stream`; both streams are drained before the call returns. */
int cbo_conformal_acq_host_f64(const cbo_gp_f64* gp, const cbo_ratio_mlp_f64* ratio, const cbo_params* params,
                               const double* Xq_host, int64_t b, const double* eps_grid, const double* z,
                               const double* eps_mc, const double* u_rand, double* acq_host,
                               double* Xq_staging /* b x q x d device */, double* acq_staging /* b device */,
                               void* workspace, size_t workspace_bytes, int64_t chunk, void* stream);

/* cEHVI (conformalbo/mobo_acquisitions.py:24-45 + acquisitions.py:225-248, q = 1): combines the per-output results of m
 * cbo_conformal_acq_f64 calls (arrays of m device pointers, each b x G or b) with the box decomposition
 * (cell_lower / cell_upper: n_cells x m, as botorch's partitioning.get_hypercell_bounds()) into one value per candidate.
 * eps_mc: S x G x m base samples of the acquisition sampler.  m <= 4. */
int cbo_cehvi_integrate_f64(int64_t m, int64_t b, int64_t G, int64_t S, const double* const* mask,
                            const double* const* logp, const double* const* grid, const double* const* mu,
                            const double* const* ood, const double* const* refit_mean, const double* const* refit_sd,
                            const double* eps_mc, const double* cell_lower, const double* cell_upper, int64_t n_cells,
                            double alpha, int32_t optimistic, double* acq, void* stream);

/* cEHVI for joint batches q <= 4 (incl. pending points) and cNEHVI (mobo_acquisitions.py:24-45, 184-219): per output k the
 * results of one cbo_conformal_acq_f64 call with q points per candidate batch -- mask / grid: b x G x q, logp: b x G, mu / ood:
 * b x q, cov: b x q x q (cbo_outputs_f64.post_var; b for q = 1), noise[k] (HOST array of m values).  The refit posterior at the
 * candidates is sampled with eps_mc (S x G x q x m) -- or, when fx != NULL, fx[k] (b x S x G x q, cbo_outputs_f64.mc_values of
 * a CBO_ACQ_JOINT call) are the samples (cNEHVI: joint with the baseline).  cells_per_sample != 0: cell_lower / cell_upper are
 * S x n_cells x m, one box decomposition per MC sample (cNEHVI), else n_cells x m.  Masks, optimistic masks and ood masks
 * multiply over outputs and over the q points, log-densities add; values are botorch's _compute_qehvi (inclusion-exclusion
 * over the q-subsets).  values_out (b x G, may be NULL) receives the MC-averaged hyper-volume improvements. */
int cbo_cehvi_integrate_q_f64(int64_t m, int64_t b, int64_t G, int64_t S, int64_t q, const double* const* mask,
                              const double* const* logp, const double* const* grid, const double* const* mu,
                              const double* const* ood, const double* const* cov, const double* noise_host,
                              const double* const* fx, const double* eps_mc, const double* cell_lower,
                              const double* cell_upper, int64_t n_cells, int32_t cells_per_sample, double alpha,
                              int32_t optimistic, double* acq, double* values_out, void* stream);

/* The same integration with its reverse pass: additionally writes, per output k, d acq[c] / d mask (b x G x q), d acq[c] / d ood
 * (b x q) and d acq[c] / d samples (b x S x G x q: of the internally sampled refit posterior, or of fx[k]) -- the
 * cbo_ext_adjoints_f64 of the m per-output cbo_conformal_acq_grad_f64 calls whose dacq_dX sum to d acq / d X (what autograd does
 * through mobo_acquisitions.py:24-45 inside optimize_acqf_sgld).  Grid, log-densities, optimistic masks and cells are constants.
 * values_out (b x G) is required (scratch between the two passes); f_bar[k] may be NULL. */
int cbo_cehvi_integrate_q_grad_f64(int64_t m, int64_t b, int64_t G, int64_t S, int64_t q, const double* const* mask,
                                   const double* const* logp, const double* const* grid, const double* const* mu,
                                   const double* const* ood, const double* const* cov, const double* noise_host,
                                   const double* const* fx, const double* eps_mc, const double* cell_lower,
                                   const double* cell_upper, int64_t n_cells, int32_t cells_per_sample, double alpha,
                                   int32_t optimistic, double* acq, double* values_out, double* const* mask_bar,
                                   double* const* ood_bar, double* const* f_bar, void* stream);

/* conf_mask_to_bounds: grid, mask: b x G x t (q = 1 squeezed) -> lb, ub: b x t.  G <= 2048. */
int cbo_conf_mask_to_bounds_f64(const double* target_grid, const double* conf_pred_mask, int64_t b, int64_t G,
                                int64_t t, double* lb, double* ub, void* stream);

/* RatioEstimator.forward on n points (n x d) -> n density ratios (unclamped below, as the reference). */
int cbo_ratio_forward_f64(const cbo_ratio_mlp_f64* ratio, const double* x, int64_t n, int64_t d, double* out,
                          void* stream);

/* Append q observations to a built model state in O(q N^2) (rank-1 extensions of L^-1 and alpha) instead of the O(N^3)
 * rebuild -- the BO loop adds q queries per round (scripts/black_box_opt.py:154-155) and fantasises pending points.
 * In place: X must have room for N + q rows, alpha / linv / linv_t are padded to Np >= N + q (a multiple of 128, the
 * leading dimension) with zero padding as cbo_model_build_f64 leaves it.  x_new: q x d, y_new: q (device).  Same
 * hyper-parameters as the build.  info as in cbo_model_build_f64.  3 launches per point, no host sync. */
size_t cbo_model_append_workspace_bytes(int64_t Np);
int cbo_model_append_f64(double* X, int64_t N, int64_t d, int64_t Np, const double* inv_lengthscale, double mean,
                         double outputscale, double noise, double* alpha, double* linv, double* linv_t,
                         const double* x_new, const double* y_new, int64_t q, int32_t* info, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Density-ratio classifier, ONE training step on the device (conformalbo/ratio_estimation.py:133-144, the body of
 * RatioEstimator.optimize_callback once the minibatch is chosen): mean BCE-with-logits over [x_neg (n_neg x d, label 0);
 * x_pos[pos_idx] (n_pos rows, label 1; pos_idx NULL = the first n_pos rows)], one torch.optim.Adam step on the classifier
 * (weight_decay is L2 added to the gradient; step is 1-based), then update_target_network (:101-107):
 * target <- (1 - ema_weight) target + ema_weight theta.  theta, adam_m, adam_v, target: flat vectors of
 * hidden*d + hidden + hidden^2 + hidden + hidden + 1 doubles in torch's parameter order (W1, b1, W2, b2, W3, b3); the
 * cbo_ratio_mlp_f64 of the acquisition calls may point straight into `target`.  hidden must be 64, d <= 64.
 * loss (device scalar, may be NULL) receives the minibatch loss.  Two launches, fixed-order reductions, no host sync. */
size_t cbo_ratio_train_workspace_bytes(int64_t n_rows, int64_t d);
int cbo_ratio_train_step_f64(int64_t d, int64_t hidden, double* theta, double* adam_m, double* adam_v, double* target,
                             const double* x_neg, int64_t n_neg, const double* x_pos, const int64_t* pos_idx,
                             int64_t n_pos, int64_t step, double lr, double beta1, double beta2, double eps,
                             double weight_decay, double ema_weight, double* loss, void* workspace,
                             size_t workspace_bytes, void* stream);

/* The positive minibatch of RatioEstimator.optimize_callback drawn on the device: k distinct indices out of [0, n), the first
 * k values of a keyed pseudo-random permutation of [0, n) (what conformalbo/ratio_estimation.py:129-131 takes from
 * np.random.permutation(num_positive)[:num_negative], without building the permutation and without the host).  The same
 * (n, k, seed) gives the same indices; out_idx (k int64) feeds cbo_ratio_train_step_f64's pos_idx.  n <= 2^31.  One launch. */
int cbo_sample_subset_i64(int64_t n, int64_t k, uint64_t seed, int64_t* out_idx, void* stream);

/* One SGLD / SGHMC update of the candidates inside optimize_acqf_sgld (conformalbo/optim/optimize.py:224-235 with
 * conformalbo/optim/sgld.py:9-47): gradient of  -sum(acq) + penalty_scale * (sigmoid((lb - X)/w).mean() +
 * sigmoid((X - ub)/w).mean())  from dacq_dX, then  buf <- (1 - momentum) buf - lr g + sqrt(2 lr momentum T) noise,
 * X <- X + buf  (momentum == 0: X <- X - lr g + sqrt(2 lr T) noise).  X, momentum_buf, dacq_dX, noise: n_rows x d; lb, ub: d.
 * The reference uses penalty_scale = 1e2, w = 1e-2, momentum = .1, T = temperature / sqrt(d). */
int cbo_sgld_update_f64(double* X, double* momentum_buf, const double* dacq_dX, const double* noise, const double* lb,
                        const double* ub, int64_t n_rows, int64_t d, double lr, double momentum, double temperature,
                        double penalty_scale, double penalty_width, void* stream);

/* Model-state builder (the step before the hot path; conformalbo/utils.py:382-390 refits SingleTaskGP and GPyTorch caches
 * the same factors): raw training data + fixed hyper-parameters -> alpha (Np), linv and linv_t (Np x Np) of cbo_gp_f64.
 * Blocked Cholesky and triangular inverse on the fp64 tensor path.  info (device int, may be NULL) receives 0, or
 * 1 + the first row with a non-positive pivot (A not positive definite; outputs are then NaN).  No host sync. */
size_t cbo_model_workspace_bytes(int64_t N, int64_t d);
int cbo_model_build_f64(const double* X, const double* y, int64_t N, int64_t d, const double* inv_lengthscale, double mean,
                        double outputscale, double noise, double* alpha, double* linv, double* linv_t, int32_t* info,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Building blocks exported for tests and profiling (same kernels the hot path launches). */
/* C[m, n] = sum_k A[m, k] * B[n, k] over the K-range implied by tri: 0 full, 1 lower (k < n-tile end),
 * 2 upper (k >= n-tile begin).  A: M x K (lda), B: Nn x K (ldb), C: M x Nn (ldc); K, Nn multiples of 128.
 * rownorm_partial (M x Nn/128) receives per-tile sums of squares of C when non-NULL. */
int cbo_gemm_tn_f64(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                    int64_t M, int64_t Nn, int64_t K, int32_t tri, double* rownorm_partial, void* stream);
/* U[c, i] = k(x_c, X_i) (Np-padded rows, zeros beyond N); mu[c] = mean + u.alpha.  scratch: d x Np doubles. */
int cbo_kernel_rows_f64(const cbo_gp_f64* gp, const double* Xq, int64_t n, double* U, double* mu, double* scratch,
                        void* stream);

/* Optional per-kernel-class timing with CUDA events on the caller's stream (bench.py's roofline leg).
 * Classes: 0 per-call tables, 1 kernel rows (+candidate ratio MLP), 2 GEMM pass 1, 3 GEMM pass 2, 4 score,
 * 5 score value+reverse, 6 reverse GEMM passes, 7 kernel-row reverse (+ratio MLP reverse). */
void cbo_timing_enable(int on);
int cbo_timing_collect(double* ms_out /* 8 */, int64_t* launches_out /* 8 */);

#ifdef __cplusplus
}
#endif
#endif /* CBO_H_ */

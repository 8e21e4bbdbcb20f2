/* =============================================================================
 * bpcuda.h — C ABI of libbpcuda: the B200 (sm_100a) engine behind bpinference's
 * posterior log-density path.
 *
 * The reference (jproney/bpinference, paths below are under /root/reference) has no
 * FFI of its own: its engine boundary is "R objects handed to rstan".  Every entry
 * point here replaces one of those hand-offs and cites it.  Nothing in this header
 * mentions R, Python or torch: plain pointers, sizes and opaque handles.  The R
 * `.Call` shim (r-package/src/bpcuda_rcall.c) and the Python ctypes mirror
 * (bpinference_b200/_ffi.py) are thin translations of these signatures.
 *
 * Conventions
 *   - every function returns 0 on success, a BPC_E_* code otherwise; the message of
 *     the last failure on the calling thread is bpc_last_error().
 *   - input buffers are borrowed for the duration of the call only; outputs are
 *     written into caller-allocated arrays; handles are freed with bpc_*_free.
 *   - matrices handed over in the Stan data list keep R's column-major layout;
 *     index vectors are 1-based exactly as create_stan_data produces them
 *     (R/stan_utils.R:57-62).
 *   - per-chain arrays are chain-major: upars[c*dim + k].
 *   - pointers named *_dev are device pointers on the handle's device; all others
 *     are host pointers.
 *   - handles are NOT thread-safe and a data handle has ONE owner at a time: its device workspace is
 *     shared by every entry point that takes it (bpc_log_prob*, bpc_log_lik, bpc_moments, a sampler
 *     created on it).  Sequential use from one thread is safe in any order, also across streams and
 *     chain counts (a sampler notices that its workspace was re-tiled and rebuilds its CUDA graph);
 *     concurrent use from two threads, or two samplers advancing on one data handle, is not.
 * ============================================================================= */
#ifndef BPCUDA_H
#define BPCUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define BPC_ABI_VERSION 2   /* 2: bpc_sample_args.init_opt_steps and .metric; pool buffers carry a trailing "rank finished" flag */

/* error codes */
enum {
  BPC_OK = 0,
  BPC_E_INVALID = 1,   /* bad argument / failed validation (message mirrors the reference's stop() text) */
  BPC_E_EXPR = 2,      /* rate expression rejected (R/stan_utils.R:205-364) */
  BPC_E_PRIOR = 3,     /* prior rejected (R/stan_utils.R:372-459) */
  BPC_E_COMPILE = 4,   /* NVRTC failed; bpc_last_error() holds the compiler log */
  BPC_E_CUDA = 5,      /* CUDA runtime error (includes "no device": there is no CPU fallback) */
  BPC_E_UNSUPPORTED = 6
};

/* per-chain status bits written by bpc_log_prob / the sampler (a non-zero status means the
 * chain's lp is -inf and its gradient 0, i.e. what makes Stan reject a proposal) */
enum {
  BPC_ST_OK = 0,
  BPC_ST_NOT_PD = 1,            /* multi_normal: Sigma not positive definite (multitype_template.stan:152) */
  BPC_ST_RATE_NONFINITE = 2,    /* a generated rate expression produced NaN/Inf */
  BPC_ST_ONETYPE_SINGULAR = 4,  /* b == d or variance <= 0 in one_type_template.stan:43 */
  BPC_ST_PRIOR_SUPPORT = 8      /* Theta_k outside the support of its prior (e.g. uniform(0,1) with bounds (0,2)) */
};

/* which Stan template the model stands for — the `simple_bd` / `noise_model` switches of
 * generate(), R/stan_utils.R:130,167-187 */
typedef enum {
  BPC_MULTITYPE = 0,        /* stanfiles/multitype_template.stan */
  BPC_MULTITYPE_NOISE = 1,  /* stanfiles/multitype_noise_template.stan */
  BPC_SIMPLE_BD = 2         /* stanfiles/one_type_template.stan */
} bpc_template_kind;

/* one prior_dist object: list(name=, params=, bounds=)  (R/stan_utils.R:467-487).
 * has_bounds == 0 stands for bounds = NA. */
typedef struct {
  const char* name;
  int nparams;
  double params[3];
  int has_bounds;
  double lower, upper;
} bpc_prior;

/* a bp_model object (R/bp_model.R:12-17) plus the priors list and the template switches that
 * generate() receives (R/stan_utils.R:130). */
typedef struct {
  int ntypes;                     /* ncol(e_mat) */
  int nevents;                    /* nrow(e_mat) */
  int nparams;                    /* model$nparams */
  int ndep;                       /* model$ndep (0 allowed; the data list then carries the dummy ndep = 1) */
  const double* e_mat;            /* nevents x ntypes, column-major */
  const int* p_vec;               /* nevents, 1-based parent types */
  const char* const* func_deps;   /* nevents strings as written by the user / deparse(model$func_deps[[e]]) */
  const bpc_prior* priors;        /* nparams entries */
  int kind;                       /* bpc_template_kind */
} bpc_model_desc;

/* the Stan data list.  Multitype templates: the fields of multitype_template.stan:64-80 as built at
 * R/stan_utils.R:57-58.  One-type template: one_type_template.stan:2-12 as built at R/stan_utils.R:61-62
 * (there `times` has one entry per observation and times_idx / ntimes_unique are ignored). */
typedef struct {
  int ndatapts;
  int ntimes_unique;
  int ndep_levels;
  int ndep;                    /* columns of function_var (>= 1: dummy column when the model has ndep == 0) */
  const double* pop_vec;       /* ndatapts x ntypes column-major (vector of length ndatapts for SIMPLE_BD) */
  const double* init_pop;      /* same shape */
  const double* times;         /* ntimes_unique durations (ndatapts durations for SIMPLE_BD) */
  const int* times_idx;        /* ndatapts, 1-based (NULL for SIMPLE_BD) */
  const double* function_var;  /* ndep_levels x ndep column-major */
  const int* var_idx;          /* ndatapts, 1-based */
} bpc_stan_data;

typedef struct bpc_model bpc_model;
typedef struct bpc_data bpc_data;
typedef struct bpc_sampler bpc_sampler;

/* ---- library ------------------------------------------------------------------------------- */
int bpc_abi_version(void);
const char* bpc_last_error(void);
/* number of visible CUDA devices (0 without a GPU; never an error) */
int bpc_device_count(void);
/* measured FP64 FMA throughput of `device` in TFLOP/s (2 flop per DFMA); a register-resident DFMA loop
 * run for `ms` milliseconds.  This is the roofline denominator SURVEY.md §8(d) asks the builder to measure. */
int bpc_fp64_peak(int device, double ms, double* tflops, double* sm_clock_mhz);
/* the same for the fp64 tensor path (mma.sync m8n8k4 f64 = DMMA, which shares the FP64 datapath with DFMA on this part): the
 * denominator for the kernels whose flops are mostly DMMA (bp_fusedo, bp_fusedc, bp_fusedw, bp_toep) */
int bpc_fp64_tensor_peak(int device, double ms, double* tflops);

/* ---- expression front end: replaces exp_to_stan / check_valid (R/stan_utils.R:205-364) ------ */
/* validates one rate expression; on failure returns BPC_E_EXPR and bpc_last_error() is the text of the
 * reference's stop() (tests/testthat/test_expression_parsing.R:4-47). */
int bpc_check_expr(const char* expr, int max_params, int max_dep);
/* writes what exp_to_stan returns (tests/testthat/test_expression_parsing.R:50-54) */
int bpc_expr_to_stan(const char* expr, int max_params, int max_dep, char* out, int out_cap);
/* writes R's deparse() of the parsed expression (tests/testthat/test_bpmodel.R:68-69) */
int bpc_expr_deparse(const char* expr, char* out, int out_cap);
/* validates a prior_dist (R/stan_utils.R:418-459) */
int bpc_check_prior(const bpc_prior* prior);

/* ---- model: replaces rstan::stan_model(model_code = generate(model, priors, ...)) ------------ */
/* (examples/two_type.R:18,25).  Parses func_deps, emits the CUDA translation unit (rate functions, their
 * reverse-mode derivatives, the generator assembly, and the kernels specialised for this model), compiles
 * it with NVRTC for sm_100a.  Works without a GPU (compilation only); the module is loaded on first use. */
int bpc_model_create(const bpc_model_desc* desc, bpc_model** out);
void bpc_model_free(bpc_model* m);
/* number of unconstrained dimensions: nparams, +1 (sigma_obs, last) for the noise template
 * (multitype_noise_template.stan:99-104) */
int bpc_model_dim(const bpc_model* m);
/* the generated CUDA source / the compiled cubin (for inspection, caching and cuobjdump) */
const char* bpc_model_cuda_source(const bpc_model* m);
int bpc_model_cubin(const bpc_model* m, const void** data, unsigned long long* size);
/* warnings the front end produced (e.g. Stan integer division of two integer literals) or "" */
const char* bpc_model_warnings(const bpc_model* m);

/* ---- data: replaces the `data = dat` argument of rstan::sampling (examples/two_type.R:26) ---- */
/* copies the host arrays, validates index ranges, repacks to the device layout.
 * grouping: 0 = auto, 1 = force the per-observation fused kernel, 2 = force the grouped two-phase path. */
int bpc_data_create(const bpc_model* m, const bpc_stan_data* d, int device, int grouping, bpc_data** out);
void bpc_data_free(bpc_data* d);
int bpc_data_ngroups(const bpc_data* d);    /* distinct (env level, duration) pairs */
int bpc_data_is_grouped(const bpc_data* d); /* which path log_prob takes */
/* the kernels bpc_log_prob launches for `nchains` chains: *path = 0 bp_fused (term ring), 1 bp_grouped, 2 bp_fusedw + bp_wpost
 * (W path, bp_fused as sub-step fallback), 3 bp_onetype, 4 bp_ctable + bp_fusedc + bp_wpost (centred W path: sorted durations
 * dense enough for short series around shared centres; bp_fused as sub-step fallback), 5 bp_otable + bp_fusedo + bp_toep +
 * bp_wpost (octet path: the centred series with eight chains per tensor tile, from 16 chains); *ntiles = CTAs per chain (octet path:
 * per octet of chains); *scratch_bytes = per-launch partial-sum traffic */
int bpc_data_path(bpc_data* d, int nchains, int* path, int* ntiles, unsigned long long* scratch_bytes);

/* ---- evaluate: replaces rstan::log_prob / rstan::grad_log_prob -------------------------------- */
/* upars: nchains x dim unconstrained vectors (Theta1..ThetaP, then sigma_obs); jacobian = rstan's
 * adjust_transform.  lp[nchains]; grad[nchains*dim] or NULL; status[nchains] or NULL.  Host buffers:
 * the call copies in, launches, copies out and synchronises. */
int bpc_log_prob(bpc_model* m, bpc_data* d, const double* upars, int nchains, int jacobian,
                 double* lp, double* grad, int* status);
/* same with device-resident buffers; asynchronous on `stream` (a cudaStream_t; NULL is CUDA's default stream). */
int bpc_log_prob_dev(bpc_model* m, bpc_data* d, const double* upars_dev, int nchains, int jacobian,
                     double* lp_dev, double* grad_dev, int* status_dev, void* stream);
/* per-observation log_lik with normalising constants: the generated-quantities LOO blocks
 * (stanfiles/multitype_loo_block.stan:31, simple_birth_death_loo_block.stan:11).
 * log_lik: nchains x ndatapts, observation order = the order of the data list.  As in the reference, the noise template gets
 * the same block (R/stan_utils.R:189-191): its log_lik has no sigma_obs term.  NaN where the density does not exist. */
int bpc_log_lik(bpc_model* m, bpc_data* d, const double* upars, int nchains, double* log_lik);
/* mean vector and covariance matrix of the final population of every (chain, observation): what
 * calculate_moments returns (R/calculate_moments.R:11-88: mu_mat, sigma_mat) for the rates the chain's parameters imply.
 * mu: nchains x ndatapts x ntypes; sigma: nchains x ndatapts x ntypes x ntypes (row-major), data-list order. */
int bpc_moments(bpc_model* m, bpc_data* d, const double* upars, int nchains, double* mu, double* sigma);
/* the `transformed parameters` of the templates (multitype_template.stan:105-117, one_type_template.stan:19-27) for `ndraws`
 * CONSTRAINED parameter vectors theta [ndraws x dim]: rates[(draw * ndep_levels + v) * nevents + e] = f_e(Theta, function_var[v, ]),
 * evaluated on the device by the model's generated rate functions.  r_mat and theta[i] of the stanfit are these numbers laid
 * out by the caller (bpinference_b200/fit.py, r-package/R/bpcuda.R): r_mat[e, p_vec[e]] of the LAST level, theta[i] =
 * c(e_mat, r_mat of level i). */
int bpc_transformed_parameters(bpc_model* m, bpc_data* d, const double* theta, int ndraws, double* rates);
int bpc_data_nlevels(const bpc_data* d);   /* ndep_levels of the data list */
/* constrained values: Theta [nchains x nparams] (+ sigma_obs) from unconstrained upars, and back */
int bpc_constrain(const bpc_model* m, const double* upars, int nchains, double* theta);
int bpc_unconstrain(const bpc_model* m, const double* theta, int nchains, double* upars);
/* kernel launches issued by this thread's handles since the last reset (bench.py's gpu_launches) */
long long bpc_launch_count(int reset);
/* device time in ms of the dominant likelihood kernel accumulated over launches since the last reset,
 * measured with CUDA events on the launching stream (only when enabled; bench.py's roofline.achieved) */
int bpc_kernel_timing(bpc_data* d, int enable);
int bpc_kernel_time_ms(bpc_data* d, int reset, double* total_ms, long long* launches);
/* algorithmic operation counts of the shipped kernels (FMA = 2): per application of L in the forward series, per
 * application of L^T in the reverse sweep, and per observation outside the series.  No GPU needed. */
int bpc_flop_model(const bpc_model* m, double* per_app_fwd, double* per_app_bwd, double* per_obs);
/* the constants of the centred / octet paths: per term of the short series (coefficient, X row, forward product, one row of the S'
 * update) and per observation outside the terms on bp_fusedc and on bp_fusedo */
int bpc_flop_model_centred(const bpc_model* m, double* per_term, double* per_obs_centred, double* per_obs_octet);
/* exact algorithmic flop count of one evaluation of the likelihood for chain 0 of `upars`, counted by
 * the shipped kernels themselves in a counting build (SURVEY.md §8(d) convention: FMA = 2) */
int bpc_count_flops(bpc_model* m, bpc_data* d, const double* upars, int nchains, double* flops_per_chain);

/* ---- sample: replaces rstan::sampling(stan_mod, data, chains, iter, warmup, init, control) ---- */
/* (examples/two_type.R:26; sampler settings of examples/<script>.R).  Device-resident NUTS (multinomial, diagonal
 * metric, Stan's windowed warm-up and dual averaging): every chain is a state machine advanced by one leapfrog
 * step per pass; one pass = one fused log-density+gradient evaluation of ALL chains + one bookkeeping kernel.
 * Random numbers are Philox4x32-10 keyed by (seed, global chain id), so a chain's draws do not depend on how
 * chains are sharded over GPUs. */
typedef struct {
  int chains;            /* chains on THIS device / rank */
  int iter;              /* total iterations per chain, warm-up included (rstan's iter) */
  int warmup;            /* rstan's warmup */
  double adapt_delta;    /* control$adapt_delta */
  int max_treedepth;     /* control$max_treedepth (<= 12): a trajectory has at most 2^max_treedepth - 1 leapfrog steps */
  unsigned long long seed;
  int chain_id_offset;   /* global id of this rank's first chain */
  const double* inits;   /* chains x nparams CONSTRAINED values (uniform_initialize, R/stan_utils.R:74-87; sigma_obs starts at 0.05);
                            NULL = uniform(-2,2) on the unconstrained scale as rstan does */
  int pooled_adaptation; /* 0: every chain adapts its own metric (rstan).  1: the metric of each warm-up window is estimated from
                            the draws of all chains of all ranks — the one collective on the path (see bpc_sampler_pool_*) */
  double stepsize;       /* initial step size (0 = 1.0 as rstan) */
  int init_opt_steps;    /* 0: NUTS starts at the initial values (rstan).  > 0: every chain first climbs from its initial value with at
                            most this many steps of a monotone Adam ascent on the unconstrained log density, using the same fused
                            log-density + gradient passes (for data whose posterior is orders of magnitude narrower than the range the
                            initial values are drawn from: uniform_initialize(0, 1) against a posterior sd of 1e-3 with 1e5 observations) */
  int metric;            /* 0: diag_e, rstan's default (per-chain variances, or pooled ones with pooled_adaptation).  1: dense_e -- the full
                            covariance of the unconstrained parameters, estimated at every warm-up window from the draws of ALL chains of
                            all ranks (needs pooled_adaptation = 1: a single chain's window is far too short for a covariance matrix).  The
                            sampler then works in the whitened coordinates q = mu + L w, Sigma = L L^T, which is HMC with M^-1 = Sigma */
  bpc_data* init_opt_data; /* NULL, or (with init_opt_steps > 0 and pooled_adaptation) a data handle holding a SUBSAMPLE of the observations, on the
                            same device: the ascent then has two stages -- on the subsample until it converges (far from the mode every evaluation
                            needs sub-steps and costs many times one near it, so the long way is travelled at a fraction of the cost), then, after
                            an exchange in which all chains wait for each other, on the full data from there.  The handle stays the caller's */
} bpc_sample_args;

int bpc_sampler_create(bpc_model* m, bpc_data* d, const bpc_sample_args* a, bpc_sampler** out);
void bpc_sampler_free(bpc_sampler* s);
/* advances all chains by at most max_passes passes.  *state: 0 = more passes needed, 1 = all chains finished,
 * 2 = (pooled adaptation) all chains reached the end of a warm-up window and wait for the pooled metric. */
int bpc_sampler_run(bpc_sampler* s, long long max_passes, int* state);
/* pooled adaptation, state 2: bpc_sampler_pool_stats writes this rank's sums (diagonal metric: count, sum q_k, sum q_k^2;
 * bpc_sampler_pool_len doubles in all) into the caller's DEVICE buffer stats_dev; the caller adds the buffers of all ranks in place (ncclAllReduce /
 * torch.distributed.all_reduce) and hands the result to bpc_sampler_pool_apply.  With count >= bpc_sampler_pool_len the entry after the sums
 * is 1.0 when every chain of this rank has finished (or failed) and 0.0 otherwise: its sum over ranks equals the number of ranks
 * exactly when the whole run is over.  The protocol that cannot deadlock: every rank calls stats + all-reduce each time its
 * bpc_sampler_run returns (state 1 or 2), applies the result when its state was 2, and leaves when the flag sum says everybody is
 * done -- a finished rank keeps contributing zeros.  On one rank both pointers may be NULL: an internal buffer is used. */
int bpc_sampler_pool_stats(bpc_sampler* s, double* stats_dev, int count);
/* doubles of the pool buffer including the trailing flag: 2 + 2*dim (diagonal metric) or 2 + dim + dim*(dim+1)/2 (dense metric:
 * count, sum dq, lower triangle of sum dq dq^T with dq = q - centre of the metric in force, flag) */
int bpc_sampler_pool_len(const bpc_sampler* s);
int bpc_sampler_pool_apply(bpc_sampler* s, const double* reduced_dev);
/* results (host copies).  draws: chains x (iter - warmup) x dim CONSTRAINED values (Theta1..P, sigma_obs);
 * lp: lp__ ; per-draw sampler parameters as rstan::get_sampler_params reports them.  Any pointer may be NULL. */
int bpc_sampler_draws(bpc_sampler* s, double* draws, double* lp, int* divergent, int* treedepth,
                      int* n_leapfrog, double* accept_stat, double* stepsize);
/* sums for cross-chain convergence diagnostics: for every dimension (count, sum of means, sum of squared means, sum of
 * variances) over the HALF chains of the saved draws (split R-hat, as rstan's summary reports it), 4*dim doubles, written to
 * the caller's device buffer (to be added over ranks) and/or to a host buffer; either may be NULL.  Then bpc_rhat. */
int bpc_sampler_chain_moments(bpc_sampler* s, double* stats_dev, int count, double* stats_host);
/* split potential scale reduction (the statistic check_rhat reads, R/stan_utils.R:516-533) from the reduced sums */
int bpc_rhat(const double* reduced_moments, int dim, int draws_per_chain, double* rhat);
long long bpc_sampler_total_leapfrogs(bpc_sampler* s);
long long bpc_sampler_passes(const bpc_sampler* s);
/* where every chain is (what rstan's `refresh` prints): iterations completed, current step size, lp__ of the current point and the
 * state-machine phase (0 initialising, 1-2 step-size search, 3 sampling, 4 waiting at a pooled window, 5 finished, 6 failed,
 * 7 optimising its initial point).  Host arrays of `chains` entries; any pointer may be NULL. */
int bpc_sampler_progress(bpc_sampler* s, int* iter, double* stepsize, double* lp, int* phase);
/* convenience: whole run on one device, single rank */
int bpc_sample(bpc_model* m, bpc_data* d, const bpc_sample_args* a, double* draws, double* lp,
               int* divergent, int* treedepth, int* n_leapfrog, double* accept_stat, double* stepsize);

/* the same run on SEVERAL devices of this process (SURVEY.md §8(b) threading, §8(e)): `a->chains` chains on each of the `ndev`
 * devices listed (global chain ids a->chain_id_offset + device_index * chains + c), the data list and the module replicated, one
 * host thread per device, and one single-process NCCL communicator (ncclCommInitAll; NCCL is bound at run time) for the two
 * exchanges of the path: the pooled warm-up statistics (when a->pooled_adaptation) and the half-chain moments behind split R-hat.
 * Replaces `options(mc.cores = parallel::detectCores())` + `rstan::sampling(..., chains =)` (examples/two_type.R:20,26) for a
 * single-process caller such as R.  a->inits, when given, holds ndev * chains rows.  Outputs hold ndev * chains chains, device-major,
 * laid out as for bpc_sample; rhat [dim] (split R-hat over all chains) and info [4] (passes of the slowest device, leapfrogs of all
 * chains, exchanges over NCCL, NCCL version code) may be NULL. */
int bpc_sample_multi(bpc_model* m, const bpc_stan_data* data, const int* devices, int ndev, const bpc_sample_args* a, double* draws,
                     double* lp, int* divergent, int* treedepth, int* n_leapfrog, double* accept_stat, double* stepsize, double* rhat,
                     long long* info);
/* version code of the NCCL library bpc_sample_multi binds (e.g. 22703), 0 when none can be loaded */
int bpc_nccl_version(void);

/* ---- simulate: replaces bp() / bpsims() (R/simulation.R:9-56, 69-131) -------------------------- */
/* Gillespie simulation of `nrep` replicates on the device (Philox); pops: nrep x ntimes x ntypes. */
int bpc_simulate(int ntypes, int nevents, const double* e_mat, const int* p_vec, const double* r_vec,
                 const double* z0, const double* times, int ntimes, int nrep, unsigned long long seed,
                 int device, double* pops);

#ifdef __cplusplus
}
#endif
#endif /* BPCUDA_H */

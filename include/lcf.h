/*
 * lcf.h -- C ABI of liblcf_b200.so, the B200 (sm_100a) implementation of the
 * lightcurve_fitting MCMC hot path.
 *
 * The reference (griffin-h/lightcurve_fitting) is pure Python and has no FFI; the
 * boundary it exposes for this path is a set of Python call signatures.  Each entry
 * point below names the reference interface it replaces (file:line relative to
 * /root/reference/lightcurve_fitting/).  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain C types only: opaque handles, int/int64/double scalars, caller-owned host
 *     buffers (C-contiguous).  No torch / numpy types cross this boundary.
 *   - every function returns 0 on success, <0 on error; lcf_last_error() returns a
 *     thread-local message.
 *   - handles are not thread-safe; each handle owns one CUDA stream.
 *   - there is NO CPU fallback: every compute entry point fails with LCF_ERR_CUDA when
 *     no sm_100-class device is usable.
 */
#ifndef LCF_H
#define LCF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCF_ABI_VERSION 1

/* error codes */
#define LCF_OK            0
#define LCF_ERR_ARG      -1   /* invalid argument                                     */
#define LCF_ERR_CUDA     -2   /* CUDA runtime / no device                             */
#define LCF_ERR_NAN      -3   /* log-posterior evaluated to NaN (emcee: ValueError)   */
#define LCF_ERR_STATE    -4   /* call sequence error (e.g. run before set_state)      */
#define LCF_ERR_NWALKERS -5   /* nwalkers < 2*ndim (emcee RedBlueMove: RuntimeError)  */

/* model ids: models.py classes */
#define LCF_MODEL_SHOCKCOOLING       1   /* models.py:301-353  */
#define LCF_MODEL_SHOCKCOOLING2      2   /* models.py:356-411  */
#define LCF_MODEL_SHOCKCOOLING3      3   /* models.py:433-496  */
#define LCF_MODEL_SHOCKCOOLING4      4   /* models.py:507-632  */
#define LCF_MODEL_COMPANIONSHOCKING  5   /* models.py:848-918  */
#define LCF_MODEL_COMPANIONSHOCKING2 6   /* models.py:921-980  */
#define LCF_MODEL_COMPANIONSHOCKING3 7   /* models.py:983-1045 */
#define LCF_MODEL_BLACKBODY_SED      8   /* bolometric.py:154-164 with spectrum=planck_fast */

/* prior kinds: models.py:1048-1098 */
#define LCF_PRIOR_UNIFORM    0
#define LCF_PRIOR_LOGUNIFORM 1
#define LCF_PRIOR_GAUSSIAN   2

#define LCF_PRECISION_FP64 0   /* parity mode: rtol 1e-9 vs the reference arithmetic */
#define LCF_PRECISION_FP32 1   /* throughput mode: rtol 1e-4                          */

#define LCF_MAX_NDIM 12
#define LCF_NMODEL_CONSTS 16

/* per-filter role bits (CompanionShocking*, models.py:811-815, 913-916) */
#define LCF_ROLE_KASEN_RU  1   /* filt.char == 'U': shock component times r_U         */
#define LCF_ROLE_SIFTO_RR  2   /* filt.char == 'r': SiFTO times r_r                   */
#define LCF_ROLE_SIFTO_RI  4   /* filt.char == 'i': SiFTO times r_i                   */
#define LCF_ROLE_DT_U      8   /* filter is filtdict['U']: SiFTO epoch offset dt_U    */
#define LCF_ROLE_DT_I     16   /* filter is filtdict['i']: SiFTO epoch offset dt_i    */

typedef struct lcf_problem  lcf_problem;    /* one light curve / SED epoch + model + priors   */
typedef struct lcf_ensemble lcf_ensemble;   /* walker ensemble resident in HBM                */
typedef struct lcf_batch    lcf_batch;      /* many independent (problem, ensemble) pairs     */

/*
 * Problem description.  All arrays are host pointers, copied by lcf_problem_create.
 *
 * Filter bank (replaces Filter.trans / Filter.synthesize, filters.py:181-214, 288-310):
 *   for filter f the samples k in [bank_offsets[f], bank_offsets[f+1]) carry
 *     bank_alpha[k] = c1 * nu_k * (1+z)                      [kK]   (exp argument = alpha / T)
 *     bank_w[k]     = c2 * nu'_k^3 * min(1, nu_c/nu'_k) * T_norm_per_freq_k * trapz_weight_k
 *                     (times 10^(-0.4*ebv*kappa_k) for a fixed ebv)
 *     bank_kappa[k] = A_F99(lambda'_k; a_v = 3.1) per unit E(B-V)  (only read by ShockCooling3)
 *   so that  synthesize(planck_fast, T, R) == R^2 * sum_k bank_w[k] / (exp(bank_alpha[k]/T) - 1).
 *
 * Points (replaces lc['MJD'], lc['filter'], lc[q], lc['d'+q]; models.py:116-119):
 *   must be grouped by filter: point_filter[] non-decreasing.
 */
typedef struct {
    int32_t model_id;
    int32_t precision;
    int32_t ndim;             /* model parameters (+1 when use_sigma)                         */
    int32_t use_sigma;        /* models.py:128-133                                            */
    int32_t sigma_type;       /* 0 = 'relative' (sigma_units = dy), 1 = 'absolute' (median dy) */
    int32_t npoints;
    int32_t nfilters;
    int32_t reserved0;
    double  sigma_unit_abs;   /* np.median(dy) when sigma_type == 1                           */
    double  model_consts[LCF_NMODEL_CONSTS];
                              /* BaseShockCooling: A, a, alpha, epsilon_1, epsilon_2, L_0, T_0,
                                 Tph_to_Tcol (models.py:195-224); others: unused               */
    const int32_t *bank_offsets;  /* [nfilters+1] */
    const double  *bank_alpha;    /* [bank_offsets[nfilters]] */
    const double  *bank_w;
    const double  *bank_kappa;    /* may be NULL unless model_id == SHOCKCOOLING3 */
    const int32_t *filter_role;   /* [nfilters] LCF_ROLE_* bits; may be NULL */
    /* SiFTO cubic splines (models.py:701-717): uniform knots x0 + i*dx, i in [0, n_knots);
       sifto_coef[f][i][0..3] = scipy PPoly coefficients (highest power first) of interval i */
    int32_t       sifto_nknots;
    int32_t       reserved1;
    double        sifto_x0, sifto_dx;
    const double *sifto_coef;     /* [nfilters][sifto_nknots-1][4]; NULL unless CompanionShocking* */
    const double  *t;             /* [npoints] */
    const int32_t *point_filter;  /* [npoints] index into the bank */
    const double  *y;             /* [npoints] */
    const double  *dy;            /* [npoints] */
    /* priors (models.py:1048-1098) */
    const int32_t *prior_kind;    /* [ndim] */
    const double  *prior_min;     /* [ndim] strict bounds */
    const double  *prior_max;
    const double  *prior_mean;    /* Gaussian only */
    const double  *prior_std;
} lcf_problem_desc;

int         lcf_abi_version(void);
const char *lcf_last_error(void);
int         lcf_device_count(void);          /* number of usable CUDA devices (0 on a CPU box) */
int         lcf_set_device(int device);

int  lcf_problem_create(const lcf_problem_desc *desc, lcf_problem **out);
void lcf_problem_destroy(lcf_problem *p);

/* Model.__call__(t, f, *params) pointwise (models.py:86-91, 1161-1162): nsets parameter
   vectors of length n_model_params -> out[nsets][npoints].                                 */
int lcf_model_eval(lcf_problem *p, int64_t nsets, const double *params, double *out);

/* Model.log_likelihood(lc, p, use_sigma, sigma_type) (models.py:93-136): params[nsets][ndim] */
int lcf_log_likelihood(lcf_problem *p, int64_t nsets, const double *params, double *out);

/* log_posterior closure (fitting.py:121-128, bolometric.py:154-164).  *nan_count receives the
   number of NaN results (the reference's emcee raises ValueError on any).                    */
int lcf_log_posterior(lcf_problem *p, int64_t nsets, const double *params, double *out, int64_t *nan_count);

/* emcee.EnsembleSampler(nwalkers, ndim, log_posterior) with the default StretchMove(a=2)
   (fitting.py:130, bolometric.py:167).  rank/world shard one ensemble across GPUs: this
   process proposes/evaluates/accepts only its slice of each half-ensemble (world=1: all). */
int  lcf_ensemble_create(lcf_problem *p, int64_t nwalkers, uint64_t seed, int rank, int world,
                         lcf_ensemble **out);
void lcf_ensemble_destroy(lcf_ensemble *e);
/* sampler.run_mcmc(initial, ...) first evaluates log_prob(initial): coords[nwalkers][ndim];
   log_prob may be NULL (computed on device).  Returns LCF_ERR_NAN like emcee.              */
int  lcf_ensemble_set_state(lcf_ensemble *e, const double *coords, const double *log_prob);
/* Shared ensemble (world > 1): this rank's own walkers only, logical [first, first + count) = the range
   lcf_ensemble_device_view reports (2 * own_begin, 2 * own_count); [count][ndim].  Uploads them, evaluates their
   log-posteriors and, with the fused exchange attached, stores rows + log-probabilities into every peer's replica over NVLink.
   Callers put one barrier between this call (on all ranks) and the first step.  Same checks as lcf_ensemble_set_state. */
int  lcf_ensemble_set_state_slice(lcf_ensemble *e, int64_t first, int64_t count, const double *coords);
int  lcf_ensemble_get_state(lcf_ensemble *e, double *coords, double *log_prob);
/* sampler.reset(): forget the stored chain and acceptance counts, keep the position.       */
int  lcf_ensemble_reset(lcf_ensemble *e);
/* nsteps stretch-move iterations with device counter-based RNG (Philox4x32-10 keyed by
   (seed, iteration, half, walker)); store != 0 appends to the HBM-resident chain.          */
int  lcf_ensemble_run(lcf_ensemble *e, int64_t nsteps, int store);
/* run(store = 1) that also streams every finished step into host buffers (ideally page-locked) on a copy stream,
   overlapping the device-to-host transfer of the chain with the sampling of the next steps.               */
int  lcf_ensemble_run_to_host(lcf_ensemble *e, int64_t nsteps, double *chain_host /* [nsteps][nwalkers][ndim] */,
                              double *log_prob_host /* [nsteps][nwalkers] */);
/* the same for the walkers [first, first + count) only (a rank's own walkers of a shared ensemble)        */
int  lcf_ensemble_run_to_host_slice(lcf_ensemble *e, int64_t nsteps, int64_t first, int64_t count,
                                    double *chain_host /* [nsteps][count][ndim] */, double *log_prob_host /* [nsteps][count] */);
/* the same iterations driven by caller-supplied draws, in emcee's order:
   split[s][w] in {0,1}; for each step the Ns0 entries for split 0 (ascending walker index)
   then the Ns1 entries for split 1: z (stretch factors), partner (index into the
   complementary set, ascending walker order), logu (log of the acceptance uniform).        */
int  lcf_ensemble_run_replay(lcf_ensemble *e, int64_t nsteps, int store, const int32_t *split,
                             const double *z, const int32_t *partner, const double *logu);
/* pre-allocate HBM for nsteps more stored iterations (otherwise grown geometrically on demand) */
int  lcf_ensemble_reserve(lcf_ensemble *e, int64_t nsteps);
/* one half-step only (multi-GPU drivers interleave the all-gather between half-steps)      */
int  lcf_ensemble_half_step(lcf_ensemble *e, int half, int store);
int  lcf_ensemble_end_step(lcf_ensemble *e, int store);
int64_t lcf_ensemble_nstored(lcf_ensemble *e);
int  lcf_ensemble_get_chain(lcf_ensemble *e, double *chain /* [nstored][nwalkers][ndim] */);
int  lcf_ensemble_get_log_prob(lcf_ensemble *e, double *log_prob /* [nstored][nwalkers] */);
/* the stored chain of walkers [first, first+count) only: chain[nstored][count][ndim], log_prob[nstored][count]
   (either may be NULL).  A rank of a sharded ensemble owns the contiguous walkers [2*own_begin, 2*(own_begin+own_count)). */
int  lcf_ensemble_get_chain_slice(lcf_ensemble *e, int64_t first, int64_t count, double *chain, double *log_prob);
int  lcf_ensemble_get_accepted(lcf_ensemble *e, int64_t *accepted /* [nwalkers] */);
/* Convergence diagnostics computed on the device-resident chain, steps [discard, nstored) (no reference equivalent;
   definitions: emcee.autocorr.integrated_time (c = 5 there) and the split Gelman-Rubin statistic).  tau/window/rhat are
   [ndim] and may be NULL; max_lag <= 0 means all lags; window[d] = -1 when no automatic window exists in range.      */
int  lcf_ensemble_diagnostics(lcf_ensemble *e, int64_t discard, double c, int64_t max_lag, double *tau, int64_t *window,
                              double *rhat);
/* device-side view for the multi-GPU exchange (torch.distributed wraps these raw pointers):
   coords are stored colour-major: rows [0, n0) even walkers, [n0, nwalkers) odd walkers.    */
int  lcf_ensemble_device_view(lcf_ensemble *e, void **d_coords, void **d_log_prob, void **stream,
                              int64_t *n0, int64_t *own_begin /* [2] */, int64_t *own_count /* [2] */);
/* launch on a caller-owned CUDA stream (cudaStream_t as void*), e.g. torch's current stream, so that
   collectives issued by the caller are stream-ordered with the half-step kernels.           */
int  lcf_ensemble_set_stream(lcf_ensemble *e, void *stream);

/* Fused multi-GPU exchange of a shared ensemble: every rank holds a full replica of the walker
   positions; once the peers' replicas are attached, the accept epilogue of the half-step kernel stores accepted
   walkers straight into them over NVLink and the ranks order their half-steps through device-side flags -- no NCCL
   call, no host round trip, lcf_ensemble_run() runs whole chains in lock-step.  Same node only (cudaIpc).
   export: three 64-byte cudaIpc handles (coords, log_prob, flags) to send to the other ranks;
   attach_ipc: the handles of ALL ranks, [world][3][64] bytes (entry `rank` is ignored);
   attach_ptrs: the same with raw device pointers valid in this process (arrays of length world).          */
int  lcf_ensemble_ipc_export(lcf_ensemble *e, unsigned char *handles /* [3][64] */);
int  lcf_ensemble_peers_attach_ipc(lcf_ensemble *e, const unsigned char *handles /* [world][3][64] */);
int  lcf_ensemble_peers_attach_ptrs(lcf_ensemble *e, void *const *d_coords, void *const *d_log_prob, void *const *d_flags);
int  lcf_ensemble_peers_detach(lcf_ensemble *e);
int  lcf_ensemble_exchange_view(lcf_ensemble *e, void **d_flags, int *npeers);
int  lcf_ensemble_sync(lcf_ensemble *e);
/* device time (ms) spent in the last lcf_ensemble_run* call, measured with CUDA events on the
   handle's stream; kernel launches issued by that call.                                    */
int  lcf_ensemble_last_timing(lcf_ensemble *e, double *ms, int64_t *launches);

/* Batched independent ensembles: one CTA per (problem, ensemble); the whole burn-in +
   sampling chain runs inside ONE launch (calculate_bolometric's per-epoch loop
   bolometric.py:735-798; survey-scale light-curve batches).  All problems must share
   model_id / precision / ndim / use_sigma.                                                 */
int  lcf_batch_create(int64_t nproblems, lcf_problem *const *problems, int64_t nwalkers, uint64_t seed,
                      lcf_batch **out);
void lcf_batch_destroy(lcf_batch *b);
int  lcf_batch_set_state(lcf_batch *b, const double *coords /* [nproblems][nwalkers][ndim] */);
/* run nburn unstored + nsteps stored iterations for every problem                           */
int  lcf_batch_run(lcf_batch *b, int64_t nburn, int64_t nsteps);
int  lcf_batch_get_chain(lcf_batch *b, double *chain /* [nproblems][nsteps][nwalkers][ndim] */);
int  lcf_batch_get_log_prob(lcf_batch *b, double *log_prob /* [nproblems][nsteps][nwalkers] */);
int  lcf_batch_get_accepted(lcf_batch *b, int64_t *accepted /* [nproblems][nwalkers] */);
int  lcf_batch_get_status(lcf_batch *b, int32_t *status /* [nproblems]: 0 ok, LCF_ERR_NAN */);
int  lcf_batch_last_timing(lcf_batch *b, double *ms, int64_t *launches);

/* calculate_bolometric at survey scale (bolometric.py:735-798).
   lcf_sed_batch_create: every SED epoch of a table as one batch, from flat arrays: epoch e owns points [offsets[e], offsets[e+1]),
   point_filter indexes the shared packed bank (bank_offsets[nfilters+1], bank_alpha, bank_w as in lcf_problem_desc); model =
   LCF_MODEL_BLACKBODY_SED (the spectrum_mcmc closure, bolometric.py:154-164), priors / use_sigma / sigma_type common to all
   epochs.  The batch owns the per-epoch problems; their device arrays share one allocation and one host-to-device copy.
   lcf_batch_summary: after lcf_batch_run, per epoch the median and the +- perc_contained/2 percentile distances
   (median_and_unc, bolometric.py:456-480) of T, R, L_bol = stefan_boltzmann(T, R) (:422-447) and L_pseudo = pseudo(T, R)
   (:32-59; `comb` = a one-point LCF_MODEL_BLACKBODY_SED problem whose single filter is the 1-THz comb) over every stored
   sample, computed on the HBM-resident chain: out[nproblems][4][3] = (median, median - lower, upper - median).                  */
int  lcf_sed_batch_create(int64_t nepochs, const int32_t *offsets, const int32_t *point_filter, const double *y, const double *dy,
                          int32_t nfilters, const int32_t *bank_offsets, const double *bank_alpha, const double *bank_w, int32_t ndim,
                          int32_t use_sigma, int32_t sigma_type, const int32_t *prior_kind, const double *prior_min,
                          const double *prior_max, const double *prior_mean, const double *prior_std, int32_t precision,
                          int64_t nwalkers, uint64_t seed, lcf_batch **out);
int  lcf_batch_summary(lcf_batch *b, lcf_problem *comb, double sigma_sb, double perc_contained, double *out);

/* Batched box-bounded least-squares fits of planck_fast(nu; T, R) to SEDs, one thread per epoch (replaces the per-epoch
   scipy.optimize.curve_fit of bolometric.py:483-531).  offsets[nepochs+1] delimit each epoch's points in nu [THz, already
   multiplied by (1+z)] and lum; c1, c2 are the constants of models.py:1101-1102; p0/lower/upper are (T, R).
   popt [nepochs][2], pcov [nepochs][2][2] (curve_fit's (J^T J)^-1 RSS/(n-2); inf when n <= 2), status 0 = converged.   */
int  lcf_blackbody_lstsq_batch(int64_t nepochs, const int32_t *offsets, const double *nu, const double *lum, double c1, double c2,
                               double cutoff_freq, const double *p0, const double *lower, const double *upper, double *popt,
                               double *pcov, int32_t *status);

/* The launch the library last issued for this problem (tests assert which kernel a workload ran on): walkers per CTA,
   warps per CTA, cluster size, grid size in CTAs, and the kernel instantiation: 0 = generic k_pass<MODEL, real, -1, false>,
   1 = k_pass<MODEL, real, 5, false> (32 walkers per CTA at compile time), 2 = k_pass<MODEL, real, 5, true> (32 walkers, no
   intrinsic-scatter / model-grid branches), 3 = k_ring (persistent cooperative kernel for one small ensemble),
   4 = k_ring with look-ahead rounds (one grid barrier per step: both outcomes of the partner's move are evaluated),
   5 = k_pass_seg (a filter bank larger than shared memory, streamed through it in segments of consecutive filters; also forced,
   for tests, by LCF_SEG_SAMPLES=<samples per segment> in the environment when the problem's first launch is shaped).      */
int  lcf_problem_last_launch(lcf_problem *p, int *walkers_per_cta, int *warps_per_cta, int *cluster_size, int64_t *grid,
                             int *kernel_variant);

/* launch-shape overrides for tuning (0 = heuristic): walkers per CTA (power of two <= 32)
   and warps per CTA.                                                                       */
int  lcf_set_tuning(int walkers_per_cta, int warps_per_cta);
/* same, plus the thread-block-cluster size (power of two <= 8): the CTAs of a cluster share one
   walker group and split its light curve between them (small ensembles on many SMs).          */
int  lcf_set_tuning_ex(int walkers_per_cta, int warps_per_cta, int cluster_size);
/* split-K override: the transmission samples of a (walker, point pair) are swept by `sample_chunks` lanes (power of two, walkers per
   CTA x chunks <= 32) whose partial sums are combined with warp shuffles -- small ensembles / SED epochs; 0 = heuristic.        */
int  lcf_set_tuning_split(int sample_chunks);
/* flat split of a half-step over the co-resident CTA slots of the device: the tile rows of every walker group are dealt to 8 units,
   each CTA of a `slots`-CTA grid takes the same number of (group, unit) pairs and a group shared by several CTAs is finished by the
   one that delivers its last units -- no partial last wave, no idle slots when there are fewer groups than slots.  The chi-square
   sums keep one fixed order (units, then warps) whichever CTA computed what, so chains do not depend on the split.
   -1 = the launch-shape cost model decides (default), 0 = never, 1 = whenever the shape allows it.                              */
int  lcf_set_tuning_flat(int mode);
/* lcf_problem_last_launch plus: walker groups of the launch, units of the structured sums (1: plain per-warp sums) and photometry
   points per lane and tile (4 in the 32-walker plain FP32 kernels of ShockCooling 1-3, else 2); a launch with
   grid != groups x cluster_size was a flat split.                                                                               */
int  lcf_problem_last_launch_ex(lcf_problem *p, int64_t *groups, int *sum_units, int *points_per_lane);

/* Plan of a segmented launch (kernel variant 5; host only, no device needed).  A filter bank larger than shared memory is cut into
   runs of consecutive filters of at most `cap_samples` transmission samples (two per pair record): greedy, each run as long as
   fits.  segs_out[4 i .. 4 i + 3] = (first filter, end filter, first pair record, pair records) of run i; returns the number of
   runs, or -1 when one filter alone exceeds `cap_samples` / `max_segs` runs are not enough.  The reference has no counterpart:
   Filter.synthesize (filters.py:288-310) integrates one transmission curve at a time from host memory.                       */
int  lcf_plan_bank_segments(const int *pair_records, int nfilters, int64_t cap_samples, int *segs_out, int max_segs);

#ifdef __cplusplus
}
#endif
#endif /* LCF_H */

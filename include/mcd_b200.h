/*
 * mcd_b200.h -- C ABI of the B200-native mcmc-dynamics likelihood path.
 *
 * One shared library (libmcd_b200.so, built from mcmc_dynamics_b200/csrc/ by
 * __graft_entry__.build()) replaces the per-walker Python likelihood of the
 * reference with one batched CUDA launch per ensemble call.  Plain pointers and
 * sizes only: no torch, numpy or C++ types cross this boundary.  Every entry
 * point names the reference interface it stands in for (paths relative to
 * /root/reference/mcmc_dynamics/).
 *
 * Conventions
 *   - return value 0 = success, negative = error; mcd_last_error() then gives a
 *     message (thread-local).  No C++ exception crosses the ABI.
 *   - a handle owns all of its device memory.  Host columns passed to
 *     mcd_pack_create() are copied; the caller keeps ownership.
 *   - a handle is thread-compatible: use it from one host thread at a time.  All launches through
 *     one handle share its reduction scratch, so they must be ordered: issue them on one stream, or
 *     synchronise between streams (the host-buffer entry points run on the handle's own stream and
 *     return synchronised).
 *   - *_device entry points are asynchronous on the given CUDA stream
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream); results
 *     are valid after that stream is synchronised.  Entry points without the
 *     suffix take HOST buffers and return when the result is in `out`.
 *   - theta is row-major float64 [n_walkers][n_theta]: exactly the array emcee
 *     hands to a vectorised log_prob_fn (analysis/runner.py:403, 162-175).
 */
#ifndef MCD_B200_H
#define MCD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCD_ABI_VERSION 3

/* Rotation / dispersion model.
 *   CONSTANT: analysis/constant.py:52-111  (ConstantFit.rotation_model / dispersion_model)
 *   RADIAL  : analysis/model.py:93-180     (ModelFit: Lynden-Bell rotation, Plummer dispersion) */
enum { MCD_ROT_CONSTANT = 0, MCD_ROT_RADIAL = 1 };

/* Background treatment.
 *   NONE         : analysis/runner.py:264-271
 *   FIXED_PMEMBER: analysis/runner.py:272-286 (lnlike_background[N], pmember[N] columns)
 *   FIXED_DENSITY: analysis/model.py:586-623  (ModelFitConstantBackground: lnlike_background[N],
 *                  m = density/(density+f_back))
 *   GAUSSIAN     : analysis/constant.py:326-364, analysis/model.py:414-456 (fitted v_back, sigma_back, f_back) */
enum { MCD_BG_NONE = 0, MCD_BG_FIXED_PMEMBER = 1, MCD_BG_FIXED_DENSITY = 2, MCD_BG_GAUSSIAN = 3 };

/* Model parameter slots: union of MODEL_PARAMETERS of the five classes
 * (constant.py:19,256; model.py:71,354,526). */
enum {
    MCD_P_V_SYS = 0, MCD_P_SIGMA_MAX, MCD_P_V_MAXX, MCD_P_V_MAXY, MCD_P_RA_CENTER, MCD_P_DEC_CENTER,
    MCD_P_A, MCD_P_R_PEAK, MCD_P_V_BACK, MCD_P_SIGMA_BACK, MCD_P_F_BACK, MCD_NPARAM
};

#define MCD_MAX_THETA 16       /* free parameters per walker (the largest shipped config has 11) */

/* Arithmetic variant of the kernels.
 *   FAST : division-free restructuring, MUFU-seeded Newton reciprocals (quadratic steps, 1.3e-12
 *          per term), logs taken of running products, table-driven exponentials in the mixture
 *          variants (6.5e-12 per term, zero mean); FP64 throughout (default).  lnprob agrees with
 *          the reference's formulas to ~2e-14 relative on the BASELINE configurations (tolerance of
 *          the path: 1e-9).  Range: the combined variance verr^2 + sigma^2 of every star must lie
 *          within 2^+-250 (1e-75 .. 1e75 km^2/s^2) -- the kernels multiply four of them before
 *          folding the exponent; beyond that the result is NaN (never a wrong finite number).
 *   PLAIN: the same per-star formulas with the compiler's div / sqrt / log / exp, no range limit
 *          beyond the reference's own; kept to A/B the restructuring on the device and for the
 *          per-star entry points. */
enum { MCD_MATH_FAST = 0, MCD_MATH_PLAIN = 1 };

/* What mcd_pack_create() compiles: a model, its parameter routing and the star columns.
 * Stands in for the per-call work of Runner.fetch_parameter_values (analysis/runner.py:143-180),
 * the inspect-based kwargs routing (constant.py:140-147, model.py:208-215) and the astropy unit
 * conversions of SURVEY.md section 3.3, all hoisted out of the sampling loop. */
typedef struct mcd_pack_desc {
    int32_t rotation;                 /* MCD_ROT_*                                              */
    int32_t background;               /* MCD_BG_*                                               */
    int32_t n_theta;                  /* P: number of free parameters = columns of theta        */
    int32_t math_mode;                /* MCD_MATH_*                                             */
    int64_t n_stars;                  /* N (of this shard)                                      */
    const double *ra;                 /* [N] degrees  (runner.py:75-81)                         */
    const double *dec;                /* [N] degrees                                            */
    const double *v;                  /* [N] km/s                                               */
    const double *verr;               /* [N] km/s                                               */
    const double *pmember;            /* [N] or NULL  (runner.py:103)                           */
    const double *density;            /* [N] or NULL  (constant.py:257, model.py:355,527)       */
    const double *lnlike_background;  /* [N] or NULL  (runner.py:102, model.py:563)             */
    int32_t slot[MCD_NPARAM];         /* column of theta that feeds the parameter, or -1: fixed */
    double fixed_value[MCD_NPARAM];   /* current value (own unit): used when slot < 0; for a    */
                                      /* sampled ra_center it is the expansion point ra0 of the */
                                      /* free-centre geometry (any value is valid)              */
    double unit_scale[MCD_NPARAM];    /* own unit -> km/s | deg | arcmin | 1                    */
    double lower[MCD_MAX_THETA];      /* box prior of theta column j (parameter.py:691-692)     */
    double upper[MCD_MAX_THETA];
    int32_t fixed_prior_ok;           /* 0: a fixed parameter violates its bounds => all -inf   */
    int32_t device;                   /* CUDA device ordinal                                    */
    int64_t n_stars_total;            /* N of the whole catalogue when star-sharded, else 0     */
    /* Segments: n_segments > 1 turns the handle into a batch of independent problems that share the
     * model and the parameter routing -- the per-radial-bin fits of bin/run.py:179-190 and
     * bin/run_tests.py:81-97, one ConstantFit per bin.  Stars [segment_offsets[s],
     * segment_offsets[s+1]) of the columns belong to segment s; theta is then
     * [n_segments][n_walkers][n_theta], results are [n_segments][n_walkers], and `n_walkers` in
     * every call is the number of walkers PER segment.  Available for MCD_BG_NONE and
     * MCD_BG_FIXED_PMEMBER (ConstantFit(data_i, parameters, background=background), bin/run.py:186). */
    int32_t n_segments;               /* 0 or 1: one problem                                    */
    const int64_t *segment_offsets;   /* [n_segments + 1] host array, or NULL                   */
} mcd_pack_desc;

typedef struct mcd_handle mcd_handle;

typedef struct mcd_info {
    int64_t n_stars;
    int32_t n_theta;
    int32_t n_columns;            /* device columns read per star by the lnlike kernel          */
    int32_t bytes_per_star;       /* = 8 * n_columns: algorithmic HBM bytes per star per launch */
    int32_t flops_per_term;       /* nominal FP64 operations per (walker, star) term*/
    int32_t free_centre;          /* 1 if ra_center or dec_center is sampled                    */
    int32_t sm_count;
    int32_t last_grid_x, last_grid_y, last_block;   /* geometry of the most recent launch       */
    int32_t last_walker_tile;
    int64_t launches;             /* kernels launched through this handle so far                */
    int32_t n_segments;
} mcd_info;

int mcd_abi_version(void);
const char *mcd_last_error(void);

/* Model construction: Runner.__init__ + ConstantFit/ModelFit.__init__ (analysis/runner.py:40-106,
 * constant.py:24-50, model.py:76-91) as far as device state is concerned. */
int mcd_pack_create(const mcd_pack_desc *desc, mcd_handle **out);
/* Re-route parameters (fixed <-> free, new fixed values, bounds, math mode, background mode) on
 * the star columns already resident on the device: what `parameters[name].set(...)` between
 * construction and the run amounts to (bin/run_tests.py:88-93,137-148).  Column pointers and
 * n_stars of `desc` are ignored / must match. */
int mcd_pack_reconfigure(mcd_handle *h, const mcd_pack_desc *desc);
void mcd_destroy(mcd_handle *h);
int mcd_get_info(const mcd_handle *h, mcd_info *info);

/* lnlike for a whole (half-)ensemble: <Model>.lnlike (constant.py:113-154,293-324;
 * model.py:182-223,391-456,565-623) called once per walker by emcee in the reference. */
int mcd_lnlike(mcd_handle *h, const double *theta_host, int32_t n_walkers, double *out_host);
int mcd_lnlike_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev, void *stream);

/* lnprob = box prior + lnlike with exact -inf for rejected walkers: Runner.lnprob
 * (analysis/runner.py:288-306) with Runner.lnprior (runner.py:182-217) fused in.
 * Host-buffer entry points block until the result is in out_host.  Calls of <= 384 theta values carry
 * theta inside the kernel arguments; larger ones stage it through pinned memory and replay copy-in ->
 * kernel as one CUDA graph from the third call of a shape on.  Either way the kernel writes every walker's
 * result straight into pinned memory as two self-validating words (half of the double + the call's tag
 * each), which the host polls: no copy-out, no fence or flag on the device, no stream synchronisation.
 * Environment overrides for tests and A/B runs: MCD_HOST_CALL=graph (always the staged path) | sync
 * (staged path with copy-out node and stream synchronisation), MCD_GEOMETRY=<tile>,<tiles per CTA>,
 * MCD_XCHG=flags (cross-GPU exchange with flags + fences instead of self-validating words). */
int mcd_lnprob(mcd_handle *h, const double *theta_host, int32_t n_walkers, double *out_host);
int mcd_lnprob_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev, void *stream);

/* Star-sharded partial: out = (sum over THIS shard's stars) + (0 | -inf prior mask), so that a
 * plain sum over shards (NCCL allreduce) yields lnprob.  Replaces walker-parallel
 * multiprocessing (analysis/runner.py:398-403) at multi-GPU scale. */
int mcd_lnprob_partial_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev,
                              void *stream);

/* Fused compute + collective for star-sharded catalogues: the likelihood kernel itself exchanges the
 * per-walker shard sums over NVLink peer mappings (symmetric memory) and adds them in rank order, so
 * no separate all-reduce is launched and every GPU returns bit-identical values.
 *   mcd_exchange_bytes : size of the per-rank exchange buffer (zero-filled by the caller)
 *   mcd_exchange_attach: peer_buffers[r] = address, valid in THIS process, of rank r's buffer
 *   mcd_lnprob_allreduce_device: lnprob of the whole catalogue on every rank; a collective call --
 *       every rank must make the same sequence of calls with the same theta and n_walkers. */
int mcd_exchange_bytes(int32_t world, int32_t max_walkers, int64_t *bytes_out);
int mcd_exchange_attach(mcd_handle *h, int32_t rank, int32_t world, const uint64_t *peer_buffers, int32_t max_walkers);
int mcd_lnprob_allreduce_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev, void *stream);
/* The same with HOST buffers -- what a host sampler calls per half-ensemble on every rank (the
 * vectorised log_prob_fn of analysis/runner.py:403 on a star-sharded catalogue): pinned copy-in and the
 * shard kernel with the exchange in its tail, replayed as ONE CUDA graph from the third call of a shape on
 * (the call's exchange tag travels to the device inside the theta copy); the kernel writes the results
 * into pinned memory as self-validating words, which the host polls.  Returns -5 if a peer
 * never published its sums (the values are NaN then). */
int mcd_lnprob_allreduce(mcd_handle *h, const double *theta_host, int32_t n_walkers, double *out_host);
/* After the caller synchronised its stream: 0, or -5 (and a message) if a kernel launched through one of
 * the *_device collective entry points gave up waiting for a peer since the last check. */
int mcd_exchange_status(mcd_handle *h);

/* Per-star log-likelihood of ONE parameter vector (`no_sum=True`, model.py:565,620-621). */
int mcd_lnlike_per_star(mcd_handle *h, const double *theta_host, double *out_host /* [N] */);
int mcd_lnlike_per_star_device(mcd_handle *h, const double *theta_dev, double *out_dev, void *stream);

/* A-posteriori membership probability of every star at ONE parameter vector (the posterior median
 * in the reference): m e^lc / (m e^lc + (1-m) e^lb), max-shifted -- constant.py:366-374,
 * model.py:458-510,625-687.  Needs a background mode other than MCD_BG_NONE. */
int mcd_membership_per_star(mcd_handle *h, const double *theta_host, double *out_host /* [N] */);
int mcd_membership_per_star_device(mcd_handle *h, const double *theta_dev, double *out_dev, void *stream);

/* The model curves at the stars for ONE parameter vector: v_los[N] and sigma_los[N] in km/s, i.e. what
 * <Model>.rotation_model / <Model>.dispersion_model return (constant.py:52-74,76-111; model.py:93-128,130-180),
 * evaluated with library division / sqrt exactly as the per-star likelihood does.  theta as for mcd_lnlike
 * (the free parameters of ONE walker); either output may be NULL. */
int mcd_model_per_star(mcd_handle *h, const double *theta_host, double *v_los_host /* [N] */,
                       double *sigma_los_host /* [N] */);
int mcd_model_per_star_device(mcd_handle *h, const double *theta_dev, double *v_los_dev, double *sigma_los_dev,
                              void *stream);

/* Runner._calculate_lnlike(v_los, sigma_los) (analysis/runner.py:240-286) for model curves the CALLER computed,
 * e.g. a user-defined model class on top of Runner: the Gaussian sum (runner.py:264-271) or, when the handle was
 * packed with pmember and lnlike_background (MCD_BG_FIXED_PMEMBER), the max-shifted mixture (runner.py:272-286).
 * v_los[N], sigma_los[N] in km/s, HOST arrays; out_host: one double.  Not available for the classes with a fitted
 * background fraction (they do not go through _calculate_lnlike in the reference either). */
int mcd_calculate_lnlike(mcd_handle *h, const double *v_los_host, const double *sigma_los_host, double *out_host);

/* Background precompute: SingleStars.__call__ (background/single_stars.py:42-77) without the
 * M x N intermediate.  v_bg[M], v[N], verr[N] are HOST arrays; out[N]. */
int mcd_single_stars_lnlike(int32_t device, const double *v_bg, int64_t m, const double *v, const double *verr,
                            int64_t n, double sigma_int, double *out_host);
/* The same on DEVICE arrays, asynchronous on `stream`; v_bg_sorted_dev[M] must be in ascending order
 * (the kernel finds each star's nearest background velocity -- the exponent shift of
 * single_stars.py:74 -- by binary search). */
int mcd_single_stars_lnlike_device(int32_t device, const double *v_bg_sorted_dev, int64_t m, const double *v_dev,
                                   const double *verr_dev, int64_t n, double sigma_int, double *out_dev, void *stream);

/* Gaussian.__call__ (background/gaussian.py:23-28); v[N], verr[N] HOST arrays, out[N]. */
int mcd_gaussian_lnlike(int32_t device, const double *v, const double *verr, int64_t n, double mean, double sigma,
                        double *out_host);

/* Device-resident affine-invariant ensemble sampler (emcee's default red/blue StretchMove(a=2)
 * as driven by analysis/runner.py:403,416-419).  All state lives on the device.  mcd_ensemble_run picks
 * one of two engines per call: the resident chain kernel (the whole run is one cooperative launch, every
 * SM keeps a slice of the stars in shared memory; catalogues up to ~9e5 stars, <= 1024 walkers, one GPU)
 * or a CUDA graph of likelihood launches with proposal and acceptance fused in (any size, star shards).
 * Environment overrides for tests and sweeps: MCD_NO_RESIDENT_CHAIN=1, MCD_FORCE_RESIDENT_CHAIN=1,
 * MCD_CHAIN_GROUP=<CTAs per segment>, MCD_CHAIN_EXCHANGE=tagged|counter. */
typedef struct mcd_ensemble mcd_ensemble;
int mcd_ensemble_create(mcd_handle *h, int32_t n_walkers, uint64_t seed, double stretch_a, mcd_ensemble **out);
void mcd_ensemble_destroy(mcd_ensemble *e);
/* set positions [n_walkers][n_theta] (host) and compute their lnprob */
int mcd_ensemble_set_state(mcd_ensemble *e, const double *pos_host);
/* advance n_steps; if chain_host != NULL store every step: chain [n_steps][n_walkers][n_theta],
 * lnprob [n_steps][n_walkers]; n_accepted (may be NULL) [n_walkers] cumulative */
int mcd_ensemble_run(mcd_ensemble *e, int32_t n_steps, double *chain_host, double *lnprob_host,
                     int64_t *n_accepted_host);
int mcd_ensemble_get_state(mcd_ensemble *e, double *pos_host, double *lnprob_host);
/* which engine the last mcd_ensemble_run took (diagnostics): *engine 1 = resident chain kernel (the
 * catalogue stays in shared memory for the whole run), 2 = CUDA graph of fused likelihood launches,
 * 0 = nothing run yet; *ctas_per_segment = CTAs sharing one segment's stars in the resident kernel */
int mcd_ensemble_engine(const mcd_ensemble *e, int32_t *engine, int32_t *ctas_per_segment);

/* Roofline denominators measured on the device the caller is on: dependent-free DFMA chains
 * (TFLOP/s, FMA = 2) and a streaming read (GB/s). */
int mcd_measure_fp64_peak(int32_t device, double *tflops_out, double *ms_out);
int mcd_measure_read_bandwidth(int32_t device, int64_t bytes, double *gbs_out);

#ifdef __cplusplus
}
#endif
#endif /* MCD_B200_H */

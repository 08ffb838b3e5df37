// Internal declarations shared by the translation units of libmcd_b200.so.
// Nothing in here crosses the C ABI (include/mcd_b200.h does).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcd_b200.h"

namespace mcd {

constexpr int kBlock = 256;        // threads per CTA of the lnlike kernel
constexpr int kMaxCols = 8;        // packed float64 columns per star
constexpr int kMaxTile = 256;      // stars per shared-memory stage
constexpr int kStages = 2;         // TMA bulk-copy stages in flight per CTA
constexpr int kMaxRanks = 8;        // GPUs of one box that can share a star-sharded catalogue
constexpr int kMaxXchgGroups = 64;  // walker groups per call of the fused cross-GPU reduction
constexpr int kXchgSlots = 4;       // exchange buffers in rotation: 0/1 host-counted calls, 2/3 sampler half-steps
constexpr int kGatherDepth = 8;     // per-CTA sums one thread of the reducing CTA loads per L2 round trip
constexpr int kWaves = 8;          // CTA waves a large catalogue is cut into (tail balance)
constexpr double kDeg2Rad = 0.017453292519943295769236907684886;
constexpr double kR0Arcmin = 3437.7467707849392526078892888463;   // 10800/pi, calc_xy_offset.py:11

// Raw star columns as uploaded (reference units: deg, km/s) -- kept on the device so that a
// change of the parameter routing (fixed centre moved, parameter fixed/freed) re-packs without
// another host transfer.
struct RawColumns {
    const double *ra, *dec, *v, *verr, *pmember, *density, *lbg;
};

// Everything the pack kernel needs to turn raw columns into the per-variant packed columns.
struct PackParams {
    RawColumns raw;
    int n_segments;                 // > 1: stars are scattered to 16-aligned segment positions
    const long long *seg_begin;     // [n_segments + 1]
    const long long *seg_packed;    // [n_segments]
    double *cols[kMaxCols];
    int32_t *icol;
    long long n_stars;
    int rotation, background, free_centre, math_mode;
    double ra_c_deg, dec_c_deg;     // fixed centre (ignored when free_centre)
    double ra0_deg;                 // reference right ascension of the free-centre expansion
};

// Fused ensemble half-step (device-resident sampler): the likelihood kernel draws the stretch-move
// proposals of the active half itself and its finishing CTA accepts or rejects them in place, so a
// half-step is ONE launch instead of propose -> lnprob -> accept.
struct FuseParams {
    int enabled;
    int half;                     // 0: red walkers move, 1: blue
    int n0;                       // size of the red half; LaunchParams::n_walkers is the active half's size
    int walkers_total;            // W per segment
    double a;                     // stretch scale
    unsigned long long seed;
    double *pos;                  // [S][W][P] current positions, updated in place on acceptance
    double *lnp;                  // [S][W]
    const int *perm;              // [S][W] red/blue partition of this step
    long long *n_accepted;        // [S][W]
    const unsigned int *step;     // [0]: global step counter (Philox counter word)
    unsigned long long tag_base;  // exchange tag of this ensemble's half-steps: (1 << 62) | nonce << 36
};

// Whole chains inside one kernel (small catalogues): one CTA per segment keeps the segment's packed
// columns and the ensemble state in shared memory and runs `n_steps` stretch-move iterations without
// leaving the SM.
constexpr int kChainBlock = 512;      // threads per CTA of the resident-chain kernel
constexpr int kChainMaxWalkers = 1024;
struct TaggedSum;
struct ChainParams {
    int n_steps;
    int n_walkers;                // W per segment
    int n0;                       // red half
    int max_segment_padded;       // column stride in shared memory (stars, multiple of 16)
    int group;                    // CTAs sharing one segment's stars (1: one CTA per segment, nothing exchanged)
    int stars_per_cta;            // even; CTA r of a group owns stars [r * stars_per_cta, ...)
    struct TaggedSum *group_sums; // [2][S][group][sum_stride] per-CTA sums (tagged words or doubles), zero at launch
    unsigned long long *group_arrivals;   // [S] CTAs arrived so far, zero at launch (tagged == 0)
    int sum_stride;               // walkers per exchange = min(kChainBlock, larger half of the ensemble)
    int tagged;                   // 1: sums travel as self-validating tagged words; 0: plain doubles behind a counter
    int *status;                  // set non-zero if a wait for a group member ran into its time limit
    unsigned int step0;           // global step counter at entry
    double a;
    unsigned long long seed;
    double *pos;                  // [S][W][P] in/out
    double *lnp;                  // [S][W] in/out
    long long *n_accepted;        // [S][W] accumulated
    double *chain;                // [n_steps][S][W][P] or nullptr
    double *chain_lnp;            // [n_steps][S][W] or nullptr
};

// Small host-buffer calls (C1 / C2-sized: a few hundred theta values) skip the copy nodes altogether: theta
// travels INSIDE the kernel arguments (constant bank), the result is written straight into pinned host memory
// and the finishing CTA raises a flag there that the host spins on -- one launch, no cudaMemcpy, no stream
// synchronisation.  kThetaInline doubles = 3 KiB of the 32 KiB argument space.
constexpr int kThetaInline = 384;
struct ThetaBlock {
    double v[kThetaInline];
};

// Kernel argument block of one lnlike / lnprob launch (passed by value, __grid_constant__).
struct LaunchParams {
    const double *cols[kMaxCols];
    const int32_t *icol;
    long long n_stars;
    int tile;                 // stars per stage (multiple of 16, <= kMaxTile)
    int n_tiles;              // of the largest segment
    int tiles_per_chunk;
    int n_chunks;             // chunks of the largest segment (gridDim.x = n_chunks * n_groups)
    int n_segments;           // gridDim.y: independent (star range, walker set) problems, >= 1
    const long long *seg_begin;   // [n_segments + 1] star index boundaries, or nullptr for one segment
    const long long *seg_packed;  // [n_segments] 16-aligned position of each segment in the packed columns
    int n_groups;             // walker groups
    int n_walkers;
    int wl;                   // walkers per CTA
    int slices;               // star slices per CTA: kBlock / wl
    int n_theta;
    int apply_prior;          // 1: lnprob (box prior fused), 0: lnlike
    int fixed_prior_ok;
    double verr2_unscale;         // per-star kernel: factor that takes the packed verr^2 column back to verr^2
    const double *theta;      // [n_segments][n_walkers][n_theta]
    int super;                // chunks per super-chunk (level 1 of the cross-CTA reduction)
    int n_super;              // super-chunks per walker group
    double *partials;         // [n_segments][n_chunks][n_walkers]
    double *partials2;        // [n_segments][n_super][n_walkers]
    unsigned int *counters;   // [n_segments][n_groups][n_super + 1], zero between launches
    double *out;              // [n_segments][n_walkers]
    // Host-buffer calls: the result goes straight into pinned host memory as self-validating words -- row r is
    // {low half | tag << 32, high half | tag << 32} at host_words[2 r], one 16-byte store per walker, no fence,
    // no counter, no flag: the host polls the rows until both tags of each carry the call's tag.
    unsigned long long *host_words;           // non-null: host-buffer call (`out` is not written)
    unsigned int host_tag;                    // low 32 bits of the call's sequence number, never 0
    const unsigned long long *host_seq_ptr;   // non-null: the sequence number is read from device memory (graph replays)
    int slot[MCD_NPARAM];
    double fixed_scaled[MCD_NPARAM];   // fixed value already multiplied by its unit scale
    double scale[MCD_NPARAM];
    double lower[MCD_MAX_THETA], upper[MCD_MAX_THETA];
    double ra0_deg;
    // fused cross-GPU reduction (star-sharded catalogues): the CTA that finishes a walker group
    // publishes this shard's sums into every rank's exchange buffer over NVLink peer mappings, waits
    // for the other shards' sums and adds them in rank order.  xchg_world <= 1: disabled.
    int xchg_world, xchg_rank, xchg_capacity;
    unsigned long long xchg_epoch;                 // call counter, identical on all ranks, starts at 1
    const unsigned long long *xchg_epoch_ptr;      // non-null: the epoch is read from device memory (graph replays)
    int *xchg_status;                              // set to 1 when the wait for a peer ran into its time limit
    double *xchg_data[kMaxRanks];                  // rank p's data region  [kXchgSlots][world][capacity]
    unsigned long long *xchg_flags[kMaxRanks];     // rank p's flag region  [kXchgSlots][world][kMaxXchgGroups]
    // tagged exchange (xchg_tagged_mode = 1): every sum travels as two 8-byte words, each carrying half of the
    // double and the call's 32-bit tag, so a word validates itself -- no flag, no system fence, no barrier
    unsigned long long *xchg_words[kMaxRanks];     // rank p's tagged region [kXchgSlots][world][capacity][2]
    int xchg_tagged_mode;
    FuseParams fuse;
};

struct Variant {
    int rotation, free_centre, background, math_mode;
};

// number of packed float64 columns of a variant / whether it has the int32 exponent column
int variant_columns(const Variant &v);
bool variant_has_icol(const Variant &v);
int variant_flops_per_term(const Variant &v);

double mix_var_scale();      // factor carried by the packed verr^2 column of the FAST mixture variants (mcd_math.cuh)
cudaError_t launch_pack(const PackParams &p, cudaStream_t stream);
// `inline_theta` (may be nullptr) is handed to the kernel by value next to `p`
cudaError_t launch_lnlike(const Variant &v, const LaunchParams &p, cudaStream_t stream, const ThetaBlock *inline_theta = nullptr);
// per-star lnlike (membership = 0) or membership probability (1) of walker 0 of p.theta into out[N]
// (always PLAIN arithmetic)
// what the per-star kernel writes: lnlike (no_sum), membership probability, or the model curves (v_los, sigma_los)
enum { kPerStarLnlike = 0, kPerStarMembership = 1, kPerStarModel = 2 };
cudaError_t launch_per_star(const Variant &v, const LaunchParams &p, double *out, double *out2, int mode,
                            cudaStream_t stream);
// resident-chain kernel: shared memory it needs for this problem, or 0 if the problem does not fit
size_t chain_shared_bytes(const Variant &v, long long stars_per_cta, int n_walkers, int n_theta);
cudaError_t launch_chain(const Variant &v, const LaunchParams &p, const ChainParams &c, size_t smem, cudaStream_t stream);
// SingleStars background column (mcd_background.cu): device pointers, v_bg sorted ascending
// Runner._calculate_lnlike for caller-supplied model curves (csrc/mcd_background.cu)
int curve_lnlike_blocks(long long n, int sm_count);
cudaError_t launch_curve_lnlike(const double *v_dev, const double *verr_dev, const double *pmember_dev, const double *lbg_dev,
                                const double *v_los_dev, const double *sigma_los_dev, long long n, double *scratch_dev,
                                double *out_dev, int sm_count, cudaStream_t stream);
cudaError_t launch_single_stars(const double *v_bg_sorted_dev, long long m, const double *v_dev, const double *verr_dev,
                                long long n, double sigma_int, double *out_dev, int sm_count, cudaStream_t stream);
// resident CTAs per SM of the lnlike kernel of this variant (occupancy API)
int lnlike_blocks_per_sm(const Variant &v);

// one lnlike (apply_prior = 0) / lnprob (1) launch on `stream`, device pointers (mcd_api.cu)
int launch_ensemble(mcd_handle *h, const double *theta_dev, int n_walkers, double *out_dev, int apply_prior,
                    cudaStream_t stream);
// the same with the proposal and the acceptance of an ensemble half-step fused in (theta is drawn in
// the kernel; n_walkers = size of the active half)
int launch_ensemble_fused(mcd_handle *h, int n_walkers, const FuseParams &fuse, cudaStream_t stream);
// whole chains in one launch when the catalogue fits shared memory; returns 1 if not eligible
int launch_resident_chain(mcd_handle *h, const ChainParams &chain, cudaStream_t stream);
// after the stream was synchronised: < 0 if a CTA group of the last resident launch lost a member
int resident_chain_status(mcd_handle *h, cudaStream_t stream);
int resident_chain_group(const mcd_handle *h);   // CTAs per segment of the last resident launch
int handle_device(const mcd_handle *h);
// make `stream` wait for whatever was last launched through the handle on another stream (mcd_api.cu)
int order_on_stream(mcd_handle *h, cudaStream_t stream);
// `stream` is about to be destroyed (and has been synchronised): never record an event on it again
void forget_stream(mcd_handle *h, cudaStream_t stream);
// changes whenever something a captured CUDA graph bakes in changes (scratch, columns, routing, exchange)
unsigned long long handle_generation(const mcd_handle *h);
// per-ensemble nonce of the fused half-step exchange tags (identical on every rank: same creation order)
unsigned long long next_fuse_nonce(mcd_handle *h);
// after `stream` was synchronised: < 0 if a kernel gave up waiting for a peer's shard sums
int exchange_status(mcd_handle *h, cudaStream_t stream);
// record the thread-local message returned by mcd_last_error() and hand back `code`
int set_error(int code, const char *fmt, ...);

}  // namespace mcd

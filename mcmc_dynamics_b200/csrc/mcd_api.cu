// C ABI of libmcd_b200.so (declared in include/mcd_b200.h): handle management, packing, launch
// geometry and the host-buffer convenience entry points.  No torch, no C++ types across the ABI.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "mcd_internal.h"

using namespace mcd;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_error[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int mcd::set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

#define MCD_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t err__ = (call);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            return fail(-2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

enum { RAW_RA = 0, RAW_DEC, RAW_V, RAW_VERR, RAW_PMEMBER, RAW_DENSITY, RAW_LBG, RAW_COUNT };

struct mcd_handle {
    int device = 0;
    Variant var{};
    mcd_pack_desc desc{};          // routing part of the descriptor (column pointers cleared)
    long long n = 0, n_alloc = 0;
    int n_segments = 1;            // independent star ranges (radial bins) evaluated by one launch
    long long max_segment = 0;     // stars of the largest segment
    long long *seg_begin = nullptr, *seg_packed = nullptr;   // device copies, n_segments > 1 only
    double *raw[RAW_COUNT] = {};
    double *cols[kMaxCols] = {};
    int32_t *icol = nullptr;
    double ra0_deg = 0.0;
    // launch scratch
    double *partials = nullptr, *partials2 = nullptr;
    size_t partials_cap = 0, partials2_cap = 0;
    unsigned int *counters = nullptr;
    int counters_cap = 0;
    // resident chains with several CTAs per segment
    unsigned char *group_sums = nullptr;            // status word (16 bytes) followed by the tagged sums
    size_t group_sums_cap = 0;
    int chain_group = 0;           // CTAs per segment of the last resident launch
    // staging for the host-buffer entry points
    double *theta_dev = nullptr, *out_dev = nullptr, *theta_pin = nullptr, *out_pin = nullptr;
    size_t theta_cap = 0, out_cap = 0;
    // launch geometry of the most recent walker counts (the search costs ~3 us: too much to redo on every small call)
    struct Geometry {
        int n_walkers = -1;
        unsigned long long generation = 0;
        int wl, slices, n_groups, tile, n_tiles, tiles_per_chunk, n_chunks, super, n_super;
    };
    Geometry geometry[4];
    int geometry_next = 0;
    // inline host-buffer calls: completion flag in pinned memory, counter of finished walker groups on the device
    unsigned long long host_seq = 0;            // host-buffer calls: sequence number, its low 32 bits tag the result words
    double *star_dev = nullptr;    // [n] scratch of the per-star entry point
    cudaStream_t stream = nullptr;
    int sm_count = 0, blocks_per_sm = 1;
    // host-buffer calls replayed as CUDA graphs (H2D theta -> kernel -> D2H result), one per
    // (walkers per call, prior on/off); dropped whenever a pointer or the routing baked into them changes
    struct HostGraph {
        int n_walkers = 0, apply_prior = 0, exchange = 0, seen = 0;
        cudaGraphExec_t exec = nullptr;
    };
    HostGraph host_graphs[8];
    // fused cross-GPU reduction (mcd_exchange_attach)
    int xchg_world = 0, xchg_rank = 0, xchg_capacity = 0;
    unsigned long long xchg_epoch = 0;
    unsigned long long fuse_nonce = 0;          // ensembles created on this handle so far (exchange tag of their half-steps)
    double *xchg_data[kMaxRanks] = {};
    unsigned long long *xchg_flags[kMaxRanks] = {};
    unsigned long long *xchg_words[kMaxRanks] = {};
    int xchg_tagged_mode = 1;                   // MCD_XCHG=flags selects the flag + fence protocol (A/B)
    int *xchg_status = nullptr;                 // device word: set to 1 by a kernel whose wait for a peer timed out
    // Everything a captured graph bakes in (scratch pointers, packed columns, routing, exchange buffers)
    // belongs to one generation; whoever replays a graph compares the generation it captured at.
    unsigned long long generation = 1;
    // launches through one handle share its reduction scratch: they are ordered across streams by making
    // a launch on a new stream wait for everything issued so far on the previous one
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    cudaEvent_t order_event = nullptr;
    mcd_info info{};
};

static void drop_host_graphs(mcd_handle *h);

extern "C" int mcd_abi_version(void) { return MCD_ABI_VERSION; }
extern "C" const char *mcd_last_error(void) { return g_error; }

// ------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------
static int validate_routing(const mcd_pack_desc *d) {
    if (d->rotation != MCD_ROT_CONSTANT && d->rotation != MCD_ROT_RADIAL) return fail(-1, "unknown rotation model %d", d->rotation);
    if (d->background < MCD_BG_NONE || d->background > MCD_BG_GAUSSIAN) return fail(-1, "unknown background mode %d", d->background);
    if (d->math_mode != MCD_MATH_FAST && d->math_mode != MCD_MATH_PLAIN) return fail(-1, "unknown math mode %d", d->math_mode);
    if (d->n_theta < 0 || d->n_theta > MCD_MAX_THETA) return fail(-1, "n_theta = %d outside [0, %d]", d->n_theta, MCD_MAX_THETA);
    if (d->n_stars < 0) return fail(-1, "n_stars < 0");
    for (int k = 0; k < MCD_NPARAM; ++k)
        if (d->slot[k] >= d->n_theta) return fail(-1, "slot[%d] = %d but theta has %d columns", k, d->slot[k], d->n_theta);
    return 0;
}

static int repack(mcd_handle *h) {
    const mcd_pack_desc &d = h->desc;
    drop_host_graphs(h);
    h->generation += 1;
    h->var.rotation = d.rotation;
    h->var.background = d.background;
    h->var.math_mode = d.math_mode;
    h->var.free_centre = (d.slot[MCD_P_RA_CENTER] >= 0 || d.slot[MCD_P_DEC_CENTER] >= 0) ? 1 : 0;

    const int nc = variant_columns(h->var);
    for (int c = 0; c < kMaxCols; ++c) {
        if (c < nc && !h->cols[c]) {
            MCD_CUDA(cudaMalloc(&h->cols[c], sizeof(double) * h->n_alloc));
            MCD_CUDA(cudaMemsetAsync(h->cols[c], 0, sizeof(double) * h->n_alloc, h->stream));
        }
    }
    if (variant_has_icol(h->var) && !h->icol) {
        MCD_CUDA(cudaMalloc(&h->icol, sizeof(int32_t) * h->n_alloc));
        MCD_CUDA(cudaMemsetAsync(h->icol, 0, sizeof(int32_t) * h->n_alloc, h->stream));
    }
    if ((d.background == MCD_BG_FIXED_PMEMBER) && !h->raw[RAW_PMEMBER]) return fail(-1, "background mode needs the pmember column");
    if ((d.background == MCD_BG_FIXED_DENSITY || d.background == MCD_BG_GAUSSIAN) && !h->raw[RAW_DENSITY])
        return fail(-1, "background mode needs the density column");
    if ((d.background == MCD_BG_FIXED_PMEMBER || d.background == MCD_BG_FIXED_DENSITY) && !h->raw[RAW_LBG])
        return fail(-1, "background mode needs the lnlike_background column");

    PackParams p{};
    p.raw.ra = h->raw[RAW_RA];
    p.raw.dec = h->raw[RAW_DEC];
    p.raw.v = h->raw[RAW_V];
    p.raw.verr = h->raw[RAW_VERR];
    p.raw.pmember = h->raw[RAW_PMEMBER];
    p.raw.density = h->raw[RAW_DENSITY];
    p.raw.lbg = h->raw[RAW_LBG];
    for (int c = 0; c < kMaxCols; ++c) p.cols[c] = h->cols[c];
    p.icol = h->icol;
    p.n_stars = h->n;
    p.n_segments = h->n_segments;
    p.seg_begin = h->seg_begin;
    p.seg_packed = h->seg_packed;
    p.rotation = d.rotation;
    p.background = d.background;
    p.free_centre = h->var.free_centre;
    p.math_mode = d.math_mode;
    p.ra_c_deg = d.fixed_value[MCD_P_RA_CENTER] * d.unit_scale[MCD_P_RA_CENTER];
    p.dec_c_deg = d.fixed_value[MCD_P_DEC_CENTER] * d.unit_scale[MCD_P_DEC_CENTER];
    // reference right ascension of the free-centre expansion: the parameter's current value
    h->ra0_deg = p.ra_c_deg;
    p.ra0_deg = h->ra0_deg;
    MCD_CUDA(launch_pack(p, h->stream));
    MCD_CUDA(cudaStreamSynchronize(h->stream));

    h->blocks_per_sm = lnlike_blocks_per_sm(h->var);
    h->info.n_stars = h->n;
    h->info.n_theta = d.n_theta;
    h->info.n_columns = nc;
    h->info.bytes_per_star = 8 * nc + (variant_has_icol(h->var) ? 4 : 0);
    h->info.flops_per_term = variant_flops_per_term(h->var);
    h->info.free_centre = h->var.free_centre;
    h->info.sm_count = h->sm_count;
    h->info.n_segments = h->n_segments;
    return 0;
}

static void drop_host_graphs(mcd_handle *h) {
    h->generation += 1;      // whatever invalidates the host-call graphs invalidates every other captured graph too
    for (auto &g : h->host_graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = mcd_handle::HostGraph();
    }
}

extern "C" void mcd_destroy(mcd_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    drop_host_graphs(h);
    for (auto &p : h->raw) cudaFree(p);
    for (auto &p : h->cols) cudaFree(p);
    cudaFree(h->icol);
    cudaFree(h->seg_begin);
    cudaFree(h->seg_packed);
    cudaFree(h->partials);
    cudaFree(h->partials2);
    cudaFree(h->counters);
    cudaFree(h->group_sums);
    cudaFree(h->theta_dev);
    cudaFree(h->out_dev);
    cudaFree(h->star_dev);
    cudaFree(h->xchg_status);
    if (h->order_event) cudaEventDestroy(h->order_event);
    cudaFreeHost(h->theta_pin);
    cudaFreeHost(h->out_pin);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// Segments: every star range starts at a 16-star boundary of the packed columns so that the bulk
// copies of each segment are 16-byte aligned for both the float64 and the int32 column.
static int setup_segments(mcd_handle *h, const mcd_pack_desc *desc) {
    const int S = desc->n_segments;
    if (!desc->segment_offsets) return fail(-1, "n_segments > 1 needs segment_offsets");
    // the per-bin fits of bin/run.py:186 are ConstantFit(data_i, parameters, background=background):
    // no background component, or the fixed-background mixture with per-star pmember
    if (desc->background != MCD_BG_NONE && desc->background != MCD_BG_FIXED_PMEMBER)
        return fail(-1, "segmented handles support the models without fitted background parameters");
    if (S > 65535) return fail(-1, "at most 65535 segments per handle");
    std::vector<long long> begin(S + 1), packed(S);
    long long pos = 0, longest = 0;
    for (int s = 0; s <= S; ++s) begin[s] = desc->segment_offsets[s];
    if (begin[0] != 0 || begin[S] != h->n) return fail(-1, "segment_offsets must run from 0 to n_stars");
    for (int s = 0; s < S; ++s) {
        const long long count = begin[s + 1] - begin[s];
        if (count < 0) return fail(-1, "segment_offsets must be non-decreasing");
        if (count > 2000000000LL) return fail(-1, "a segment holds at most 2e9 stars");
        packed[s] = pos;
        pos += ((count + 15) / 16) * 16;
        longest = std::max(longest, count);
    }
    h->n_segments = S;
    h->max_segment = longest;
    h->n_alloc = ((pos + kMaxTile - 1) / kMaxTile + 1) * kMaxTile;
    MCD_CUDA(cudaMalloc(&h->seg_begin, sizeof(long long) * (S + 1)));
    MCD_CUDA(cudaMalloc(&h->seg_packed, sizeof(long long) * S));
    MCD_CUDA(cudaMemcpy(h->seg_begin, begin.data(), sizeof(long long) * (S + 1), cudaMemcpyHostToDevice));
    MCD_CUDA(cudaMemcpy(h->seg_packed, packed.data(), sizeof(long long) * S, cudaMemcpyHostToDevice));
    return 0;
}

static int upload_column(mcd_handle *h, int which, const double *host) {
    if (!host) return 0;
    MCD_CUDA(cudaMalloc(&h->raw[which], sizeof(double) * std::max<long long>(h->n, 1)));
    if (h->n > 0)
        MCD_CUDA(cudaMemcpyAsync(h->raw[which], host, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

extern "C" int mcd_pack_create(const mcd_pack_desc *desc, mcd_handle **out) {
    if (!desc || !out) return fail(-1, "null argument");
    *out = nullptr;
    if (int rc = validate_routing(desc)) return rc;
    if (desc->n_stars > 0 && (!desc->ra || !desc->dec || !desc->v || !desc->verr))
        return fail(-1, "ra, dec, v and verr columns are required");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return fail(-3, "no CUDA device is visible: the B200 path has no CPU fallback");
    if (desc->device < 0 || desc->device >= count) return fail(-1, "device %d not in [0, %d)", desc->device, count);
    mcd_handle *h = new (std::nothrow) mcd_handle();
    if (!h) return fail(-4, "out of host memory");
    h->device = desc->device;
    int rc = 0;
    do {
        if (cudaSetDevice(h->device) != cudaSuccess) { rc = fail(-2, "cudaSetDevice(%d) failed", h->device); break; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) { rc = fail(-2, "cudaGetDeviceProperties failed"); break; }
        h->sm_count = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(-2, "cudaStreamCreate failed"); break; }
        if (cudaEventCreateWithFlags(&h->order_event, cudaEventDisableTiming) != cudaSuccess) { rc = fail(-2, "cudaEventCreate failed"); break; }
        {   // keep what the stream-ordered allocations (ensemble state, chains, background scratch) give back in the
            // device's pool instead of returning it to the driver at every synchronisation
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess) {
                unsigned long long keep = 1ull << 30;
                (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            (void)cudaGetLastError();
        }
        h->n = desc->n_stars;
        h->n_alloc = ((h->n + kMaxTile - 1) / kMaxTile + 1) * kMaxTile;   // full bulk copies at the tail
        h->max_segment = h->n;
        if (desc->n_segments > 1) {
            if ((rc = setup_segments(h, desc))) break;
        }
        h->desc = *desc;
        h->desc.ra = h->desc.dec = h->desc.v = h->desc.verr = nullptr;
        h->desc.pmember = h->desc.density = h->desc.lnlike_background = nullptr;
        if ((rc = upload_column(h, RAW_RA, desc->ra))) break;
        if ((rc = upload_column(h, RAW_DEC, desc->dec))) break;
        if ((rc = upload_column(h, RAW_V, desc->v))) break;
        if ((rc = upload_column(h, RAW_VERR, desc->verr))) break;
        if ((rc = upload_column(h, RAW_PMEMBER, desc->pmember))) break;
        if ((rc = upload_column(h, RAW_DENSITY, desc->density))) break;
        if ((rc = upload_column(h, RAW_LBG, desc->lnlike_background))) break;
        if ((rc = repack(h))) break;
    } while (0);
    if (rc) {
        mcd_destroy(h);
        return rc;
    }
    *out = h;
    return 0;
}

extern "C" int mcd_pack_reconfigure(mcd_handle *h, const mcd_pack_desc *desc) {
    if (!h || !desc) return fail(-1, "null argument");
    if (int rc = validate_routing(desc)) return rc;
    if (desc->n_stars != h->n) return fail(-1, "reconfigure cannot change the catalogue (n_stars %lld != %lld)", (long long)desc->n_stars, h->n);
    if (desc->n_segments > 1 && desc->n_segments != h->n_segments) return fail(-1, "reconfigure cannot change the segments");
    MCD_CUDA(cudaSetDevice(h->device));
    h->desc = *desc;
    h->desc.ra = h->desc.dec = h->desc.v = h->desc.verr = nullptr;
    h->desc.pmember = h->desc.density = h->desc.lnlike_background = nullptr;
    return repack(h);
}

extern "C" int mcd_get_info(const mcd_handle *h, mcd_info *info) {
    if (!h || !info) return fail(-1, "null argument");
    *info = h->info;
    return 0;
}

// ------------------------------------------------------------------------------------------
// launch geometry
// ------------------------------------------------------------------------------------------
static void search_geometry(const mcd_handle *h, int n_walkers, LaunchParams &p);

// cached per (walker count, pack generation); MCD_GEOMETRY (experiments) always searches
static void choose_geometry(mcd_handle *h, int n_walkers, LaunchParams &p) {
    if (!getenv("MCD_GEOMETRY")) {
        for (const auto &g : h->geometry)
            if (g.n_walkers == n_walkers && g.generation == h->generation) {
                p.wl = g.wl; p.slices = g.slices; p.n_groups = g.n_groups; p.tile = g.tile; p.n_tiles = g.n_tiles;
                p.tiles_per_chunk = g.tiles_per_chunk; p.n_chunks = g.n_chunks; p.super = g.super; p.n_super = g.n_super;
                return;
            }
    }
    search_geometry(h, n_walkers, p);
    if (getenv("MCD_GEOMETRY")) return;
    mcd_handle::Geometry &g = h->geometry[h->geometry_next++ % 4];
    g.n_walkers = n_walkers;
    g.wl = p.wl; g.slices = p.slices; g.n_groups = p.n_groups; g.tile = p.tile; g.n_tiles = p.n_tiles;
    g.tiles_per_chunk = p.tiles_per_chunk; g.n_chunks = p.n_chunks; g.super = p.super; g.n_super = p.n_super;
    g.generation = h->generation;
}

static void search_geometry(const mcd_handle *h, int n_walkers, LaunchParams &p) {
    // walkers per CTA: minimise groups / slices (CTA time ~ stars / slices, CTAs ~ groups)
    int best_g = 1, best_wl = std::max(1, std::min(n_walkers, kBlock)), best_s = 1;
    double best_cost = 1e300;
    const int g_min = std::max(1, (n_walkers + kBlock - 1) / kBlock);
    for (int g = g_min; g <= std::max(g_min, std::min(n_walkers, 4 * g_min + 8)); ++g) {
        const int wl = (n_walkers + g - 1) / g;
        const int gg = (n_walkers + wl - 1) / wl;
        const int s = std::max(1, kBlock / wl);
        const double cost = (double)gg / (double)s;
        if (cost < best_cost * (1.0 - 1e-9)) {
            best_cost = cost;
            best_g = gg;
            best_wl = wl;
            best_s = s;
        }
    }
    p.wl = best_wl;
    p.slices = best_s;
    p.n_groups = best_g;
    // Star tiling: stars per stage (`tile`, multiple of 16) and tiles per CTA are chosen together.  The
    // kernel is FP64-pipe bound, so co-resident CTAs share the pipe and the launch takes as long as the most
    // loaded SM needs for its CTAs one after the other.  Cost of a candidate, in star iterations per thread:
    //   cta   = tiles_per_chunk * (tile / slices + 3) + overhead      (3: barrier + stage hand-over per tile;
    //           overhead: walker set-up and reduction, ~16 iterations' worth of pipe time)
    //   load  = ceil(CTAs / SMs): CTAs on the most loaded SM.  Equal-sized CTAs are dealt out evenly, so a grid
    //           of 8.26 CTAs per SM costs 9 (measured: +9 %), whether or not it fits one wave of resident CTAs.
    //           A partly filled last round of resident CTAs costs extra (see below).
    //   cost  = load * cta * (1 + 2 % / waves^1.5): SMs differ by a per cent or two in speed; several waves of
    //           CTAs even that out, one wave cannot.  Measured on star shards of C5 with the round-2 kernel
    //           (gpurun_out/r2r_geometry.log): 5e6 stars, one wave 3121 us, two 3074, four 3056, eight 3053;
    //           1.25e6 stars, one wave 788.6 us, two 779.3, 3.7 waves 778.2
    // Fitted to (tile, tiles_per_chunk) sweeps of BASELINE configs C3, C4 and of one of eight C5 shards
    // (profiles/r02_ab_runs.md): measured time / cost is constant within +-8 % over 24 mid-size geometries and
    // within 2 % over the 7 shard-sized ones, where the round-1 model (whole waves of SMs x CTAs-per-SM, blind
    // to a partly filled wave) was off by up to 35 % and 9 %.
    const int sms = std::max(1, h->sm_count);
    const int problems = std::max(1, p.n_groups * h->n_segments);      // CTAs per star chunk
    const double overhead = 16.0;      // 24 before the round-2 kernels; 1e7 stars: 24 CTAs per SM 6056 us, 12 per SM 6084 us
    const long long n = std::max<long long>(h->max_segment, 1);     // grid sized for the largest segment
    int min_tile = ((2 * p.slices + 15) / 16) * 16;
    min_tile = std::min(std::max(min_tile, 16), kMaxTile);
    double best = 1e300;
    int best_tile = kMaxTile, best_tpc = 1;
    for (int tile = min_tile; tile <= kMaxTile; tile += 16) {
        const long long n_tiles = (n + tile - 1) / tile;
        // candidates: k CTAs per SM for k = 1 .. kWaves * CTAs-per-SM
        for (int k = 1; k <= kWaves * std::max(1, h->blocks_per_sm); ++k) {
            const long long slots = std::max<long long>(1, (long long)k * sms / problems);
            const long long tpc = std::max<long long>(1, (n_tiles + slots - 1) / slots);
            const long long chunks = (n_tiles + tpc - 1) / tpc;
            const long long ctas = chunks * problems;
            const double cta = (double)tpc * ((double)tile / p.slices + 3.0) + overhead;
            const long long per_sm = (ctas + sms - 1) / sms;
            // An SM works through its CTAs `room` at a time.  A last, partly filled round hides less latency: one
            // CTA alone on an SM runs 25 % slower per term (measured: 8 warps do not fill the FP64 pipe), 2 of 3
            // about 3 % (an estimate that only breaks ties between otherwise equal geometries).
            const int room = std::max(1, h->blocks_per_sm);
            const long long full = per_sm / room, rem = per_sm % room;
            const double rem_cost = rem == 0 ? 0.0 : (rem == 1 ? 1.25 : (double)rem * (1.0 + 0.03 * (double)(room - rem)));
            const double base = ((double)(full * room) + rem_cost) * cta;
            const double waves = std::max(1.0, (double)per_sm / (double)room);
            const double cost = base * (1.0 + 0.02 / (waves * std::sqrt(waves)));
            if (cost < best * (1.0 - 1e-12)) {
                best = cost;
                best_tile = tile;
                best_tpc = (int)tpc;
            }
        }
    }
    if (const char *env = getenv("MCD_GEOMETRY")) {   // "tile,tiles_per_chunk": experiments only
        int t = 0, c = 0;
        if (sscanf(env, "%d,%d", &t, &c) == 2 && t >= min_tile && t <= kMaxTile && t % 16 == 0 && c >= 1) {
            best_tile = t;
            best_tpc = c;
        }
    }
    p.tile = best_tile;
    p.n_tiles = (int)((h->max_segment + best_tile - 1) / best_tile);
    p.tiles_per_chunk = best_tpc;
    p.n_chunks = std::max(1, (p.n_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk);
    // Cross-CTA reduction: the CTA that takes the last ticket adds the per-CTA sums with all its threads,
    // kGatherDepth * slices rows per L2 round trip (measured ~0.9 us each, tools/probe/launch_timeline.sh).
    // One level while that is a single round trip; otherwise two levels with super-chunks of at least one
    // round trip's worth of rows, sqrt(chunks) for large launches.
    const int per_trip = kGatherDepth * p.slices;
    p.super = p.n_chunks <= per_trip ? std::max(1, p.n_chunks)
                                     : std::max(per_trip, (int)std::ceil(std::sqrt((double)p.n_chunks)));
    p.n_super = (p.n_chunks + p.super - 1) / p.super;
}

static int ensure_scratch(mcd_handle *h, const LaunchParams &p) {
    const size_t need = (size_t)p.n_chunks * p.n_walkers * p.n_segments;
    if (need > h->partials_cap) {
        drop_host_graphs(h);
        MCD_CUDA(cudaFree(h->partials));
        h->partials = nullptr;
        h->partials_cap = 0;
        MCD_CUDA(cudaMalloc(&h->partials, sizeof(double) * need));
        h->partials_cap = need;
    }
    const size_t need2 = (size_t)p.n_super * p.n_walkers * p.n_segments;
    if (need2 > h->partials2_cap) {
        drop_host_graphs(h);
        MCD_CUDA(cudaFree(h->partials2));
        h->partials2 = nullptr;
        h->partials2_cap = 0;
        MCD_CUDA(cudaMalloc(&h->partials2, sizeof(double) * need2));
        h->partials2_cap = need2;
    }
    const int n_counters = p.n_segments * p.n_groups * (p.n_super + 1);
    if (n_counters > h->counters_cap) {
        drop_host_graphs(h);
        MCD_CUDA(cudaFree(h->counters));
        h->counters = nullptr;
        h->counters_cap = 0;
        const int cap = std::max(256, n_counters);
        MCD_CUDA(cudaMalloc(&h->counters, sizeof(unsigned int) * cap));
        MCD_CUDA(cudaMemset(h->counters, 0, sizeof(unsigned int) * cap));
        h->counters_cap = cap;
    }
    return 0;
}

static void fill_params(const mcd_handle *h, LaunchParams &p) {
    const mcd_pack_desc &d = h->desc;
    for (int c = 0; c < kMaxCols; ++c) p.cols[c] = h->cols[c];
    p.icol = h->icol;
    p.n_stars = h->n;
    p.n_segments = h->n_segments;
    p.seg_begin = h->seg_begin;
    p.seg_packed = h->seg_packed;
    p.n_theta = d.n_theta;
    p.fixed_prior_ok = d.fixed_prior_ok;
    for (int k = 0; k < MCD_NPARAM; ++k) {
        p.slot[k] = d.slot[k];
        p.scale[k] = d.unit_scale[k];
        p.fixed_scaled[k] = d.fixed_value[k] * d.unit_scale[k];
    }
    for (int j = 0; j < MCD_MAX_THETA; ++j) {
        p.lower[j] = d.lower[j];
        p.upper[j] = d.upper[j];
    }
    p.ra0_deg = h->ra0_deg;
}

// Launches through one handle share its reduction scratch, counters and exchange slots (mcd_b200.h:
// "they must be ordered").  Callers use several streams -- the handle's own for the host-buffer entry
// points, torch's current stream for the tensor ops, an ensemble's stream for the sampler -- so the
// first launch on a different stream than the previous one waits for everything issued there so far.
// Free in the common case (same stream as before).  Inside a stream capture nothing is added: a
// capture follows un-captured launches on the same stream, and a cross-stream edge out of a capture
// is an error (cudaErrorStreamCaptureIsolation).
int mcd::order_on_stream(mcd_handle *h, cudaStream_t stream) {
    if (h->last_stream_valid && h->last_stream != stream) {
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &capturing) != cudaSuccess) (void)cudaGetLastError();
        if (capturing != cudaStreamCaptureStatusNone) return 0;
        if (cudaEventRecord(h->order_event, h->last_stream) == cudaSuccess) {
            MCD_CUDA(cudaStreamWaitEvent(stream, h->order_event, 0));
        } else {
            (void)cudaGetLastError();      // the previous stream no longer exists: its owner synchronised it
        }
    }
    h->last_stream = stream;
    h->last_stream_valid = true;
    return 0;
}
void mcd::forget_stream(mcd_handle *h, cudaStream_t stream) {
    if (h && h->last_stream_valid && h->last_stream == stream) h->last_stream_valid = false;
}
unsigned long long mcd::handle_generation(const mcd_handle *h) { return h ? h->generation : 0ull; }
unsigned long long mcd::next_fuse_nonce(mcd_handle *h) { return ++h->fuse_nonce; }

static int launch(mcd_handle *h, const double *theta_dev, int n_walkers, double *out_dev, int apply_prior,
                  cudaStream_t stream, bool exchange = false, const FuseParams *fuse = nullptr,
                  const unsigned long long *epoch_dev = nullptr, const ThetaBlock *inline_theta = nullptr,
                  const unsigned long long *host_seq_dev = nullptr) {
    if (!h) return fail(-1, "null handle");
    if (n_walkers < 0) return fail(-1, "n_walkers < 0");
    if (n_walkers == 0) return 0;
    if (!fuse) {
        if (!theta_dev && !inline_theta && h->desc.n_theta > 0) return fail(-1, "null theta");
        if (!out_dev) return fail(-1, "null out");
    }
    MCD_CUDA(cudaSetDevice(h->device));
    if (int rc = order_on_stream(h, stream)) return rc;
    LaunchParams p{};
    fill_params(h, p);
    p.n_walkers = n_walkers;
    p.apply_prior = apply_prior;
    p.theta = theta_dev;
    p.out = out_dev;
    choose_geometry(h, n_walkers, p);
    if (int rc = ensure_scratch(h, p)) return rc;
    p.partials = h->partials;
    p.partials2 = h->partials2;
    p.counters = h->counters;
    if (fuse) {
        p.fuse = *fuse;
        p.fuse.enabled = 1;
    }
    // a handle with an exchange attached is a star shard: every ensemble launch of the device sampler
    // (fused half-steps, the initial lnprob, the dry runs) is then a collective over the shards
    if (fuse && h->xchg_world > 1) exchange = true;
    if (exchange) {
        if (h->xchg_world < 2) return fail(-1, "mcd_exchange_attach has not been called on this handle");
        if (h->n_segments > 1) return fail(-1, "the fused cross-GPU reduction does not support segmented handles");
        if (n_walkers > h->xchg_capacity) return fail(-1, "n_walkers %d exceeds the exchange capacity %d", n_walkers, h->xchg_capacity);
        if (p.n_groups > kMaxXchgGroups) return fail(-1, "too many walker groups for the exchange buffer");
        p.xchg_world = h->xchg_world;
        p.xchg_rank = h->xchg_rank;
        p.xchg_capacity = h->xchg_capacity;
        // fused half-steps take their tag from the device-side step counter; host-buffer calls replayed as
        // a graph read it from device memory (epoch_dev, refreshed by the same copy that brings theta)
        p.xchg_epoch = (fuse || epoch_dev) ? 0 : ++h->xchg_epoch;
        p.xchg_epoch_ptr = epoch_dev;
        p.xchg_status = h->xchg_status;
        p.xchg_tagged_mode = h->xchg_tagged_mode;
        for (int r = 0; r < h->xchg_world; ++r) {
            p.xchg_data[r] = h->xchg_data[r];
            p.xchg_flags[r] = h->xchg_flags[r];
            p.xchg_words[r] = h->xchg_words[r];
        }
    }
    if (inline_theta) {
        // theta rides in the kernel arguments, the result goes into pinned host memory as tagged words
        p.theta = nullptr;
        p.host_words = reinterpret_cast<unsigned long long *>(h->out_pin);
        p.host_tag = (unsigned int)h->host_seq;
    }
    if (host_seq_dev) {
        // graph-replayed host call: the same, with the sequence number arriving in device memory beside theta
        p.host_words = reinterpret_cast<unsigned long long *>(h->out_pin);
        p.host_seq_ptr = host_seq_dev;
    }
    MCD_CUDA(launch_lnlike(h->var, p, stream, inline_theta));
    h->info.last_grid_x = p.n_chunks;
    h->info.last_grid_y = p.n_groups;
    h->info.last_block = kBlock;
    h->info.last_walker_tile = p.wl;
    h->info.launches += 1;
    return 0;
}

static int ensure_staging(mcd_handle *h, size_t theta_doubles, size_t out_doubles) {
    if (theta_doubles > h->theta_cap) {
        drop_host_graphs(h);
        cudaFree(h->theta_dev);
        cudaFreeHost(h->theta_pin);
        h->theta_dev = h->theta_pin = nullptr;
        h->theta_cap = 0;
        const size_t cap = std::max<size_t>(theta_doubles, 4096);
        MCD_CUDA(cudaMalloc(&h->theta_dev, sizeof(double) * cap));
        MCD_CUDA(cudaMallocHost(&h->theta_pin, sizeof(double) * cap));
        h->theta_cap = cap;
    }
    if (out_doubles > h->out_cap) {
        drop_host_graphs(h);
        cudaFree(h->out_dev);
        cudaFreeHost(h->out_pin);
        h->out_dev = h->out_pin = nullptr;
        h->out_cap = 0;
        const size_t cap = std::max<size_t>(out_doubles, 1024);
        MCD_CUDA(cudaMalloc(&h->out_dev, sizeof(double) * cap));
        // pinned: [0, 2 cap) tagged result words of the flagged paths, [2 cap, 3 cap) plain doubles of the copy-out path
        MCD_CUDA(cudaMallocHost(&h->out_pin, sizeof(double) * 3 * cap));
        memset(h->out_pin, 0, sizeof(double) * 3 * cap);
        h->out_cap = cap;
    }
    return 0;
}

// Small calls: ONE kernel launch, nothing else.  theta is copied into the kernel's argument block and the kernel
// writes every walker's result into pinned host memory as two self-validating words (half of the double plus the
// call's 32-bit tag each, one 16-byte store): no fence, no counter, no flag on the device.  The host polls the rows
// (a stream synchronisation costs more than the kernel on these sizes) and falls back to waiting on the stream
// when the kernel is a long one.
static unsigned long long next_host_seq(mcd_handle *h) {
    h->host_seq += 1;
    if ((unsigned int)h->host_seq == 0u) h->host_seq += 1;       // tag 0 is what a fresh buffer holds
    return h->host_seq;
}

// collect the rows of call `seq` from the tagged words into out_host
static int wait_for_words(mcd_handle *h, unsigned long long seq, size_t rows, double *out_host) {
    const volatile unsigned long long *w = reinterpret_cast<const volatile unsigned long long *>(h->out_pin);
    const unsigned long long tag = (unsigned long long)(unsigned int)seq;
    long long budget = 200000;                             // ~100 us of polling, then block on the stream
    bool synced = false;
    for (size_t i = 0; i < rows; ++i) {
        unsigned long long a, b;
        while (true) {
            a = w[2 * i];
            b = w[2 * i + 1];
            if ((a >> 32) == tag && (b >> 32) == tag) break;
            if (--budget > 0) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                continue;
            }
            if (synced) return fail(-2, "the likelihood kernel finished without publishing its result");
            MCD_CUDA(cudaStreamSynchronize(h->stream));
            synced = true;
            budget = 1000;
        }
        const unsigned long long bits = (a & 0xffffffffULL) | (b << 32);
        memcpy(&out_host[i], &bits, sizeof(double));
    }
    return 0;
}

static int host_call_inline(mcd_handle *h, const double *theta_host, int n_walkers, double *out_host, int apply_prior,
                            size_t rows, size_t nt) {
    if (int rc = ensure_staging(h, 1, rows)) return rc;
    ThetaBlock block;
    if (nt) memcpy(block.v, theta_host, sizeof(double) * nt);
    const unsigned long long seq = next_host_seq(h);
    if (int rc = launch(h, nullptr, n_walkers, h->out_dev, apply_prior, h->stream, false, nullptr, nullptr, &block)) return rc;
    return wait_for_words(h, seq, rows, out_host);
}

static int host_call(mcd_handle *h, const double *theta_host, int n_walkers, double *out_host, int apply_prior,
                     int exchange = 0) {
    if (!h) return fail(-1, "null handle");
    if (n_walkers < 0) return fail(-1, "n_walkers < 0");
    if (n_walkers == 0) return 0;
    if (!out_host || (!theta_host && h->desc.n_theta > 0)) return fail(-1, "null buffer");
    if (exchange && h->xchg_world < 2) return fail(-1, "mcd_exchange_attach has not been called on this handle");
    MCD_CUDA(cudaSetDevice(h->device));
    const size_t rows = (size_t)n_walkers * h->n_segments;        // theta is [segments][walkers][theta]
    const size_t nt = rows * h->desc.n_theta;
    if (!exchange && nt <= (size_t)kThetaInline && rows <= 4096) {
        const char *mode = getenv("MCD_HOST_CALL");               // "graph" / "sync": take the staged graph path (A/B, tests)
        if (!(mode && (mode[0] == 'g' || mode[0] == 's'))) {
            if (int rc = order_on_stream(h, h->stream)) return rc;
            return host_call_inline(h, theta_host, n_walkers, out_host, apply_prior, rows, nt);
        }
    }
    // Two words after theta travel to the device inside the same copy, so that the replayed graph (whose kernel
    // arguments are frozen) sees fresh values on every call: the exchange epoch (identical on all ranks) and
    // this handle's call sequence number, whose low 32 bits tag the result words the kernel writes straight into
    // pinned host memory.  The graph is copy-in -> kernel: no copy-out node, no stream synchronisation
    // (MCD_HOST_CALL=sync keeps both, for A/B).
    const char *mode = getenv("MCD_HOST_CALL");
    const bool flagged = !(mode && mode[0] == 's');
    if (int rc = ensure_staging(h, nt + 2, rows)) return rc;
    if (nt) memcpy(h->theta_pin, theta_host, sizeof(double) * nt);
    const unsigned long long *epoch_dev = nullptr, *seq_dev = nullptr;
    unsigned long long words[2] = {0ull, 0ull};
    if (exchange) {
        words[0] = ++h->xchg_epoch;
        epoch_dev = reinterpret_cast<const unsigned long long *>(h->theta_dev + nt);
    }
    if (flagged) {
        words[1] = next_host_seq(h);
        seq_dev = reinterpret_cast<const unsigned long long *>(h->theta_dev + nt + 1);
    }
    memcpy(h->theta_pin + nt, words, sizeof(words));
    const size_t n_copy = nt + 2;

    // A sampler calls with the same shape thousands of times: from the third call of a shape on, the
    // copy-in / kernel (/ copy-out) sequence is one graph launch (the first call sizes the scratch
    // buffers, the second captures).
    mcd_handle::HostGraph *slot = nullptr;
    const int kind = exchange | (flagged ? 2 : 0);
    for (auto &g : h->host_graphs)
        if (g.seen && g.n_walkers == n_walkers && g.apply_prior == apply_prior && g.exchange == kind) slot = &g;
    if (!slot) {
        for (auto &g : h->host_graphs)
            if (!g.seen && !slot) slot = &g;
        if (slot) {
            slot->n_walkers = n_walkers;
            slot->apply_prior = apply_prior;
            slot->exchange = kind;
        }
    }
    if (int rc = order_on_stream(h, h->stream)) return rc;
    auto enqueue = [&]() -> int {
        MCD_CUDA(cudaMemcpyAsync(h->theta_dev, h->theta_pin, sizeof(double) * n_copy, cudaMemcpyHostToDevice, h->stream));
        if (int rc = launch(h, h->theta_dev, n_walkers, h->out_dev, apply_prior, h->stream, exchange != 0, nullptr, epoch_dev,
                            nullptr, seq_dev))
            return rc;
        if (!flagged)
            MCD_CUDA(cudaMemcpyAsync(h->out_pin + 2 * h->out_cap, h->out_dev, sizeof(double) * rows, cudaMemcpyDeviceToHost, h->stream));
        return 0;
    };
    if (slot && slot->exec) {
        MCD_CUDA(cudaGraphLaunch(slot->exec, h->stream));
        h->info.launches += 1;
    } else if (slot && slot->seen == 1) {
        MCD_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        cudaGraph_t graph = nullptr;
        const cudaError_t end = cudaStreamEndCapture(h->stream, &graph);
        if (rc != 0 || end != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rc ? rc : fail(-2, "capturing the host-call graph failed: %s", cudaGetErrorString(end));
        }
        const cudaError_t inst = cudaGraphInstantiate(&slot->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (inst != cudaSuccess) {
            slot->exec = nullptr;
            return fail(-2, "instantiating the host-call graph failed: %s", cudaGetErrorString(inst));
        }
        MCD_CUDA(cudaGraphLaunch(slot->exec, h->stream));
    } else {
        if (int rc = enqueue()) return rc;
    }
    if (slot) slot->seen += slot->seen < 2 ? 1 : 0;
    if (flagged) {
        if (int rc = wait_for_words(h, words[1], rows, out_host)) return rc;
    } else {
        MCD_CUDA(cudaStreamSynchronize(h->stream));
        memcpy(out_host, h->out_pin + 2 * h->out_cap, sizeof(double) * rows);
    }
    if (exchange) {
        // a peer that never published turns the sums into NaN after the kernel's time limit: say so
        for (size_t i = 0; i < rows; ++i)
            if (out_host[i] != out_host[i]) {
                MCD_CUDA(cudaStreamSynchronize(h->stream));
                return exchange_status(h, h->stream);
            }
    }
    return 0;
}

// After the stream is synchronised: < 0 (with a message) if a kernel of this handle gave up waiting for a
// peer rank's shard sums since the last check.
int mcd::exchange_status(mcd_handle *h, cudaStream_t stream) {
    if (!h || !h->xchg_status) return 0;
    int status = 0;
    MCD_CUDA(cudaMemcpyAsync(&status, h->xchg_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MCD_CUDA(cudaStreamSynchronize(stream));
    if (status) {
        MCD_CUDA(cudaMemsetAsync(h->xchg_status, 0, sizeof(int), stream));
        return fail(-5, "fused cross-GPU reduction: a peer rank never published its shard sums (timed out); "
                        "the values of that call are NaN");
    }
    return 0;
}

int mcd::launch_ensemble(mcd_handle *h, const double *theta_dev, int n_walkers, double *out_dev, int apply_prior,
                         cudaStream_t stream) {
    return launch(h, theta_dev, n_walkers, out_dev, apply_prior, stream, h && h->xchg_world > 1);
}
int mcd::launch_ensemble_fused(mcd_handle *h, int n_walkers, const FuseParams &fuse, cudaStream_t stream) {
    return launch(h, nullptr, n_walkers, nullptr, 1, stream, false, &fuse);
}
// Self-validating tagged words save two L2 round trips per half-step but have every thread polling:
// measured on B200 they win while group^2 * walkers (words read per half-step) stays below ~10^5
// (C1 and the radial bins), and lose to counter + plain reads above (C2..C4).
static bool chain_tagged_exchange(int group, double walkers_per_exchange) {
    return (double)group * group * walkers_per_exchange <= 1.0e5;
}

int mcd::launch_resident_chain(mcd_handle *h, const ChainParams &chain, cudaStream_t stream) {
    if (!h) return fail(-1, "null handle");
    if (h->xchg_world > 1) return 1;         // star shards exchange sums every half-step: launch engine only
    if (chain.n_walkers > kChainMaxWalkers) return 1;
    // `group` CTAs (one per SM) hold one segment's stars in shared memory.  Take the resident kernel
    // only where its half-step is estimated to be shorter than a launch (~14 us of fixed latency + the
    // same arithmetic spread over all SMs), with the group size that minimises the estimate.  Cycle
    // model: FP64-pipe instructions ~ nominal flops per term, 64 lanes per SM, 60 % pipe efficiency,
    // 1.9 GHz; the exchange of slice sums ~1 us (tagged words) or ~2.2 us (counter, then reads) of L2
    // latency plus the traffic of every CTA reading every CTA's sums.  MCD_FORCE_RESIDENT_CHAIN=1 skips
    // the comparison with the launch engine, MCD_CHAIN_GROUP=g fixes the group size (tests, experiments).
    const char *force = getenv("MCD_FORCE_RESIDENT_CHAIN");
    const char *fixed = getenv("MCD_CHAIN_GROUP");
    const double flops = (double)variant_flops_per_term(h->var);
    const double ns = 0.5 * chain.n_walkers;
    const double per_sm = 64.0 * 1.9e9 * 0.6;
    const int sms = std::max(1, h->sm_count);
    const int max_group = h->n_segments <= sms ? std::min(sms / h->n_segments, (int)kChainBlock) : 1;
    int group = 0;
    long long per_cta = 0;
    size_t smem = 0;
    double resident = 1e30;
    for (int g = 1; g <= max_group; ++g) {
        if (fixed && fixed[0] && atoi(fixed) != g) continue;
        const long long per = (((h->max_segment + g - 1) / g) + 1) & ~1LL;
        const size_t bytes = chain_shared_bytes(h->var, per, chain.n_walkers, h->desc.n_theta);
        if (bytes == 0) continue;
        const double wl = std::min(ns, (double)kChainBlock);
        double t = 0.8e-6 + ns * (double)per * flops / per_sm;
        if (g > 1) t += (chain_tagged_exchange(g, wl) ? 1.0e-6 : 2.2e-6) + (double)g * g * wl * 8.0 / 6e12;
        else t *= std::ceil((double)h->n_segments / sms);
        if (t < resident) {
            resident = t;
            group = g;
            per_cta = per;
            smem = bytes;
        }
    }
    if (group == 0) return 1;                // does not fit the shared memory of the SMs it may use
    const double launched = 14e-6 + ns * (double)h->n * flops / (per_sm * sms);
    if (!(force && force[0] == '1') && resident > launched) return 1;

    MCD_CUDA(cudaSetDevice(h->device));
    if (int rc = order_on_stream(h, stream)) return rc;
    LaunchParams p{};
    fill_params(h, p);
    p.apply_prior = 1;
    ChainParams c = chain;
    c.max_segment_padded = (int)(((per_cta + 15) / 16) * 16);
    c.group = group;
    c.stars_per_cta = (int)per_cta;
    if (group > 1) {
        const int larger_half = chain.n_walkers - chain.n0 > chain.n0 ? chain.n_walkers - chain.n0 : chain.n0;
        c.sum_stride = std::min((int)kChainBlock, larger_half);
        const size_t slots = (size_t)2 * h->n_segments * group * c.sum_stride;
        const size_t header = 16 + (((size_t)h->n_segments * 8 + 15) & ~(size_t)15);     // status word, arrival counters
        const size_t need = header + slots * 16;
        if (need > h->group_sums_cap) {
            if (h->group_sums) cudaFree(h->group_sums);
            h->group_sums = nullptr;
            h->group_sums_cap = 0;
            MCD_CUDA(cudaMalloc(&h->group_sums, need));
            h->group_sums_cap = need;
        }
        MCD_CUDA(cudaMemsetAsync(h->group_sums, 0, need, stream));      // tag 0 = nothing published yet
        c.status = reinterpret_cast<int *>(h->group_sums);
        c.group_arrivals = reinterpret_cast<unsigned long long *>(h->group_sums + 16);
        c.group_sums = reinterpret_cast<TaggedSum *>(h->group_sums + header);
        c.tagged = chain_tagged_exchange(group, c.sum_stride) ? 1 : 0;
        if (const char *mode = getenv("MCD_CHAIN_EXCHANGE"))      // "tagged" / "counter": experiments, tests
            c.tagged = mode[0] == 't' ? 1 : (mode[0] == 'c' ? 0 : c.tagged);
    }
    const cudaError_t err = launch_chain(h->var, p, c, smem, stream);
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        if (group > 1) return 1;             // e.g. no cooperative launch in this context: launch engine instead
        return fail(-2, "resident chain launch: %s", cudaGetErrorString(err));
    }
    h->info.launches += 1;
    h->chain_group = group;
    return 0;
}
int mcd::resident_chain_group(const mcd_handle *h) { return h ? h->chain_group : 0; }
int mcd::resident_chain_status(mcd_handle *h, cudaStream_t stream) {
    if (!h || !h->group_sums || h->chain_group <= 1) return 0;
    int status = 0;
    MCD_CUDA(cudaMemcpyAsync(&status, h->group_sums, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MCD_CUDA(cudaStreamSynchronize(stream));
    if (status) return fail(-2, "resident chain: a CTA of a group never published its sums");
    return 0;
}
int mcd::handle_device(const mcd_handle *h) { return h->device; }

extern "C" int mcd_lnlike(mcd_handle *h, const double *theta_host, int32_t n_walkers, double *out_host) {
    return host_call(h, theta_host, n_walkers, out_host, 0);
}
extern "C" int mcd_lnprob(mcd_handle *h, const double *theta_host, int32_t n_walkers, double *out_host) {
    return host_call(h, theta_host, n_walkers, out_host, 1);
}
extern "C" int mcd_lnlike_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev, void *stream) {
    return launch(h, theta_dev, n_walkers, out_dev, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int mcd_lnprob_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev, void *stream) {
    return launch(h, theta_dev, n_walkers, out_dev, 1, static_cast<cudaStream_t>(stream));
}
extern "C" int mcd_lnprob_partial_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev,
                                         void *stream) {
    // a shard's lnprob kernel already returns (sum over its stars) or -inf: additive across shards
    return launch(h, theta_dev, n_walkers, out_dev, 1, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------
// fused cross-GPU reduction over symmetric (peer-mapped) memory
// ------------------------------------------------------------------------------------------
static size_t exchange_flag_bytes(int world) { return sizeof(unsigned long long) * kXchgSlots * world * kMaxXchgGroups; }

extern "C" int mcd_exchange_bytes(int32_t world, int32_t max_walkers, int64_t *bytes_out) {
    if (world < 2 || world > kMaxRanks || max_walkers < 1 || !bytes_out) return fail(-1, "bad exchange geometry");
    // flags | plain sums (flag protocol) | tagged words, two per sum (tagged protocol)
    *bytes_out = (int64_t)(exchange_flag_bytes(world) + 3 * sizeof(double) * kXchgSlots * world * (size_t)max_walkers);
    return 0;
}

extern "C" int mcd_exchange_attach(mcd_handle *h, int32_t rank, int32_t world, const uint64_t *peer_buffers,
                                   int32_t max_walkers) {
    if (!h || !peer_buffers) return fail(-1, "null argument");
    if (world < 2 || world > kMaxRanks || rank < 0 || rank >= world || max_walkers < 1)
        return fail(-1, "bad exchange geometry (world %d, rank %d)", world, rank);
    MCD_CUDA(cudaSetDevice(h->device));
    drop_host_graphs(h);                 // bumps the generation: exchange pointers are baked into captured graphs
    if (!h->xchg_status) {
        MCD_CUDA(cudaMalloc(&h->xchg_status, sizeof(int)));
        MCD_CUDA(cudaMemset(h->xchg_status, 0, sizeof(int)));
    }
    h->xchg_world = world;
    h->xchg_rank = rank;
    h->xchg_capacity = max_walkers;
    h->xchg_epoch = 0;
    for (int r = 0; r < world; ++r) {
        if (!peer_buffers[r]) return fail(-1, "null peer buffer");
        char *base = reinterpret_cast<char *>(static_cast<uintptr_t>(peer_buffers[r]));
        h->xchg_flags[r] = reinterpret_cast<unsigned long long *>(base);
        h->xchg_data[r] = reinterpret_cast<double *>(base + exchange_flag_bytes(world));
        h->xchg_words[r] = reinterpret_cast<unsigned long long *>(
            base + exchange_flag_bytes(world) + sizeof(double) * kXchgSlots * world * (size_t)max_walkers);
    }
    const char *mode = getenv("MCD_XCHG");
    h->xchg_tagged_mode = (mode && mode[0] == 'f') ? 0 : 1;
    return 0;
}

extern "C" int mcd_lnprob_allreduce_device(mcd_handle *h, const double *theta_dev, int32_t n_walkers, double *out_dev,
                                           void *stream) {
    return launch(h, theta_dev, n_walkers, out_dev, 1, static_cast<cudaStream_t>(stream), true);
}

extern "C" int mcd_lnprob_allreduce(mcd_handle *h, const double *theta_host, int32_t n_walkers, double *out_host) {
    return host_call(h, theta_host, n_walkers, out_host, 1, 1);
}

extern "C" int mcd_exchange_status(mcd_handle *h) {
    if (!h) return fail(-1, "null handle");
    MCD_CUDA(cudaSetDevice(h->device));
    return exchange_status(h, h->stream);
}

// ------------------------------------------------------------------------------------------
// per-star lnlike (no_sum)
// ------------------------------------------------------------------------------------------
static int per_star(mcd_handle *h, const double *theta_dev, double *out_dev, int mode, cudaStream_t stream,
                    double *out2_dev = nullptr) {
    const int membership = mode == kPerStarMembership;
    LaunchParams p{};
    fill_params(h, p);
    p.n_walkers = 1;
    p.theta = theta_dev;
    // the per-star kernel evaluates the reference's formulas literally and wants lbg itself where
    // the FAST packing keeps a mantissa
    const int nb = variant_columns(h->var) - (h->var.background == MCD_BG_NONE ? 0 : (h->var.background == MCD_BG_GAUSSIAN ? 1 : 2));
    if (h->var.background == MCD_BG_FIXED_PMEMBER || h->var.background == MCD_BG_FIXED_DENSITY) p.cols[nb + 1] = h->raw[RAW_LBG];
    // ... and the velocity itself where the FAST mixture packing keeps it in units of 1 / kExpArgScale
    p.verr2_unscale = 1.0;
    if (h->var.background != MCD_BG_NONE && h->var.math_mode == MCD_MATH_FAST) {
        p.cols[nb - 2] = h->raw[RAW_V];
        p.verr2_unscale = 1.0 / mix_var_scale();      // ... and verr^2 times kMixVarScale
    }
    if (membership && h->var.background == MCD_BG_NONE) return fail(-1, "membership probabilities need a background component");
    if (h->n_segments > 1) return fail(-1, "per-star entry points are not available for segmented handles");
    if (!theta_dev && h->desc.n_theta > 0) return fail(-1, "null theta");
    MCD_CUDA(launch_per_star(h->var, p, out_dev, out2_dev, mode, stream));
    h->info.launches += 1;
    return 0;
}

extern "C" int mcd_lnlike_per_star_device(mcd_handle *h, const double *theta_dev, double *out_dev, void *stream) {
    if (!h || !out_dev) return fail(-1, "null argument");
    MCD_CUDA(cudaSetDevice(h->device));
    return per_star(h, theta_dev, out_dev, kPerStarLnlike, static_cast<cudaStream_t>(stream));
}

extern "C" int mcd_membership_per_star_device(mcd_handle *h, const double *theta_dev, double *out_dev, void *stream) {
    if (!h || !out_dev) return fail(-1, "null argument");
    MCD_CUDA(cudaSetDevice(h->device));
    return per_star(h, theta_dev, out_dev, kPerStarMembership, static_cast<cudaStream_t>(stream));
}

extern "C" int mcd_model_per_star_device(mcd_handle *h, const double *theta_dev, double *v_los_dev, double *sigma_los_dev,
                                         void *stream) {
    if (!h) return fail(-1, "null argument");
    if (!v_los_dev && !sigma_los_dev) return 0;
    MCD_CUDA(cudaSetDevice(h->device));
    return per_star(h, theta_dev, v_los_dev, kPerStarModel, static_cast<cudaStream_t>(stream), sigma_los_dev);
}

static int per_star_host(mcd_handle *h, const double *theta_host, double *out_host, int mode, double *out2_host = nullptr) {
    if (!h || (!out_host && !out2_host)) return fail(-1, "null argument");
    if (!theta_host && h->desc.n_theta > 0) return fail(-1, "null theta");
    MCD_CUDA(cudaSetDevice(h->device));
    if (h->n == 0) return 0;
    if (int rc = ensure_staging(h, (size_t)h->desc.n_theta, 1)) return rc;
    if (!h->star_dev) MCD_CUDA(cudaMalloc(&h->star_dev, sizeof(double) * 2 * h->n));      // [n] values (+ [n] second curve)
    if (h->desc.n_theta) {
        memcpy(h->theta_pin, theta_host, sizeof(double) * h->desc.n_theta);
        MCD_CUDA(cudaMemcpyAsync(h->theta_dev, h->theta_pin, sizeof(double) * h->desc.n_theta, cudaMemcpyHostToDevice, h->stream));
    }
    double *first = out_host ? h->star_dev : nullptr, *second = out2_host ? h->star_dev + h->n : nullptr;
    if (int rc = per_star(h, h->theta_dev, first, mode, h->stream, second)) return rc;
    if (out_host) MCD_CUDA(cudaMemcpyAsync(out_host, first, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
    if (out2_host) MCD_CUDA(cudaMemcpyAsync(out2_host, second, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
    MCD_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// Runner._calculate_lnlike(v_los, sigma_los) for curves computed by the caller (analysis/runner.py:240-286)
extern "C" int mcd_calculate_lnlike(mcd_handle *h, const double *v_los_host, const double *sigma_los_host, double *out_host) {
    if (!h || !out_host) return fail(-1, "null argument");
    if (h->n > 0 && (!v_los_host || !sigma_los_host)) return fail(-1, "null model curve");
    if (h->n_segments > 1) return fail(-1, "mcd_calculate_lnlike is not available for segmented handles");
    const int bg = h->var.background;
    if (bg == MCD_BG_FIXED_DENSITY || bg == MCD_BG_GAUSSIAN)
        return fail(-1, "the classes with a fitted background fraction do not use _calculate_lnlike (model.py:565-623, "
                        "constant.py:326-364)");
    MCD_CUDA(cudaSetDevice(h->device));
    if (h->n == 0) {            // np.sum over nothing
        *out_host = 0.0;
        return 0;
    }
    if (int rc = order_on_stream(h, h->stream)) return rc;
    const int blocks = curve_lnlike_blocks(h->n, h->sm_count);
    double *buf = nullptr;      // [n] v_los | [n] sigma_los | [blocks] partial sums | result
    const size_t count = 2 * (size_t)h->n + (size_t)blocks + 1;
    MCD_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&buf), sizeof(double) * count, h->stream));
    double *v_los = buf, *sigma_los = buf + h->n, *partial = buf + 2 * h->n, *out = partial + blocks;
    cudaError_t err = cudaMemcpyAsync(v_los, v_los_host, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(sigma_los, sigma_los_host, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream);
    const bool mixture = bg == MCD_BG_FIXED_PMEMBER;
    if (err == cudaSuccess)
        err = launch_curve_lnlike(h->raw[RAW_V], h->raw[RAW_VERR], mixture ? h->raw[RAW_PMEMBER] : nullptr,
                                  mixture ? h->raw[RAW_LBG] : nullptr, v_los, sigma_los, h->n, partial, out, h->sm_count, h->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out_host, out, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    cudaFreeAsync(buf, h->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
    if (err != cudaSuccess) return fail(-2, "mcd_calculate_lnlike: %s", cudaGetErrorString(err));
    h->info.launches += 2;
    return 0;
}

extern "C" int mcd_lnlike_per_star(mcd_handle *h, const double *theta_host, double *out_host) {
    if (!out_host) return fail(-1, "null argument");
    return per_star_host(h, theta_host, out_host, kPerStarLnlike);
}
extern "C" int mcd_membership_per_star(mcd_handle *h, const double *theta_host, double *out_host) {
    if (!out_host) return fail(-1, "null argument");
    return per_star_host(h, theta_host, out_host, kPerStarMembership);
}
extern "C" int mcd_model_per_star(mcd_handle *h, const double *theta_host, double *v_los_host, double *sigma_los_host) {
    if (h && !v_los_host && !sigma_los_host) return 0;
    return per_star_host(h, theta_host, v_los_host, kPerStarModel, sigma_los_host);
}

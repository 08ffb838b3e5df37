// Likelihood kernels of the B200-native mcmc-dynamics hot path (sm_100a).
//
// One launch evaluates a whole (half-)ensemble: every walker of the call against every star of
// the catalogue shard held by this GPU, reduced to one float64 per walker.  It replaces the
// per-walker Python call chain of the reference
//   Runner.lnprob -> lnprior + <Model>.lnlike -> rotation_model / dispersion_model
//                 -> calc_xy_offset -> Runner._calculate_lnlike
//   (analysis/runner.py:182-306, constant.py:52-154,293-364, model.py:93-223,391-456,565-623,
//    utils/coordinates/calc_xy_offset.py:9-33)
// and the walker-parallel process pool around it (analysis/runner.py:398-403).
//
// Work decomposition
//   grid  = (star chunks x walker groups, segments); a CTA owns `wl` walkers x `slices` star slices
//           (wl * slices <= 256 threads) and a contiguous run of star tiles.
//   tile  = `tile` stars of every packed column, brought into shared memory by TMA bulk copies
//           (cp.async.bulk, mbarrier completion), double buffered; every thread of the CTA then
//           reads the stars of its slice as shared-memory broadcasts.
//   thread= one walker: its derived constants live in registers for the whole launch, its
//           partial sums are register accumulators (no shuffles in the star loop).
//   reduce= slices -> shared memory, chunks -> `partials` in global memory, then two ticketed
//           levels (last CTA of a super-chunk, last super-chunk of a walker group) that add in
//           index order, so the result does not depend on CTA scheduling.
#include <math.h>
#include <stdio.h>

#include <algorithm>

#include "mcd_internal.h"
#include "mcd_math.cuh"
#include "mcd_rng.cuh"

// tuning knobs (overridable at compile time for A/B runs: -DMCD_MIN_BLOCKS=.. -DMCD_PAIRS=..)
// Measured on the headline workload (gpurun_out/ab*.log): 6.27e11 terms/s with
// (1 pair, 4 CTA/SM), 6.44e11 with (2 pairs, 3 CTA/SM); more occupancy or ILP changes nothing
// further -- the kernel sits on the FP64 pipe's operand-bandwidth limit.
#ifndef MCD_MIN_BLOCKS
#define MCD_MIN_BLOCKS 3      // __launch_bounds__ minimum resident CTAs per SM, no-background variants
#endif
#ifndef MCD_PAIRS
#define MCD_PAIRS 2           // star pairs per inner-loop iteration (1 or 2), no-background variants
#endif
#ifndef MCD_BG_MIN_BLOCKS
#define MCD_BG_MIN_BLOCKS 2   // the same two knobs for the background-mixture variants
#endif
#ifndef MCD_GROUP4
#define MCD_GROUP4 1          // no-background variants: exponent of the variance product folded every four stars (0: two)
#endif
#ifndef MCD_FLAG_LATE
#define MCD_FLAG_LATE 1       // fixed-background mixtures: per-star fast-path flags checked after the fast evaluation
#endif
#ifndef MCD_BG_PAIRS
#define MCD_BG_PAIRS 2        // round 2 (lean arithmetic): C3 45.1 -> 43.1 us, mixgb 4286 -> 4159 us, mix +0.9 % (r02_ab_runs.md)
#endif

// This file is compiled three times (see __graft_entry__.py), so that the template instantiations
// build in parallel:  MCD_TU_PART 0 = pack kernel, variant tables and the dispatch front ends,
// 1 = every FAST-arithmetic kernel, 2 = every PLAIN-arithmetic kernel (+ the per-star kernel).
#ifndef MCD_TU_PART
#error "compile with -DMCD_TU_PART=0|1|2"
#endif

namespace mcd {

// ------------------------------------------------------------------------------------------
// packed column layout
// ------------------------------------------------------------------------------------------
//   (FAST mixture variants store v * kExpArgScale instead of v)
//   free centre            : p1 = cos(dec) sin(ra-ra0), p2 = cos(dec) cos(ra-ra0), sin(dec), v, verr^2
//   fixed centre, constant : cos(theta_i), sin(theta_i), v, verr^2
//   fixed centre, radial   : x, y [arcmin], r^2, v, verr^2
//   + FIXED_PMEMBER        : pmember, then FAST: mantissa of (1-p) sqrt(2pi) exp(lbg) (+ int32 exponent)
//                                           PLAIN: lbg
//   + FIXED_DENSITY        : density, then FAST: mantissa of sqrt(2pi) exp(lbg) (+ int32 exponent)
//                                           PLAIN: lbg
//   + GAUSSIAN             : density
__host__ __device__ constexpr int base_columns(int rot, int free_centre) {
    return free_centre ? 5 : (rot == MCD_ROT_RADIAL ? 5 : 4);
}
__host__ __device__ constexpr int total_columns(int rot, int free_centre, int bg) {
    return base_columns(rot, free_centre) +
           (bg == MCD_BG_NONE ? 0 : (bg == MCD_BG_GAUSSIAN ? 1 : 2));
}
__host__ __device__ constexpr bool has_icol(int bg, int math) {
    return math == MCD_MATH_FAST && (bg == MCD_BG_FIXED_PMEMBER || bg == MCD_BG_FIXED_DENSITY);
}

#if MCD_TU_PART == 0
int variant_columns(const Variant &v) { return total_columns(v.rotation, v.free_centre, v.background); }
bool variant_has_icol(const Variant &v) { return has_icol(v.background, v.math_mode); }

// Nominal FP64 operations per (walker, star) term: add/sub/mul/compare = 1, fma = 2, every
// div / sqrt / rsqrt / log / exp = 1 (DESIGN.md "work per term").
int variant_flops_per_term(const Variant &v) {
    int f;
    if (v.rotation == MCD_ROT_CONSTANT) f = v.free_centre ? 29 : 11;
    else f = v.free_centre ? 35 : 19;
    if (v.background == MCD_BG_FIXED_PMEMBER || v.background == MCD_BG_FIXED_DENSITY) f += 12;
    if (v.background == MCD_BG_FIXED_DENSITY) f += 3;
    if (v.background == MCD_BG_GAUSSIAN) f += 23;
    return f;
}

// ------------------------------------------------------------------------------------------
// pack: raw catalogue columns -> packed columns (one thread per star, once per model object)
// ------------------------------------------------------------------------------------------
__global__ void pack_kernel(const PackParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_stars) return;
    const double ra = P.raw.ra[i], dec = P.raw.dec[i];
    const double v = P.raw.v[i], verr = P.raw.verr[i];
    const double e2 = verr * verr;
    // output position: identity, or the star's slot inside its 16-aligned segment
    long long o = i;
    if (P.n_segments > 1) {
        int lo = 0, hi = P.n_segments;          // seg_begin[lo] <= i < seg_begin[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) / 2;
            if (P.seg_begin[mid] <= i) lo = mid; else hi = mid;
        }
        o = P.seg_packed[lo] + (i - P.seg_begin[lo]);
    }
    int c = 0;
    if (P.free_centre) {
        double sa, ca, sd, cd;
        sincos((ra - P.ra0_deg) * kDeg2Rad, &sa, &ca);
        sincos(dec * kDeg2Rad, &sd, &cd);
        P.cols[c++][o] = cd * sa;
        P.cols[c++][o] = cd * ca;
        P.cols[c++][o] = sd;
    } else {
        // utils/coordinates/calc_xy_offset.py:30-31, arcmin
        double sda, cda, sd, cd, sdc, cdc;
        sincos((ra - P.ra_c_deg) * kDeg2Rad, &sda, &cda);
        sincos(dec * kDeg2Rad, &sd, &cd);
        sincos(P.dec_c_deg * kDeg2Rad, &sdc, &cdc);
        const double dx = -kR0Arcmin * cd * sda;
        const double dy = kR0Arcmin * (sd * cdc - cd * sdc * cda);
        if (P.rotation == MCD_ROT_CONSTANT) {
            // theta_i = atan2(dy, dx) (constant.py:107).  A star exactly at the centre has
            // dx = -r0 cos(dec) sin(+0) = -0.0 and dy = +0.0, i.e. theta_i = atan2(+0, -0) = pi.
            const double r = sqrt(dx * dx + dy * dy);
            P.cols[c++][o] = r > 0.0 ? dx / r : -1.0;
            P.cols[c++][o] = r > 0.0 ? dy / r : 0.0;
        } else {
            P.cols[c++][o] = dx;
            P.cols[c++][o] = dy;
            P.cols[c++][o] = dx * dx + dy * dy;
        }
    }
    // FAST mixture kernels take the velocity in units of 1 / kExpArgScale (see `term`)
    const bool scaled = P.math_mode == MCD_MATH_FAST && P.background != MCD_BG_NONE;
    P.cols[c++][o] = scaled ? v * kExpArgScale : v;
    // ... and verr^2 times kMixVarScale (a power of two: exact)
    P.cols[c++][o] = scaled ? e2 * kMixVarScale : e2;
    if (P.background == MCD_BG_FIXED_PMEMBER || P.background == MCD_BG_FIXED_DENSITY) {
        const double w = P.background == MCD_BG_FIXED_PMEMBER ? P.raw.pmember[i] : P.raw.density[i];
        const double lbg = P.raw.lbg[i];
        P.cols[c++][o] = w;
        if (P.math_mode == MCD_MATH_FAST) {
            double m;
            int e, invalid = 0;
            exp_split(lbg, m, e, invalid);
            const double wb = P.background == MCD_BG_FIXED_PMEMBER ? (1.0 - w) : 1.0;
            const double mant = invalid ? __longlong_as_double(0x7ff8000000000000LL) : wb * kSqrt2Pi * m;
            // Background term B = mant * 2^e.  Where it is a comfortable plain double (the rule, not the
            // exception: |log2 B| <= kMixComfort means lbg within ~ +-140 nats) the star takes the kernel's
            // fast path -- B is stored as it is and the exponent column carries kMixFastFlag; otherwise
            // (p = 1, a star hundreds of sigma from every background velocity, invalid input) the
            // (mantissa, exponent) pair is kept for the extended-range path.
            const int mexp = ((__double2hiint(mant) >> 20) & 0x7ff) - 1023;
            const bool comfortable = !invalid && mant > 0.0 && mexp > -1000 && abs(mexp + e) <= kMixComfort && abs(e) < 2000;
            P.cols[c++][o] = comfortable ? ldexp(mant, e) : mant;
            P.icol[o] = comfortable ? kMixFastFlag : e;
        } else {
            P.cols[c++][o] = lbg;
        }
    } else if (P.background == MCD_BG_GAUSSIAN) {
        P.cols[c++][o] = P.raw.density[i];
    }
}

double mix_var_scale() { return kMixVarScale; }

cudaError_t launch_pack(const PackParams &p, cudaStream_t stream) {
    if (p.n_stars <= 0) return cudaSuccess;
    const int block = 256;
    const long long grid = (p.n_stars + block - 1) / block;
    pack_kernel<<<(unsigned)grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}
#endif  // MCD_TU_PART == 0

#if MCD_TU_PART != 0

// ------------------------------------------------------------------------------------------
// per-walker derived constants
// ------------------------------------------------------------------------------------------
struct Walker {
    double vsys, s2;            // systemic velocity, sigma_max^2
    double s2h;                 // sigma_max^2 / 2 (pairs with rsqrt_twice in the FAST radial variants)
    double cx, cy;              // rotation: v_rot numerator = x*cx + y*cy
    double ip2, ia2;            // 1/r_peak^2, 1/a^2 in units of the stored coordinates
    double sb, cb, cdc;         // free centre: sin/cos(ra_c - ra0), cos(dec_c)
    double sbs, cbs;            //              -sin(ra_c - ra0) sin(dec_c), -cos(ra_c - ra0) sin(dec_c)
    double vb, sb2, fb;         // background: v_back, sigma_back^2, f_back
    int prior_ok;
    int slow;                   // FAST mixtures: every term of this walker takes the extended-range path
    // FAST mixture kernels (SCALED): vsys, cx, cy, vb are multiplied by kExpArgScale, so that residuals come
    // out as u = z * kExpArgScale, the argument of exp_neg_sq_split; s2, s2h, sb2 by kMixVarScale like the
    // packed verr^2 column
};

// stretch-move proposal of active walker k of segment `seg` (emcee RedBlueMove/StretchMove):
// z = ((a-1) u + 1)^2 / a, q = c_j - (c_j - s) z with c_j drawn uniformly from the other half.
// Pure function of (seed, step, half, walker): every CTA of the walker and the accepting CTA
// recompute the same q.  Returns the global walker index (row of pos / lnp) through `row_index`.
__device__ __forceinline__ double draw_proposal(const LaunchParams &P, int seg, int k, double *q, int &row_index) {
    const FuseParams &F = P.fuse;
    const int ns = P.n_walkers, nc = F.walkers_total - ns;
    const int *perm = F.perm + (size_t)seg * F.walkers_total;
    const int *active = perm + (F.half == 0 ? 0 : F.n0);
    const int *other = perm + (F.half == 0 ? F.n0 : 0);
    double u0, u1;
    uniforms(F.seed, F.step[0], (uint32_t)F.half, (uint32_t)(seg * F.walkers_total + k), 0u, u0, u1);
    const double t = (F.a - 1.0) * u0 + 1.0;
    const double z = t * t / F.a;
    int j = (int)(u1 * nc);
    j = j >= nc ? nc - 1 : j;
    row_index = seg * F.walkers_total + active[k];
    const double *s = F.pos + (size_t)row_index * P.n_theta;
    const double *c = F.pos + ((size_t)seg * F.walkers_total + other[j]) * P.n_theta;
    for (int p = 0; p < P.n_theta; ++p) q[p] = c[p] - (c[p] - s[p]) * z;
    return z;
}

template <int ROT, int FREE, int BG, bool SCALED = false>
__device__ __forceinline__ void load_walker(const LaunchParams &P, const double *row, Walker &W) {
    double par[MCD_NPARAM];
#pragma unroll
    for (int k = 0; k < MCD_NPARAM; ++k) {
        const int s = P.slot[k];
        par[k] = s >= 0 ? row[s] * P.scale[k] : P.fixed_scaled[k];
    }
    int ok = P.fixed_prior_ok;
    // box prior, bounds inclusive (parameter.py:691-692); NaN compares false and passes, as it
    // does in the reference
    for (int j = 0; j < P.n_theta; ++j) {
        const double t = row[j];
        ok &= !(t < P.lower[j] || t > P.upper[j]);
    }
    W.prior_ok = ok;
    W.vsys = par[MCD_P_V_SYS];
    W.s2 = par[MCD_P_SIGMA_MAX] * par[MCD_P_SIGMA_MAX];
    W.s2h = 0.5 * W.s2;
    const double vmx = par[MCD_P_V_MAXX], vmy = par[MCD_P_V_MAXY];
    if constexpr (ROT == MCD_ROT_CONSTANT) {
        // v_max sin(theta_i - theta_0) = sin(theta_i) v_maxx - cos(theta_i) v_maxy (constant.py:109-111)
        W.cx = -vmy;
        W.cy = vmx;
        W.ip2 = 0.0;
        W.ia2 = 0.0;
    } else {
        // coordinates are stored in arcmin (fixed centre) or in units of r0 = 10800/pi arcmin (free)
        const double L = FREE ? kR0Arcmin : 1.0;
        const double rp = par[MCD_P_R_PEAK], a = par[MCD_P_A];
        W.cx = -2.0 * L * vmy / rp;
        W.cy = 2.0 * L * vmx / rp;
        W.ip2 = (L / rp) * (L / rp);
        W.ia2 = (L / a) * (L / a);
    }
    if constexpr (FREE) {
        double sdc;
        sincos((par[MCD_P_RA_CENTER] - P.ra0_deg) * kDeg2Rad, &W.sb, &W.cb);
        sincos(par[MCD_P_DEC_CENTER] * kDeg2Rad, &sdc, &W.cdc);
        W.sbs = -W.sb * sdc;
        W.cbs = -W.cb * sdc;
    } else {
        W.sb = 0.0; W.cb = 1.0; W.cdc = 1.0; W.sbs = 0.0; W.cbs = 0.0;
    }
    W.vb = par[MCD_P_V_BACK];
    W.sb2 = par[MCD_P_SIGMA_BACK] * par[MCD_P_SIGMA_BACK];
    W.fb = par[MCD_P_F_BACK];
    W.slow = 0;
    if constexpr (SCALED) {
        W.vsys *= kExpArgScale;
        W.cx *= kExpArgScale;
        W.cy *= kExpArgScale;
        W.vb *= kExpArgScale;
        W.s2 *= kMixVarScale;
        W.s2h *= kMixVarScale;
        W.sb2 *= kMixVarScale;
        // opaque to the optimiser: it would otherwise redo these multiplications inside the star loop
        // (v - v_sys * c as an FMA with the constant rebuilt in uniform registers per pair)
        asm volatile("" : "+d"(W.vsys), "+d"(W.vb));
        // The fast path keeps the per-star factor as a plain double and renormalises the running product
        // every four stars: it needs the background weight f_back inside a sane range.  A walker with a
        // (nearly) vanishing or non-finite f_back evaluates all its terms in extended range instead.
        if constexpr (BG == MCD_BG_FIXED_DENSITY || BG == MCD_BG_GAUSSIAN) W.slow = !(W.fb >= 1e-9 && W.fb <= 1e9);
    }
}

// ------------------------------------------------------------------------------------------
// per-thread accumulators
// ------------------------------------------------------------------------------------------
template <int BG, int MATH>
struct Accum;

template <>
struct Accum<MCD_BG_NONE, MCD_MATH_FAST> {
    double chi;
    LogProduct norm;
    __device__ __forceinline__ void reset() { chi = 0.0; norm.reset(); }
    // called after every group of at most four factors multiplied in with mul_raw(): variances within
    // 2^+-250 (1e-75 .. 1e75 km^2/s^2) keep the product of four a normal number; anything else raises `bad`
    __device__ __forceinline__ void end_group() { norm.renormalise_light(); }
    __device__ __forceinline__ void end_tile() { norm.fold(); }
    __device__ __forceinline__ double value() {
        norm.fold();
        return -0.5 * (chi + norm.ln());
    }
};
template <int BG>
struct Accum<BG, MCD_MATH_FAST> {
    LogProduct num, den;
    int dead;            // a star's likelihood is exactly zero (extended-range path): lnlike = -inf
    __device__ __forceinline__ void reset() { num.reset(); den.reset(); dead = 0; }
    // Called after at most four stars.  Fast-path factors lie within 2^+-(kMixComfort + a few), so the
    // product of four stays a normal number; anything else (NaN, inf, zero, negative) raises `bad`.
    // (renormalise_light() is 1.2 - 1.6 % slower here: these kernels are latency-, not issue-bound, and the two
    // extra accumulator registers cost more than the five integer instructions, profiles/r02_ab_runs.md)
    __device__ __forceinline__ void end_group() {
        num.renormalise_checked();
        if (BG != MCD_BG_FIXED_PMEMBER) den.renormalise_checked();
    }
    __device__ __forceinline__ void end_tile() {}
    // one factor m * 2^e from the extended-range path
    __device__ __forceinline__ void mul_slow(double m, int e) {
        if (m > 0.0) {
            num.mul_ext(m, e);
            num.renormalise_ext();
        } else if (m == 0.0) {
            dead = 1;
        } else {
            num.bad = 1;        // negative weight or NaN
        }
    }
    __device__ __forceinline__ double value() {
        const double n = num.ln();
        const double r = BG == MCD_BG_FIXED_PMEMBER ? n : n - den.ln();
        return (dead && r == r) ? __longlong_as_double(0xfff0000000000000LL) : r;
    }
};
template <int BG>
struct Accum<BG, MCD_MATH_PLAIN> {
    double sum;
    double pmember;     // a-posteriori membership of the last star (per-star kernel only; dead elsewhere)
    double vlos, sig2;  // model velocity and squared model dispersion of the last star (per-star kernel only)
    __device__ __forceinline__ void reset() { sum = 0.0; pmember = 1.0; vlos = 0.0; sig2 = 0.0; }
    __device__ __forceinline__ void end_group() {}
    __device__ __forceinline__ void end_tile() {}
    __device__ __forceinline__ double value() { return sum; }
};

// The packed values of one star, in registers.
template <int NC>
struct Star {
    double c[NC];
    int e;        // exponent column (FAST fixed-background variants)
};

// Extended-range evaluation of one mixture factor (times sqrt(2 pi)) as mantissa * 2^exponent: the
// exponentials come as (mantissa, exponent) pairs without the clamp of the fast path and the two
// components are added in that form, so nothing can underflow -- the analogue of the reference's
// max-shifted log-sum-exp (analysis/runner.py:279-284) including its -inf corner.  Out of line and by
// value: the rare path must not take part in the register allocation of the star loop.
//   u, ub : member / background residual in units of 1 / kExpArgScale;  y : norm^-1/2;  wm : member weight
//   b     : fixed backgrounds: the packed background column (mantissa, or the plain term when bexp is
//           kMixFastFlag); fitted background: yb = (verr^2 + sigma_back^2)^-1/2
struct ExtFactor {
    double m;
    int e;
};
template <int BG>
static __device__ __noinline__ ExtFactor mix_term_slow(double fb, double u, double y, double wm, double b, int bexp, double ub) {
    const double zz = u * (1.0 / kExpArgScale);
    double sm, bm;
    int se, be;
    exp_neg_half(zz * zz, sm, se);
    if constexpr (BG == MCD_BG_GAUSSIAN) {
        const double zb = ub * (1.0 / kExpArgScale);
        double ebm;
        exp_neg_half(zb * zb, ebm, be);
        bm = fb * b * ebm;
    } else {
        bm = BG == MCD_BG_FIXED_DENSITY ? fb * b : b;
        be = bexp == kMixFastFlag ? 0 : bexp;
    }
    ExtFactor f;
    ext_add(wm * y * sm, se, bm, be, f.m, f.e);
    return f;
}

// 2^(j / kMixTableSize) into shared memory (called by the first kMixTableSize threads' worth of a CTA before
// its first barrier)
__device__ __forceinline__ void fill_exp2_table(double *table) {
#if MCD_MIX_LEAN == 2
    // one library exp2 per four entries, the other three by a multiplication (1.5 ulp)
    for (int j = threadIdx.x; j < kMixTableSize / 4; j += blockDim.x) {
        const double e = exp2((double)j * (4.0 / kMixTableSize));
        table[4 * j] = e;
        table[4 * j + 1] = e * 1.0006771306930664;      // 2^(1/1024)
        table[4 * j + 2] = e * 1.0013547198921082;      // 2^(2/1024)
        table[4 * j + 3] = e * 1.002032767907594;       // 2^(3/1024)
    }
#elif MCD_MIX_LEAN
    for (int j = threadIdx.x; j < kMixTableSize; j += blockDim.x) table[j] = exp2((double)j * (1.0 / kMixTableSize));
#else
    if (threadIdx.x < 64) table[threadIdx.x] = kExp2Table[threadIdx.x];
#endif
}

// ------------------------------------------------------------------------------------------
// one (walker, star) term
// ------------------------------------------------------------------------------------------
// Fast-path value of one mixture term: nothing is applied to the accumulators yet, so that the caller can run
// several stars through one basic block (independent dependency chains for the scheduler) and decide afterwards.
struct MixFast {
    double factor;      // wm y exp(-z^2/2) + background term
    double den;         // density + f_back (variants that normalise by it)
    int slow;           // fitted background: both components tiny, the term needs the extended-range path
};

// FAST_VALUE (FAST mixtures): the caller has checked that this star may take the fast path; its value goes to
// *fast and `A` is left alone.  Otherwise the term is evaluated and applied to `A` (fast or slow path as needed).
template <int ROT, int FREE, int BG, int MATH, bool FAST_VALUE = false>
__device__ __forceinline__ void term(const Walker &W, const Star<total_columns(ROT, FREE, BG)> &S, Accum<BG, MATH> &A,
                                     uint32_t exp2_table = 0, MixFast *fast = nullptr) {
    constexpr int NB = base_columns(ROT, FREE);
    constexpr bool FAST = MATH == MCD_MATH_FAST;
    // ---- geometry: numerator `num` of the rotation term, r^2 ---------------------------------
    double num, r2 = 0.0;
    if constexpr (FREE) {
        const double p1 = S.c[0], p2 = S.c[1], sd = S.c[2];
        const double dx = fma(p2, W.sb, -p1 * W.cb);            // -cos(dec) sin(ra - ra_c)
        // sin(dec) cos(dec_c) - cos(dec) cos(ra - ra_c) sin(dec_c), the middle factor expanded into the
        // stored p1, p2 with the walker's products sin/cos(ra_c - ra0) sin(dec_c): three instructions
        const double dy = fma(sd, W.cdc, fma(p2, W.cbs, p1 * W.sbs));
        r2 = fma(dx, dx, dy * dy);
        num = fma(dy, W.cy, dx * W.cx);
        if constexpr (ROT == MCD_ROT_CONSTANT) {
            // sin(theta_i - theta_0) needs the unit vector: divide by r; at r = 0 the reference has
            // theta_i = atan2(+0, -0) = pi, i.e. the unit vector (-1, 0)
            const bool origin = (__double2hiint(r2) | __double2loint(r2)) == 0;
            const double rinv = FAST ? fast_rsqrt(origin ? 1.0 : r2) : 1.0 / sqrt(origin ? 1.0 : r2);
            num = origin ? -W.cx : num * rinv;
        }
    } else if constexpr (ROT == MCD_ROT_CONSTANT) {
        num = fma(S.c[1], W.cy, S.c[0] * W.cx);
    } else {
        num = fma(S.c[1], W.cy, S.c[0] * W.cx);
        r2 = S.c[2];
    }
    const double v = S.c[NB - 2], e2 = S.c[NB - 1];
    // FAST mixtures: the packed v and the walker's v_sys both carry kExpArgScale (pack kernel, load_walker)
    const double dv = v - W.vsys;

    if constexpr (FAST) {
        // ---- dispersion and the combined variance ----------------------------------------
        double norm, D1 = 1.0;
        if constexpr (ROT == MCD_ROT_RADIAL) {
            D1 = fma(r2, W.ip2, 1.0);
            const double D2 = fma(r2, W.ia2, 1.0);
            // sigma_max^2 / sqrt(1 + r^2/a^2) + verr^2
            // (mixtures: W.s2h, W.s2 and e2 carry kMixVarScale, see load_walker)
            if constexpr (MCD_NEWTON == 2 && (BG == MCD_BG_NONE || MCD_MIX_LEAN != 0))
                norm = fma(W.s2h, rsqrt_twice(D2), e2);
            else
                norm = fma(W.s2, BG == MCD_BG_NONE ? fast_rsqrt(D2) : mix_rsqrt(D2), e2);
        } else {
            norm = e2 + W.s2;
        }
        // residual times D1: (v - v_sys) D1 - num  (= (v - v_los) D1)
        const double t = ROT == MCD_ROT_RADIAL ? fma(dv, D1, -num) : dv - num;
        const double q = ROT == MCD_ROT_RADIAL ? (D1 * D1) * norm : norm;
        if constexpr (BG == MCD_BG_NONE) {
            // chi^2 = t^2 / (D1^2 norm): one reciprocal per term, no log (running product)
            A.chi = fma(t * t, fast_rcp(q), A.chi);
            A.norm.mul_raw(norm);
        } else {
            // Mixture factor times sqrt(2 pi):  wm y exp(-z^2/2) + (background term),  y = norm^-1/2,
            // z = t y / D1, u = z kExpArgScale.  Fast path: both components as plain doubles, one FMA, the
            // factor multiplied into the running product (exponent folded every four stars).  It is valid
            // when the background term is a comfortable double (flag set by the pack kernel per star for the
            // fixed backgrounds, checked per term for the fitted one); everything else -- weights of exactly
            // 0 or 1, components hundreds of sigma out, the reference's -inf corner of runner.py:280-286 --
            // goes through the (mantissa, exponent) arithmetic below, which cannot underflow.
            const double yq = mix_var_rsqrt(q);
            const double u = t * yq;
            const double y = ROT == MCD_ROT_RADIAL ? yq * D1 : yq;
            const double wm = S.c[NB];
            int Nm, Nb = 0;
            const double em = exp_neg_sq_split(u, exp2_table, Nm);
            double bterm, yb = 0.0, ub = 0.0;
            bool tiny = false;
            if constexpr (BG == MCD_BG_FIXED_PMEMBER) {
                bterm = S.c[NB + 1];
            } else if constexpr (BG == MCD_BG_FIXED_DENSITY) {
                bterm = W.fb * S.c[NB + 1];
            } else {
                yb = mix_var_rsqrt(e2 + W.sb2);
                ub = (v - W.vb) * yb;
                const double eb = exp_neg_sq_split(ub, exp2_table, Nb);
                bterm = (W.fb * yb) * scale_by_table_exponent(eb, Nb);
                tiny = max(Nm, Nb) < kMixSlowExp * kMixTableSize;
            }
            const double factor = fma(wm * y, scale_by_table_exponent(em, Nm), bterm);
            if constexpr (FAST_VALUE) {
                fast->factor = factor;
                fast->den = wm + W.fb;
                fast->slow = tiny ? 1 : 0;
            } else {
                bool slow = W.slow != 0 || tiny;
                if constexpr (has_icol(BG, MATH)) slow |= (S.e != kMixFastFlag);
                if constexpr (BG != MCD_BG_FIXED_PMEMBER) A.den.mul_raw(wm + W.fb);
                if (!slow) {
                    A.num.mul_raw(factor);
                } else {
                    const ExtFactor f = mix_term_slow<BG>(W.fb, u, y, wm, BG == MCD_BG_GAUSSIAN ? yb : S.c[NB + 1],
                                                          BG == MCD_BG_GAUSSIAN ? 0 : S.e, ub);
                    A.mul_slow(f.m, f.e);
                }
            }
        }
    } else {
        // ---- PLAIN: the reference's formulas with library div / sqrt / log / exp ----------
        double sig2, vlos;
        if constexpr (ROT == MCD_ROT_RADIAL) {
            vlos = W.vsys + num / (1.0 + r2 * W.ip2);                 // model.py:180
            sig2 = W.s2 / sqrt(1.0 + r2 * W.ia2);                     // model.py:128, squared
        } else {
            vlos = W.vsys + num;                                      // constant.py:111
            sig2 = W.s2;                                              // constant.py:74
        }
        A.vlos = vlos;
        A.sig2 = sig2;
        const double norm = e2 + sig2;                                // runner.py:261
        const double resid = v - vlos;
        const double lm = -0.5 * log(kTwoPi * norm) + (-0.5 * (resid * resid) / norm);   // runner.py:262-271
        if constexpr (BG == MCD_BG_NONE) {
            A.sum += lm;
        } else {
            double wm, lb;
            if constexpr (BG == MCD_BG_FIXED_PMEMBER) {
                wm = S.c[NB];
                lb = S.c[NB + 1];
            } else if constexpr (BG == MCD_BG_FIXED_DENSITY) {
                const double d = S.c[NB];
                wm = d / (d + W.fb);                                  // model.py:589
                lb = S.c[NB + 1];
            } else {
                const double d = S.c[NB];
                wm = d / (d + W.fb);                                  // constant.py:339, model.py:427
                const double nb = e2 + W.sb2;                         // constant.py:333-336
                const double rb = v - W.vb;
                lb = -0.5 * log(kTwoPi * nb) + (-0.5 * (rb * rb) / nb);
            }
            // runner.py:279-284, constant.py:320-323, model.py:452-454,614-618
            const double mx = fmax(lm, lb);
            const double pm = wm * exp(lm - mx), pb = (1.0 - wm) * exp(lb - mx);
            A.sum += mx + log(pm + pb);
            A.pmember = pm / (pm + pb);     // constant.py:374, model.py:509-510,686-687
        }
    }
}

// two adjacent stars of the current stage: one 16-byte shared-memory load per column
template <int NC, bool ICOL>
__device__ __forceinline__ void load_pair(const double *__restrict__ c, const int32_t *__restrict__ ci, int tile, int i,
                                          Star<NC> &a, Star<NC> &b) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const double2 v = *reinterpret_cast<const double2 *>(c + k * tile + i);
        a.c[k] = v.x;
        b.c[k] = v.y;
    }
    if (ICOL) {
        const int2 e = *reinterpret_cast<const int2 *>(ci + i);
        a.e = e.x;
        b.e = e.y;
    } else {
        a.e = b.e = 0;
    }
}

template <int NC, bool ICOL>
__device__ __forceinline__ void load_one(const double *__restrict__ c, const int32_t *__restrict__ ci, int tile, int i,
                                         Star<NC> &a) {
#pragma unroll
    for (int k = 0; k < NC; ++k) a.c[k] = c[k * tile + i];
    a.e = ICOL ? ci[i] : 0;
}

// N adjacent stars (2 or 4).  The FAST mixtures decide once for the group whether all take the fast path: the
// per-star flags of the fixed backgrounds are checked first (integer compares), the per-term condition of the
// fitted background after the values are known; the N fast evaluations form one basic block.
template <int ROT, int FREE, int BG, int MATH, int N>
__device__ __forceinline__ void term_group(const Walker &W, const Star<total_columns(ROT, FREE, BG)> *const (&s)[N],
                                           Accum<BG, MATH> &A, uint32_t exp2_table) {
    if constexpr (MATH == MCD_MATH_FAST && BG != MCD_BG_NONE) {
        int slow = W.slow;
#if !MCD_FLAG_LATE
        if constexpr (has_icol(BG, MATH)) {
#pragma unroll
            for (int k = 0; k < N; ++k) slow |= s[k]->e ^ kMixFastFlag;
        }
#endif
        if (slow == 0) {
            MixFast f[N];
#pragma unroll
            for (int k = 0; k < N; ++k) term<ROT, FREE, BG, MATH, true>(W, *s[k], A, exp2_table, &f[k]);
#if MCD_FLAG_LATE
            // the per-star flags are looked at after the (harmless) fast evaluation: the first instructions of
            // the block then wait for the star columns only, not for a flag load, compare and branch
            if constexpr (has_icol(BG, MATH)) {
#pragma unroll
                for (int k = 0; k < N; ++k) slow |= s[k]->e ^ kMixFastFlag;
            }
#endif
            if constexpr (BG == MCD_BG_GAUSSIAN) {
#pragma unroll
                for (int k = 0; k < N; ++k) slow |= f[k].slow;
            }
            if (slow == 0) {
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    A.num.mul_raw(f[k].factor);
                    if constexpr (BG != MCD_BG_FIXED_PMEMBER) A.den.mul_raw(f[k].den);
                }
                return;
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) term<ROT, FREE, BG, MATH, false>(W, *s[k], A, exp2_table);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) term<ROT, FREE, BG, MATH>(W, *s[k], A, exp2_table);
    }
}

template <int ROT, int FREE, int BG, int MATH>
__device__ __forceinline__ void term_pair(const Walker &W, const Star<total_columns(ROT, FREE, BG)> &a,
                                          const Star<total_columns(ROT, FREE, BG)> &b, Accum<BG, MATH> &A, uint32_t exp2_table) {
    const Star<total_columns(ROT, FREE, BG)> *const group[2] = {&a, &b};
    term_group<ROT, FREE, BG, MATH, 2>(W, group, A, exp2_table);
}

// ------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier plumbing
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy through the TMA unit; completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Ticket of the cross-CTA reductions: one acq_rel read-modify-write at GPU scope.  Together with the
// bar.sync before (all partials of this CTA are written) and after (the ticket is known) it orders
// "every earlier CTA's partials" before "the last CTA's reads" -- release/acquire cumulativity --
// without the sequentially consistent fence + L1 invalidation that __threadfence() costs (about a
// microsecond each, which is what small catalogues spend most of their kernel time on).  The
// partials are read with ld.cg (L2), so no L1 line can be stale.
__device__ __forceinline__ unsigned int take_ticket(unsigned int *counter) {
    unsigned int ticket;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(counter) : "memory");
    return ticket;
}

// ------------------------------------------------------------------------------------------
// shards -> catalogue: one-shot all-reduce over NVLink peer memory, fused into the kernel tail
// ------------------------------------------------------------------------------------------
// Called by every thread of the CTA that finished a walker group.  Every rank runs the kernel on its
// shard with the same theta; `owner` threads hold the shard's sum for walker `w`.  data/flags live in
// symmetric memory; parity double-buffers consecutive calls (a rank can be at most one call ahead,
// because it cannot finish call k+1 before every peer has published call k+1).  Kept out of line so
// that this cold path does not take part in the register allocation of the star loop.
// `buffer` selects one of kXchgSlots buffers, `epoch` is the unique, never repeating tag of this call.
// Consecutive calls must use different slots (a rank can be one call ahead of a peer that is still
// reading): host-counted calls alternate 0/1 by epoch parity, the sampler's half-steps alternate 2/3.
#ifdef MCD_KERNEL_PROFILE
// [0] sums stored to every peer, [1] flags published, [2] every peer's flag seen (thread 0 of the exchanging CTA)
__device__ unsigned long long g_xchg_stamp[3];
#define MCD_XSTAMP(i) do { if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_xchg_stamp[i])); } while (0)
#else
#define MCD_XSTAMP(i) do { } while (0)
#endif

static __device__ __noinline__ double exchange_shard_sums(const LaunchParams &P, double total, int w, int group, bool owner,
                                                   int *timed_out, int buffer, unsigned long long epoch) {
    const int tid = threadIdx.x;
    const int par = buffer;
    const size_t slot = ((size_t)par * P.xchg_world + P.xchg_rank) * P.xchg_capacity + w;
    if (tid == 0) *timed_out = 0;
    if (owner) {
        for (int peer = 0; peer < P.xchg_world; ++peer) P.xchg_data[peer][slot] = total;   // st over NVLink
        __threadfence_system();
    }
    __syncthreads();
    MCD_XSTAMP(0);
    if (tid < P.xchg_world) {
        // publish: flag[par][my rank][group] on rank `tid` <- epoch (release, system scope)
        unsigned long long *dst = P.xchg_flags[tid] + ((size_t)par * P.xchg_world + P.xchg_rank) * kMaxXchgGroups + group;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
        // wait: flag[par][rank tid][group] in MY buffer == epoch (acquire, system scope)
        const unsigned long long *src =
            P.xchg_flags[P.xchg_rank] + ((size_t)par * P.xchg_world + tid) * kMaxXchgGroups + group;
        unsigned long long seen;
        const long long t0 = clock64();
        MCD_XSTAMP(1);
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
            // a peer that never arrives (crashed rank) must not hang the GPU: give up after ~10 s
            if (clock64() - t0 > 20000000000LL) {
                *timed_out = 1;
                if (P.xchg_status) *P.xchg_status = 1;       // read by the host after the run (exchange_status)
                break;
            }
        } while (seen != epoch);
    }
    __syncthreads();
    MCD_XSTAMP(2);
    if (owner) {
        double s = *timed_out ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
        for (int r = 0; r < P.xchg_world; ++r)
            s += __ldcg(&P.xchg_data[P.xchg_rank][((size_t)par * P.xchg_world + r) * P.xchg_capacity + w]);
        total = s;      // same order on every rank: bit-identical results across the box
    }
    return total;
}

// The same exchange with self-validating words (the LL scheme of NCCL, as the resident chain kernel uses inside
// one GPU): the owner of walker w stores {low half | tag, high half | tag} into every rank's slot with one
// 16-byte store, then polls its OWN rank's slots of all ranks until both tags match and adds in rank order.
// Nothing orders the data before a separate flag, so there is no __threadfence_system, no block barrier and no
// second round trip (8 GPUs, 512 walkers over 50 000-star shards: 67.2 us per call against 69.9 us with the flag
// protocol, MCD_XCHG=flags).  `tag` must differ
// from the tag the slot carried two calls earlier and never be 0 (the buffer starts zero-filled).
static __device__ __noinline__ double exchange_shard_sums_tagged(const LaunchParams &P, double total, int w, bool owner,
                                                          int buffer, unsigned int tag) {
    if (!owner) return total;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(total);
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long lo = (bits & 0xffffffffULL) | t, hi = (bits >> 32) | t;
    const size_t mine = (((size_t)buffer * P.xchg_world + P.xchg_rank) * P.xchg_capacity + w) * 2;
    MCD_XSTAMP(0);
    for (int peer = 0; peer < P.xchg_world; ++peer)
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(P.xchg_words[peer] + mine), "l"(lo), "l"(hi) : "memory");
    MCD_XSTAMP(1);
    double s = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < P.xchg_world; ++r) {
        const unsigned long long *src = P.xchg_words[P.xchg_rank] + (((size_t)buffer * P.xchg_world + r) * P.xchg_capacity + w) * 2;
        unsigned long long a, b;
        unsigned int polls = 0u;
        while (true) {
            asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
            if ((unsigned int)(a >> 32) == tag && (unsigned int)(b >> 32) == tag) break;
            // a peer that never arrives (crashed rank) must not hang the GPU: give up after ~10 s
            if ((++polls & 1023u) == 0u && clock64() - t0 > 20000000000LL) {
                if (P.xchg_status) *P.xchg_status = 1;
                return __longlong_as_double(0x7ff8000000000000LL);
            }
        }
        s += __longlong_as_double((long long)((a & 0xffffffffULL) | (b << 32)));     // rank order: bit-identical everywhere
    }
    MCD_XSTAMP(2);
    return s;
}

// acceptance of the fused half-step: (P-1) ln z + lnp(q) - lnp(s) > ln u'; NaN never accepts
static __device__ __noinline__ void accept_proposal(const LaunchParams &P, int seg, int k, double lnp_new) {
    const FuseParams &F = P.fuse;
    double q[MCD_MAX_THETA];
    int row;
    const double z = draw_proposal(P, seg, k, q, row);
    double u0, u1;
    uniforms(F.seed, F.step[0], (uint32_t)F.half, (uint32_t)(seg * F.walkers_total + k), 1u, u0, u1);
    const double diff = (P.n_theta - 1.0) * log(z) + lnp_new - F.lnp[row];
    if (diff > log(u0)) {
        double *s = F.pos + (size_t)row * P.n_theta;
        for (int p = 0; p < P.n_theta; ++p) s[p] = q[p];
        F.lnp[row] = lnp_new;
        F.n_accepted[row] += 1;
    }
}

// ------------------------------------------------------------------------------------------
// chunks -> result
// ------------------------------------------------------------------------------------------
// Called by every thread of every CTA after its chunk partial is stored.  Two ticketed levels, both
// adding in index order (independent of CTA scheduling): the last CTA of a super-chunk (P.super
// consecutive chunks) adds their partials, the last super-chunk to finish adds the super-chunk sums,
// applies the prior mask and hands the result on (output vector, cross-GPU exchange or the fused
// acceptance).  `owner` threads (slice 0 of a valid walker) carry walker `w`.  Out of line on
// purpose: the register allocation of the star loop must not depend on this cold code.
// -DMCD_KERNEL_PROFILE: thread 0 of every CTA stamps %globaltimer at the stages of a likelihood launch;
// the CTA that finishes the launch prints its own timeline relative to the earliest CTA start
// (tools/profile_config.py with MCD_B200_LIB pointing at such a build).  Not in the product library.
#ifdef MCD_KERNEL_PROFILE
__device__ unsigned long long g_kernel_start[2] = {~0ull, ~0ull};
__device__ unsigned int g_kernel_launch = 0;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MCD_KSTAMP(arr, i) do { if (threadIdx.x == 0) (arr)[i] = global_ns(); } while (0)
#else
#define MCD_KSTAMP(arr, i) do { } while (0)
#endif

// sum over the star slices of one walker held in red[slice * wl + lane]; fixed order, four chains
__device__ __forceinline__ double sum_slices(const double *red, int lane, int wl, int slices) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 4 <= slices; j += 4) {
        a0 += red[j * wl + lane];
        a1 += red[(j + 1) * wl + lane];
        a2 += red[(j + 2) * wl + lane];
        a3 += red[(j + 3) * wl + lane];
    }
    for (; j < slices; ++j) a0 += red[j * wl + lane];
    return (a0 + a1) + (a2 + a3);
}

// Sum rows [first, last) of a [rows][n_walkers] array of per-CTA sums for walker w, with every thread
// of the CTA loading (thread (lane, slice) takes rows first + slice, first + slice + slices, ...:
// kGatherDepth independent loads in flight per thread, one L2 round trip for up to kGatherDepth * slices rows).  The order
// of the additions depends on the launch geometry only, not on which CTA runs this.  All threads of
// the CTA must call; the result is valid in the owner threads (slice 0).
__device__ __forceinline__ double gather_rows(const double *rows, int first, int last, const LaunchParams &P, int w,
                                              double *red) {
    const int tid = threadIdx.x;
    const int lane = tid % P.wl, slice = tid / P.wl;
    const bool valid = slice < P.slices && w < P.n_walkers;
    double acc = 0.0;
    if (valid) {
        const size_t stride = (size_t)P.n_walkers;
        const int step = P.slices;
        for (int c = first + slice; c < last; c += kGatherDepth * step) {
            double v[kGatherDepth];
#pragma unroll
            for (int u = 0; u < kGatherDepth; ++u) {
                const int row = c + u * step;
                v[u] = row < last ? __ldcg(&rows[(size_t)row * stride + w]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kGatherDepth; ++u) acc += v[u];
        }
    }
    __syncthreads();                       // red[] may still be read by the caller's slice sum
    red[tid] = acc;
    __syncthreads();
    return (valid && slice == 0) ? sum_slices(red, lane, P.wl, P.slices) : 0.0;
}

template <int MATH, bool FUSE>
__device__ __noinline__ void finish_walker_group(const LaunchParams &P, int seg, int chunk, int group, int w, int n_chunks,
                                                 long long seg_stars, bool owner, int prior_ok, int *s_last, double *red,
                                                 unsigned long long *stamps = nullptr) {
    const int tid = threadIdx.x;
    const int n_super = (n_chunks + P.super - 1) / P.super;
    unsigned int *cnt = P.counters + ((size_t)seg * P.n_groups + group) * (P.n_super + 1);
    const int sup = chunk / P.super;
    const int c_begin = sup * P.super;
    const int c_end = min(n_chunks, c_begin + P.super);
    __syncthreads();
    if (tid == 0) *s_last = (take_ticket(&cnt[sup]) == (unsigned int)(c_end - c_begin) - 1u);
    __syncthreads();
    MCD_KSTAMP(stamps, 7);
    if (!*s_last) return;
    // words of a graph-replayed call that live in device memory (refreshed by the copy that brought theta): asked
    // for here, so that their round trip to L2 overlaps the gathers instead of following them
    const unsigned long long epoch_word = (P.xchg_world > 1 && P.xchg_epoch_ptr) ? *P.xchg_epoch_ptr : P.xchg_epoch;
    const unsigned int host_tag = (P.host_words && P.host_seq_ptr) ? (unsigned int)*P.host_seq_ptr : P.host_tag;
    const double level1 = gather_rows(P.partials + (size_t)seg * P.n_chunks * P.n_walkers, c_begin, c_end, P, w, red);
    if (tid == 0) cnt[sup] = 0u;
    MCD_KSTAMP(stamps, 8);
    double total = 0.0;
    if (n_super > 1) {
        if (owner) P.partials2[((size_t)seg * P.n_super + sup) * P.n_walkers + w] = level1;
        __syncthreads();
        if (tid == 0) *s_last = (take_ticket(&cnt[P.n_super]) == (unsigned int)n_super - 1u);
        __syncthreads();
        MCD_KSTAMP(stamps, 9);
        if (!*s_last) return;
        total = gather_rows(P.partials2 + (size_t)seg * P.n_super * P.n_walkers, 0, n_super, P, w, red);
        if (tid == 0) cnt[P.n_super] = 0u;
    } else {
        total = level1;            // at most P.super chunks: one level is the whole reduction
    }
    if (owner) {
        if (MATH == MCD_MATH_FAST) total = fma((double)seg_stars, -0.5 * kLn2Pi, total);
        const bool rejected = P.apply_prior && !prior_ok;
        total = rejected ? __longlong_as_double(0xfff0000000000000LL) : total;
    }
    if (P.xchg_world > 1) {
        // host-counted calls: the epoch is a kernel argument, or -- when the launch is replayed from a
        // CUDA graph -- a word in device memory refreshed by the copy that brought theta
        unsigned long long epoch = epoch_word;
        int slot = (int)(epoch & 1ull);
        if constexpr (FUSE) {
            // sampler half-steps: the tag comes from the device-side step counter plus the ensemble's
            // nonce (the step counter restarts with every ensemble, the flags persist on the handle)
            slot = 2 + P.fuse.half;
            epoch = P.fuse.tag_base | (2ull * P.fuse.step[0] + (unsigned long long)P.fuse.half);
        }
        if (P.xchg_tagged_mode) {
            // 32-bit tag: host-counted calls use the epoch itself; half-steps mix the ensemble's nonce (bits
            // 36.. of tag_base) into the top byte, so consecutive users of a slot never share a tag
            unsigned int tag = (unsigned int)epoch;
            if constexpr (FUSE) tag = ((unsigned int)((P.fuse.tag_base >> 36) % 255ull + 1ull) << 24) ^ ((unsigned int)epoch & 0xffffffu);
            total = exchange_shard_sums_tagged(P, total, w, owner, slot, tag == 0u ? 1u : tag);
        } else {
            total = exchange_shard_sums(P, total, w, group, owner, s_last, slot, epoch);
        }
    }
    if constexpr (FUSE) {
        // accept or reject in place.  Safe without further synchronisation: the positions of the active
        // half are read only by the CTAs of their own walker group, all of which have finished (ticket),
        // and the other half is read-only during this half-step.
        if (owner) accept_proposal(P, seg, w, total);
    } else {
        if (P.host_words) {
            // pinned host memory, self-validating words (see LaunchParams): nothing to order, nothing to count
            if (owner) {
                const unsigned int tag = host_tag;
                const unsigned long long bits = (unsigned long long)__double_as_longlong(total);
                const unsigned long long t = (unsigned long long)tag << 32;
                asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(P.host_words + 2 * ((size_t)seg * P.n_walkers + w)),
                             "l"((bits & 0xffffffffULL) | t), "l"((bits >> 32) | t)
                             : "memory");
            }
        } else if (owner) {
            P.out[(size_t)seg * P.n_walkers + w] = total;
        }
    }
#ifdef MCD_KERNEL_PROFILE
    if (tid == 0 && stamps && group == P.n_groups - 1) {
        const unsigned long long end = global_ns();
        const unsigned int launch = g_kernel_launch;
        const unsigned long long t0 = g_kernel_start[launch & 1u];
        printf("launch %u (grid %u x %u, chunks %d, super %d): finishing CTA %u entered +%llu ns | barriers ready +%llu | "
               "walker loaded +%llu | first tile +%llu | star loop done +%llu | slices summed +%llu | partial stored +%llu | "
               "ticket 1 +%llu | level 1 summed +%llu | ticket 2 +%llu | end +%llu\n",
               launch, gridDim.x, gridDim.y, n_chunks, n_super, blockIdx.x, stamps[0] - t0, stamps[1] - t0, stamps[2] - t0,
               stamps[3] - t0, stamps[4] - t0, stamps[5] - t0, stamps[6] - t0, stamps[7] - t0, stamps[8] - t0,
               n_super > 1 ? stamps[9] - t0 : 0ull, end - t0);
        if (P.xchg_world > 1)
            printf("  rank %d exchange: begin / sums stored to peers +%llu ns | stores issued / flags published +%llu | all %d ranks seen +%llu (waited %llu ns)\n",
                   P.xchg_rank, g_xchg_stamp[0] - t0, g_xchg_stamp[1] - t0, P.xchg_world, g_xchg_stamp[2] - t0,
                   g_xchg_stamp[2] - g_xchg_stamp[1]);
        g_kernel_start[launch & 1u] = ~0ull;
        g_kernel_launch = launch + 1u;
    }
#endif
}

// ------------------------------------------------------------------------------------------
// the lnlike / lnprob kernel
// ------------------------------------------------------------------------------------------
// SEG: segmented launch (blockIdx.y = segment).  A separate instantiation so that the single-catalogue
// kernel carries none of the segment bookkeeping (it changed the star loop's register allocation
// and cost 2.8 % on the headline workload when it was a run-time branch).
// FUSE: ensemble half-step fused in (proposal drawn in the prologue, acceptance in the finishing CTA).
template <int ROT, int FREE, int BG, int MATH, bool SEG, bool FUSE>
__global__ void __launch_bounds__(kBlock, (BG == MCD_BG_NONE ? MCD_MIN_BLOCKS : MCD_BG_MIN_BLOCKS)) lnlike_kernel(const __grid_constant__ LaunchParams P, const __grid_constant__ ThetaBlock T) {
    constexpr int NC = total_columns(ROT, FREE, BG);
    constexpr bool ICOL = has_icol(BG, MATH);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[kStages];
    __shared__ double red[kBlock];
    __shared__ int s_last;
    // 2^(j/64) for the exponentials of the mixture terms (mcd_math.cuh); filled before the first barrier
    constexpr bool EXP_TABLE = BG != MCD_BG_NONE && MATH == MCD_MATH_FAST;
    __shared__ double s_exp2[EXP_TABLE ? kMixTableSize : 1];
    if constexpr (EXP_TABLE) fill_exp2_table(s_exp2);
    uint32_t exp2_addr = EXP_TABLE ? smem_u32(s_exp2) : 0u;
    asm volatile("" : "+r"(exp2_addr));      // keep the address in a register (rebuilt per pair from the CTA id otherwise)

#ifdef MCD_KERNEL_PROFILE
    __shared__ unsigned long long s_stamp[12];
    if (threadIdx.x == 0) {
        s_stamp[0] = global_ns();
        atomicMin(&g_kernel_start[g_kernel_launch & 1u], s_stamp[0]);
    }
#else
    unsigned long long *s_stamp = nullptr;
#endif
    const int tile = P.tile;                     // stars per stage actually copied (<= kMaxTile)
    constexpr int TS = kMaxTile;                 // column stride in shared memory: compile-time, so
                                                 // that the LDS offsets of the star loop are immediates
    double *sd = reinterpret_cast<double *>(smem_raw);                       // [kStages][NC][TS]
    int32_t *si = reinterpret_cast<int32_t *>(sd + (size_t)kStages * NC * TS);   // [kStages][TS]

    const int tid = threadIdx.x;
    // walker groups of one star chunk are adjacent in launch order: the second group finds the chunk's
    // tiles in L2 instead of re-reading them from HBM
    const int chunk = blockIdx.x / P.n_groups, group = blockIdx.x % P.n_groups, seg = SEG ? blockIdx.y : 0;
    const int lane = tid % P.wl, slice = tid / P.wl;
    const int w = group * P.wl + lane;
    const bool valid = slice < P.slices && w < P.n_walkers;

    // Segments (blockIdx.y): independent star ranges with their own walkers, e.g. the radial bins of
    // bin/run.py:179-190 / bin/run_tests.py:81-97 fitted in one launch.  One segment = whole shard.
    long long seg_stars = P.n_stars, seg_offset = 0;
    int n_tiles = P.n_tiles, n_chunks = P.n_chunks;
    if constexpr (SEG) {                                // segment sizes fit 32 bits (checked at pack time)
        const int count = (int)(P.seg_begin[seg + 1] - P.seg_begin[seg]);
        seg_stars = count;
        seg_offset = P.seg_packed[seg];                 // 16-star aligned position in the packed columns
        n_tiles = (count + tile - 1) / tile;
        n_chunks = max(1, (n_tiles + P.tiles_per_chunk - 1) / P.tiles_per_chunk);
    }
    if (chunk >= n_chunks) return;          // the grid is sized for the largest segment

    const int t_begin = chunk * P.tiles_per_chunk;
    const int t_end = min(n_tiles, t_begin + P.tiles_per_chunk);
    const int n_my_tiles = max(0, t_end - t_begin);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    MCD_KSTAMP(s_stamp, 1);
    const uint32_t stage_bytes = (uint32_t)tile * (NC * 8u + (ICOL ? 4u : 0u));
    auto issue = [&](int k) {   // tid 0 only
        const int stage = k % kStages;
        const size_t off = (size_t)seg_offset + (size_t)(t_begin + k) * tile;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[stage], stage_bytes);
#pragma unroll
        for (int c = 0; c < NC; ++c)
            bulk_g2s(sd + ((size_t)stage * NC + c) * TS, P.cols[c] + off, (uint32_t)tile * 8u, &bars[stage]);
        if (ICOL) bulk_g2s(si + (size_t)stage * TS, P.icol + off, (uint32_t)tile * 4u, &bars[stage]);
    };

    if (tid == 0 && n_my_tiles > 0) issue(0);

    Walker W;
    W.prior_ok = 0;
    if constexpr (FUSE) {
        if (valid) {
            double q[MCD_MAX_THETA];
            int row_index;
            draw_proposal(P, seg, w, q, row_index);
            load_walker<ROT, FREE, BG, EXP_TABLE>(P, q, W);
        }
    } else {
        // theta in device memory, or inside the kernel arguments (small host-buffer calls)
        const double *theta = P.theta ? P.theta : T.v;
        if (valid) load_walker<ROT, FREE, BG, EXP_TABLE>(P, theta + (size_t)(seg * P.n_walkers + w) * P.n_theta, W);
    }
    // a walker outside its box prior is never evaluated by the reference (runner.py:303-306)
    const bool active = valid && (W.prior_ok || !P.apply_prior);
    MCD_KSTAMP(s_stamp, 2);

    Accum<BG, MATH> A;
    A.reset();

    for (int k = 0; k < n_my_tiles; ++k) {
        const int stage = k % kStages;
        if (tid == 0 && k + 1 < n_my_tiles) issue(k + 1);
        mbar_wait(&bars[stage], (uint32_t)(k / kStages) & 1u);
#ifdef MCD_KERNEL_PROFILE
        if (k == 0) MCD_KSTAMP(s_stamp, 3);
#endif
        const long long first = (long long)(t_begin + k) * tile;
        const int n = (int)min((long long)tile, seg_stars - first);
        if (active) {
            const double *c = sd + (size_t)stage * NC * TS;
            const int32_t *ci = si + (size_t)stage * TS;
            const int step = 2 * P.slices;
            // pairs of adjacent stars: (2 slice, 2 slice + 1), then stride 2 slices
            const int n2 = n & ~1;
            int i = 2 * slice;
            // two pairs per iteration: four independent dependency chains per thread
            if constexpr ((BG == MCD_BG_NONE ? MCD_PAIRS : MCD_BG_PAIRS) == 2) for (; i + step < n2; i += 2 * step) {
                Star<NC> s0, s1, s2, s3;
                load_pair<NC, ICOL>(c, ci, TS, i, s0, s1);
                load_pair<NC, ICOL>(c, ci, TS, i + step, s2, s3);
                if constexpr (EXP_TABLE) {
                    // FAST mixtures: four stars through one basic block, exponent folded every four stars
                    const Star<NC> *const group[4] = {&s0, &s1, &s2, &s3};
                    term_group<ROT, FREE, BG, MATH, 4>(W, group, A, exp2_addr);
                } else {
                    term_pair<ROT, FREE, BG, MATH>(W, s0, s1, A, exp2_addr);
                    if constexpr (MCD_GROUP4 == 0) A.end_group();
                    term_pair<ROT, FREE, BG, MATH>(W, s2, s3, A, exp2_addr);
                }
                A.end_group();
            }
            for (; i < n2; i += step) {
                Star<NC> s0, s1;
                load_pair<NC, ICOL>(c, ci, TS, i, s0, s1);
                term_pair<ROT, FREE, BG, MATH>(W, s0, s1, A, exp2_addr);
                A.end_group();
            }
            if ((n & 1) && (n2 / 2) % P.slices == slice) {   // odd tail of the last tile
                Star<NC> s0;
                load_one<NC, ICOL>(c, ci, TS, n2, s0);
                term<ROT, FREE, BG, MATH>(W, s0, A, exp2_addr);
                A.end_group();
            }
            A.end_tile();
        }
        __syncthreads();   // everyone is done with `stage` before it is refilled
    }

    // ---- slices -> one value per walker of this CTA ----------------------------------------
    MCD_KSTAMP(s_stamp, 4);
    red[tid] = active ? A.value() : 0.0;
    __syncthreads();
    MCD_KSTAMP(s_stamp, 5);
    if (valid && slice == 0)
        P.partials[((size_t)seg * P.n_chunks + chunk) * P.n_walkers + w] = sum_slices(red, lane, P.wl, P.slices);
    MCD_KSTAMP(s_stamp, 6);

    // ---- chunks -> result (and shards -> catalogue, proposal acceptance): cold path, out of line ----
    finish_walker_group<MATH, FUSE>(P, seg, chunk, group, w, n_chunks, seg_stars, valid && slice == 0, W.prior_ok,
                                    &s_last, red, s_stamp);
}

// ------------------------------------------------------------------------------------------
// resident chains: whole stretch-move runs with the catalogue held in shared memory
// ------------------------------------------------------------------------------------------
// For catalogues of a few thousand stars (the per-radial-bin fits of bin/run.py:179-190 and
// bin/run_tests.py:81-97, BASELINE config 1) a likelihood launch is < 1 us of arithmetic inside
// ~13 us of launch, first-tile and cross-CTA reduction latency.  Here `group` CTAs per segment load
// the segment's packed columns into their shared memory ONCE (each CTA an even-sized slice of the
// stars), every CTA keeps positions and log-probabilities of the whole ensemble in shared memory, and
// every emcee iteration (red/blue split, proposals, likelihood over all stars, acceptance) runs inside
// the one launch.  With group == 1 __syncthreads is the only synchronisation.  With group > 1 (a
// cooperative launch: all CTAs are resident together) the CTAs of a segment exchange their slice sums
// through L2 once per half-step (tagged words, below), every CTA adds the slices in the same order and
// so takes the same accept/reject decisions on its own copy of the ensemble.  Same move, same Philox counters and the same per-term arithmetic (`term<>`) as the
// launch-per-half-step sampler.
// Small groups: slice sums travel through L2 as two 8-byte words, each carrying half of the double and
// the 32-bit number of the half-step it belongs to (the scheme of NCCL's LL protocol): a word is valid
// as soon as its own tag matches, so plain relaxed stores and loads are enough -- no flag, fence or
// barrier, one L2 round trip between "last CTA published" and "every CTA has every sum".  With many CTAs
// x many walkers every thread of the GPU polling its words costs more than it saves (measured: C2..C4
// of BASELINE.json); there the CTAs arrive on a counter per segment, one thread per CTA waits for it
// and the sums are plain doubles read once (ChainParams::tagged = 0).
// -DMCD_CHAIN_PROFILE: thread 0 of CTA 0 accumulates clock64() spent in each stage of a half-step and
// prints the averages when the kernel ends (tools/chain_group_sweep.py with MCD_B200_LIB pointing at
// such a build).  Not compiled into the product library.
#ifdef MCD_CHAIN_PROFILE
#define MCD_STAMP(i)                                   \
    do {                                               \
        if (blockIdx.x == 0 && tid == 0) {             \
            const long long now_ = clock64();          \
            prof[i] += now_ - last_;                   \
            last_ = now_;                              \
        }                                              \
    } while (0)
#else
#define MCD_STAMP(i) do { } while (0)
#endif

struct alignas(16) TaggedSum {
    unsigned long long lo, hi;
};
__device__ __forceinline__ void publish_sum(TaggedSum *slot, double value, unsigned int tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"((bits & 0xffffffffULL) | t),
                 "l"((bits >> 32) | t)
                 : "memory");
}
__device__ __forceinline__ bool read_sum(const TaggedSum *slot, unsigned int tag, double &value) {
    unsigned long long lo, hi;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(slot) : "memory");
    if ((unsigned int)(lo >> 32) != tag || (unsigned int)(hi >> 32) != tag) return false;
    value = __longlong_as_double((long long)((lo & 0xffffffffULL) | (hi << 32)));
    return true;
}
// A wait for the group gives up after ~1 s (a member never published: cannot happen in a cooperative
// launch, but a spin without a bound would hang the GPU if it did) and raises *status; every 1024
// failed polls a waiter also looks at *status, so once one wait has failed the rest of the kernel
// drains at once (`group_lost`) instead of timing out half-step by half-step.
constexpr long long kGroupWaitCycles = 2000000000LL;
__device__ __forceinline__ bool group_wait_failed(long long t0, unsigned int &polls, int *status) {
    if ((++polls & 1023u) != 0u) return false;
    if (clock64() - t0 > kGroupWaitCycles) atomicExch(status, 1);
    return *reinterpret_cast<volatile int *>(status) != 0;
}

// The counter-based random numbers of one half-step for active walker k = threadIdx.x: stretch factor
// z = ((a-1) u + 1)^2 / a, partner index, (P-1) ln z and ln u of the acceptance test (emcee
// StretchMove).  Out of line: one copy of two Philox blocks and two logs instead of three.
struct Draw {
    double z, lz, lu;
    int partner;
};
static __device__ __noinline__ Draw draw_half_step(const ChainParams &C, int n_theta, int seg, unsigned int step, int half) {
    Draw d{1.0, 0.0, 0.0, 0};
    const int W = C.n_walkers;
    const int ns = half == 0 ? C.n0 : W - C.n0;
    const int k = (int)threadIdx.x;
    if (k < ns) {
        const int nc = W - ns;
        double u0, u1;
        uniforms(C.seed, step, (uint32_t)half, (uint32_t)(seg * W + k), 0u, u0, u1);
        const double t = (C.a - 1.0) * u0 + 1.0;
        d.z = t * t / C.a;
        const int j = (int)(u1 * nc);
        d.partner = j >= nc ? nc - 1 : j;
        d.lz = (n_theta - 1.0) * log(d.z);
        uniforms(C.seed, step, (uint32_t)half, (uint32_t)(seg * W + k), 1u, u0, u1);
        d.lu = log(u0);
    }
    return d;
}

template <int ROT, int FREE, int BG, int MATH>
__global__ void __launch_bounds__(kChainBlock) chain_kernel(const __grid_constant__ LaunchParams P,
                                                            const __grid_constant__ ChainParams C) {
    constexpr int NC = total_columns(ROT, FREE, BG);
    constexpr bool ICOL = has_icol(BG, MATH);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool EXP_TABLE = BG != MCD_BG_NONE && MATH == MCD_MATH_FAST;
    __shared__ double s_exp2[EXP_TABLE ? kMixTableSize : 1];          // 2^(j / kMixTableSize), see exp_neg_sq_split
    if constexpr (EXP_TABLE) fill_exp2_table(s_exp2);
    uint32_t exp2_addr = EXP_TABLE ? smem_u32(s_exp2) : 0u;
    asm volatile("" : "+r"(exp2_addr));
    const int tid = threadIdx.x;
    const int G = C.group;
    const int seg = blockIdx.x / G;
    const int member = blockIdx.x - seg * G;          // this CTA's place in the segment's group
    const int n_segments = gridDim.x / G;
    const int W = C.n_walkers, NP = P.n_theta;
    const int stride = C.max_segment_padded;

    const long long seg_first = P.seg_begin ? P.seg_begin[seg] : 0;
    const int seg_stars = (int)((P.seg_begin ? P.seg_begin[seg + 1] : P.n_stars) - seg_first);
    const int first = min(seg_stars, member * C.stars_per_cta);
    const int n = min(seg_stars - first, C.stars_per_cta);             // stars of this CTA's slice
    const long long offset = (P.seg_begin ? P.seg_packed[seg] : 0) + first;
    unsigned int phase = 0;         // half-steps exchanged so far in this launch: tag of the slice sums
    bool group_lost = false;        // a wait for the group failed: stop waiting, the host reports the error

    // shared-memory carve-up (every block is a multiple of 16 bytes)
    double *cols = reinterpret_cast<double *>(smem_raw);                       // [NC][stride]
    int32_t *icol = reinterpret_cast<int32_t *>(cols + (size_t)NC * stride);  // [stride]  (ICOL only)
    double *red = reinterpret_cast<double *>(icol + (ICOL ? stride : 0));     // [kChainBlock]
    double *pos = red + kChainBlock;                                          // [W][NP]
    double *lnp = pos + (size_t)W * NP;                                       // [W]
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(lnp + W);   // [W]
    int *perm = reinterpret_cast<int *>(keys + W);                            // [W]
    int *nacc = perm + W;                                                     // [W]
    // proposals of the active half, [larger half][NP] (perm + nacc = 8 bytes per walker: 8-byte aligned)
    double *prop = reinterpret_cast<double *>(nacc + W);

    for (int c = 0; c < NC; ++c)
        for (int i = tid; i < n; i += kChainBlock) cols[(size_t)c * stride + i] = P.cols[c][offset + i];
    if (ICOL)
        for (int i = tid; i < n; i += kChainBlock) icol[i] = P.icol[offset + i];
    for (int i = tid; i < W * NP; i += kChainBlock) pos[i] = C.pos[(size_t)seg * W * NP + i];
    for (int i = tid; i < W; i += kChainBlock) {
        lnp[i] = C.lnp[(size_t)seg * W + i];
        nacc[i] = 0;
    }
    __syncthreads();

#ifdef MCD_CHAIN_PROFILE
    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long last_ = clock64();
#endif
    // Everything a half-step draws from the counter-based generator (`draw_half_step`) depends on (step,
    // half, walker) only, not on the state of the chain: thread k < (walkers of that half) computes it one
    // half-step ahead, while the slice sums of the current half-step travel, and keeps it in registers.
    Draw next = draw_half_step(C, NP, seg, C.step0, 0);
    // thread = (walker `lane` of the active half, star slice): the mapping of both halves, computed once
    // (three integer divisions per half otherwise)
    const int ns0 = C.n0, ns1 = W - C.n0;
    const int slices0 = kChainBlock / ns0, slices1 = kChainBlock / ns1;
    const int lane0 = tid % ns0, lane1 = tid % ns1, slice0 = tid / ns0, slice1 = tid / ns1;

    for (int it = 0; it < C.n_steps; ++it) {
        const unsigned int step = C.step0 + (unsigned int)it;
        MCD_STAMP(7);
        // ---- red/blue partition: rank of a random key ---------------------------------------------
        for (int w = tid; w < W; w += kChainBlock) {
            const uint4 r = philox4x32_10(make_uint4(step, 2u, (uint32_t)(seg * W + w), 7u),
                                          make_uint2((uint32_t)C.seed, (uint32_t)(C.seed >> 32)));
            keys[w] = (((unsigned long long)r.x << 32) | r.y);
        }
        __syncthreads();
        for (int w = tid; w < W; w += kChainBlock) {
            const unsigned long long mine = keys[w];
            int rank = 0;
            for (int o = 0; o < W; ++o) rank += (keys[o] < mine) || (keys[o] == mine && o < w);
            perm[rank] = w;
        }
        __syncthreads();
        MCD_STAMP(0);
        for (int half = 0; half < 2; ++half) {
            // each half of the ensemble fits the CTA (n_walkers <= kChainMaxWalkers = 2 kChainBlock)
            const int *active = perm + (half == 0 ? 0 : C.n0);
            const int *other = perm + (half == 0 ? C.n0 : 0);
            const int wl = half == 0 ? ns0 : ns1, slices = half == 0 ? slices0 : slices1;
            const int lane = half == 0 ? lane0 : lane1, slice = half == 0 ? slice0 : slice1;
            const bool valid = slice < slices;
            const bool owner = tid < wl;               // slice 0: proposes, adds up and accepts for walker `lane`
            const Draw cur = next;
            // ---- proposals of the active half -> shared memory (one thread per walker) -------------
            int wa = 0;
            if (owner) {
                wa = active[lane];
                const double *s = pos + (size_t)wa * NP;
                const double *c = pos + (size_t)other[cur.partner] * NP;
                for (int p = 0; p < NP; ++p) prop[(size_t)lane * NP + p] = c[p] - (c[p] - s[p]) * cur.z;
            }
            __syncthreads();
            MCD_STAMP(9);
            Walker Wk;
            Wk.prior_ok = 0;
            if (valid) load_walker<ROT, FREE, BG, EXP_TABLE>(P, prop + (size_t)lane * NP, Wk);
            MCD_STAMP(1);
            Accum<BG, MATH> A;
            A.reset();
            if (valid && Wk.prior_ok) {
                const int n2 = n & ~1;
                const int stepi = 2 * slices;
                int done = 0;
                for (int i = 2 * slice; i < n2; i += stepi) {
                    Star<NC> s0, s1;
                    load_pair<NC, ICOL>(cols, icol, stride, i, s0, s1);
                    term_pair<ROT, FREE, BG, MATH>(Wk, s0, s1, A, exp2_addr);
                    A.end_group();
                    if (++done == 64) {     // fold the running products before they can overflow
                        A.end_tile();
                        done = 0;
                    }
                }
                if ((n & 1) && (n2 / 2) % slices == slice) {
                    Star<NC> s0;
                    load_one<NC, ICOL>(cols, icol, stride, n2, s0);
                    term<ROT, FREE, BG, MATH>(Wk, s0, A, exp2_addr);
                    A.end_group();
                }
                A.end_tile();
            }
            red[tid] = (valid && Wk.prior_ok) ? A.value() : 0.0;
            MCD_STAMP(2);
            __syncthreads();
            MCD_STAMP(3);
            if (G > 1) {
                // this CTA's slice sums -> L2 ...
                ++phase;
                const size_t first_slot = ((size_t)((phase & 1u) * n_segments + seg) * G) * C.sum_stride;
                if (owner) {
                    const double mine = sum_slices(red, lane, wl, slices);
                    if (C.tagged) publish_sum(&C.group_sums[first_slot + (size_t)member * C.sum_stride + lane], mine, phase);
                    else reinterpret_cast<double *>(C.group_sums)[first_slot + (size_t)member * C.sum_stride + lane] = mine;
                }
                if (!C.tagged) {
                    // ... announced by one arrival per CTA on the segment's counter (the bar.sync before
                    // the release orders the other threads' stores before it)
                    __syncthreads();
                    if (tid == 0)
                        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(C.group_arrivals + seg) : "memory");
                }
            }
            // ... and while they travel, the random numbers of the next half-step
            next = draw_half_step(C, NP, seg, step + (unsigned int)half, half ^ 1);
            MCD_STAMP(4);
            if (G > 1) {
                // every CTA's sums in member order: thread (lane, slice) takes members slice, slice + slices, ...
                const size_t first_slot = ((size_t)((phase & 1u) * n_segments + seg) * G) * C.sum_stride;
                double acc = 0.0;
                if (!C.tagged) {
                    // large groups: one thread waits for all arrivals, then plain 8-byte reads from L2
                    if (tid == 0) {
                        const unsigned long long target = (unsigned long long)phase * (unsigned long long)G;
                        unsigned long long seen = 0;
                        unsigned int polls = 0u;
                        const long long t0 = clock64();
                        do {
                            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(C.group_arrivals + seg) : "memory");
                            if (seen < target && group_wait_failed(t0, polls, C.status)) group_lost = true;
                        } while (seen < target && !group_lost);
                    }
                    __syncthreads();           // also: red[] is read above and rewritten below
                    if (valid) {
                        const double *sums = reinterpret_cast<const double *>(C.group_sums) + first_slot;
                        constexpr int B = 16;
                        for (int g0 = slice; g0 < G; g0 += B * slices) {
                            double v[B];
#pragma unroll
                            for (int u = 0; u < B; ++u) {
                                const int g = g0 + u * slices;
                                v[u] = g < G ? __ldcg(&sums[(size_t)g * C.sum_stride + lane]) : 0.0;
                            }
#pragma unroll
                            for (int u = 0; u < B; ++u) acc += v[u];
                        }
                    }
                } else {
                    // small groups: the words carry their own tag, every thread polls the ones it adds
                    const TaggedSum *sums = C.group_sums + first_slot;
                    __syncthreads();           // red[] is read above and rewritten below
                    if (valid) {
                        constexpr int B = 4;
                        for (int g0 = slice; g0 < G; g0 += B * slices) {
                            double v[B];
                            unsigned int pending = 0u;
#pragma unroll
                            for (int u = 0; u < B; ++u) {
                                const int g = g0 + u * slices;
                                v[u] = 0.0;
                                if (g < G && !read_sum(&sums[(size_t)g * C.sum_stride + lane], phase, v[u])) pending |= 1u << u;
                            }
                            unsigned int polls = 0u;
                            const long long t0 = clock64();
                            while (pending && !group_lost) {
#pragma unroll
                                for (int u = 0; u < B; ++u)
                                    if ((pending >> u) & 1u)
                                        if (read_sum(&sums[(size_t)(g0 + u * slices) * C.sum_stride + lane], phase, v[u]))
                                            pending &= ~(1u << u);
                                if (pending && group_wait_failed(t0, polls, C.status)) group_lost = true;
                            }
#pragma unroll
                            for (int u = 0; u < B; ++u) acc += v[u];
                        }
                    }
                }
                red[tid] = acc;
                __syncthreads();
            }
            MCD_STAMP(5);
            if (owner) {
                double total = sum_slices(red, lane, wl, slices);
                if (MATH == MCD_MATH_FAST) total = fma((double)seg_stars, -0.5 * kLn2Pi, total);
                if (!Wk.prior_ok) total = __longlong_as_double(0xfff0000000000000LL);
                const double diff = cur.lz + total - lnp[wa];
                if (diff > cur.lu) {               // NaN never accepts
                    for (int p = 0; p < NP; ++p) pos[(size_t)wa * NP + p] = prop[(size_t)lane * NP + p];
                    lnp[wa] = total;
                    nacc[wa] += 1;
                }
            }
            __syncthreads();
            MCD_STAMP(6);
        }
        if (C.chain && member == 0) {
            const size_t rows = (size_t)n_segments * W;
            double *dst = C.chain + ((size_t)it * rows + (size_t)seg * W) * NP;
            for (int i = tid; i < W * NP; i += kChainBlock) dst[i] = pos[i];
            double *dl = C.chain_lnp + (size_t)it * rows + (size_t)seg * W;
            for (int i = tid; i < W; i += kChainBlock) dl[i] = lnp[i];
        }
    }
#ifdef MCD_CHAIN_PROFILE
    if (blockIdx.x == 0 && tid == 0 && C.n_steps >= 100) {
        const double h = 2.0 * C.n_steps;
        printf("chain profile (cycles per half-step, group %d): split %.0f | proposals %.0f | load_walker %.0f | "
               "stars %.0f | sync %.0f | publish + next draws %.0f | gather %.0f | accept %.0f | store %.0f\n",
               G, prof[0] / h, prof[9] / h, prof[1] / h, prof[2] / h, prof[3] / h, prof[4] / h, prof[5] / h,
               prof[6] / h, prof[7] / h);
    }
#endif
    if (member != 0) return;
    for (int i = tid; i < W * NP; i += kChainBlock) C.pos[(size_t)seg * W * NP + i] = pos[i];
    for (int i = tid; i < W; i += kChainBlock) {
        C.lnp[(size_t)seg * W + i] = lnp[i];
        C.n_accepted[(size_t)seg * W + i] += nacc[i];
    }
}

// ------------------------------------------------------------------------------------------
// per-star lnlike of one parameter vector (`no_sum=True`, model.py:565,620-621); with mode
// kPerStarMembership the a-posteriori membership probability of every star at that parameter vector
// (constant.py:366-374, model.py:458-510,625-687); with mode kPerStarModel the model curves themselves,
// v_los into `out` and sigma_los into `out2` (rotation_model / dispersion_model, constant.py:52-111,
// model.py:93-180; either pointer may be null)
// ------------------------------------------------------------------------------------------
template <int ROT, int FREE, int BG>
__global__ void per_star_kernel(const __grid_constant__ LaunchParams P, double *__restrict__ out,
                                double *__restrict__ out2, int mode) {
    constexpr int NC = total_columns(ROT, FREE, BG);
    __shared__ Walker Ws;
    if (threadIdx.x == 0) load_walker<ROT, FREE, BG>(P, P.theta, Ws);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_stars) return;
    Star<NC> S;
#pragma unroll
    for (int k = 0; k < NC; ++k) S.c[k] = P.cols[k][i];
    S.c[base_columns(ROT, FREE) - 1] *= P.verr2_unscale;      // 1, or 1 / kMixVarScale for a FAST mixture packing (exact)
    S.e = 0;
    Accum<BG, MCD_MATH_PLAIN> A;
    A.reset();
    const Walker W = Ws;
    term<ROT, FREE, BG, MCD_MATH_PLAIN>(W, S, A);
    if (mode == kPerStarModel) {
        if (out) out[i] = A.vlos;
        if (out2) {
            // the curves carry the sign of sigma_max (constant.py:74, model.py:128); the walker constants keep its square
            const int slot = P.slot[MCD_P_SIGMA_MAX];
            const double sigma_max = slot >= 0 ? P.theta[slot] * P.scale[MCD_P_SIGMA_MAX] : P.fixed_scaled[MCD_P_SIGMA_MAX];
            out2[i] = copysign(sqrt(A.sig2), sigma_max);
        }
    } else {
        out[i] = mode == kPerStarMembership ? A.pmember : A.value();
    }
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
#define MCD_DISPATCH_BG(ROT, FREE, MATH, FN, ...)                                            \
    switch (v.background) {                                                                   \
        case MCD_BG_NONE: return FN<ROT, FREE, MCD_BG_NONE, MATH>(__VA_ARGS__);               \
        case MCD_BG_FIXED_PMEMBER: return FN<ROT, FREE, MCD_BG_FIXED_PMEMBER, MATH>(__VA_ARGS__); \
        case MCD_BG_FIXED_DENSITY: return FN<ROT, FREE, MCD_BG_FIXED_DENSITY, MATH>(__VA_ARGS__); \
        default: return FN<ROT, FREE, MCD_BG_GAUSSIAN, MATH>(__VA_ARGS__);                    \
    }
#define MCD_DISPATCH_GEO(MATH, FN, ...)                                                       \
    if (v.rotation == MCD_ROT_CONSTANT) {                                                     \
        if (v.free_centre) { MCD_DISPATCH_BG(MCD_ROT_CONSTANT, 1, MATH, FN, __VA_ARGS__) }    \
        else { MCD_DISPATCH_BG(MCD_ROT_CONSTANT, 0, MATH, FN, __VA_ARGS__) }                  \
    } else {                                                                                  \
        if (v.free_centre) { MCD_DISPATCH_BG(MCD_ROT_RADIAL, 1, MATH, FN, __VA_ARGS__) }      \
        else { MCD_DISPATCH_BG(MCD_ROT_RADIAL, 0, MATH, FN, __VA_ARGS__) }                    \
    }

template <int ROT, int FREE, int BG, int MATH>
static cudaError_t launch_one(const LaunchParams &p, cudaStream_t stream, const ThetaBlock &t) {
    constexpr int NC = total_columns(ROT, FREE, BG);
    const size_t smem = (size_t)kStages * kMaxTile * (NC * 8 + (has_icol(BG, MATH) ? 4 : 0));
    dim3 grid((unsigned)p.n_chunks * (unsigned)p.n_groups, (unsigned)std::max(1, p.n_segments));
    if (p.seg_begin) {
        // segmented handles (RadialBinsFit) exist for the models without fitted background parameters
        if constexpr (BG == MCD_BG_NONE || BG == MCD_BG_FIXED_PMEMBER) {
            if (p.fuse.enabled) lnlike_kernel<ROT, FREE, BG, MATH, true, true><<<grid, kBlock, smem, stream>>>(p, t);
            else lnlike_kernel<ROT, FREE, BG, MATH, true, false><<<grid, kBlock, smem, stream>>>(p, t);
        } else {
            return cudaErrorInvalidValue;
        }
    } else if (p.fuse.enabled) {
        lnlike_kernel<ROT, FREE, BG, MATH, false, true><<<grid, kBlock, smem, stream>>>(p, t);
    } else {
        lnlike_kernel<ROT, FREE, BG, MATH, false, false><<<grid, kBlock, smem, stream>>>(p, t);
    }
    return cudaGetLastError();
}

template <int ROT, int FREE, int BG, int MATH>
static int occupancy_one() {
    constexpr int NC = total_columns(ROT, FREE, BG);
    const size_t smem = (size_t)kStages * kMaxTile * (NC * 8 + (has_icol(BG, MATH) ? 4 : 0));
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lnlike_kernel<ROT, FREE, BG, MATH, false, false>, kBlock, smem) !=
        cudaSuccess)
        return 1;
    return n > 0 ? n : 1;
}

template <int ROT, int FREE, int BG, int MATH>
static cudaError_t chain_one(const LaunchParams &p, const ChainParams &c, size_t smem, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(chain_kernel<ROT, FREE, BG, MATH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess) return err;
    const unsigned grid = (unsigned)(std::max(1, p.n_segments) * std::max(1, c.group));
    if (c.group > 1) {
        // the CTAs of a group wait for each other: cooperative launch, all of them resident together
        void *args[] = {const_cast<LaunchParams *>(&p), const_cast<ChainParams *>(&c)};
        return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(&chain_kernel<ROT, FREE, BG, MATH>), dim3(grid),
                                           dim3(kChainBlock), args, smem, stream);
    }
    chain_kernel<ROT, FREE, BG, MATH><<<grid, kChainBlock, smem, stream>>>(p, c);
    return cudaGetLastError();
}

#if MCD_TU_PART == 1
cudaError_t launch_lnlike_fast(const Variant &v, const LaunchParams &p, cudaStream_t stream, const ThetaBlock &t) {
    MCD_DISPATCH_GEO(MCD_MATH_FAST, launch_one, p, stream, t)
}
int occupancy_fast(const Variant &v) { MCD_DISPATCH_GEO(MCD_MATH_FAST, occupancy_one) }
cudaError_t launch_chain_fast(const Variant &v, const LaunchParams &p, const ChainParams &c, size_t smem, cudaStream_t stream) {
    MCD_DISPATCH_GEO(MCD_MATH_FAST, chain_one, p, c, smem, stream)
}
#else
cudaError_t launch_lnlike_plain(const Variant &v, const LaunchParams &p, cudaStream_t stream, const ThetaBlock &t) {
    MCD_DISPATCH_GEO(MCD_MATH_PLAIN, launch_one, p, stream, t)
}
int occupancy_plain(const Variant &v) { MCD_DISPATCH_GEO(MCD_MATH_PLAIN, occupancy_one) }
cudaError_t launch_chain_plain(const Variant &v, const LaunchParams &p, const ChainParams &c, size_t smem, cudaStream_t stream) {
    MCD_DISPATCH_GEO(MCD_MATH_PLAIN, chain_one, p, c, smem, stream)
}

template <int ROT, int FREE, int BG, int MATH_UNUSED>
static cudaError_t per_star_one(const LaunchParams &p, double *out, double *out2, int mode, cudaStream_t stream) {
    if (p.n_stars <= 0) return cudaSuccess;
    const int block = 256;
    const long long grid = (p.n_stars + block - 1) / block;
    per_star_kernel<ROT, FREE, BG><<<(unsigned)grid, block, 0, stream>>>(p, out, out2, mode);
    return cudaGetLastError();
}

cudaError_t launch_per_star(const Variant &v, const LaunchParams &p, double *out, double *out2, int mode,
                            cudaStream_t stream) {
    MCD_DISPATCH_GEO(MCD_MATH_PLAIN, per_star_one, p, out, out2, mode, stream)
}
#endif
#endif  // MCD_TU_PART != 0

#if MCD_TU_PART == 0
cudaError_t launch_lnlike_fast(const Variant &v, const LaunchParams &p, cudaStream_t stream, const ThetaBlock &t);
cudaError_t launch_lnlike_plain(const Variant &v, const LaunchParams &p, cudaStream_t stream, const ThetaBlock &t);
int occupancy_fast(const Variant &v);
int occupancy_plain(const Variant &v);
cudaError_t launch_chain_fast(const Variant &v, const LaunchParams &p, const ChainParams &c, size_t smem, cudaStream_t stream);
cudaError_t launch_chain_plain(const Variant &v, const LaunchParams &p, const ChainParams &c, size_t smem, cudaStream_t stream);

cudaError_t launch_lnlike(const Variant &v, const LaunchParams &p, cudaStream_t stream, const ThetaBlock *inline_theta) {
    static const ThetaBlock kNoTheta{};
    const ThetaBlock &t = inline_theta ? *inline_theta : kNoTheta;
    return v.math_mode == MCD_MATH_PLAIN ? launch_lnlike_plain(v, p, stream, t) : launch_lnlike_fast(v, p, stream, t);
}

int lnlike_blocks_per_sm(const Variant &v) {
    return v.math_mode == MCD_MATH_PLAIN ? occupancy_plain(v) : occupancy_fast(v);
}

static size_t chain_bytes(int nc, bool icol, long long stride, int n_walkers, int n_theta) {
    size_t b = (size_t)nc * stride * 8 + (icol ? (size_t)stride * 4 : 0);
    b += (size_t)kChainBlock * 8;                          // red
    b += (size_t)n_walkers * n_theta * 8 + (size_t)n_walkers * 8;   // pos, lnp
    b += (size_t)n_walkers * (8 + 4 + 4) + 8;              // keys, perm, nacc (+ alignment of prop)
    b += (size_t)((n_walkers + 1) / 2) * n_theta * 8;      // prop
    return (b + 15) & ~(size_t)15;
}

size_t chain_shared_bytes(const Variant &v, long long stars_per_cta, int n_walkers, int n_theta) {
    if (n_walkers > kChainMaxWalkers) return 0;
    const long long stride = ((stars_per_cta + 15) / 16) * 16;
    const size_t b = chain_bytes(variant_columns(v), variant_has_icol(v), stride, n_walkers, n_theta);
    return b <= (size_t)220 * 1024 ? b : 0;                // 227 KB per CTA on sm_100a, minus static use
}

cudaError_t launch_chain(const Variant &v, const LaunchParams &p, const ChainParams &c, size_t smem, cudaStream_t stream) {
    return v.math_mode == MCD_MATH_PLAIN ? launch_chain_plain(v, p, c, smem, stream)
                                         : launch_chain_fast(v, p, c, smem, stream);
}
#endif  // MCD_TU_PART == 0

}  // namespace mcd

// FP64 device arithmetic for the likelihood kernels (sm_100a).
//
// The B200 FP64 pipe issues one warp-wide DFMA every 2 cycles per SM sub-partition, and the
// compiler's IEEE division / sqrt / log / exp expand to 10-30 such instructions each plus
// slow-path branches.  The likelihood only needs ~1e-14 relative accuracy, not correct
// rounding, so the hot loop uses:
//   * MUFU.RCP64H / MUFU.RSQ64H seeds (rel. error ~2^-22, issued on the otherwise idle XU
//     pipe) refined by ONE cubic Newton step (3 resp. 5 DFMA) -> rel. error < 2^-60;
//   * no per-star log at all: sum_i ln(x_i) = ln(prod_i x_i), the running product is kept
//     as a mantissa in [1, 2^k) plus an integer exponent that is maintained with integer-pipe
//     bit operations on the high word;
//   * exp() only in the mixture variants, as 2^n * 2^f with n in an integer register and
//     2^f a degree-12 polynomial; the two mixture components are combined in this
//     (mantissa, exponent) form, i.e. a base-2 log-sum-exp without log or exp calls.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcd {

constexpr double kLn2 = 0.693147180559945309417232121458;
constexpr double kLog2e = 1.44269504088896340735992468100;
constexpr double kLn2Pi = 1.83787706640934548356065947281;      // ln(2 pi)
constexpr double kSqrt2Pi = 2.50662827463100050241576528481;    // sqrt(2 pi)
constexpr double kInvSqrt2Pi = 0.398942280401432677939946059934;
constexpr double kTwoPi = 6.28318530717958647692528676656;

__device__ __forceinline__ double rcp_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// Newton order of the reciprocal / reciprocal square root refinements:
//   2 (quadratic step, default): measured 9.8e-13 / 1.3e-12 relative (seed error 2^-20 on B200 for both
//                            MUFU seeds); 2 / 3 FP64 instructions (+1 integer op)
//   3 (cubic step)         : 2.2e-16 / 2.7e-16; 3 / 5 FP64 instructions
// The tolerance of the path is 1e-9 relative on lnprob (BASELINE.json north_star); a per-term relative error
// of 1.3e-12 ends up far below that on lnprob whatever the catalogue size (it is relative; measured against the
// oracle on every BASELINE configuration: 0.6e-14 .. 2.3e-14, profiles/r02_configs.md), five orders inside it.  On the
// headline workload the quadratic step is worth 8.3 % (7.395 -> 6.827 ms per 512-walker call over 1e7 stars,
// profiles/r02_ab_runs.md); round 1 measured 2.5 % on an earlier, less FP64-bound loop and kept the cubic
// step.  -DMCD_NEWTON=3 restores it (A/B builds: tools/build_variant.py).
// Measurements: tools/microbench/seeds.cu, DESIGN.md ("arithmetic").
#ifndef MCD_NEWTON
#define MCD_NEWTON 2
#endif

// 1/x for normal positive x.  e = 1 - x*y0, 1/x = y0 (1 + e + e^2 + O(e^3)).
__device__ __forceinline__ double fast_rcp(double x) {
    const double y0 = rcp_seed(x);
    const double e = fma(-x, y0, 1.0);
#if MCD_NEWTON == 3
    const double p = fma(e, e, e);
    return fma(y0, p, y0);
#else
    return fma(y0, e, y0);
#endif
}

// x^(-1/2) for normal positive x.  e = 1 - x*y0^2, x^(-1/2) = y0 (1 + e/2 + 3 e^2/8 + O(e^3)).
__device__ __forceinline__ double fast_rsqrt(double x) {
    const double y0 = rsqrt_seed(x);
    const double t = x * y0;
    const double e = fma(-t, y0, 1.0);
#if MCD_NEWTON == 3
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    return fma(ye, p, y0);
#else
    // y0/2 by an exponent decrement on the integer pipe (y0 is a normal number far from the
    // denormal range: it is the rsqrt of a normal double)
    const double h = __hiloint2double(__double2hiint(y0) - 0x00100000, __double2loint(y0));
    return fma(h, e, y0);
#endif
}

// 2 x^(-1/2) for normal positive x: the quadratic Newton step in the form y0 (3 - x y0^2), which needs no halved
// seed (no integer-pipe work); the caller folds the factor 2 into a constant.  Same error as fast_rsqrt with
// MCD_NEWTON 2 (3/8 e^2, e = the seed's relative error 2^-20).
#ifndef MCD_RSQRT3
#define MCD_RSQRT3 1
#endif
__device__ __forceinline__ double rsqrt_twice(double x) {
#if MCD_RSQRT3
    const double y0 = rsqrt_seed(x);
    const double t = x * y0;
    const double g = fma(-t, y0, 3.0);
    return y0 * g;
#else
    const double y = fast_rsqrt(x);
    return y + y;
#endif
}

// FAST mixture variants: the packed verr^2 column and the walker's sigma^2 constants carry this factor, so that
// every reciprocal square root of the mixture term is rsqrt_twice of four times its argument -- no halved seed
// anywhere in the loop.  (1: round-2 arithmetic with mix_rsqrt; A/B builds.)
#if MCD_RSQRT3 && MCD_NEWTON == 2
constexpr double kMixVarScale = 4.0;
#else
constexpr double kMixVarScale = 1.0;
#endif

// 2^d for integer d <= 0; exact, flushed to zero below the normal range.
__device__ __forceinline__ double pow2_nonpos(int d) {
    const int hi = (d + 1023) << 20;
    return d < -1022 ? 0.0 : __hiloint2double(hi, 0);
}

// Running product kept as mantissa * 2^exponent: sum_i ln(x_i) without a log per factor.
// The exponent bookkeeping is integer-pipe work; the FP64 pipe sees one DMUL per factor.
struct LogProduct {
    double mant;        // product of the factors' mantissas since the last renormalisation
    long long expo;     // sum of the factors' unbiased exponents
    int bad;            // a factor of mul()/renormalise_checked() left the normal positive range (result: NaN)
    int signs;          // OR of the high words of the mul_ext() factors: sign bit set = a negative factor
    int expo32;         // renormalise_light(): exponents since the last fold() (32-bit: one add per group)
    unsigned range;     // renormalise_light(): largest (high word - 0x00100000) seen, as unsigned: >= 0x7fe00000
                        // means some group product left the normal positive range (result: NaN)

    __device__ __forceinline__ void reset() { mant = 1.0; expo = 0; bad = 0; signs = 0; expo32 = 0; range = 0u; }

    // multiply by x, a normal positive double
    __device__ __forceinline__ void mul(double x) {
        const int hi = __double2hiint(x);
        bad |= ((unsigned)(hi - 0x00100000) >= 0x7fe00000u);
        expo += (hi >> 20) - 1023;
        mant *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    }

    // Multiply by m * 2^e, m >= 0 (hot loop of the mixture kernels).  m is brought into [1, 2) by a
    // multiplication with 2^-(its exponent) instead of by rewriting its exponent bits, so the special
    // values take care of themselves: m == 0 makes the product 0 for good (lnlike = -inf), inf and NaN
    // make it NaN or negative, and ln() sorts those out once at the end -- no compare or select per
    // factor.  A negative m (a weight outside [0, 1]) is remembered through the sign bit of `signs`.
    // Denormal m stay exact (they only use up headroom below 1).
    __device__ __forceinline__ void mul_ext(double m, int e) {
        const int hi = __double2hiint(m);
        const int biased = (hi >> 20) & 0x7ff;
        signs |= hi;
        expo += (long long)(biased - 1023 + e);
        mant *= m * __hiloint2double((2046 - biased) << 20, 0);
    }

    // multiply by x without touching the exponent: the caller renormalises after a small group of
    // factors (each within 2^+-500 for a group of two), which moves the exponent bookkeeping from
    // once per factor to once per group
    __device__ __forceinline__ void mul_raw(double x) { mant *= x; }

    // the same for a product of mul_ext() factors, which may be zero (and must stay zero), inf or NaN
    __device__ __forceinline__ void renormalise_ext() {
        const int biased = (__double2hiint(mant) >> 20) & 0x7ff;
        expo += biased - 1023;
        mant *= __hiloint2double((2046 - biased) << 20, 0);
    }

    // fold the mantissa's own exponent into `expo`; call at least every 1000 mul() factors
    __device__ __forceinline__ void renormalise() {
        const int hi = __double2hiint(mant);
        expo += (hi >> 20) - 1023;
        mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(mant));
    }

    // the same after mul_raw(): additionally flags a product that left the normal positive range
    // (zero, denormal, negative, inf, NaN factor)
    __device__ __forceinline__ void renormalise_checked() {
        const int hi = __double2hiint(mant);
        bad |= ((unsigned)(hi - 0x00100000) >= 0x7fe00000u);
        expo += (hi >> 20) - 1023;
        mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(mant));
    }

    // The hot-loop form of renormalise_checked(): the exponent goes into a 32-bit sum and the range check into
    // a running maximum (6 integer instructions instead of 11); fold() moves both into `expo` / `bad` and must
    // be called at least every 2^20 calls and before ln().
    __device__ __forceinline__ void renormalise_light() {
        const int hi = __double2hiint(mant);
        range = max(range, (unsigned)(hi - 0x00100000));
        expo32 += (hi >> 20) - 1023;
        mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(mant));
    }
    __device__ __forceinline__ void fold() {
        expo += expo32;
        expo32 = 0;
        bad |= range >= 0x7fe00000u;
    }

    // ln of the product: -inf if it is exactly zero, NaN if a factor was bad or negative or the
    // product is not a normal positive number (inf, NaN, negative, denormal)
    __device__ __forceinline__ double ln() {
        const int hi = __double2hiint(mant);
        const bool is_zero = (hi | __double2loint(mant)) == 0;
        const bool out_of_range = (unsigned)(hi - 0x00100000) >= 0x7fe00000u;
        renormalise();
        const double r = fma((double)expo, kLn2, log(mant));
        return (bad || signs < 0 || (out_of_range && !is_zero))
                   ? __longlong_as_double(0x7ff8000000000000LL)
                   : (is_zero ? __longlong_as_double(0xfff0000000000000LL) : r);
    }
};

// Taylor coefficients of 2^f = exp(f ln 2): ln2^k / k!, k = 1..12.  In constant memory so that the
// FMAs of the polynomial take them as constant-bank operands (no register, no move per use).
__constant__ double kExp2Coef[12] = {
    6.93147180559945286e-01, 2.40226506959100694e-01, 5.55041086648215762e-02, 9.61812910762847687e-03,
    1.33335581464284411e-03, 1.54035303933816099e-04, 1.52527338040598403e-05, 1.32154867901443095e-06,
    1.01780860092396998e-07, 7.05491162080112333e-09, 4.44553827187081150e-10, 2.56784359934882051e-11};

// 2^f for |f| <= 0.5 (Taylor in f*ln2, degree 12: truncation 2e-16 relative).  Even and odd
// coefficients are two independent Horner chains in f^2 (depth 8 instead of 12 for one more
// instruction): the mixture kernels run at low occupancy and are latency-, not issue-bound.
__device__ __forceinline__ double exp2_frac(double f) {
    const double g = f * f;
    double even = kExp2Coef[11], odd = kExp2Coef[10];
    even = fma(even, g, kExp2Coef[9]);
    odd = fma(odd, g, kExp2Coef[8]);
    even = fma(even, g, kExp2Coef[7]);
    odd = fma(odd, g, kExp2Coef[6]);
    even = fma(even, g, kExp2Coef[5]);
    odd = fma(odd, g, kExp2Coef[4]);
    even = fma(even, g, kExp2Coef[3]);
    odd = fma(odd, g, kExp2Coef[2]);
    even = fma(even, g, kExp2Coef[1]);
    odd = fma(odd, g, kExp2Coef[0]);
    even = fma(even, g, 1.0);
    return fma(odd, f, even);
}

// exp(x) = mant * 2^expo for x <= 0 (moderate x > 0 works too); mant in [2^-0.5, 2^0.5].
// x <= -2^30 ln2 (including -inf) saturates to an exponent that flushes to zero downstream;
// NaN or x >= 2^30 ln2 sets `invalid`.  All range handling is integer-pipe work on the high
// word, the FP64 pipe only sees the 15 multiply-adds.
__device__ __forceinline__ void exp_split(double x, double &mant, int &expo, int &invalid) {
    const double kMagic = 6755399441055744.0;            // 1.5 * 2^52: round-to-nearest-integer trick
    const double t = x * kLog2e;
    const int thi = __double2hiint(t);
    const unsigned uthi = (unsigned)thi & 0x7fffffffu;
    const bool huge = uthi >= 0x41d00000u;               // |t| >= 2^30, inf or NaN
    const bool isnan = uthi > 0x7ff00000u || (uthi == 0x7ff00000u && __double2loint(t) != 0);
    const double shifted = t + kMagic;
    const int n = __double2loint(shifted);
    const double nf = shifted - kMagic;
    // f = x*log2e - n with the rounding error of the product recovered by the fma
    const double f = fma(x, kLog2e, -nf);
    const double m = exp2_frac(f);
    expo = huge ? -(1 << 30) : n;
    mant = huge ? 1.0 : m;
    invalid |= (huge && (thi >= 0 || isnan)) ? 1 : 0;
}

// exp(-w/2) = mant * 2^expo for w >= 0 (w = z^2 of a Gaussian), the hot-loop variant of exp_split:
// no range checks at all.  w is clamped to 2^31 so that the exponent always fits an int32 (a star
// 46 000 sigma away contributes exp(-1e9), i.e. nothing, either way); NaN propagates into the
// mantissa and is caught where the mantissa enters the running product.
__device__ __forceinline__ void exp_neg_half(double w, double &mant, int &expo) {
    const double kMagic = 6755399441055744.0;            // 1.5 * 2^52
    const double kC = -0.5 * kLog2e;
    const double wc = fmin(w, 2147483648.0);
    const double shifted = fma(wc, kC, kMagic);
    expo = __double2loint(shifted);
    const double nf = shifted - kMagic;
    const double f = fma(wc, kC, -nf);                  // exact product minus the integer part
    mant = exp2_frac(f);
}

// Table-driven variant for the hot loops of the mixture kernels (the FP64 pipe's issue rate is their
// bound, so instructions count, not latency): 2^t with t*64 = N + r, |r| <= 1/2, N = 64 expo + j,
//     exp(-w/2) = 2^expo * T[j] * (1 + q(r)),   T[j] = 2^(j/64),   q(r) = 2^(r/64) - 1  (degree 5, 4e-17)
// -- 10 FP64 instructions instead of 17.  `table` is a copy of kExp2Table in shared memory (64 lanes'
// worth of divergent indices: constant memory would serialise).  mant in [0.99, 2.01).
__constant__ double kExp2Table[64] = {
    1.00000000000000000e+00, 1.01088928605170048e+00, 1.02189714865411663e+00, 1.03302487902122841e+00,
    1.04427378242741375e+00, 1.05564517836055716e+00, 1.06714040067682370e+00, 1.07876079775711986e+00,
    1.09050773266525769e+00, 1.10238258330784089e+00, 1.11438674259589243e+00, 1.12652161860824185e+00,
    1.13878863475669156e+00, 1.15118922995298267e+00, 1.16372485877757748e+00, 1.17639699165028122e+00,
    1.18920711500272103e+00, 1.20215673145270308e+00, 1.21524735998046896e+00, 1.22848053610687002e+00,
    1.24185781207348400e+00, 1.25538075702469110e+00, 1.26905095719173322e+00, 1.28287001607877826e+00,
    1.29683955465100964e+00, 1.31096121152476441e+00, 1.32523664315974132e+00, 1.33966752405330292e+00,
    1.35425554693689265e+00, 1.36900242297459052e+00, 1.38390988196383202e+00, 1.39897967253831124e+00,
    1.41421356237309515e+00, 1.42961333839197002e+00, 1.44518080697704665e+00, 1.46091779418064704e+00,
    1.47682614593949935e+00, 1.49290772829126484e+00, 1.50916442759342284e+00, 1.52559815074453842e+00,
    1.54221082540794074e+00, 1.55900440023783693e+00, 1.57598084510788650e+00, 1.59314215134226700e+00,
    1.61049033194925428e+00, 1.62802742185734783e+00, 1.64575547815396495e+00, 1.66367658032673638e+00,
    1.68179283050742900e+00, 1.70010635371852348e+00, 1.71861929812247793e+00, 1.73733383527370622e+00,
    1.75625216037329945e+00, 1.77537649252652119e+00, 1.79470907500310717e+00, 1.81425217550039886e+00,
    1.83400808640934243e+00, 1.85397912508338547e+00, 1.87416763411029996e+00, 1.89457598158696561e+00,
    1.91520656139714740e+00, 1.93606179349229435e+00, 1.95714412417540018e+00, 1.97845602638795093e+00};
// (ln2 / 64)^k / k!, k = 1..5
__constant__ double kExp2StepCoef[5] = {1.08304246962491451e-02, 5.86490495505616997e-05, 2.11731371554647763e-07,
                                        5.73285168864040189e-10, 1.24178437017169253e-12};

__device__ __forceinline__ void exp_neg_half_table(double w, const double *__restrict__ table, double &mant, int &expo) {
    const double kMagic = 6755399441055744.0;            // 1.5 * 2^52
    const double kC = -32.0 * kLog2e;                    // -1/2 * log2(e) * 64
    // clamp w to 2^25 on the integer pipe (w >= 0: the high words order like the doubles), so that N
    // stays inside an int32; a NaN with the sign bit clear is clamped too -- the caller's mantissa is
    // NaN through its other factors then
    const double wc = __hiloint2double(min(__double2hiint(w), 0x41800000), __double2loint(w));
    const double shifted = fma(wc, kC, kMagic);
    const int N = __double2loint(shifted);
    const double nf = shifted - kMagic;
    const double r = fma(wc, kC, -nf);                   // exact product minus the integer part
    double p = fma(kExp2StepCoef[4], r, kExp2StepCoef[3]);
    p = fma(p, r, kExp2StepCoef[2]);
    p = fma(p, r, kExp2StepCoef[1]);
    p = fma(p, r, kExp2StepCoef[0]);
    const double t = table[N & 63];
    expo = N >> 6;
    mant = fma(t, p * r, t);
}

// ---- mixture fast path -------------------------------------------------------------------------
// exp(-z^2/2) as ONE plain double, for u = z * kExpArgScale (kExpArgScale^2 = T log2(e) / 2 for a table of T
// entries, so that -z^2/2 * log2(e) * T = -u^2 and the square is formed inside the two FMAs that split it
// into integer and fraction -- no separate multiplication, and the walker constants carry the scale for free).
// |u| is clamped on the integer pipe (kMixClampHi), i.e. the result never drops below 2^-1010 and its exponent
// can be set by an integer addition on the high word: the caller only takes this path when the other
// mixture component is at least 2^-480, where a member term of 2^-1010 and one of 2^-100000 are the same
// thing.  A NaN u is clamped too (the NaN reaches the factor through y = norm^-1/2 or the walker's
// `bad` flag).
// MCD_MIX_LEAN selects the arithmetic of the mixture kernels (same tolerance argument as MCD_NEWTON above):
//   2 (default): 1024-entry table + quadratic polynomial (truncation 6.5e-12, zero-mean) and quadratic Newton
//      steps (1.3e-12); the table unit is folded into kExpArgScale, so the argument needs no rescaling:
//      26 FP64 instructions per term of the fixed-background mixture, 42 with the fitted Gaussian background.
//   1: 256-entry table + cubic polynomial (1.4e-13), the argument doubled in the loop (28 / 46; round 2's
//      first version: +14 % over 0 on the 2e6-star mixture workload, 2.87 -> 2.51 ms).
//   0: 64-entry table + degree-5 polynomial (4e-17) and cubic Newton steps (34 / 59).
#ifndef MCD_MIX_LEAN
#define MCD_MIX_LEAN 2
#endif
#if MCD_MIX_LEAN == 2
constexpr int kMixTableBits = 10;
constexpr double kExpArgScale = 27.17829760921661;      // sqrt(512 / ln 2): -z^2/2 log2(e) 1024 = -u^2
constexpr int kMixClampHi = 0x408fc000;                 // |u| <= 1016: the result never drops below 2^-1008.1
#elif MCD_MIX_LEAN
constexpr int kMixTableBits = 8;
constexpr double kExpArgScale = 6.7945744023041525;     // sqrt(32 / ln 2), doubled in exp_neg_sq_split
constexpr int kMixClampHi = 0x406fc000;                 // |u| <= 254
#else
constexpr int kMixTableBits = 6;
constexpr double kExpArgScale = 6.7945744023041525;     // sqrt(32 / ln 2)
constexpr int kMixClampHi = 0x406fc000;                 // |u| <= 254
#endif
constexpr int kMixTableSize = 1 << kMixTableBits;
constexpr int kMixFastFlag = (int)0x80000000;           // exponent-column value of a fast-path star
constexpr int kMixComfort = 200;                        // fast path: |log2(background term)| <= this
constexpr int kMixSlowExp = -200;                       // fitted background: both components below 2^this -> slow path

// (ln2 / 256)^k / k!, k = 1..3, and (ln2 / 1024)^k / k!, k = 1..2
__constant__ double kExp2LeanCoef[3] = {0.0027076061740622863, 3.6655655969101062e-06, 3.3083026805413713e-09};
__constant__ double kExp2Lean2Coef[2] = {0.0006769015435155716, 2.290978498068816e-07};

// returns the mantissa in [0.99, 2.01); N = table-units exponent: exp(-z^2/2) = mant * 2^(N >> kMixTableBits)
// `table` is the 32-bit shared-memory address of 2^(j / kMixTableSize), j = 0 .. kMixTableSize - 1 (a generic
// pointer costs four uniform-datapath instructions per load to rebuild the shared window base).
__device__ __forceinline__ double exp_neg_sq_split(double u, uint32_t table, int &N) {
    const double kMagic = 6755399441055744.0;            // 1.5 * 2^52
    const int hi = min(__double2hiint(u) & 0x7fffffff, kMixClampHi);
    const double uc = __hiloint2double(hi, __double2loint(u));
#if MCD_MIX_LEAN == 2
    const double shifted = fma(-uc, uc, kMagic);
    N = __double2loint(shifted);
    const double nf = shifted - kMagic;
    const double r = fma(-uc, uc, -nf);                  // exact square minus its integer part, |r| <= 1/2
    const double p = fma(kExp2Lean2Coef[1], r, kExp2Lean2Coef[0]);
#elif MCD_MIX_LEAN
    // table units of 1/256: -u^2 * 4
    const double s = uc * 2.0;                            // exact
    const double shifted = fma(-s, s, kMagic);
    N = __double2loint(shifted);
    const double nf = shifted - kMagic;
    const double r = fma(-s, s, -nf);
    double p = fma(kExp2LeanCoef[2], r, kExp2LeanCoef[1]);    // constant-bank operands: no register, no move
    p = fma(p, r, kExp2LeanCoef[0]);
#else
    const double shifted = fma(-uc, uc, kMagic);
    N = __double2loint(shifted);
    const double nf = shifted - kMagic;
    const double r = fma(-uc, uc, -nf);                  // exact square minus its integer part
    double p = fma(kExp2StepCoef[4], r, kExp2StepCoef[3]);
    p = fma(p, r, kExp2StepCoef[2]);
    p = fma(p, r, kExp2StepCoef[1]);
    p = fma(p, r, kExp2StepCoef[0]);
#endif
    double t;
    asm("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(table + ((uint32_t)(N & (kMixTableSize - 1)) << 3)));
    return fma(t, p * r, t);
}

// mant * 2^(N >> kMixTableBits) by an integer addition on the exponent field (N >= -64 * 1010: stays normal)
__device__ __forceinline__ double scale_by_table_exponent(double mant, int N) {
    return __hiloint2double(__double2hiint(mant) + ((N >> kMixTableBits) << 20), __double2loint(mant));
}

// x^(-1/2) for the mixture fast path: cubic Newton step, or quadratic under MCD_MIX_LEAN
__device__ __forceinline__ double mix_rsqrt(double x) {
#if MCD_MIX_LEAN
    const double y0 = rsqrt_seed(x);
    const double t = x * y0;
    const double e = fma(-t, y0, 1.0);
    const double h = __hiloint2double(__double2hiint(y0) - 0x00100000, __double2loint(y0));
    return fma(h, e, y0);
#else
    return fast_rsqrt(x);
#endif
}

// x^(-1/2) of a variance that carries kMixVarScale
__device__ __forceinline__ double mix_var_rsqrt(double x_scaled) {
    if constexpr (kMixVarScale == 4.0 && MCD_MIX_LEAN != 0) return rsqrt_twice(x_scaled);
    else return mix_rsqrt(x_scaled * (1.0 / kMixVarScale));
}

// 2^d for d <= 0, exactly zero below 2^-960 (integer pipe only)
__device__ __forceinline__ double pow2_flush(int d) {
    const int hi = (d + 1023) << 20;
    return __hiloint2double(d < -960 ? 0 : hi, 0);
}

// a_m*2^a_e + b_m*2^b_e as (mantissa, exponent); both mantissas >= 0 and O(1).
// Both terms are scaled to the larger exponent; the smaller one is flushed to zero beyond 2^-960 --
// the analogue of exp() underflowing in the reference's max-shifted form (analysis/runner.py:280-286:
// a component more than ~745 nats below the other contributes exactly nothing, and if the other one
// has zero weight the star's likelihood is 0, i.e. lnlike = -inf).  Branch- and select-free.
__device__ __forceinline__ void ext_add(double a_m, int a_e, double b_m, int b_e, double &m, int &e) {
    e = max(a_e, b_e);
    m = fma(a_m, pow2_flush(a_e - e), b_m * pow2_flush(b_e - e));
}

}  // namespace mcd

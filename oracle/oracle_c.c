/*
 * Plain-C restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
 * a second, independently written checker beside oracle/reference_np.py, and the compiled CPU
 * baseline of bench.py.  Literal: like the reference it re-evaluates the sky -> tangent-plane
 * geometry with sin/cos/atan2 for every star on every call, takes one log and one division per
 * term, and uses the max-shifted two-exp form of the mixture.  Walkers are independent: the outer
 * loop is parallelised over walkers with OpenMP, which is what the reference's process pool does
 * (analysis/runner.py:398-403).
 *
 * All paths relative to /root/reference/mcmc_dynamics.  Units as in the reference's defaults:
 * velocities km/s, centre in degrees, a and r_peak in THEIR OWN unit with the factor to arcmin
 * passed in (astropy's implicit conversion).
 *
 * Built by __graft_entry__.build_oracle():  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads of the walker loops; returns what is in force (torchrun exports OMP_NUM_THREADS=1, bench.py
 * asks for the cores it reports) */
int oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define TWO_PI 6.283185307179586476925286766559
#define DEG (3.14159265358979323846264338327950288 / 180.0)

/* parameter vector of one walker, every model parameter resolved (fixed ones merged in) */
typedef struct {
    double v_sys, sigma_max, v_maxx, v_maxy, ra_center, dec_center, a, r_peak, v_back, sigma_back, f_back;
} oracle_params;

/* utils/coordinates/calc_xy_offset.py:30-31, arcmin */
static void xy_offset(double ra, double dec, double ra_c, double dec_c, double *dx, double *dy) {
    const double r0 = 10800. / 3.14159265358979323846264338327950288;
    const double d = dec * DEG, dc = dec_c * DEG, dra = (ra - ra_c) * DEG;
    *dx = -r0 * cos(d) * sin(dra);
    *dy = r0 * (sin(d) * cos(dc) - cos(d) * sin(dc) * cos(dra));
}

/* one star's member log-likelihood and, through *out_norm, nothing else: runner.py:261-271 */
static double member_lnlike(int radial, const oracle_params *p, double a_to_arcmin, double rp_to_arcmin, double ra,
                            double dec, double v, double verr) {
    double dx, dy, v_los, sigma_los;
    xy_offset(ra, dec, p->ra_center, p->dec_center, &dx, &dy);
    const double theta = atan2(dy, dx);
    const double v_max = sqrt(p->v_maxx * p->v_maxx + p->v_maxy * p->v_maxy);
    const double theta_0 = atan2(p->v_maxy, p->v_maxx);
    if (!radial) {
        v_los = p->v_sys + v_max * sin(theta - theta_0);                     /* constant.py:109-111 */
        sigma_los = p->sigma_max;                                            /* constant.py:74 */
    } else {
        /* the reference evaluates calc_xy_offset a second time in dispersion_model (model.py:126) */
        double dx2, dy2;
        xy_offset(ra, dec, p->ra_center, p->dec_center, &dx2, &dy2);
        const double r = sqrt(dx * dx + dy * dy), r_disp = sqrt(dx2 * dx2 + dy2 * dy2);
        const double x_pa = r * sin(theta - theta_0);
        const double lin = 1.0 / rp_to_arcmin;                               /* arcmin per unit(r_peak) */
        const double ratio = (r / p->r_peak) * lin;
        v_los = p->v_sys + (2. * (v_max / p->r_peak) * x_pa / (1. + ratio * ratio)) * lin;   /* model.py:180 */
        const double ra2 = (r_disp / p->a) / a_to_arcmin;
        sigma_los = p->sigma_max / pow(1. + ra2 * ra2, 0.25);               /* model.py:128 */
    }
    const double norm = verr * verr + sigma_los * sigma_los;
    const double exponent = -0.5 * (v - v_los) * (v - v_los) / norm;
    return -0.5 * log(TWO_PI * norm) + exponent;
}

/* runner.py:279-284, constant.py:320-323, model.py:452-454,614-618 */
static double mixture(double lm, double lb, double m) {
    const double mx = lm > lb ? lm : lb;
    return mx + log(m * exp(lm - mx) + (1. - m) * exp(lb - mx));
}

/*
 * lnlike of `nw` walkers.
 *   radial     0: ConstantFit kinematics, 1: ModelFit kinematics
 *   background 0: none (runner.py:264-271)                1: fixed lbg + pmember (runner.py:272-286)
 *              2: fixed lbg + density, f_back (model.py:586-623)
 *              3: fitted Gaussian background + density, f_back (constant.py:326-364, model.py:414-456)
 */
void oracle_lnlike(int radial, int background, long n, const double *ra, const double *dec, const double *v,
                   const double *verr, const double *pmember, const double *density, const double *lbg, int nw,
                   const oracle_params *params, double a_to_arcmin, double rp_to_arcmin, double *out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < nw; ++w) {
        const oracle_params *p = &params[w];
        double sum = 0.0;
        for (long i = 0; i < n; ++i) {
            const double lm = member_lnlike(radial, p, a_to_arcmin, rp_to_arcmin, ra[i], dec[i], v[i], verr[i]);
            if (background == 0) {
                sum += lm;
            } else if (background == 1) {
                sum += mixture(lm, lbg[i], pmember[i]);
            } else {
                const double m = density[i] / (density[i] + p->f_back);
                double lb;
                if (background == 2) {
                    lb = lbg[i];
                } else {
                    const double nb = verr[i] * verr[i] + p->sigma_back * p->sigma_back;
                    lb = -0.5 * log(TWO_PI * nb) + -0.5 * (v[i] - p->v_back) * (v[i] - p->v_back) / nb;
                }
                sum += mixture(lm, lb, m);
            }
        }
        out[w] = sum;
    }
}

/* background/gaussian.py:23-28 */
void oracle_gaussian_background(long n, const double *v, const double *verr, double mean, double sigma, double *out) {
    for (long i = 0; i < n; ++i) {
        const double norm = verr[i] * verr[i] + sigma * sigma;
        out[i] = -0.5 * log(TWO_PI * norm) + -0.5 * (v[i] - mean) * (v[i] - mean) / norm;
    }
}

/* background/single_stars.py:72-77 */
void oracle_single_stars_background(long m, const double *v_bg, long n, const double *v, const double *verr,
                                    double sigma_int, double *out) {
#pragma omp parallel for
    for (long i = 0; i < n; ++i) {
        const double norm = sigma_int * sigma_int + verr[i] * verr[i];
        double best = -INFINITY;
        for (long j = 0; j < m; ++j) {
            const double e = -(v_bg[j] - v[i]) * (v_bg[j] - v[i]) / (2. * norm);
            if (e > best) best = e;
        }
        double s = 0.0;
        for (long j = 0; j < m; ++j) {
            const double e = -(v_bg[j] - v[i]) * (v_bg[j] - v[i]) / (2. * norm);
            s += exp(e - best) / sqrt(TWO_PI * norm);
        }
        out[i] = best + log(s) - log((double)m);
    }
}

/*
 * ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the serial loops the reference JIT-compiles through weave.inline
 * (SURVEY.md section 2a), single-threaded exactly as the reference runs them (it passes
 * -fopenmp but contains no "#pragma omp").  Used (1) to cross-check the NumPy restatement in
 * oracle/operators.py, (2) as the timed CPU baseline of bench.py (cpu_baseline.kind = "port").
 * Paths cited are relative to /root/reference.  Built by oracle/cloops.py with gcc -O3.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* interfaces/linearoperators.py:368-375 (pol=1), :424-430 (pol=2), :483-489 (pol=3) */
void orc_pointing_mult(const int64_t *pix, const double *c, const double *s, int64_t nt, int pol,
                       const double *v, double *x)
{
    int64_t i;
    if (pol == 1) {
        for (i = 0; i < nt; ++i) { if (pix[i] == -1) continue; x[i] += v[pix[i]]; }
    } else if (pol == 2) {
        for (i = 0; i < nt; ++i) {
            if (pix[i] == -1) continue;
            x[i] += v[2 * pix[i]] * c[i] + v[2 * pix[i] + 1] * s[i];
        }
    } else {
        for (i = 0; i < nt; ++i) {
            if (pix[i] == -1) continue;
            x[i] += v[3 * pix[i]] + v[3 * pix[i] + 1] * c[i] + v[3 * pix[i] + 2] * s[i];
        }
    }
}

/* interfaces/linearoperators.py:394-401 (pol=1), :447-454 (pol=2), :509-517 (pol=3) */
void orc_pointing_rmult(const int64_t *pix, const double *c, const double *s, int64_t nt, int pol,
                        const double *v, double *x)
{
    int64_t i;
    if (pol == 1) {
        for (i = 0; i < nt; ++i) { if (pix[i] == -1) continue; x[pix[i]] += v[i]; }
    } else if (pol == 2) {
        for (i = 0; i < nt; ++i) {
            if (pix[i] == -1) continue;
            x[2 * pix[i]] += v[i] * c[i];
            x[2 * pix[i] + 1] += v[i] * s[i];
        }
    } else {
        for (i = 0; i < nt; ++i) {
            if (pix[i] == -1) continue;
            x[3 * pix[i]] += v[i];
            x[3 * pix[i] + 1] += v[i] * c[i];
            x[3 * pix[i] + 2] += v[i] * s[i];
        }
    }
}

/* utilities/process_ces.py:480-486, 505-513, 527-538: weighted per-pixel moments */
void orc_moments(const int64_t *pix, const double *w, const double *c, const double *s, int64_t nt,
                 int pol, double *counts, double *cosine, double *sine, double *cos2, double *sin2,
                 double *sincos)
{
    int64_t i, p;
    for (i = 0; i < nt; ++i) {
        p = pix[i];
        if (p == -1) continue;
        if (pol == 1) { counts[p] += w[i]; continue; }
        if (pol == 3) {
            counts[p] += w[i];
            cosine[p] += w[i] * c[i];
            sine[p] += w[i] * s[i];
        }
        cos2[p] += w[i] * c[i] * c[i];
        sin2[p] += w[i] * s[i] * s[i];
        sincos[p] += w[i] * s[i] * c[i];
    }
}

/* utilities/process_ces.py:411-417 */
void orc_relabel(int64_t *pix, const int64_t *old2new, int64_t nt)
{
    int64_t i;
    for (i = 0; i < nt; ++i) { if (pix[i] == -1) continue; pix[i] = old2new[pix[i]]; }
}

/* interfaces/linearoperators.py:796-806 (pol=3) and :822-831 (pol=2); det and mask as :792-795 */
void orc_bd_apply(int64_t npix, int pol, const double *hits, const double *c, const double *s,
                  const double *c2, const double *s2, const double *cs, const double *x, double *y)
{
    int64_t j;
    if (pol == 1) {
        for (j = 0; j < npix; ++j) y[j] = hits[j] > 0 ? x[j] / hits[j] : 0.0;
    } else if (pol == 2) {
        for (j = 0; j < npix; ++j) {
            double det = (c2[j] * s2[j]) - (cs[j] * cs[j]);
            if (fabs(det) > 1e-5) {
                y[2 * j] = (s2[j] * x[2 * j] - cs[j] * x[2 * j + 1]) / det;
                y[2 * j + 1] = (-cs[j] * x[2 * j] + c2[j] * x[2 * j + 1]) / det;
            } else { y[2 * j] = 0.0; y[2 * j + 1] = 0.0; }
        }
    } else {
        for (j = 0; j < npix; ++j) {
            double det = hits[j] * (c2[j] * s2[j] - cs[j] * cs[j]) - c[j] * c[j] * s2[j]
                         - s[j] * s[j] * c2[j] + 2. * c[j] * s[j] * cs[j];
            if (fabs(det) > 1e-5) {
                y[3*j]   = ((c2[j]*s2[j]-cs[j]*cs[j])*x[3*j] + (s[j]*cs[j]-c[j]*s2[j])*x[3*j+1] + (c[j]*cs[j]-s[j]*c2[j])*x[3*j+2]) / det;
                y[3*j+1] = ((s[j]*cs[j]-c[j]*s2[j])*x[3*j] + (hits[j]*s2[j]-s[j]*s[j])*x[3*j+1] + (s[j]*c[j]-hits[j]*cs[j])*x[3*j+2]) / det;
                y[3*j+2] = ((c[j]*cs[j]-s[j]*c2[j])*x[3*j] + (-hits[j]*cs[j]+c[j]*s[j])*x[3*j+1] + (hits[j]*c2[j]-c[j]*c[j])*x[3*j+2]) / det;
            } else { y[3*j] = 0.0; y[3*j+1] = 0.0; y[3*j+2] = 0.0; }
        }
    }
}

/* interfaces/linearoperators.py:587-595: y = a0 v ; y[:-k] += a_k v[k:] ; y[k:] += a_k v[:-k].
 * Same lag-major accumulation order as the NumPy loop of the reference. */
void orc_toeplitz(const double *a, int64_t na, const double *v, int64_t n, double *y)
{
    int64_t k, j;
    for (j = 0; j < n; ++j) y[j] = a[0] * v[j];
    for (k = 1; k < na && k < n; ++k) {
        const double ak = a[k];
        for (j = 0; j < n - k; ++j) y[j] += ak * v[j + k];
        for (j = k; j < n; ++j) y[j] += ak * v[j - k];
    }
}

/* interfaces/linearoperators.py:141-156: masked mean over [0, n); returns sum, writes count */
double orc_seq_sum(const double *d, int64_t n)
{
    double acc = 0.;
    int64_t j;
    for (j = 0; j < n; ++j) acc += d[j];
    return acc;
}

/* interfaces/linearoperators.py:129-168 for one (start,end) list already expanded by the caller */
void orc_filter_offset(const int64_t *pix, const double *d, double *out, const int64_t *start,
                       const int64_t *end, int64_t nseg)
{
    int64_t k, j;
    for (k = 0; k < nseg; ++k) {
        double mean = 0., counter = 0.;
        for (j = start[k]; j < end[k]; ++j) {
            if (pix[j] == -1) continue;
            mean += d[j];
            counter += 1.;
        }
        mean = mean / counter;
        if (isinf(mean) || isnan(mean)) continue;
        for (j = start[k]; j < end[k]; ++j) out[j] = d[j] - mean;
    }
}

/* SciPy-style PCG vector work on the host (scipy/sparse/linalg/_isolve/iterative.py:405-431),
 * used only by the CPU baseline so that its vector updates are not NumPy-temporary bound. */
double orc_dot(const double *a, const double *b, int64_t n)
{
    double acc = 0.;
    int64_t j;
    for (j = 0; j < n; ++j) acc += a[j] * b[j];
    return acc;
}

/* CPU oracle (plain C restatement) of the HPX-FFT 2-D r2c hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library; the product path
 * (hpx-fft_b200/csrc) never links or calls it.
 *
 * Restated (citations relative to /root/reference):
 *   hpxfft::shared::loop::initialize      core/src/shared/loop.cpp:158-189  (dimension inference :163-165)
 *   hpxfft::shared::loop::fft_2d_r2c_par  core/src/shared/loop.cpp:56-113   (4 parallel phases, 5 timers)
 *   fft_1d_r2c_inplace / fft_1d_c2c_inplace  loop.cpp:6-15  ->  core/src/util/adapter_fftw.cpp:12-15, 32-35
 *   transpose_shared_y_to_x / _x_to_y     loop.cpp:18-25, 46-53  (element-wise strided copies, as written there)
 *   for_loop(par, 0, n, f)                loop.cpp:61-102  ->  `#pragma omp parallel for` over the same index
 *
 * The arithmetic is FFTW 3.3.10's (un-vendored dependency, spack-repo/environments/hpxfft_ci.yaml:4),
 * absent from this image.  Its published definition is restated with a self-contained Stockham
 * mixed-radix FFT:  forward, unnormalised  Y[k] = sum_j x[j] exp(-2 pi i j k / n); r2c keeps k=0..n/2
 * and is computed in place through the half-length complex transform (the layout fftw_execute_dft_r2c
 * uses on a padded row, adapter_fftw.cpp:14).
 *
 * PINNING: checked in tests/test_oracle.py against the reference's single known-answer vector
 * (test/src/test_shared_loop.cpp:15-34,53: 4x6 rows [1,2,3,4,0,0] -> row0 [40,0,-8,8,-8,0], exact ==)
 * and against the independent pocketfft restatement in oracle/oracle.py.  Beyond that vector the
 * reference's tests leave parity unpinned.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

typedef struct {
    int n;
    int npass;
    int radix[64];
    cplx *tw[64];   /* tw[p][k*R + r] = exp(-2 pi i r k / (Ns R)),  k < Ns */
    cplx *root[64]; /* root[p][q]     = exp(-2 pi i q / R)  (generic odd radix) */
} plan1d;

static const long double PI_L = 3.14159265358979323846264338327950288L;

static cplx expi_l(long double num, long double den)
{ /* exp(-2 pi i num/den), exact on the axes */
    long double f = fmodl(num, den) / den; /* in [0,1) */
    long double e = 8.0L * f;
    if (e == floorl(e)) {
        static const double s2 = 0.70710678118654752440084436210484903928L;
        switch ((int) e) {
        case 0: return 1.0;
        case 1: return s2 - s2 * I;
        case 2: return -1.0 * I;
        case 3: return -s2 - s2 * I;
        case 4: return -1.0;
        case 5: return -s2 + s2 * I;
        case 6: return 1.0 * I;
        case 7: return s2 + s2 * I;
        }
    }
    long double a = -2.0L * PI_L * f;
    return (double) cosl(a) + (double) sinl(a) * I;
}

static void plan1d_init(plan1d *p, int n)
{
    p->n = n;
    p->npass = 0;
    int m = n;
    while (m % 4 == 0) { p->radix[p->npass++] = 4; m /= 4; }
    while (m % 2 == 0) { p->radix[p->npass++] = 2; m /= 2; }
    for (int f = 3; m > 1; f += 2)
        while (m % f == 0) { p->radix[p->npass++] = f; m /= f; }
    int Ns = 1;
    for (int q = 0; q < p->npass; ++q) {
        int R = p->radix[q];
        p->tw[q] = (cplx *) malloc(sizeof(cplx) * (size_t) Ns * R);
        for (int k = 0; k < Ns; ++k)
            for (int r = 0; r < R; ++r)
                p->tw[q][(size_t) k * R + r] = expi_l((long double) r * k, (long double) Ns * R);
        p->root[q] = (cplx *) malloc(sizeof(cplx) * R);
        for (int r = 0; r < R; ++r) p->root[q][r] = expi_l(r, R);
        Ns *= R;
    }
}

static void plan1d_free(plan1d *p)
{
    for (int q = 0; q < p->npass; ++q) { free(p->tw[q]); free(p->root[q]); }
    p->npass = 0;
}

/* forward c2c of length n; data in x, scratch y (both n); result left in x */
static void fft1d_exec(const plan1d *p, cplx *x, cplx *y)
{
    const int n = p->n;
    cplx *in = x, *out = y;
    int Ns = 1;
    for (int q = 0; q < p->npass; ++q) {
        const int R = p->radix[q];
        const int T = n / R;
        const cplx *tw = p->tw[q];
        const int nblk = T / Ns;
        if (R == 4) {
            for (int b = 0; b < nblk; ++b) {
                const cplx *i0 = in + (size_t) b * Ns;
                cplx *o0 = out + (size_t) b * Ns * 4;
                for (int k = 0; k < Ns; ++k) {
                    cplx a0 = i0[k];
                    cplx a1 = i0[k + T] * tw[4 * k + 1];
                    cplx a2 = i0[k + 2 * T] * tw[4 * k + 2];
                    cplx a3 = i0[k + 3 * T] * tw[4 * k + 3];
                    cplx s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
                    cplx md13 = cimag(d13) - creal(d13) * I; /* -i * d13 */
                    o0[k] = s02 + s13;
                    o0[k + Ns] = d02 + md13;
                    o0[k + 2 * Ns] = s02 - s13;
                    o0[k + 3 * Ns] = d02 - md13;
                }
            }
        } else if (R == 2) {
            for (int b = 0; b < nblk; ++b) {
                const cplx *i0 = in + (size_t) b * Ns;
                cplx *o0 = out + (size_t) b * Ns * 2;
                for (int k = 0; k < Ns; ++k) {
                    cplx a0 = i0[k];
                    cplx a1 = i0[k + T] * tw[2 * k + 1];
                    o0[k] = a0 + a1;
                    o0[k + Ns] = a0 - a1;
                }
            }
        } else {
            const cplx *root = p->root[q];
            cplx v[R];
            for (int b = 0; b < nblk; ++b)
                for (int k = 0; k < Ns; ++k) {
                    for (int r = 0; r < R; ++r) v[r] = in[(size_t) b * Ns + k + (size_t) r * T] * tw[(size_t) k * R + r];
                    for (int s = 0; s < R; ++s) {
                        cplx acc = v[0];
                        for (int r = 1; r < R; ++r) acc += v[r] * root[(r * s) % R];
                        out[(size_t) b * Ns * R + k + (size_t) s * Ns] = acc;
                    }
                }
        }
        cplx *t = in; in = out; out = t;
        Ns *= R;
    }
    if (in != x) memcpy(x, in, sizeof(cplx) * (size_t) n);
}

/* in-place r2c of a padded row: n reals (+2 pad) -> n/2+1 complex, n even.
 * half-length complex FFT + Hermitian split. */
static void r2c_row(const plan1d *half, const cplx *wn, double *row, int n, cplx *scratch)
{
    const int m = n / 2;
    cplx *z = (cplx *) row;
    if (m >= 1) fft1d_exec(half, z, scratch);
    cplx z0 = z[0];
    for (int k = 1; k <= m / 2; ++k) {
        cplx a = z[k], b = z[m - k];
        cplx e = 0.5 * (a + conj(b)), o = 0.5 * (a - conj(b));
        cplx eb = 0.5 * (b + conj(a)), ob = 0.5 * (b - conj(a));
        cplx mi_o = cimag(o) - creal(o) * I;    /* -i * o  */
        cplx mi_ob = cimag(ob) - creal(ob) * I; /* -i * ob */
        z[k] = e + wn[k] * mi_o;
        if (m - k != k) z[m - k] = eb + wn[m - k] * mi_ob;
    }
    z[0] = creal(z0) + cimag(z0);
    z[m] = creal(z0) - cimag(z0);
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

int hpxfft_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* shared::loop: vals is n_row x n_col doubles (n_col = 2*(ny/2+1)), transformed in place.
 * timings[5] = total, first_fftw, first_trans, second_fftw, second_trans (seconds), may be NULL.
 * returns 0, or -1 on bad arguments / allocation failure. */
int hpxfft_oracle_shared_loop(double *vals, size_t n_row, size_t n_col, int nthreads, double *timings)
{
    if (!vals || n_row == 0 || n_col < 4 || (n_col & 1)) return -1;
    /* core/src/shared/loop.cpp:163-165 */
    const size_t dim_c_x = n_row, dim_c_y = n_col / 2, dim_r_y = 2 * dim_c_y - 2;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    /* loop.cpp:167  trans_values_vec_(dim_c_y_, 2*dim_c_x_) */
    cplx *trans = (cplx *) malloc(sizeof(cplx) * dim_c_y * dim_c_x);
    if (!trans) return -1;
    plan1d p_half, p_x;
    plan1d_init(&p_half, (int) (dim_r_y / 2));
    plan1d_init(&p_x, (int) dim_c_x);
    cplx *wn = (cplx *) malloc(sizeof(cplx) * (dim_r_y / 2 + 1));
    for (size_t k = 0; k <= dim_r_y / 2; ++k) wn[k] = expi_l((long double) k, (long double) dim_r_y);
    size_t smax = dim_c_x > dim_r_y / 2 ? dim_c_x : dim_r_y / 2;
    cplx *scratch = (cplx *) malloc(sizeof(cplx) * smax * (size_t) nthreads);
    if (!scratch || !wn) return -1;
    cplx *vc = (cplx *) vals;

    double t0 = now_s();
    /* phase 1 (loop.cpp:61-68): 1-D r2c in y on every row */
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long i = 0; i < (long) dim_c_x; ++i) {
#ifdef _OPENMP
        cplx *s = scratch + smax * (size_t) omp_get_thread_num();
#else
        cplx *s = scratch;
#endif
        r2c_row(&p_half, wn, vals + (size_t) i * n_col, (int) dim_r_y, s);
    }
    double t1 = now_s();
    /* phase 2 (loop.cpp:72-80, 18-25): trans(ky, x) = vals(x, ky), one task per ky */
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long ky = 0; ky < (long) dim_c_y; ++ky)
        for (size_t x = 0; x < dim_c_x; ++x) trans[(size_t) ky * dim_c_x + x] = vc[x * dim_c_y + (size_t) ky];
    double t2 = now_s();
    /* phase 3 (loop.cpp:83-91): forward c2c in x on every transposed row */
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long ky = 0; ky < (long) dim_c_y; ++ky) {
#ifdef _OPENMP
        cplx *s = scratch + smax * (size_t) omp_get_thread_num();
#else
        cplx *s = scratch;
#endif
        fft1d_exec(&p_x, trans + (size_t) ky * dim_c_x, s);
    }
    double t3 = now_s();
    /* phase 4 (loop.cpp:94-102, 46-53): vals(kx, ky) = trans(ky, kx) */
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long ky = 0; ky < (long) dim_c_y; ++ky)
        for (size_t x = 0; x < dim_c_x; ++x) vc[x * dim_c_y + (size_t) ky] = trans[(size_t) ky * dim_c_x + x];
    double t4 = now_s();
    if (timings) {
        timings[0] = t4 - t0;
        timings[1] = t1 - t0;
        timings[2] = t2 - t1;
        timings[3] = t3 - t2;
        timings[4] = t4 - t3;
    }
    free(scratch);
    free(wn);
    plan1d_free(&p_half);
    plan1d_free(&p_x);
    free(trans);
    return 0;
}

/* batched 1-D forward c2c (rows of length n, contiguous), used by the 1-D kernel parity tests */
int hpxfft_oracle_c2c_rows(double *data, size_t n_rows, size_t n, int nthreads)
{
    if (!data || n == 0) return -1;
    plan1d p;
    plan1d_init(&p, (int) n);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    cplx *scratch = (cplx *) malloc(sizeof(cplx) * n * (size_t) nthreads);
    if (!scratch) return -1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long i = 0; i < (long) n_rows; ++i) {
#ifdef _OPENMP
        cplx *s = scratch + n * (size_t) omp_get_thread_num();
#else
        cplx *s = scratch;
#endif
        fft1d_exec(&p, (cplx *) data + (size_t) i * n, s);
    }
    free(scratch);
    plan1d_free(&p);
    return 0;
}
